"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the counter-based RNG used by the kernels.

The reference draws from torch's global generator in a data-dependent order
(legged_gym/envs/base/legged_robot.py:353-366, 405, 425, 431, 442, 465, 230), which no
kernel can mirror.  The product therefore defines its own stream: Philox4x32-10
(Salmon et al., SC'11; the published algorithm, same constants as Random123/cuRAND)
keyed by ``seed`` with counter ``(env, block, stream, step)``.  This file is the
bit-exact integer restatement the tests pin ``lgk_rng_uniforms`` against; the uniforms it
produces are what the oracle consumes through its explicit ``U`` tables (the "RNG tap").

Known-answer vectors (Random123 kat_vectors, philox4x32-10) are checked in
tests/test_philox.py.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)

# stream ids (must match legged_games_gym_b200/csrc/lgk_rng.cuh)
STREAM_CMD = 0         # callback command resample: x, y, heading|yaw
STREAM_PUSH = 1        # push velocities: x, y
STREAM_RESET_DOF = 2   # 12 dof position factors
STREAM_RESET_ROOT = 3  # xy offset (2) then base velocity (6)
STREAM_RESET_CMD = 4   # command resample inside reset_idx: x, y, heading|yaw
STREAM_TERRAIN = 5     # raw uint32 for the terrain-level wrap randint
STREAM_OBS = 6         # observation noise, one uniform per obs column
STREAM_ACT = 7         # policy action noise (Box-Muller pairs)
STREAM_PREDATOR = 8    # low_level_game predator spawn: offset xyz (3) then the sign draw
# the high-level games reset root / dof state of the SAME low-level env a second time in one step (HLG:326-347,
# DHLG:270-296): their draws come from streams of their own
STREAM_GAME_ROOT = 9       # xy offset (2) then base velocity (6)
STREAM_GAME_PREDATOR = 10  # predator spawn: offset xyz (3) then the sign draw
STREAM_GAME_DOF = 11       # 12 dof position factors (DecHighLevelGame only)


def philox4x32_10(counter, key):
    """counter: uint32 [..., 4]; key: uint32 [..., 2] (broadcastable) -> uint32 [..., 4]."""
    c = np.array(counter, dtype=np.uint32, copy=True)
    k = np.array(np.broadcast_to(np.asarray(key, dtype=np.uint32), c.shape[:-1] + (2,)), copy=True)
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    k0, k1 = k[..., 0].copy(), k[..., 1].copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = PHILOX_M0 * c0.astype(np.uint64)
            p1 = PHILOX_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = k0 + PHILOX_W0
            k1 = k1 + PHILOX_W1
    return np.stack([c0, c1, c2, c3], axis=-1)


def raw_u32(seed, step, env_ids, stream, count):
    """uint32 table [len(env_ids), count]: word j of env e = philox((e, j//4, stream, step))[j%4]."""
    env_ids = np.asarray(env_ids, dtype=np.uint32)
    nblk = (count + 3) // 4
    ctr = np.zeros((env_ids.shape[0], nblk, 4), dtype=np.uint32)
    ctr[..., 0] = env_ids[:, None]
    ctr[..., 1] = np.arange(nblk, dtype=np.uint32)[None, :]
    ctr[..., 2] = np.uint32(stream)
    ctr[..., 3] = np.uint32(step & 0xFFFFFFFF)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    out = philox4x32_10(ctr, key).reshape(env_ids.shape[0], nblk * 4)
    return out[:, :count]


def uniforms(seed, step, env_ids, stream, count):
    """float32 uniforms in [0,1): (word >> 8) * 2**-24 -- exact in fp32."""
    w = raw_u32(seed, step, env_ids, stream, count)
    return (w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def obs_uniforms(seed, step, env_ids, num_obs):
    """OBS stream: column j uses word (j//32)%4 of block 32*(j//128) + j%32, so that one warp lane owns one
    Philox block per 128 columns and writes coalesced row segments (csrc/lgk_step_device.cuh obs_block_of)."""
    env_ids = np.asarray(env_ids, dtype=np.uint32)
    nsuper = (num_obs + 127) // 128
    nblk = 32 * nsuper
    ctr = np.zeros((env_ids.shape[0], nblk, 4), dtype=np.uint32)
    ctr[..., 0] = env_ids[:, None]
    ctr[..., 1] = np.arange(nblk, dtype=np.uint32)[None, :]
    ctr[..., 2] = np.uint32(STREAM_OBS)
    ctr[..., 3] = np.uint32(step & 0xFFFFFFFF)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    words = philox4x32_10(ctr, key)                      # [n, nblk, 4]
    j = np.arange(num_obs)
    w = words[:, 32 * (j // 128) + (j % 32), (j // 32) % 4]
    return (w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def normals(seed, step, env_ids, count):
    """Box-Muller on the ACT stream: pair (2i, 2i+1) of uniforms -> normals (2i, 2i+1).
    u1 is mapped to (0,1] so the log is finite.  fp32 arithmetic like the kernel
    (tolerance-level agreement only: logf/cosf differ in the last ulps)."""
    npair = (count + 1) // 2
    u = uniforms(seed, step, env_ids, STREAM_ACT, 2 * npair)
    u1 = np.float32(1.0) - u[:, 0::2]
    u2 = u[:, 1::2]
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    th = (np.float32(2.0 * np.pi) * u2).astype(np.float32)
    z = np.empty((u.shape[0], 2 * npair), dtype=np.float32)
    z[:, 0::2] = r * np.cos(th)
    z[:, 1::2] = r * np.sin(th)
    return z[:, :count]
