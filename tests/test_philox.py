"""Known-answer vectors for the counter-based RNG restatement (Random123 kat_vectors, philox4x32-10)."""
import numpy as np

from oracle import philox


def test_philox_kat():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(np.array(ctr, np.uint32), np.array(key, np.uint32))
        assert tuple(int(x) for x in got) == want


def test_uniform_range_and_layout():
    u = philox.uniforms(7, 3, np.arange(5), philox.STREAM_CMD, 3)
    assert u.shape == (5, 3) and u.dtype == np.float32 and (u >= 0).all() and (u < 1).all()
    # word j of env e is philox((e, j//4, stream, step))[j%4]
    w = philox.raw_u32(7, 3, np.array([2]), philox.STREAM_RESET_DOF, 12)
    blk1 = philox.philox4x32_10(np.array([2, 1, philox.STREAM_RESET_DOF, 3], np.uint32), np.array([7, 0], np.uint32))
    assert (w[0, 4:8] == blk1).all()


def test_obs_stream_mapping():
    o = philox.obs_uniforms(9, 4, np.array([1, 6]), 235)
    assert o.shape == (2, 235)
    # column 200: block 32*(200//128) + 200%32 = 40, word (200//32)%4 = 2
    blk = philox.philox4x32_10(np.array([6, 40, philox.STREAM_OBS, 4], np.uint32), np.array([9, 0], np.uint32))
    assert o[1, 200] == np.float32(blk[2] >> 8) * np.float32(2.0 ** -24)
