"""World-size-2 gloo tests (CPU) of the host-side multi-process logic: env sharding keeps RNG streams disjoint and the
PPO gradient bucket all-reduce + KL all-reduce keep the ranks' parameters identical."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ppo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                   # identical initial weights on every rank
    from legged_games_gym_b200.rsl_rl.modules import ActorCritic
    from legged_games_gym_b200.rsl_rl.algorithms import PPO
    ac = ActorCritic(20, 20, 4, [32, 16, 8], [32, 16, 8])
    alg = PPO(ac, num_learning_epochs=2, num_mini_batches=2, schedule="adaptive", device="cpu")
    alg.init_storage(16, 6, [20], [None], [4])
    g = torch.Generator().manual_seed(100 + rank)          # different data per rank (its own env shard)
    st = alg.storage
    st.observations.copy_(torch.randn(st.observations.shape, generator=g))
    st.actions.copy_(torch.randn(st.actions.shape, generator=g))
    st.values.copy_(torch.randn(st.values.shape, generator=g))
    st.returns.copy_(torch.randn(st.returns.shape, generator=g))
    st.advantages.copy_(torch.randn(st.advantages.shape, generator=g))
    st.actions_log_prob.copy_(torch.randn(st.actions_log_prob.shape, generator=g))
    st.mu.copy_(torch.randn(st.mu.shape, generator=g))
    st.sigma.copy_(torch.rand(st.sigma.shape, generator=g) + 0.5)
    st.step = 6
    torch.manual_seed(7)                                   # same mini-batch permutation on both ranks
    alg.update()
    flat = torch.cat([p.detach().reshape(-1) for p in ac.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        out.put((alg.allreduce_calls, float((gathered[0] - gathered[1]).abs().max()), alg.learning_rate))
    dist.destroy_process_group()


def test_ppo_update_allreduce_keeps_ranks_in_sync():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ppo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    calls, max_diff, lr = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert calls == 4                      # 2 epochs x 2 mini-batches, one bucket all-reduce each
    assert max_diff == 0.0                 # identical parameters after the update although the data differed


def test_env_shards_have_disjoint_rng_streams():
    from oracle import philox
    n = 64
    a = philox.uniforms(1, 5, np.arange(0, n), philox.STREAM_RESET_DOF, 12)         # rank 0: env_id_offset 0
    b = philox.uniforms(1, 5, np.arange(n, 2 * n), philox.STREAM_RESET_DOF, 12)     # rank 1: env_id_offset n
    whole = philox.uniforms(1, 5, np.arange(0, 2 * n), philox.STREAM_RESET_DOF, 12)
    assert np.array_equal(np.concatenate([a, b]), whole)        # sharded job == single-process job, stream-wise
    assert not np.array_equal(a, b)
