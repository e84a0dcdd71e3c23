"""2-GPU NCCL check of the sharded training flow (run under torchrun on a multi-GPU box; not collected by pytest):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_train_check.py
Every rank trains 2 PPO iterations on its own env shard; the gradient all-reduce must keep the policies bit-identical,
the shards must have consumed different RNG streams (different observations), and the episode statistics are global."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from legged_games_gym_b200.scripts.train import train  # noqa: E402
from legged_games_gym_b200.utils import get_args  # noqa: E402

args = get_args(["--task", "anymal_c_flat", "--num_envs", "512", "--headless", "--max_iterations", "2", "--seed", "5"])
runner = train(args, log_root=None)
rank, world = dist.get_rank(), dist.get_world_size()
flat = torch.cat([p.detach().reshape(-1) for p in runner.alg.actor_critic.parameters()])
gathered = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
obs = runner.env.obs_buf.flatten()[:4096].contiguous()
obs_all = [torch.zeros_like(obs) for _ in range(world)]
dist.all_gather(obs_all, obs)
if rank == 0:
    diffs = [float((gathered[0] - g).abs().max()) for g in gathered[1:]]
    assert all(d == 0.0 for d in diffs), f"policies diverged across ranks: {diffs}"
    assert not torch.equal(obs_all[0], obs_all[1]), "shards must see different environments"
    assert runner.alg.allreduce_calls == 2 * runner.alg.num_learning_epochs * runner.alg.num_mini_batches
    assert runner.env.env_id_offset == 0 and runner.tot_timesteps == 2 * world * 512 * runner.num_steps_per_env
    print(f"multi-GPU train check ok: world {world}, {runner.alg.allreduce_calls} gradient all-reduces, "
          f"max parameter difference {max(diffs)}, global episode stats {getattr(runner, 'global_episode_stats', None)}")
# a captured update graph holds NCCL work: release it before the communicator goes away (destroy_process_group waits
# on it otherwise)
runner.alg.release_graph()
del runner
import gc
gc.collect()
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
