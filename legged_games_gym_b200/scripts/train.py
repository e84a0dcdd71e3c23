"""python -m legged_games_gym_b200.scripts.train --task anymal_c_rough --num_envs 4096 --headless --max_iterations 100
(same entry point and flags as the reference's legged_gym/scripts/train.py:40-47).

Multi-GPU (one process per GPU, environments sharded, no collective inside env.step):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        -m legged_games_gym_b200.scripts.train --task anymal_c_rough --num_envs 65536 --headless
--num_envs is PER GPU; rank r owns the global env ids [r*N, (r+1)*N) (disjoint Philox streams), PPO gradients and the KL
estimate are all-reduced over NCCL (rsl_rl/algorithms/ppo.py), episode statistics are all-reduced for logging."""
import os

import torch

from legged_games_gym_b200.envs import *  # noqa: F401,F403  (registers the tasks)
from legged_games_gym_b200.utils import get_args, task_registry


def setup_distributed(args):
    """Reads torchrun's environment; returns (rank, world).  Single process: (0, 1) and nothing is initialised."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1
    import torch.distributed as dist
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    args.sim_device = args.rl_device = f"cuda:{local}"
    args.sim_device_id = args.compute_device_id = local
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    return rank, world


def train(args, log_root="default"):
    rank, world = setup_distributed(args)
    env_cfg, _ = task_registry.get_cfgs(name=args.task)
    env, env_cfg = task_registry.make_env(name=args.task, args=args, env_cfg=env_cfg)
    if world > 1:
        env.set_env_id_offset(rank * env.num_envs)
    ppo_runner, train_cfg = task_registry.make_alg_runner(env=env, name=args.task, args=args,
                                                          log_root=log_root)
    ppo_runner.learn(num_learning_iterations=train_cfg.runner.max_iterations, init_at_random_ep_len=True)
    return ppo_runner


if __name__ == "__main__":
    import sys
    runner = train(get_args(sys.argv[1:]))
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        runner.alg.release_graph()          # the captured update graph holds NCCL work
        del runner
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
