import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from legged_games_gym_b200.envs import task_registry
from legged_games_gym_b200.utils import get_args
task = sys.argv[1] if len(sys.argv) > 1 else "anymal_c_rough"
n = sys.argv[2] if len(sys.argv) > 2 else "4096"
args = get_args(["--task", task, "--num_envs", n, "--headless", "--max_iterations", "6"])
env, env_cfg = task_registry.make_env(name=args.task, args=args)
runner, train_cfg = task_registry.make_alg_runner(env=env, name=args.task, args=args, log_root=None)
for it in range(6):
    torch.cuda.synchronize(); t0 = time.time()
    runner.learn(num_learning_iterations=1, init_at_random_ep_len=(it == 0))
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"iter {it}: {dt*1e3:.1f} ms  collection {runner.collection_time*1e3:.1f} ms  learning {runner.learn_time*1e3:.1f} ms  "
          f"-> {runner.num_steps_per_env * env.num_envs / dt / 1e6:.2f} M env-steps/s", flush=True)
