"""Which kernels make up PPO.update (5 epochs x 4 mini-batches of 24 576 samples, captured graph)?  torch.profiler over one
runner.learn iteration at 4096 envs; prints the top kernels by total device time.
    python profiles/update_profile.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

from legged_games_gym_b200.envs import task_registry  # noqa: E402
from legged_games_gym_b200.utils import get_args  # noqa: E402

dev = "cuda:0"
task = "anymal_c_rough"
a = get_args(["--task", task, "--num_envs", "4096", "--headless", "--sim_device", dev, "--rl_device", dev])
env, _ = task_registry.make_env(name=task, args=a)
runner, _ = task_registry.make_alg_runner(env=env, name=task, args=a, log_root=None)
for it in range(3):
    runner.learn(num_learning_iterations=1, init_at_random_ep_len=(it == 0))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    runner.learn(num_learning_iterations=1, init_at_random_ep_len=False)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"device time of one iteration: {tot / 1e3:.2f} ms over {sum(e.count for e in rows)} kernel launches; "
      f"collection {runner.collection_time * 1e3:.2f} ms, learning {runner.learn_time * 1e3:.2f} ms (wall)")
for e in rows[:45]:
    print(f"{e.device_time_total / 1e3:8.3f} ms  {e.count:5d} x {e.device_time_total / e.count:8.1f} us  {e.key[:110]}")
