"""Host-side cost of the rollout loop (OnPolicyRunner.learn's collection phase): cProfile over rollout steps at 4096 envs.
    python profiles/rollout_host_profile.py
"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from legged_games_gym_b200.envs import task_registry  # noqa: E402
from legged_games_gym_b200.utils import get_args  # noqa: E402

dev = "cuda:0"
task = "anymal_c_rough"
a = get_args(["--task", task, "--num_envs", "4096", "--headless", "--sim_device", dev, "--rl_device", dev])
env, _ = task_registry.make_env(name=task, args=a)
runner, _ = task_registry.make_alg_runner(env=env, name=task, args=a, log_root=None)
runner.learn(num_learning_iterations=2, init_at_random_ep_len=True)
alg = runner.alg
obs = env.get_observations()


def rollout(steps):
    global obs
    with torch.inference_mode():
        for s in range(steps):
            alg.storage.step = s % 24
            alg.actor_critic.set_rng(1, s, 0)
            actions = alg.act(obs, obs)
            obs, _, rew, dones, infos = env.step(actions)
            alg.process_env_step(rew, dones, infos)


rollout(48)
torch.cuda.synchronize()
t0 = time.perf_counter()
rollout(240)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host time per rollout step {(t1 - t0) / 240 * 1e6:.1f} us (GPU drained {(t2 - t1) * 1e6:.0f} us after the loop)")
pr = cProfile.Profile()
pr.enable()
rollout(240)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
