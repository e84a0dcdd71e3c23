"""Logger with the reference's interface (legged_gym/utils/logger.py:36-128): log_state(s), log_rewards, print_rewards,
plot_states.  matplotlib is not a dependency here: plot_states prints a per-signal summary instead of opening a window."""
from collections import defaultdict

import numpy as np


class Logger:
    def __init__(self, dt):
        self.state_log = defaultdict(list)
        self.rew_log = defaultdict(list)
        self.dt = dt
        self.num_episodes = 0
        self.plot_process = None

    def log_state(self, key, value):
        self.state_log[key].append(value)

    def log_states(self, dict):
        for key, value in dict.items():
            self.log_state(key, value)

    def log_rewards(self, dict, num_episodes):
        for key, value in dict.items():
            if "rew" in key:
                self.rew_log[key].append(value.item() * num_episodes)
        self.num_episodes += num_episodes

    def reset(self):
        self.state_log.clear()
        self.rew_log.clear()

    def plot_states(self):
        for key, values in self.state_log.items():
            a = np.asarray(values, dtype=np.float64)
            print(f" - {key}: n={a.shape[0]} mean={a.mean():.4f} min={a.min():.4f} max={a.max():.4f}")

    def print_rewards(self):
        print("Average rewards per second:")
        for key, values in self.rew_log.items():
            mean = np.sum(np.array(values)) / max(self.num_episodes, 1)
            print(f" - {key}: {mean}")
        print(f"Total number of episodes: {self.num_episodes}")
