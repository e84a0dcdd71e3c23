"""BaseTask: device selection, rsl_rl VecEnv buffers, reset()/get_observations()
(mirror of reference legged_gym/envs/base/base_task.py:38-118; the viewer/keyboard half, :84-102 and
:123-147, is graphics and out of scope -- render() is a no-op)."""
import torch

from ...sim.state_feeder import SimBackend


class BaseTask:
    def __init__(self, cfg, sim_params, physics_engine, sim_device, headless, sim_backend=None):
        self.sim_params = sim_params
        self.physics_engine = physics_engine
        self.sim_device = sim_device
        self.headless = headless
        dev_type = str(sim_device).split(":")[0]
        self.sim_device_id = int(str(sim_device).split(":")[1]) if ":" in str(sim_device) else 0
        # reference: env device is the GPU only with the GPU pipeline (base_task.py:49-53).  The kernels ARE the
        # GPU pipeline, so anything else is an error rather than a silent CPU path.
        if dev_type != "cuda" or not getattr(sim_params, "use_gpu_pipeline", True):
            raise RuntimeError("legged_games_gym_b200 runs the env step on a CUDA device only (no CPU fallback); "
                               f"got sim_device={sim_device!r}")
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA device required: the env step is implemented as sm_100a kernels only")
        self.device = str(sim_device) if ":" in str(sim_device) else "cuda:0"
        self.graphics_device_id = -1
        self.num_envs = cfg.env.num_envs
        self.num_obs = cfg.env.num_observations
        self.num_privileged_obs = cfg.env.num_privileged_obs
        self.num_actions = cfg.env.num_actions
        dev = self.device
        self.obs_buf = torch.zeros(self.num_envs, self.num_obs, device=dev, dtype=torch.float)
        self.rew_buf = torch.zeros(self.num_envs, device=dev, dtype=torch.float)
        self.reset_buf = torch.ones(self.num_envs, device=dev, dtype=torch.long)   # becomes bool at the first step
        self.episode_length_buf = torch.zeros(self.num_envs, device=dev, dtype=torch.long)
        self.time_out_buf = torch.zeros(self.num_envs, device=dev, dtype=torch.bool)
        if self.num_privileged_obs is not None:
            self.privileged_obs_buf = torch.zeros(self.num_envs, self.num_privileged_obs, device=dev, dtype=torch.float)
        else:
            self.privileged_obs_buf = None
        self.extras = {}
        self.gym = sim_backend            # same attribute name as the reference's Isaac Gym handle
        self.create_sim()
        if not isinstance(self.gym, SimBackend):
            raise TypeError("sim backend must implement legged_games_gym_b200.sim.state_feeder.SimBackend")
        self.enable_viewer_sync = True
        self.viewer = None

    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def reset_idx(self, env_ids):
        raise NotImplementedError

    def reset(self):
        """Reset all robots, then one zero-action step (base_task.py:114-118)."""
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        obs, privileged_obs, _, _, _ = self.step(
            torch.zeros(self.num_envs, self.num_actions, device=self.device, requires_grad=False))
        return obs, privileged_obs

    def step(self, actions):
        raise NotImplementedError

    def render(self, sync_frame_time=True):
        return None
