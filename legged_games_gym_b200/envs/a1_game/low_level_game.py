"""LowLevelGame -- the reference's predator/prey low-level task (legged_gym/envs/a1_game/low_level_game.py:52-1045).

In the reference this class is a stand-alone copy of LeggedRobot whose differences are (a) every env holds TWO actors,
the A1 "prey" and a sphere "predator" (LLG:710-812), so root_states has 2N rows addressed through ``prey_indices`` /
``predator_indices`` (LLG:123-125, 139, 228, 410-432, 457, 470, 536-585, 856), (b) the contact view has 18 bodies per env
(17 robot links + the sphere, LLG:561), (c) a reset re-spawns the predator at ``prey_pos - sign * U(1,10)^3`` with z = 0.3
(LLG:419-432).  Here it is LeggedRobot with those three facts handed to the kernels: ``actors_per_env = 2``,
``root_actor_offset = 0``, ``predator_spawn = 1`` (lgk.h).  Nothing else of the step changes."""
import torch

from ..base.legged_robot import LeggedRobot


class LowLevelGame(LeggedRobot):
    def __init__(self, cfg, sim_params, physics_engine, sim_device, headless, **kw):
        print("[LowLevelGame] initializing... ")
        self.aggregate_mode = 1
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless, **kw)

    def _actors_per_env(self):
        return 2

    def _extra_bodies_per_env(self):
        return 1                                   # the predator sphere closes the env's body list

    def _create_envs(self):
        super()._create_envs()
        n = self.num_envs
        # actor creation order per env is robot, then sphere (LLG:799-812): DOMAIN_SIM indices 2i and 2i+1
        self.prey_indices = torch.arange(n, dtype=torch.long, device=self.device) * 2
        self.predator_indices = self.prey_indices + 1

    def _root_rows(self):
        return self.prey_indices

    def _init_buffers(self):
        super()._init_buffers()
        n = self.num_envs
        # the reference re-copies the prey quaternions at the top of every post_physics_step (LLG:123) and keeps that copy
        # past reset_idx (the high-level games read it then, HLG:432): the step kernel writes it (LgkStepParams.base_quat)
        self.base_quat = self.root_states.view(n, 2, 13)[:, 0, 3:7].clone()
        # LLG:538-558 -- initial predator position (init-time draw from torch's global generator, like the reference)
        init_prey_pos = self.root_states[self.prey_indices, :3].detach().clone()
        rand_offset = torch.zeros_like(init_prey_pos).uniform_(1.0, 10.0)
        rand_sign = torch.rand(n, dtype=torch.float, device=self.device)
        rand_sign = torch.where(rand_sign < 0.5, -torch.ones_like(rand_sign), torch.ones_like(rand_sign)).unsqueeze(1)
        self.init_predator_pos = init_prey_pos - rand_sign * rand_offset
        self.init_predator_pos[:, 2] = 0.3

    def _configure_native(self, p):
        p.predator_spawn, p.predator_actor_offset = 1, 1
        p.base_quat = self.base_quat.data_ptr()

    def _push_resets_to_sim(self):
        # LLG:441-451: dof states and root states of the prey actors, then root states of the predator actors
        gym = self.gym
        gym.set_dof_state_tensor_indexed(self.dof_state, self.reset_env_ids, self.reset_count)
        gym.set_actor_root_state_tensor_indexed(self.root_states, self.reset_env_ids, self.reset_count, actor_stride=2, actor_offset=0)
        gym.set_actor_root_state_tensor_indexed(self.root_states, self.reset_env_ids, self.reset_count, actor_stride=2, actor_offset=1)
