"""TaskRegistry (mirror of reference legged_gym/utils/task_registry.py:44-224: register / get_task_class / get_cfgs /
make_env / make_alg_runner / make_dec_alg_runner, same signatures)."""
import os
from datetime import datetime

from .. import LEGGED_GYM_ROOT_DIR
from .helpers import get_args, update_cfg_from_args, class_to_dict, get_load_path, get_dec_load_path, set_seed, parse_sim_params


class TaskRegistry:
    def __init__(self):
        self.task_classes = {}
        self.env_cfgs = {}
        self.train_cfgs = {}

    def register(self, name, task_class, env_cfg, train_cfg):
        self.task_classes[name] = task_class
        self.env_cfgs[name] = env_cfg
        self.train_cfgs[name] = train_cfg

    def get_task_class(self, name):
        return self.task_classes[name]

    def get_cfgs(self, name):
        train_cfg = self.train_cfgs[name]
        env_cfg = self.env_cfgs[name]
        env_cfg.seed = train_cfg.seed
        return env_cfg, train_cfg

    def make_env(self, name, args=None, env_cfg=None, **env_kwargs):
        if args is None:
            args = get_args()
        if name not in self.task_classes:
            raise ValueError(f"Task with name: {name} was not registered")
        task_class = self.get_task_class(name)
        if env_cfg is None:
            env_cfg, _ = self.get_cfgs(name)
        env_cfg, _ = update_cfg_from_args(env_cfg, None, args)
        set_seed(getattr(env_cfg, "seed", 1))
        sim_params = parse_sim_params(args, {"sim": class_to_dict(env_cfg.sim)})
        env = task_class(cfg=env_cfg, sim_params=sim_params, physics_engine=args.physics_engine,
                         sim_device=args.sim_device, headless=args.headless, **env_kwargs)
        return env, env_cfg

    def make_alg_runner(self, env, name=None, args=None, train_cfg=None, log_root="default"):
        return self._make_runner(False, env, name, args, train_cfg, log_root)

    def make_dec_alg_runner(self, env, name=None, args=None, train_cfg=None, log_root="default"):
        """task_registry.py:172-221: the two-agent runner of dec_high_level_game"""
        return self._make_runner(True, env, name, args, train_cfg, log_root)

    def _make_runner(self, dec, env, name, args, train_cfg, log_root):
        from ..rsl_rl.runners import OnPolicyRunner, DecGamePolicyRunner
        if args is None:
            args = get_args()
        if train_cfg is None:
            if name is None:
                raise ValueError("Either 'name' or 'train_cfg' must be not None")
            _, train_cfg = self.get_cfgs(name)
        elif name is not None:
            print(f"'train_cfg' provided -> Ignoring 'name={name}'")
        _, train_cfg = update_cfg_from_args(None, train_cfg, args)
        if log_root == "default":
            log_root = os.path.join(LEGGED_GYM_ROOT_DIR, "logs", train_cfg.runner.experiment_name)
            log_dir = os.path.join(log_root, datetime.now().strftime("%b%d_%H-%M-%S") + "_" + train_cfg.runner.run_name)
        elif log_root is None:
            log_dir = None
        else:
            log_dir = os.path.join(log_root, datetime.now().strftime("%b%d_%H-%M-%S") + "_" + train_cfg.runner.run_name)
        # multi-GPU: every rank resolves the resume path from the real log root, rank 0 alone writes logs / checkpoints
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
            log_dir = None
        if dec:
            runner = DecGamePolicyRunner(env, class_to_dict(train_cfg), log_dir, device=args.rl_device)
            if train_cfg.runner.resume:
                for agent_id, label in ((0, "PREDATOR"), (1, "PREY")):
                    path = get_dec_load_path(log_root, agent_id=agent_id, load_run=train_cfg.runner.load_run,
                                             checkpoint=train_cfg.runner.checkpoint)
                    print(f"Loading {label} model from: {path}")
                    runner.load(agent_id=agent_id, path=path)
            return runner, train_cfg
        runner = OnPolicyRunner(env, class_to_dict(train_cfg), log_dir, device=args.rl_device)
        if train_cfg.runner.resume:
            resume_path = get_load_path(log_root, load_run=train_cfg.runner.load_run, checkpoint=train_cfg.runner.checkpoint)
            print(f"Loading model from: {resume_path}")
            runner.load(resume_path)
        return runner, train_cfg


task_registry = TaskRegistry()
