"""Host-side helpers (mirror of the parts of reference legged_gym/utils/helpers.py the hot path
and the registry need: class_to_dict :41-56, update_class_from_dict :58-65, set_seed :67-77,
parse_sim_params :79-101, get_load_path :103-125, update_cfg_from_args :159-182, get_args :184-210).
``gymutil.parse_arguments`` (Isaac Gym) is replaced by argparse with the same flag names."""
import argparse
import os
import random
import types

import numpy as np
import torch


def class_to_dict(obj) -> dict:
    """dir()-ordered (i.e. ALPHABETICAL) dict of an object's public attributes, recursively.
    The alphabetical order is load-bearing: it fixes the reward summation order (LR:583-607)."""
    if not hasattr(obj, "__dict__"):
        return obj
    out = {}
    for name in dir(obj):
        if name.startswith("_"):
            continue
        attr = getattr(obj, name)
        if isinstance(attr, list):
            out[name] = [class_to_dict(x) for x in attr]
        else:
            out[name] = class_to_dict(attr)
    return out


def update_class_from_dict(obj, d):
    for name, v in d.items():
        cur = getattr(obj, name, None)
        if isinstance(cur, type):
            update_class_from_dict(cur, v)
        else:
            setattr(obj, name, v)


def set_seed(seed):
    if seed == -1:
        seed = np.random.randint(0, 10000)
    print("Setting seed: {}".format(seed))
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    return seed


class SimParams(types.SimpleNamespace):
    """Stand-in for gymapi.SimParams: only ``dt`` (and the flags below) are read by the hot path."""


def parse_sim_params(args, cfg):
    sp = SimParams(dt=1. / 60., substeps=2, use_gpu_pipeline=getattr(args, "use_gpu_pipeline", True),
                   physx=SimParams(use_gpu=getattr(args, "use_gpu", True),
                                   num_subscenes=getattr(args, "subscenes", 0), num_threads=0))
    if "sim" in cfg:
        for k, v in cfg["sim"].items():
            if isinstance(v, dict):
                sub = getattr(sp, k, SimParams())
                for kk, vv in v.items():
                    setattr(sub, kk, vv)
                setattr(sp, k, sub)
            else:
                setattr(sp, k, v)
    if getattr(args, "num_threads", 0) > 0:
        sp.physx.num_threads = args.num_threads
    return sp


def get_load_path(root, load_run=-1, checkpoint=-1):
    try:
        runs = sorted(os.listdir(root))
        if "exported" in runs:
            runs.remove("exported")
        last_run = os.path.join(root, runs[-1])
    except Exception:
        raise ValueError("No runs in this directory: " + root)
    load_run = last_run if load_run == -1 else os.path.join(root, load_run)
    if checkpoint == -1:
        models = [f for f in os.listdir(load_run) if "model" in f]
        models.sort(key=lambda m: "{0:0>15}".format(m))
        model = models[-1]
    else:
        model = "model_{}.pt".format(checkpoint)
    return os.path.join(load_run, model)


def get_dec_load_path(root, agent_id, load_run=-1, checkpoint=-1):
    """helpers.py:127-156: checkpoints of the decentralised game, ``pred_model_<n>.pt`` (agent 0) / ``prey_model_<n>.pt``"""
    try:
        runs = sorted(os.listdir(root))
        if "exported" in runs:
            runs.remove("exported")
        last_run = os.path.join(root, runs[-1])
    except Exception:
        raise ValueError("No runs in this directory: " + root)
    load_run = last_run if load_run == -1 else os.path.join(root, load_run)
    agent_name = "pred" if agent_id == 0 else "prey"
    if checkpoint == -1:
        models = [f for f in os.listdir(load_run) if agent_name + "_model_" in f]
        models.sort(key=lambda m: "{0:0>15}".format(m))
        model = models[-1]
    else:
        model = (agent_name + "_model_{}.pt").format(checkpoint)
    return os.path.join(load_run, model)


def update_cfg_from_args(env_cfg, cfg_train, args):
    if env_cfg is not None and getattr(args, "num_envs", None) is not None:
        env_cfg.env.num_envs = args.num_envs
    if cfg_train is not None:
        if getattr(args, "seed", None) is not None:
            cfg_train.seed = args.seed
        r = cfg_train.runner
        if getattr(args, "max_iterations", None) is not None:
            r.max_iterations = args.max_iterations
        if getattr(args, "resume", False):
            r.resume = args.resume
        for k in ("experiment_name", "run_name", "load_run", "checkpoint"):
            if getattr(args, k, None) is not None:
                setattr(r, k, getattr(args, k))
    return env_cfg, cfg_train


def export_policy_as_jit(actor_critic, path):
    """reference legged_gym/utils/helpers.py:212-222: the actor MLP as TorchScript, <path>/policy_1.pt (run from C++)."""
    import copy
    import torch
    os.makedirs(path, exist_ok=True)
    path = os.path.join(path, "policy_1.pt")
    model = copy.deepcopy(actor_critic.actor).to("cpu")
    torch.jit.script(model).save(path)
    return path


def get_args(argv=None):
    p = argparse.ArgumentParser(description="RL Policy")
    p.add_argument("--task", type=str, default="anymal_c_flat")
    p.add_argument("--resume", action="store_true", default=False)
    p.add_argument("--experiment_name", type=str)
    p.add_argument("--run_name", type=str)
    p.add_argument("--load_run", type=str)
    p.add_argument("--checkpoint", type=int)
    p.add_argument("--headless", action="store_true", default=False)
    p.add_argument("--horovod", action="store_true", default=False)
    p.add_argument("--rl_device", type=str, default="cuda:0")
    p.add_argument("--num_envs", type=int)
    p.add_argument("--seed", type=int)
    p.add_argument("--max_iterations", type=int)
    # the subset of gymutil.parse_arguments' own flags that the reference reads
    p.add_argument("--sim_device", type=str, default="cuda:0")
    p.add_argument("--pipeline", type=str, default="gpu")
    p.add_argument("--physics_engine", type=str, default="physx")
    p.add_argument("--num_threads", type=int, default=0)
    p.add_argument("--subscenes", type=int, default=0)
    args = p.parse_args([] if argv is None else argv)
    args.use_gpu_pipeline = args.pipeline.lower() == "gpu"
    args.use_gpu = args.sim_device.startswith("cuda")
    args.sim_device_type = args.sim_device.split(":")[0]
    args.compute_device_id = int(args.sim_device.split(":")[1]) if ":" in args.sim_device else 0
    args.sim_device_id = args.compute_device_id
    return args
