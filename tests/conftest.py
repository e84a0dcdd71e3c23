import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist for every test session (CPU tests only load it and check its symbols)."""
    from legged_games_gym_b200.csrc.build import build
    build()


@pytest.fixture(autouse=True)
def _strict_fp32_matmul():
    """PPO switches cuBLAS to TF32 for its update (a process-wide torch flag); parity tests that run after a training test
    in the same process must see strict fp32 again, whatever the test order."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


@pytest.fixture(autouse=True)
def _pristine_task_registry():
    """task_registry.make_env applies command-line overrides to the REGISTERED cfg objects in place, as the reference's
    does (utils/task_registry.py:96-99): a training test would otherwise leave its num_envs / terrain overrides behind for
    whatever test reads the registry next."""
    import copy
    from legged_games_gym_b200.envs import task_registry
    env, train = copy.deepcopy(task_registry.env_cfgs), copy.deepcopy(task_registry.train_cfgs)
    yield
    task_registry.env_cfgs.clear()
    task_registry.env_cfgs.update(env)
    task_registry.train_cfgs.clear()
    task_registry.train_cfgs.update(train)
