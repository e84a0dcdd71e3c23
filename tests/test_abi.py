"""The C-ABI library loads on a CPU-only box and exports every symbol include/lgk.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lgk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lgk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from legged_games_gym_b200 import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"liblgk.so does not export {s}"
    assert sorted(_native.EXPORTS) == syms


def test_struct_mirrors_match():
    from legged_games_gym_b200 import _native as nat
    for which, cls in enumerate((nat.TorqueParams, nat.LstmWeights, nat.StepParams, nat.PolicyParams, nat.GameParams)):
        assert nat.lib.lgk_struct_size(which) == ctypes.sizeof(cls)
    import re
    header = open(os.path.join(ROOT, "include", "lgk.h")).read()
    declared = int(re.search(r"#define LGK_ABI_VERSION (\d+)", header).group(1))
    assert nat.lib.lgk_abi_version() == nat.ABI_VERSION == declared


def test_argument_errors_are_codes_not_crashes():
    from legged_games_gym_b200 import _native as nat
    p = nat.StepParams()
    rc = nat.lib.lgk_post_physics(ctypes.byref(p), None)
    assert rc == 1 and b"num_envs" in nat.lib.lgk_last_error_string()
    t = nat.TorqueParams()
    assert nat.lib.lgk_compute_torques(ctypes.byref(t), None) == 1
    # the one-call form of the step validates like the two calls it stands for
    assert nat.lib.lgk_post_physics_finalize(ctypes.byref(p), None, None, None, None, None) == 1
    assert b"num_envs" in nat.lib.lgk_last_error_string()
    assert nat.lib.lgk_post_physics_finalize(None, None, None, None, None, None) == 1


def test_reward_order_is_alphabetical():
    from legged_games_gym_b200 import _native as nat
    assert nat.REWARD_TERMS == sorted(nat.REWARD_TERMS)


def test_env_fails_loudly_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from legged_games_gym_b200.envs import task_registry
    from legged_games_gym_b200.utils.helpers import get_args
    import copy
    with pytest.raises(RuntimeError):
        task_registry.make_env("a1", get_args([]), env_cfg=copy.deepcopy(task_registry.env_cfgs["a1"]))


def test_graft_entry_build_runs_on_cpu():
    """the driver's "does it build" check: compiles (or finds up to date) liblgk.so, loads it, imports the checker"""
    import __graft_entry__ as g
    g.build()


def test_unknown_task_raises_value_error():
    """task_registry.make_env on an unregistered name: ValueError like the reference (task_registry.py:87)"""
    import pytest
    from legged_games_gym_b200.envs import task_registry
    from legged_games_gym_b200.utils.helpers import get_args
    with pytest.raises(ValueError, match="was not registered"):
        task_registry.make_env("no_such_task", get_args([]))
    with pytest.raises(ValueError, match="Either 'name' or 'train_cfg'"):
        task_registry.make_alg_runner(env=None, name=None, args=get_args([]), train_cfg=None)


def test_actor_critic_configurations_outside_the_kernel_use_the_modules():
    """rsl_rl's ActorCritic options the fused kernel does not implement (other activations, depths, unequal widths) keep
    the API and the state_dict layout and run through the torch modules; the kernel's own configuration refuses CPU
    tensors loudly instead of falling back."""
    import pytest
    import torch
    from legged_games_gym_b200.rsl_rl.modules import ActorCritic
    ac = ActorCritic(10, 12, 3, [32, 16], [64, 32, 16], activation="tanh")
    assert not ac.fusable and sorted(ac.state_dict())[:2] == ["actor.0.bias", "actor.0.weight"]
    o, c = torch.randn(5, 10), torch.randn(5, 12)
    with torch.inference_mode():
        out = ac.act_and_evaluate(o, c)
        assert {k: tuple(v.shape) for k, v in out.items()} == dict(actions=(5, 3), mean=(5, 3), sigma=(5, 3), values=(5, 1), logp=(5,))
        assert torch.equal(ac.act_inference(o), ac.actor(o)) and torch.equal(ac.evaluate(c), ac.critic(c))
        assert torch.allclose(ac.get_actions_log_prob(out["actions"]), out["logp"])
    with pytest.raises(ValueError):
        ActorCritic(10, 10, 3, [32, 16, 8], [32, 16, 8], activation="swish")
    fused = ActorCritic(10, 10, 3, [128, 64, 32], [128, 64, 32])
    assert fused.fusable
    with torch.inference_mode(), pytest.raises(RuntimeError, match="CUDA"):
        fused.act_inference(torch.randn(4, 10))


def test_affinity_helpers_degrade_quietly_without_a_gpu():
    """utils/affinity.py: with no NVML answer (this container has no GPU) nothing is bound and the affinity is untouched."""
    import os
    from legged_games_gym_b200.utils import affinity
    before = sorted(os.sched_getaffinity(0))
    assert affinity.gpu_cpu_affinity(0) == [] or set(affinity.gpu_cpu_affinity(0)) <= set(range(os.cpu_count()))
    prev, new = affinity.bind_to_gpu(0)
    assert prev == before
    affinity.restore_affinity(prev)
    assert sorted(os.sched_getaffinity(0)) == before
