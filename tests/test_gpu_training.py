"""-m gpu: the reference's user-facing flow on top of the hot path:
scripts/train.py (task_registry.make_env + make_alg_runner + OnPolicyRunner.learn, PPO.update in torch autograd),
checkpoints with the reference's model_<it>.pt layout, scripts/play.py (resume, inference policy, TorchScript export)."""
import glob
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_checkpoint_play_export(tmp_path):
    from legged_games_gym_b200.envs import task_registry
    from legged_games_gym_b200.utils import get_args
    from legged_games_gym_b200.scripts import play as play_mod
    args = get_args(["--task", "anymal_c_flat", "--num_envs", "256", "--headless", "--max_iterations", "2", "--seed", "3"])
    env, env_cfg = task_registry.make_env(name=args.task, args=args)
    assert env.num_envs == 256 and env.num_obs == 48
    log_root = str(tmp_path / "logs")
    runner, train_cfg = task_registry.make_alg_runner(env=env, name=args.task, args=args, log_root=log_root)
    before = [p.detach().clone() for p in runner.alg.actor_critic.parameters()]
    vl, sl = runner.learn(num_learning_iterations=2, init_at_random_ep_len=True)
    assert all(map(lambda x: x == x, (vl, sl)))                               # finite losses
    after = list(runner.alg.actor_critic.parameters())
    assert any(not torch.equal(a, b) for a, b in zip(after, before))          # PPO.update moved the weights
    assert all(torch.isfinite(p).all() for p in after)
    ckpts = sorted(glob.glob(os.path.join(log_root, "*", "model_*.pt")))
    assert [os.path.basename(c) for c in ckpts] == ["model_0.pt", "model_2.pt"]   # save_interval + final, HLP:103-125 layout
    d = torch.load(ckpts[-1], map_location="cpu")
    assert set(d) == {"model_state_dict", "optimizer_state_dict", "iter", "infos"} and d["iter"] == 2
    assert "actor.0.weight" in d["model_state_dict"] and "std" in d["model_state_dict"]
    # play: resume from the last run, roll the inference policy, export TorchScript
    args2 = get_args(["--task", "anymal_c_flat", "--headless"])
    env2, logger, exported = play_mod.play(args2, num_steps=120, log_root=log_root)
    assert env2.num_envs == 50 and len(logger.state_log["dof_pos"]) == 100
    jit = torch.jit.load(os.path.join(exported, "policy_1.pt"))
    obs = env2.get_observations().cpu()
    want = runner.alg.actor_critic.actor(obs.to(env.device)).cpu()
    assert torch.allclose(jit(obs), want, rtol=1e-4, atol=1e-5)
    # the rollout-time fused kernel agrees with the autograd modules it replaces
    with torch.inference_mode():
        fused = runner.alg.actor_critic.act_inference(obs.to(env.device)).cpu()
    assert torch.allclose(fused, want, rtol=1e-3, atol=1e-3)
