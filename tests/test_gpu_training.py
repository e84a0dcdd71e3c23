"""-m gpu: the reference's user-facing flow on top of the hot path:
scripts/train.py (task_registry.make_env + make_alg_runner + OnPolicyRunner.learn, PPO.update in torch autograd),
checkpoints with the reference's model_<it>.pt layout, scripts/play.py (resume, inference policy, TorchScript export)."""
import glob
import os

import numpy as np
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
    assert torch.allclose(jit(obs), want, rtol=1e-3, atol=1e-3)    # CPU fp32 TorchScript vs the GPU modules (TF32 matmuls)
    # the rollout-time fused kernel agrees with the autograd modules it replaces
    with torch.inference_mode():
        fused = runner.alg.actor_critic.act_inference(obs.to(env.device)).cpu()
    assert torch.allclose(fused, want, rtol=1e-3, atol=1e-3)


def test_graphed_ppo_update_matches_eager_update():
    """PPO.update as one CUDA graph per mini-batch (device-side Adam step counter, learning-rate rule, loss accumulators)
    must do what the eager rsl_rl loop does: same data, same permutation -> same parameters, losses and learning rate."""
    from legged_games_gym_b200.rsl_rl.modules import ActorCritic
    from legged_games_gym_b200.rsl_rl.algorithms import PPO
    dev = "cuda:0"

    def make(graph):
        torch.manual_seed(0)
        ac = ActorCritic(48, 48, 12, [128, 64, 32], [128, 64, 32]).to(dev)
        alg = PPO(ac, num_learning_epochs=2, num_mini_batches=4, schedule="adaptive", desired_kl=0.01, learning_rate=1e-3,
                  entropy_coef=0.01, device=dev, tf32_matmul=False)      # strict fp32: the two paths must agree tightly
        alg.use_cuda_graph = graph
        alg.init_storage(256, 8, [48], [None], [12])
        g = torch.Generator().manual_seed(1)
        st = alg.storage
        for name in ("observations", "actions", "values", "returns", "advantages", "actions_log_prob", "mu"):
            t = getattr(st, name)
            t.copy_(torch.randn(t.shape, generator=g))
        st.sigma.copy_(torch.rand(st.sigma.shape, generator=g) + 0.5)
        st.step = 8
        return alg

    results = []
    for graph in (False, True):
        alg = make(graph)
        out = []
        for it in range(3):                         # three updates: the graph is built in the first, replayed in the others
            alg.storage.step = 8
            torch.manual_seed(100 + it)             # same mini-batch permutation in both runs
            out.append(alg.update())
        results.append((out, [p.detach().clone() for p in alg.actor_critic.parameters()], alg.learning_rate))
    (l0, p0, lr0), (l1, p1, lr1) = results
    assert lr0 == pytest.approx(lr1, rel=1e-6)
    for a, b in zip(l0, l1):
        assert a[0] == pytest.approx(b[0], rel=1e-4, abs=1e-6) and a[1] == pytest.approx(b[1], rel=1e-4, abs=1e-6)
    for a, b in zip(p0, p1):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)
    assert any(not torch.equal(a, torch.zeros_like(a)) for a in p1)


def _random_ll_policy(dev, seed=0):
    """a frozen low-level locomotion policy with the A1 cfg's shape (the reference loads a trained one from a checkpoint
    that is not in its tree)"""
    from legged_games_gym_b200.rsl_rl.modules import ActorCritic
    torch.manual_seed(seed)
    ac = ActorCritic(235, 235, 12, [512, 256, 128], [512, 256, 128]).to(dev).eval()
    return ac.act_inference


def test_high_level_game_trains_and_plays(tmp_path):
    """scripts/train.py and scripts/play_game.py flow on high_level_game: OnPolicyRunner on the 19-float observation /
    6-command action space, checkpoint, reload, TorchScript export, roll-out."""
    from legged_games_gym_b200.scripts.play_game import play_game
    from legged_games_gym_b200.envs import task_registry
    from legged_games_gym_b200.utils import get_args
    dev = "cuda:0"
    root = str(tmp_path / "logs")
    a = get_args(["--task", "high_level_game", "--num_envs", "256", "--headless", "--max_iterations", "2", "--sim_device", dev, "--rl_device", dev])
    env, _ = task_registry.make_env(name=a.task, args=a, ll_policy=_random_ll_policy(dev))
    runner, train_cfg = task_registry.make_alg_runner(env=env, name=a.task, args=a, log_root=root)
    v, s_ = runner.learn(num_learning_iterations=2, init_at_random_ep_len=True)
    assert np.isfinite(v) and np.isfinite(s_)
    assert env.obs_buf.shape == (256, 19) and torch.isfinite(env.obs_buf).all() and torch.isfinite(env.rew_buf).all()
    assert int(env.curr_episode_step.max()) > 0
    env2, exported = play_game(a, num_steps=20, log_root=root, ll_policy=_random_ll_policy(dev))
    assert os.path.exists(os.path.join(exported, "policy_1.pt"))
    assert env2.cfg.terrain.mesh_type == "plane"         # play_game's overrides (an explicit --num_envs wins, as in the reference)
    assert torch.isfinite(env2.obs_buf).all()


def test_dec_high_level_game_trains_and_plays(tmp_path):
    """scripts/train_dec_game.py and scripts/play_dec_game.py flow: two PPO agents in alternating evolutions, per-agent
    checkpoints (pred_model_*.pt / prey_model_*.pt), reload through make_dec_alg_runner."""
    from legged_games_gym_b200.scripts.play_dec_game import play_dec_game
    from legged_games_gym_b200.envs import task_registry
    from legged_games_gym_b200.utils import get_args
    dev = "cuda:0"
    root = str(tmp_path / "logs")
    a = get_args(["--task", "dec_high_level_game", "--num_envs", "256", "--headless", "--sim_device", dev, "--rl_device", dev])
    env, _ = task_registry.make_env(name=a.task, args=a, ll_policy=_random_ll_policy(dev))
    runner, train_cfg = task_registry.make_dec_alg_runner(env=env, name=a.task, args=a, log_root=root)
    before = [p.detach().clone() for alg in runner.algs for p in alg.actor_critic.actor.parameters()]
    losses = runner.learn(max_num_evolutions=2, num_learning_iterations=1, init_at_random_ep_len=True)
    assert all(np.isfinite(x) for x in losses)
    after = [p.detach() for alg in runner.algs for p in alg.actor_critic.actor.parameters()]
    assert any(not torch.equal(x, y) for x, y in zip(before[:len(before) // 2], after[:len(after) // 2])), "predator did not learn"
    assert any(not torch.equal(x, y) for x, y in zip(before[len(before) // 2:], after[len(after) // 2:])), "prey did not learn"
    run = os.path.join(root, sorted(os.listdir(root))[-1])
    assert {"pred_model_1.pt", "prey_model_1.pt"} <= set(os.listdir(run))
    assert "rew_prey_evasion" in env.extras["episode"] and "rew_pred_pursuit" in env.extras["episode"]
    env2 = play_dec_game(a, num_steps=20, log_root=root, ll_policy=_random_ll_policy(dev))
    assert env2.cfg.noise.add_noise is False
    assert torch.isfinite(env2.obs_buf_prey).all() and torch.isfinite(env2.obs_buf_pred).all()
