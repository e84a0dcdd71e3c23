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


# ---------------------------------------------------------------------------------------------- rollout bookkeeping
def test_rollout_storage_holds_the_observation_the_policy_saw():
    """The env's obs_buf is ONE persistent tensor that every step rewrites in place (the reference's env rebinds it, so
    rsl_rl may keep a reference): PPO.act must snapshot it, otherwise storage.observations[s] would hold obs_{s+1} next to
    the actions / log-probs / values computed from obs_s."""
    from legged_games_gym_b200.envs import task_registry
    from legged_games_gym_b200.utils import get_args
    args = get_args(["--task", "anymal_c_flat", "--num_envs", "128", "--headless", "--seed", "5"])
    env, _ = task_registry.make_env(name=args.task, args=args)
    import copy
    train_cfg = copy.deepcopy(task_registry.train_cfgs[args.task])
    train_cfg.runner.resume = False            # (play() of an earlier test flips the registry's shared cfg object)
    runner, _ = task_registry.make_alg_runner(env=env, args=args, train_cfg=train_cfg, log_root=None)
    alg = runner.alg
    obs = env.get_observations()
    fed, acted = [], []
    with torch.inference_mode():
        for s in range(6):
            alg.actor_critic.set_rng(1, s + 1, 0)
            fed.append(obs.clone())
            a = alg.act(obs, obs)
            acted.append(a.clone())
            obs, _, rew, dones, infos = env.step(a)
            alg.process_env_step(rew, dones, infos)
    assert obs.data_ptr() == env.obs_buf.data_ptr()
    for s in range(6):
        assert torch.equal(alg.storage.observations[s], fed[s]), f"slot {s} does not hold the observation fed to act()"
        assert torch.equal(alg.storage.actions[s], acted[s])
    assert not torch.equal(fed[0], fed[1])           # the env really moved between the steps
    # the stored pairs are consistent: the stored mean is the actor's output on the STORED observation
    mu = alg.actor_critic.actor(alg.storage.observations[3])
    assert torch.allclose(mu, alg.storage.mu[3], rtol=1e-3, atol=1e-3)


def test_rollout_policy_follows_graphed_updates():
    """A CUDA-graph replay moves no torch version counter: after several graphed PPO updates of ONE policy (no other
    policy evicting the packed-weight cache) the fused rollout kernel must run the CURRENT weights."""
    from legged_games_gym_b200.rsl_rl.modules import ActorCritic
    from legged_games_gym_b200.rsl_rl.algorithms import PPO
    dev = "cuda:0"
    torch.manual_seed(0)
    ac = ActorCritic(235, 235, 12, [512, 256, 128], [512, 256, 128]).to(dev)
    alg = PPO(ac, num_learning_epochs=2, num_mini_batches=2, learning_rate=1e-2, device=dev)
    alg.init_storage(256, 8, [235], [None], [12])
    g = torch.Generator().manual_seed(1)
    obs = torch.randn(256, 235, generator=g).to(dev)
    with torch.inference_mode():
        first = ac.act_inference(obs).clone()
    for it in range(4):
        st = alg.storage
        for name in ("observations", "actions", "values", "returns", "advantages", "actions_log_prob", "mu"):
            t = getattr(st, name)
            t.copy_(torch.randn(t.shape, generator=g))
        st.sigma.copy_(torch.rand(st.sigma.shape, generator=g) + 0.5)
        st.step = 8
        alg.update()
        assert alg._graph is not None                 # the graphed path is the one under test
        with torch.inference_mode():
            fused = ac.act_inference(obs)
            want = ac.actor(obs)
        assert torch.allclose(fused, want, rtol=1e-3, atol=2e-3), f"update {it}: rollout kernel runs stale weights"
    assert not torch.allclose(first, fused, atol=1e-2)  # and the weights really moved


def test_graph_build_keeps_adam_state_and_checkpoint_is_portable(tmp_path):
    """Switching from the eager optimizer to the capturable one (first graphed update, e.g. after --resume) must carry the
    Adam moments and step over; the saved optimizer state loads into a stock torch.optim.Adam (the reference's format)."""
    from legged_games_gym_b200.rsl_rl.modules import ActorCritic
    from legged_games_gym_b200.rsl_rl.algorithms import PPO
    dev = "cuda:0"
    torch.manual_seed(0)
    ac = ActorCritic(48, 48, 12, [128, 64, 32], [128, 64, 32]).to(dev)
    alg = PPO(ac, num_learning_epochs=1, num_mini_batches=2, learning_rate=1e-3, device=dev, tf32_matmul=False)
    alg.init_storage(64, 4, [48], [None], [12])
    g = torch.Generator().manual_seed(1)

    def fill():
        st = alg.storage
        for name in ("observations", "actions", "values", "returns", "advantages", "actions_log_prob", "mu"):
            t = getattr(st, name)
            t.copy_(torch.randn(t.shape, generator=g))
        st.sigma.copy_(torch.rand(st.sigma.shape, generator=g) + 0.5)
        st.step = 4
    alg.use_cuda_graph = False
    fill(); alg.update(); fill(); alg.update()                    # 4 eager Adam steps
    p0 = next(iter(ac.parameters()))
    m_before = alg.optimizer.state[p0]["exp_avg"].clone()
    assert float(alg.optimizer.state[p0]["step"]) == 4 and m_before.abs().sum() > 0
    alg.use_cuda_graph = True
    alg._build_graph()
    st = alg.optimizer.state[p0]
    assert float(st["step"]) == 4, "Adam step counter was reset by the graph build"
    assert torch.equal(st["exp_avg"], m_before), "Adam first moment was reset by the graph build"
    fill(); alg.update()
    assert float(alg.optimizer.state[p0]["step"]) == 6
    sd = alg.optimizer_state_dict()
    torch.save(sd, tmp_path / "opt.pt")
    sd = torch.load(tmp_path / "opt.pt", map_location="cpu")
    assert isinstance(sd["param_groups"][0]["lr"], float) and not sd["param_groups"][0]["capturable"]
    stock = torch.optim.Adam([torch.nn.Parameter(p.detach().cpu().clone()) for p in ac.parameters()], lr=1e-3)
    stock.load_state_dict(sd)                                       # the reference's rsl_rl does exactly this on resume
    assert float(stock.state[stock.param_groups[0]["params"][0]]["step"]) == 6
    # and back: a stock checkpoint loads into the live graphed optimizer IN PLACE (the graph keeps the tensors' addresses)
    ptr = alg.optimizer.state[p0]["exp_avg"].data_ptr()
    alg.load_optimizer_state_dict(stock.state_dict())
    assert alg.optimizer.state[p0]["exp_avg"].data_ptr() == ptr and float(alg.optimizer.state[p0]["step"]) == 6
    fill(); alg.update()
    assert float(alg.optimizer.state[p0]["step"]) == 8


def test_actor_only_calls_never_touch_the_critic_input():
    """act() / act_inference() with num_critic_obs != num_obs (privileged observations): the kernel must not read
    num_critic_obs columns through the actor's observation pointer; alternating policies keep their packed weights."""
    from legged_games_gym_b200 import _native as nat
    from legged_games_gym_b200.rsl_rl.modules import ActorCritic
    dev = "cuda:0"
    torch.manual_seed(0)
    ac = ActorCritic(48, 235, 12, [512, 256, 128], [512, 256, 128]).to(dev)
    other = ActorCritic(235, 235, 12, [512, 256, 128], [512, 256, 128]).to(dev)
    obs = torch.randn(300, 48, device=dev)                 # exactly sized: an out-of-bounds read would run past it
    cobs = torch.randn(300, 235, device=dev)
    big = torch.randn(300, 235, device=dev)
    with torch.inference_mode():
        mean = ac.act_inference(obs)
        assert torch.allclose(mean, ac.actor(obs), rtol=1e-3, atol=2e-3)
        out = ac.act_and_evaluate(obs, cobs)
        assert torch.allclose(out["values"], ac.critic(cobs), rtol=1e-3, atol=2e-3)
        assert torch.allclose(out["mean"], mean, rtol=0, atol=0)
        # evaluate() is never served from a stale cache keyed by the tensor's address
        cobs.mul_(0.5)
        assert torch.allclose(ac.evaluate(cobs), ac.critic(cobs))
        other.act_inference(big)
        torch.cuda.synchronize()
        l0 = nat.launch_count()
        for _ in range(5):                                  # the games alternate 2-3 policies every step
            ac.act_inference(obs)
            other.act_inference(big)
        assert nat.launch_count() - l0 == 10, "alternating policies re-packed their weights"
    with pytest.raises(ValueError):
        ac.act_and_evaluate(obs, obs)                       # wrong critic width is refused, not read out of bounds
