"""Small end-to-end target for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every kernel of the hot path
at sizes the tools finish in seconds -- K1 fast path (full tiles, several tiles per persistent CTA via LGK_K1_CTAS_PER_SM=1),
K1 generic path (ragged tile, two actors per env, split PRE / POST phases, reset_idx override), K2 hot + generic kernels,
flat-task K1 tail, explicit reset, the game kernel, the tcgen05 policy kernel, GAE.
    compute-sanitizer --tool racecheck python profiles/sanitize_target.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LGK_K1_CTAS_PER_SM", "1")
import numpy as np  # noqa: E402
import torch  # noqa: E402
from oracle import harness  # noqa: E402
from tests.util import product_env, product_game, feeder_state  # noqa: E402

DEV = "cuda:0"


def steps(task, n, ov=None, k=3, graph=False):
    case = harness.build_case(task, n, seed=3, overrides=ov)
    env, feeder = product_env(case, graph=graph)
    for s in range(1, k + 1):
        a = torch.from_numpy(np.random.default_rng(s).normal(0, 1, (n, 12)).astype(np.float32)).to(DEV)
        env.step(a)
    torch.cuda.synchronize()
    return env


short = {"env.episode_length_s": 0.06, "domain_rand.push_interval_s": 0.04}
steps("anymal_c_rough", 148 * 32 * 2 + 64, short)            # fast path, > 2 tiles per CTA at one CTA per SM
steps("anymal_c_rough", 77, short)                          # ragged tile: generic instantiation
steps("anymal_c_flat", 96, short)                           # flat: K1 finishes the rows
steps("cassie", 130, short)
steps("low_level_game", 90, short)                          # two actors per env, predator re-spawn
steps("a1", 160, dict(short, **{"commands.curriculum": True}), k=6)
env = steps("anymal_c_rough", 200, short, k=1)
env.reset()                                                  # explicit reset_idx + one step
from legged_games_gym_b200.envs import LeggedRobot, task_registry  # noqa: E402
from legged_games_gym_b200.envs.a1.a1_config import A1RoughCfg  # noqa: E402


class Override(LeggedRobot):
    def reset_idx(self, env_ids):
        super().reset_idx(env_ids)


task_registry.register("a1_sanitize_override", Override, A1RoughCfg(), None)
case = harness.build_case("a1", 100, seed=4, overrides=short)
case["task"] = "a1_sanitize_override"
env, _ = product_env(case)
for s in range(1, 4):
    env.step(torch.zeros(100, 12, device=DEV))             # PRE / POST_REWARD / reset_idx / POST_OBS
# games
case = harness.build_case("low_level_game", 64, seed=9, overrides={"env.episode_length_s": 0.1})
st = harness.torch_state(case)
harness.place_predators(st, 9)
case["state"]["root_states"] = st["root_states"].numpy().copy()
box = {"a": torch.zeros(64, 12, device=DEV)}
game, feeder = product_game(case, "hl", None, ll_policy=lambda obs: box["a"])
game.predator_pos.copy_(feeder_state(feeder)["root_states"][1::2, :3])
for s in range(3):
    game.step(torch.randn(64, 6, device=DEV))
# policy (tcgen05) + GAE
from legged_games_gym_b200.rsl_rl.modules import ActorCritic  # noqa: E402
from legged_games_gym_b200 import _native as nat  # noqa: E402
ac = ActorCritic(235, 235, 12, [512, 256, 128], [512, 256, 128]).to(DEV)
with torch.inference_mode():
    out = ac.act_and_evaluate(torch.randn(300, 235, device=DEV), torch.randn(300, 235, device=DEV))
    ac.act_inference(torch.randn(300, 235, device=DEV))
T, N = 8, 300
r, v = torch.randn(T, N, 1, device=DEV), torch.randn(T, N, 1, device=DEV)
d = (torch.rand(T, N, 1, device=DEV) < 0.1).to(torch.uint8)
lv = torch.randn(N, 1, device=DEV)
ret, adv = torch.empty_like(r), torch.empty_like(r)
scr = torch.zeros(4, dtype=torch.float64, device=DEV)
nat.check(nat.lib.lgk_gae(r.data_ptr(), v.data_ptr(), d.data_ptr(), lv.data_ptr(), T, N, 0.99, 0.95, ret.data_ptr(), adv.data_ptr(),
                          scr.data_ptr(), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("sanitize target ok", nat.launch_count(), "launches")
