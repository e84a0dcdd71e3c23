import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from oracle import harness
from tests.util import product_env, feeder_state
for graph in (True, False):
    case = harness.build_case("a1", 1000, seed=6, overrides={"domain_rand.push_interval_s": 0.04})
    st_or = harness.torch_state(case); orc = harness.make_oracle(case, st_or)
    env, feeder = product_env(case, graph=graph, tile=16); st_gpu = feeder_state(feeder)
    for step in range(1, 5):
        tables = harness.step_tables(case["seed"], step, 1000, orc.num_obs)
        acts = torch.from_numpy(np.random.default_rng(step).normal(0, 1, (1000, 12)).astype(np.float32))
        root_before = st_or["root_states"].clone()
        orc.step(acts.clone(), tables); env.step(acts.cuda()); torch.cuda.synchronize()
        d = (env.measured_heights.cpu() != orc.measured_heights)
        idx = d.nonzero()
        print("graph", graph, "step", step, "n mismatching heights", len(idx), idx[:6].tolist())
        for e, j in idx[:3].tolist():
            px = orc.last_px.view(1000, -1)[e, j].item(); py = orc.last_py.view(1000, -1)[e, j].item()
            print("   env", e, "pt", j, "oracle idx", px, py, "root", root_before[e, :7].tolist(), "got", env.measured_heights[e, j].item(), "want", orc.measured_heights[e, j].item())
        noise = harness.make_noise(case, step, 5); harness.apply_noise(st_or, noise); harness.apply_noise(st_gpu, noise)
