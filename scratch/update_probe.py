import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from legged_games_gym_b200.rsl_rl.modules import ActorCritic
from legged_games_gym_b200.rsl_rl.algorithms import PPO
dev = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = os.environ.get("TF32", "0") == "1"
torch.manual_seed(0)
ac = ActorCritic(235, 235, 12, [512, 256, 128], [512, 256, 128]).to(dev)
alg = PPO(ac, num_learning_epochs=5, num_mini_batches=4, schedule="adaptive", desired_kl=0.01, learning_rate=1e-3, entropy_coef=0.01, device=dev)
alg.use_cuda_graph = os.environ.get("GRAPH", "1") == "1"
alg.init_storage(4096, 24, [235], [None], [12])
st = alg.storage
for name in ("observations", "actions", "values", "returns", "advantages", "actions_log_prob", "mu"):
    getattr(st, name).normal_()
st.sigma.uniform_(0.5, 1.5)
for it in range(4):
    st.step = 24
    torch.cuda.synchronize(); t0 = time.perf_counter()
    alg.update()
    torch.cuda.synchronize(); print(f"update {it}: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
