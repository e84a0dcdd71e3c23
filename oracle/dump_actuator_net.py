"""Extract the SEA actuator-net tensors from the reference's TorchScript asset
(/root/reference/resources/actuator_nets/anydrive_v3_lstm.pt, pointed to by
anymal_c_rough_config.py:69) into a plain .npz.  The weights are an input ASSET of the path
(SURVEY.md section 2 row 3), not source; the .npz travels with the repo because /root/reference
does not exist on the GPU box.  Run once in the build container:

    python -m oracle.dump_actuator_net
"""
import os
import numpy as np
import torch

SRC = "/root/reference/resources/actuator_nets/anydrive_v3_lstm.pt"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "legged_games_gym_b200", "resources", "actuator_nets", "anydrive_v3_lstm.npz")


def extract(path=SRC):
    m = torch.jit.load(path, map_location="cpu")
    out = {}
    for n, p in m.named_parameters():
        out[n.replace("lstm.", "").replace("linear.", "linear_")] = p.detach().numpy().astype(np.float32)
    for n, b in m.named_buffers():
        out[n] = b.detach().numpy().astype(np.float32).reshape(-1)
    return out


if __name__ == "__main__":
    w = extract()
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    np.savez(DST, **w)
    print({k: v.shape for k, v in w.items()}, "->", DST)
