"""Anymal: LeggedRobot + the ANYdrive SEA actuator network (mirror of reference
legged_gym/envs/anymal_c/anymal.py:46-81).  The LSTM runs in lgk_compute_torques; its h/c state keeps the reference
layout ``sea_hidden_state / sea_cell_state [2, N*12, 8]`` and is zeroed for reset envs inside the fused step."""
import ctypes as C
import os

import numpy as np
import torch

from ... import LEGGED_GYM_ROOT_DIR, _native as nat
from ..base.legged_robot import LeggedRobot, _stream_ptr


def load_actuator_weights(path):
    """Accepts the reference's TorchScript archive (.pt) or the .npz dump shipped under resources/."""
    if not os.path.exists(path) and path.endswith(".pt") and os.path.exists(path[:-3] + ".npz"):
        path = path[:-3] + ".npz"
    if path.endswith(".npz"):
        return {k: np.asarray(v, dtype=np.float32) for k, v in np.load(path).items()}
    m = torch.jit.load(path, map_location="cpu")
    w = {n.replace("lstm.", "").replace("linear.", "linear_"): p.detach().numpy().astype(np.float32)
         for n, p in m.named_parameters()}
    w.update({n: b.detach().numpy().astype(np.float32).reshape(-1) for n, b in m.named_buffers()})
    return w


def upload_actuator_weights(w):
    lw = nat.LstmWeights()
    for dst, src in (("w_ih0", "weight_ih_l0"), ("w_hh0", "weight_hh_l0"), ("b_ih0", "bias_ih_l0"),
                     ("b_hh0", "bias_hh_l0"), ("w_ih1", "weight_ih_l1"), ("w_hh1", "weight_hh_l1"),
                     ("b_ih1", "bias_ih_l1"), ("b_hh1", "bias_hh_l1"), ("lin_w", "linear_weight"),
                     ("lin_b", "linear_bias"), ("in_scale", "in_scale"), ("out_scale", "out_scale")):
        getattr(lw, dst)[:] = [float(x) for x in np.asarray(w[src], dtype=np.float32).reshape(-1)]
    nat.check(nat.lib.lgk_set_lstm_weights(C.byref(lw), _stream_ptr()), "lgk_set_lstm_weights")
    torch.cuda.current_stream().synchronize()


class Anymal(LeggedRobot):
    def __init__(self, cfg, sim_params, physics_engine, sim_device, headless, **kw):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless, **kw)
        if self.cfg.control.use_actuator_network:
            path = self.cfg.control.actuator_net_file.format(LEGGED_GYM_ROOT_DIR=LEGGED_GYM_ROOT_DIR)
            self.actuator_weights = load_actuator_weights(path)
            upload_actuator_weights(self.actuator_weights)

    def _init_buffers(self):
        super()._init_buffers()
        n = self.num_envs * self.num_actions
        self.sea_input = torch.zeros(n, 1, 2, device=self.device, requires_grad=False)
        self.sea_hidden_state = torch.zeros(2, n, 8, device=self.device, requires_grad=False)
        self.sea_cell_state = torch.zeros(2, n, 8, device=self.device, requires_grad=False)
        self.sea_hidden_state_per_env = self.sea_hidden_state.view(2, self.num_envs, self.num_actions, 8)
        self.sea_cell_state_per_env = self.sea_cell_state.view(2, self.num_envs, self.num_actions, 8)

    def _build_native_params(self):
        super()._build_native_params()
        if self.cfg.control.use_actuator_network:
            self._tq_params.use_lstm = 1
            self._tq_params.sea_hidden_state = self.sea_hidden_state.data_ptr()
            self._tq_params.sea_cell_state = self.sea_cell_state.data_ptr()
        # ANY:56-60 zeroes the state on reset whether or not the network is in use
        self._params.zero_lstm_on_reset = 1
        self._params.sea_hidden_state = self.sea_hidden_state.data_ptr()
        self._params.sea_cell_state = self.sea_cell_state.data_ptr()
