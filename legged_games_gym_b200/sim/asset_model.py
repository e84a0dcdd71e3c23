"""Robot constants the reference obtains from Isaac Gym's URDF loader
(legged_robot.py:685-750: dof/body names, limits, index sets).  There is no PhysX here, so the few
numbers the hot path needs are tabulated from the URDFs under resources/robots/ of the reference
(joint <limit> tags; bodies after collapse_fixed_joints with the ``dont_collapse`` feet kept).
Body/DOF ORDER is decided by Isaac Gym's loader and is not derivable from the repo; the tables use
the depth-first, alphabetical-sibling order Isaac Gym is known to produce and every index set stays a
runtime tensor exactly as in the reference, so a real sim backend can overwrite them."""
from dataclasses import dataclass
from typing import List


@dataclass
class RobotModel:
    name: str
    dof_names: List[str]
    body_names: List[str]
    dof_lower: List[float]
    dof_upper: List[float]
    dof_vel_limits: List[float]
    torque_limits: List[float]

    @property
    def num_dof(self):
        return len(self.dof_names)

    @property
    def num_bodies(self):
        return len(self.body_names)

    def indices(self, substrings):
        if isinstance(substrings, str):
            substrings = [substrings]
        out = []
        for s in substrings:
            out.extend(i for i, b in enumerate(self.body_names) if s in b)
        return out

    def consts(self, asset_cfg):
        """dict in the shape the oracle / parity tests use."""
        return dict(dof_names=list(self.dof_names), num_bodies=self.num_bodies,
                    feet_indices=self.indices(asset_cfg.foot_name),
                    penalised_contact_indices=self.indices(list(asset_cfg.penalize_contacts_on)),
                    termination_contact_indices=self.indices(list(asset_cfg.terminate_after_contacts_on)),
                    dof_lower=list(self.dof_lower), dof_upper=list(self.dof_upper),
                    dof_vel_limits=list(self.dof_vel_limits), torque_limits=list(self.torque_limits))


def _quadruped(name, legs, joints, links, base, lower, upper, vel, effort):
    dofs, bodies = [], [base]
    for leg in legs:
        dofs += [leg + "_" + j for j in joints]
        bodies += [leg + "_" + l for l in links]
    n = len(legs)
    return RobotModel(name, dofs, bodies, lower * n, upper * n, vel * n, effort * n)


_ANY_LEGS = ["LF", "LH", "RF", "RH"]
_A1_LEGS = ["FL", "FR", "RL", "RR"]

MODELS = {
    # anymal_c.urdf declares effort/velocity only (no lower/upper): +-9.42 as for anymal_b
    "anymal_c": _quadruped("anymal_c", _ANY_LEGS, ["HAA", "HFE", "KFE"], ["HIP", "THIGH", "SHANK", "FOOT"], "base",
                           [-9.42] * 3, [9.42] * 3, [20.] * 3, [80.] * 3),
    "anymal_b": _quadruped("anymal_b", _ANY_LEGS, ["HAA", "HFE", "KFE"], ["HIP", "THIGH", "SHANK", "FOOT"], "base",
                           [-9.42] * 3, [9.42] * 3, [15., 20., 20.], [80.] * 3),
    "a1": _quadruped("a1", _A1_LEGS, ["hip_joint", "thigh_joint", "calf_joint"], ["hip", "thigh", "calf", "foot"],
                     "base", [-0.802851455917, -1.0471975512, -2.69653369433],
                     [0.802851455917, 4.18879020479, -0.916297857297], [52.4, 28.6, 28.6], [20., 55., 55.]),
}

_CASSIE_J = ["hip_abduction", "hip_rotation", "hip_flexion", "thigh_joint", "ankle_joint", "toe_joint"]
MODELS["cassie"] = RobotModel(
    "cassie",
    [j + "_left" for j in _CASSIE_J] + [j + "_right" for j in _CASSIE_J],
    ["pelvis", "left_pelvis_rotation", "left_hip", "left_thigh", "left_shin", "left_tarsus", "left_toe",
     "right_pelvis_rotation", "right_hip", "right_thigh", "right_shin", "right_tarsus", "right_toe"],
    [-0.2618, -0.3927, -0.8727, -2.8623, 0.6458, -2.4435, -0.3927, -0.3927, -0.8727, -2.8623, 0.6458, -2.4435],
    [0.3927, 0.3927, 1.3963, -0.6458, 2.8623, -0.5236, 0.2618, 0.3927, 1.3963, -0.6458, 2.8623, -0.5236],
    [20.1475, 20.1475, 20.5085, 20.5085, 20.5085, 20.5192] * 2,
    [112., 112., 195., 195., 195., 45.] * 2)


def model_for_asset(asset_cfg) -> RobotModel:
    """Pick the table from cfg.asset.file (``.../robots/<name>/urdf/<name>.urdf``) or cfg.asset.name."""
    import os
    stem = os.path.splitext(os.path.basename(asset_cfg.file))[0]
    for key in (stem, asset_cfg.name):
        if key in MODELS:
            return MODELS[key]
    raise ValueError(f"no robot table for asset file '{asset_cfg.file}' / name '{asset_cfg.name}'")
