"""B200-side stand-in for the parts of rsl_rl (v1.0.2 API; third-party, not vendored by the reference -- call sites
legged_gym/utils/task_registry.py:37-38, 154) that sit on or next to the hot path: ActorCritic.act / evaluate and
RolloutStorage.compute_returns run as CUDA kernels; PPO.update stays plain PyTorch autograd and is where the NCCL
gradient all-reduce hooks in."""
