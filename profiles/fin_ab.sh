# A/B of the finalize pass: fused into K2's grid (default) against the stand-alone kernel (LGK_NO_FUSED_FINALIZE=1)
set -x
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "env_step or explicit_reset or full_size or user_reward or reset_idx_override or deterministic or host_sim or host_result" 2>&1 | tail -5
for v in "" 1 "" 1; do
  if [ -n "$v" ]; then export LGK_NO_FUSED_FINALIZE=1; else unset LGK_NO_FUSED_FINALIZE; fi
  python bench.py --steps 400 --warmup 10 --no-train --no-cpu-baseline 2>/dev/null > gpurun_out/fin_ab.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/fin_ab.json").read().strip().splitlines()[-1])
print("no_fused_finalize='$v' us/step", round(d["ms_per_step"]*1e3, 2), "launches", d["gpu_launches"],
      {k: round(v["ms_per_step"]*1e3, 2) for k, v in d.get("sweep", {}).items()},
      {k: round(v["ms_per_step"]*1e3, 2) for k, v in d.get("configs", {}).items() if "ms_per_step" in v}, "e2e", d["e2e"]["us_per_step"])
PY
done
