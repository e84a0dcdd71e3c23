# A/B of the finalize pass: riding in K2's grid as CTA 0 (default on rough terrain up to 16 384 envs) against the stand-alone
# kernel (LGK_NO_FUSED_FINALIZE=1), over the bench's batch sizes and tasks
for v in "" 1; do
  if [ -n "$v" ]; then export LGK_NO_FUSED_FINALIZE=1; else unset LGK_NO_FUSED_FINALIZE; fi
  timeout 200 python bench.py --steps 400 --warmup 10 --no-train --no-cpu-baseline 2>/dev/null > gpurun_out/fin_ab.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/fin_ab.json").read().strip().splitlines()[-1])
c=d.get("configs", {})
print("no_fused_finalize='$v' us/step", round(d["ms_per_step"]*1e3, 2), "launches", d["gpu_launches"],
      {k: round(v["ms_per_step"]*1e3, 2) for k, v in d.get("sweep", {}).items()},
      {k: round(v["ms_per_step"]*1e3, 2) for k, v in c.items() if "ms_per_step" in v},
      "flat64 lstm/pd", c["anymal_c_flat_64"]["lstm"]["gpu_us_per_step"], c["anymal_c_flat_64"]["pd"]["gpu_us_per_step"],
      "game", (d.get("game_phase") or {}).get("us_per_step"), "e2e", d["e2e"]["us_per_step"])
PY
done
