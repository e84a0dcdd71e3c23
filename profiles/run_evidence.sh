#!/bin/bash
# Evidence run of a round (under gpurun, ONE GPU): sanitizer passes, ncu captures under replica rotation, launch list.
#   bash profiles/run_evidence.sh r2
set -u
R=${1:-r2}
O=gpurun_out
mkdir -p $O
python profiles/sanitize_target.py > $O/${R}_sanitize_plain.log 2>&1 || { echo "target fails without the sanitizer"; tail -5 $O/${R}_sanitize_plain.log; }
if [ "${SANITIZE:-0}" = "1" ]; then      # refused on this GPU pool (see README.md); kept for pools that allow it
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_target.py > $O/${R}_sanitize_$tool.log 2>&1
  echo "$tool rc=$?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/${R}_sanitize_$tool.log | tail -1)"
done
fi
[ -x profiles/umma_rate_probe ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/umma_rate_probe profiles/umma_rate_probe.cu
profiles/umma_rate_probe > $O/${R}_umma_rates.txt 2>&1; tail -2 $O/${R}_umma_rates.txt
python profiles/policy_probe.py 4096 16384 65536 > $O/${R}_policy_probe.txt 2>&1
python profiles/e2e_probe.py > $O/${R}_e2e_probe.txt 2>&1; python profiles/pcie_duplex_probe.py >> $O/${R}_e2e_probe.txt 2>&1
python profiles/update_profile.py > $O/${R}_update_profile.txt 2>&1
K='regex:torque_kernel|post_kernel|scan_obs|finalize'
ncu --set full --clock-control none --import-source on -k "$K" -s 14 -c 7 -o $O/${R}_env_65536 -f python profiles/prof_target.py --num-envs 65536 --steps 3 --rotate > $O/${R}_ncu_65536.log 2>&1; tail -1 $O/${R}_ncu_65536.log
ncu --set full --clock-control none --import-source on -k "$K" -s 42 -c 7 -o $O/${R}_env_4096 -f python profiles/prof_target.py --num-envs 4096 --steps 3 --rotate > $O/${R}_ncu_4096.log 2>&1; tail -1 $O/${R}_ncu_4096.log
ncu --set full --clock-control none --import-source on -k 'regex:policy_tc|policy_pack' -c 3 -o $O/${R}_policy_4096 -f python profiles/prof_policy.py --num-envs 4096 > $O/${R}_ncu_policy.log 2>&1; tail -1 $O/${R}_ncu_policy.log
ncu --set full --clock-control none --import-source on -k 'regex:policy_tc' -s 1 -c 2 -o $O/${R}_policy_65536 -f python profiles/prof_policy.py --num-envs 65536 > $O/${R}_ncu_policy65536.log 2>&1; tail -1 $O/${R}_ncu_policy65536.log
python bench.py --steps 60 --warmup 5 --no-sweep --no-cpu-baseline --no-train > $O/${R}_bench_plain.json 2> $O/${R}_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${R}_launches_bench_4096.csv python bench.py --steps 60 --warmup 5 --no-sweep --no-cpu-baseline --no-train > $O/${R}_ncu_launches.log 2>&1
echo "launch list rows: $(wc -l < $O/${R}_launches_bench_4096.csv)"
