#!/bin/bash
# Final evidence of round 2 after ABI 7 (finalize riding in K2's grid): the default bench line, the launch list of the
# bench command and fresh ncu --set full captures of the env-step kernels.    bash profiles/run_evidence_final.sh
set -u
O=gpurun_out
mkdir -p $O
python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; echo "bench rc=$?"; tail -c 300 $O/r2_bench_n1.json; echo
K='regex:torque_kernel|post_kernel|scan_obs|finalize'
ncu --set full --clock-control none --import-source on -k "$K" -s 36 -c 12 -o $O/r2_env_4096 -f python profiles/prof_target.py --num-envs 4096 --steps 3 --rotate > $O/r2_ncu_4096.log 2>&1; tail -1 $O/r2_ncu_4096.log
ncu --set full --clock-control none --import-source on -k "$K" -s 14 -c 7 -o $O/r2_env_65536 -f python profiles/prof_target.py --num-envs 65536 --steps 3 --rotate > $O/r2_ncu_65536.log 2>&1; tail -1 $O/r2_ncu_65536.log
python bench.py --steps 60 --warmup 5 --no-sweep --no-cpu-baseline --no-train > $O/r2_bench_plain.json 2> $O/r2_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r2_launches_bench_4096.csv python bench.py --steps 60 --warmup 5 --no-sweep --no-cpu-baseline --no-train > $O/r2_ncu_launches.log 2>&1
echo "launch list rows: $(wc -l < $O/r2_launches_bench_4096.csv)"
