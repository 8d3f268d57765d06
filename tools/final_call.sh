#!/bin/bash
# Last one-GPU call of the round: GPU tests, the default bench line, pass-2 grid A/B, launch list, ncu --set full.
set -u
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -q -m gpu -rf -x > gpurun_out/pytest_gpu.txt 2>&1; tail -6 gpurun_out/pytest_gpu.txt | cut -c1-220
T0=$(date +%s)
timeout -k 10 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "default bench wall seconds: $(( $(date +%s) - T0 ))"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_default.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "steps", "recall_at_k", "recall_queries", "parity", "self_parity")})
    print({k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
    print(d["e2e"]["value"], d["concurrent"]["value"], d["clocks"], d["cpu_baseline"])
except Exception as e:
    print("bench output unreadable:", e)
PY
timeout 300 python tools/variant_sweep.py --steps 10 --only "pass-2 grid 16,pass-2 grid 32" 2>&1 | tee gpurun_out/variant_sweep.txt
B="python bench.py --steps 2 --warmup 1 --no-cpu --recall-queries 0 --threads 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_" -c 700 --csv \
    --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
timeout 500 ncu --set full --import-source on --clock-control none \
    -k "regex:^k_(maxsim_tc|pair_exact|scores16_tc|recheck_pairs|recheck_dots|approx16|select_u32)" -s 7 -c 8 -o gpurun_out/r02_final2 \
    $B > gpurun_out/ncu_final2.log 2>&1
tail -c 200 gpurun_out/ncu_final2.log
