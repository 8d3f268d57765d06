#!/bin/bash
# Third GPU call, one box, ~7 minutes: GPU tests, short bench + nq = 48 / exact twins, launch list.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -rf -x > gpurun_out/pytest_gpu.txt 2>&1; tail -12 gpurun_out/pytest_gpu.txt | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --recall-queries 32 --parity-queries 16 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
tail -2 gpurun_out/bench_c3.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_c3.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "recall_at_k", "parity", "self_parity")})
    print({k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
    print(d["e2e"]["value"], d["concurrent"]["value"], d["clocks"])
except Exception as e:
    print("bench output unreadable:", e)
PY
timeout 300 python tools/variant_sweep.py --steps 10 --only "exact fp32,nq=48" 2>&1 | tee gpurun_out/variant_sweep.txt
B="python bench.py --steps 2 --warmup 1 --no-cpu --recall-queries 0 --threads 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_" -c 600 --csv \
    --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_" -c 600 --csv \
    --log-file gpurun_out/r02_launches_nq48.csv $B --nq 48 > gpurun_out/ncu_launches48.log 2>&1
tail -c 200 gpurun_out/ncu_launches48.log
