#!/bin/bash
# Short one-GPU check: GPU tests, a 10-step bench, pass-2 grid A/B, launch list, sanitizer slice.
set -u
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -q -m gpu -rf -x > gpurun_out/pytest_gpu.txt 2>&1; tail -6 gpurun_out/pytest_gpu.txt | cut -c1-220
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --recall-queries 32 --parity-queries 16 > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_check.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "recall_at_k", "parity", "self_parity")})
    print({k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
    print(d["e2e"]["value"], d["concurrent"]["value"], d["clocks"])
except Exception as e:
    print("bench output unreadable:", e)
PY
timeout 400 python tools/variant_sweep.py --steps 10 --only "pass-2 grid" 2>&1 | tee gpurun_out/variant_sweep.txt
B="python bench.py --steps 2 --warmup 1 --no-cpu --recall-queries 0 --threads 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_" -c 700 --csv \
    --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
bash tools/sanitize_call.sh
