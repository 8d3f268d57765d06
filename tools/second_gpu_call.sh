#!/bin/bash
# Second GPU call, one box, ~10 minutes: GPU test-suite (all of it), the default bench line, nq = 48, exact-a2 twin, then
# the launch list of two steps and full ncu captures of the five main kernels.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu -rf --durations=6 > gpurun_out/pytest_gpu.txt 2>&1; tail -25 gpurun_out/pytest_gpu.txt | cut -c1-200
timeout 500 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err
tail -2 gpurun_out/bench_r02_n1.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r02_n1.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "recall_at_k", "parity", "self_parity", "cpu_baseline")})
    print({k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
    print(d["e2e"]["value"], d["concurrent"]["value"], d["clocks"])
except Exception as e:
    print("bench output unreadable:", e)
PY
timeout 300 python tools/variant_sweep.py --steps 10 --only "exact fp32,FILTER_V1,nq=48" 2>&1 | tee gpurun_out/variant_sweep.txt
B="python bench.py --steps 2 --warmup 1 --no-cpu --recall-queries 0 --threads 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_" -c 600 --csv \
    --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
tail -c 300 gpurun_out/ncu_launches.log
timeout 400 ncu --set full --import-source on --clock-control none \
    -k "regex:^k_(scores16_tc|approx16|approx_recheck|exact_tc2|exact)$" -s 5 -c 5 -o gpurun_out/r02_top5 \
    $B > gpurun_out/ncu_top5.log 2>&1
tail -c 300 gpurun_out/ncu_top5.log
