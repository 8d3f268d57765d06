#!/bin/bash
# Profiling call, one box, ~10 minutes: launch list of two steps, full ncu captures of the five main kernels, the
# deselected long test, a default-length bench line.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu --recall-queries 0 --threads 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_" -c 600 --csv \
    --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
tail -c 400 gpurun_out/ncu_launches.log
timeout 400 ncu --set full --import-source on --clock-control none \
    -k "regex:^k_(scores16_tc|approx16|approx_recheck|exact_tc2|exact)$" -s 5 -c 5 -o gpurun_out/r02_top5 \
    $B > gpurun_out/ncu_top5.log 2>&1
tail -c 300 gpurun_out/ncu_top5.log
timeout 300 python -m pytest tests/test_gpu_create_index.py::test_baseline_config_a -q 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err
tail -2 gpurun_out/bench_r02_n1.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_r02_n1.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "recall_at_k", "parity", "self_parity", "cpu_baseline")})
    print({k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, d["e2e"]["value"], d["clocks"])
    print(json.dumps(d["roofline"])[:600])
except Exception as e:
    print("bench output unreadable:", e)
PY
