#!/bin/bash
# First GPU call of a round, one box, ~6 minutes: the GPU test-suite, the default bench, every measurement knob
# (tools/variant_sweep.py, parity checked per variant), then a launch list and full captures of the kernels that
# changed last.  Everything lands in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/first_gpu_call.sh'
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -25
timeout 200 python -m pytest tests/test_loader_cpu.py -x -q 2>&1 | tail -3
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 250 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step", "recall_at_k", "parity")},
      {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, d["e2e"]["value"], d["clocks"])
PY
timeout 600 python tools/variant_sweep.py --steps 10 --only "default,exact,E=2,FFMA2,cg" 2>&1 | tee gpurun_out/variant_sweep.txt
timeout 200 ncu --set full --import-source on --clock-control none \
    -k "regex:^k_(scores16_tc|approx_recheck|approx16)$" -s 3 -c 3 -o gpurun_out/ncu_k1 \
    python bench.py --steps 3 --warmup 2 --no-cpu --recall-queries 0 > gpurun_out/ncu_k1.log 2>&1
tail -c 300 gpurun_out/ncu_k1.log
