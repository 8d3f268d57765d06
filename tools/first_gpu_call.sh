#!/bin/bash
# First GPU call of a round, one box, ~15 minutes: the GPU test-suite, smoke, a short default bench and the variant
# sweep (tools/variant_sweep.py, self-parity checked per variant).  Everything lands in gpurun_out/.
#   tools/gpurun_retry.sh 1100 gpurun_out/call1.txt 'bash tools/first_gpu_call.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
free -g | head -2; nproc
timeout 600 python -m pytest tests -q -m gpu -rf --durations=8 --deselect tests/test_gpu_create_index.py::test_baseline_config_a > gpurun_out/pytest_gpu.txt 2>&1; tail -45 gpurun_out/pytest_gpu.txt | cut -c1-220
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 --recall-queries 64 --parity-queries 16 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_n1.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "wall_ms_per_step", "recall_at_k", "parity", "self_parity")})
    print({k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
    print({k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()}, d["e2e"]["value"], d["clocks"])
    print(d["work_per_step"])
except Exception as e:
    print("bench output unreadable:", e)
PY
timeout 420 python tools/variant_sweep.py --steps 10 --only "exact fp32,code-diff,E=2,FFMA2 in k_exact,cg,FILTER_V1,nq=48,signature" 2>&1 | tee gpurun_out/variant_sweep.txt
# a5 design question: random-row gather rate of an L2-resident table, LSU vs TMA gather4 (tools/tma_gather_bench.cu)
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/tma_gather_bench.cu -o /tmp/tma_gather_bench 2>&1 | tail -3
for rb in 64 32; do timeout 60 /tmp/tma_gather_bench $rb 18 256 1; done 2>&1 | tee gpurun_out/tma_gather_bench.txt
timeout 60 /tmp/tma_gather_bench 64 18 256 4 2>&1 | tail -1 | tee -a gpurun_out/tma_gather_bench.txt
