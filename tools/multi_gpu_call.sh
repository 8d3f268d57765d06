#!/bin/bash
# Multi-GPU call (gpurun --gpus N): NCCL parity check, the doc-sharded bench (BASELINE config C: fixed 10M-doc corpus),
# the data-parallel build benchmark.  N from the argument (default 2).
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tests/gpu_sharded_check.py > gpurun_out/sharded_check_n$N.log 2>&1
grep -E "sharded check|mismatch|failed|Error" gpurun_out/sharded_check_n$N.log | head -8
timeout 1500 $TR --master-port 29512 bench.py --gpus $N > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err
tail -3 gpurun_out/bench_r02_n$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_r02_n$N.json") if l.startswith("{")][-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling", "recall_at_k", "parity", "self_parity")})
    print({k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, d["e2e"]["value"], d["clocks"], d["index_build_s"])
except Exception as e:
    print("bench output unreadable:", e)
PY
timeout 600 $TR --master-port 29513 tools/bench_build.py --tokens $((4194304 * N)) --kmeans-points $((2097152 * N)) --log2k 18 \
    2>&1 | tail -2 | tee gpurun_out/bench_build_n$N.json
