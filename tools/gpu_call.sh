#!/bin/bash
# One-GPU call (gpurun): GPU tests, the default bench line (timed), reference arm, config E probe sweep on one shard,
# build benchmark on one GPU, launch list.
set -u
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -q -m gpu -rf -x > gpurun_out/pytest_gpu.txt 2>&1; tail -15 gpurun_out/pytest_gpu.txt | cut -c1-220
T0=$(date +%s)
timeout -k 10 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "default bench wall seconds: $(( $(date +%s) - T0 ))"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_default.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "steps", "recall_at_k", "recall_queries", "parity", "self_parity")})
    print({k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
    print(d["e2e"]["value"], d["concurrent"]["value"], d["clocks"], d["cpu_baseline"], d["roofline"])
except Exception as e:
    print("bench output unreadable:", e)
PY
timeout 900 python tools/recall_sweep.py --queries 64 --docs-total 6250000 --doclen 220 --nbits 2 --top-k 1000 \
    --n-full-scores 8192 --batch 32 --threshold -1 --probes 1,2,4,8,16,32,64 > gpurun_out/config_e_shard.jsonl 2> gpurun_out/config_e_shard.err
cat gpurun_out/config_e_shard.jsonl | cut -c1-420; tail -3 gpurun_out/config_e_shard.err
timeout 600 python tools/bench_build.py --tokens 4194304 --kmeans-points 2097152 --log2k 18 2>&1 | tail -2 | tee gpurun_out/bench_build_n1.json
B="python bench.py --steps 2 --warmup 1 --no-cpu --recall-queries 0 --threads 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_" -c 700 --csv \
    --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
tail -c 200 gpurun_out/ncu_launches.log
