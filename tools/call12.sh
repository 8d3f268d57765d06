#!/bin/bash
# GPU call 12, one box: GPU tests, the default bench line, config E probe sweep on one shard, nq = 48, launch list.
set -u
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -q -m gpu -rf -x > gpurun_out/pytest_gpu.txt 2>&1; tail -15 gpurun_out/pytest_gpu.txt | cut -c1-220
/usr/bin/time -v timeout -k 10 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
grep -E "Elapsed|Maximum resident" gpurun_out/bench_default.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_default.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "recall_at_k", "recall_queries", "parity", "self_parity")})
    print({k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
    print(d["e2e"]["value"], d["concurrent"]["value"], d["clocks"], d["cpu_baseline"], d["roofline"])
except Exception as e:
    print("bench output unreadable:", e)
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json
timeout 900 python tools/recall_sweep.py --queries 64 --docs-total 6250000 --doclen 220 --nbits 2 --top-k 1000 \
    --n-full-scores 8192 --batch 32 --probes 1,2,4,8,16,32,64 > gpurun_out/config_e_shard.jsonl 2> gpurun_out/config_e_shard.err
cat gpurun_out/config_e_shard.jsonl | cut -c1-400; tail -3 gpurun_out/config_e_shard.err
timeout 300 python tools/variant_sweep.py --steps 10 --only "nq=48" 2>&1 | tee gpurun_out/variant_sweep.txt
B="python bench.py --steps 2 --warmup 1 --no-cpu --recall-queries 0 --threads 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_" -c 700 --csv \
    --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
tail -c 200 gpurun_out/ncu_launches.log
