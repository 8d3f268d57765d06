#!/bin/bash
# A/B of one switch with the whole GPU suite under it: PB_K1_EARLY (k_scores16_tc epilogue).
set -u
mkdir -p gpurun_out
PB_K1_EARLY=1 timeout -k 10 600 python -m pytest tests -q -m gpu -rf -x > gpurun_out/pytest_gpu_early.txt 2>&1; tail -5 gpurun_out/pytest_gpu_early.txt | cut -c1-220
timeout 400 python tools/variant_sweep.py --steps 10 --only "accumulator release" 2>&1 | tee gpurun_out/variant_sweep.txt
