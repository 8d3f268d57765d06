#!/bin/bash
# smoke() and the GPU suite on the committed binary.
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -k 10 500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu_last.txt 2>&1; tail -3 gpurun_out/pytest_gpu_last.txt | cut -c1-200
