#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "mrf" 2>&1 | tail -3
timeout 200 python tools/bench_mrf.py 2>&1 | tee gpurun_out/bench_mrf_step2.log
FV_MRF_C32_CTAS=1 timeout 200 python tools/bench_mrf.py --shapes 32x24064x64 2>&1 | tee -a gpurun_out/bench_mrf_step2.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mrf_fused -c 2 -f -o gpurun_out/prof_mrf_step2 \
    python tools/bench_mrf.py --acts silu --iters 1 > gpurun_out/ncu_mrf_step2.log 2>&1
