#!/usr/bin/env bash
# Quick confirmation run: GPU tests + the three single-GPU workloads (no CPU baseline).  Usage: tools/run_gpu_check.sh TAG
set -uo pipefail
TAG="${1:-chk}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_hifigan.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload bigvgan_b32 > gpurun_out/${TAG}_bigvgan.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload vocos_huge_b128 > gpurun_out/${TAG}_vocos.log 2>&1
python - "$TAG" <<'PY'
import json,glob,sys
for f in sorted(glob.glob(f"gpurun_out/{sys.argv[1]}_*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, round(d["ms_per_step"],3), "ms", round(d["value"]/1e6,1), "Ms/s e2e", round(d["e2e"]["value"]/1e6,1))
        for k,v in (r.get("families") or {}).items():
            print("   ", k, v["launches"], round(v["ms_per_step"],3), "ms", round(v["tflops"],1), "TF/s", round(v["gbs"],1), "GB/s roof", round(v["roofline_frac"],3))
    except Exception as e:
        print(f, "FAIL", open(f).read()[-600:])
PY
