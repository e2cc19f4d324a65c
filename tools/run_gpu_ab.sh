#!/usr/bin/env bash
# A/B of an environment switch over the four single-GPU workloads: tools/run_gpu_ab.sh VAR "0 1" TAG
set -uo pipefail
VAR="$1"; VALS="$2"; TAG="${3:-ab}"
mkdir -p gpurun_out
for v in $VALS; do
  for w in hifigan_b64 hifigan_b1 bigvgan_b32 vocos_huge_b128; do
    env "$VAR=$v" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/${TAG}_${w}_$v.log 2>&1
  done
done
python - "$TAG" <<'PY'
import json,glob,sys
for f in sorted(glob.glob(f"gpurun_out/{sys.argv[1]}_*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), "ms", round(d["value"]/1e6,1), "Ms/s e2e", round(d["e2e"]["value"]/1e6,1), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "FAIL", open(f).read()[-600:])
PY
