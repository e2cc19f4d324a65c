#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "passed|failed|error|Error|assert|small_stress:|small_ref:|full-width" | head -30
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/g_hifigan.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload vocos_huge_b128 > gpurun_out/g_vocos.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload bigvgan_b32 > gpurun_out/g_bigvgan.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:snake_aa -s 40 -c 1 -f -o gpurun_out/prof_snake_r1g \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --workload bigvgan_b32 > gpurun_out/ncu_snake_g.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/g_*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, round(d["ms_per_step"],3), "ms", round(d["value"]/1e6,1), "Ms/s e2e", round(d["e2e"]["value"]/1e6,1))
        for k,v in (r.get("families") or {}).items():
            print("   ", k, v["launches"], round(v["ms_per_step"],3), "ms", round(v["tflops"],1), "TF/s", round(v["gbs"],1), "GB/s roof", round(v["roofline_frac"],3))
    except Exception as e:
        print(f, "FAIL", open(f).read()[-600:])
PY
