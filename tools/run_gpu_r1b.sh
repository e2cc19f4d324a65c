#!/usr/bin/env bash
# Session-3 evidence run: GPU tests, benches (fused / layer-wise MRF), ncu launch lists + full capture of fv_mrf_fused.
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_hifigan_b64.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-fuse-mrf --no-cpu-baseline > gpurun_out/bench_hifigan_b64_nofuse.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --workload bigvgan_b32 > gpurun_out/bench_bigvgan_b32.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --workload vocos_huge_b128 > gpurun_out/bench_vocos_huge_b128.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --workload hifigan_b1 --no-cpu-baseline > gpurun_out/bench_hifigan_b1.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_hifigan_b64.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mrf_fused -s 2 -c 2 -f -o gpurun_out/prof_mrf_fused \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_mrf.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, round(d["ms_per_step"],3), "ms", round(d["value"]/1e6,1), "Ms/s e2e", round(d["e2e"]["value"]/1e6,1))
        for k,v in (r.get("families") or {}).items():
            print("   ", k, v["launches"], round(v["ms_per_step"],3), "ms", round(v["tflops"],1), "TF/s", round(v["gbs"],1), "GB/s roof", round(v["roofline_frac"],3))
    except Exception as e:
        print(f, "FAIL", open(f).read()[-400:])
PY
