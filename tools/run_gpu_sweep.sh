python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu5.log
for mb in 64 32 16 8 0; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --micro-batch $mb > gpurun_out/bench_hifigan_mb$mb.log 2>&1; done
for mb in 32 0; do python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload bigvgan_b32 --micro-batch $mb > gpurun_out/bench_bigvgan_mb$mb.log 2>&1; done
cat gpurun_out/pytest_gpu5.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*_mb*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), "ms", round(d["value"]/1e6,1), "Msamples/s", "launches", d["launches_per_step"], "conv share", round(d["roofline"]["share_of_step"],3), "hbm_frac", round(d["roofline"]["hbm_frac"],3))
    except Exception as e:
        print(f, "FAIL", open(f).read()[-300:])
PY
