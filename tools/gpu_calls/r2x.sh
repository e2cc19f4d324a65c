# round 2: alternating tile direction (producer -> consumer L2 reuse): suite + A/B
O=gpurun_out/r2x; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -n 3 $O/pytest.log
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --steps 20 --warmup 3"
for v in 1 0; do
  for wl in hifigan_b64 bigvgan_b32; do
    FV_SERPENTINE=$v timeout 120 python bench.py $BA --workload $wl > $O/bench_${wl}_serp$v.json 2>> $O/err.log
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2x/bench_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"], 4), {k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["families"].items()})
PY
