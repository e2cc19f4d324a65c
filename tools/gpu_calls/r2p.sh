# round 2, call P: Snake ring depth (two vs three windows in flight, two blocks per SM) + full suite on the final library
set -uo pipefail
O=gpurun_out/r2p; mkdir -p $O
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --steps 20 --warmup 3 --workload bigvgan_b32"
for v in 2 3 2 3; do
  FV_SNAKE_RING=$v timeout 300 python bench.py $BA > $O/bench_ring${v}_$RANDOM.json 2> $O/bench.err
done
FV_SNAKE_RING=3 timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "snake or bigvgan" > $O/pytest_ring3.log 2>&1; tail -n 2 $O/pytest_ring3.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/pytest.log 2>&1; tail -n 2 $O/pytest.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2p/bench_ring*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        fam = d["roofline"]["families"]
        print(f, round(d["ms_per_step"], 4), {k: round(x["ms_per_step"], 3) for k, x in fam.items()})
    except Exception as e:
        print(f, "failed", e)
PY
