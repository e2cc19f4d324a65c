# round 2, call O: Snake look-ahead depth at equal occupancy (shared-memory ring with two blocks per SM)
set -uo pipefail
O=gpurun_out/r2o; mkdir -p $O
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --steps 20 --warmup 3 --workload bigvgan_b32"
for v in 0 2 1; do
  FV_SNAKE_RING=$v timeout 300 python bench.py $BA > $O/bench_ring$v.json 2> $O/bench_ring$v.err
done
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "snake or bigvgan" > $O/pytest.log 2>&1; tail -n 2 $O/pytest.log
FV_SNAKE_RING=2 timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "snake or bigvgan" > $O/pytest_ring2.log 2>&1; tail -n 2 $O/pytest_ring2.log
python - <<'PY'
import json
for v in (0, 2, 1):
    try:
        d = json.loads(open(f"gpurun_out/r2o/bench_ring{v}.json").read().strip().splitlines()[-1])
        fam = d["roofline"]["families"]
        print("FV_SNAKE_RING", v, round(d["ms_per_step"], 4), {k: round(x["ms_per_step"], 3) for k, x in fam.items()})
    except Exception as e:
        print(v, "failed", e)
PY
