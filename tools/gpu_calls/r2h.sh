# round 2, call H: chain streams at B = 1 (tests + A/B bench), row-pair rule by tap count (A/B bench), full GPU suite
set -uo pipefail
O=gpurun_out/r2h; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider -s > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --steps 50 --warmup 5"
for cs in auto off; do
  timeout 300 python bench.py $BA --workload hifigan_b1 --chain-streams $cs > $O/bench_b1_$cs.json 2> $O/bench_b1_$cs.err
done
timeout 300 python bench.py --extra none --no-cpu-baseline --no-sustained --no-stress-parity --steps 20 --warmup 3 --workload bigvgan_b32 > $O/bench_bigvgan.json 2> $O/bench_bigvgan.err
timeout 300 python bench.py --extra none --no-cpu-baseline --no-sustained --no-stress-parity --steps 20 --warmup 3 --workload hifigan_b64 > $O/bench_hifigan.json 2> $O/bench_hifigan.err
tail -n 8 $O/pytest.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2h/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        fam = (d.get("roofline") or {}).get("families") or {}
        print(f, d["ms_per_step"], d.get("launches_per_step"), {k: round(v["ms_per_step"], 3) for k, v in fam.items()},
              (d.get("parity") or {}).get("max_abs_err"))
    except Exception as e:
        print(f, "failed", e)
PY
