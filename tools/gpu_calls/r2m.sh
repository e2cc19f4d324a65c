# round 2, call M: full suite on the final library, B = 1 with N splitting on top of one stream per chain
set -uo pipefail
O=gpurun_out/r2m; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider -s > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -n 4 $O/pytest.log
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --steps 50 --warmup 5"
timeout 200 python bench.py $BA --workload hifigan_b1 > $O/bench_b1.json 2> $O/b1.err
FV_TC_SPLITN=1 timeout 200 python bench.py $BA --workload hifigan_b1 > $O/bench_b1_splitn.json 2>> $O/b1.err
FV_PDL=1 timeout 200 python bench.py $BA --workload hifigan_b1 > $O/bench_b1_pdl.json 2>> $O/b1.err
FV_TC_SPLITN=1 timeout 200 python bench.py $BA --workload bigvgan_b1 > $O/bench_bigvgan_b1_splitn.json 2>> $O/b1.err
timeout 200 python bench.py $BA --workload bigvgan_b1 > $O/bench_bigvgan_b1.json 2>> $O/b1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2m/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 4), d.get("launches_per_step"))
    except Exception as e:
        print(f, "failed", e)
PY
