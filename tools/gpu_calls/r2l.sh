# round 2, call L: one stream per chain + 256-row whole-stage tiles for short launches (B = 1 latency), template golden
set -uo pipefail
O=gpurun_out/r2l; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider -s > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -n 6 $O/pytest.log
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --steps 50 --warmup 5"
timeout 200 python bench.py $BA --workload hifigan_b1 > $O/bench_b1.json 2> $O/b1.err
FV_MRF_SMALL=0 timeout 200 python bench.py $BA --workload hifigan_b1 > $O/bench_b1_nosmall.json 2>> $O/b1.err
timeout 200 python bench.py $BA --workload hifigan_b1 --chain-streams off > $O/bench_b1_off.json 2>> $O/b1.err
timeout 200 python bench.py $BA --workload bigvgan_b1 > $O/bench_bigvgan_b1.json 2>> $O/b1.err
timeout 300 python bench.py --extra none --no-cpu-baseline --no-sustained --no-stress-parity --steps 20 --warmup 3 --workload hifigan_b64 > $O/bench_hifigan.json 2> $O/bench_hifigan.err
python tools/bench_conv.py --C 128 --L 6016 --B 64 --ks 7,11 --configs 0:0:0:0,0:0:0:2 > $O/bench_conv_c128.log 2>&1
cat $O/bench_conv_c128.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2l/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        fam = (d.get("roofline") or {}).get("families") or {}
        print(f, round(d["ms_per_step"], 4), d.get("launches_per_step"), {k: round(v["ms_per_step"], 3) for k, v in fam.items()},
              (d.get("parity") or {}).get("max_abs_err"))
    except Exception as e:
        print(f, "failed", e)
PY
