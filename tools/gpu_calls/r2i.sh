# round 2, call I: bulk-copy dwconv+LN (tests, sanitizer, A/B bench on Vocos-huge), B = 1 latency switches
set -uo pipefail
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider -s > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -n 4 $O/pytest.log
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity"
FV_DWLN_BULK=1 timeout 300 python bench.py $BA --steps 20 --warmup 3 --workload vocos_huge_b128 > $O/bench_vocos_bulk.json 2> $O/bench_vocos_bulk.err
FV_DWLN_BULK=0 timeout 300 python bench.py $BA --steps 20 --warmup 3 --workload vocos_huge_b128 > $O/bench_vocos_nobulk.json 2> $O/bench_vocos_nobulk.err
B1="$BA --steps 50 --warmup 5 --workload hifigan_b1"
timeout 200 python bench.py $B1 > $O/bench_b1_base.json 2> $O/b1.err
FV_PDL=1 timeout 200 python bench.py $B1 > $O/bench_b1_pdl.json 2>> $O/b1.err
FV_TC_SPLITN=1 timeout 200 python bench.py $B1 > $O/bench_b1_splitn.json 2>> $O/b1.err
timeout 200 python bench.py $B1 --pairwise-c64 > $O/bench_b1_pairs64.json 2>> $O/b1.err
FV_PDL=1 timeout 200 python bench.py $B1 --pairwise-c64 > $O/bench_b1_pairs64_pdl.json 2>> $O/b1.err
timeout 200 python bench.py $BA --steps 50 --warmup 5 --workload bigvgan_b1 > $O/bench_bigvgan_b1.json 2>> $O/b1.err
timeout 200 python bench.py $BA --steps 50 --warmup 5 --workload bigvgan_b1 --chain-streams off > $O/bench_bigvgan_b1_off.json 2>> $O/b1.err
bash tools/run_sanitizer.sh $O/sanitizer > $O/sanitizer_run.log 2>&1
tail -n 6 $O/sanitizer_run.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2i/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        fam = (d.get("roofline") or {}).get("families") or {}
        print(f, round(d["ms_per_step"], 4), d.get("launches_per_step"), {k: round(v["ms_per_step"], 3) for k, v in fam.items()})
    except Exception as e:
        print(f, "failed", e)
PY
