# round 2, call J: conv_post (batched loads, [hi | lo] rows) and dwconv+LN (hoisted loads vs bulk-copy pipeline): tests, A/B, captures
set -uo pipefail
O=gpurun_out/r2j; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider -s > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -n 4 $O/pytest.log
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity"
FV_DWLN_BULK=1 timeout 300 python bench.py $BA --steps 20 --warmup 3 --workload vocos_huge_b128 > $O/bench_vocos_bulk.json 2> $O/bench_vocos_bulk.err
FV_DWLN_BULK=0 timeout 300 python bench.py $BA --steps 20 --warmup 3 --workload vocos_huge_b128 > $O/bench_vocos_vec.json 2> $O/bench_vocos_vec.err
timeout 300 python bench.py $BA --steps 20 --warmup 3 --workload hifigan_b64 > $O/bench_hifigan.json 2> $O/bench_hifigan.err
timeout 300 python bench.py $BA --steps 20 --warmup 3 --workload bigvgan_b32 > $O/bench_bigvgan.json 2> $O/bench_bigvgan.err
for v in 1 0; do
  FV_DWLN_BULK=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_ln -s 190 -c 2 -f -o $O/prof_dwln_$v \
      python bench.py $BA --no-graph --steps 1 --warmup 3 --workload vocos_huge_b128 > $O/ncu_dwln_$v.log 2>&1
  ncu -i $O/prof_dwln_$v.ncu-rep --page raw --csv > $O/prof_dwln_${v}_raw.csv 2>/dev/null
  ncu -i $O/prof_dwln_$v.ncu-rep --page source --csv > $O/prof_dwln_${v}_source.csv 2>/dev/null
  rm -f $O/prof_dwln_$v.ncu-rep
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2j/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        fam = (d.get("roofline") or {}).get("families") or {}
        print(f, round(d["ms_per_step"], 4), d.get("launches_per_step"), {k: round(v["ms_per_step"], 3) for k, v in fam.items()})
    except Exception as e:
        print(f, "failed", e)
PY
