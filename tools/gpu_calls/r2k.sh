# round 2, call K: dwconv+LN bulk kernel, second version (channel-pair threads, packed math): tests, A/B, capture
set -uo pipefail
O=gpurun_out/r2k; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider -s > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -n 4 $O/pytest.log
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity"
FV_DWLN_PIPE=1 timeout 300 python bench.py $BA --steps 20 --warmup 3 --workload vocos_huge_b128 > $O/bench_vocos_pipe.json 2> $O/bench_vocos_pipe.err
FV_DWLN_PIPE=0 timeout 300 python bench.py $BA --steps 20 --warmup 3 --workload vocos_huge_b128 > $O/bench_vocos_vec.json 2> $O/bench_vocos_vec.err
timeout 300 python bench.py $BA --steps 20 --warmup 3 --workload firefly_b32 > $O/bench_firefly.json 2> $O/bench_firefly.err
FV_DWLN_PIPE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_ln -s 190 -c 2 -f -o $O/prof_dwln \
    python bench.py $BA --no-graph --steps 1 --warmup 3 --workload vocos_huge_b128 > $O/ncu_dwln.log 2>&1
ncu -i $O/prof_dwln.ncu-rep --page raw --csv > $O/prof_dwln_raw.csv 2>/dev/null
ncu -i $O/prof_dwln.ncu-rep --page source --csv > $O/prof_dwln_source.csv 2>/dev/null
rm -f $O/prof_dwln.ncu-rep
timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "test_cuda_core_kernels or (test_generator_matches_reference_golden and vocos_small_stress)" -p no:cacheprovider > $O/racecheck_dwln.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" $O/racecheck_dwln.log | tail -3
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2k/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        fam = (d.get("roofline") or {}).get("families") or {}
        print(f, round(d["ms_per_step"], 4), d.get("launches_per_step"), {k: round(v["ms_per_step"], 3) for k, v in fam.items()})
    except Exception as e:
        print(f, "failed", e)
PY
