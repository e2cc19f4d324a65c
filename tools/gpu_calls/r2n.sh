# round 2, call N: B = 1 after the host-side fingerprint fix: host time per forward call, PDL / N splitting again
set -uo pipefail
O=gpurun_out/r2n; mkdir -p $O
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --steps 50 --warmup 5"
timeout 200 python bench.py $BA --workload hifigan_b1 > $O/bench_b1.json 2> $O/b1.err
FV_PDL=1 timeout 200 python bench.py $BA --workload hifigan_b1 > $O/bench_b1_pdl.json 2>> $O/b1.err
FV_TC_SPLITN=1 timeout 200 python bench.py $BA --workload hifigan_b1 > $O/bench_b1_splitn.json 2>> $O/b1.err
FV_PDL=1 FV_TC_SPLITN=1 timeout 200 python bench.py $BA --workload hifigan_b1 > $O/bench_b1_pdl_splitn.json 2>> $O/b1.err
timeout 200 python bench.py $BA --workload hifigan_b1 --chain-streams off > $O/bench_b1_off.json 2>> $O/b1.err
timeout 200 python bench.py $BA --workload bigvgan_b1 > $O/bench_bigvgan_b1.json 2>> $O/b1.err
FV_PDL=1 timeout 200 python bench.py $BA --workload bigvgan_b1 > $O/bench_bigvgan_b1_pdl.json 2>> $O/b1.err
timeout 300 python bench.py --extra none --no-cpu-baseline --no-sustained --no-stress-parity --steps 20 --warmup 3 --workload hifigan_b64 > $O/bench_hifigan.json 2>> $O/b1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2n/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"], 4), d.get("launches_per_step"), "host us", round(d.get("host_us_per_forward_call", -1), 1), "e2e ms", round(d["e2e"]["ms_per_step"], 4))
    except Exception as e:
        print(f, "failed", e)
PY
