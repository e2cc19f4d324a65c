# round 2: default bench + reference arm + smoke on the final code (host timing moved behind the timed regions)
set -uo pipefail
O=gpurun_out/r2y; mkdir -p $O
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference.json 2>> $O/bench.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
tail -n 2 $O/smoke.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2y/bench.json").read().strip().splitlines()[-1])
print("headline", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "host", d.get("host_us_per_forward_call"), d["clocks"])
for k, w in d["workloads"].items():
    print(k, w["ms_per_step"], w["value"], w.get("host_us_per_forward_call"))
r = json.loads(open("gpurun_out/r2y/bench_reference.json").read().strip().splitlines()[-1])
print("reference", r["ms_per_step"], r["value"], "ratio", d["e2e"]["value"] / r["value"])
PY
