# round 2: chain streams for the C = 256 stage at full batch (row threshold 148 x 256 against 148 x 512 / 1024)
O=gpurun_out/r2v; mkdir -p $O
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --steps 20 --warmup 3"
for r in 37888 75776 37888 75776; do
  timeout 200 python bench.py $BA --workload hifigan_b64 --chain-rows $r > $O/bench_hifigan_${r}_$RANDOM.json 2>> $O/err.log
done
for r in 37888 75776 151552; do
  timeout 200 python bench.py $BA --workload bigvgan_b32 --chain-rows $r > $O/bench_bigvgan_${r}_$RANDOM.json 2>> $O/err.log
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2v/bench_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"], 4), d.get("launches_per_step"))
PY
