# round 2: polyphase tile order (phase fastest): full suite + A/B-less bench of the three main workloads
O=gpurun_out/r2w; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -n 4 $O/pytest.log
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --steps 20 --warmup 3"
for wl in hifigan_b64 bigvgan_b32 hifigan_b64 bigvgan_b32; do
  timeout 200 python bench.py $BA --workload $wl > $O/bench_${wl}_$RANDOM.json 2>> $O/err.log
done
RX='regex:conv_tc|mrf_fused|snake_aa|conv_post|pack_input'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$RX" -c 1000 --csv --log-file $O/launches_hifigan_b64.csv \
    python bench.py $BA --no-graph --steps 2 --workload hifigan_b64 > $O/ncu.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2w/bench_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"], 4), {k: round(v["ms_per_step"], 3) for k, v in d["roofline"]["families"].items()}, d["parity"]["max_abs_err"])
PY
