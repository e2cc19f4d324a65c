# round 2, call G: row-pair convs of the C = 16 Snake stage (tests, A/B bench, launch list) + a full capture of the Snake
# kernel in that stage (51 us per launch against 41 us for the same element count at C = 64)
set -uo pipefail
O=gpurun_out/r2g; mkdir -p $O
timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -q -p no:cacheprovider -s -k "row_pair or benched_shape" > $O/pytest.log 2>&1
echo "pytest exit $?" >> $O/pytest.log
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --steps 20 --warmup 3 --workload bigvgan_b32"
timeout 300 python bench.py $BA > $O/bench_pairs.json 2> $O/bench_pairs.err
timeout 300 python bench.py $BA --no-row-pairs > $O/bench_nopairs.json 2> $O/bench_nopairs.err
RX='regex:conv_tc|mrf_fused|snake_aa|dwconv_ln|conv_post|pack_input|istft_ola|act_cast|resample|noise_conv|unpack_output|conv_simt'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$RX" -c 4000 --csv --log-file $O/launches_bigvgan_b32.csv \
    python bench.py --extra none --no-cpu-baseline --no-sustained --no-stress-parity --no-graph --steps 2 --warmup 3 \
    --workload bigvgan_b32 > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:snake_aa -s 436 -c 2 -f -o $O/prof_snake_c16 \
    python bench.py --extra none --no-cpu-baseline --no-sustained --no-stress-parity --no-graph --steps 1 --warmup 3 \
    --workload bigvgan_b32 > $O/ncu_full_snake.log 2>&1
ncu -i $O/prof_snake_c16.ncu-rep --page raw --csv > $O/prof_snake_c16_raw.csv 2>/dev/null
ncu -i $O/prof_snake_c16.ncu-rep --page source --csv > $O/prof_snake_c16_source.csv 2>/dev/null
rm -f $O/prof_snake_c16.ncu-rep
tail -n 5 $O/pytest.log
python - <<'PY'
import json
for n in ("pairs", "nopairs"):
    try:
        d = json.loads(open(f"gpurun_out/r2g/bench_{n}.json").read().strip().splitlines()[-1])
        f = d["roofline"]["families"]
        print(n, d["ms_per_step"], {k: round(v["ms_per_step"], 3) for k, v in f.items()}, d.get("parity", {}).get("max_abs_err"))
    except Exception as e:
        print(n, "failed", e)
PY
