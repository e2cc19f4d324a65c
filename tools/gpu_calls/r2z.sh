# round 2, final evidence call: full GPU suite, sanitizer, default bench + reference arm, smoke, launch lists of the final kernels
set -uo pipefail
O=gpurun_out/r2z; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider -s > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference.json 2>> $O/bench.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
bash tools/run_sanitizer.sh $O/sanitizer > $O/sanitizer_run.log 2>&1
RX='regex:conv_tc|mrf_fused|snake_aa|dwconv_ln|conv_post|pack_input|istft_ola|act_cast|resample|noise_conv|unpack_output|conv_simt'
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --no-graph --steps 2 --warmup 3"
for wl in hifigan_b64 bigvgan_b32 vocos_huge_b128 hifigan_b1; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$RX" -c 4000 --csv --log-file $O/launches_$wl.csv \
      python bench.py $BA --workload $wl > $O/ncu_$wl.log 2>&1
done
tail -n 3 $O/pytest.log $O/sanitizer_run.log $O/smoke.log
head -c 600 $O/bench.json; echo; cat $O/bench_reference.json | head -c 400
