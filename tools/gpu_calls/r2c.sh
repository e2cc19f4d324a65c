# round 2, call C: snake shared-memory-ring variant A/B, sanitizer runs
O=gpurun_out/r2c; mkdir -p $O
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity --workload bigvgan_b32"
FV_SNAKE_RING=1 timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "cuda_core or bigvgan or snake" > $O/pytest_ring.log 2>&1; echo "exit $?" >> $O/pytest_ring.log
timeout 200 python bench.py $BA > $O/bench_bigvgan.json 2>> $O/bench.err
FV_SNAKE_RING=1 timeout 200 python bench.py $BA > $O/bench_bigvgan_ring.json 2>> $O/bench.err
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "five_stage or cli_end" > $O/pytest_fix.log 2>&1; echo "exit $?" >> $O/pytest_fix.log
bash tools/run_sanitizer.sh $O/sanitizer > $O/sanitizer_run.log 2>&1
tail -3 $O/pytest_ring.log $O/pytest_fix.log $O/sanitizer_run.log
