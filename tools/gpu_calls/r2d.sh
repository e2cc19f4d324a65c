# round 2, call D: full GPU suite, sanitizer after the explicit staging barriers, full bench + reference arm
O=gpurun_out/r2d; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider -s > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
bash tools/run_sanitizer.sh $O/sanitizer > $O/sanitizer_run.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference.json 2>> $O/bench.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
tail -n 3 $O/pytest.log $O/sanitizer_run.log $O/smoke.log
