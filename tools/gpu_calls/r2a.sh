# round 2, call A: full GPU test-suite, hardware probes, precision report, first bench of the reworked path
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 120 python tools/umma_rate.py > $O/umma_rate.md 2>&1
timeout 1000 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider -s > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
timeout 200 python tools/bench_mrf.py > $O/bench_mrf.log 2>&1
FV_MRF_WS=1 timeout 200 python tools/bench_mrf.py --shapes 64x12032x64 > $O/bench_mrf_ws.log 2>&1
timeout 200 python tools/bench_mrf.py pairs > $O/bench_mrf_pairs.log 2>&1
timeout 400 python tests/diag/precision_report.py > $O/precision_report.md 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 200 python bench.py --extra none --no-cpu-baseline --no-sustained --no-fuse-pairs > $O/bench_nopairs.json 2>> $O/bench.err
timeout 200 python bench.py --extra none --no-cpu-baseline --no-sustained --mrf-silu-h2 > $O/bench_h2.json 2>> $O/bench.err
timeout 300 python bench.py --workload bigvgan_b32 --extra none --precision strict --steps 10 --no-sustained > $O/bench_bigvgan_strict.json 2>> $O/bench.err
timeout 300 python bench.py --workload bigvgan_b32 --extra none --precision fp16 --steps 10 --no-sustained > $O/bench_bigvgan_fp16.json 2>> $O/bench.err
tail -3 $O/pytest.log
