# round 2, call B: verify the fixes, hardware probes (UMMA modes, fp32 FMA rate), C=64 pair configuration, B=1 split-N
O=gpurun_out/r2b; mkdir -p $O
timeout 300 python tools/umma_rate.py > $O/umma_rate.md 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider -s > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
timeout 200 python tools/bench_mrf.py --acts tanh --shapes 64x12032x64 > $O/bench_mrf.log 2>&1
timeout 200 python tools/bench_mrf.py pairs64 > $O/bench_mrf_pairs64.log 2>&1
FV_MRF_C64_PAIR=0 timeout 200 python tools/bench_mrf.py pairs64 > $O/bench_mrf_pairs64_1cta.log 2>&1
BA="--extra none --no-cpu-baseline --no-sustained --no-stress-parity"
timeout 200 python bench.py $BA > $O/bench_hifigan.json 2>> $O/bench.err
timeout 200 python bench.py $BA --pairwise-c64 > $O/bench_hifigan_pairs64.json 2>> $O/bench.err
timeout 200 python bench.py $BA --workload hifigan_b1 > $O/bench_b1.json 2>> $O/bench.err
FV_TC_SPLITN=0 timeout 200 python bench.py $BA --workload hifigan_b1 > $O/bench_b1_nosplit.json 2>> $O/bench.err
FV_PDL=1 timeout 200 python bench.py $BA --workload hifigan_b1 > $O/bench_b1_pdl.json 2>> $O/bench.err
tail -3 $O/pytest.log
