# round 2: GPU suite only
O=gpurun_out/r2t; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -n 15 $O/pytest.log
