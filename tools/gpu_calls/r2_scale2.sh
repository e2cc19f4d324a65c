# round 2: the driver's N > 1 launch line on two GPUs (headline cfg B + configs[4] BigVGAN 32 / GPU), plus the reference arm under torchrun
O=gpurun_out/r2s; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_n2_reference.json 2> $O/bench_n2_reference.err
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "another_device or shard" > $O/pytest_2gpu.log 2>&1
tail -c 600 $O/bench_n2.json; tail -n 3 $O/pytest_2gpu.log; tail -n 5 $O/bench_n2.err
