# round 2: the driver's N > 1 launch line on four GPUs (headline cfg B + configs[4] BigVGAN 32 / GPU)
O=gpurun_out/r2s4; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29527 \
    bench.py --gpus 4 --steps 10 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err
tail -c 400 $O/bench_n4.json; tail -n 3 $O/bench_n4.err
