python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2h_bench_2gpu.json 2> gpurun_out/r2h_bench_2gpu.err
tail -c 300 gpurun_out/r2h_bench_2gpu.json
