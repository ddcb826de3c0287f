# 8-GPU evidence of the time-sharded cfg3 render on the final kernels (one-hour and ten-minute renders)
for s in 3600 600; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/cfg3_timeshard.py --seconds $s > gpurun_out/r2h_cfg3_timeshard_8gpu_${s}s.json 2> gpurun_out/r2h_cfg3_8gpu_${s}.err
tail -c 700 gpurun_out/r2h_cfg3_timeshard_8gpu_${s}s.json; echo
done
