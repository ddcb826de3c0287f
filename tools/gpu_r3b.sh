bash tools/gpu_spec_ab.sh r3b base a16 occb2 base > /dev/null 2>&1
cut -c1-150 gpurun_out/specab_r3b.txt
TSB200_LIB=$PWD/build/var/a16/libtsb200.so python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "spectral" 2>&1 | tail -n 2
