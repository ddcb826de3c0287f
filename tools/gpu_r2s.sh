# cfg5: the frame pipeline (kernel A of frame f+1 beside kernel B of frame f): scratch slots x stream priority
set -x
out=gpurun_out/r2t_cfg5.txt; : > $out
for sl in 1 2 3 4; do for pr in 0 1; do
  [ $sl = 1 ] && [ $pr = 1 ] && continue
  echo "== slots $sl prio $pr" >> $out; TSB_SPEC_OVERLAP=$sl TSB_SPEC_PRIO=$pr python tools/cfg5_bench.py 65536 64 24 nobuf >> $out 2>&1
done; done
echo "== slots 3, 96 frames" >> $out; TSB_SPEC_OVERLAP=3 python tools/cfg5_bench.py 65536 64 96 nobuf >> $out 2>&1
echo "== slots 2, 96 frames" >> $out; TSB_SPEC_OVERLAP=2 python tools/cfg5_bench.py 65536 64 96 nobuf >> $out 2>&1
cat $out
TSB_SPEC_OVERLAP=3 python -m pytest tests/test_gpu_ops.py tests/test_expr.py -m gpu -x -q -k "spectral or expr or hack" > gpurun_out/r2t_pytest_overlap3.log 2>&1; tail -3 gpurun_out/r2t_pytest_overlap3.log
python -m pytest tests/test_gpu_ops.py tests/test_expr.py tests/test_facade.py -m gpu -x -q > gpurun_out/r2t_pytest_default.log 2>&1; tail -3 gpurun_out/r2t_pytest_default.log
