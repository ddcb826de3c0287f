# cfg5: spectrum-from-HBM timing and the frame pipeline (kernel A of frame f+1 beside kernel B of frame f)
set -x
out=gpurun_out/r2s_cfg5.txt; : > $out
echo "== base" >> $out; python tools/cfg5_bench.py >> $out 2>&1
echo "== overlap" >> $out; TSB_SPEC_OVERLAP=1 python tools/cfg5_bench.py >> $out 2>&1
echo "== overlap, B 1 CTA/SM" >> $out; TSB_SPEC_OVERLAP=1 TSB_SPEC_BPERSM=1 python tools/cfg5_bench.py >> $out 2>&1
echo "== base, B 1 CTA/SM" >> $out; TSB_SPEC_BPERSM=1 python tools/cfg5_bench.py 65536 64 24 nobuf >> $out 2>&1
cat $out
TSB_SPEC_OVERLAP=1 python -m pytest tests/test_gpu_ops.py tests/test_expr.py -m gpu -x -q -k "spectral or expr or hack" > gpurun_out/r2s_pytest_overlap.log 2>&1; tail -3 gpurun_out/r2s_pytest_overlap.log
