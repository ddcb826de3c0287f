# developer aid: can kernel B of frame f really run beside kernel A of frame f+1?  (B on a high-priority stream,
# one or two persistent CTAs per SM, compiled for 2 or 3 CTAs per SM)
out=gpurun_out/r2w_cfg5.txt; : > $out
for v in base occb3; do
  if [ $v = base ]; then lib=tinyspec-cling_b200/libtsb200.so; else lib=build/var/$v/libtsb200.so; fi
  for pr in 0 2; do for bp in 0 1 2; do
    echo "== $v prio $pr bpersm $bp" >> $out
    if [ $bp = 0 ]; then TSB200_LIB=$PWD/$lib TSB_SPEC_PRIO=$pr python tools/cfg5_bench.py 65536 64 24 nobuf >> $out 2>&1
    else TSB200_LIB=$PWD/$lib TSB_SPEC_PRIO=$pr TSB_SPEC_BPERSM=$bp python tools/cfg5_bench.py 65536 64 24 nobuf >> $out 2>&1; fi
  done; done
done
TSB_SPEC_PRIO=2 TSB_SPEC_BPERSM=1 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "spectral" > gpurun_out/r2w_pytest.log 2>&1
cut -c1-130 $out; tail -n 2 gpurun_out/r2w_pytest.log
