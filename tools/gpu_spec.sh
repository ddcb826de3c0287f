# developer aid: spectral-synthesis tests on the regular build, then cfg5 timings of library variants
# usage: bash tools/gpu_spec.sh TAG [variant ...]   (variants: build/var/<name>/libtsb200.so)
tag=$1; shift
python -m pytest tests/test_gpu_ops.py tests/test_expr.py -m gpu -x -q -k "spectral" > gpurun_out/spec_${tag}_pytest.log 2>&1; tail -3 gpurun_out/spec_${tag}_pytest.log
out=gpurun_out/spec_$tag.txt; : > $out
for v in base "$@" base; do
  if [ $v = base ]; then lib=tinyspec-cling_b200/libtsb200.so; else lib=build/var/$v/libtsb200.so; fi
  echo "== $v" >> $out
  TSB200_LIB=$PWD/$lib python tools/cfg5_bench.py >> $out 2>&1
done
cat $out
