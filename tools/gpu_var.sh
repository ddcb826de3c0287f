# developer aid: time chain-kernel variants (build/var/*/libtsb200.so, tools/build_variant.sh) against the regular build
# usage: bash tools/gpu_var.sh TAG PRECS LANES variant [variant ...]
tag=$1; precs=$2; lanes=$3; shift 3
out=gpurun_out/var_$tag.txt; : > $out
for v in base "$@" base; do
  if [ $v = base ]; then lib=tinyspec-cling_b200/libtsb200.so; else lib=build/var/$v/libtsb200.so; fi
  echo "== $v" >> $out
  TSB200_LIB=$PWD/$lib python tools/quick_bench.py 1024 $lanes 20 $precs >> $out 2>&1
done
cat $out
