# developer aid: ncu --set full of the spectral kernels for library variants.  usage: bash tools/gpu_spec_ncu.sh TAG variant...
tag=$1; shift
for v in "$@"; do
  if [ $v = base ]; then lib=tinyspec-cling_b200/libtsb200.so; else lib=build/var/$v/libtsb200.so; fi
  TSB200_LIB=$PWD/$lib python tools/cfg5_bench.py 65536 64 8 > gpurun_out/ncu_${tag}_${v}_plain.txt 2>&1
  TSB200_LIB=$PWD/$lib ncu --set full --clock-control none --import-source on -k regex:spectral -s 8 -c 2 -o gpurun_out/ncu_${tag}_${v} -f python tools/cfg5_bench.py 65536 64 8 > /dev/null 2>&1
done
ls -la gpurun_out | tail -5
