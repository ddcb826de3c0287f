out=gpurun_out/var_carve2.txt; : > $out
for v in ip96 ip96_tw1 ip96_tw0; do
  echo "== $v carve 192" >> $out
  TSB_CARVEOUT_KB=192 TSB200_LIB=$PWD/build/var/$v/libtsb200.so python tools/quick_bench.py 1024 740 20 f64 >> $out 2>&1
done
echo "== ip96 carve 164, 4 lanes per SM" >> $out
TSB_CARVEOUT_KB=164 TSB200_LIB=$PWD/build/var/ip96/libtsb200.so python tools/quick_bench.py 1024 592 20 f64 >> $out 2>&1
echo "== ip96_tw1 carve 164, 4 lanes per SM" >> $out
TSB_CARVEOUT_KB=164 TSB200_LIB=$PWD/build/var/ip96_tw1/libtsb200.so python tools/quick_bench.py 1024 592 20 f64 >> $out 2>&1
cat $out
