TSB_SPEC_OVERLAP=1 ncu --set full --clock-control none --import-source on -k regex:spectral -s 8 -c 2 -o gpurun_out/ncu_r2y_r16 -f python tools/cfg5_bench.py 65536 64 8 nobuf > /dev/null 2>&1
ls -la gpurun_out | tail -3
