for i in 1 2 3 4; do python tools/cfg5_bench.py 65536 64 24 nobuf | cut -c1-120; done
for i in 1 2 3; do python tools/cfg5_bench.py 65536 32 24 nobuf | cut -c1-120; done
for i in 1 2; do python tools/cfg5_bench.py 65536 16 24 nobuf | cut -c1-120; done
for i in 1 2; do TSB_SPEC_OVERLAP=1 python tools/cfg5_bench.py 65536 64 24 nobuf | cut -c1-120; done
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
