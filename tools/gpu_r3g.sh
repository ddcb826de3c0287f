for pr in 1 2; do for i in 1 2 3 4; do echo -n "prio $pr: "; TSB_SPEC_PRIO=$pr python tools/cfg5_bench.py 65536 64 24 nobuf | cut -c1-100; done; done
for i in 1 2 3; do echo -n "slots 3: "; TSB_SPEC_OVERLAP=3 python tools/cfg5_bench.py 65536 64 24 nobuf | cut -c1-100; done
