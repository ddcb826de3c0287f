# round-2 evidence, second batch: full GPU suite + smoke + bench on HEAD, ncu captures of the final cfg5 / cfg4 kernels
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; tail -3 gpurun_out/r2m_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; tail -3 gpurun_out/r2m_smoke.log
python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; tail -c 400 gpurun_out/r2m_bench.json
python tools/cfg5_bench.py > gpurun_out/r2m_cfg5.txt 2>&1; cat gpurun_out/r2m_cfg5.txt
ncu --set full --clock-control none --import-source on -k regex:spectral -s 8 -c 2 -o gpurun_out/r2m_cfg5_full python tools/cfg5_bench.py 65536 64 8 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spectral -c 64 --csv --log-file gpurun_out/r2m_cfg5_launches.csv python tools/cfg5_bench.py 65536 64 8 > /dev/null 2>&1
ls -la gpurun_out | tail -8
