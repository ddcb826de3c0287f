# full GPU evidence pass: the gpu test suite, smoke, the contract bench, cfg5 timing.  usage: bash tools/gpu_full.sh TAG
tag=$1
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -3 gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.json
python tools/cfg5_bench.py > gpurun_out/${tag}_cfg5.txt 2>&1; cat gpurun_out/${tag}_cfg5.txt
