# final pass of the round on HEAD: GPU suite, smoke, contract bench (both arms).  usage: bash tools/gpu_final.sh TAG
T=${1:-final}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 200 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
python tools/cfg5_bench.py > gpurun_out/${T}_cfg5.txt 2>&1; cat gpurun_out/${T}_cfg5.txt
