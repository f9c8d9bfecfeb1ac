#!/bin/bash
# What the driver runs at round end on one GPU: whole suite, smoke(), both bench arms with default flags.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/final_smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
echo "reference arm exit $?"; tail -1 gpurun_out/final_bench_ref.json | cut -c1-300
timeout 1200 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
echo "bench exit $?"; tail -1 gpurun_out/final_bench_n1.json | cut -c1-400
