#!/bin/bash
mkdir -p gpurun_out
for c in 2 3; do
  timeout 900 python tests/gpu_tools/run_config.py --config $c > gpurun_out/config$c.json 2> gpurun_out/config$c.err
  echo "config $c exit $?"; cat gpurun_out/config$c.json; tail -3 gpurun_out/config$c.err
done
timeout 600 python tests/gpu_tools/run_config.py --config 5 --scale 0.1 > gpurun_out/config5_small.json 2> gpurun_out/config5_small.err
echo "config 5 (0.1) exit $?"; cat gpurun_out/config5_small.json; tail -3 gpurun_out/config5_small.err
timeout 600 python tests/gpu_tools/run_config.py --config 4 --scale 0.1 > gpurun_out/config4_small.json 2> gpurun_out/config4_small.err
echo "config 4 (0.1) exit $?"; cat gpurun_out/config4_small.json; tail -3 gpurun_out/config4_small.err
