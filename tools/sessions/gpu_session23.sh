#!/bin/bash
# merge_crystal_maps on the device: parity tests + timing at a 1000x1000 map
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_merge_maps.py -q -m gpu -x 2>&1 | tail -15
