#!/bin/bash
# last seconds of the round's GPU budget: the NORDIF-to-device test and a refinement sanity check of the final build
timeout 70 python -m pytest tests/test_io_nordif.py tests/test_refinement.py -q -m gpu -x -k "device or edge" 2>&1 | tail -5
