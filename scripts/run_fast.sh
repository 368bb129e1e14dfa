#!/bin/bash
O=gpurun_out/r2c; mkdir -p $O
timeout 300 python scripts/abl_edge.py fast > $O/fast_timing.log 2>&1
cat $O/fast_timing.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > $O/pytest_gpu.log
cat $O/pytest_gpu.log
