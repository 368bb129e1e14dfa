#!/bin/bash
O=gpurun_out/r02_2gpu_b; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -4 > $O/pytest_dist_2gpu.log; cat $O/pytest_dist_2gpu.log
