#!/bin/bash
O=gpurun_out/r02o; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu -x 2>&1 | tail -12 > $O/pytest.log; cat $O/pytest.log
