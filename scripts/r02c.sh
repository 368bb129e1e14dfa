#!/bin/bash
O=gpurun_out/r02c; mkdir -p $O
./scripts/ubench/x2_rate > $O/x2_rate.log 2>&1; cat $O/x2_rate.log
timeout 300 python scripts/abl_edge.py base > $O/fast_timing.log 2>&1; cat $O/fast_timing.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | tail -5 > $O/pytest.log; cat $O/pytest.log
