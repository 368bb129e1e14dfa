#!/bin/bash
# 2-GPU validation of the final library (overlapped launches next to NCCL kernels): dist tests + the driver-style bench line
O=gpurun_out/${OUT:-r02_2gpu_final}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_launch_overlap.py -q -m gpu 2>&1 | tail -4 > $O/pytest_dist_2gpu.log; cat $O/pytest_dist_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; tail -2 $O/bench_2gpu.err; python scripts/print_bench.py $O/bench_2gpu.json | head -1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/'+__import__("os").environ.get("OUT","r02_2gpu_final")+'/bench_2gpu.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('strong_4m'))[:600])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench_ref_2gpu.json 2> $O/bench_ref_2gpu.err; cut -c1-200 $O/bench_ref_2gpu.json
