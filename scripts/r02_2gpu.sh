#!/bin/bash
O=gpurun_out/r02_2gpu; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -4 > $O/pytest_dist_2gpu.log; cat $O/pytest_dist_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_train_2gpu.json 2> $O/bench_train_2gpu.err; tail -5 $O/bench_train_2gpu.err; grep '^{' $O/bench_train_2gpu.json | python scripts/print_bench.py /dev/stdin; grep '^{' $O/bench_train_2gpu.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d.get('strong_4m'), indent=1))"
