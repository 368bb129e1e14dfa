#!/bin/bash
O=gpurun_out/r02_8gpu; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_train_8gpu.out 2> $O/bench_train_8gpu.err; tail -5 $O/bench_train_8gpu.err; grep '^{' $O/bench_train_8gpu.out > $O/bench_train_8gpu.json; python scripts/print_bench.py $O/bench_train_8gpu.json; python -c "import json,sys; d=json.load(open('$O/bench_train_8gpu.json')); print(json.dumps(d.get('strong_4m'), indent=1))"
