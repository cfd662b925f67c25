#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/train_trace.py > gpurun_out/train_trace_s3i.txt 2>&1; head -24 gpurun_out/train_trace_s3i.txt | tail -22 | cut -c1-150
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 1 --workload train --steps 3 --warmup 3 > gpurun_out/bench_train_s3i.json 2> gpurun_out/bench_train_s3i.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/bench_train_s3i.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], d['train']['phase_ms'], d['train']['gpu_launches_per_step'])"
