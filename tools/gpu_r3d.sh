#!/bin/bash
# 8-GPU box, profiling-only setup: N = 8 with the final code
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29808 bench.py --gpus 8 --fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity > gpurun_out/r3d_cfg5_n8.json 2> gpurun_out/r3d_cfg5_n8.err
python -c "
import json,sys; d=json.load(open('gpurun_out/r3d_cfg5_n8.json')); print('r3d_cfg5_n8', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))" || tail -c 600 gpurun_out/r3d_cfg5_n8.err
