#!/bin/bash
# 4-GPU box, profiling-only setup: N = 4 with the final code
run() { # N port extra-args out
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 $3 > gpurun_out/$4.json 2> gpurun_out/$4.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/$4.json')); print('$4', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))" || tail -c 600 gpurun_out/$4.err
}
run 4 29803 "--fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity" r3b_cfg5_n4
