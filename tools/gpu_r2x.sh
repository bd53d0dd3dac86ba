#!/bin/bash
# 8-GPU box, profiling-only setup: graph timelines at N = 8 and N = 4
run() { # N port extra-args out
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 $3 > gpurun_out/$4.json 2> gpurun_out/$4.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/$4.json')); print('$4', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))" || tail -c 600 gpurun_out/$4.err
  grep -h "graph timeline" gpurun_out/$4.err | sort
}
A5="--fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity"
run 8 29801 "$A5" r2x_cfg5_n8
CONP_TRACE=1 run 8 29802 "$A5" r2x_cfg5_n8_trace
CONP_TRACE=1 run 4 29803 "$A5" r2x_cfg5_n4_trace
