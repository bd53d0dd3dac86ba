#!/bin/bash
# 8-GPU box, profiling-only setup (random S): stage splits at N = 8 and N = 4
run() { # N port extra-args out
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 $3 > gpurun_out/$4.json 2> gpurun_out/$4.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/$4.json')); print('$4', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})" || tail -c 600 gpurun_out/$4.err
}
CONP_DEBUG=1 run 8 29604 "--fast-setup --steps 300 --warmup 20 --blocks 3 --no-parity" r2t_cfg5_n8_fast
CONP_DEBUG=1 run 4 29605 "--fast-setup --steps 300 --warmup 20 --blocks 3 --no-parity" r2t_cfg5_n4_fast
CONP_ROUTE=0 CONP_DEBUG=1 run 8 29606 "--fast-setup --steps 300 --warmup 20 --blocks 3 --no-parity" r2t_cfg5_n8_fast_noroute
grep -h "k-space stage" gpurun_out/r2t_*.err | grep "rank 0" | head
