#!/bin/bash
# 2-GPU box: pair math + trace facility; graph timeline at N = 2 and N = 1
python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4 > gpurun_out/r2u_tests.log
cat gpurun_out/r2u_tests.log
run() { # N port extra-args out
  if [ $1 -eq 1 ]; then python bench.py $3 > gpurun_out/$4.json 2> gpurun_out/$4.err; else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 $3 > gpurun_out/$4.json 2> gpurun_out/$4.err; fi
  python -c "
import json,sys; d=json.load(open('gpurun_out/$4.json')); print('$4', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})" || tail -c 600 gpurun_out/$4.err
  grep -h "graph timeline" gpurun_out/$4.err
}
CONP_DEBUG=1 run 1 0 "--fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity" r2u_cfg5_n1
CONP_TRACE=1 run 1 0 "--fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity" r2u_cfg5_n1_trace
CONP_DEBUG=1 run 2 29701 "--fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity" r2u_cfg5_n2
CONP_TRACE=1 run 2 29702 "--fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity" r2u_cfg5_n2_trace
