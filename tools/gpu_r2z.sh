#!/bin/bash
# 2-GPU box: flags raised by the consumer kernels (no stand-alone signal kernels); N = 2 evidence line (real setup, parity)
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -4 > gpurun_out/r2z_tests.log
cat gpurun_out/r2z_tests.log
run() { # N port extra-args out
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 $3 > gpurun_out/$4.json 2> gpurun_out/$4.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/$4.json')); p=d.get('parity') or {}; print('$4', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), p.get('ok'), p.get('max_rel_dq'))" || tail -c 600 gpurun_out/$4.err
}
A5="--fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity"
run 2 29701 "$A5" r2z_cfg5_n2_fast
CONP_FUSED_SIGNAL=0 run 2 29702 "$A5" r2z_cfg5_n2_fast_signal_kernels
run 2 29703 "" r2z_bench_cfg5_n2
