#!/bin/bash
# 2-GPU: routed position exchange
python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -25 > gpurun_out/r2i_tests.log
cat gpurun_out/r2i_tests.log
CONP_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload cfg4 --steps 200 --warmup 10 --blocks 3 > gpurun_out/r2i_cfg4_n2.json 2> gpurun_out/r2i_cfg4_n2.err
CONP_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg4 --kspace ewald --steps 50 --warmup 5 --blocks 2 > gpurun_out/r2i_cfg4_ewald_n2.json 2> gpurun_out/r2i_cfg4_ewald_n2.err
CONP_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --fast-setup --steps 200 --warmup 10 --blocks 3 --no-parity > gpurun_out/r2i_cfg5_n2_fast.json 2> gpurun_out/r2i_cfg5_n2_fast.err
CONP_ROUTE=0 CONP_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --fast-setup --steps 200 --warmup 10 --blocks 3 --no-parity > gpurun_out/r2i_cfg5_n2_fast_noroute.json 2> gpurun_out/r2i_cfg5_n2_fast_noroute.err
for f in gpurun_out/r2i_cfg4_n2.json gpurun_out/r2i_cfg4_ewald_n2.json gpurun_out/r2i_cfg5_n2_fast.json gpurun_out/r2i_cfg5_n2_fast_noroute.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity'), d['roofline']['stage_ms'])"; done
grep -h "k-space stage" gpurun_out/r2i_*.err | head
tail -c 400 gpurun_out/r2i_cfg4_n2.err
