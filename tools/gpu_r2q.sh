#!/bin/bash
python -m pytest tests/test_multi_gpu.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -6 > gpurun_out/r2q_tests.log
cat gpurun_out/r2q_tests.log
CONP_DEBUG=1 python bench.py --fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity > gpurun_out/r2q_cfg5_n1.json 2> gpurun_out/r2q_cfg5_n1.err
CONP_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity > gpurun_out/r2q_cfg5_n2.json 2> gpurun_out/r2q_cfg5_n2.err
CONP_SPREAD=atomic CONP_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity > gpurun_out/r2q_cfg5_n2_atomic.json 2> gpurun_out/r2q_cfg5_n2_atomic.err
grep -H "k-space stage" gpurun_out/r2q_*.err | grep "rank 0"
for f in gpurun_out/r2q_cfg5_n1.json gpurun_out/r2q_cfg5_n2.json gpurun_out/r2q_cfg5_n2_atomic.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})"; done
