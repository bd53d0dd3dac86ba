#!/bin/bash
# 2-GPU validation: multi-GPU parity tests, all GPU tests, bench (cfg4 pppm + ewald) at N=2 with parity in the line
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py::test_cfg5_headline_size_matches_oracle 2>&1 | tail -25 > gpurun_out/r2h_tests.log
cat gpurun_out/r2h_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload cfg4 --steps 200 --warmup 10 --blocks 3 > gpurun_out/r2h_cfg4_n2.json 2> gpurun_out/r2h_cfg4_n2.err
tail -c 300 gpurun_out/r2h_cfg4_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg4 --kspace ewald --steps 50 --warmup 5 --blocks 2 > gpurun_out/r2h_cfg4_ewald_n2.json 2> gpurun_out/r2h_cfg4_ewald_n2.err
tail -c 300 gpurun_out/r2h_cfg4_ewald_n2.err
python bench.py --workload cfg4 --steps 200 --warmup 10 --blocks 3 --no-cpu-baseline > gpurun_out/r2h_cfg4_n1.json 2> gpurun_out/r2h_cfg4_n1.err
for f in gpurun_out/r2h_cfg4_n2.json gpurun_out/r2h_cfg4_ewald_n2.json gpurun_out/r2h_cfg4_n1.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity'), d['roofline']['stage_ms'])"; done
