#!/bin/bash
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py::test_cfg5_headline_size_matches_oracle 2>&1 | tail -6 > gpurun_out/r2o_tests.log
cat gpurun_out/r2o_tests.log
CONP_DEBUG=1 python bench.py --fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity > gpurun_out/r2o_cfg5.json 2> gpurun_out/r2o_cfg5.err
CONP_DEBUG=1 CONP_SPREAD=sweep python bench.py --workload cfg4 --fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity > gpurun_out/r2o_cfg4_sweep.json 2> gpurun_out/r2o_cfg4_sweep.err
CONP_DEBUG=1 python bench.py --workload cfg4 --fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity > gpurun_out/r2o_cfg4.json 2> gpurun_out/r2o_cfg4.err
grep -H "k-space stage" gpurun_out/r2o_*.err
for f in gpurun_out/r2o_cfg5.json gpurun_out/r2o_cfg4_sweep.json gpurun_out/r2o_cfg4.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})"; done
