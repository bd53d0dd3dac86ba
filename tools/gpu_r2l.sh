#!/bin/bash
# 1-GPU: full GPU test suite with the final code + smoke + cfg5 / cfg4 stage check
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2l_tests.log
cat gpurun_out/r2l_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
CONP_DEBUG=1 python bench.py --fast-setup --steps 200 --warmup 10 --blocks 1 --no-parity > gpurun_out/r2l_cfg5.json 2> gpurun_out/r2l_cfg5.err
CONP_DEBUG=1 CONP_SPREAD=mma python bench.py --workload cfg4 --fast-setup --steps 200 --warmup 10 --blocks 1 --no-parity > gpurun_out/r2l_cfg4_mma.json 2> gpurun_out/r2l_cfg4_mma.err
grep -H "k-space stage" gpurun_out/r2l_*.err
python -c "
import json; d=json.load(open('gpurun_out/r2l_cfg5.json')); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
