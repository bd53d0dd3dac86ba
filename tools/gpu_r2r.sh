#!/bin/bash
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2r_tests.log
cat gpurun_out/r2r_tests.log
CONP_DEBUG=1 python bench.py --fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity > gpurun_out/r2r_cfg5_n1.json 2> gpurun_out/r2r_cfg5_n1.err
CONP_DEBUG=1 python bench.py --config 4 --fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity > gpurun_out/r2r_cfg4_n1.json 2> gpurun_out/r2r_cfg4_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2r_launches_cfg5.csv python bench.py --fast-setup --steps 2 --warmup 3 --no-parity > gpurun_out/r2r_ncu.log 2>&1
grep -H "k-space stage" gpurun_out/r2r_*.err | grep "rank 0"
for f in gpurun_out/r2r_cfg5_n1.json gpurun_out/r2r_cfg4_n1.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})"; done
python tools/parse_launches.py gpurun_out/r2r_launches_cfg5.csv 2>/dev/null | tail -40
