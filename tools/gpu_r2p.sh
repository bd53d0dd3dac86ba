#!/bin/bash
python -m pytest tests/test_gpu_edge_cases.py -x -q -k "spread_kernels" 2>&1 | tail -3
CONP_DEBUG=1 python bench.py --fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity > gpurun_out/r2p_cfg5.json 2> gpurun_out/r2p_cfg5.err
grep -H "k-space stage" gpurun_out/r2p_*.err
python -c "
import json; d=json.load(open('gpurun_out/r2p_cfg5.json')); print(round(d['value'],1), round(d['ms_per_step'],4), d['timing']['ms_per_step_blocks'], round(d['e2e']['value'],1), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2p_launches_cfg5.csv python bench.py --fast-setup --steps 3 --warmup 3 --blocks 1 --no-parity > /dev/null 2>&1
python tools/parse_launches.py gpurun_out/r2p_launches_cfg5.csv | head -12
