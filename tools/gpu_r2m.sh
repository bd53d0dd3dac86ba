#!/bin/bash
python -m pytest tests/test_gpu_edge_cases.py -x -q -k "spread_kernels" 2>&1 | tail -15 > gpurun_out/r2m_tests.log
cat gpurun_out/r2m_tests.log
CONP_DEBUG=1 CONP_SPREAD=sweep python bench.py --fast-setup --steps 200 --warmup 10 --blocks 1 --no-parity > gpurun_out/r2m_cfg5_sweep.json 2> gpurun_out/r2m_cfg5_sweep.err
CONP_DEBUG=1 CONP_SPREAD=sweep python bench.py --workload cfg4 --fast-setup --steps 200 --warmup 10 --blocks 1 --no-parity > gpurun_out/r2m_cfg4_sweep.json 2> gpurun_out/r2m_cfg4_sweep.err
grep -H "k-space stage\|sweep spread" gpurun_out/r2m_*.err
python -c "
import json; d=json.load(open('gpurun_out/r2m_cfg5_sweep.json')); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
CONP_SPREAD=sweep python bench.py --fast-setup --steps 3 --warmup 3 --blocks 1 --no-parity > /dev/null 2>&1 && \
CONP_SPREAD=sweep ncu --set full --import-source on --clock-control none -k regex:spread_sweep -c 1 -o gpurun_out/r2m_spread_sweep python bench.py --fast-setup --steps 3 --warmup 3 --blocks 1 --no-parity > gpurun_out/r2m_ncu.log 2>&1
CONP_SPREAD=sweep ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2m_launches_cfg5.csv python bench.py --fast-setup --steps 3 --warmup 3 --blocks 1 --no-parity > /dev/null 2>&1
tail -2 gpurun_out/r2m_ncu.log
