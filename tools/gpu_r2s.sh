#!/bin/bash
# per-thread scatter rebuilt; timing + ncu --set full of the short kernels of the step
python -m pytest tests/test_gpu_edge_cases.py -x -q -k "spread_kernels" 2>&1 | tail -3
CONP_DEBUG=1 python bench.py --fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity > gpurun_out/r2s_cfg5_n1.json 2> gpurun_out/r2s_cfg5_n1.err
grep -H "k-space stage" gpurun_out/r2s_*.err | grep "rank 0"
python -c "
import json; d=json.load(open('gpurun_out/r2s_cfg5_n1.json')); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})"
ncu --set full --import-source on --clock-control none -k regex:'zconv|gather_b|symv_reduce|ele_spread|pair_kernel|cell_scan|cell_scatter|mesh_scatter|mesh_bin|pack_count|update_charge|spread_sweep' --launch-skip 60 -c 14 -o gpurun_out/r2s_small python bench.py --fast-setup --steps 3 --warmup 3 --blocks 1 --no-parity > gpurun_out/r2s_ncu.log 2>&1
tail -2 gpurun_out/r2s_ncu.log
