#!/bin/bash
# 1-GPU evidence run with the final code: smoke, tests (incl. cfg5 oracle test), headline bench, cfg4 benches, launch list
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r3a_tests.log
cat gpurun_out/r3a_tests.log
python bench.py > gpurun_out/r3a_bench_cfg5_n1.json 2> gpurun_out/r3a_bench_cfg5_n1.err
python bench.py --workload cfg4 > gpurun_out/r3a_bench_cfg4_n1.json 2> gpurun_out/r3a_bench_cfg4_n1.err
python bench.py --workload cfg4 --kspace ewald --steps 100 --warmup 5 --blocks 3 > gpurun_out/r3a_bench_cfg4_ewald_n1.json 2> gpurun_out/r3a_bench_cfg4_ewald_n1.err
for f in gpurun_out/r3a_bench_*.json; do python -c "
import json,sys; d=json.load(open('$f')); p=d.get('parity') or {}; print('$f'.split('/')[-1], round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), p.get('ok'), p.get('max_rel_dq'), p.get('sum_q'), round(d['roofline']['frac'],3), round(d['roofline']['update_frac'],3), d['setup']['build_A_ms'], d['setup']['invert_project_ms'], {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3a_launches_cfg5.csv python bench.py --fast-setup --steps 3 --warmup 3 --blocks 1 --no-parity > gpurun_out/r3a_ncu_launches.log 2>&1
python tools/parse_launches.py gpurun_out/r3a_launches_cfg5.csv > gpurun_out/r3a_step_cfg5.txt; cat gpurun_out/r3a_step_cfg5.txt
