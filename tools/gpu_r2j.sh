#!/bin/bash
# 1-GPU evidence run: tests (incl. cfg5 oracle test), headline bench, cfg4 benches, launch list, ncu captures
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2j_tests.log
cat gpurun_out/r2j_tests.log
python bench.py > gpurun_out/r2j_bench_cfg5_n1.json 2> gpurun_out/r2j_bench_cfg5_n1.err
python bench.py --workload cfg4 > gpurun_out/r2j_bench_cfg4_n1.json 2> gpurun_out/r2j_bench_cfg4_n1.err
python bench.py --workload cfg4 --kspace ewald --steps 100 --warmup 5 --blocks 3 > gpurun_out/r2j_bench_cfg4_ewald_n1.json 2> gpurun_out/r2j_bench_cfg4_ewald_n1.err
for f in gpurun_out/r2j_bench_*.json; do python -c "
import json,sys; d=json.load(open('$f')); p=d.get('parity') or {}; print('$f'.split('/')[-1], round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), p.get('ok'), p.get('max_rel_dq'), p.get('sum_q'), round(d['roofline']['frac'],3), round(d['roofline']['update_frac'],3), d['setup']['build_A_ms'], d['setup']['invert_project_ms'], {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})"; done
# launch list of one short run (kernel share of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2j_launches_cfg5.csv python bench.py --fast-setup --steps 3 --warmup 3 --blocks 1 --no-parity > gpurun_out/r2j_ncu_launches.log 2>&1
# full captures: symv, gram (during a cfg4 setup), the Ewald tn gemm
ncu --set full --import-source on --clock-control none -k regex:symv_tma -c 1 -o gpurun_out/r2j_symv python bench.py --fast-setup --steps 3 --warmup 3 --blocks 1 --no-parity > gpurun_out/r2j_ncu_symv.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:gram_kernel --launch-skip 3 -c 1 -o gpurun_out/r2j_gram python bench.py --workload cfg4 --steps 3 --warmup 3 --blocks 1 --no-parity --no-cpu-baseline > gpurun_out/r2j_ncu_gram.log 2>&1
tail -2 gpurun_out/r2j_ncu_gram.log
