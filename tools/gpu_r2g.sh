#!/bin/bash
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py::test_cfg5_headline_size_matches_oracle 2>&1 | tail -25 > gpurun_out/r2g_tests.log
cat gpurun_out/r2g_tests.log
for tz in 4 8; do
CONP_SPREAD_TILE=$tz,8,32 CONP_DEBUG=1 python bench.py --fast-setup --steps 200 --warmup 10 --blocks 1 --no-parity > gpurun_out/r2g_cfg5_tz$tz.json 2> gpurun_out/r2g_cfg5_tz$tz.err
CONP_SPREAD_TILE=$tz,8,32 CONP_DEBUG=1 python bench.py --workload cfg4 --fast-setup --steps 200 --warmup 10 --blocks 1 --no-parity > gpurun_out/r2g_cfg4_tz$tz.json 2> gpurun_out/r2g_cfg4_tz$tz.err
done
grep -H "k-space stage" gpurun_out/r2g_*.err
python bench.py --fast-setup --steps 3 --warmup 3 --blocks 1 --no-parity > /dev/null 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:spread_mma -c 1 -o gpurun_out/r2g_spread_mma python bench.py --fast-setup --steps 3 --warmup 3 --blocks 1 --no-parity > gpurun_out/r2g_ncu.log 2>&1
tail -2 gpurun_out/r2g_ncu.log
