#!/bin/bash
# Ewald tensor-core path parity + tile-spread ncu capture
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py::test_cfg5_headline_size_matches_oracle 2>&1 | tail -15 > gpurun_out/r2d_tests.log
cat gpurun_out/r2d_tests.log
python bench.py --workload cfg4 --kspace ewald --steps 50 --warmup 5 --blocks 2 --no-cpu-baseline > gpurun_out/r2d_cfg4_ewald.json 2> gpurun_out/r2d_cfg4_ewald.err
tail -c 400 gpurun_out/r2d_cfg4_ewald.err
CONP_SPREAD_TILE=4,8,32 python bench.py --fast-setup --steps 3 --warmup 3 --blocks 1 --no-parity > /dev/null 2>&1 && \
CONP_SPREAD_TILE=4,8,32 ncu --set full --import-source on --clock-control none -k regex:spread_tile -c 1 -o gpurun_out/r2d_spread_tile python bench.py --fast-setup --steps 3 --warmup 3 --blocks 1 --no-parity > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log
