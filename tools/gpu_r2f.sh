#!/bin/bash
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py::test_cfg5_headline_size_matches_oracle 2>&1 | tail -25 > gpurun_out/r2f_tests.log
cat gpurun_out/r2f_tests.log
CONP_DEBUG=1 python bench.py --fast-setup --steps 200 --warmup 10 --blocks 1 --no-parity > gpurun_out/r2f_cfg5.json 2> gpurun_out/r2f_cfg5.err
CONP_DEBUG=1 python bench.py --workload cfg4 --fast-setup --steps 200 --warmup 10 --blocks 1 --no-parity > gpurun_out/r2f_cfg4.json 2> gpurun_out/r2f_cfg4.err
grep -H "k-space stage" gpurun_out/r2f_*.err
python bench.py --fast-setup --steps 3 --warmup 3 --blocks 1 --no-parity > /dev/null 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:spread_mma -c 1 -o gpurun_out/r2f_spread_mma python bench.py --fast-setup --steps 3 --warmup 3 --blocks 1 --no-parity > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log
