#!/bin/bash
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py::test_cfg5_headline_size_matches_oracle 2>&1 | tail -15 > gpurun_out/r2c_tests.log
for cfg in "8,8,32" "4,8,32" "4,16,32"; do
  CONP_DEBUG=1 CONP_SPREAD_TILE=$cfg python bench.py --fast-setup --steps 200 --warmup 10 --blocks 1 --no-parity > gpurun_out/r2c_tile_${cfg//,/_}.json 2> gpurun_out/r2c_tile_${cfg//,/_}.err
done
CONP_DEBUG=1 python bench.py --workload cfg4 --fast-setup --steps 200 --warmup 10 --blocks 1 --no-parity > gpurun_out/r2c_cfg4.json 2> gpurun_out/r2c_cfg4.err
cat gpurun_out/r2c_tests.log
grep -H "k-space stage" gpurun_out/r2c_*.err
