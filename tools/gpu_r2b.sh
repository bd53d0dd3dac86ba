#!/bin/bash
# round 2: tile spread A/B
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py::test_cfg5_headline_size_matches_oracle 2>&1 | tail -15 > gpurun_out/r2b_tests.log
for cfg in "8,8,32" "4,16,32" "8,16,32" "16,8,32" "4,8,32"; do
  CONP_DEBUG=1 CONP_SPREAD_TILE=$cfg python bench.py --fast-setup --steps 200 --warmup 10 --blocks 1 --no-parity > gpurun_out/r2b_tile_${cfg//,/_}.json 2> gpurun_out/r2b_tile_${cfg//,/_}.err
done
CONP_DEBUG=1 CONP_SPREAD_ATOMIC=1 python bench.py --fast-setup --steps 200 --warmup 10 --blocks 1 --no-parity > gpurun_out/r2b_atomic.json 2> gpurun_out/r2b_atomic.err
CONP_DEBUG=1 python bench.py --workload cfg4 --fast-setup --steps 200 --warmup 10 --blocks 1 --no-parity > gpurun_out/r2b_cfg4.json 2> gpurun_out/r2b_cfg4.err
cat gpurun_out/r2b_tests.log
grep -h "k-space stage" gpurun_out/r2b_*.err
