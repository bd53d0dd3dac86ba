#!/bin/bash
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py::test_cfg5_headline_size_matches_oracle 2>&1 | tail -25 > gpurun_out/r2e_tests.log
cat gpurun_out/r2e_tests.log
for cfg in "8,8,32" "4,8,32"; do
  CONP_DEBUG=1 CONP_SPREAD_TILE=$cfg python bench.py --fast-setup --steps 200 --warmup 10 --blocks 1 --no-parity > gpurun_out/r2e_tile_${cfg//,/_}.json 2> gpurun_out/r2e_tile_${cfg//,/_}.err
done
CONP_DEBUG=1 python bench.py --workload cfg4 --fast-setup --steps 200 --warmup 10 --blocks 1 --no-parity > gpurun_out/r2e_cfg4.json 2> gpurun_out/r2e_cfg4.err
grep -H "k-space stage" gpurun_out/r2e_*.err
python bench.py --workload cfg4 --kspace ewald --steps 50 --warmup 5 --blocks 2 --no-cpu-baseline > gpurun_out/r2e_cfg4_ewald.json 2> gpurun_out/r2e_cfg4_ewald.err
python -c "
import json; d=json.load(open('gpurun_out/r2e_cfg4_ewald.json')); print(d['value'], d['ms_per_step'], d['parity']['ok'], d['roofline']['stage_ms'])"
