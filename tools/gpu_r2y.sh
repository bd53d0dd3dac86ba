#!/bin/bash
# pair kernel with batched run descriptors + candidate prefetch: parity and timing
python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -4 > gpurun_out/r2y_tests.log
cat gpurun_out/r2y_tests.log
run() {
  python bench.py $1 > gpurun_out/$2.json 2> gpurun_out/$2.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/$2.json')); print('$2', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})" || tail -c 600 gpurun_out/$2.err
}
run "--fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity" r2y_cfg5_n1
run "--workload cfg4 --fast-setup --steps 500 --warmup 10 --blocks 3 --no-parity" r2y_cfg4_n1
