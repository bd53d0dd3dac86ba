#!/bin/bash
# co-residency experiments: caps on the pair / spread kernels' blocks per SM, unsorted red.global spread
run() { # N port extra-args out
  if [ $1 -eq 1 ]; then python bench.py $3 > gpurun_out/$4.json 2> gpurun_out/$4.err; else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 $3 > gpurun_out/$4.json 2> gpurun_out/$4.err; fi
  python -c "
import json,sys; d=json.load(open('gpurun_out/$4.json')); print('$4', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))" || tail -c 600 gpurun_out/$4.err
  grep -h "graph timeline" gpurun_out/$4.err | head -1
}
A4="--workload cfg4 --fast-setup --steps 500 --warmup 10 --blocks 3 --no-parity"
A5="--fast-setup --steps 300 --warmup 10 --blocks 3 --no-parity"
run 1 0 "$A4" r2w_cfg4_A
CONP_SPREAD_UNSORTED=1 run 1 0 "$A4" r2w_cfg4_B_unsorted
CONP_SPREAD_UNSORTED=1 CONP_PAIR_BLOCKS_PER_SM=2 run 1 0 "$A4" r2w_cfg4_C_unsorted_pair2
CONP_SPREAD_UNSORTED=1 CONP_PAIR_BLOCKS_PER_SM=2 CONP_SPREAD_BLOCKS_PER_SM=4 run 1 0 "$A4" r2w_cfg4_D_unsorted_pair2_spread4
CONP_PAIR_BLOCKS_PER_SM=2 run 1 0 "$A4" r2w_cfg4_E_pair2
CONP_SPREAD_UNSORTED=1 CONP_SPREAD_BLOCKS_PER_SM=4 run 1 0 "$A4" r2w_cfg4_F_unsorted_spread4
CONP_SPREAD_UNSORTED=1 CONP_PAIR_BLOCKS_PER_SM=3 CONP_SPREAD_BLOCKS_PER_SM=5 run 1 0 "$A4" r2w_cfg4_G_unsorted_pair3_spread5
CONP_TRACE=1 CONP_SPREAD_UNSORTED=1 CONP_PAIR_BLOCKS_PER_SM=2 CONP_SPREAD_BLOCKS_PER_SM=4 run 1 0 "$A4" r2w_cfg4_D_trace
CONP_PAIR_BLOCKS_PER_SM=2 run 2 29711 "$A5" r2w_cfg5_n2_pair2
CONP_PAIR_BLOCKS_PER_SM=2 CONP_SPREAD_BLOCKS_PER_SM=4 run 2 29712 "$A5" r2w_cfg5_n2_pair2_spread4
CONP_TRACE=1 CONP_PAIR_BLOCKS_PER_SM=2 CONP_SPREAD_BLOCKS_PER_SM=4 run 2 29713 "$A5" r2w_cfg5_n2_pair2_spread4_trace
CONP_PAIR_BLOCKS_PER_SM=2 run 1 0 "$A5" r2w_cfg5_n1_pair2
