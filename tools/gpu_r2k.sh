#!/bin/bash
# 8-GPU box: scaling with parity in every line, plus stage splits with / without the routed exchange
run() { # N port extra-args out
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 $3 > gpurun_out/$4.json 2> gpurun_out/$4.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/$4.json')); p=d.get('parity') or {}; print('$4', round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), p.get('ok'), p.get('max_rel_dq'), p.get('max_dq_between_ranks'), d['setup']['build_A_ms'], {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})" || tail -c 600 gpurun_out/$4.err
}
run 8 29601 "--steps 300 --warmup 20 --blocks 5" r2k_cfg5_n8
run 4 29602 "--steps 300 --warmup 20 --blocks 5" r2k_cfg5_n4
run 2 29603 "--steps 300 --warmup 20 --blocks 5" r2k_cfg5_n2
CONP_DEBUG=1 run 8 29604 "--fast-setup --steps 300 --warmup 20 --blocks 3 --no-parity" r2k_cfg5_n8_fast
CONP_ROUTE=0 CONP_DEBUG=1 run 8 29605 "--fast-setup --steps 300 --warmup 20 --blocks 3 --no-parity" r2k_cfg5_n8_fast_noroute
CONP_SPREAD=mma CONP_DEBUG=1 run 8 29606 "--fast-setup --steps 300 --warmup 20 --blocks 3 --no-parity" r2k_cfg5_n8_fast_mma
run 8 29607 "--workload cfg4 --steps 300 --warmup 20 --blocks 5" r2k_cfg4_n8
grep -h "k-space stage" gpurun_out/r2k_cfg5_n8_fast*.err | grep "rank 0" | head
