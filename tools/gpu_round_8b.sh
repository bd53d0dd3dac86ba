#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "n=%d value %.1f e2e %.1f ms %.4f | frac %.3f in-pipe %.4f alone %.4f | setup %.1f | %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], r["frac"], r["launch_ms_in_pipeline"], r["launch_ms_alone"], d["setup"]["total_s"], {k: round(v,4) for k,v in r["stage_ms"].items()}))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 500 --warmup 20 > gpurun_out/bench_cfg5_n8_$tag.json 2> gpurun_out/bench_cfg5_n8_$tag.err; echo "cfg5 n=8 rc=$?"
show gpurun_out/bench_cfg5_n8_$tag.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_cfg5_n8_$tag.err | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 500 --warmup 20 --fast-setup > gpurun_out/bench_cfg5_n4_fast_$tag.json 2> gpurun_out/bench_cfg5_n4_fast_$tag.err; echo "cfg5 n=4 fast rc=$?"
show gpurun_out/bench_cfg5_n4_fast_$tag.json
