#!/bin/bash
tag=${1:-x}; n=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all_$tag.log 2>&1; echo "pytest all rc=$?"
tail -3 gpurun_out/pytest_all_$tag.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "n=%d value %.1f e2e %.1f ms %.4f | in-pipe %.4f | %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], r["launch_ms_in_pipeline"], {k: round(v,4) for k,v in r["stage_ms"].items()}))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
CONP_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 300 --warmup 20 --fast-setup --no-cpu-baseline > gpurun_out/bench_cfg5_n${n}_$tag.json 2> gpurun_out/bench_cfg5_n${n}_$tag.err; echo "cfg5 n=$n rc=$?"
show gpurun_out/bench_cfg5_n${n}_$tag.json
grep "k-space stage" gpurun_out/bench_cfg5_n${n}_$tag.err
