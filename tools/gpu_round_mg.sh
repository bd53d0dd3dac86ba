#!/bin/bash
# multi-GPU visit: bash tools/gpu_round_mg.sh <tag> <ngpus>
tag=${1:-x}; n=${2:-2}
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_mg_$tag.log 2>&1; echo "pytest mg rc=$?"
tail -3 gpurun_out/pytest_mg_$tag.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "n=%d value %.1f e2e %.1f ms %.4f | frac %.3f in-pipe %.4f alone %.4f | %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], r["frac"], r["launch_ms_in_pipeline"], r["launch_ms_alone"], {k: round(v,4) for k,v in r["stage_ms"].items()}))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
for p2p in 1 0; do
CONP_P2P=$p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 300 --warmup 20 --fast-setup --no-cpu-baseline > gpurun_out/bench_cfg5_n${n}_p2p${p2p}_$tag.json 2> gpurun_out/bench_cfg5_n${n}_p2p${p2p}_$tag.err; echo "cfg5 n=$n p2p=$p2p rc=$?"
show gpurun_out/bench_cfg5_n${n}_p2p${p2p}_$tag.json
done
