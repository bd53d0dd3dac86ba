#!/bin/bash
# One GPU-box visit: parity tests, then the benches whose numbers go under profiles/.
# usage (from the repo root, via gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$tag.log
tail -5 gpurun_out/pytest_$tag.log
CONP_DEBUG=1 python bench.py --workload cfg4 --steps 500 > gpurun_out/bench_cfg4_$tag.json 2> gpurun_out/bench_cfg4_$tag.err; echo "cfg4 rc=$?"
CONP_NO_SYMV=1 python bench.py --workload cfg4 --steps 500 --no-cpu-baseline > gpurun_out/bench_cfg4_nosymv_$tag.json 2> gpurun_out/bench_cfg4_nosymv_$tag.err; echo "cfg4 nosymv rc=$?"
CONP_NO_OVERLAP=1 python bench.py --workload cfg4 --steps 500 --no-cpu-baseline > gpurun_out/bench_cfg4_noovl_$tag.json 2> gpurun_out/bench_cfg4_noovl_$tag.err; echo "cfg4 nooverlap rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_cfg4_*%s.json" % "")):
    pass
PY
for f in gpurun_out/bench_cfg4_$tag.json gpurun_out/bench_cfg4_nosymv_$tag.json gpurun_out/bench_cfg4_noovl_$tag.json; do
python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[1], "value %.1f e2e %.1f ms %.4f | %s frac %.3f in-pipe %.4f alone %.4f | upd_frac %.3f | %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel"], r["frac"], r["launch_ms_in_pipeline"], r["launch_ms_alone"], r["update_frac"], {k: round(v,4) for k,v in r["stage_ms"].items()}))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
