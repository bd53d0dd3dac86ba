#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$tag.log
for wl in cfg5 cfg4; do
python bench.py --workload $wl --steps 500 --fast-setup --no-cpu-baseline > gpurun_out/bench_${wl}_fast_$tag.json 2> gpurun_out/bench_${wl}_fast_$tag.err; echo "$wl rc=$?"
python - gpurun_out/bench_${wl}_fast_$tag.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print(sys.argv[1], "value %.1f e2e %.1f ms %.4f | in-pipe %.4f alone %.4f | clocks %s | %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["launch_ms_in_pipeline"], r["launch_ms_alone"], d["clocks"], {k: round(v,4) for k,v in r["stage_ms"].items()}))
PY
done
