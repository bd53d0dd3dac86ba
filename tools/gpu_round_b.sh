#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_$tag.log
for wl in cfg4 cfg5; do
python bench.py --workload $wl --steps 300 --fast-setup --no-cpu-baseline > gpurun_out/bench_${wl}_fast_$tag.json 2> gpurun_out/bench_${wl}_fast_$tag.err; echo "$wl fast rc=$?"
python - gpurun_out/bench_${wl}_fast_$tag.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print(sys.argv[1], "value %.1f e2e %.1f ms %.4f | %s frac %.3f in-pipe %.4f alone %.4f | upd_frac %.3f | %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel"], r["frac"], r["launch_ms_in_pipeline"], r["launch_ms_alone"], r["update_frac"], {k: round(v,4) for k,v in r["stage_ms"].items()}))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${wl}_$tag.csv python bench.py --workload $wl --steps 3 --warmup 3 --fast-setup --no-cpu-baseline > gpurun_out/ncu_${wl}_$tag.log 2>&1; echo "ncu $wl rc=$?"
python tools/parse_launches.py gpurun_out/launches_${wl}_$tag.csv | tail -24
done
