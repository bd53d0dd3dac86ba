#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.1f e2e %.1f ms %.4f | frac %.3f in-pipe %.4f alone %.4f | upd_frac %.3f | %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["frac"], r["launch_ms_in_pipeline"], r["launch_ms_alone"], r["update_frac"], {k: round(v,4) for k,v in r["stage_ms"].items()}))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_$tag.log
for occ in 0 4 3 2; do
for wl in cfg4 cfg5; do
CONP_SPREAD_BLOCKS_PER_SM=$occ python bench.py --workload $wl --steps 300 --fast-setup --no-cpu-baseline > gpurun_out/bench_${wl}_occ${occ}_$tag.json 2> gpurun_out/bench_${wl}_occ${occ}_$tag.err; echo "$wl occ$occ rc=$?"
show gpurun_out/bench_${wl}_occ${occ}_$tag.json
done; done
for wl in cfg4 cfg5; do
CONP_DEBUG=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${wl}_$tag.csv python bench.py --workload $wl --steps 3 --warmup 3 --fast-setup --no-cpu-baseline > gpurun_out/ncu_${wl}_$tag.log 2>&1; echo "ncu $wl rc=$?"
grep "conp\]" gpurun_out/ncu_${wl}_$tag.log | head -4
python tools/parse_launches.py gpurun_out/launches_${wl}_$tag.csv | tail -24
done
