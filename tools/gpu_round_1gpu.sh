#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.1f e2e %.1f ms %.4f | %s frac %.3f in-pipe %.4f alone %.4f | upd_frac %.3f | setup %.1fs | cpu %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel"], r["frac"], r["launch_ms_in_pipeline"], r["launch_ms_alone"], r["update_frac"], d["setup"]["total_s"], d.get("cpu_baseline")))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
python bench.py > gpurun_out/bench_cfg5_n1_$tag.json 2> gpurun_out/bench_cfg5_n1_$tag.err; echo "cfg5 rc=$?"; show gpurun_out/bench_cfg5_n1_$tag.json
python bench.py --workload cfg4 > gpurun_out/bench_cfg4_n1_$tag.json 2> gpurun_out/bench_cfg4_n1_$tag.err; echo "cfg4 rc=$?"; show gpurun_out/bench_cfg4_n1_$tag.json
for wl in cfg5 cfg4; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${wl}_$tag.csv python bench.py --workload $wl --steps 3 --warmup 3 --fast-setup --no-cpu-baseline > gpurun_out/ncu_${wl}_$tag.log 2>&1; echo "ncu $wl rc=$?"
python tools/parse_launches.py gpurun_out/launches_${wl}_$tag.csv | tail -20
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
