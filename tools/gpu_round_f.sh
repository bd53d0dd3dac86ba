#!/bin/bash
# final single-GPU numbers of the round: real setup benches + ncu full capture of the dominant kernel
tag=${1:-x}
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.1f e2e %.1f ms %.4f | %s frac %.3f in-pipe %.4f alone %.4f asym %.2e | upd_frac %.3f | setup %.1fs | cpu %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel"], r["frac"], r["launch_ms_in_pipeline"], r["launch_ms_alone"], r["matrix_asymmetry"], r["update_frac"], d["setup"]["total_s"], d.get("cpu_baseline")))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
python bench.py > gpurun_out/bench_cfg5_n1_$tag.json 2> gpurun_out/bench_cfg5_n1_$tag.err; echo "cfg5 rc=$?"; show gpurun_out/bench_cfg5_n1_$tag.json
python bench.py --workload cfg4 > gpurun_out/bench_cfg4_n1_$tag.json 2> gpurun_out/bench_cfg4_n1_$tag.err; echo "cfg4 rc=$?"; show gpurun_out/bench_cfg4_n1_$tag.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_cfg5_$tag.json 2> gpurun_out/bench_ref_cfg5_$tag.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_cfg5_$tag.json
for wl in cfg5 cfg4; do
ncu --set full --clock-control none --import-source on -k regex:symv_tma_kernel --launch-skip 6 -c 1 -o gpurun_out/symv_${wl}_$tag -f python bench.py --workload $wl --steps 3 --warmup 3 --fast-setup --no-cpu-baseline > gpurun_out/ncu_full_${wl}_$tag.log 2>&1; echo "ncu full $wl rc=$?"
ncu -i gpurun_out/symv_${wl}_$tag.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY'
import csv,sys
rows=list(csv.reader(sys.stdin))
if len(rows)>=3:
    h=rows[0]; u=rows[1]; v=rows[2]
    want=["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","dram__throughput.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","launch__registers_per_thread","sm__warps_active.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","smsp__inst_executed.sum","sm__inst_executed_pipe_fp64.sum"]
    for w in want:
        if w in h:
            i=h.index(w); print(w, v[i], u[i])
PY
done
