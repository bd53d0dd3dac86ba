#!/usr/bin/env python
"""bench.py -- electrode charge updates/sec of the per-step solve (b_cal +
S.b + epilogue + electrode re-spread) on synthetic graphite capacitors.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload cfg4|cfg5|...] [--kspace pppm|ewald]

One "step" = one `conp_pre_force` (FixConp::pre_force, fix_conp.cpp:543-573)
for the whole electrode with freshly jittered electrolyte positions.
Prints ONE JSON line (contract in the task statement):
  value   device-resident updates/s (inputs already in HBM), CUDA events on the
          library's stream, max over ranks
  e2e     the same through the host-pointer C-ABI call, H2D of positions and D2H
          of charges inside the timed region
  roofline      GEMV launch (dominant kernel): algorithmic bytes / event time
  cpu_baseline  the CPU oracle (port of the reference algorithm) on a bounded
                sample of the same workload, all host cores
`--impl reference` times only that CPU port (the reference needs LAMMPS, which
is not available) and never touches the CUDA library.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "lammps-user-conp2_b200"))
# keep stdout to the single JSON line: NCCL's version/debug banner goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

from conp_b200 import MockLammps, make_workload  # noqa: E402
from conp_b200.mockhost import mesh_for_spacing  # noqa: E402
from conp_b200.system import WORKLOADS  # noqa: E402

METRIC = "electrode charge updates/sec"
UNIT = "updates/s"
G_EWALD, CUT, ETA, DV, ACC, MESH_H, SLAB = 0.26, 12.0, 1.979, 2.0, 1e-4, 1.0, 3.0
JITTER = 0.05
NSETS = 8


def describe(name):
    w = WORKLOADS[name]
    n_ele = 2 * w["nlayers"] * 4 * w["ncx"] * w["ncy"]
    return (f"{name}: synthetic graphite capacitor, {n_ele} electrode atoms / {w['n_elyte']} electrolyte charges, "
            f"conp, slab {SLAB}, g_ewald {G_EWALD}, cut {CUT} A, PPPM order 5 mesh<= {MESH_H} A, Nevery=1")


def make_case(name, kspace):
    s = make_workload(name)
    lmp = MockLammps(s, "p p f")
    lmp.pair_style_coul_long(CUT)
    mesh = mesh_for_spacing(s.prd, SLAB, MESH_H) if kspace == "pppm" else None
    lmp.kspace("pppm/conp" if kspace == "pppm" else "pppm", ACC, G_EWALD, slab=SLAB, mesh=mesh)
    lmp.group_molecule("eleleft", 1)
    lmp.group_molecule("eleright", 2)
    arg = f"e eleleft conp 1 eleright {ETA} {DV} log_conp etypes 1 3".split() + (["pppm"] if kspace == "pppm" else [])
    return lmp, arg


def config_dict(name, kspace, lmp, n_ele, kcount_a):
    """`config` of the JSON line: identical for the b200 arm and the reference arm (it names the
    workload, not the implementation)."""
    return {"workload": describe(name), "kspace": kspace, "mesh": [int(v) for v in lmp.mesh] if lmp.mesh else None,
            "kcount_A": int(kcount_a),
            "l2_policy": f"inputs larger than any cache: the {8.0 * n_ele * n_ele / 1e9:.2f} GB S matrix is streamed "
                         f"from memory every step; {NSETS} jittered position sets (sigma {JITTER} A) in rotation"}


# measured DRAM bytes per launch (ncu --set full, profiles/): (workload, kernel) -> read + written
TRAFFIC = {("cfg5", "gemv"): 12.801121e9 + 6.872832e6, ("cfg4", "gemv"): 800.11392e6 + 3.297536e6,
           ("cfg5", "symv"): 6.615391e9 + 11.123200e6, ("cfg4", "symv"): 406.715392e6 + 5.890816e6}

DEFAULT_WORKLOAD = "cfg5"  # BASELINE configs[4]: the configuration the 1/2/4/8-GPU metric is quoted on; fits one GPU


def synthetic_matrix(n):
    """Symmetric random stand-in for S (profiling / reference-arm timing only): u_i u_j-type low-rank
    blocks, cheap to build even at n = 40 000 (12.8 GB) and symmetric to the last bit, like the
    projected inverse the library builds itself."""
    rng = np.random.default_rng(1234)
    u = rng.standard_normal((n, 4)) * 3e-2
    S = u @ u.T
    S = np.minimum(S, S.T)  # BLAS may not return an exactly symmetric product
    return S


def jitter_sets(x, nsets, seed):
    """Jittered copies of ALL non-electrode positions from one seeded stream: every rank generates the
    same global sets and slices its owned atoms out of them, so an N-rank run solves exactly the
    systems the 1-rank run and the CPU oracle solve."""
    rng = np.random.default_rng(seed)
    return [x + rng.normal(0.0, JITTER, x.shape) for _ in range(nsets)]


class ClockSampler:
    """SM clock / throttle-reason samples of one GPU taken by a thread through NVML (no fork, no
    nvidia-smi start-up inside the run).  Started BEFORE the warm-up, so nothing is spawned between
    the barrier and the timed loop; only the samples inside [mark_begin, mark_end] are reported."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, device, period_s=0.02):
        import threading
        self.samples, self.t0, self.t1 = [], None, None
        self.period = period_s
        self.err = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.nv = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = f"NVML unavailable: {e}"
            return
        self._stop = threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.samples.append((time.perf_counter(), sm, rs))
            except Exception as e:  # noqa: BLE001
                self.err = str(e)
                return
            self._stop.wait(self.period)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.err and not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err]}
        self._stop.set()
        self.th.join(timeout=2)
        t0, t1 = self.t0 or 0.0, self.t1 or float("inf")
        rows = [r for r in self.samples if t0 <= r[0] <= t1] or self.samples
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no samples"]}
        sm = [r[1] for r in rows]
        bits = 0
        for r in rows:
            bits |= r[2]
        reasons = sorted(nm for nm, b in self.REASONS if bits & b)
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": self.sm_max,
                "reasons": reasons, "samples": len(rows), "source": "NVML thread, timed region only"}


def make_cpu_port(lmp, arg, S):
    """CPU oracle (port of the reference algorithm) set up on the `inv`-file path with matrix S."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import conp_oracle as O
    cores = os.cpu_count() or 1
    try:  # torchrun exports OMP_NUM_THREADS=1: undo it for this process' BLAS/OpenMP pools
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=cores)
    except Exception:  # noqa: BLE001
        pass
    O.lib().orc_set_num_threads(cores)
    fix = O.OracleFixConp(lmp, arg, fft_workers=cores)
    fix.setup_preinverted(S)
    return fix, cores


def cpu_port_updates_per_s(fix, lmp, sets, budget_s):
    """Times the CPU oracle's per-step path (b_cal + matvec + epilogue [+ electrode re-spread]) on
    jittered positions; returns (updates/s, n_updates)."""
    lmp.system.x[fix.oth_idx] = sets[0]
    fix.pre_force()  # warm-up (FFT plans, page faults)
    n, t0 = 0, time.perf_counter()
    while True:
        lmp.system.x[fix.oth_idx] = sets[(n + 1) % len(sets)]
        fix.pre_force()
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= 50:
            break
    return n / el, n


def run_reference(args, rank, world, out):
    """Reference arm: CPU port of the reference's per-step path, all host cores."""
    if rank != 0:
        return
    name = args.workload or DEFAULT_WORKLOAD
    lmp, arg = make_case(name, args.kspace)
    n_ele = int((lmp.system.mol > 0).sum())
    cores = os.cpu_count() or 1
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=cores)  # also first-touches S from all cores (NUMA)
    except Exception:  # noqa: BLE001
        pass
    # S only feeds the O(N^2) matvec, whose cost does not depend on its values; the
    # true S needs the O(N^2 K) A build, which no CPU finishes in minutes at this size.
    S = synthetic_matrix(n_ele)
    fix, cores = make_cpu_port(lmp, arg, S)
    import conp_oracle as O
    s_ = lmp.system
    kcount_a = O.OracleEwald(lmp.g_ewald, lmp.accuracy, lmp.q2(), s_.natoms, s_.prd, lmp.slabflag,
                             lmp.slab_volfactor).kcount
    sets = jitter_sets(lmp.system.x[fix.oth_idx].copy(), 3, 99)
    per_step_budget = 8.0
    ups, n = cpu_port_updates_per_s(fix, lmp, sets, per_step_budget * max(1, min(args.steps, 3)))
    line = {
        "impl": "reference", "metric": METRIC, "value": ups, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / ups, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(name, args.kspace, lmp, n_ele, kcount_a),
        "parallelism": f"OpenMP threads x{cores} on one shared S",
        "cpu_baseline": {"value": ups, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} full updates of the same workload (random stand-in S, true b path)"},
        "e2e": {"value": ups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference algorithm (oracle/), not the reference binary: LAMMPS is not available",
    }
    out.emit(json.dumps(line))


def gather_matrix_to_rank0(ctx, info, N, rank, world):
    """The GPU-built S (row blocks, one per rank) assembled on rank 0's host."""
    import torch
    import torch.distributed as dist
    mine = ctx.get_matrix()
    if world == 1:
        return mine
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"  # (gloo: the CPU test of this plumbing)
    rows = torch.tensor([info.row_begin, info.row_end], dtype=torch.int64, device=dev)
    allrows = [torch.zeros_like(rows) for _ in range(world)]
    dist.all_gather(allrows, rows)
    allrows = [tuple(int(v) for v in t.cpu()) for t in allrows]
    S = None
    if rank == 0:
        S = np.empty((N, N))
        S[info.row_begin:info.row_end] = mine
        for r in range(1, world):
            a, b = allrows[r]
            if b > a:
                buf = torch.empty((b - a, N), dtype=torch.float64, device=dev)
                dist.recv(buf, src=r)
                S[a:b] = buf.cpu().numpy()
                del buf
    elif info.row_end > info.row_begin:
        dist.send(torch.from_numpy(np.ascontiguousarray(mine)).to(dev), dst=0)
    return S


class QuietStdout:
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torch.distributed) write to
    file descriptor 1 on their own, so everything goes to stderr until the line is printed."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    with QuietStdout() as out:
        run(out)


def run(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--kspace", default="pppm", choices=["pppm", "ewald"])
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--blocks", type=int, default=5, help="timed blocks of --steps (median block is `value`)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--fast-setup", action="store_true",
                    help="PROFILING ONLY: load a random S instead of building/inverting A (line is marked invalid)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    args.blocks = max(1, args.blocks)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return

    import torch
    import torch.distributed as dist
    from conp_b200 import abi
    from conp_b200.fix_conp import make_fix

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        buf = torch.zeros(abi.CONP_UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(abi.get_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())

    name = args.workload or DEFAULT_WORKLOAD
    lmp, arg = make_case(name, args.kspace)
    t0 = time.perf_counter()
    fix = make_fix(lmp, arg, device=local_rank, rank=rank, nranks=world, unique_id=uid)
    if args.fast_setup:
        fix.setup_post_neighbor()
        if fix.args.pppmflag:
            t = lmp.pppm_tables()
            fix.ctx.pppm_setup(t.mesh, t.order, t.rho_coeff, t.greensfn, t.shift, t.shiftone)
        Sr = synthetic_matrix(fix.N)
        fix.ctx.load_matrix(Sr, True)
        del Sr
        fix.totsetq = fix.ctx.set_unit_voltage(fix.evscale)
        fix.runstage = 3
        fix.post_neighbor()
    else:
        fix.setup()
    setup_s = time.perf_counter() - t0
    ctx = fix.ctx
    info = ctx.info()
    kmode = fix.kspace_mode
    N, M = fix.N, info.n_elyte
    nlocal = len(fix.owned)

    # jittered position sets of the WHOLE system from one seed, then this rank's owned atoms:
    # pinned host copies (e2e leg) and device copies (resident leg)
    oth = np.nonzero(fix.side_all == 0)[0]
    own_in_oth = np.searchsorted(oth, fix.owned)
    global_sets = jitter_sets(lmp.system.x[oth], NSETS, 20261018)
    host_sets = [torch.from_numpy(np.ascontiguousarray(s[own_in_oth])).pin_memory() for s in global_sets]
    parity_set = global_sets[0].copy()
    del global_sets
    dev_sets = [h.cuda() for h in host_sets]
    q_host = torch.zeros(N, dtype=torch.float64).pin_memory()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def over_ranks(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(v):
        return over_ranks(v, dist.ReduceOp.MAX) if world > 1 else v

    def all_ranks(v):
        if world == 1:
            return [v]
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    # the clock sampler is a thread of rank 0, started before any warm-up: nothing is spawned between a
    # barrier and a timed loop (round 1's nvidia-smi fork there stalled the peers of rank 0)
    sampler = ClockSampler(local_rank) if rank == 0 else None

    # ---- device-resident leg: `blocks` timed regions of exactly `steps` solves each ----------------
    for k in range(args.warmup):
        ctx.solve_device(dev_sets[k % NSETS].data_ptr(), kmode, 0, DV)
    barrier()
    if sampler:
        sampler.mark_begin()
    block_ms, block_rank_ms, launches = [], [], 0
    for blk in range(args.blocks):
        l0 = ctx.info().launches
        ctx.timer_record(0)
        for k in range(args.steps):
            ctx.solve_device(dev_sets[k % NSETS].data_ptr(), kmode, 0, DV)
        ctx.timer_record(1)
        barrier()
        mine = ctx.timer_elapsed_ms(0, 1)
        block_rank_ms.append([v / args.steps for v in all_ranks(mine)])
        block_ms.append(max(block_rank_ms[-1]))
        launches = int(ctx.info().launches - l0)
    if sampler:
        sampler.mark_end()
    order_ = sorted(range(args.blocks), key=lambda i: block_ms[i])
    med = order_[len(order_) // 2]
    ms_step = block_ms[med]
    value = 1e3 / ms_step

    # ---- end-to-end leg through the host-pointer ABI call ---------------------
    for k in range(max(3, args.warmup // 4)):
        ctx.pre_force_into(host_sets[k % NSETS].data_ptr(), kmode, 0, DV, q_host.data_ptr())
    e2e_blocks = []
    for blk in range(args.blocks):
        barrier()
        ctx.timer_record(2)
        t_wall = time.perf_counter()
        for k in range(args.steps):
            ctx.pre_force_into(host_sets[k % NSETS].data_ptr(), kmode, 0, DV, q_host.data_ptr())
        ctx.timer_record(3)
        ctx.sync()
        wall_ms = (time.perf_counter() - t_wall) * 1e3
        barrier()
        e2e_blocks.append(max_over_ranks(max(ctx.timer_elapsed_ms(2, 3), wall_ms)) / args.steps)
    e2e_ms = sorted(e2e_blocks)[len(e2e_blocks) // 2]
    clocks = sampler.stop() if sampler else None

    dgemm_tf = ctx.bench_dgemm_tflops(8192) if (rank == 0 and not args.fast_setup) else None
    # ---- per-stage event timing inside the pipeline (roofline of the GEMV) -------
    ctx.stage_times(True)
    nst = min(args.steps, 50)
    for k in range(nst):
        ctx.solve_device(dev_sets[k % NSETS].data_ptr(), kmode, 0, DV)
    _, st = ctx.stage_times(False)
    gemv_alone_ms = ctx.bench_gemv(20)
    stage_names = ["pack", "bin", "pair", "kspace", "gather", "exchange_b", "gemv", "epilogue"]
    gemv_ms = max_over_ranks(float(st[6]))
    nrows = info.row_end - info.row_begin
    sym = bool(ctx.info().symmetric_matvec)
    # Algorithmic bytes of the matvec launch.  General kernel: the row block of S once.
    # Symmetric kernel: S = S^T, so the half band of every row (N/2 + 1 columns) is all the product
    # needs; the same work expressed in the 8d figure (full rows) is reported as gemv_equivalent_gbs.
    full_bytes = 8.0 * nrows * N + 8.0 * N + 8.0 * nrows
    gemv_bytes = (8.0 * nrows * (N // 2 + 1) + 8.0 * N + 8.0 * N) if sym else full_bytes
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = gemv_bytes / (gemv_ms * 1e-3) / 1e9
    G = int(np.prod(lmp.mesh)) if lmp.mesh else 0
    order = lmp.order
    rest = (8.0 * N + 16.0 * nrows + 32.0 * M
            + ((24.0 * G + 24.0 * order * nrows) if kmode == 1 else
               (16.0 * info.kcount + 16.0 * info.kcount_flat * nrows)))
    b_contract = 8.0 * nrows * N + rest                       # SURVEY 8d contract figure
    b_update = (gemv_bytes - 16.0 * N if sym else 8.0 * nrows * N) + rest
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel from the
    # ncu --set full captures under profiles/ (one GPU); other shapes: null
    traffic = TRAFFIC.get((name, "symv" if sym else "gemv")) if world == 1 else None
    roofline = {"bound": "hbm", "kernel": "symv_tma_kernel" if sym else "gemv_tma_kernel",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_launch": gemv_bytes, "launch_ms_in_pipeline": gemv_ms, "launch_ms_alone": gemv_alone_ms,
                "symmetric_matvec": sym, "matrix_asymmetry": ctx.info().asymmetry,
                "gemv_equivalent_gbs": full_bytes / (gemv_ms * 1e-3) / 1e9,
                "update_bytes": b_update, "update_achieved_gbs": b_update / (ms_step * 1e-3) / 1e9,
                "update_frac": b_update / (ms_step * 1e-3) / 1e9 / peak,
                "update_contract_bytes": b_contract,
                "update_contract_gbs": b_contract / (ms_step * 1e-3) / 1e9,
                "stage_ms": {n_: float(v) for n_, v in zip(stage_names, st)},
                "stage_sum_ms": float(sum(st)),
                "stage_note": "stages timed one after another (eager, no overlap); in the timed run the pair "
                              "kernel and the brick clears run beside the k-space chain in the CUDA graph"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_dict(name, args.kspace, lmp, N, info.kcount),
        "parallelism": (f"S rows sharded x{world}; electrolyte: "
                        + ("z-slabs of the PPPM mesh" if kmode == 1 else "structure-factor chunks") + f" x{world}"),
        "matvec_stream": f"{gemv_bytes/1e6:.0f} MB of the S row block per GPU per step"
                         f"{' (half band of the symmetric matrix)' if sym else ''}",
        "timing": {"blocks": args.blocks, "steps_per_block": args.steps, "value_is": "median block",
                   "ms_per_step_blocks": block_ms, "ms_per_step_min": min(block_ms), "ms_per_step_max": max(block_ms),
                   "ms_per_step_per_rank": block_rank_ms[med], "e2e_ms_per_step_blocks": e2e_blocks},
        "electrode_atom_updates_per_s": value * N,
        "e2e": {"value": 1e3 / e2e_ms, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(nlocal * 24), "d2h_bytes_per_step": int(N * 8 + 16)},
        "gpu_launches": launches,
        "roofline": roofline,
        "clocks": clocks,
        "setup": {"total_s": setup_s, "build_A_ms": info.setup_build_ms, "invert_project_ms": info.setup_invert_ms,
                  "gram_flops_full": 4.0 * nrows * N * info.kcount,
                  "gram_tflops_full_equiv": 4.0 * nrows * N * info.kcount / max(info.setup_build_ms, 1e-9) / 1e9,
                  "gram_note": "FP64 DMMA; the lower triangle only is computed, so the full-Gram-equivalent "
                               "rate can exceed the DGEMM ceiling; includes panel generation and the real-space part",
                  "cublas_dgemm_tflops_ceiling": dgemm_tf},
    }

    # ---- parity at this size and this N: GPU (through the host-pointer ABI) vs the CPU oracle on the
    # same positions, with the GPU-built S (the oracle's `inv`-file setup, fix_conp.cpp:442-445) ------
    if args.fast_setup:
        line["INVALID"] = "fast-setup profiling run: S is random, not a bench value"
    line["cpu_baseline"] = None
    if not args.fast_setup and not args.no_parity:
        xp = torch.from_numpy(np.ascontiguousarray(parity_set[own_in_oth])).pin_memory()
        scal_gpu = ctx.pre_force_into(xp.data_ptr(), kmode, 0, DV, q_host.data_ptr())
        q_gpu = q_host.numpy().copy()
        b_gpu, _ = ctx.get_b()
        rank_dq = 0.0
        if world > 1:  # the replicated epilogue must give identical charges on every rank
            q0 = torch.from_numpy(q_gpu).cuda()
            dist.broadcast(q0, 0)
            rank_dq = max_over_ranks(float(np.abs(q0.cpu().numpy() - q_gpu).max()))
        S = gather_matrix_to_rank0(ctx, info, N, rank, world)
        if rank == 0:
            lmp2, arg2 = make_case(name, args.kspace)
            ofix, cores = make_cpu_port(lmp2, arg2, S)
            lmp2.system.x[ofix.oth_idx] = parity_set
            q_ref = ofix.pre_force().copy()
            b_ref = ofix.bbb_all
            dq = float(np.abs(q_gpu - q_ref).max())
            qmax = float(np.abs(q_ref).max())
            db = float(np.abs(b_gpu - b_ref).max())
            bmax = float(np.abs(b_ref).max())
            tol_q, tol_b = 1e-9 * qmax + 1e-12, 5e-11 * max(bmax, 1.0)
            ok = (dq <= tol_q and db <= tol_b and abs(float(q_gpu.sum())) < 1e-11 and rank_dq == 0.0
                  and abs(scal_gpu - ofix.scalar_output) <= 1e-9 * abs(ofix.scalar_output) + 1e-12)
            line["parity"] = {"max_abs_dq": dq, "max_rel_dq": dq / qmax, "b_max_err": db, "b_max": bmax,
                              "q_max": qmax, "sum_q": float(q_gpu.sum()), "scalar_gpu": scal_gpu,
                              "scalar_ref": float(ofix.scalar_output), "max_dq_between_ranks": rank_dq,
                              "tol": {"dq": "1e-9*max|q| + 1e-12 e", "b": "5e-11*max|b|", "sum_q": 1e-11},
                              "against": "CPU oracle per-step path on the same jittered positions, GPU-built S "
                                         f"(assembled from {world} row block(s))", "ok": bool(ok)}
            if world == 1 and not args.no_cpu_baseline:
                sets3 = jitter_sets(lmp2.system.x[ofix.oth_idx].copy(), 3, 99)
                ups, n = cpu_port_updates_per_s(ofix, lmp2, sets3, args.cpu_budget)
                line["cpu_baseline"] = {"value": ups, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": f"{n} full updates of the same workload with the GPU-built S"}
            del S
    if rank == 0:
        out.emit(json.dumps(line))
    fix.close()
    bad = rank == 0 and "parity" in line and not line["parity"]["ok"]
    if world > 1:
        dist.destroy_process_group()
    if bad:
        raise SystemExit("bench.py: PARITY FAILED " + json.dumps(line["parity"]))


if __name__ == "__main__":
    main()
