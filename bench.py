#!/usr/bin/env python
"""bench.py -- AMG-PCG theta-sweep hot path on B200 (contract: task prompt section 4).

Headline workload (BASELINE.json configs[1]): 3D diffusion, Q1 on [-1,1]^3, piecewise-
discontinuous coefficients (checkerboard, contrast 1e6), m=200 -> 8 120 601 DoFs,
217 081 801 nnz, theta sweep 0.05..0.95 step 0.05 (19 systems sharing one matrix,
ref testcase2-diffusion-structured/src/main.cpp:440-467), tol 1e-8 absolute.

One "step" = one theta sweep = 19 x (AMG setup + PCG solve) on the resident matrix.
metric = seconds per system (setup + solve), lower is better.

  value  : device-resident (CSR, rhs, x0 already in HBM), timed with CUDA events
  e2e    : through the reference-facing API with HOST buffers (matrix upload once
           per sweep, x/b H2D and x/residual-history D2H per system)
  N > 1  : one process per GPU, each rank sweeps its own system (weak scaling,
           no data-path collective); time = max over ranks.  The same line then carries a
           `partitioned` block: BASELINE config 5 (m=464, 100.5 M DoFs), ONE system
           row-partitioned over the N GPUs with halo exchange over NVLink.
  N = 1  : the line also carries a `configs` block with the other BASELINE configs
           (1: m=100; 3: Q1 elasticity m=187 with and without aggressive levels; 4: batch of
           m=46 systems through the C++ driver; 5: the m=464 system on one GPU as the anchor
           of the partitioned curve) and the roofline of the pooling kernel.
  --impl reference : the reference's own CPU flavour (Falgout + hybrid symmetric
           GS, the CPU restatement in oracle/; hypre itself is not installable
           here), one system per host thread, on a bounded sample.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

THETA = ("0.05", "0.96", "0.05")  # ref testcase2-diffusion-structured/datagen.py:47
TOL = 1e-8                        # ref datagen.py:9 (absolute)
PATTERN, MODE, CONTRAST = 4, 3, 6.0
DEVICE_ALGO = "PMIS + classical interp + C/F l1-Jacobi V(1,1)"
REFERENCE_ALGO = "Falgout + classical interp + hybrid symmetric Gauss-Seidel V(1,1) (PCHYPRE defaults)"


_T0 = time.perf_counter()


def _progress(what):
    """one line per phase on stderr: if a run ever stops responding, its log says where"""
    print(f"[bench {time.perf_counter() - _T0:7.1f} s] {what}", file=sys.stderr, flush=True)


def workload_name(m, algo=DEVICE_ALGO):
    n = (m + 1) ** 3
    return (f"3D diffusion Q1, checkerboard mu in {{1,1e6}} on a {PATTERN}^3 pattern, m={m} "
            f"({n} DoFs), theta sweep 0.05..0.95 step 0.05 (19 systems/matrix), "
            f"{algo}, PCG tol 1e-8 abs")


def make_system(ab, m, seed_shift=0):
    epsv = ab.gen.checkerboard_epsv(PATTERN, MODE, CONTRAST)
    if seed_shift:
        epsv = np.roll(epsv, seed_shift)
    return ab.gen.poisson_q1(m, PATTERN, MODE, epsv)


def device_options(ab, theta, agg=0):
    R = ab.RelaxationType
    return ab.AdditionalData(True, theta, 0.9, agg, True, relaxation_type_up=R.l1scaledJacobi,
                             relaxation_type_down=R.l1scaledJacobi)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_traffic(m):
    """DRAM bytes per launch of the roofline's launch set (the three level-0 Jacobi half
    sweeps of one V(1,1) cycle at theta = 0.25: F, F, C) from the committed ncu capture
    (profiles/traffic.json); only valid for the mesh it was captured on."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if m != 200 or not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f).get("smooth_l0_theta0.25")
    return t["traffic_bytes_per_launch"] if t else None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def family_table(fam):
    kern = {}
    total_ms = sum(v["ms"] for k, v in fam.items() if not k.endswith("_l0")) or 1.0
    for k, v in fam.items():
        if v["launches"] and v["ms"] > 0:
            kern[k] = {"ms": round(v["ms"], 3), "launches": v["launches"],
                       "GBps": round(v["bytes"] / v["ms"] / 1e6, 1),
                       "share": round(v["ms"] / total_ms, 4) if not k.endswith("_l0") else None}
    return kern


# --------------------------------------------------------------------------- CPU legs
def oracle_system_seconds(s, theta, flavour):
    """One (matrix, theta) system on one host core: setup + solve wall seconds."""
    import amg_ann_b200 as ab
    from oracle import binding as orc
    if flavour == "reference":   # Falgout + hybrid symmetric GS (PCHYPRE defaults, A.2)
        d = ab.AdditionalData(True, theta, 0.9, 0, True, coarsen_type=ab.COARSEN_FALGOUT)
    else:                         # same algorithm as the device path
        d = device_options(ab, theta)
    t0 = time.perf_counter()
    H = orc.Hierarchy(s.rowptr32(), s.col, s.val, d.to_struct())
    t1 = time.perf_counter()
    rc, _, nit, _ = H.cg_solve(s.rhs, s.x0, abs_tol=TOL)
    t2 = time.perf_counter()
    H.close()
    return dict(theta=theta, setup_s=t1 - t0, solve_s=t2 - t1, iters=int(nit), rc=int(rc))


def cpu_sample(s, thetas, flavour, threads):
    """`threads` host threads, each solving its share of `thetas` on system s.
    Returns (wall seconds per system for the whole job, per-system records)."""
    work = list(thetas)
    out, lock = [], threading.Lock()

    def run():
        while True:
            with lock:
                if not work:
                    return
                th = work.pop()
            r = oracle_system_seconds(s, th, flavour)
            with lock:
                out.append(r)
    t0 = time.perf_counter()
    ts = [threading.Thread(target=run) for _ in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    wall = time.perf_counter() - t0
    return wall / len(out), out


def scaling_exponent(t_small, nnz_small, t_big, nnz_big):
    if t_small <= 0 or t_big <= 0 or nnz_big == nnz_small:
        return 1.0
    return math.log(t_big / t_small) / math.log(nnz_big / nnz_small)


def run_reference(args):
    """--impl reference: CPU restatement of the reference flavour (Falgout + symmetric GS),
    one system per host thread on all host cores.  The mesh of the timed sample is the
    largest m <= --cpu-m whose step fits the wall-clock budget; seconds are extrapolated to
    the m=200 system with the nnz-scaling exponent MEASURED between a calibration size and
    the sample size (not assumed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import amg_ann_b200 as ab
    cores = os.cpu_count() or 1
    thetas = ab.gen.theta_sweep(*map(float, THETA))
    n_full, nnz_full = ab.gen.sizes(0, args.m)
    # calibration (also the W untimed warm-up steps): a small mesh, same theta mix
    m_cal = min(40, args.cpu_m)
    s_cal = make_system(ab, m_cal)
    cal = []
    for step in range(max(1, args.warmup)):
        sel = [thetas[(step * cores + i) % len(thetas)] for i in range(cores)]
        sec, recs = cpu_sample(s_cal, sel, "reference", cores)
        cal.append(max(r["setup_s"] + r["solve_s"] for r in recs))   # step wall = slowest thread
    t_cal = float(np.median(cal))
    _, nnz_cal = ab.gen.sizes(0, m_cal)
    # the largest sample mesh whose K steps fit the budget (time ~ nnz^1.1 between these sizes)
    budget = args.cpu_budget_s / max(1, args.steps)
    m_s = m_cal
    for m_try in range(args.cpu_m, m_cal, -4):
        _, nnz_try = ab.gen.sizes(0, m_try)
        if t_cal * (nnz_try / nnz_cal) ** 1.1 <= budget:
            m_s = m_try
            break
    s = make_system(ab, m_s) if m_s != m_cal else s_cal
    _, nnz_s = ab.gen.sizes(0, m_s)
    per_step, per_theta = [], {}
    for step in range(args.steps):
        # one step = `cores` systems of the sweep in parallel (the reference fans
        # independent runs out over processes: 00_data-generation.py:105-116)
        sel = [thetas[((step + args.warmup) * cores + i) % len(thetas)] for i in range(cores)]
        sec, recs = cpu_sample(s, sel, "reference", cores)
        per_step.append(sec)
        for r in recs:
            per_theta.setdefault(f"{r['theta']:.2f}", []).append(r)
    sec_s = float(np.mean(per_step))
    # exponent from the two measured sizes, on the per-system CPU seconds of the common thetas
    one_cal = float(np.mean([r["setup_s"] + r["solve_s"] for r in cpu_sample(s_cal, thetas[::4], "reference", cores)[1]]))
    one_s = float(np.mean([np.mean([r["setup_s"] + r["solve_s"] for r in per_theta[k]])
                           for k in per_theta if k in {f"{t:.2f}" for t in thetas[::4]}] or
                          [np.mean([r["setup_s"] + r["solve_s"] for rs in per_theta.values() for r in rs])]))
    expo = scaling_exponent(one_cal, nnz_cal, one_s, nnz_s) if m_s != m_cal else 1.0
    expo_used = min(max(expo, 1.0), 1.3)
    scale = (nnz_full / nnz_s) ** expo_used
    v = sec_s * scale
    prof = {k: {"iters": int(np.median([r["iters"] for r in rs])),
                "setup_s": round(float(np.mean([r["setup_s"] for r in rs])), 3),
                "solve_s": round(float(np.mean([r["solve_s"] for r in rs])), 3)}
            for k, rs in sorted(per_theta.items())}
    sample = (f"m={m_s} ({(m_s + 1) ** 3} DoFs, {nnz_s} nnz) instead of m={args.m}; {cores} systems per step, "
              f"one per host thread, theta values cycling through the 19-value sweep; measured {sec_s:.3f} "
              f"s/system at the sample size; scaled to m={args.m} by (nnz ratio {nnz_full / nnz_s:.2f})^{expo_used:.3f}, "
              f"the exponent measured between m={m_cal} and m={m_s} ({expo:.3f}, clamped to [1, 1.3]); warm-up "
              f"steps run at m={m_cal}; CPU restatement of hypre {REFERENCE_ALGO} (hypre itself not installable)")
    line = {"impl": "reference", "metric": "AMG-PCG setup+solve seconds per system",
            "value": v, "unit": "s/system", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec_s * cores * 1e3,
            "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args.m, REFERENCE_ALGO) + f"; timed on a bounded sample at m={m_s}",
                       "sample_m": m_s, "sample_nnz": nnz_s, "calibration_m": m_cal,
                       "nnz_scaling_exponent_measured": round(expo, 4), "nnz_scaling_exponent_used": round(expo_used, 4),
                       "per_theta_at_sample": prof},
            "cpu_baseline": {"value": v, "unit": "s/system", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": v, "unit": "s/system", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def cpu_baseline_port(ab, m_full, nnz_full):
    """The oracle with the DEVICE algorithm on one core over the sweep's theta mix (every
    third value) at m=64, scaled with the exponent measured against m=44."""
    thetas = ab.gen.theta_sweep(*map(float, THETA))
    sel = thetas[::3]
    t0 = time.perf_counter()
    s_a, s_b = make_system(ab, 44), make_system(ab, 64)
    sec_a, _ = cpu_sample(s_a, sel, "port", 1)
    sec_b, recs = cpu_sample(s_b, sel, "port", 1)
    expo = scaling_exponent(sec_a, s_a.nnz, sec_b, s_b.nnz)
    used = min(max(expo, 1.0), 1.3)
    scale = (nnz_full / s_b.nnz) ** used
    return {"value": sec_b * scale, "unit": "s/system", "cores": 1, "kind": "port",
            "sample": (f"oracle ({DEVICE_ALGO}, same options as the device) at m=64 ({s_b.n} DoFs), theta in "
                       f"{[round(t, 2) for t in sel]} (every third value of the sweep), 1 thread; measured "
                       f"{sec_b:.3f} s/system, scaled to m={m_full} by (nnz ratio {nnz_full / s_b.nnz:.1f})^{used:.3f} "
                       f"(exponent measured between m=44 and m=64: {expo:.3f})"),
            "measured_s_per_system_at_sample": sec_b,
            "iters_at_sample": {f"{r['theta']:.2f}": r["iters"] for r in recs},
            "wall_s": round(time.perf_counter() - t0, 2)}


# --------------------------------------------------------------------------- extras (N = 1)
def timed_solve(ab, L, ctx, A, n, d_b, d_x0, d_x, data, reps=3):
    """setup + solve of one system, device-resident, CUDA events on ctx's stream; best of reps."""
    import torch
    from amg_ann_b200._native import c_f64p
    best = None
    hist, nit = np.zeros(8192), C.c_int64()
    for _ in range(reps):
        d_x.copy_(d_x0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        P = ab.PreconditionBoomerAMG()
        P.initialize(A, data, ctx)
        ev[1].record()
        rc = L.amgb_cg_solve_device(ctx._h, A._h, C.c_void_p(d_x.data_ptr()), C.c_void_p(d_b.data_ptr()), P._h,
                                    n, TOL, hist.ctypes.data_as(c_f64p), len(hist), C.byref(nit))
        ev[2].record()
        torch.cuda.synchronize()
        if rc != 0:
            raise RuntimeError(f"amgb_cg_solve_device -> {rc}: {L.amgb_last_error(ctx._h).decode()}")
        st = P.level_stats()
        P.close()
        r = dict(setup_ms=ev[0].elapsed_time(ev[1]), solve_ms=ev[1].elapsed_time(ev[2]), iters=int(nit.value),
                 levels=[int(v) for v in st["rows"]], operator_complexity=round(float(st["operator"]), 4))
        if best is None or r["setup_ms"] + r["solve_ms"] < best["setup_ms"] + best["solve_ms"]:
            best = r
    best["s_per_system"] = (best["setup_ms"] + best["solve_ms"]) / 1e3
    return best


def extra_config1_and_pooling(ab, L, ctx, peak):
    """BASELINE config 1 (m=100, mu=1, theta=0.25, device assembly) + pooled image V=75."""
    import torch
    m = 100
    n = (m + 1) ** 3
    d_b = torch.empty(n, dtype=torch.float64, device="cuda")
    d_x0 = torch.empty(n, dtype=torch.float64, device="cuda")
    A = ab.SparseMatrix.assemble_poisson_q1(ctx, m, 1, 1, None, d_b.data_ptr(), d_x0.data_ptr())
    d_x = torch.empty_like(d_x0)
    out = {"config1_m100": dict(workload=f"3D Poisson Q1, m=100 ({n} DoFs), theta=0.25, {DEVICE_ALGO}",
                                **timed_solve(ab, L, ctx, A, n, d_b, d_x0, d_x, device_options(ab, 0.25)))}
    A.close()
    return out


def extra_pooling(ab, ctx, A, n, nnz, peak, V=75):
    """Pooling kernel on the resident config-2 matrix: algorithmic bytes = CSR read + 28 V^2 written."""
    vm = ab.ViewMaker(V)
    ts = []
    for _ in range(5):
        vm.make_view(A)
        ts.append(vm.t_device_us)
    t = float(np.median(ts[1:]))
    by = 12.0 * nnz + 4.0 * (n + 1) + 28.0 * V * V
    return {"pooling": {"workload": f"matrix -> {V}x{V} image (sum, count, max+, max-), m=200 CSR resident",
                        "device_us": round(t, 1),
                        "roofline": {"bound": "hbm", "achieved": round(by / t / 1e3, 1), "peak": peak, "unit": "GB/s",
                                     "frac": round(by / t / 1e3 / peak, 4), "algorithmic_bytes_per_launch": round(by),
                                     "traffic": None}}}


def extra_config4(systems=64, threads=8):
    """BASELINE config 4 through the C++ driver: a batch of m=46 systems x 19 theta, then pooled images."""
    exe = os.path.join(ROOT, "amg-ann_b200", "host", "amgb_datagen")
    if not os.path.exists(exe):
        return {"config4_m46": {"skipped": "amgb_datagen not built"}}
    out = {}
    with tempfile.TemporaryDirectory() as d:
        base = [exe, "--m", "46", "--systems", str(systems), "--threads", str(threads), "--seed", "0",
                "--device-assembly", "1"]
        r = subprocess.run(base + ["--out", os.path.join(d, "stats.csv")], capture_output=True, text=True, timeout=240)
        if r.returncode != 0:
            return {"config4_m46": {"failed": r.stderr[-300:]}}
        j = json.loads(r.stdout.strip().splitlines()[-1])
        try:
            v = subprocess.run(base + ["--make-view", "1", "--out", os.path.join(d, "views.csv")], capture_output=True,
                               text=True, timeout=120)
            jv = json.loads(v.stdout.strip().splitlines()[-1]) if v.returncode == 0 else None
        except subprocess.TimeoutExpired:
            jv = None
        out["config4_m46"] = {
            "workload": f"{systems} systems of 103823 DoFs (seeds 0..{systems - 1}) x 19 theta, C++ driver amgb_datagen, "
                        f"{threads} host threads / contexts on one GPU, device assembly, CSV rows written",
            "solves": j["solves"], "wall_s": round(j["wall_s"], 3),
            "systems_per_s": round(j["solves"] / j["wall_s"], 1),
            "steady_systems_per_s": (round(j["steady_solves"] / j["steady_s"], 1) if j.get("steady_s", 0) > 0 else None),
            "steady_note": "from the moment every host thread has finished its first system (process start-up, "
                           "contexts, first launch of every kernel) to the end",
            "sum_setup_s": round(j["setup_s"], 3), "sum_solve_s": round(j["solve_s"], 3),
            "images": jv["views"] if jv else None,
            "images_per_s": round(jv["views"] / jv["wall_s"], 1) if jv else None,
            "image_device_us": round(1e6 * jv["view_device_s"] / max(1, jv["views"]), 1) if jv else None}
    return out


def extra_config3(ab, L, ctx, m=187):
    """BASELINE config 3: Q1 elasticity, 3 DoFs/node, scalar AMG; agg_nl = 0 and 2 (ref t3 main.cpp:456)."""
    import torch
    t0 = time.perf_counter()
    s = ab.gen.elasticity_q1(m, PATTERN, MODE, 10.0 ** ab.gen.checkerboard_epsv(PATTERN, MODE, 2.0))
    t_gen = time.perf_counter() - t0
    A = ab.SparseMatrix(ctx, s.rowptr32(), s.col, s.val)
    d_b = torch.from_numpy(s.rhs).cuda()
    d_x0 = torch.from_numpy(s.x0).cuda()
    d_x = torch.empty_like(d_x0)
    out = {}
    ctx.reserve(int(3.2 * 12 * s.nnz))     # the pool grown once instead of allocation by allocation
    for agg in (0, 2):
        # the first pass pays for kernel loading and whatever the pool still has to grow: reported, not hidden
        t0 = time.perf_counter()
        first = timed_solve(ab, L, ctx, A, s.n, d_b, d_x0, d_x, device_options(ab, 0.25, agg), reps=1)
        first_wall = time.perf_counter() - t0
        r = timed_solve(ab, L, ctx, A, s.n, d_b, d_x0, d_x, device_options(ab, 0.25, agg), reps=1)
        out[f"config3_m{m}_agg{agg}"] = dict(
            workload=f"Q1 elasticity m={m} ({s.n} DoFs, {s.nnz} nnz), theta=0.25, aggressive levels {agg}, {DEVICE_ALGO}",
            host_generation_s=round(t_gen, 1), first_pass={"setup_ms": round(first["setup_ms"], 1),
                                                          "solve_ms": round(first["solve_ms"], 1),
                                                          "wall_s": round(first_wall, 2)}, **r)
    A.close()
    return out


def config5_threads_on_one_gpu(ab, m, nblocks, thetas):
    """The row-partitioned path on ONE GPU: `nblocks` row blocks handled by host threads that share
    the device (local nnz must stay < 2^31, and config 5 has 2.7e9 entries).  The anchor of the
    strong-scaling curve of the `partitioned` block."""
    import torch
    from amg_ann_b200 import dist
    from amg_ann_b200._native import amgb_lib, c_f64p
    L = amgb_lib()
    starts = dist.slab_partition(m, nblocks)
    n = (m + 1) ** 3
    barrier = threading.Barrier(nblocks)

    def fn(rank, comm):
        ctx = comm.ctx
        b0, e0 = starts[rank], starts[rank + 1]
        d_b = torch.empty(e0 - b0, dtype=torch.float64, device="cuda")
        d_x0 = torch.empty(e0 - b0, dtype=torch.float64, device="cuda")
        A = dist.DistSparseMatrix.assemble_poisson_q1(comm, m, b0, e0, 1, 1, None, d_b.data_ptr(), d_x0.data_ptr())
        ctx.synchronize()
        d_x = torch.empty_like(d_x0)
        hist, nit = np.zeros(4096), C.c_int64()
        res = {}
        for rep in range(2):           # first pass = warm-up (pool growth, windows)
            for th in thetas:
                d_x.copy_(d_x0)
                torch.cuda.synchronize()
                barrier.wait()
                t0 = time.perf_counter()
                P = dist.DistPreconditionBoomerAMG()
                P.initialize(A, device_options(ab, th))
                ctx.synchronize()
                barrier.wait()
                t1 = time.perf_counter()
                rc = L.amgb_dist_cg_solve_device(ctx._h, C.c_void_p(d_x.data_ptr()), C.c_void_p(d_b.data_ptr()), P._h,
                                                 n, TOL, hist.ctypes.data_as(c_f64p), len(hist), C.byref(nit))
                barrier.wait()
                t2 = time.perf_counter()
                if rc != 0:
                    raise RuntimeError(f"amgb_dist_cg_solve_device -> {rc}: {L.amgb_last_error(ctx._h).decode()}")
                st = P.level_stats()
                res[th] = dict(setup_s=t1 - t0, solve_s=t2 - t1, iters=int(nit.value),
                               levels=[int(v) for v in st["rows"]], nnz=int(st["nnz"][0]))
                P.close()
        A.close()
        return res

    parts = dist.run_local_group(nblocks, fn)
    r0 = parts[0]
    per = {f"{th:.2f}": {"setup_s": round(max(p[th]["setup_s"] for p in parts), 4),
                         "solve_s": round(max(p[th]["solve_s"] for p in parts), 4),
                         "iters": r0[th]["iters"]} for th in thetas}
    value = sum(v["setup_s"] + v["solve_s"] for v in per.values()) / len(thetas)
    return {"workload": f"3D Poisson Q1, m={m} ({n} DoFs, {r0[thetas[0]]['nnz']} nnz), ONE system in {nblocks} row "
                        f"blocks on ONE GPU (host threads; local nnz < 2^31), theta in {thetas}, {DEVICE_ALGO}",
            "value": round(value, 4), "unit": "s/system", "n_gpus": 1, "row_blocks": nblocks,
            "timing": "host wall clock around synchronised calls, max over blocks, second pass",
            "per_theta": per, "levels": r0[thetas[0]]["levels"]}


# --------------------------------------------------------------------------- partitioned block (N > 1)
def partitioned_block(ab, args, rank, world, local, stream, m):
    """BASELINE config 5: ONE system row-partitioned in z-slabs over the N GPUs of this run (device
    assembly; halo exchange, replication all-gather and dot-product reductions as put/flag kernels over
    NVLink peer windows, NCCL when AMGB_PEER=0).  step = setup + PCG solve for theta in {0.25, 0.5}
    (SURVEY.md 8d config 5); strong scaling (total work fixed)."""
    import torch
    import torch.distributed as tdist
    from amg_ann_b200 import dist
    from amg_ann_b200._native import amgb_lib, c_f64p
    L = amgb_lib()
    ctx = ab.Context(local, stream.cuda_stream)
    comm = dist.Communicator.nccl_from_torch(ctx)
    thetas = [0.25, 0.5]
    exch = "NCCL" if os.environ.get("AMGB_PEER", "1").startswith("0") else "NVLink peer-window (put/flag kernels)"
    starts = dist.slab_partition(m, world)
    b0, e0 = starts[rank], starts[rank + 1]
    n = (m + 1) ** 3
    d_b = torch.empty(e0 - b0, dtype=torch.float64, device="cuda")
    d_x0 = torch.empty(e0 - b0, dtype=torch.float64, device="cuda")
    A = dist.DistSparseMatrix.assemble_poisson_q1(comm, m, b0, e0, 1, 1, None, d_b.data_ptr(), d_x0.data_ptr())
    ctx.synchronize()
    d_x = torch.empty_like(d_x0)
    hist, nit = np.zeros(4096), C.c_int64()
    info, split = {}, {}

    def tmax(v):
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    def step(record=False):
        for th in thetas:
            d_x.copy_(d_x0)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record(stream)
            P = dist.DistPreconditionBoomerAMG()
            P.initialize(A, device_options(ab, th))
            ev[1].record(stream)
            rc = L.amgb_dist_cg_solve_device(ctx._h, C.c_void_p(d_x.data_ptr()), C.c_void_p(d_b.data_ptr()), P._h,
                                             n, TOL, hist.ctypes.data_as(c_f64p), len(hist), C.byref(nit))
            ev[2].record(stream)
            if rc != 0:
                raise RuntimeError(f"amgb_dist_cg_solve_device -> {rc}: {L.amgb_last_error(ctx._h).decode()}")
            if record:
                torch.cuda.synchronize()
                info[th] = (nit.value, P.level_stats())
                split[th] = (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]))
            P.close()

    def barrier():
        tdist.barrier()
        torch.cuda.synchronize()

    t_cold = time.perf_counter()
    step()
    barrier()
    cold_s = tmax(time.perf_counter() - t_cold)
    steps = max(1, min(args.steps, args.part_steps))
    barrier()
    ctx.reset_kernel_launches()
    e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0_.record(stream)
    for k in range(steps):
        step(record=(k == steps - 1))
    e1_.record(stream)
    barrier()
    launches = ctx.kernel_launches()
    ms_step = tmax(e0_.elapsed_time(e1_)) / steps
    setup_s = {th: tmax(split[th][0]) / 1e3 for th in thetas}
    solve_s = {th: tmax(split[th][1]) / 1e3 for th in thetas}
    # per-family device times of one more step (plain launches: the timers disable the captured cycle)
    ctx.enable_timers(True)
    ctx.reset_timers()
    step()
    fam = ctx.timers()
    ctx.enable_timers(False)
    fam_ms = sum(v["ms"] for k, v in fam.items() if not k.endswith("_l0")) or 1.0
    exch_ms = fam.get("exchange", {"ms": 0.0})["ms"]
    exch_share = tmax(exch_ms / fam_ms)
    peak, _ = measured_peaks()
    d = fam["smooth_l0"]
    block = None
    if rank == 0:
        st = info[thetas[0]][1]
        block = {"workload": f"3D Poisson Q1, m={m} ({n} DoFs, {int(st['nnz'][0])} nnz), ONE system row-partitioned in "
                             f"z-slabs over {world} GPUs, device assembly, {exch} halo exchange and dot-product "
                             f"reductions, theta in {thetas}, {DEVICE_ALGO}, PCG tol 1e-8 abs",
                 "metric": "AMG-PCG setup+solve seconds per system", "value": ms_step / 1e3 / len(thetas),
                 "unit": "s/system", "n_gpus": world, "steps": steps, "warmup": 1, "scaling": "strong",
                 "ms_per_step": ms_step, "timing": "CUDA events on the library's stream, max over ranks",
                 "setup_s": {f"{th:.2f}": round(setup_s[th], 4) for th in thetas},
                 "solve_s": {f"{th:.2f}": round(solve_s[th], 4) for th in thetas},
                 "iters": {f"{th:.2f}": info[th][0] for th in thetas},
                 "levels": [int(r) for r in st["rows"]], "operator_complexity": round(float(st["operator"]), 4),
                 "first_step_s": round(cold_s, 3), "gpu_launches_rank0": launches,
                 "exchange_share_of_kernel_time": round(exch_share, 4),
                 "smooth_l0_GBps_rank0": round(d["bytes"] / max(d["ms"], 1e-9) / 1e6, 1),
                 "smooth_l0_frac_of_peak": round(d["bytes"] / max(d["ms"], 1e-9) / 1e6 / peak, 4),
                 "families_rank0": family_table(fam)}
    A.close()
    comm.close()
    ctx.close()
    return block


# --------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import amg_ann_b200 as ab
    from amg_ann_b200._native import amgb_lib, c_f64p

    t_start = time.perf_counter()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU implementation")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    L = amgb_lib()
    thetas = ab.gen.theta_sweep(*map(float, THETA))
    nsys = len(thetas)
    s = make_system(ab, args.m, seed_shift=rank)   # each rank its own system (weak scaling)
    n, nnz = s.n, s.nnz
    rp32 = s.rowptr32()

    # a real (non-legacy) stream shared by torch and the library, so the CUDA
    # events below are recorded on the stream the kernels are launched on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = ab.Context(local, stream.cuda_stream)

    # ---- device-resident inputs for `value`
    d_rp = torch.from_numpy(rp32).cuda()
    d_col = torch.from_numpy(s.col).cuda()
    d_val = torch.from_numpy(s.val).cuda()
    d_b = torch.from_numpy(s.rhs).cuda()
    d_x0 = torch.from_numpy(s.x0).cuda()
    d_x = torch.empty_like(d_x0)
    A_dev = ab.SparseMatrix.wrap_device(ctx, n, nnz, d_rp.data_ptr(), d_col.data_ptr(), d_val.data_ptr())
    results = {}

    # The systems of a sweep are independent (the reference fans them out over processes,
    # 00_data-generation.py:105-116): `lanes` host threads, one context + stream each, take
    # them from a queue, so one system's latency-bound phases (coarse levels, PMIS rounds,
    # host round trips) overlap another's bandwidth-bound ones on the same GPU.
    lanes = max(1, args.streams)
    lane_streams = [stream] + [torch.cuda.Stream() for _ in range(lanes - 1)]
    lane_ctx = [ctx] + [ab.Context(local, st.cuda_stream) for st in lane_streams[1:]]
    lane_x = [d_x] + [torch.empty_like(d_x0) for _ in range(lanes - 1)]
    _progress(f"system generated and resident (n = {n}, nnz = {nnz}), {lanes} lanes")
    # every lane's pool grown once, up front, instead of allocation by allocation: measured at m = 200 with
    # 3 lanes, a sweep takes 4.53 +- 0.005 s with the reserve and 4.5 ... 7.4 s without (pool growth
    # stalls every lane; which lane grows when depends on the order the lanes pick their systems in)
    reserve = int(args.reserve_gb * 2 ** 30) if args.reserve_gb >= 0 else 100 * nnz
    if reserve > 0:
        for c in lane_ctx:
            c.reserve(reserve)

    # phase locks (--phases 1): of the systems in flight at most one is in its setup and one in
    # its solve, so what overlaps on the GPU is always an instruction/latency-bound setup with a
    # bandwidth-bound solve (two or three solves side by side only share the same HBM)
    class _Free:
        def __enter__(self): return self
        def __exit__(self, *a): return False
    phases = bool(args.phases) and lanes > 1
    setup_lock = threading.Lock() if phases else _Free()
    solve_lock = threading.Lock() if phases else _Free()
    phase_locks = (setup_lock, solve_lock) if phases else None

    def solve_device(w, th, profile=None):
        c = lane_ctx[w]
        h, k = np.zeros(4096), C.c_int64()
        with setup_lock:
            with torch.cuda.stream(lane_streams[w]):
                lane_x[w].copy_(d_x0)                # solution = m_zero_solution (t2 main.cpp:446)
            t0 = time.perf_counter()
            P = ab.PreconditionBoomerAMG()
            P.initialize(A_dev, device_options(ab, th), c)
            c.synchronize()
            t1 = time.perf_counter()
        with solve_lock:
            t1s = time.perf_counter()
            rc = L.amgb_cg_solve_device(c._h, A_dev._h, C.c_void_p(lane_x[w].data_ptr()),
                                        C.c_void_p(d_b.data_ptr()), P._h, n, TOL,
                                        h.ctypes.data_as(c_f64p), len(h), C.byref(k))
            t2 = time.perf_counter()
        if rc != 0:
            raise RuntimeError(f"amgb_cg_solve_device -> {rc}: {L.amgb_last_error(c._h).decode()}")
        if os.environ.get("BENCH_DEBUG"):
            print(f"[bench]   lane {w} theta {th:.2f}: setup {1e3 * (t1 - t0):7.1f} ms, solve {1e3 * (t2 - t1s):7.1f} ms "
                  f"({k.value} it), started {t0 - t_start:8.3f}", file=sys.stderr, flush=True)
        results[th] = (k.value, P.level_stats() if th == thetas[0] else None)
        if profile is not None:
            profile[th] = (t1 - t0, t2 - t1s, k.value)
        P.close()

    def fan_out(solve):
        """run solve(lane, theta) for the whole sweep; most expensive systems (large theta) first"""
        if lanes == 1:
            for th in thetas:
                solve(0, th)
            return
        todo, lock, errs = list(thetas), threading.Lock(), []

        def worker(w):
            torch.cuda.set_device(local)
            try:
                while True:
                    with lock:
                        if not todo or errs:
                            return
                        th = todo.pop()
                    solve(w, th)
            except BaseException as e:  # noqa: BLE001 - re-raised below
                errs.append(e)
        ts = [threading.Thread(target=worker, args=(w,)) for w in range(lanes)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        if errs:
            raise errs[0]

    def sweep_device():
        fan_out(solve_device)

    # ---- host buffers for `e2e` (pinned)
    def pinned(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_rp, h_col, h_val = pinned(rp32), pinned(s.col), pinned(s.val)
    h_b, h_x0 = pinned(s.rhs), pinned(s.x0)
    h_xs = [pinned(s.x0) for _ in range(lanes)]
    e2e_bytes = {"h2d": 0, "d2h": 0}

    def sweep_e2e():
        A = ab.SparseMatrix(ctx, h_rp, h_col, h_val)     # H2D of the CSR, once per sweep
        moved = {"h2d": h_rp.nbytes + h_col.nbytes + h_val.nbytes, "d2h": 0}
        mlock = threading.Lock()

        def solve_host(w, th):
            h_x = h_xs[w]
            h_x[...] = h_x0
            row = ab.amg_solve(device_options(ab, th), TOL, A, h_b, h_x, lane_ctx[w], phase_locks)  # x,b H2D; x,hist D2H
            with mlock:
                moved["h2d"] += h_b.nbytes + h_x.nbytes
                moved["d2h"] += h_x.nbytes + 8 * (row["niters"] + 1)
        fan_out(solve_host)
        A.close()
        e2e_bytes["h2d"], e2e_bytes["d2h"] = moved["h2d"], moved["d2h"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        for c in lane_ctx:
            c.reset_kernel_launches()
        # every lane's stream is idle here and again when fn() returns (each solve ends with a
        # synchronised read of its result), so the two events bracket the device work of all lanes
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            t_fn = time.perf_counter()
            fn()
            if os.environ.get("BENCH_DEBUG"):
                print(f"[bench] {fn.__name__}: {time.perf_counter() - t_fn:.3f} s wall", file=sys.stderr, flush=True)
        e1.record(stream)
        barrier()
        launches = sum(c.kernel_launches() for c in lane_ctx)
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, launches, clocks

    ms_step, launches, clocks = timed(sweep_device, args.steps, args.warmup, sample_clocks=True)
    _progress("device-resident sweeps timed")
    value = ms_step / 1e3 / (nsys * world)          # whole-job seconds per system
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    ms_e2e, _, _ = timed(sweep_e2e, e2e_steps, 1)
    _progress("host-buffer (e2e) sweeps timed")
    e2e_value = ms_e2e / 1e3 / (nsys * world)

    # ---- t(theta) of the device flavour: one sweep on ONE lane, host wall clock around the
    # synchronised calls (the quantity the reference's ANN learns, data_preprocessing.py:110)
    profile = {}
    for th in thetas:
        solve_device(0, th, profile)
    theta_profile = {"theta": [round(t, 2) for t in thetas],
                     "iters": [profile[t][2] for t in thetas],
                     "t_setup_ms": [round(1e3 * profile[t][0], 2) for t in thetas],
                     "t_solve_ms": [round(1e3 * profile[t][1], 2) for t in thetas],
                     "one_lane_s_per_system": round(sum(profile[t][0] + profile[t][1] for t in thetas) / nsys, 4)}

    # ---- per-kernel-family device times (CUDA events around every launch, on the
    # launching stream) over one more sweep on one lane
    ctx.enable_timers(True)
    ctx.reset_timers()
    for th in thetas:
        solve_device(0, th)
    fam = ctx.timers()
    # roofline of the dominant kernel on ONE launch set: the level-0 Jacobi half sweeps of the
    # theta = 0.25 system (per cycle: F rows, F rows, C rows), the same launches the ncu capture
    # behind `traffic` holds
    ctx.reset_timers()
    solve_device(0, 0.25)
    fam25 = ctx.timers()
    _progress("per-family timers collected")
    ctx.enable_timers(False)
    peak, peak_src = measured_peaks()
    kern = family_table(fam)
    d = fam25["smooth_l0"]
    ach = d["bytes"] / d["ms"] / 1e6
    roofline = {"bound": "hbm",
                "kernel": "sell_rows_kernel<1,SPLIT,EpiJacobi>: level-0 C/F Jacobi half sweeps of the theta=0.25 system "
                          "(2 F-row sweeps : 1 C-row sweep per V(1,1) cycle)",
                "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "peak_source": peak_src, "frac_of_nominal_8TBps": round(ach / 8000.0, 4),
                "traffic": measured_traffic(args.m),
                "avg_launch_ms": round(d["ms"] / d["launches"], 4), "launches": d["launches"],
                "algorithmic_bytes_per_launch": round(d["bytes"] / d["launches"]),
                "sweep_wide": {"achieved": round(fam["smooth_l0"]["bytes"] / fam["smooth_l0"]["ms"] / 1e6, 1),
                               "launches": fam["smooth_l0"]["launches"]},
                "families": kern}

    # ---- extras
    extras, part = {}, None
    want = args.extras
    if want == "auto":
        want = "configs" if world == 1 else "partitioned"
    # An extra must not take the headline down: it runs in a helper thread under a watchdog.  One that
    # does not come back in time is recorded as such, the remaining extras are skipped, the line is
    # printed and the process leaves through os._exit (a stuck CUDA call cannot be cancelled).
    hung = []

    def watched(name, fn, timeout_s):
        """(result, None) | (None, reason)"""
        if hung:
            return None, f"skipped: `{hung[0]}` did not return"
        box = {}

        def run():
            try:
                torch.cuda.set_device(local)
                torch.cuda.set_stream(stream)   # (torch's current stream is per thread: same one as the main thread)
                box["result"] = fn()
            except BaseException as e:  # noqa: BLE001
                box["error"] = f"{type(e).__name__}: {e}"[:300]
        th = threading.Thread(target=run, daemon=True)
        th.start()
        th.join(timeout_s)
        if th.is_alive():
            hung.append(name)
            print(f"[bench] extra `{name}` did not return within {timeout_s:.0f} s", file=sys.stderr, flush=True)
            return None, f"did not return within {timeout_s:.0f} s"
        if "error" in box:
            return None, box["error"]
        return box["result"], None

    def guarded(name, fn, timeout_s=300.0):
        if time.perf_counter() - t_start > args.extras_budget_s:
            extras[name] = {"skipped": f"wall-clock budget of {args.extras_budget_s} s for the whole run reached"}
            return
        _progress(f"extra `{name}` ...")
        res, why = watched(name, fn, timeout_s)
        if why is None:
            extras.update(res)
        else:
            extras[name] = {"failed": why}
    if world == 1 and want in ("configs", "all"):
        guarded("pooling", lambda: extra_pooling(ab, ctx, A_dev, n, nnz, peak), 120.0)
        guarded("config1_m100", lambda: extra_config1_and_pooling(ab, L, ctx, peak), 180.0)
        guarded("config4_m46", lambda: extra_config4(args.batch_systems, args.batch_threads), 420.0)
    # release the sweep's device memory before the large extras
    cpu = cpu_baseline_port(ab, args.m, nnz) if rank == 0 else None
    st0 = results[thetas[0]][1]
    iters = {f"{th:.2f}": results[th][0] for th in thetas}
    if not hung:
        A_dev.close()
        for c in lane_ctx[1:]:
            c.close()
        del d_rp, d_col, d_val, d_b, d_x0, d_x, lane_x, h_rp, h_col, h_val, h_xs
        torch.cuda.empty_cache()
    if world == 1 and want in ("configs", "all"):
        guarded("config5_m464_1gpu",
                lambda: {"config5_m464_1gpu": config5_threads_on_one_gpu(ab, args.part_m, 2, [0.25, 0.5])}, 420.0)
        if not hung:
            torch.cuda.empty_cache()
        guarded("config3_m187", lambda: extra_config3(ab, L, ctx), 420.0)
    if world > 1 and want in ("partitioned", "all"):
        # (on the calling thread, as measured on 2 and 8 GPUs: the block is collective, a watchdog on one rank
        # would not help the others)
        _progress("partitioned block ...")
        try:
            part = partitioned_block(ab, args, rank, world, local, stream, args.part_m)
        except Exception as e:  # noqa: BLE001
            part = {"failed": f"{type(e).__name__}: {e}"[:300]}

    line = None
    if rank == 0:
        line = {"metric": "AMG-PCG setup+solve seconds per system", "value": value, "unit": "s/system",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(args.m), "n": n, "nnz": nnz, "systems_per_step": nsys,
                           "per_gpu": "one matrix + full theta sweep per rank",
                           "systems_in_flight": lanes,
                           "phase_locks": phases,
                           "l2": "inputs (2.6 GB CSR) exceed the 126 MB L2; no flush needed",
                           "iters": iters,
                           "levels_theta0.05": [int(r) for r in st0["rows"]] if st0 else None,
                           "operator_complexity_theta0.05": st0["operator"] if st0 else None},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": "s/system", "steps": e2e_steps,
                        "h2d_bytes_per_step": e2e_bytes["h2d"], "d2h_bytes_per_step": e2e_bytes["d2h"]},
                "roofline": roofline, "cpu_baseline": cpu, "theta_profile": theta_profile}
        if extras:
            line["configs"] = extras
        if part is not None:
            line["partitioned"] = part
        line["wall_s"] = round(time.perf_counter() - t_start, 1)
    if line:
        print(json.dumps(line), flush=True)
    if hung:  # a helper thread still sits in a CUDA call: nothing below would return either
        os._exit(0)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------- config 5 alone
def run_partitioned(args):
    """--workload partitioned: only the config-5 block (see partitioned_block), printed as the line."""
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // max(1, int(os.environ.get("WORLD_SIZE", 1)))))
    import torch
    import torch.distributed as tdist
    import amg_ann_b200 as ab

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU implementation")
    torch.cuda.set_device(local)
    if world == 1:
        blk = config5_threads_on_one_gpu(ab, args.m, 2, [0.25, 0.5])
        print(json.dumps(dict(blk, metric="AMG-PCG setup+solve seconds per system", higher_is_better=False,
                              scaling="strong", dtype="f64", data="synthetic", vs_baseline=None)))
        return 0
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    tdist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    args.part_steps = args.steps
    blk = partitioned_block(ab, args, rank, world, local, stream, args.m)
    if rank == 0:
        print(json.dumps(dict(blk, higher_is_better=False, dtype="f64", data="synthetic", vs_baseline=None,
                              config={"workload": blk["workload"]})))
    tdist.barrier()
    tdist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--m", "--cells", dest="m", type=int, default=200, help="cells per direction (default: config 2); use --cells under torchrun")
    ap.add_argument("--cpu-m", type=int, default=100, help="largest mesh of the bounded CPU sample (reference arm)")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0, help="wall-clock budget of the reference arm's timed steps")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--streams", type=int, default=3,
                    help="independent systems of the sweep kept in flight per GPU (host threads, one stream each)")
    ap.add_argument("--reserve-gb", type=float, default=-1.0,
                    help="memory pool of every lane pre-grown to this size (default: 100 bytes per matrix entry)")
    ap.add_argument("--phases", type=int, default=0,
                    help="1: the systems in flight take turns per phase (one setup and one solve at a time); 0: free-running lanes")
    ap.add_argument("--workload", default="sweep", choices=["sweep", "partitioned"],
                    help="sweep: config 2 theta sweep, one system per GPU (default, the headline metric); "
                         "partitioned: config 5 only, one system row-partitioned over all GPUs (use --cells 464)")
    ap.add_argument("--extras", default="auto", choices=["auto", "none", "configs", "partitioned", "all"],
                    help="auto: N=1 adds the `configs` block, N>1 the `partitioned` block")
    ap.add_argument("--extras-budget-s", type=float, default=600.0,
                    help="no further extra is started once the whole run has taken this long")
    ap.add_argument("--part-m", type=int, default=464, help="mesh of the partitioned block (config 5)")
    ap.add_argument("--part-steps", type=int, default=2)
    ap.add_argument("--batch-systems", type=int, default=64, help="systems of the config-4 batch")
    ap.add_argument("--batch-threads", type=int, default=8)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "partitioned":
        return run_partitioned(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
