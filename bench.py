#!/usr/bin/env python
"""bench.py -- AMG-PCG theta-sweep hot path on B200 (contract: task prompt section 4).

Workload (BASELINE.json configs[1]): 3D diffusion, Q1 on [-1,1]^3, piecewise-
discontinuous coefficients (checkerboard, contrast 1e6), m=200 -> 8 120 601 DoFs,
217 081 801 nnz, theta sweep 0.05..0.95 step 0.05 (19 systems sharing one matrix,
ref testcase2-diffusion-structured/src/main.cpp:440-467), tol 1e-8 absolute.

One "step" = one theta sweep = 19 x (AMG setup + PCG solve) on the resident matrix.
metric = seconds per system (setup + solve), lower is better.

  value  : device-resident (CSR, rhs, x0 already in HBM), timed with CUDA events
  e2e    : through the reference-facing API with HOST buffers (matrix upload once
           per sweep, x/b H2D and x/residual-history D2H per system)
  N > 1  : one process per GPU, each rank sweeps its own system (weak scaling,
           no data-path collective); time = max over ranks
  --impl reference : the reference's own CPU flavour (Falgout + hybrid symmetric
           GS, the CPU restatement in oracle/; hypre itself is not installable
           here), one system per host thread, on a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

THETA = ("0.05", "0.96", "0.05")  # ref testcase2-diffusion-structured/datagen.py:47
TOL = 1e-8                        # ref datagen.py:9 (absolute)
PATTERN, MODE, CONTRAST = 4, 3, 6.0


def workload_name(m):
    n = (m + 1) ** 3
    return (f"3D diffusion Q1, checkerboard mu in {{1,1e6}} on a {PATTERN}^3 pattern, m={m} "
            f"({n} DoFs), theta sweep 0.05..0.95 step 0.05 (19 systems/matrix), "
            f"PMIS + classical interp + C/F l1-Jacobi V(1,1), PCG tol 1e-8 abs")


def make_system(ab, m, seed_shift=0):
    epsv = ab.gen.checkerboard_epsv(PATTERN, MODE, CONTRAST)
    if seed_shift:
        epsv = np.roll(epsv, seed_shift)
    return ab.gen.poisson_q1(m, PATTERN, MODE, epsv)


def device_options(ab, theta):
    R = ab.RelaxationType
    return ab.AdditionalData(True, theta, 0.9, 0, True, relaxation_type_up=R.l1scaledJacobi,
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


def measured_traffic(family, m):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture
    (profiles/traffic.json); only valid for the mesh it was captured on."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if m != 200 or not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f).get(family)
    return t["traffic_bytes_per_launch"] if t else None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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
    return t1 - t0, t2 - t1, nit, rc


def cpu_sample(m_sample, thetas, flavour, threads):
    """`threads` host threads, each solving its share of `thetas` on an m_sample mesh.
    Returns wall seconds per system (whole job) at the sample size."""
    import amg_ann_b200 as ab
    s = make_system(ab, m_sample)
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
    return wall / len(out), s, out


def run_reference(args):
    """--impl reference: CPU restatement of the reference flavour, all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import amg_ann_b200 as ab
    cores = os.cpu_count() or 1
    thetas = ab.gen.theta_sweep(*map(float, THETA))
    m_s = args.cpu_m
    n_full, nnz_full = ab.gen.sizes(0, args.m)
    _, nnz_s = ab.gen.sizes(0, m_s)
    scale = nnz_full / nnz_s
    per_step = []
    for step in range(args.warmup + args.steps):
        # one step = `cores` systems of the sweep in parallel (the reference fans
        # independent runs out over processes: 00_data-generation.py:105-116)
        sel = [thetas[(step * cores + i) % len(thetas)] for i in range(cores)]
        sec, s, _ = cpu_sample(m_s, sel, "reference", cores)
        if step >= args.warmup:
            per_step.append(sec)
    v = float(np.mean(per_step)) * scale
    sample = (f"m={m_s} ({(m_s + 1) ** 3} DoFs) instead of m={args.m}; {cores} systems per step, one "
              f"per host thread; seconds scaled by nnz ratio {scale:.1f} (AMG work is O(nnz)); "
              f"CPU restatement of hypre Falgout + symmetric GS (hypre itself not installable)")
    line = {"impl": "reference", "metric": "AMG-PCG setup+solve seconds per system",
            "value": v, "unit": "s/system", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(per_step)) * cores * 1e3,
            "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload_name(args.m)},
            "cpu_baseline": {"value": v, "unit": "s/system", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": v, "unit": "s/system", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import amg_ann_b200 as ab
    from amg_ann_b200._native import amgb_lib, c_f64p

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
    hist = np.zeros(4096)
    nit = C.c_int64()
    results = {}

    # The systems of a sweep are independent (the reference fans them out over processes,
    # 00_data-generation.py:105-116): `lanes` host threads, one context + stream each, take
    # them from a queue, so one system's latency-bound phases (coarse levels, PMIS rounds,
    # host round trips of PCG) overlap another's bandwidth-bound ones on the same GPU.
    lanes = max(1, args.streams)
    lane_streams = [stream] + [torch.cuda.Stream() for _ in range(lanes - 1)]
    lane_ctx = [ctx] + [ab.Context(local, st.cuda_stream) for st in lane_streams[1:]]
    lane_x = [d_x] + [torch.empty_like(d_x0) for _ in range(lanes - 1)]

    def solve_device(w, th):
        c = lane_ctx[w]
        h, k = np.zeros(4096), C.c_int64()
        with torch.cuda.stream(lane_streams[w]):
            lane_x[w].copy_(d_x0)                # solution = m_zero_solution (t2 main.cpp:446)
        P = ab.PreconditionBoomerAMG()
        P.initialize(A_dev, device_options(ab, th), c)
        rc = L.amgb_cg_solve_device(c._h, A_dev._h, C.c_void_p(lane_x[w].data_ptr()),
                                    C.c_void_p(d_b.data_ptr()), P._h, n, TOL,
                                    h.ctypes.data_as(c_f64p), len(h), C.byref(k))
        if rc != 0:
            raise RuntimeError(f"amgb_cg_solve_device -> {rc}: {L.amgb_last_error(c._h).decode()}")
        results[th] = (k.value, P.level_stats() if th == thetas[0] else None)
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

    def sweep_device_one_lane():
        for th in thetas:
            solve_device(0, th)

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
            row = ab.amg_solve(device_options(ab, th), TOL, A, h_b, h_x, lane_ctx[w])  # x,b H2D; x,hist D2H
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
            fn()
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
    value = ms_step / 1e3 / (nsys * world)          # whole-job seconds per system
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    ms_e2e, _, _ = timed(sweep_e2e, e2e_steps, 1)
    e2e_value = ms_e2e / 1e3 / (nsys * world)

    # ---- per-kernel-family device times (CUDA events around every launch, on the
    # launching stream) over one more sweep: roofline of the dominant kernel
    ctx.enable_timers(True)
    ctx.reset_timers()
    sweep_device_one_lane()
    fam = ctx.timers()
    ctx.enable_timers(False)
    peak, peak_src = measured_peaks()
    kern = {}
    total_ms = sum(v["ms"] for k, v in fam.items() if not k.endswith("_l0")) or 1.0
    for k, v in fam.items():
        if v["launches"] and v["ms"] > 0:
            kern[k] = {"ms": round(v["ms"], 3), "launches": v["launches"],
                       "GBps": round(v["bytes"] / v["ms"] / 1e6, 1),
                       "share": round(v["ms"] / total_ms, 4) if not k.endswith("_l0") else None}
    # dominant solve kernel: the level-0 Jacobi half sweeps (csr_rows_kernel<8,EpiJacobi>)
    dom = "smooth_l0" if "smooth_l0" in kern else max(kern, key=lambda k: kern[k]["ms"])
    d = fam[dom]
    ach = d["bytes"] / d["ms"] / 1e6
    roofline = {"bound": "hbm", "kernel": f"{dom} (sell_rows_kernel<1,SPLIT,EpiJacobi>, level-0 C/F half sweeps)",
                "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "peak_source": peak_src, "frac_of_nominal_8TBps": round(ach / 8000.0, 4),
                "traffic": measured_traffic(dom, args.m),
                "avg_launch_ms": round(d["ms"] / d["launches"], 4),
                "algorithmic_bytes_per_launch": round(d["bytes"] / d["launches"]),
                "families": kern}

    line = None
    if rank == 0:
        # ---- CPU baseline beside it: the oracle (same algorithm as the device path),
        # one core, bounded sample
        t0 = time.perf_counter()
        m_s = args.cpu_m
        _, nnz_s = ab.gen.sizes(0, m_s)
        scale = nnz / nnz_s
        sel = [0.25, 0.5] if args.cpu_m > 40 else thetas[::6]
        sec, _, detail = cpu_sample(m_s, sel, "port", 1)
        cpu = {"value": sec * scale, "unit": "s/system", "cores": 1, "kind": "port",
               "sample": (f"oracle (PMIS + classical + C/F l1-Jacobi, same options as the device) at "
                          f"m={m_s} ({(m_s + 1) ** 3} DoFs), theta in {sel}, 1 thread; measured "
                          f"{sec:.3f} s/system, scaled by nnz ratio {scale:.1f} to the m={args.m} system"),
               "measured_s_per_system_at_sample": sec,
               "wall_s": round(time.perf_counter() - t0, 2)}
        st = results[thetas[0]][1]
        line = {"metric": "AMG-PCG setup+solve seconds per system", "value": value, "unit": "s/system",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(args.m), "n": n, "nnz": nnz, "systems_per_step": nsys,
                           "per_gpu": "one matrix + full theta sweep per rank",
                           "systems_in_flight": lanes,
                           "l2": "inputs (2.6 GB CSR) exceed the 126 MB L2; no flush needed",
                           "iters": {f"{th:.2f}": results[th][0] for th in thetas},
                           "levels_theta0.05": [int(r) for r in st["rows"]] if st else None,
                           "operator_complexity_theta0.05": st["operator"] if st else None},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": "s/system", "steps": e2e_steps,
                        "h2d_bytes_per_step": e2e_bytes["h2d"], "d2h_bytes_per_step": e2e_bytes["d2h"]},
                "roofline": roofline, "cpu_baseline": cpu}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line:
        print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- config 5
def run_partitioned(args):
    """--workload partitioned: BASELINE config 5, ONE system row-partitioned over the N GPUs
    (z-slabs; halo exchange and dot-product reductions as put/flag kernels over NVLink peer
    windows, NCCL when AMGB_PEER=0).  step = setup + PCG solve for theta in
    {0.25, 0.5}; strong scaling (total work fixed)."""
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // max(1, int(os.environ.get("WORLD_SIZE", 1)))))
    import torch
    import torch.distributed as tdist
    import amg_ann_b200 as ab
    from amg_ann_b200 import dist
    from amg_ann_b200._native import amgb_lib, c_f64p

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU implementation")
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    tdist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = ab.Context(local, stream.cuda_stream)
    comm = dist.Communicator.nccl_from_torch(ctx)
    L = amgb_lib()
    m = args.m
    thetas = [0.25, 0.5]
    exch = "NCCL" if os.environ.get("AMGB_PEER", "1").startswith("0") else "NVLink peer-window (put/flag kernels)"
    starts = dist.slab_partition(m, world)
    b0, e0 = starts[rank], starts[rank + 1]
    sl = ab.gen.poisson_q1(m, row_begin=b0, row_end=e0)          # mu = 1: 3D Poisson
    A = dist.DistSparseMatrix(comm, sl.n, b0, e0, sl.rowptr, sl.col, sl.val)
    d_b = torch.from_numpy(sl.rhs).cuda()
    d_x0 = torch.from_numpy(sl.x0).cuda()
    d_x = torch.empty_like(d_x0)
    hist = np.zeros(4096)
    nit = C.c_int64()
    info = {}

    def step_device():
        for th in thetas:
            d_x.copy_(d_x0)
            P = dist.DistPreconditionBoomerAMG()
            P.initialize(A, device_options(ab, th))
            rc = L.amgb_dist_cg_solve_device(ctx._h, C.c_void_p(d_x.data_ptr()), C.c_void_p(d_b.data_ptr()), P._h,
                                             sl.n, TOL, hist.ctypes.data_as(c_f64p), len(hist), C.byref(nit))
            if rc != 0:
                raise RuntimeError(f"amgb_dist_cg_solve_device -> {rc}: {L.amgb_last_error(ctx._h).decode()}")
            info[th] = (nit.value, P.level_stats())
            P.close()

    h_x = sl.x0.copy()
    e2e_bytes = {}

    def step_e2e():
        A2 = dist.DistSparseMatrix(comm, sl.n, b0, e0, sl.rowptr, sl.col, sl.val)   # slab H2D
        h2d = sl.rowptr.nbytes // 2 + sl.col.nbytes + sl.val.nbytes
        d2h = 0
        for th in thetas:
            h_x[...] = sl.x0
            P = dist.DistPreconditionBoomerAMG()
            P.initialize(A2, device_options(ab, th))
            ctl = ab.SolverControl(sl.n, TOL)
            dist.DistSolverCG(ctl).solve(A2, h_x, sl.rhs, P)
            h2d += 2 * h_x.nbytes
            d2h += h_x.nbytes + 8 * (ctl.last_step() + 1)
            P.close()
        A2.close()
        e2e_bytes["h2d"], e2e_bytes["d2h"] = h2d, d2h

    def barrier():
        tdist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sample=False):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local) if sample else None
        if sampler:
            sampler.start()
        ctx.reset_kernel_launches()
        e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0_.record(stream)
        for _ in range(steps):
            fn()
        e1_.record(stream)
        barrier()
        launches = ctx.kernel_launches()
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([e0_.elapsed_time(e1_)], device="cuda", dtype=torch.float64)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item()) / steps, launches, clocks

    ms_step, launches, clocks = timed(step_device, args.steps, args.warmup, sample=True)
    ms_e2e, _, _ = timed(step_e2e, 1, 1)
    ctx.enable_timers(True)
    ctx.reset_timers()
    step_device()
    fam = ctx.timers()
    ctx.enable_timers(False)
    peak, peak_src = measured_peaks()
    kern = {k: {"ms": round(v["ms"], 3), "launches": v["launches"], "GBps": round(v["bytes"] / v["ms"] / 1e6, 1)}
            for k, v in fam.items() if v["launches"] and v["ms"] > 0}
    d = fam["smooth_l0"]
    ach = d["bytes"] / d["ms"] / 1e6
    if rank == 0:
        st = info[thetas[0]][1]
        line = {"metric": "AMG-PCG setup+solve seconds per system", "value": ms_step / 1e3 / len(thetas),
                "unit": "s/system", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"3D Poisson Q1, m={m} ({sl.n} DoFs, {int(st['nnz'][0])} nnz), ONE system "
                                       f"row-partitioned in z-slabs over {world} GPUs ({exch} halo exchange and "
                                       f"dot-product reductions), theta in {thetas}, PMIS + classical interp + C/F l1-Jacobi V(1,1), "
                                       f"PCG tol 1e-8 abs",
                           "n": sl.n, "nnz": int(st["nnz"][0]), "systems_per_step": len(thetas),
                           "l2": "per-GPU slab exceeds the 126 MB L2; no flush needed",
                           "iters": {f"{th:.2f}": info[th][0] for th in thetas},
                           "levels": [int(r) for r in st["rows"]], "operator_complexity": st["operator"]},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": ms_e2e / 1e3 / len(thetas), "unit": "s/system", "steps": 1,
                        "h2d_bytes_per_step": e2e_bytes["h2d"], "d2h_bytes_per_step": e2e_bytes["d2h"],
                        "note": "per rank"},
                "roofline": {"bound": "hbm", "kernel": "smooth_l0 (sell_rows_kernel<1,*,EpiJacobi>, level 0, rank 0)",
                             "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                             "peak_source": peak_src, "frac_of_nominal_8TBps": round(ach / 8000.0, 4),
                             "traffic": None, "families": kern},
                "cpu_baseline": None}
        print(json.dumps(line))
    A.close()
    comm.close()
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
    ap.add_argument("--cpu-m", type=int, default=56, help="mesh of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--streams", type=int, default=3,
                    help="independent systems of the sweep kept in flight per GPU (host threads, one stream each)")
    ap.add_argument("--workload", default="sweep", choices=["sweep", "partitioned"],
                    help="sweep: config 2 theta sweep, one system per GPU (default, the headline metric); "
                         "partitioned: config 5, one system row-partitioned over all GPUs (use --cells 464)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "partitioned":
        return run_partitioned(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
