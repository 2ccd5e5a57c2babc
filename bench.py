#!/usr/bin/env python
"""bench.py -- SBCGrQ time-to-solution on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    torchrun ... bench.py --gpus N ...            (one rank per GPU, slab decomposition)

A *step* is one complete solve of the workload (default = configs[2] of BASELINE.json, the
configuration the metric is quoted on: SBCGrQ, 24^4 sites, N=12 right-hand sides, mass 1e-3,
tol 1e-10, the nine benchmark shifts of benchmark.cpp:12-13) on synthetic inputs (links and
sources uniform in [-1,1]+i[-1,1], the reference's distribution, inc/dirac_op.hpp:27-32).

 value     time-to-solution with inputs resident in HBM (CUDA events around setup + loop, taken
           inside the library on its own stream), mean over the K timed steps, max over ranks.
 e2e       the same solves measured through the host-buffer C-ABI call a reference user makes
           (bcg_solve_sbcgrq <- SBCGrQ<N>): pinned host B in, nine host X out, wall clock around
           the call, copies inside the timed region.
 parity    evidence taken in THIS run: K=4 iterations of the GPU loop against 4 iterations of the
           unmodified reference on the same inputs (per-shift relative difference), and the TRUE
           residual |B - (A + sigma) X| / |B| of every shift of the timed solve (benchmark.cpp:93-103).
           At N > 1 GPUs: the slab-decomposed loop against the single-domain loop on the same inputs.
 loop      what the timed loop really did: histogram of active systems per iteration (shifts retire
           at eps_shifts), bytes moved, unfused algorithmic bytes, both / value; the
           per-stage device times measured INSIDE the loop (CUDA events, first iterations of the
           last warm-up solve) and the iteration time predicted from the histogram beside the measured one.
 roofline  the dominant kernel (multishift update): histogram-weighted average launch, measured live
           with CUDA events (bcg_bench_kernel at every active count); algorithmic bytes per SURVEY 8(d).
 cpu_baseline  the unmodified reference (oracle/_ref) timed on this box's host cores on a bounded
           sample (seconds per iteration with all shifts and with one, from K=1 and K=4 runs),
           weighted with the active-system histogram -- an extrapolation, labelled as such.

--impl reference runs the reference's own CPU implementation (oracle/_ref, else the oracle port)
on the same workload: the same bounded sample per step, extrapolated the same way.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

# stdout carries the ONE JSON line and nothing else: libraries that chat on file descriptor 1 (NCCL prints its
# version banner there) are sent to stderr, the JSON line goes to a private copy of the original stdout
JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BENCH_SHIFTS = [0, 0, 1e-10, 1e-8, 1e-6, 1e-5, 1e-4, 1e-2, 1e-1]  # benchmark.cpp:12-13


def _sb(V, N=12):
    return dict(solver="sbcgrq", V=V, N=N, mass=1e-3, eps=1e-10, eps_shifts=1e-15, shifts=BENCH_SHIFTS)


def _bq(V, N):
    return dict(solver="bcgrq", V=V, N=N, mass=1e-3, eps=1e-10, eps_shifts=0.0, shifts=[0.0])


WORKLOADS = {
    # configs[2] of BASELINE.json: the configuration the metric is quoted on
    "sbcgrq_24^4_N12": _sb(24 ** 4),
    # configs[0]: the reference's README default (CPU-runnable: the reference arm is a FULL solve, no extrapolation)
    "sbcgrq_V1000_N12": dict(_sb(1000), cpu_full=True),
    "sbcgrq_16^4_N12": _sb(16 ** 4),
    # configs[1]: BCGrQ single-shift block solve, 16^4 sites, N = 4 / 8 / 12
    "bcgrq_16^4_N4": _bq(16 ** 4, 4),
    "bcgrq_16^4_N8": _bq(16 ** 4, 8),
    "bcgrq_16^4_N12": _bq(16 ** 4, 12),
    # configs[3]: 48^3 x 96 sites on 2 / 4 / 8 GPUs (64 / 32 / 16 GB of fields per GPU); a full solve takes
    # ~25 000 iterations of ~40 ms / n_gpus, so this one is meant to be run with --max-it
    "sbcgrq_48^3x96_N12": _sb(48 ** 3 * 96),
}
ITER_FILE = os.path.join(ROOT, "profiles", "bench_iterations.json")
CLOCK_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
               "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
               "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
REF_BUILD = ("oracle/_ref: the unmodified reference headers compiled by oracle/Makefile with the reference's Release "
             "flags (-O3 -DEIGEN_NO_DEBUG) except -march=x86-64-v3 (AVX2+FMA, portable between this container and the "
             "GPU box) in place of -march=native; 1 thread (the reference is single-threaded)")


def make_inputs(V, N, seed=1):
    rng = np.random.default_rng(seed)
    U = rng.uniform(-1, 1, (V, 3, 3)) + 1j * rng.uniform(-1, 1, (V, 3, 3))
    B = rng.uniform(-1, 1, (V, N, 3)) + 1j * rng.uniform(-1, 1, (V, N, 3))
    return U, B


def make_inputs_slab(V, N, rank, world, chunks=8):
    """Large volumes: the global problem is defined as `chunks` independently seeded pieces, each rank
    generates only the pieces of its own slab (same global inputs for every world size dividing `chunks`)."""
    assert chunks % world == 0 and V % chunks == 0
    per, Vc = chunks // world, V // chunks
    Us, Bs = [], []
    for c in range(rank * per, (rank + 1) * per):
        u, b = make_inputs(Vc, N, seed=1000 + c)
        Us.append(u)
        Bs.append(b)
    return np.concatenate(Us), np.concatenate(Bs)


class ClockSampler:
    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = "/tmp/bcg_clocks_%d_%d.csv" % (os.getpid(), device)

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + CLOCK_QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def recorded(workload):
    """Iteration count and active-system histogram of the last recorded GPU-arm run of this workload."""
    try:
        d = json.load(open(ITER_FILE)).get(workload)
    except Exception:
        return None
    if isinstance(d, int):
        return {"iterations": d, "active_hist": None}
    return d


# ---- the reference on host cores (checker side: oracle/_ref, else the oracle port) -------------------
def reference_run(w, U, B, shifts, max_it):
    """The reference's own solver for `max_it` iterations.  Returns (X [S][V][N][3], iterations, seconds, kind)."""
    from oracle.pyoracle import Oracle, RefShim, build as build_oracle
    if RefShim.available(w["N"]):
        r = RefShim(w["N"])
        if w["solver"] == "bcgrq":
            X, it, sec = r.BCGrQ(U, B, w["mass"], w["eps"], max_it=max_it)
            return X[None], it, sec, "reference"
        X, it, sec = r.SBCGrQ(U, B, w["mass"], shifts, w["eps"], w["eps_shifts"], max_it=max_it)
        return X, it, sec, "reference"
    build_oracle()
    o = Oracle()
    if w["solver"] == "bcgrq":
        X, it, sec = o.BCGrQ(U, B, w["mass"], w["eps"], max_it=max_it)
        return X[None], it, sec, "port"
    X, it, sec, _ = o.SBCGrQ(U, B, w["mass"], shifts, w["eps"], w["eps_shifts"], max_it=max_it)
    return X, it, sec, "port"


def reference_sample(w, U, B, K):
    """Seconds per iteration of the reference with all S systems active and with one, each from the
    difference of a K-iteration and a 1-iteration run (so the set-up -- thinQR of B, S copies of Q --
    is not charged to the iterations).  Returns dict + the K-iteration solutions (lock-step parity)."""
    S = len(w["shifts"])
    X_K, itK, tK, kind = reference_run(w, U, B, w["shifts"], K)
    _, it1, t1, _ = reference_run(w, U, B, w["shifts"], 1)
    spi_all = (tK - t1) / max(itK - it1, 1)
    out = {"kind": kind, "cores": 1, "K": K, "s_per_iteration_all_systems": spi_all, "setup_s": max(t1 - spi_all, 0.0),
           "sample_cpu_s": tK + t1}
    if S > 1:
        _, itK1, tK1, _ = reference_run(w, U, B, w["shifts"][:1], K)
        _, it11, t11, _ = reference_run(w, U, B, w["shifts"][:1], 1)
        out["s_per_iteration_one_system"] = (tK1 - t11) / max(itK1 - it11, 1)
        out["sample_cpu_s"] += tK1 + t11
    else:
        out["s_per_iteration_one_system"] = spi_all
    return out, X_K, itK


def extrapolate_cpu(sample, S, iters, hist):
    """Reference time-to-solution = set-up + sum over iterations of t(a), a = systems active in that iteration,
    t(a) interpolated linearly between the measured one-system and all-systems iteration (the reference's
    cost is 19F + U_b + 6F (a-1) bytes per iteration, SURVEY 8a: affine in a)."""
    t1, tS = sample["s_per_iteration_one_system"], sample["s_per_iteration_all_systems"]

    def t_of(a):
        return tS if S <= 1 else t1 + (tS - t1) * (a - 1) / (S - 1)
    if hist and sum(hist) > 0:
        total = sum(n * t_of(max(a, 1)) for a, n in enumerate(hist) if n)
        how = "weighted with the GPU run's active-system histogram"
    else:
        total = iters * tS
        how = "ALL systems assumed active in every iteration (no histogram recorded: an upper bound)"
    return sample["setup_s"] + total, how


def run_reference(args, w, wname):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    U, B = make_inputs(w["V"], w["N"])
    S = len(w["shifts"])
    rec = recorded(wname) or {}
    times, last = [], None
    if w.get("cpu_full"):
        for i in range(args.warmup + args.steps):
            _, it, sec, kind = reference_run(w, U, B, w["shifts"], 1000000)
            if i >= args.warmup:
                times.append(sec)
        value = statistics.mean(times)
        sample = ("FULL reference %s<%d> solve per step on the same inputs (%d iterations, no extrapolation); %s"
                  % (w["solver"].upper().replace("RQ", "rQ"), w["N"], it, REF_BUILD))
        extra = {"extrapolated": False, "iterations": it}
        cores = 1
    else:
        for i in range(args.warmup + args.steps):
            if i < args.warmup:  # warm-up step: one iteration (pages the library and the inputs in; a CPU has no clocks to ramp)
                reference_run(w, U, B, w["shifts"], 1)
                continue
            smp, _, _ = reference_sample(w, U, B, args.cpu_iters)
            if i >= args.warmup:
                v, how = extrapolate_cpu(smp, S, rec.get("iterations") or args.cpu_iters, rec.get("active_hist"))
                times.append(v)
                last = smp
        value = statistics.mean(times)
        kind, cores = last["kind"], last["cores"]
        sample = ("per step: the reference's own solver for %d and for 1 iteration(s) on the same inputs, with all %d "
                  "systems and with one (%.1f s of CPU work): %.3f / %.3f s per iteration; extrapolated to the GPU arm's "
                  "%s iterations (profiles/bench_iterations.json), %s; %s"
                  % (args.cpu_iters, S, last["sample_cpu_s"], last["s_per_iteration_all_systems"],
                     last["s_per_iteration_one_system"], rec.get("iterations"), how, REF_BUILD))
        extra = {"extrapolated": True, "s_per_iteration_all_systems": last["s_per_iteration_all_systems"],
                 "s_per_iteration_one_system": last["s_per_iteration_one_system"], "iterations_assumed": rec.get("iterations"),
                 "note": "the reference itself needs a few % MORE iterations than the GPU loop (sequential Gram sums, "
                         "SURVEY F7b); the GPU count is used, which favours the reference"}
    line = {"impl": "reference", "metric": "sbcgrq_time_to_solution", "value": value, "unit": "s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * value,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_of(wname, w, args.gpus),
            "cpu_baseline": {"value": value, "unit": "s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line.update(extra)
    print(json.dumps(line), file=JSON_OUT, flush=True)


EXCHANGE = ("halo sites and Gram blocks exchanged by P2P stores over NVLink from inside the kernels "
            "(NCCL only in the set-up)")


def config_of(wname, w, gpus):
    nf = 2 * len(w["shifts"]) + 3   # X_s, P_s, Q, Q', T (the deep deferral keeps up to two more Q fields)
    return {"workload": wname, "solver": w["solver"], "V": w["V"], "n_rhs": w["N"], "n_shifts": len(w["shifts"]),
            "mass": w["mass"], "eps": w["eps"], "eps_shifts": w["eps_shifts"],
            "operator": "reference 1-D chain (inc/dirac_op.hpp:13-21)",
            "partition": "1 slab" if gpus == 1 else "%d contiguous site slabs; %s" % (gpus, EXCHANGE),
            "l2_policy": ("at least %d fields of %.0f MB each per GPU stream through every group of 2-4 iterations (working set %s"
                          " the 126 MB L2); no explicit flush"
                          % (nf, 48.0 * w["N"] * w["V"] / gpus / 1e6,
                             "exceeds" if nf * 48.0 * w["N"] * w["V"] / gpus > 126e6 else "FITS IN"))}


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def run_ours(args, w, wname):
    global EXCHANGE
    import torch

    import blockcg_b200
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torchrun --nproc-per-node %d" % (args.gpus, world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    V, N, S = w["V"], w["N"], len(w["shifts"])
    block = w["solver"] == "bcgrq"
    if V % world:
        raise SystemExit("V=%d not divisible by %d ranks" % (V, world))
    Vl = V // world
    big = V > 2 ** 21
    if big:
        Ul, Bl = make_inputs_slab(V, N, rank, world)
        U = B = None
    else:
        U, B = make_inputs(V, N)
        Ul = np.ascontiguousarray(U[rank * Vl:(rank + 1) * Vl])
        Bl = B[rank * Vl:(rank + 1) * Vl]
    # pinned host buffers for the end-to-end path
    Bh = torch.empty((Vl, N, 3), dtype=torch.complex128).pin_memory()
    Bh.numpy()[...] = Bl
    del Bl
    Xh = [torch.empty((Vl, N, 3), dtype=torch.complex128).pin_memory() for _ in range(S)]
    Bn, Xn = Bh.numpy(), [x.numpy() for x in Xh]

    ctx = blockcg_b200.Context(Vl, N, max_shifts=S, device=local, rank=rank, nranks=world)
    if world > 1:
        uid = torch.zeros(blockcg_b200.capi.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(ctx.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()))
        if not args.no_p2p:
            from blockcg_b200.distributed import exchange_ipc_handles
            p2p_on = exchange_ipc_handles(dist, ctx, torch.device("cuda", local))
            if not p2p_on:
                EXCHANGE = "NCCL send/recv halo + all-reduce of the Gram blocks in the loop (peer mapping unavailable)"
        else:
            EXCHANGE = "NCCL send/recv halo + all-reduce of the Gram blocks in the loop (--no-p2p)"
    ctx.set_links(Ul, w["mass"])

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def solve(max_it):
        if block:
            return ctx.solve_bcgrq(Xn[0], Bn, w["eps"], max_it)
        return ctx.solve_sbcgrq(Xn, Bn, w["shifts"], w["eps"], w["eps_shifts"], max_it)

    def step():
        t0 = time.perf_counter()
        info = solve(args.max_it)
        t1 = time.perf_counter()
        return info, t1 - t0

    profile = None
    for i in range(args.warmup):
        if i == args.warmup - 1 and args.profile_iters > 0:
            ctx.set_loop_profile(args.profile_iters, args.profile_after)  # stage-by-stage device times inside the loop (untimed step)
            step()
            profile = ctx.loop_profile()
        else:
            step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t_begin = time.perf_counter()
    dev_ms, wall_s, launches, iters, resid = [], [], 0, None, None
    for _ in range(args.steps):
        info, dt = step()
        dev_ms.append(info.setup_ms + info.solve_ms)
        wall_s.append(dt)
        launches += info.kernel_launches
        iters, resid = info.iterations, info.residual
    barrier()
    t_end = time.perf_counter()
    clocks = sampler.stop()
    stats = ctx.last_solve_stats()

    dev_s = statistics.mean(dev_ms) / 1e3
    e2e_s = statistics.mean(wall_s)
    region_s = (t_end - t_begin) / args.steps
    if dist is not None:
        t = torch.tensor([dev_s, e2e_s, region_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s, region_s = t.tolist()
        ln = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(ln)
        launches = int(ln.item())

    # ---- parity evidence, part 1: TRUE residuals of the timed solve (device verification path, all ranks) ----
    hb = ctx.field(Bn)
    hx = ctx.field()
    true_res = []
    for s in range(S):
        ctx.upload(hx, Xn[s])
        true_res.append(float(ctx.true_residual(hx, hb, w["shifts"][s]).max()))
    ctx.free(hx)
    parity = {"true_residual": true_res,
              "true_residual_rule": "max over right-hand sides of |B - (A + sigma_s) X_s| / |B| per shift, computed on the "
                                    "device as benchmark.cpp:93-103 does; the reference's own test gate is 2*eps at kappa~10 "
                                    "(test/solvers.cpp:116); at this kappa~1e7 the reference itself reaches 3.1e-10 at V=1000 "
                                    "(tests/golden/bench_V1000_N12.npz) and 1.6e-10 at V=8^4 (profiles/r01_parity_full_solve_8x4.json)",
              "recurrence_residual": resid}

    # ---- part 2 (N > 1): the slab-decomposed loop against the single-domain loop, same inputs, same iterations ----
    if world > 1 and not big:
        K = args.multi_lockstep_iters
        Xconv = [x.copy() for x in Xn[:S]] if args.max_it >= 1000000 else None   # the timed (converged) slab solutions
        solve(K)
        Xslab = [x.copy() for x in Xn[:S]]
        ho = ctx.field()
        ctx.op(ho, hb, w["shifts"][0])
        op_slab = ctx.download(ho)
        ctx.free(ho)
        with blockcg_b200.Context(V, N, max_shifts=S, device=local) as c1:  # every rank: the whole lattice on its own GPU
            c1.set_links(U, w["mass"])
            h1 = c1.field(B)
            xs1 = [c1.field() for _ in range(S)]
            if block:
                c1.solve_bcgrq_dev(xs1[0], h1, w["eps"], K)
            else:
                c1.solve_sbcgrq_dev(xs1, h1, w["shifts"], w["eps"], w["eps_shifts"], K)
            sl = slice(rank * Vl, (rank + 1) * Vl)
            num = [float(np.abs(c1.download(xs1[s])[sl] - Xslab[s]).max()) for s in range(S)]
            den = [float(np.abs(c1.download(xs1[s])).max()) for s in range(S)]
            conv = None
            if Xconv is not None:   # ... and the CONVERGED solutions: the whole solve again on one GPU (every rank its own copy)
                if block:
                    i1 = c1.solve_bcgrq_dev(xs1[0], h1, w["eps"])
                else:
                    i1 = c1.solve_sbcgrq_dev(xs1, h1, w["shifts"], w["eps"], w["eps_shifts"])
                cnum = [float(np.abs(c1.download(xs1[s])[sl] - Xconv[s]).max()) for s in range(S)]
                cden = [float(np.abs(c1.download(xs1[s])).max()) for s in range(S)]
                conv = (cnum, cden, i1.iterations)
            h2 = c1.field()
            c1.op(h2, h1, w["shifts"][0])
            o1 = c1.download(h2)
            op_num, op_den = float(np.abs(o1[sl] - op_slab).max()), float(np.abs(o1).max())
        t = torch.tensor(num + [op_num] + (conv[0] if conv else []), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        num = t.tolist()
        mg = {"lockstep_iterations": K, "lockstep_x_rel_vs_single_domain": [n_ / d_ for n_, d_ in zip(num[:S], den)],
              "op_rel_vs_single_domain": num[S] / op_den, "tol": 1e-9,
              "note": "same inputs: the block Dirac apply, %d iterations of the %d-slab loop (peer-memory exchange) and the "
                      "converged solutions of the timed solve, each against the single-domain loop on one GPU; max over "
                      "ranks.  (The Gram sums are grouped by slab, so un-converged iterates of this kappa ~ 1e7 problem "
                      "drift apart at the rounding level and the drift is amplified until convergence -- 1e-3 after 200 "
                      "iterations, measured -- which is why the lock-step window is short; the converged solutions agree.)"
                      % (K, world)}
        ok = max(mg["lockstep_x_rel_vs_single_domain"]) < 1e-9 and mg["op_rel_vs_single_domain"] < 1e-13
        if conv:
            mg["converged_x_rel_vs_single_domain"] = [n_ / d_ for n_, d_ in zip(num[S + 1:], conv[1])]
            mg["iterations_slabs"], mg["iterations_single_domain"] = iters, conv[2]
            ok = ok and max(mg["converged_x_rel_vs_single_domain"]) < 1e-9
        mg["pass"] = bool(ok)
        parity["multi_gpu"] = mg

    # ---- weak-scaling probe: the same sites PER GPU as the 1-GPU headline run, capped iterations (extra key) ----
    weak = None
    if args.weak_iters > 0 and not big and w["solver"] == "sbcgrq":
        Vw = V   # per GPU; the global chain has world * V sites, inputs drawn on the device (counter-based generator)
        cw = blockcg_b200.Context(Vw, N, max_shifts=S, device=local, rank=rank, nranks=world)
        try:
            if world > 1:
                uid = torch.zeros(blockcg_b200.capi.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
                if rank == 0:
                    uid.copy_(torch.frombuffer(bytearray(cw.unique_id()), dtype=torch.uint8))
                dist.broadcast(uid, 0)
                cw.comm_init(bytes(uid.cpu().numpy().tobytes()))
                if not args.no_p2p:
                    from blockcg_b200.distributed import exchange_ipc_handles
                    exchange_ipc_handles(dist, cw, torch.device("cuda", local))
            cw.set_links_random(7, w["mass"])
            hbw = cw.field()
            cw.field_random(hbw, 8)
            xsw = [cw.field() for _ in range(S)]
            cw.solve_sbcgrq_dev(xsw, hbw, w["shifts"], w["eps"], w["eps_shifts"], args.weak_iters)   # warm-up (graph capture)
            barrier()
            iw = cw.solve_sbcgrq_dev(xsw, hbw, w["shifts"], w["eps"], w["eps_shifts"], args.weak_iters)
            tw = torch.tensor([iw.solve_ms], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            weak = {"sites_per_gpu": Vw, "global_sites": Vw * world, "iterations": iw.iterations,
                    "ms_per_iteration": float(tw.item()) / max(iw.iterations, 1),
                    "note": "weak scaling: the headline volume PER GPU (global chain of n_gpus x that), device-generated inputs, "
                            "first %d iterations (all %d systems active), max over ranks; compare ms_per_iteration across n_gpus"
                            % (args.weak_iters, S)}
        finally:
            cw.close()

    # ---- per-kernel roofline, measured live with CUDA events on same-size fields ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    F, Ub = 48.0 * N * Vl, 144.0 * Vl
    hs = [hb, ctx.field(Bn)] + [ctx.field() for _ in range(2 * S - 1)]
    for h in hs[2:]:
        ctx.copy(h, hs[0])
    kern = {}
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(wname, {}) if world == 1 else {}
    except Exception:
        pass
    sched, depth = (stats["schedule"], stats["depth"]) if S > 1 else (0, 1)  # what the timed solve ran (capi.cu: pair_default)
    paired = sched != 0

    def bench(name, which, nh, ns, nbytes, per_rep=1, reps=20):
        ms, _ = ctx.bench_kernel(which, max(reps // per_rep, 2), hs[:nh], ns)
        ms /= per_rep
        kern[name] = {"ms": ms, "alg_bytes": nbytes, "achieved": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak,
                      "traffic": traffic.get(name)}
        return ms
    bench("dirac_gram", 0, 2, 1, 2 * F + Ub)
    bench("dirac", 1, 2, 1, 2 * F + Ub)
    bench("axpy_gram", 3, 2, 1, 3 * F)
    # The multishift update at every active-system count a = 1..S.  Numerator: the fixed per-unit figure of
    # SURVEY 8(d) x the units a launch processes: one back-substitution (2 F) and a system updates (4 F each).
    # "shift_pair[a]" is the AVERAGE launch of an odd + even pair (the loop serves the shifted systems every
    # second iteration, shift_pair.cuh); "shift_update[a]" the plain kernel (every system every iteration).
    dmma = os.environ.get("BCG_DMMA", "1") != "0" and N % 4 == 0   # shift_dmma.cuh: the same update on the FP64 tensor instruction
    upd = ("shift_dmma_pair" if paired else "shift_dmma") if dmma else ("shift_pair" if paired else "shift_update")
    if sched == 3:
        upd = "shift_stag%d" % depth   # shift_stag.cuh: the shifted systems every depth-th iteration
    plain = "shift_dmma" if dmma else "shift_update"
    t_upd = {}
    for a in range(1, S + 1):
        if sched == 3:   # one repetition = `depth` consecutive launches (every group of systems served once)
            t_upd[a] = bench("%s[%d]" % (upd, a), 14, 1 + 2 * a, a, (2 + 4 * a) * F, depth, 12)
        elif paired:
            t_upd[a] = bench("%s[%d]" % (upd, a), 13, 1 + 2 * a, a, (2 + 4 * a) * F, 2, 12)
        else:
            t_upd[a] = bench("%s[%d]" % (upd, a), 4, 1 + 2 * a, a, (2 + 4 * a) * F, 1, 12)
    if paired:
        bench("%s[%d]" % (plain, S), 4, 1 + 2 * S, S, (2 + 4 * S) * F)  # every system every iteration: for comparison
    for h in hs[1:]:
        ctx.free(h)
    ctx.free(hb)

    # ---- what the timed loop did ----
    hist = stats["active_hist"]
    n_it = max(sum(hist), 1)
    mean_active = sum(a * n for a, n in enumerate(hist)) / n_it
    moved = iters * (2 * F + Ub + 3 * F) + stats["shift_update_field_passes"] * F       # K1 + K3 + multishift
    unfused = sum(n * (19 * F + Ub + 6 * F * (max(a, 1) - 1)) for a, n in enumerate(hist))
    loop_s = dev_s  # set-up (thinQR of B, S copies) is < 0.1 % of a full solve
    w_ms = sum(n * t_upd[max(min(a, S), 1)] for a, n in enumerate(hist)) / n_it   # histogram-weighted update launch
    w_bytes = sum(n * (2 + 4 * max(a, 1)) * F for a, n in enumerate(hist)) / n_it
    chain_ms = None
    predicted = None
    upd_scale = 1.0
    if profile and profile["iterations"] > 0:
        pm = profile["ms"]
        chain_ms = pm["step_a"] + pm["step_b"] + pm["halo"]
        # field kernels as measured inside the loop where the window has them (stencil, Q update); the
        # multishift update per active count from the micro-benchmark, scaled by (in-loop / micro) at a = S
        inloop_upd = 0.5 * (pm["shift_odd"] + pm["shift_even"]) if pm["shift_even"] > 0 else pm["shift_odd"]
        a_w = max(min(profile["active_systems"], S), 1)   # systems active in the profiled window
        upd_scale = inloop_upd / t_upd[a_w] if t_upd[a_w] > 0 else 1.0
        predicted = pm["dirac_gram"] + pm["axpy_gram"] + chain_ms + upd_scale * w_ms
    loop = {"active_hist": {str(a): n for a, n in enumerate(hist) if n}, "mean_active_systems": mean_active,
            "moved_bytes_per_solve": moved, "moved_GBps": moved / loop_s / 1e9, "moved_frac_of_hbm_peak": moved / loop_s / 1e9 / peak,
            "unfused_algorithmic_bytes_per_solve": unfused, "unfused_GBps": unfused / loop_s / 1e9,
            "unfused_frac_of_hbm_peak": unfused / loop_s / 1e9 / peak,
            "bytes_note": "moved = what the kernels read + write (stencil 2F+U_b, Q update 3F, multishift update as counted "
                          "on the device per launch); unfused = the reference's own 19F + U_b + 6F(a-1) per iteration "
                          "(SURVEY 8a), a = active systems; both summed over the iterations of the timed solve, / value",
            "in_loop_profile": profile,
            "microbench_ms": {"dirac_gram": kern["dirac_gram"]["ms"], "axpy_gram": kern["axpy_gram"]["ms"],
                              upd + "[a]": {str(a): t_upd[a] for a in t_upd}},
            "predicted_ms_per_iteration": predicted, "measured_ms_per_iteration": 1e3 * dev_s / max(iters, 1),
            "prediction": "stage times measured inside the loop (window of --profile-iters iterations after --profile-after, "
                          "sustained clocks) with the multishift update re-weighted by the histogram: sum_a hist[a] * "
                          "t_update(a) / iterations, t_update(a) from the micro-benchmark scaled by in-loop / micro-benchmark "
                          "at the window's active count"}
    if predicted:
        loop["predicted_over_measured"] = predicted / loop["measured_ms_per_iteration"]
    roofline = {"kernel": upd + " (multishift update), histogram-weighted average launch of the timed loop",
                "bound": "hbm", "achieved": w_bytes / w_ms / 1e6, "peak": peak, "unit": "GB/s",
                "frac": w_bytes / w_ms / 1e6 / peak, "traffic": (traffic.get(upd) if S > 1 else traffic.get("shift_update")),
                "peak_source": peak_src, "alg_bytes_per_launch": w_bytes, "ms_per_launch": w_ms,
                "moved_bytes_per_launch": stats["shift_update_field_passes"] * F / max(iters, 1),
                "moved_GBps": stats["shift_update_field_passes"] * F / max(iters, 1) / w_ms / 1e6,
                "share_of_iteration": (upd_scale * w_ms / loop["measured_ms_per_iteration"]) if predicted else None,
                "all_systems_active": kern["%s[%d]" % (upd, S)],
                "note": "numerator = fixed per-unit bytes of SURVEY 8(d) (2F per back-substitution + 4F per system update) x "
                        "units per launch, so fusing / pairing shows as bandwidth; moved_* = bytes the launches really moved "
                        "(device-side count); traffic = ncu dram bytes of an all-systems-active launch (profiles/)"}
    # The update kernel is bound by the FP64 (tensor-instruction) pipe rather than by HBM since the deferred schedules
    # (ncu: profiles/r02_ncu_full_top3_depth4.txt), so the same launch is also put against the FP64 peak measured on
    # B200 by tools/micro/fp64_pipes.cu (DFMA 35.5, DMMA 36.8 TFLOP/s on the same pipe).  Real flops of the complex
    # arithmetic: a system update is two N x N products per site and colour (2 * 3 * N^2 complex multiply-adds of
    # 8 flops), the back-substitution 3 * N(N+1)/2 of them.
    try:
        w_act = sum(n * max(a, 1) for a, n in enumerate(hist)) / n_it
        flops_launch = 8.0 * Vl * (3.0 * N * (N + 1) / 2 + 6.0 * N * N * w_act)
        fp64_peak = 35.5
        roofline["fp64"] = {"achieved": flops_launch / w_ms / 1e9, "peak": fp64_peak, "unit": "TFLOP/s",
                            "frac": flops_launch / w_ms / 1e9 / fp64_peak,
                            "peak_source": "tools/micro/fp64_pipes.cu on B200 (DFMA issue rate; DMMA shares the pipe), DESIGN.md 4",
                            "note": "complex-arithmetic flops of the histogram-weighted launch / its time; the kernel's ncu "
                                    "figures with all systems active: DMMA pipe 63 %, DRAM 42 %"}
    except Exception:   # explanatory key only: never let it take the bench line down
        pass
    dirac = kern["dirac_gram"]

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- parity evidence, part 3 + CPU baseline: the reference on this box's host cores (rank 0, N = 1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline and not big:
        if w.get("cpu_full"):
            Xr, it_cpu, sec, kind = reference_run(w, U, B, w["shifts"], 1000000)
            cpu = {"value": sec, "unit": "s", "cores": 1, "kind": kind,
                   "sample": "FULL reference solve on the same inputs (%d iterations, no extrapolation); %s" % (it_cpu, REF_BUILD)}
            solve(1000000)
            parity["full_solve_vs_reference"] = {"iterations_gpu": iters, "iterations_reference": it_cpu,
                                                 "x_rel": [rel(Xn[s], Xr[s]) for s in range(S)], "tol": 1e-9}
        else:
            K = args.cpu_iters
            smp, Xr, itK = reference_sample(w, U, B, K)
            v, how = extrapolate_cpu(smp, S, iters, hist)
            cpu = {"value": v, "unit": "s", "cores": smp["cores"], "kind": smp["kind"], "extrapolated": True,
                   "s_per_iteration_all_systems": smp["s_per_iteration_all_systems"],
                   "s_per_iteration_one_system": smp["s_per_iteration_one_system"],
                   "sample": "the reference's own solver for %d and for 1 iteration(s) on the same inputs, with all %d systems "
                             "and with one (%.1f s of CPU work): %.3f / %.3f s per iteration; x the %d GPU iterations, %s "
                             "(the reference itself needs a few %% more iterations, SURVEY F7b); host has %d cores; %s"
                             % (K, S, smp["sample_cpu_s"], smp["s_per_iteration_all_systems"],
                                smp["s_per_iteration_one_system"], iters, how, os.cpu_count(), REF_BUILD)}
            # lock-step: the same K iterations on the GPU, per-shift relative difference to the reference's X
            solve(itK)
            parity["lockstep"] = {"iterations": itK, "against": smp["kind"], "tol": 1e-10,
                                  "x_rel": [rel(Xn[s], Xr[s]) for s in range(S)]}
            parity["lockstep"]["pass"] = bool(max(parity["lockstep"]["x_rel"]) < 1e-10)
    line = {"metric": "sbcgrq_time_to_solution", "value": dev_s, "unit": "s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(wname, w, world),
            "iterations": iters, "residual": resid, "ms_per_iteration": 1e3 * dev_s / max(iters, 1),
            "timed_region_s_per_step": region_s,
            "e2e": {"value": e2e_s, "unit": "s", "h2d_bytes_per_step": int(F) * world,
                    "d2h_bytes_per_step": int(S * F) * world},
            "gpu_launches": launches, "clocks": clocks, "parity": parity, "loop": loop, "weak_scaling": weak,
            "roofline": roofline,
            "dirac_op": {"kernel": "dirac_chain_kernel (block Dirac apply; +gram = with the fused P^dag T epilogue)",
                         "GBps": kern["dirac"]["achieved"], "frac_of_hbm_peak": kern["dirac"]["frac"],
                         "ms": kern["dirac"]["ms"], "GBps_with_gram": dirac["achieved"],
                         "frac_of_hbm_peak_with_gram": dirac["frac"], "ms_with_gram": dirac["ms"],
                         "ms_with_gram_in_loop": profile["ms"]["dirac_gram"] if profile and profile["iterations"] else None,
                         "alg_bytes": dirac["alg_bytes"]},
            "kernels": kern, "cpu_baseline": cpu}
    print(json.dumps(line), file=JSON_OUT, flush=True)
    if world == 1 and args.record_iterations:
        os.makedirs(os.path.dirname(ITER_FILE), exist_ok=True)
        d = {}
        try:
            d = json.load(open(ITER_FILE))
        except Exception:
            pass
        d[wname] = {"iterations": iters, "active_hist": hist}
        json.dump(d, open(ITER_FILE, "w"), indent=1)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sbcgrq_24^4_N12", choices=sorted(WORKLOADS))
    ap.add_argument("--max-it", type=int, default=1000000)
    ap.add_argument("--cpu-iters", type=int, default=4, help="iterations in the bounded CPU sample")
    ap.add_argument("--profile-iters", type=int, default=64,
                    help="iterations of the last warm-up solve timed stage by stage inside the loop (0: off)")
    ap.add_argument("--profile-after", type=int, default=3000,
                    help="iterations to run before the profiled window starts (sustained clocks)")
    ap.add_argument("--multi-lockstep-iters", type=int, default=8,
                    help="N > 1 GPUs: iterations of the slab loop compared with the single-domain loop")
    ap.add_argument("--weak-iters", type=int, default=200,
                    help="iterations of the weak-scaling probe (the headline volume per GPU; 0: off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: NCCL halo / all-reduce instead of peer-memory stores")
    ap.add_argument("--record-iterations", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, args.workload)
    else:
        run_ours(args, w, args.workload)


if __name__ == "__main__":
    main()
