#!/usr/bin/env python
"""bench.py -- SBCGrQ time-to-solution on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun ... bench.py --gpus N ...            (one rank per GPU, slab decomposition)

A *step* is one complete multishift solve of the headline workload (configs[2]:
SBCGrQ, 24^4 sites, N=12 right-hand sides, mass 1e-3, tol 1e-10, the nine benchmark
shifts of benchmark.cpp:12-13) on synthetic inputs (links and sources uniform in
[-1,1]+i[-1,1], the reference's distribution, inc/dirac_op.hpp:27-32).

 value   time-to-solution with inputs resident in HBM (CUDA events around setup + loop,
         taken inside the library on its own stream), mean over the K timed steps,
         max over ranks.
 e2e     the same solves measured through the host-buffer C-ABI call a reference user
         makes (bcg_solve_sbcgrq <- SBCGrQ<N>): pinned host B in, nine host X out, wall
         clock around the call, copies inside the timed region.
 roofline  per-kernel device time measured live here with CUDA events
         (bcg_bench_kernel) on fields of the same size; algorithmic bytes per SURVEY 8(d).
 cpu_baseline  the unmodified reference (oracle/_ref) timed on this box's host cores on a
         bounded sample (first few iterations), extrapolated to the GPU iteration count.

--impl reference runs the reference's own CPU implementation (oracle/_ref, else the
oracle port) on the same workload: a bounded sample per step, extrapolated.
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
WORKLOADS = {
    # configs[2] of BASELINE.json: the configuration the metric is quoted on
    "sbcgrq_24^4_N12": dict(V=24 ** 4, N=12, mass=1e-3, eps=1e-10, eps_shifts=1e-15, shifts=BENCH_SHIFTS),
    # configs[0]: the reference's README default (CPU-runnable)
    "sbcgrq_V1000_N12": dict(V=1000, N=12, mass=1e-3, eps=1e-10, eps_shifts=1e-15, shifts=BENCH_SHIFTS),
    "sbcgrq_16^4_N12": dict(V=16 ** 4, N=12, mass=1e-3, eps=1e-10, eps_shifts=1e-15, shifts=BENCH_SHIFTS),
    # configs[3]: 48^3 x 96 sites on 2 / 4 / 8 GPUs (64 / 32 / 16 GB of fields per GPU); a full solve takes
    # ~25 000 iterations of ~40 ms / n_gpus, so this one is meant to be run with --max-it
    "sbcgrq_48^3x96_N12": dict(V=48 ** 3 * 96, N=12, mass=1e-3, eps=1e-10, eps_shifts=1e-15, shifts=BENCH_SHIFTS),
}
ITER_FILE = os.path.join(ROOT, "profiles", "bench_iterations.json")
CLOCK_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
               "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
               "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


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


def known_iterations(workload):
    try:
        return json.load(open(ITER_FILE)).get(workload)
    except Exception:
        return None


def reference_sample(w, U, B, sample_iters):
    """Time the reference's own SBCGrQ (oracle/_ref; else the oracle port) for the first
    `sample_iters` iterations on host cores.  Returns (seconds, kind, cores)."""
    from oracle.pyoracle import Oracle, RefShim, build as build_oracle
    if RefShim.available(w["N"]):
        r = RefShim(w["N"])
        _, it, sec = r.SBCGrQ(U, B, w["mass"], w["shifts"], w["eps"], w["eps_shifts"], max_it=sample_iters)
        return sec, it, "reference", 1
    build_oracle()
    o = Oracle()
    _, it, sec, _ = o.SBCGrQ(U, B, w["mass"], w["shifts"], w["eps"], w["eps_shifts"], max_it=sample_iters)
    return sec, it, "port", 1


def run_reference(args, w, wname):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    U, B = make_inputs(w["V"], w["N"])
    iters_full = known_iterations(wname)
    sample_iters = args.cpu_iters
    times = []
    kind = cores = None
    for i in range(args.warmup + args.steps):
        sec, it, kind, cores = reference_sample(w, U, B, sample_iters)
        if i >= args.warmup:
            times.append(sec / max(it, 1))
    s_per_iter = statistics.mean(times)
    extrap = iters_full if iters_full else sample_iters
    value = s_per_iter * extrap
    sample = ("first %d iterations of the reference SBCGrQ<12> per step on the same inputs; %.3f s/iteration x %s"
              % (sample_iters, s_per_iter,
                 ("%d iterations (GPU-arm count, profiles/bench_iterations.json) = extrapolated time-to-solution"
                  % iters_full) if iters_full else "sample only (no iteration count recorded yet)"))
    line = {"impl": "reference", "metric": "sbcgrq_time_to_solution", "value": value, "unit": "s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * value,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_of(wname, w, args.gpus),
            "cpu_baseline": {"value": value, "unit": "s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "extrapolated": True, "s_per_iteration": s_per_iter}
    print(json.dumps(line), file=JSON_OUT, flush=True)


EXCHANGE = ("halo sites and Gram blocks exchanged by P2P stores over NVLink from inside the kernels "
            "(NCCL only in the set-up)")


def config_of(wname, w, gpus):
    return {"workload": wname, "V": w["V"], "n_rhs": w["N"], "n_shifts": len(w["shifts"]), "mass": w["mass"],
            "eps": w["eps"], "eps_shifts": w["eps_shifts"], "operator": "reference 1-D chain (inc/dirac_op.hpp:13-21)",
            "partition": "1 slab" if gpus == 1 else "%d contiguous site slabs; %s" % (gpus, EXCHANGE),
            "l2_policy": ("%d fields of %.0f MB each per GPU stream through every pair of iterations (working set %s"
                          " the 126 MB L2); no explicit flush"
                          % (2 * len(w["shifts"]) + 3, 48.0 * w["N"] * w["V"] / gpus / 1e6,
                             "exceeds" if (2 * len(w["shifts"]) + 3) * 48.0 * w["N"] * w["V"] / gpus > 126e6
                             else "FITS IN"))}


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

    def step():
        t0 = time.perf_counter()
        info = ctx.solve_sbcgrq(Xn, Bn, w["shifts"], w["eps"], w["eps_shifts"], args.max_it)
        t1 = time.perf_counter()
        return info, t1 - t0

    for _ in range(args.warmup):
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

    # ---- per-kernel roofline, measured live with CUDA events on same-size fields ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    F, Ub = 48.0 * N * Vl, 144.0 * Vl
    hs = [ctx.field(Bn) for _ in range(2)] + [ctx.field() for _ in range(2 * S - 1)]
    for h in hs[2:]:
        ctx.copy(h, hs[0])
    kern = {}
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(wname, {}) if world == 1 else {}
    except Exception:
        pass
    # The loop serves the shifted systems every second iteration (shift_pair.cuh).  "shift_pair" is the
    # AVERAGE launch of an odd + even pair.  Its numerator is the fixed per-unit figure of SURVEY 8(d) x the
    # units a pair processes -- 2 back-substitutions (2 F each) and 2 S system updates (4 F each), i.e.
    # (2 + 4 S) F per launch, the same figure as for the plain kernel -- so that serving X_s, P_s once per two
    # iterations shows up as bandwidth; the bytes the pair really moves ((14 + 4 (S-1)) F / 2 per launch) are
    # the ncu figure in `traffic`.  "shift_update" (plain kernel, every system every iteration): for comparison.
    paired = S > 1 and os.environ.get("BCG_PAIR", "1") != "0"
    todo = [("dirac_gram", 0, 2, 1, 2 * F + Ub, 1), ("dirac", 1, 2, 1, 2 * F + Ub, 1), ("axpy_gram", 3, 2, 1, 3 * F, 1),
            ("shift_update", 4, 1 + 2 * S, S, (2 + 4 * S) * F, 1)]
    if paired:
        todo.append(("shift_pair", 13, 1 + 2 * S, S, (2 + 4 * S) * F, 2))
    for name, which, nh, ns, nbytes, per_rep in todo:
        ms, _ = ctx.bench_kernel(which, 20 // per_rep, hs[:nh], ns)
        ms /= per_rep
        kern[name] = {"ms": ms, "alg_bytes": nbytes, "achieved": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak,
                      "traffic": traffic.get(name)}
    for h in hs:
        ctx.free(h)
    # one iteration = stencil+Gram, Q update, multishift update ("dirac" = the stencil without its Gram epilogue)
    in_loop = ["dirac_gram", "axpy_gram", "shift_pair" if paired else "shift_update"]
    it_ms = sum(kern[n_]["ms"] for n_ in in_loop)
    dom = max(in_loop, key=lambda k: kern[k]["ms"])
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kern[dom]["achieved"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["frac"], "traffic": kern[dom]["traffic"], "peak_source": peak_src,
                "share_of_iteration": kern[dom]["ms"] / it_ms,
                "alg_bytes_per_launch": kern[dom]["alg_bytes"], "ms_per_launch": kern[dom]["ms"]}
    if dom == "shift_pair":
        moved = (14 + 4 * (S - 1)) * F / 2
        roofline["note"] = ("average launch of an odd+even pair; numerator = fixed per-unit bytes (2 F per back-substitution, "
                            "4 F per system update) x units; the pair touches X_s, P_s once per two iterations, so it moves "
                            "%.3g B per launch (%.0f GB/s = %.2f of peak) and is limited on chip: ncu of the even launch "
                            "L1/shared 78 %%, FP64 57 %%, DRAM 45 %%" % (moved, moved / kern[dom]["ms"] / 1e6,
                                                                      moved / kern[dom]["ms"] / 1e6 / peak))
        roofline["moved_bytes_per_launch"] = moved
    dirac = kern["dirac_gram"]

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline and not big:
        sec, it_cpu, kind, cores = reference_sample(w, U, B, args.cpu_iters)
        spi = sec / max(it_cpu, 1)
        cpu = {"value": spi * iters, "unit": "s", "cores": cores, "kind": kind,
               "sample": "first %d iterations of the reference SBCGrQ<12> on the same inputs (%.2f s, %.3f s/iteration)"
                         " x %d GPU iterations = extrapolated time-to-solution; host has %d cores, the reference is"
                         " single-threaded" % (it_cpu, sec, spi, iters, os.cpu_count())}
    line = {"metric": "sbcgrq_time_to_solution", "value": dev_s, "unit": "s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(wname, w, world),
            "iterations": iters, "residual": resid, "ms_per_iteration": 1e3 * dev_s / max(iters, 1),
            "timed_region_s_per_step": region_s,
            "e2e": {"value": e2e_s, "unit": "s", "h2d_bytes_per_step": int(F) * world,
                    "d2h_bytes_per_step": int(S * F) * world},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "dirac_op": {"kernel": "dirac_chain_kernel (block Dirac apply; +gram = with the fused P^dag T epilogue)",
                         "GBps": kern["dirac"]["achieved"], "frac_of_hbm_peak": kern["dirac"]["frac"],
                         "ms": kern["dirac"]["ms"], "GBps_with_gram": dirac["achieved"],
                         "frac_of_hbm_peak_with_gram": dirac["frac"], "ms_with_gram": dirac["ms"],
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
        d[wname] = iters
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
