#!/usr/bin/env python
"""bench.py -- throughput of the accelerated hot path on B200: batched find_leaf_at + per-leaf quadratic fits + model
evaluation over a CSp-tree (BASELINE.json: "queries/sec (leaf lookup + quadratic eval) and leaf fits/sec").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C2|C3] [--queries M]

A step is one pass of the hot path over one resident batch: re-fit every leaf model (K2, csp_models_refit) and locate +
evaluate M queries (K1+K3a fused, csp_locate_eval_dev).  N=1 workload: BASELINE config C2 (CS2, n=10, 1e5 leaves, 1e8
queries = 8 GB of resident float64 coordinates -- far larger than L2, so every step streams from HBM).  N>1: the same
per-GPU work on every rank (weak scaling); rank 0 grows the tree, the flattened blob is NCCL-broadcast once, queries and
fits then run with no further communication.

`--impl reference` times the reference's algorithm on the host cores instead (the C++ restatement in oracle/, since
Julia is not installed in this image), all host threads, bounded samples.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_QUERY = lambda n: 8 * n + 4 + 8 + 1  # coordinates in, leaf id + value + status out  (8n+12, +1 status)


def model_bytes(n):
    return 8 * (1 + 2 * n + n * (n + 1) // 2)


def fit_bytes(n, visited):
    """SURVEY 8(d) per-leaf algorithmic bytes of a CS2 fit."""
    return model_bytes(n) + 64 * n * (n - 1) // 2 + 8 * n * n + 48 * ((n + 1) // 2) + 8 * visited


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.lines, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not shrink the CPU baseline)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def grow_config(cb, name):
    root, obj = cb.synthetic.config_tree(name)
    return root, obj


def grow_oracle_tree(cb, O, name):
    """The same tree (same points, dimlists, objective) in the CPU oracle's pointer representation."""
    import ctypes as C
    c = cb.synthetic.CONFIGS[name]
    obj = c["objective"]()
    n, p = c["n"], c["p"]
    rng = np.random.default_rng(1000)
    lo, up = np.full(n, float(c.get("lower", -1.0))), np.full(n, float(c.get("upper", 1.0)))
    base = c.get("base", 0.0)
    pos = np.full(n, float(base)) if np.isscalar(base) else np.asarray(base, np.float64)
    X = lo + (up - lo) * rng.random((c["npoints"], n))
    cycle = cb.synthetic.cs1_dimlists(n) if p == 1 else cb.synthetic.round_robin_dimlists(n)
    oroot = O.Box(O.World(lo, up, pos, obj(pos)), p)
    O.addpoints_cycle(oroot, X, cycle, O.OBJECTIVE(obj.fn_ptr), C.c_void_p(obj.ctx))
    O.index_tree(oroot)
    return oroot, obj


def cpu_reference_leg(cb, name, nq, nfit, seed=7):
    """Times the reference algorithm (oracle port) on the host: locate+eval of nq queries and nfit leaf fits."""
    from oracle import oracle as O
    oroot, _ = grow_oracle_tree(cb, O, name)
    threads = host_threads()
    lv = O.leaves(oroot)
    rng = np.random.default_rng(seed)
    sel = rng.choice(len(lv), size=min(nfit, len(lv)), replace=False)
    t0 = time.perf_counter()
    fit_sample = O.fit_boxes_batch([lv[i] for i in sel], nthreads=threads)
    fit_s = time.perf_counter() - t0
    n = oroot.n
    # models for evaluation: fitted for the sampled leaves, zeros elsewhere (evaluation cost is independent of the values)
    nl = len(lv)
    c, g, Q, b = np.zeros(nl), np.zeros((nl, n)), np.zeros((nl, n, n)), np.zeros((nl, n))
    c[sel], g[sel], Q[sel], b[sel] = fit_sample["c"], fit_sample["g"], np.nan_to_num(fit_sample["Q"]), fit_sample["b"]
    w = oroot.world
    X = w.lower + (w.upper - w.lower) * rng.random((nq, n))
    _, _, _, q_s = O.locate_eval_batch(oroot, X, c, g, Q, b, nthreads=threads)
    return dict(threads=threads, q_per_s=nq / q_s, fits_per_s=len(sel) / fit_s, nq=nq, nfit=len(sel), q_s=q_s, fit_s=fit_s,
                oroot=oroot, models=(c, g, Q, b))


def run_reference(args):
    """--impl reference: K timed steps of a bounded sample on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import csp_b200 as cb
    from oracle import oracle as O
    name = args.config
    oroot, _ = grow_oracle_tree(cb, O, name)
    threads = host_threads()
    n = oroot.n
    lv = O.leaves(oroot)
    nl = len(lv)
    rng = np.random.default_rng(11)
    nq = args.ref_step_queries
    sel = rng.choice(nl, size=min(args.ref_fits, nl), replace=False)
    boxes = [lv[i] for i in sel]
    c, g, Q, b = np.zeros(nl), np.zeros((nl, n)), np.zeros((nl, n, n)), np.zeros((nl, n))
    w = oroot.world
    times, fit_times = [], []
    for it in range(args.warmup + args.steps):
        X = w.lower + (w.upper - w.lower) * rng.random((nq, n))
        t0 = time.perf_counter()
        fit = O.fit_boxes_batch(boxes, nthreads=threads)
        t1 = time.perf_counter()
        c[sel], g[sel], Q[sel], b[sel] = fit["c"], fit["g"], np.nan_to_num(fit["Q"]), fit["b"]
        _, _, _, q_s = O.locate_eval_batch(oroot, X, c, g, Q, b, nthreads=threads)
        if it >= args.warmup:
            times.append(q_s)
            fit_times.append(t1 - t0)
    qps = nq * len(times) / sum(times)
    out = {"impl": "reference", "metric": "queries/sec (leaf lookup + quadratic eval)", "value": qps, "unit": "queries/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(name, nq, nl, n),
           "fits": {"value": len(boxes) * len(fit_times) / sum(fit_times), "unit": "leaf fits/s"},
           "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                            "sample": f"{nq} uniform queries per step located + evaluated, {len(boxes)} leaf fits per step; C++ "
                                      "restatement of the Julia reference (oracle/), not Julia"},
           "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def workload_config(name, m, nl, n):
    desc = {"C2": "C2: CS2 tree (p=2), n=10, 1e5 leaves, synthetic quadratic + noise", "C3": "C3: CS2 tree (p=2), n=20, 1e6 leaves",
            "C1": "C1: CS1 tree (p=1), n=2, 1k leaves, Rosenbrock"}.get(name, name)
    return {"workload": desc, "queries_per_step_per_gpu": int(m), "leaves": int(nl), "n": int(n),
            "step": "refit all leaf models + locate/evaluate all queries", "l2": "inputs larger than L2 (no flush needed)"}


_REAL_STDOUT = None


def capture_stdout():
    """Route everything written to fd 1 (NCCL / library banners included) to stderr; the JSON line goes to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--queries", type=int, default=100_000_000)
    ap.add_argument("--e2e-queries", type=int, default=None)
    ap.add_argument("--ref-queries", type=int, default=20_000_000)
    ap.add_argument("--ref-fits", type=int, default=100_006)
    ap.add_argument("--ref-step-queries", type=int, default=5_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import csp_b200 as cb
    from csp_b200.device import DeviceTree

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the accelerated path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # ---- tree: grown on rank 0 (host), flattened, broadcast once as a blob, uploaded per GPU
    name = args.config
    t0 = time.perf_counter()
    blob = None
    if rank == 0:
        root, _ = grow_config(cb, name)
        blob = cb.pack_blob(root.tree.flat())
    bcast_ms = 0.0
    if world > 1:
        dist.all_reduce(torch.zeros(1, device=dev))  # communicator set-up is not part of the broadcast time
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dblob = cb.shard.broadcast_blob(blob, src=0, device=dev)  # C1: the one collective of the path (NCCL over NVLink)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        tree = DeviceTree(blob=dblob, device=local)  # upload straight from the device buffer the broadcast landed in
    else:
        tree = DeviceTree(blob=blob, device=local)
    setup_s = time.perf_counter() - t0
    n, nl = tree.n, tree.n_leaves

    # ---- resident queries (synthetic, uniform in the world box [-1,1]^n of the configs), per-rank seed
    m = args.queries
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    X = torch.empty((m, n), dtype=torch.float64, device=dev)
    chunk = 1 << 22
    for o in range(0, m, chunk):
        X[o:o + chunk] = torch.rand((min(chunk, m - o), n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    leaf = torch.empty(m, dtype=torch.int32, device=dev)
    val = torch.empty(m, dtype=torch.float64, device=dev)
    status = torch.empty(m, dtype=torch.uint8, device=dev)
    models = tree.fit_models()

    def step(evs=None):
        if evs:
            evs[0].record()
        tree.refit_models(models)
        if evs:
            evs[1].record()
        tree.locate_eval(models, X, leaf=leaf, val=val, status=status)
        if evs:
            evs[2].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = cb.kernel_launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        step(evs[k])
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = cb.kernel_launches() - launches0
    total_ms = evs[0][0].elapsed_time(end)
    fit_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    loc_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    tms = torch.tensor([total_ms, fit_ms, loc_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    total_ms, fit_ms, loc_ms = [float(v) for v in tms.cpu()]
    ms_per_step = total_ms / args.steps
    value = world * m / (ms_per_step * 1e-3)

    # second pass, instrumented: per-kernel device times
    from csp_b200 import device as cdev
    cdev.timing_read()
    cdev.timing_enable(True)
    for k in range(args.steps):
        step()
    KT = cdev.timing_read()
    cdev.timing_enable(False)
    if world > 1:
        dist.barrier()

    # distinct leaves touched (the model bytes that must cross HBM at least once per step)
    touched = int(torch.unique(leaf[: min(m, 20_000_000)]).numel())
    bad = int((status != 0).sum().item())
    nanvals = int(torch.isnan(val).sum().item())

    # ---- e2e: the same metric through the host-buffer C-ABI call, pinned host memory, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        me = args.e2e_queries or m
        try:
            Xh = torch.empty((me, n), dtype=torch.float64, pin_memory=True)
            Xh.copy_(X[:me])
            lh = torch.empty(me, dtype=torch.int32, pin_memory=True)
            vh = torch.empty(me, dtype=torch.float64, pin_memory=True)
            sh = torch.empty(me, dtype=torch.uint8, pin_memory=True)
            Xn, ln, vn, sn = Xh.numpy(), lh.numpy(), vh.numpy(), sh.numpy()
            esteps = max(2, min(args.steps, 5))
            tree.locate_eval(models, Xn, leaf=ln, val=vn, status=sn)  # warm-up
            barrier()
            t0 = time.perf_counter()
            for _ in range(esteps):
                tree.refit_models(models)
                tree.locate_eval(models, Xn, leaf=ln, val=vn, status=sn)  # returns when results are in host memory
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            same = bool(np.array_equal(ln[:100000], leaf[:100000].cpu().numpy()))
            e2e = {"value": world * me * esteps / dt, "unit": "queries/s", "h2d_bytes_per_step": int(me * n * 8),
                   "d2h_bytes_per_step": int(me * 13), "steps": esteps, "queries_per_step_per_gpu": int(me),
                   "matches_resident_path": same, "api": "csp_locate_eval (host buffers, pinned)"}
            del Xh, lh, vh, sh
        except Exception as ex:  # e.g. not enough pinnable host memory
            e2e = {"value": None, "unit": "queries/s", "error": str(ex)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: algorithmic bytes / measured kernel time (CUDA events on the launching stream, library-side)
    peak, peak_src = peaks()
    kt = KT
    per_step = {k: (v[0] / args.steps, v[1] / args.steps) for k, v in kt.items() if v[1] > 0}
    pipe_ms = sum(v[0] for k, v in per_step.items() if k != "fit")
    alg_q = {"locate": 8 * n + 5, "locate_eval_fused": BYTES_PER_QUERY(n), "scan": 0, "scatter": 4 + 8, "eval_binned": 8 + 8 * n + 8,
             "eval": 4 + 8 * n + 8}
    kernels = {}
    for k, (ms, cnt) in per_step.items():
        if k == "fit":
            continue
        b = m * alg_q.get(k, 0) + (touched * model_bytes(n) if k in ("locate_eval_fused", "eval_binned") else 0) + \
            (tree.kib * 1024 if k in ("locate", "locate_eval_fused") else 0)
        kernels[k] = {"ms_per_step": ms, "launches_per_step": cnt, "share": ms / max(pipe_ms, 1e-9),
                      "algorithmic_GBps": b / (ms * 1e-3) / 1e9 if ms > 0 else None, "frac": b / (ms * 1e-3) / 1e9 / peak if ms > 0 else None}
    dom = max((k for k in kernels), key=lambda k: kernels[k]["ms_per_step"])
    traffic, traffic_src = None, None
    try:  # DRAM traffic of the dominant kernel from the committed ncu --set full capture, scaled to this run's launch size
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        key = {"locate": "k_locate_only", "eval_binned": "k_eval_binned", "scatter": "k_scatter"}.get(dom)
        if key and n == 10:
            traffic = tj[key]["dram_bytes_per_query"] * m / max(kernels[dom]["launches_per_step"], 1)
            traffic_src = "profiles/r1_%s.ncu-rep: %.1f DRAM bytes/query x queries per launch" % (key, tj[key]["dram_bytes_per_query"])
    except Exception:
        pass
    p_ = tree.p
    eval_name = ("k_eval_dmma (FP64 tensor cores, grouped dense modelvalue)" if n > 32 else
                 "k_eval_binned_mma<%d> (model-stationary modelvalue on the FP64 tensor cores)" % n if (n >= 16 and n % 2 == 0) else
                 "k_eval_binned_exact<%d> (model-stationary modelvalue)" % n)
    names = {"locate": ("k_locate_wide<%d>" if n > 32 else "k_locate_only<%d>") % p_ + " (find_leaf_at descent + per-leaf histogram)",
             "eval_binned": eval_name, "locate_eval_fused": "k_locate<%d,...> (fused find_leaf_at + modelvalue)" % p_,
             "scatter": "k_scatter + k_unpermute", "scan": "k_scan", "eval": "k_eval"}
    alg_bytes = m * BYTES_PER_QUERY(n) + touched * model_bytes(n) + tree.kib * 1024
    roofline = {"bound": "hbm", "kernel": names.get(dom, dom), "achieved": kernels[dom]["algorithmic_GBps"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": m * alg_q.get(dom, 0) / max(kernels[dom]["launches_per_step"], 1),
                "ms_per_launch": kernels[dom]["ms_per_step"] / max(kernels[dom]["launches_per_step"], 1),
                "launches_per_step": kernels[dom]["launches_per_step"], "share_of_pipeline": kernels[dom]["share"],
                "bytes_per_query": alg_q.get(dom), "kernels": kernels,
                "pipeline": {"what": "whole locate+evaluate pipeline of one step (all its kernels)", "algorithmic_bytes_per_step": int(alg_bytes),
                             "bytes_per_query": BYTES_PER_QUERY(n), "distinct_leaves_touched": touched, "kernel_ms_per_step": pipe_ms,
                             "achieved": alg_bytes / (loc_ms * 1e-3) / 1e9, "frac": alg_bytes / (loc_ms * 1e-3) / 1e9 / peak},
                "note": "kernel times from a second, instrumented pass of the same K steps (csp_timing_*: CUDA events around "
                        "every launch on the launching stream); value/ms_per_step come from the uninstrumented pass"}
    out = {"metric": "queries/sec (leaf lookup + quadratic eval)", "value": value, "unit": "queries/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(name, m, nl, n),
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
           "fits": {"value": world * nl / (fit_ms * 1e-3), "unit": "leaf fits/s", "ms_per_launch": fit_ms, "leaves": nl,
                    "kernel_ms": KT["fit"][0] / args.steps, "kernel_launches_per_step": KT["fit"][1] / args.steps,
                    "algorithmic_bytes_per_leaf": fit_bytes(n, 136 if n == 10 else 574),
                    "frac": nl * fit_bytes(n, 136 if n == 10 else 574) / (max(KT["fit"][0] / args.steps, 1e-9) * 1e-3) / 1e9 / peak,
                    "kernel": "k_fit_chain (chaintop + fit_poly1) + k_fit<2> (coefficients_p + fit_poly2_diag + canonicalize)"},
           "locate_eval_ms": loc_ms, "tree_broadcast_ms": bcast_ms, "setup_s": setup_s, "outside_world": bad, "nan_values": nanvals}

    if not args.no_cpu_baseline:
        try:
            leg = cpu_reference_leg(cb, name, nq=args.ref_queries, nfit=args.ref_fits)
            out["cpu_baseline"] = {"value": leg["q_per_s"], "unit": "queries/s", "cores": leg["threads"], "kind": "port",
                                   "fits_per_s": leg["fits_per_s"],
                                   "sample": f"{leg['nq']} queries located+evaluated in {leg['q_s']:.2f}s and {leg['nfit']} leaf fits in "
                                             f"{leg['fit_s']:.2f}s on the same tree; C++ restatement of the Julia reference, not Julia"}
        except Exception as ex:
            out["cpu_baseline"] = {"value": None, "error": str(ex)[:200]}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
