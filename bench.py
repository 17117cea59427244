#!/usr/bin/env python
"""bench.py -- throughput of the accelerated hot path on B200: batched find_leaf_at + per-leaf quadratic fits + model
evaluation over a CSp-tree (BASELINE.json: "queries/sec (leaf lookup + quadratic eval) and leaf fits/sec, 1/2/4/8 B200").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C3] [--queries M]

Workload (every N): BASELINE config C3, the north-star configuration -- ONE CS2 tree, n=20, 1 000 021 leaves, grown on rank 0,
packed, NCCL-broadcast once and uploaded on every GPU from the device buffer; ONE batch of 1e9 queries per step, sharded by
contiguous ranges over the N GPUs (strong scaling; resident in HBM, in chunks of <= 1e8 = 16 GB, so every pass streams from
HBM); leaf fits sharded by leaf-id range with the fitted records all-gathered; (leaf id, value) of every query gathered to
rank 0 over NCCL inside the timed region, pipelined behind the kernels.  All collectives are the library's own
(csp_shard_*, csrc/comm.cu); torch.distributed only carries the NCCL id, the barriers and the max-over-ranks of the timings.
A step = sharded re-fit of all leaf models + all-gather + locate/evaluate the whole batch + gather.

Secondary keys (N=1 unless noted): "c2" (BASELINE C2: CS2 n=10, 1e5 leaves, 1e8 queries), "c4" (CS1 n=100, dense evaluation on the
FP64 tensor cores), "c5" (CS2 n=300 leaf fits, sharded over the N GPUs).

`--impl reference` times the reference's algorithm on the host cores instead (the C++ restatement in oracle/, since Julia is not
installed in this image), all host threads, bounded samples of the same workload.  Prints ONE JSON line (rank 0).
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


WORKLOADS = {
    "C1": "C1: CS1 tree (p=1), n=2, 1k leaves, Rosenbrock",
    "C2": "C2: CS2 tree (p=2), n=10, 1e5 leaves, synthetic quadratic + noise; 1e8 queries",
    "C3": "C3: CS2 tree (p=2), n=20, 1e6 leaves (tree NCCL-broadcast once); 1e9 queries sharded across the GPUs",
}


def workload_config(name, total, nl, n):
    """Identical for both arms (the driver compares it); arm-specific detail goes under "partition" / "cpu_baseline"."""
    return {"workload": WORKLOADS.get(name, name), "queries_per_step": int(total), "leaves": int(nl), "n": int(n),
            "step": "re-fit all leaf models + locate/evaluate every query of the batch (results on rank 0)",
            "l2": "inputs larger than L2 (no flush needed): every resident chunk is >= 1 GB"}


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


# ------------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: K timed steps of a bounded sample of the same workload on the host cores (rank 0 only)."""
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
    # models: fitted for the sampled leaves, zeros elsewhere (the cost of an evaluation does not depend on the values);
    # kept in the oracle's column-major layout so that no per-step conversion is timed or repeated
    c, g, Qf, b = np.zeros(nl), np.zeros((nl, n)), np.zeros((nl, n, n)), np.zeros((nl, n))
    w = oroot.world
    times, fit_times = [], []
    for it in range(args.warmup + args.steps):
        X = w.lower + (w.upper - w.lower) * rng.random((nq, n))
        t0 = time.perf_counter()
        fit = O.fit_boxes_batch(boxes, nthreads=threads)
        t1 = time.perf_counter()
        c[sel], g[sel], b[sel] = fit["c"], fit["g"], fit["b"]
        Qf[sel] = np.nan_to_num(fit["Q"]).transpose(0, 2, 1)
        _, _, _, q_s = O.locate_eval_batch(oroot, X, c, g, None, b, nthreads=threads, Qf=Qf)
        if it >= args.warmup:
            times.append(q_s)
            fit_times.append(t1 - t0)
    qps = nq * len(times) / sum(times)
    total = args.queries or DEFAULT_QUERIES.get(name, 100_000_000)
    out = {"impl": "reference", "metric": "queries/sec (leaf lookup + quadratic eval)", "value": qps, "unit": "queries/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(name, total, nl, n),
           "fits": {"value": len(boxes) * len(fit_times) / sum(fit_times), "unit": "leaf fits/s"},
           "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                            "sample": f"{nq} uniform queries per step located + evaluated (a bounded sample of the {total}-query batch), "
                                      f"{len(boxes)} leaf fits per step, on the same tree; C++ restatement of the Julia reference "
                                      "(oracle/), not Julia"},
           "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


DEFAULT_QUERIES = {"C3": 1_000_000_000, "C2": 100_000_000, "C1": 100_000}


# ------------------------------------------------------------------------------------------------------------------------
# helpers of the GPU arm
# ------------------------------------------------------------------------------------------------------------------------
def fill_uniform(torch, X, seed, lo=-1.0, hi=1.0):
    gen = torch.Generator(device=X.device)
    gen.manual_seed(seed)
    step = 1 << 22
    for o in range(0, X.shape[0], step):
        k = min(step, X.shape[0] - o)
        X[o:o + k] = torch.rand((k, X.shape[1]), dtype=torch.float64, device=X.device, generator=gen) * (hi - lo) + lo


def kernel_table(KT, steps, m_per_step, n, touched, tree_bytes, peak, p_):
    """Per-kernel-class device time -> achieved algorithmic GB/s and roofline fraction (HBM)."""
    per_step = {k: (v[0] / steps, v[1] / steps) for k, v in KT.items() if v[1] > 0}
    pipe_ms = sum(v[0] for k, v in per_step.items() if k != "fit")
    alg_q = {"locate": 8 * n + 5, "locate_eval_fused": BYTES_PER_QUERY(n), "scan": 0, "scatter": 4 + 8, "eval_binned": 8 + 8 * n + 8,
             "eval": 4 + 8 * n + 8}
    kernels = {}
    for k, (ms, cnt) in per_step.items():
        if k == "fit":
            continue
        b = m_per_step * alg_q.get(k, 0) + (touched * model_bytes(n) if k in ("locate_eval_fused", "eval_binned") else 0) + \
            (tree_bytes if k in ("locate", "locate_eval_fused") else 0)
        kernels[k] = {"ms_per_step": ms, "launches_per_step": cnt, "share": ms / max(pipe_ms, 1e-9),
                      "algorithmic_GBps": b / (ms * 1e-3) / 1e9 if ms > 0 else None, "frac": b / (ms * 1e-3) / 1e9 / peak if ms > 0 else None}
    eval_name = ("k_eval_dmma (FP64 tensor cores, grouped dense modelvalue)" if n > 32 else
                 "k_eval_binned_mma<%d> (model-stationary modelvalue on the FP64 tensor cores)" % n if (n >= 16 and n % 2 == 0) else
                 "k_eval_binned_exact<%d> (model-stationary modelvalue)" % n)
    names = {"locate": ("k_locate_wide<%d>" if n > 32 else "k_locate_ring<%d>" if n >= 8 else "k_locate_only<%d>") % p_ +
                       " (find_leaf_at descent + per-leaf histogram)",
             "eval_binned": eval_name, "locate_eval_fused": "k_locate<%d,...> (fused find_leaf_at + modelvalue)" % p_,
             "scatter": "k_scatter + k_unpermute", "scan": "k_scan", "eval": "k_eval"}
    return kernels, pipe_ms, alg_q, names


def traffic_entry(config, kernel_class):
    """DRAM bytes per query of a kernel class from the committed `ncu --set full` capture of this round (profiles/r2_traffic.json,
    keyed by config); None when that config / kernel was not captured."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        e = tj[config][kernel_class]
        return float(e["dram_bytes_per_query"]), e.get("source", "profiles/r2_traffic.json")
    except Exception:
        return None, None


def roofline_block(config, KT, steps, m_per_step, n, p_, touched, tree_bytes, loc_ms):
    peak, peak_src = peaks()
    kernels, pipe_ms, alg_q, names = kernel_table(KT, steps, m_per_step, n, touched, tree_bytes, peak, p_)
    if not kernels:
        return None
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    lp = max(kernels[dom]["launches_per_step"], 1)
    tq, tsrc = traffic_entry(config, dom)
    alg_bytes = m_per_step * BYTES_PER_QUERY(n) + touched * model_bytes(n) + tree_bytes
    return {"bound": "hbm", "kernel": names.get(dom, dom), "achieved": kernels[dom]["algorithmic_GBps"], "peak": peak, "unit": "GB/s",
            "frac": kernels[dom]["frac"], "peak_source": peak_src,
            "traffic": None if tq is None else tq * m_per_step / lp, "traffic_source": tsrc,
            "algorithmic_bytes_per_launch": m_per_step * alg_q.get(dom, 0) / lp, "ms_per_launch": kernels[dom]["ms_per_step"] / lp,
            "launches_per_step": kernels[dom]["launches_per_step"], "share_of_pipeline": kernels[dom]["share"],
            "bytes_per_query": alg_q.get(dom), "kernels": kernels,
            "pipeline": {"what": "whole locate+evaluate pipeline of one step on one GPU (all its kernels)", "algorithmic_bytes_per_step": int(alg_bytes),
                         "bytes_per_query": BYTES_PER_QUERY(n), "distinct_leaves_touched": touched, "kernel_ms_per_step": pipe_ms,
                         "achieved": alg_bytes / (loc_ms * 1e-3) / 1e9, "frac": alg_bytes / (loc_ms * 1e-3) / 1e9 / peak},
            "note": "per-GPU figures (rank 0's shard); kernel times from a separate instrumented pass (csp_timing_*: CUDA events around "
                    "every launch on the launching stream); value/ms_per_step come from the uninstrumented pass"}


def mean_visited(dt, leaf_nodes, n):
    """Mean number of splits coefficients_p examines per leaf (the V of SURVEY 8(d)), measured on a sample of leaves."""
    if n > 64:
        return None
    sample = np.ascontiguousarray(leaf_nodes[:: max(1, len(leaf_nodes) // 4000)], np.int32)
    _, vis = dt.coefficients_p(node_ids=sample)
    return float(vis.mean())


def parity_and_cpu(cb, name, dt, models, nq, nfit, nerr=1_000_000):
    """Rank 0: the oracle (reference algorithm on the host cores) on the SAME tree -- timed as the CPU baseline and compared with
    the GPU's results: a sample of leaf fits bit for bit, nq queries leaf id for leaf id and value for value."""
    from oracle import oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import eval_errors
    oroot, _ = grow_oracle_tree(cb, O, name)
    threads = host_threads()
    lv = O.leaves(oroot)
    nl, n = len(lv), oroot.n
    rng = np.random.default_rng(7)
    sel = np.sort(rng.choice(nl, size=min(nfit, nl), replace=False)).astype(np.int32)
    t0 = time.perf_counter()
    ofit = O.fit_boxes_batch([lv[i] for i in sel], nthreads=threads)
    fit_s = time.perf_counter() - t0
    gfit = models.download_rows(sel)
    fit_equal = all(np.array_equal(gfit[k], ofit[k], equal_nan=True) for k in ("status", "c", "g", "Q", "b"))
    mo = models.download()
    w = oroot.world
    X = w.lower + (w.upper - w.lower) * rng.random((nq, n))
    gleaf, gval, gstatus = dt.locate_eval(models, X)  # GPU, host-buffer path
    Qf = np.ascontiguousarray(np.nan_to_num(mo["Q"]).transpose(0, 2, 1)) if "Q" in mo else np.zeros((nl, n, n))
    oleaf, oval, ostatus, q_s = O.locate_eval_batch(oroot, X, mo["c"], mo["g"], None, mo["b"], nthreads=threads, Qf=Qf)
    del Qf
    k = min(nerr, nq)
    e = eval_errors(gval[:k], oval[:k], X[:k], oleaf[:k], mo["c"], mo["g"], mo.get("Q"), mo["b"])
    fin = np.isfinite(oval)
    parity = {"leaf_id_mismatches": int((gleaf != oleaf).sum()), "status_mismatches": int((gstatus != ostatus).sum()),
              "queries_compared": int(nq), "fit_bits_equal": bool(fit_equal), "leaf_fits_compared": int(len(sel)),
              "nan_pattern_equal": bool(np.array_equal(np.isfinite(gval), fin)),
              "max_abs_err": float(np.abs(gval[fin] - oval[fin]).max()) if fin.any() else 0.0,
              "max_err_over_condition_scale": e["max_scaled"], "max_rel_err": e["max_rel_wellcond"],
              "max_rel_err_note": f"plain |err|/|value| over the well-conditioned values (sum of term magnitudes <= 16 |value|: "
                                  f"{100 * e['frac_wellcond']:.1f}% of {e['n']} compared); tolerance 1e-12",
              "against": "oracle/ (C++ restatement of the Julia reference) on the same tree and inputs"}
    cpu = {"value": nq / q_s, "unit": "queries/s", "cores": threads, "kind": "port", "fits_per_s": len(sel) / fit_s,
           "sample": f"{nq} queries located+evaluated in {q_s:.2f}s and {len(sel)} leaf fits in {fit_s:.2f}s on the same tree; "
                     "C++ restatement of the Julia reference, not Julia"}
    return parity, cpu


# ------------------------------------------------------------------------------------------------------------------------
# secondary configurations
# ------------------------------------------------------------------------------------------------------------------------
def run_single_gpu_config(cb, torch, name, m, steps, local, with_parity=True):
    """One-GPU line for a whole configuration (C2): re-fit + locate/evaluate of m resident queries per step."""
    from csp_b200 import device as cdev
    from csp_b200.device import DeviceTree
    dev = torch.device("cuda", local)
    root, _ = cb.synthetic.config_tree(name)
    flat = root.tree.flat()
    tree = DeviceTree(flat, device=local)
    n, nl = tree.n, tree.n_leaves
    X = torch.empty((m, n), dtype=torch.float64, device=dev)
    fill_uniform(torch, X, 4321)
    leaf = torch.empty(m, dtype=torch.int32, device=dev)
    val = torch.empty(m, dtype=torch.float64, device=dev)
    status = torch.empty(m, dtype=torch.uint8, device=dev)
    models = tree.fit_models()

    def step(evs=None):
        if evs: evs[0].record()
        tree.refit_models(models)
        if evs: evs[1].record()
        tree.locate_eval(models, X, leaf=leaf, val=val, status=status)
        if evs: evs[2].record()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for k in range(steps):
        step(evs[k])
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    torch.cuda.synchronize()
    ms = evs[0][0].elapsed_time(end) / steps
    fit_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / steps
    loc_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / steps
    # per-kernel times: one chunk in flight, so that every kernel is timed alone (for 8 <= n <= 12 the timed pass above keeps two
    # chunks in flight on two internal streams: the descent of one shares the SMs with the evaluation of the other)
    prev_streams = os.environ.get("CSP_BIN_STREAMS")
    os.environ["CSP_BIN_STREAMS"] = "1"
    cdev.timing_read(); cdev.timing_enable(True)
    for k in range(steps):
        step()
    KT = cdev.timing_read(); cdev.timing_enable(False)
    if prev_streams is None:
        os.environ.pop("CSP_BIN_STREAMS")
    else:
        os.environ["CSP_BIN_STREAMS"] = prev_streams
    touched = int(torch.unique(leaf[: min(m, 20_000_000)]).numel())
    peak, _ = peaks()
    V = mean_visited(tree, flat["leaf_nodes"], n)
    out = {"workload": WORKLOADS.get(name, name), "value": m / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms, "queries_per_step": m,
           "leaves": nl, "n": n, "locate_eval_ms": loc_ms,
           "roofline": roofline_block(name, KT, steps, m, n, tree.p, touched, tree.kib * 1024, loc_ms),
           "fits": {"value": nl / (fit_ms * 1e-3), "unit": "leaf fits/s", "ms": fit_ms, "kernel_ms": KT["fit"][0] / steps,
                    "mean_splits_examined": V, "algorithmic_bytes_per_leaf": fit_bytes(n, V) if V else None,
                    "frac": nl * fit_bytes(n, V) / (max(KT["fit"][0] / steps, 1e-9) * 1e-3) / 1e9 / peak if V else None}}
    if out["roofline"] and 8 <= n <= 12:
        out["roofline"]["note"] += ("; per-kernel times with one chunk in flight (CSP_BIN_STREAMS=1): their sum exceeds locate_eval_ms, "
                                    "which overlaps two chunks")
    if with_parity:
        try:
            out["parity"], out["cpu_baseline"] = parity_and_cpu(cb, name, tree, models, nq=2_000_000, nfit=nl)
        except Exception as ex:
            out["parity"] = {"error": str(ex)[:200]}
    del X, leaf, val, status
    models.close(); tree.close()
    torch.cuda.empty_cache()
    return out


def run_c4_dense(cb, torch, m, steps, local):
    """C4: CS1 tree, n=100, 1e4 leaves, models supplied; locate (streaming descent) + grouping + dense evaluation on the
    FP64 tensor cores (K3b)."""
    from csp_b200 import device as cdev
    from csp_b200.device import DeviceTree, LeafModels
    dev = torch.device("cuda", local)
    root, _ = cb.synthetic.config_tree("C4")
    tree = DeviceTree(root.tree.flat(), device=local)
    n, nl = tree.n, tree.n_leaves
    rng = np.random.default_rng(0)
    c, g, b = rng.standard_normal(nl), rng.standard_normal((nl, n)), 0.1 * rng.standard_normal((nl, n))
    Q = np.triu(rng.standard_normal((nl, n, n)) / n)
    models = LeafModels.upload(c, g, Q, b, device=local)
    X = torch.empty((m, n), dtype=torch.float64, device=dev)
    fill_uniform(torch, X, 99)
    leaf = torch.empty(m, dtype=torch.int32, device=dev)
    val = torch.empty(m, dtype=torch.float64, device=dev)
    status = torch.empty(m, dtype=torch.uint8, device=dev)
    run = lambda: tree.locate_eval(models, X, leaf=leaf, val=val, status=status)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    cdev.timing_read(); cdev.timing_enable(True)
    for _ in range(steps):
        run()
    kt = cdev.timing_read(); cdev.timing_enable(False)
    flop_q = 2 * n * n + 5 * n + 2
    eval_ms = kt["eval_binned"][0] / steps if kt["eval_binned"][1] else None
    fp64_peak = 37.1  # TFLOP/s, DMMA pipe, measured on this pool's B200 by tools/ubench/fp64_peaks.cu (profiles/r1_fp64_peaks.txt)
    out = {"workload": "C4: CS1 tree (p=1), n=100, 1e4 leaves, models supplied; dense x'Qx on the FP64 tensor cores", "queries_per_step": m,
           "value": m / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms, "useful_tflops_end_to_end": m * flop_q / (ms * 1e-3) / 1e12,
           "kernel_ms_per_step": {k: v[0] / steps for k, v in kt.items() if v[1]},
           "dense_eval": None if not eval_ms else {"kernel": "k_eval_dmma<64,NTJ,4>", "ms": eval_ms, "useful_tflops": m * flop_q / (eval_ms * 1e-3) / 1e12,
                                                   "bound": "tensor (FP64 DMMA)", "peak_tflops": fp64_peak,
                                                   "frac_useful": m * flop_q / (eval_ms * 1e-3) / 1e12 / fp64_peak,
                                                   "peak_source": "tools/ubench/fp64_peaks.cu on this pool (builder-measured; MEASURED_PEAKS.json has no FP64 entry)"}}
    del X, leaf, val, status
    models.close(); tree.close()
    torch.cuda.empty_cache()
    return out


def run_c5_fits(cb, torch, shard, rank, world, steps, cur, dist):
    """C5: CS2 tree, n=300, ~1e4 leaves: all-leaf fits sharded by leaf-id range over the GPUs + all-gather of the records."""
    blob = None
    if rank == 0:
        root, _ = cb.synthetic.config_tree("C5")
        blob = cb.pack_blob(root.tree.flat())
    st = shard.ShardedTree(blob, root=0)
    sm = st.fit(streams=[cur])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(steps):
        st.fit(models=sm, streams=[cur])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    t = torch.tensor([ms], dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    tm = shard.timing()
    out = {"workload": "C5: CS2 tree (p=2), n=300, 23 addpoint!s (structurally under-determined, SURVEY M5); all-leaf fits sharded by "
                       "leaf-id range over the GPUs, fitted records all-gathered", "leaves": st.n_leaves, "n": st.n, "n_gpus": world,
           "value": st.n_leaves / (ms * 1e-3), "unit": "leaf fits/s", "ms_per_fit_pass": ms, "fit_kernel_ms_rank0": tm["fit_ms"],
           "allgather_ms_rank0": tm["allgather_ms"] if world > 1 else 0.0,
           "record_bytes": 8 * ((st.n + (st.n + 1) * (st.n + 2) // 2 + 3) // 4 * 4)}
    sm.close(); st.close()
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------------------------
def main():
    capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3")
    ap.add_argument("--queries", type=int, default=None, help="queries per step, whole job (default 1e9 for C3, 1e8 for C2)")
    ap.add_argument("--chunk", type=int, default=100_000_000, help="resident chunk per GPU (queries)")
    ap.add_argument("--resident-chunks", type=int, default=4, help="distinct resident chunks per GPU when a shard exceeds them (cycled)")
    ap.add_argument("--gather-chunk", type=int, default=25_000_000, help="queries per gather pipeline step (also the binning chunk)")
    ap.add_argument("--e2e-queries", type=int, default=None)
    ap.add_argument("--ref-queries", type=int, default=20_000_000)
    ap.add_argument("--ref-fits", type=int, default=100_000)
    ap.add_argument("--ref-step-queries", type=int, default=5_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import csp_b200 as cb
    from csp_b200 import device as cdev
    from csp_b200 import shard

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
        dist.all_reduce(torch.zeros(1, device=torch.device("cuda", local)))
        info = shard.init_from_torch_distributed(local)
    else:
        info = shard.init_single_process(1, [local])
    dev = torch.device("cuda", local)
    cur = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    # ---- tree: grown on rank 0 (host), flattened, broadcast ONCE as a blob, uploaded per GPU from the device buffer
    name = args.config
    t0 = time.perf_counter()
    blob, flat = None, None
    if rank == 0:
        root, _ = cb.synthetic.config_tree(name)
        flat = root.tree.flat()
        blob = cb.pack_blob(flat)
    grow_s = time.perf_counter() - t0
    # warm the communicator (connection set-up is not part of the broadcast time)
    if world > 1:
        tiny = cb.pack_blob(cb.synthetic.config_tree("C1")[0].tree.flat()) if rank == 0 else None
        shard.ShardedTree(tiny, root=0).close()
    barrier()
    st = shard.ShardedTree(blob, root=0)  # C1: the one collective per tree version
    tm_tree = shard.timing()
    tree = st.local[0]
    n, nl = tree.n, tree.n_leaves
    setup_s = time.perf_counter() - t0

    # ---- queries: ONE batch of M per step, sharded by contiguous ranges; resident in chunks of <= args.chunk per GPU
    M = args.queries or DEFAULT_QUERIES.get(name, 100_000_000)
    counts = shard.shard_counts(M, world)
    mine = counts[rank]
    chunk = max(1, min(args.chunk, max(counts)))
    ncalls = -(-max(counts) // chunk)                       # identical on every rank (collective calls)
    nres = min(ncalls, max(1, args.resident_chunks))
    if ncalls * chunk * n * 8 <= 100e9:                     # the whole shard fits: every query of the batch is distinct and resident
        nres = ncalls
    Xres = []
    for k in range(nres):
        rows = max(0, min(chunk, mine - k * chunk)) if nres == ncalls else min(chunk, mine)
        Xk = torch.empty((max(rows, 1), n), dtype=torch.float64, device=dev)
        fill_uniform(torch, Xk, 1234 + 131 * rank + k)
        Xres.append(Xk)
    status = torch.empty(max(mine, 1), dtype=torch.uint8, device=dev)
    gather = world > 1
    leaf_all = val_all = None
    gather_path = "none"
    if gather:
        try:     # library-owned gather buffers: the blocks move with the copy engines (peer DMA over NVLink), no SMs involved
            leaf_all, val_all = shard.gather_alloc(M, root=0, device=local)
            gather_path = "copy engines (csp_shard_gather_alloc: device-to-device DMA over NVLink into rank 0's arrays; NCCL carries one completion word per call)"
        except cb.CspError:
            if rank == 0:
                leaf_all = torch.empty(M, dtype=torch.int32, device=dev)
                val_all = torch.empty(M, dtype=torch.float64, device=dev)
            gather_path = "NCCL send/recv kernels"
        if os.environ.get("CSP_GATHER") == "nccl":
            gather_path = "NCCL send/recv kernels (CSP_GATHER=nccl)"
    inplace = gather and rank == 0   # rank 0 computes straight into its slots of the gathered arrays: no copy of its own part
    leaf = None if inplace else torch.empty(max(mine, 1), dtype=torch.int32, device=dev)
    val = None if inplace else torch.empty(max(mine, 1), dtype=torch.float64, device=dev)
    sm = st.fit(streams=[cur])
    torch.cuda.synchronize()

    # per call j: every rank processes queries [j*chunk, (j+1)*chunk) of its shard; rank 0 receives them chunk-major
    call_counts = [[max(0, min(chunk, c - j * chunk)) for c in counts] for j in range(ncalls)]
    call_base = [sum(sum(cc) for cc in call_counts[:j]) for j in range(ncalls)]

    def out_slices(j):
        """(leaf, val) output buffers of call j on this rank."""
        if inplace:
            return leaf_all[call_base[j]:], val_all[call_base[j]:]
        return leaf[j * chunk:], val[j * chunk:]

    if gather:
        shard.defer_join(True)   # the calls of a step write disjoint result ranges: one join per step instead of one per call

    def locate_all(with_gather):
        for j in range(ncalls):
            cc = call_counts[j]
            o = j * chunk
            Xj = Xres[j % nres]
            groot = 0 if (with_gather and gather) else -1
            la = leaf_all[call_base[j]:] if (groot == 0 and rank == 0) else None
            va = val_all[call_base[j]:] if (groot == 0 and rank == 0) else None
            lj, vj = out_slices(j)
            st.locate_eval_resident(sm, [Xj], cc, [lj], [vj], [status[o:]], gather_root=groot, leaf_all=la, val_all=va,
                                    chunk=args.gather_chunk, streams=[cur], gather_base=call_base[j])
        if with_gather and gather:
            shard.join([cur])

    def step(evs=None, with_gather=True):
        if evs: evs[0].record()
        st.fit(models=sm, streams=[cur])                     # sharded re-fit + all-gather of the records
        if evs: evs[1].record()
        locate_all(with_gather)
        if evs: evs[2].record()

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
    total_ms, fit_ms, loc_ms = maxranks([total_ms, fit_ms, loc_ms])
    ms_per_step = total_ms / args.steps
    value = M / (ms_per_step * 1e-3)
    tm_fit = shard.timing()

    # ---- the same without the gather, and the gather alone (explain the headline; not the headline)
    nx = max(2, min(args.steps, 5))
    nogather_ms, gather_ms = ms_per_step, 0.0
    if gather:
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(nx):
            step(with_gather=False)
        b.record()
        barrier()
        nogather_ms = maxranks([a.elapsed_time(b) / nx])[0]
        a.record()
        for _ in range(nx):
            for j in range(ncalls):
                lj, vj = out_slices(j)
                st.gather_resident([lj], [vj], call_counts[j], 0, leaf_all[call_base[j]:] if rank == 0 else None,
                                   val_all[call_base[j]:] if rank == 0 else None, streams=[cur], gather_base=call_base[j])
        b.record()
        barrier()
        gather_ms = maxranks([a.elapsed_time(b) / nx])[0]
    # the gathered arrays hold what the owners computed: every rank's checksums of its own results against rank 0's copy
    gather_ok = None
    if gather:
        lj, vj = out_slices(0)
        k0 = call_counts[0][rank]
        mine_ck = torch.stack([lj[:k0].sum(dtype=torch.int64).double(), vj[:k0].nan_to_num().sum()])
        cks = [torch.zeros_like(mine_ck) for _ in range(world)]
        dist.all_gather(cks, mine_ck)
        if rank == 0:
            offs = np.concatenate([[0], np.cumsum(call_counts[0])])
            gather_ok = all(bool(torch.equal(torch.stack([leaf_all[offs[r]:offs[r + 1]].sum(dtype=torch.int64).double(),
                                                          val_all[offs[r]:offs[r + 1]].nan_to_num().sum()]), cks[r])) for r in range(world))
        lj = vj = None

    # ---- instrumented pass: per-kernel device times (rank-local)
    prev_streams = os.environ.get("CSP_BIN_STREAMS")
    os.environ["CSP_BIN_STREAMS"] = "1"  # every kernel timed alone (the default overlaps two chunks for 8 <= n <= 12)
    cdev.timing_read()
    cdev.timing_enable(True)
    ni = max(1, min(args.steps, 3))
    for k in range(ni):
        step(with_gather=False)
    KT = cdev.timing_read()
    cdev.timing_enable(False)
    if prev_streams is None:
        os.environ.pop("CSP_BIN_STREAMS")
    else:
        os.environ["CSP_BIN_STREAMS"] = prev_streams
    barrier()

    l0, v0 = out_slices(0)
    k0 = min(call_counts[0][rank], 20_000_000)
    touched = int(torch.unique(l0[:k0]).numel()) if k0 else 0
    bad = int((status[:mine] != 0).sum().item())
    nanvals = int(torch.isnan(v0[:call_counts[0][rank]]).sum().item())

    # ---- e2e: the same metric through the host-buffer C-ABI call, pinned host memory, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        me = args.e2e_queries or min(mine, max(1, int(8e9 // (8 * n))))  # ~8 GB of pinned coordinates per rank
        try:
            Xh = torch.empty((me, n), dtype=torch.float64, pin_memory=True)
            Xh.copy_(Xres[0][:me])
            lh = torch.empty(me, dtype=torch.int32, pin_memory=True)
            vh = torch.empty(me, dtype=torch.float64, pin_memory=True)
            sh = torch.empty(me, dtype=torch.uint8, pin_memory=True)
            Xn, ln, vn, sn = Xh.numpy(), lh.numpy(), vh.numpy(), sh.numpy()
            esteps = max(2, min(args.steps, 5))
            st.locate_eval(sm, Xn, leaf=ln, val=vn, status=sn)  # warm-up
            barrier()
            t0e = time.perf_counter()
            for _ in range(esteps):
                st.fit(models=sm, streams=[cur])
                torch.cuda.synchronize()
                st.locate_eval(sm, Xn, leaf=ln, val=vn, status=sn)  # csp_shard_locate_eval: returns when the results are in host memory
            torch.cuda.synchronize()
            dt = maxranks([time.perf_counter() - t0e])[0]
            same = bool(np.array_equal(ln[:100000], l0[:100000].cpu().numpy()))  # the same queries as the first resident chunk
            e2e = {"value": world * me * esteps / dt, "unit": "queries/s", "h2d_bytes_per_step": int(world * me * n * 8),
                   "d2h_bytes_per_step": int(world * me * 13), "steps": esteps, "queries_per_step": int(world * me),
                   "queries_per_step_per_gpu": int(me), "matches_resident_path": same,
                   "api": "csp_shard_fit + csp_shard_locate_eval (host buffers, pinned; results land in host arrays)",
                   "note": "a bounded sample of the batch per step (pinned host memory is the limit); throughput is per query"}
            del Xh, lh, vh, sh
        except Exception as ex:  # e.g. not enough pinnable host memory
            e2e = {"value": None, "unit": "queries/s", "error": str(ex)[:200]}

    # ---- secondary: C5 fits sharded over the same GPUs (collective: every rank takes part)
    c5 = None
    if not args.no_secondary:
        del Xres, leaf_all, val_all, l0, v0, leaf, val
        Xres = None
        if gather:
            shard.gather_free()
        torch.cuda.empty_cache()
        try:
            c5 = run_c5_fits(cb, torch, shard, rank, world, max(2, min(args.steps, 5)), cur, dist)
        except Exception as ex:
            c5 = {"error": str(ex)[:300]}

    if rank != 0:
        dist.barrier()  # rank 0 still uses its replicas for the parity / CPU-baseline leg
        sm.close()
        st.close()
        shard.shutdown()
        dist.destroy_process_group()
        return

    # ---- rank 0: roofline, parity + CPU baseline, secondary single-GPU configurations
    peak, peak_src = peaks()
    roofline = roofline_block(name, KT, ni, mine, n, tree.p, touched, tree.kib * 1024, loc_ms)
    V = mean_visited(tree, flat["leaf_nodes"], n)
    fit_kernel_ms = KT["fit"][0] / ni
    lo_r0, hi_r0 = shard.shard_range(nl, world, 0)
    ag_ms = tm_fit["allgather_ms"] if world > 1 else 0.0
    ag_overlapped = world > 1 and os.environ.get("CSP_ALLGATHER") != "nccl"
    fit_total_ms = fit_ms + (ag_ms if ag_overlapped else 0.0)  # with the copy-engine all-gather fit_ms (stream time) holds the kernels only
    fits = {"value": nl / (fit_total_ms * 1e-3), "unit": "leaf fits/s", "what": "all leaves fitted ONCE per step, sharded by leaf-id range over the GPUs, "
            "records all-gathered (value = leaves / (fit kernels + all-gather), counted serially although the copy-engine all-gather "
            "runs beside the next batch's descent)", "leaves": nl, "ms_per_step": fit_total_ms, "stream_ms_per_step": fit_ms,
            "allgather": ("copy engines: every rank pushes its record range to every other rank (peer DMA over NVLink), overlapped"
                          if ag_overlapped else ("NCCL broadcasts on the step's stream" if world > 1 else "single GPU")),
            "kernel_ms_rank0": fit_kernel_ms, "allgather_ms_rank0": ag_ms, "leaves_rank0": hi_r0 - lo_r0,
            "mean_splits_examined": V, "algorithmic_bytes_per_leaf": fit_bytes(n, V) if V else None,
            "frac": (hi_r0 - lo_r0) * fit_bytes(n, V) / (max(fit_kernel_ms, 1e-9) * 1e-3) / 1e9 / peak if V else None,
            "frac_note": "per GPU: rank 0's leaves x SURVEY 8(d) bytes / its fit-kernel time / measured HBM peak",
            "kernel": "k_fit_chain (chaintop + fit_poly1) + k_fit<2> (coefficients_p + fit_poly2_diag + canonicalize)"}
    out = {"metric": "queries/sec (leaf lookup + quadratic eval)", "value": value, "unit": "queries/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(name, M, nl, n),
           "partition": {
               "queries_per_step_per_gpu": int(max(counts)), "leaves_fitted_per_gpu": int(-(-nl // world)),
               "resident_chunk_queries": int(chunk), "calls_per_step": int(ncalls), "distinct_resident_chunks_per_gpu": int(nres),
               "resident": "every query of the batch distinct and resident" if nres == ncalls else
                           f"the shard exceeds HBM: {nres} distinct resident chunks of {chunk} queries cycled over the {ncalls} passes of a step",
               "gather": "(leaf:int32, value:f64) of every query to rank 0 inside the timed region, pipelined in steps of "
                         f"{args.gather_chunk} queries behind the kernels; transport: {gather_path}" if gather else "single GPU: results already on rank 0"},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "fits": fits,
           "locate_eval_ms": loc_ms, "fit_ms": fit_ms,
           "gather_ms": gather_ms, "value_without_gather": M / (nogather_ms * 1e-3), "ms_per_step_without_gather": nogather_ms,
           "gathered_equals_local": gather_ok, "gather_bytes_to_rank0": int(12 * (M - counts[0])) if gather else 0,
           "tree_broadcast_ms": tm_tree["tree_broadcast_ms"] if world > 1 else 0.0, "tree_blob_h2d_ms": tm_tree["blob_h2d_ms"],
           "tree_upload_ms": tm_tree["upload_ms"], "tree_blob_bytes": int(blob.nbytes), "nccl_version": info["nccl_version"],
           "grow_s": grow_s, "setup_s": setup_s, "outside_world": bad, "nan_values": nanvals}
    if c5 is not None:
        out["c5"] = c5

    if not args.no_cpu_baseline:
        try:
            out["parity"], out["cpu_baseline"] = parity_and_cpu(cb, name, tree, sm.local[0], nq=args.ref_queries, nfit=args.ref_fits)
        except Exception as ex:
            out["cpu_baseline"] = {"value": None, "error": str(ex)[:300]}
    sm.close()
    st.close()
    if world > 1:
        dist.barrier()
    shard.shutdown()
    torch.cuda.empty_cache()
    if not args.no_secondary and world == 1:
        for key, fn in (("c2", lambda: run_single_gpu_config(cb, torch, "C2", 100_000_000, max(3, min(args.steps, 10)), local,
                                                             with_parity=not args.no_cpu_baseline)),
                        ("c4", lambda: run_c4_dense(cb, torch, 20_000_000, max(2, min(args.steps, 5)), local))):
            try:
                out[key] = fn()
            except Exception as ex:
                out[key] = {"error": str(ex)[:300]}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
