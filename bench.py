#!/usr/bin/env python
"""bench.py -- throughput of the batched LCSampledFunction evaluation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c2s|c2g|c3|c4|c4b|c5|a1] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic queries (per GPU; weak scaling: the batch per GPU is fixed
as N grows, the model is replicated, queries are independent).  Rank 0 prints ONE JSON line.

Workloads = BASELINE.json configs.  Headline c2 (configs[1]): 3-D cubic scalar function, bootstrap level 5 ->
32,768 leaves / 42,875 samples / K=64, 100M queries per step.  The north star's TARGET config (configs[3], 6-D cubic
mesh-valued, workload c4b) is measured in the same run with its own cpu_baseline, e2e and roofline and reported under
`north_star_config`; c3 (configs[2]) and two skewed variants of c2 follow under `other_workloads`.

  value      queries/s with the batch already resident in HBM, CUDA events on the launching stream, max over ranks;
             N > 1: includes the path's exchange step (every batch gathered on its consumer GPU, locointerpolate_b200/shard.py)
  e2e        the same metric through the host-buffer C-ABI call (loco_eval_batch / loco_eval_batch_ring) with page-locked host
             buffers, H2D + kernels + D2H inside the timed region; beside it the box's own H2D+D2H copy ceiling for the same
             bytes and the rate with PAGEABLE host buffers (what a C++ caller of the facade passes)
  roofline   dominant kernel: the bytes it must move (its inputs + outputs, model data counted once) or the flops it must
             execute / its CUDA-event duration, against the measured HBM peak (MEASURED_PEAKS.json) or the measured fp64 rate;
             `step` = the same for the whole step on compulsory bytes (queries in, results out); frac <= ~1 by construction.
             The cache-less per-query byte count of SURVEY.md 8(d) is kept under `survey_cacheless` for reference only.
  cpu_baseline  the reference's own C++ path (oracle/_ref/liblocoref.so, fork-parallel over the host cores) or, where that
             library is absent, the oracle port, on a bounded sample of the same workload
  spot_check the outputs of the LAST TIMED step compared with the CPU restatement (oracle/) on a sample of its queries
"""
from __future__ import annotations

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

WORKLOADS = {
    # name: N, basis, bootstrap level, D (doubles per value), eval_corners, queries per step per GPU, description
    "c1": dict(N=2, basis=0, level=2, D=1, corners=True, Q=1_000_000,
               desc="2-D linear scalar, TestExamples tree (16 leaves, 25 samples, K=4), 1M queries"),
    "c2": dict(N=3, basis=1, level=5, D=1, corners=True, Q=100_000_000,
               desc="3-D cubic scalar, bootstrap 5 (32,768 leaves, 42,875 samples, K=64), 100M queries"),
    "c2s": dict(N=3, basis=1, level=5, D=1, corners=True, Q=100_000_000, dist="cluster",
                desc="c2 with a SKEWED batch: half of the 100M queries fall into 1 % of the leaves (a box of 0.215^3 of the domain)"),
    "c2g": dict(N=3, basis=1, level=5, D=1, corners=True, Q=100_000_000, dist="sweep",
                desc="c2 with a STRUCTURED batch: 100M queries from a regular 464^3 lattice sweep in scan order (TestDebug.cpp:128-146 style), "
                     "consecutive queries in the same leaf"),
    "c3": dict(N=4, basis=0, level=3, D=30_000, corners=True, Q=262_144,
               desc="4-D linear mesh-valued (10k vertices, D=30,000), bootstrap 3 (4,096 leaves, K=16); 10M queries streamed in 262,144-query batches (63 GB of results per batch)"),
    "c4": dict(N=6, basis=1, level=1, D=150_000, corners=False, Q=16_384,
               desc="6-D cubic mesh-valued (50k vertices, D=150,000), bootstrap 1 (64 leaves, K=4,096); 10M queries streamed in 16,384-query batches (19.7 GB of results per batch)"),
    "c4b": dict(N=6, basis=1, level=2, D=150_000, corners=False, Q=262_144, ring=16_384,
                desc="6-D cubic mesh-valued (50k vertices, D=150,000), bootstrap 2 (4,096 leaves at depth 12, 15,625 value rows = 18.75 GB, K=4,096); "
                     "10M queries streamed in 262,144-query steps (64 per leaf), each sorted once and passed through a 16,384-query ring (19.7 GB) with a per-query checksum"),
    # not a BASELINE config: the reference's own adaptively refined tree (its builder's output, committed as a fixture)
    "a1": dict(N=2, basis=1, level=-1, D=1, corners=True, Q=100_000_000, proto="tests/golden/cub2d_testdebug.pb.gz",
               desc="2-D cubic scalar on the reference's TestDebug tree (TestDebug.cpp:153-157: maxDepth 10, threshold 0.1, bootstrap 2 -> "
                    "724 leaves at levels 4-10, 1,063 samples, 54,419 basis functions), generic term-list kernels, 100M queries"),
    "c5": dict(N=8, basis=0, level=1, D=1, corners=True, Q=64_000_000,
               desc="8-D linear scalar error check, bootstrap 1 (256 leaves, K=256); 1B candidate points per pass in 64M-point batches"),
}
CPU_SAMPLE = {"c1": 2_000_000, "c2": 200_000, "c2s": 200_000, "c2g": 200_000, "c3": 500_000, "c4": 1024, "c4b": 1024, "c5": 100_000, "a1": 600_000}
FP64_PEAK_TFLOPS = 36.4   # measured on this pool: dependent-free DFMA loop (tools/micro/dmma_rate.cu; DMMA m8n8k4: 37.1)
NVLINK_PEER_GBS = 770.0   # B200_PROFILING.md: measured peer copy per direction


def survey_cacheless_bytes_per_query(N, depth, T, K, D):
    """a cache-less single pass of the reference's data flow, counted once (reference figure, NOT a roofline numerator)."""
    return 8 * N + 16 * depth + T * (16 * N + 8 + 4) + K * D * 8 + 8 * D + 5


def test_function(N):
    """deterministic stand-in for LCRealFunction(N, linear=false) (LCRealFunction.cpp:50-63): 1 + sum_i sum_j a_ij sin(3 j x_i + p_ij)"""
    rng = np.random.default_rng(5)
    amp = rng.uniform(-1, 1, (N, 10))
    phase = rng.uniform(-1, 1, (N, 10))
    j = np.arange(10)

    def f(p):
        return 1.0 + float((amp * np.sin(3 * j[None, :] * p[:, None] + phase)).sum())
    return f, amp, phase


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_model(capi, w, device):
    f, amp, phase = test_function(w["N"])
    if w.get("proto"):  # a committed model file (gzip-compressed proto written by the reference's builder)
        import gzip
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".pb", delete=False) as t:
            t.write(gzip.open(os.path.join(ROOT, w["proto"]), "rb").read())
        try:
            return capi.Model.from_proto(t.name, device=device), amp, phase
        finally:
            os.unlink(t.name)
    if w["D"] == 1:
        m = capi.Model.bootstrap(w["N"], w["basis"], w["level"], 1, w["corners"], fn=f, device=device)
    else:
        m = capi.Model.bootstrap(w["N"], w["basis"], w["level"], w["D"], w["corners"], fn=None, device=device)
        if device >= 0:
            m.fill_values_synthetic(0.5)
    return m, amp, phase


_TWINS = {}


def scalar_twin_graph(capi, w):
    """plain-array graph of the workload's tree with scalar values (host-only handle -> proto -> arrays): what the CPU
    restatement / the reference are built from.  A fresh copy per call (callers attach their own values)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import copy
    import protoschema
    import tempfile
    key = (w["N"], w["basis"], w["level"], w["corners"], w.get("proto"))
    if key not in _TWINS:
        m, _, _ = build_model(capi, {**w, "D": 1}, capi.HOST_ONLY)
        tmp = tempfile.mktemp(suffix=".pb")
        try:
            m.save_proto(tmp)
            _TWINS[key] = (protoschema.read_binary(tmp), m)
        finally:
            if os.path.exists(tmp):
                os.unlink(tmp)
    g, m = _TWINS[key]
    return copy.copy(g), m


def make_queries(torch, w, Q, rank):
    """the step's query batch on the device: i.i.d. uniform over the domain, or one of the skewed variants"""
    N = w["N"]
    gen = torch.Generator("cuda").manual_seed(1234 + rank)
    q = torch.rand((Q, N), dtype=torch.float64, device="cuda", generator=gen)
    kind = w.get("dist", "uniform")
    if kind == "cluster":
        # every second query lands in a box holding 1 % of the leaves (interleaved, so that every chunk of the batch is skewed)
        side = 0.01 ** (1.0 / N)
        q[::2] = q[::2] * side + 0.4
    elif kind == "sweep":
        n = int(round(Q ** (1.0 / N)))
        idx = torch.arange(Q, device="cuda", dtype=torch.int64) % (n ** N)
        for d in range(N):  # dimension 0 fastest, like the reference's nested evaluation loops
            q[:, d] = ((idx % n).double() + 0.5) / n
            idx = idx // n
    return q


def cpu_reference_arm(w, sample_q, steps, warmup, name):
    """The reference's own CPU implementation on the host cores (oracle/_ref when present, else the oracle port).
    Mesh-valued workloads: the reference has no mesh value type (LCFunctionValue.cpp:20-31), so a query costs the
    reference's own locate + influencing-set assembly + weights, then `val += w * other` over the D components
    (SURVEY.md 8d); to bound host memory the table has min(D, 3000) columns and the rate is scaled by D_used / D
    (the component loop is linear in D and dominates).
    c4b: the reference's loader needs ~10 min to re-link the 4,096-leaf 6-D cubic tree (measured in this container: 626 s in
    LCSampledFunction::addBorderCells + addAllNeighboursTopDown), more than a bench run may take; the per-query work is timed on the
    bootstrap-1 tree of the same function instead -- the same K = 4,096 influencing samples per query, one basis function each,
    the same D-component sum (only the k-d descent is 6 levels shorter, < 0.1 % of a query)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from locointerpolate_b200 import capi
    from oracle import oracleapi, refapi
    cores = len(os.sched_getaffinity(0))
    N, D = w["N"], w["D"]
    tree_note = ""
    if name == "c4b" and refapi.available():
        w = {**w, "level": 1}
        tree_note = "; timed on the bootstrap-1 tree of the same function (same K=4,096 samples per query; the reference's loader needs ~10 min for the bootstrap-2 tree)"
    g, _ = scalar_twin_graph(capi, w)
    rng = np.random.default_rng(99)
    q = rng.random((sample_q, N))
    if w.get("dist") == "cluster":
        q[::2] = q[::2] * 0.01 ** (1.0 / N) + 0.4
    scale, note = 1.0, ""
    extrapolate = None
    if refapi.available():
        fn = refapi.RefFn.from_graph(g)
        kind, used = "reference", cores
        if D == 1:
            run = lambda: fn.eval_fork(q, cores)
        else:
            Du = min(D, 3000)
            table = rng.standard_normal((fn.S, Du))
            note = (f"; mesh values: the reference's own locate + influencing set + weights (timed as its scalar evaluation) plus the component loop "
                    f"`val += w * other` timed over {Du} of the {D} columns and extrapolated linearly in D")
            run = lambda: fn.eval_mesh_fork(q, table, cores)
            extrapolate = (lambda: fn.eval_fork(q, cores), D / Du)
    else:
        if D != 1:
            raise SystemExit("mesh-valued CPU baseline needs oracle/_ref/liblocoref.so")
        orc = oracleapi.Oracle(g)
        kind, used = "port", 1
        run = lambda: orc.eval(q)
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    if extrapolate is not None:
        # t(D) = t_scalar + (t(Du) - t_scalar) * D / Du : only the component loop grows with D
        scalar_run, factor = extrapolate
        scalar_run()
        t0 = time.perf_counter()
        scalar_run()
        ts = time.perf_counter() - t0
        dt = ts + max(dt - ts, 0.0) * factor
    how = f"fork-parallel over {used} host cores" if kind == "reference" else "single thread"
    return dict(value=sample_q / dt * scale, unit="queries/s", cores=used, kind=kind,
                sample=f"{sample_q} uniform queries of workload {name} per step, {steps} steps after {warmup} warm-up, {how}{note}{tree_note}"), dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--level", type=int, default=-1, help="override the workload's bootstrap level (experiments)")
    ap.add_argument("--queries", type=int, default=0, help="queries per step per GPU (default: the workload's)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU baseline sample (default: sized for ~10-20 s)")
    ap.add_argument("--ecell-poly", type=int, default=-1, help="0/1: term-list / polynomial form of the evaluation cells (adaptive scalar trees; option ecell_poly)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-spot-check", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "ipc", "p2p"],
                    help="multi-GPU exchange transport: peer-mapped buffers + copy engines (ipc), or NCCL point-to-point (p2p)")
    ap.add_argument("--no-extra", action="store_true", help="skip the north-star target config and the secondary workloads")
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 direct, 2 leaf-sorted TMA tiles, 4 L2-resident sorted chunks, 5 leaf buckets")
    ap.add_argument("--overlap", type=int, default=0, help="streams the chunks of the sorted path are dealt over (0 = library default)")
    ap.add_argument("--chunk", type=int, default=0, help="queries per L2-resident chunk of the sorted path (0 = auto)")
    ap.add_argument("--opt", action="append", default=[], help="name=value library option (loco_set_option), repeatable")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.queries:
        w["Q"] = args.queries
    if args.level >= 0:
        w["level"] = args.level
        w["desc"] += f" [bootstrap level overridden to {args.level}]"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "interpolated_queries_per_sec"

    def config_of(name, w):
        return {"workload": f"{name}: {w['desc']}", "n_params": w["N"], "basis": "cubic" if w["basis"] else "linear",
                "value_dim": w["D"], "queries_per_step_per_gpu": w["Q"],
                "parallelism": f"query-sharded x{world}, model replicated" +
                               (f", results of batch i gathered on rank i mod {world} (peer copies, skewed by rank)" if world > 1 and w["D"] == 1 and name != "c5" else ""),
                "l2": "inputs+outputs per step exceed the 126 MB L2" if w["Q"] * (w["N"] + w["D"]) * 8 > 126e6 else "L2 flushed between steps"}
    config = config_of(args.workload, w)

    # ------------------------------------------------------------------ reference arm (CPU) ----------
    if args.impl == "reference":
        if rank != 0:
            return
        base, dt = cpu_reference_arm(w, args.cpu_sample or CPU_SAMPLE[args.workload], max(args.steps, 1), args.warmup, args.workload)
        line = {"impl": "reference", "metric": metric, "value": base["value"], "unit": "queries/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if args.workload == "c2" and not args.no_extra:
            # the north star's target config beside the headline, like the GPU arm
            try:
                wt = dict(WORKLOADS["c4b"])
                bt, dtt = cpu_reference_arm(wt, CPU_SAMPLE["c4b"], max(1, min(args.steps, 2)), min(args.warmup, 1), "c4b")
                line["north_star_config"] = {"value": bt["value"], "unit": "queries/s", "ms_per_step": dtt * 1e3, "config": config_of("c4b", wt), "cpu_baseline": bt}
            except BaseException as e:  # noqa: BLE001 -- a secondary number must never take the headline down
                line["north_star_config"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm (GPU) -----------------
    import torch
    import torch.distributed as dist
    from locointerpolate_b200 import build as product_build, capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if local == 0:
        product_build.build()
    if world > 1:
        dist.barrier()

    line = measure(args.workload, w, args, rank, world, local, config, full=True)
    if not args.no_extra and args.workload == "c2":
        # the north star's TARGET config, measured in full (own cpu_baseline, e2e, roofline, spot check), then the other
        # BASELINE configs and the skewed batches with fewer steps (secondary numbers; the headline stays `config`)
        def guarded(name, full):
            try:
                torch.cuda.empty_cache()
                wx = dict(WORKLOADS[name])
                r = measure(name, wx, args, rank, world, local, config_of(name, wx), full=full, secondary=True)
                return r
            except Exception as e:  # a secondary number must never take the headline down
                return {"error": f"{type(e).__name__}: {e}"[:300]}
        line["north_star_config"] = guarded("c4b", True)
        line["other_workloads"] = {name: guarded(name, False) for name in ("c3", "c2s", "c2g")}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# what each kernel must move per query it processes, given its interface (inputs + outputs; leaf data is model data, counted
# once per launch through `model_bytes`): the numerator of the kernel-level roofline
def kernel_bytes_per_query(kernel, N, D, rec_bytes):
    return {
        "bucket_scatter": 8 * N + rec_bytes + 5,          # query in; record, leaf, status out
        "bucket_eval": rec_bytes + 8,                     # record in; result out
        "bucket_fused": 8 * N + 8 + 5,
        "sorted_count": 8 * N + 5, "sorted_scatter": 8 * N + 4 + rec_bytes, "sorted_eval": rec_bytes + 8,
        "locate_count": 8 * N + 5, "scatter": 8 * N + 4 + rec_bytes,
        "eval_scalar_direct": 8 * N + 8 + 5, "eval_tensor_direct": 8 * N + 8 + 5,
        "mesh_tensor_mma": rec_bytes + 8 * D, "mesh_tensor_tiles": rec_bytes + 8 * D, "mesh_generic_mma": rec_bytes + 8 * D,
        "error_check_generated": 0,
    }.get(kernel)


def min_flops_per_query(name, N, F, D, Kf):
    """fp64 operations the evaluation cannot avoid: the N*F one-dimensional basis polynomials (~6 flops linear / ~12 cubic
    each) and the sum-factorised contraction (scalar: 2 * sum_l F^l) or the K x D weighted sum over the distinct value rows (mesh)"""
    basis = N * F * (12 if F == 4 else 6)
    if D == 1:
        # single-piece leaves (every bootstrap grid): the leaf's polynomial by nested Horner, F^N - 1 FMAs + N local coordinates
        horner = 2 * N + 2 * (F ** N - 1)
        extra = N * (40 + 10 * 6) if name == "c5" else 0   # the truth: one sincos + 10 rotated terms per dimension
        return horner + extra
    return basis + 2.0 * Kf * D


def measure(name, w, args, rank, world, local, config, full, secondary=False):
    """one workload on this rank's GPU: timed steps (max over ranks), per-kernel CUDA-event times, roofline, spot check,
    e2e and cpu_baseline (full only)"""
    import torch
    import torch.distributed as dist
    from locointerpolate_b200 import capi, shard
    steps = args.steps if (full and not secondary) else max(2, min(args.steps, 3))
    warmup = args.warmup if (full and not secondary) else 3
    m, amp, phase = build_model(capi, w, local)
    if not secondary:
        m.set_option("path", args.path)
        m.set_option("chunk", args.chunk)
        if args.overlap:
            m.set_option("overlap", args.overlap)
        for kv in args.opt:
            k, v = kv.split("=")
            m.set_option(k, int(v))
    if args.ecell_poly >= 0:
        m.set_option("ecell_poly", args.ecell_poly)
    info = m.info
    N, D, Q = w["N"], w["D"], w["Q"]
    q = make_queries(torch, w, Q, rank)
    stream = torch.cuda.current_stream().cuda_stream
    leaf = torch.empty(Q, dtype=torch.int32, device="cuda")
    status = torch.empty(Q, dtype=torch.uint8, device="cuda")
    flush = None
    if Q * (N + D) * 8 <= 126e6:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ex = None          # multi-GPU exchange of scalar results
    last_out = {}      # where the last step's results live (spot check)

    if name == "c5":
        err = torch.empty(info["n_leaves"], dtype=torch.float64, device="cuda")
        split = torch.empty(info["n_leaves"], dtype=torch.uint8, device="cuda")
        d_amp, d_phase = torch.from_numpy(amp).cuda(), torch.from_numpy(phase).cuda()
        ppl = Q // info["n_leaves"]
        Q = ppl * info["n_leaves"]

        def step(i):
            m.error_check_generated_device(ppl, 1000 + i, d_amp.data_ptr(), d_phase.data_ptr(), 10, 0.05, err.data_ptr(), split.data_ptr(), stream)
            if world > 1:  # error maxima: all_reduce(MAX) of the bit patterns + split flags (shard.py)
                shard.reduce_error_maxima(err, split)
    elif D == 1:
        if world > 1:
            ex = shard.RotatingExchange(Q, torch.float64, torch.device("cuda", local), transport=args.exchange)
            ex.warm()
        else:
            outs = [torch.empty(Q, dtype=torch.float64, device="cuda") for _ in range(2)]

        def step(i):
            out = ex.out_buffer(i) if ex is not None else outs[i % 2]
            m.eval_device(q.data_ptr(), Q, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), stream)
            if ex is not None:
                ex.submit(i)
            last_out["t"] = out
    elif w.get("ring"):
        # mesh-valued, streamed: results pass through a ring in leaf order and are reduced to one checksum per query
        ring = torch.empty((w["ring"], D + (D & 1)), dtype=torch.float64, device="cuda")
        chk = torch.empty(Q, dtype=torch.float64, device="cuda")

        def step(i):
            m.eval_ring_device(q.data_ptr(), Q, ring.data_ptr(), w["ring"], chk.data_ptr(), leaf.data_ptr(), status.data_ptr(), stream)
            last_out["t"] = chk
    else:
        # mesh-valued: the whole batch of results stays in HBM ([Q x D] doubles; sized for the 180 GB part)
        assert D % 2 == 0, "bench assumes an even value dimension"
        mesh_out = torch.empty((Q, D), dtype=torch.float64, device="cuda")

        def step(i):
            m.eval_device(q.data_ptr(), Q, mesh_out.data_ptr(), leaf.data_ptr(), status.data_ptr(), stream)
            last_out["t"] = mesh_out

    sampler = ClockSampler(local)
    if rank == 0 and full and not secondary:
        sampler.start()  # from the warm-up on: short timed regions would otherwise fall between two 100 ms samples
    for i in range(warmup):
        if flush is not None:
            flush.zero_()
        step(i)
        if w.get("dist"):
            # skewed batches: the library picks its sort path from the counters of the batches it has FINISHED (no synchronisation
            # inside the call); a real stream of batches has that history, a back-to-back enqueue of the warm-up would not
            torch.cuda.synchronize()
    if ex is not None:
        ex.drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = m.kernel_launches
    t_wall = time.perf_counter()
    if flush is None:
        # inputs + outputs of a step exceed the L2: time the K steps back to back (N > 1: the exchange of every batch included)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        if ex is not None:
            ex.drain()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    else:
        # small batch: flush the L2 (write 256 MB) before every step, outside the timed events
        ms = 0.0
        for i in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(i)
            if ex is not None:
                ex.drain()
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 and full and not secondary else None
    launches = m.kernel_launches - launches0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / steps
    value = world * Q / (ms_per_step * 1e-3)
    bad = int((status != 0).sum().item()) if name != "c5" else 0

    # ---- spot check of the LAST TIMED step's outputs against the CPU restatement (rank 0; the oracle is the checker only)
    spot = None
    if rank == 0 and not args.no_spot_check:
        try:
            spot = spot_check(name, w, m, q, last_out.get("t"), leaf, status, amp, phase, capi, torch,
                              (ppl, 1000 + steps - 1, err, split) if name == "c5" else None)
        except Exception as e:  # the check must not hide the measurement; a failed check is reported as such
            spot = {"ok": False, "error": f"{type(e).__name__}: {e}"[:300]}

    # ---- every kernel of the step, CUDA events around each launch on the launching stream (no flush, no exchange)
    #      (measured with the chunks of the scalar paths on ONE stream: with the default multi-stream overlap the events of
    #      a kernel would also cover the kernels running beside it)
    m.set_option("overlap", 1)
    m.profile_begin()
    n_prof = min(steps, 3)
    for i in range(n_prof):
        if name == "c5":
            m.error_check_generated_device(ppl, 1000 + i, d_amp.data_ptr(), d_phase.data_ptr(), 10, 0.05, err.data_ptr(), split.data_ptr(), stream)
        elif ex is not None:
            m.eval_device(q.data_ptr(), Q, ex.local[0].data_ptr(), leaf.data_ptr(), status.data_ptr(), stream)
        else:
            step(i)
    prof = m.profile_end()
    m.set_option("overlap", args.overlap or 3)
    tot_ms = sum(v["total_ms"] for v in prof.values()) or 1.0
    kernels = {k: {"launches_per_step": v["launches"] / n_prof, "ms_per_launch": v["total_ms"] / v["launches"], "share": v["total_ms"] / tot_ms}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["total_ms"])}
    dom = next(iter(kernels))
    dom_ms = kernels[dom]["ms_per_launch"]
    q_per_launch = Q / kernels[dom]["launches_per_step"]
    T = info["n_terms"] / info["n_leaves"]
    K = info["n_support_entries"] / info["n_leaves"]
    Kf = info["n_fold_entries"] / info["n_leaves"] if info.get("n_fold_entries") else K
    F = 4 if w["basis"] else 2
    peak, peak_src = measured_peak()
    rec_bytes = ((N * 8 + 4 + 31) // 32) * 32
    step_kernel_ms = tot_ms / n_prof
    # the two floors of the whole step: bytes that must cross HBM (queries in, results + leaf + status out, model once) and
    # fp64 operations that must execute
    compulsory = (Q * (8 * N + 8 * D + 5) if name != "c5" else 0) + info["device_bytes"]
    flops_q = min_flops_per_query(name, N, F, D, Kf)
    t_hbm = compulsory / (peak * 1e9) * 1e3
    t_fp = flops_q * Q / (FP64_PEAK_TFLOPS * 1e12) * 1e3
    bound = "hbm" if t_hbm >= t_fp else "tensor"
    # kernel-level numerator for the dominant kernel
    kb = kernel_bytes_per_query(dom, N, D, rec_bytes)
    dom_model_bytes = info["device_bytes"] if dom.startswith("mesh_") else 0   # mesh kernels stream the value table once per launch
    traffic, traffic_note = traffic_for(name, dom, q_per_launch)
    if kb is None:
        kb = 8 * N + 8 * D + 5   # a kernel without an entry: the step's compulsory bytes per query
    if bound == "hbm":
        algo = kb * q_per_launch + dom_model_bytes
        achieved, unit, pk = algo / (dom_ms * 1e-3) / 1e9, "GB/s", peak
    else:
        algo = flops_q * q_per_launch
        achieved, unit, pk = algo / (dom_ms * 1e-3) / 1e12, "TFLOP/s", FP64_PEAK_TFLOPS
    roofline = {"bound": bound, "kernel": dom, "achieved": achieved, "peak": pk, "unit": unit, "frac": achieved / pk, "traffic": traffic,
                "peak_source": peak_src if unit == "GB/s" else "measured fp64 rate on this pool (tools/micro/dmma_rate.cu: DFMA 36.4, DMMA m8n8k4 37.1 TFLOP/s); "
                               "'tensor' = the fp64 DMMA/DFMA pipe, no reduced-precision tensor-core math is used",
                "kernel_ms_per_launch": dom_ms, "queries_per_launch": q_per_launch,
                "numerator": (f"{kb} B/query the kernel must read+write (its inputs and outputs)" + (" + the value table once" if dom_model_bytes else "")) if unit == "GB/s"
                             else f"{flops_q:.0f} fp64 flops/query (basis polynomials + contraction over the {Kf:.0f} distinct value rows per leaf)",
                "traffic_note": traffic_note,
                "step": {"kernels_ms_per_step": step_kernel_ms, "timed_ms_per_step": ms_per_step,
                         "hbm": {"compulsory_bytes_per_step": compulsory, "floor_ms": t_hbm, "frac": t_hbm / ms_per_step,
                                 "achieved_gbs": compulsory / (ms_per_step * 1e-3) / 1e9},
                         "fp64": {"flops_per_query": flops_q, "floor_ms": t_fp, "frac": t_fp / ms_per_step,
                                  "achieved_tflops": flops_q * Q / (ms_per_step * 1e-3) / 1e12, "peak_tflops": FP64_PEAK_TFLOPS},
                         "bound": bound, "frac": max(t_hbm, t_fp) / ms_per_step},
                "survey_cacheless": {"bytes_per_query": survey_cacheless_bytes_per_query(N, info["depth"], T, K, D),
                                     "note": "SURVEY.md 8(d): the reference's per-query data flow without any cache or reuse; leaf data is served "
                                             "from shared memory / L2 here, so this is NOT a roofline numerator"},
                "results_written_gbs": (Q * 8 * D / (ms_per_step * 1e-3) / 1e9) if D > 1 and not w.get("ring") else None}

    result = {"value": value, "unit": "queries/s", "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup, "queries_per_step_per_gpu": Q,
              "n_gpus": world, "gpu_launches": int(launches), "kernels": kernels, "roofline": roofline, "bad_status": bad, "spot_check": spot,
              "model": {k: info[k] for k in ("n_leaves", "n_samples", "n_terms", "max_support", "depth", "device_bytes")}}
    if config is not None and secondary:
        result["config"] = config
    if ex is not None:
        result["exchange"] = exchange_stats(ex, m, q, Q, leaf, status, stream, torch, world, ms_per_step)

    if not full:
        result["workload"] = f"{name}: {w['desc']}"
        if ex is not None:
            ex.close()
        del q, leaf, status
        m.free()
        return result

    # ---- end to end through the host-buffer C-ABI call (page-locked host memory; H2D + kernels + D2H inside)
    e2e = None
    if name != "c5" and not args.no_e2e:
        e2e = measure_e2e(name, w, m, q, Q, torch, dist, world, min(steps, 3))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu, _ = cpu_reference_arm(w, args.cpu_sample or CPU_SAMPLE[name], 1, 1, name)
        except SystemExit as e:
            cpu = {"value": None, "unit": "queries/s", "cores": 0, "kind": "port", "sample": str(e)}

    if secondary:
        result.update({"e2e": e2e, "cpu_baseline": cpu, "dtype": "f64", "data": "synthetic"})
        line = result
    else:
        line = {"metric": "interpolated_queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "kernels": kernels, "wall_s": wall, "bad_status": bad, "spot_check": spot, "model": result["model"]}
        if ex is not None:
            line["exchange"] = result["exchange"]
    if ex is not None:
        ex.close()
    del q, leaf, status
    m.free()
    return line


def traffic_for(name, kernel, q_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the ncu capture recorded in
    profiles/traffic.json -- only if that capture was taken from the SAME kernel sources (build digest), else null"""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tpath):
        return None, "no ncu capture recorded"
    try:
        from locointerpolate_b200 import build as product_build
        tj = json.load(open(tpath))
        # entries are keyed "<workload>" or "<workload>_<anything>" (one per captured kernel of that workload)
        ent = next((v for k, v in tj.items() if (k == name or k.startswith(name + "_")) and v.get("kernel") == kernel), {})
        if ent.get("kernel") != kernel:
            return None, f"no ncu capture of {kernel} for workload {name}"
        digest = product_build.source_digest()[:16]
        if ent.get("source_digest") != digest:
            return None, f"stale: the recorded ncu capture is of source digest {ent.get('source_digest')}, the library is {digest}"
        return ent["dram_bytes_per_query"] * q_per_launch, f"ncu --set full, {ent.get('capture', '')} (source digest {digest})"
    except Exception as e:
        return None, f"traffic.json unreadable: {e}"


def spot_check(name, w, m, q, out_t, leaf, status, amp, phase, capi, torch, c5):
    """compare the last timed step's device outputs with the CPU restatement on a sample of its queries"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import oracleapi
    N, D = w["N"], w["D"]
    if name == "c5":
        # the fused kernel's per-leaf maxima against the three-kernel form on the same generated points
        ppl, seed, err, split = c5
        a, sa = err.clone(), split.clone()
        d_amp, d_phase = torch.from_numpy(amp).cuda(), torch.from_numpy(phase).cuda()
        m.set_option("fused_errgen", 0)
        m.error_check_generated_device(ppl, seed, d_amp.data_ptr(), d_phase.data_ptr(), 10, 0.05, err.data_ptr(), split.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        m.set_option("fused_errgen", 1)
        diff = float((a - err).abs().max())
        return {"ok": bool(diff <= 1e-12 and torch.equal(sa, split)), "leaves": int(a.numel()), "max_abs_diff": diff,
                "against": "the generate / evaluate / reduce kernels on the same points (those are pinned to the CPU restatement in tests/test_gpu_baseline_shapes.py)"}
    g, twin = scalar_twin_graph(capi, w)
    n = 4096 if D == 1 else 192
    Q = q.shape[0]
    pick = (torch.arange(n, device="cuda", dtype=torch.int64) * ((Q - 1) // max(n - 1, 1))).clamp_(max=Q - 1)
    qs = q[pick].cpu().numpy()
    got_leaf, got_status = leaf[pick].cpu().numpy(), status[pick].cpu().numpy()
    if D == 1:
        orc = oracleapi.Oracle(g)
        o_out, o_leaf, o_status, scale = orc.eval(qs)
        got = out_t[pick].cpu().numpy()
        okm = o_status == 0
        rel = float(np.max(np.abs(got[okm] - o_out[okm]) / np.maximum(np.abs(o_out[okm]), scale[okm]))) if okm.any() else 0.0
        ok = bool(np.array_equal(got_leaf, o_leaf) and np.array_equal(got_status, o_status) and rel <= 1e-12 and np.isnan(got[~okm]).all())
        return {"ok": ok, "queries": n, "max_rel_diff": rel, "leaf_equal": bool(np.array_equal(got_leaf, o_leaf)), "against": "oracle/loco_oracle.cpp (CPU restatement)"}
    # mesh-valued: the restatement on 8 sampled components of the device's own table (first 4 + last 4), the same queries
    # evaluated once more with kept results; the timed ring checksum against the checksum of those kept results
    cols = [(0, 4), (D - 4, 4)]
    rows = m.sample_value_rows()
    tab = np.concatenate([m.get_value_columns(c0, nc) for c0, nc in cols], axis=1)
    g.D = tab.shape[1]
    g.sample_value = tab[rows]
    g.sample_value_group = None
    g.cell_sample_value = g.border_sample_value = g.cell_center_value = g.border_center_value = None
    orc = oracleapi.Oracle(g.normalise())
    o_out, o_leaf, o_status, scale = orc.eval(qs)
    dq = torch.from_numpy(qs).cuda()
    kept = torch.empty((n, D), dtype=torch.float64, device="cuda")
    m.eval_device(dq.data_ptr(), n, kept.data_ptr(), 0, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = torch.cat([kept[:, c0:c0 + nc] for c0, nc in cols], dim=1).cpu().numpy()
    okm = o_status == 0
    rel = float(np.max(np.abs(got[okm] - o_out[okm]) / np.maximum(np.abs(o_out[okm]), scale[okm][:, None])))
    res = {"queries": n, "components": tab.shape[1], "max_rel_diff": rel, "leaf_equal": bool(np.array_equal(got_leaf, o_leaf)),
           "against": "oracle/loco_oracle.cpp (CPU restatement) on 8 sampled components of the device's value table"}
    ok = bool(np.array_equal(got_leaf, o_leaf) and np.array_equal(got_status, o_status) and rel <= 1e-12)
    if w.get("ring"):
        wts = (1 + torch.arange(D, device="cuda") % 7).double()
        ref_chk = (kept * wts).sum(dim=1)
        mag = (kept.abs() * wts).sum(dim=1)
        d = float(((out_t[pick] - ref_chk).abs() / mag).max())
        res["timed_checksum_rel_diff"] = d
        ok = ok and d <= 1e-12
    else:
        d = float((out_t[pick][:, :4] - kept[:, :4]).abs().max())   # the timed batch's own rows
        res["timed_rows_abs_diff"] = d
        ok = ok and d == 0.0
    res["ok"] = ok
    return res


def exchange_stats(ex, m, q, Q, leaf, status, stream, torch, world, ms_per_step):
    """the exchange's transfers timed on their own (CUDA events on the side stream around each peer copy) during a few extra steps"""
    stats = {"transport": {"ipc": "peer-mapped buffers (CUDA IPC) + cudaMemcpyAsync on a side stream: copy engines over NVLink",
                           "p2p": "NCCL point-to-point (batch_isend_irecv), same skewed schedule"}[ex.transport],
             "bytes_per_step_per_rank_sent": int(Q * 8 * (world - 1) / world), "shard_bytes": int(Q * 8),
             "schedule": "rank r ships batch b at step b + r to rank b % world: every step is a permutation"}
    if getattr(ex, "_ipc_error", None):
        stats["ipc_error"] = ex._ipc_error
    if ex.transport == "ipc":
        ex.time_copies = True
        ex.copy_events.clear()
        for i in range(2 * world):
            out = ex.out_buffer(i)
            m.eval_device(q.data_ptr(), Q, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), stream)
            ex.submit(i)
        ex.drain()
        torch.cuda.synchronize()
        ex.time_copies = False
        ts = [(a.elapsed_time(b), n) for a, b, n in ex.copy_events]
        if ts:
            gbs = sorted(n / (t * 1e-3) / 1e9 for t, n in ts if t > 0)
            stats.update({"copies_timed": len(ts), "copy_ms_median": float(np.median([t for t, _ in ts])),
                          "copy_gbs_median": float(np.median(gbs)), "copy_gbs_min": gbs[0], "nvlink_peer_gbs_measured_reference": NVLINK_PEER_GBS,
                          "link_busy_frac_of_step": float(np.median([t for t, _ in ts])) / ms_per_step})
    return stats


def measure_e2e(name, w, m, q, Q, torch, dist, world, n_e2e):
    """the metric through the host-buffer C-ABI call: page-locked buffers (headline), the same bytes as plain copies (the box's
    ceiling), and pageable buffers (what a C++ caller of the facade passes; staged inside the library)"""
    N, D = w["N"], w["D"]
    ring = w.get("ring")
    n_e2e = max(2, n_e2e)
    if ring:
        Qe, out_cols = Q, 1          # the step's result is the per-query checksum
    else:
        Qe = Q if D == 1 else min(Q, max(1, (4 << 30) // (D * 8)))
        out_cols = D
    hq = torch.empty((Qe, N), dtype=torch.float64).pin_memory()
    hq.copy_(q[:Qe].cpu())
    ho = torch.empty((Qe, out_cols), dtype=torch.float64).pin_memory()
    hl = torch.empty(Qe, dtype=torch.int32).pin_memory()
    hs = torch.empty(Qe, dtype=torch.uint8).pin_memory()

    def call(qp, op, lp, sp):
        if ring:
            m.eval_ring_raw(qp, Qe, ring, op, lp, sp)
        else:
            m.eval_raw(qp, Qe, op, lp, sp)

    def timed(fn, n):
        fn()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        dt = (time.perf_counter() - t0) / n
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())
    dt = timed(lambda: call(hq.data_ptr(), ho.data_ptr(), hl.data_ptr(), hs.data_ptr()), n_e2e)
    h2d, d2h = Qe * N * 8, Qe * (out_cols * 8 + 5)
    e2e = {"value": world * Qe / dt, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "queries_per_step_per_gpu": Qe, "ms_per_step": dt * 1e3, "host_buffers": "page-locked (torch pin_memory)",
           "result": "per-query checksums of the streamed mesh results" if ring else "values + leaf + status"}
    ref_out = ho[:min(Qe, 32_000_000), 0].clone() if (not ring and D == 1) else None   # (the ceiling test below overwrites ho)
    # ---- the box's copy ceiling for exactly these bytes: H2D and D2H streams running concurrently, nothing else
    dq = torch.empty((Qe, N), dtype=torch.float64, device="cuda")
    do = torch.empty((Qe, out_cols), dtype=torch.float64, device="cuda")
    dl = torch.empty(Qe, dtype=torch.int32, device="cuda")
    ds = torch.empty(Qe, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def copies():
        with torch.cuda.stream(s_in):
            dq.copy_(hq, non_blocking=True)
        with torch.cuda.stream(s_out):
            ho.copy_(do, non_blocking=True)
            hl.copy_(dl, non_blocking=True)
            hs.copy_(ds, non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()
    dtc = timed(copies, n_e2e)
    e2e["copy_ceiling"] = {"value": world * Qe / dtc, "unit": "queries/s", "ms_per_step": dtc * 1e3,
                           "h2d_gbs_per_gpu": h2d / dtc / 1e9, "d2h_gbs_per_gpu": d2h / dtc / 1e9,
                           "what": "the same H2D + D2H bytes as plain page-locked copies on two streams, no kernels (what this box's PCIe / host memory allows at this N)"}
    e2e["frac_of_copy_ceiling"] = e2e["value"] / e2e["copy_ceiling"]["value"]
    del dq, do, dl, ds
    # ---- pageable host buffers (numpy / std::vector): staged through page-locked slots inside loco_eval_batch
    if not ring and D == 1:
        Qp = min(Qe, 32_000_000)
        nq = np.ascontiguousarray(hq[:Qp].numpy().copy())
        no, nl, ns = np.empty(Qp), np.empty(Qp, dtype=np.int32), np.empty(Qp, dtype=np.uint8)
        dtp = timed(lambda: m.eval_raw(nq.ctypes.data, Qp, no.ctypes.data, nl.ctypes.data, ns.ctypes.data), 2)
        same = bool(np.array_equal(no, ref_out.numpy(), equal_nan=True))
        e2e["pageable"] = {"value": world * Qp / dtp, "unit": "queries/s", "queries_per_step_per_gpu": Qp, "ms_per_step": dtp * 1e3,
                           "results_equal_pinned_path": same,
                           "what": "the same call with pageable buffers (what LCSampledFunction::evalShapeInfoBatch of the C++ facade passes): staged through "
                                   "page-locked slots by host threads inside the library"}
    return e2e


if __name__ == "__main__":
    main()
