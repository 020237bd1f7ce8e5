#!/usr/bin/env python
"""bench.py -- throughput of the batched LCSampledFunction evaluation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c4b|c5|a1] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic queries (per GPU; weak scaling: the batch
per GPU is fixed as N grows, the model is replicated, queries are independent).  Rank 0 prints ONE JSON line.

Workloads = BASELINE.json configs.  Default c2 (configs[1]): 3-D cubic scalar function,
bootstrap level 5 -> 32,768 leaves / 42,875 samples / K=64, 100M queries per step.

  value      queries/s with the batch already resident in HBM, timed with CUDA events on the launching stream
  e2e        the same metric through the host-buffer C-ABI call loco_eval_batch (pinned host memory,
             H2D + kernels + D2H inside the timed region)
  roofline   dominant kernel: algorithmic bytes (SURVEY.md 8d formula) / its CUDA-event duration, against
             the measured HBM peak in MEASURED_PEAKS.json
  cpu_baseline  the reference's own C++ path (oracle/_ref/liblocoref.so, fork-parallel over the host cores)
             or, where that library is absent, the oracle port, on a bounded sample of the same workload
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
    "c3": dict(N=4, basis=0, level=3, D=30_000, corners=True, Q=262_144,
               desc="4-D linear mesh-valued (10k vertices, D=30,000), bootstrap 3 (4,096 leaves, K=16); 10M queries streamed in 262,144-query batches (63 GB of results per batch)"),
    "c4": dict(N=6, basis=1, level=1, D=150_000, corners=False, Q=16_384,
               desc="6-D cubic mesh-valued (50k vertices, D=150,000), bootstrap 1 (64 leaves, K=4,096); 10M queries streamed in 16,384-query batches (19.7 GB of results per batch)"),
    "c4b": dict(N=6, basis=1, level=2, D=150_000, corners=False, Q=262_144, ring=16_384,
                desc="6-D cubic mesh-valued (50k vertices, D=150,000), bootstrap 2 (4,096 leaves at depth 12, 15,625 value rows = 18.75 GB, K=4,096); "
                     "262,144 queries per step sorted once and streamed through a 16,384-query ring (19.7 GB) with a per-query checksum"),
    # not a BASELINE config: the reference's own adaptively refined tree (its builder's output, committed as a fixture)
    "a1": dict(N=2, basis=1, level=-1, D=1, corners=True, Q=100_000_000, proto="tests/golden/cub2d_testdebug.pb.gz",
               desc="2-D cubic scalar on the reference's TestDebug tree (TestDebug.cpp:153-157: maxDepth 10, threshold 0.1, bootstrap 2 -> "
                    "724 leaves at levels 4-10, 1,063 samples, 54,419 basis functions), generic term-list kernels, 100M queries"),
    "c5": dict(N=8, basis=0, level=1, D=1, corners=True, Q=64_000_000,
               desc="8-D linear scalar error check, bootstrap 1 (256 leaves, K=256); 1B candidate points per pass in 64M-point batches"),
}


def algorithmic_bytes_per_query(N, depth, T, K, D):
    """a cache-less single pass of the reference's data flow, counted once."""
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


def cpu_reference_arm(w, sample_q, steps, warmup, name):
    """The reference's own CPU implementation on the host cores (oracle/_ref when present, else the oracle port).
    Mesh-valued workloads: the reference has no mesh value type (LCFunctionValue.cpp:20-31), so a query costs the
    reference's own locate + influencing-set assembly + weights, then `val += w * other` over the D components
    (SURVEY.md 8d); to bound host memory the table has min(D, 3000) columns and the rate is scaled by D_used / D
    (the component loop is linear in D and dominates)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from locointerpolate_b200 import capi
    from oracle import oracleapi, refapi
    import protoschema
    import tempfile
    cores = len(os.sched_getaffinity(0))
    N, D = w["N"], w["D"]
    m, _, _ = build_model(capi, {**w, "D": 1}, capi.HOST_ONLY)  # the same tree with scalar values
    tmp = tempfile.mktemp(suffix=".pb")
    m.save_proto(tmp)
    g = protoschema.read_binary(tmp)
    os.unlink(tmp)
    rng = np.random.default_rng(99)
    q = rng.random((sample_q, N))
    scale, note = 1.0, ""
    if refapi.available():
        fn = refapi.RefFn.from_graph(g)
        kind, used = "reference", cores
        if D == 1:
            run = lambda: fn.eval_fork(q, cores)
        else:
            Du = min(D, 3000)
            table = rng.standard_normal((fn.S, Du))
            scale = Du / D
            note = f"; mesh values: reference weights + component loop over {Du} of {D} columns, rate scaled by {Du}/{D}"
            run = lambda: fn.eval_mesh_fork(q, table, cores)
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
    how = f"fork-parallel over {used} host cores" if kind == "reference" else "single thread"
    return dict(value=sample_q / dt * scale, unit="queries/s", cores=used, kind=kind,
                sample=f"{sample_q} uniform queries of workload {name} per step, {steps} steps, {how}{note}"), dt


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
    ap.add_argument("--ecell-poly", type=int, default=0, help="1: polynomial form of the evaluation cells (adaptive scalar trees; option ecell_poly)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--gather", default="rotate", choices=["rotate", "all"],
                    help="multi-GPU exchange step: results of batch i gathered on rank i mod N (default) or all_gather on every rank")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads (c3, c4) reported under other_workloads")
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 direct, 2 leaf-sorted TMA tiles, 4 L2-resident sorted chunks")
    ap.add_argument("--overlap", type=int, default=0, help="streams the chunks of the sorted path are dealt over (0 = library default)")
    ap.add_argument("--chunk", type=int, default=0, help="queries per L2-resident chunk of the sorted path (0 = auto)")
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
    config = {"workload": f"{args.workload}: {w['desc']}", "n_params": w["N"], "basis": "cubic" if w["basis"] else "linear",
              "value_dim": w["D"], "queries_per_step_per_gpu": w["Q"],
              "parallelism": f"query-sharded x{world}, model replicated" + (f", results of batch i gathered on rank i mod {world}" if world > 1 and args.gather == "rotate" and w["D"] == 1 else ""),
              "l2": "inputs+outputs per step exceed the 126 MB L2" if w["Q"] * (w["N"] + w["D"]) * 8 > 126e6 else "L2 flushed between steps"}

    # ------------------------------------------------------------------ reference arm (CPU) ----------
    if args.impl == "reference":
        if rank != 0:
            return
        cpu_default = {"c1": 2_000_000, "c2": 200_000, "c3": 500_000, "c4": 2048, "c4b": 512, "c5": 100_000, "a1": 600_000}[args.workload]
        base, dt = cpu_reference_arm(w, args.cpu_sample or cpu_default, max(args.steps, 1), args.warmup, args.workload)
        print(json.dumps({"impl": "reference", "metric": metric, "value": base["value"], "unit": "queries/s", "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": base,
                          "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ our arm (GPU) -----------------
    import torch
    import torch.distributed as dist
    from locointerpolate_b200 import build as product_build, capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        opts = None
        try:  # the result gather runs beside the next batch's kernels: give NCCL's stream priority over them
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:
            pass
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    if local == 0:
        product_build.build()
    if world > 1:
        dist.barrier()

    line = measure(args.workload, w, args, rank, world, local, config, full=True)
    # the other single-GPU configurations of BASELINE.json, measured the same way with fewer steps (secondary numbers;
    # the headline stays the workload named in `config`)
    if not args.no_extra and args.workload == "c2":
        extra = {}
        for name in ("c3", "c4"):
            try:
                torch.cuda.empty_cache()
                wx = dict(WORKLOADS[name])
                r = measure(name, wx, args, rank, world, local, None, full=False)
                extra[name] = r
            except Exception as e:  # a secondary number must never take the headline down
                extra[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
        line["other_workloads"] = extra
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


KERNEL_ROLE = {  # which part of SURVEY.md 8(d)'s per-query byte formula a kernel covers
    "sorted_count": "locate", "locate_count": "locate", "sorted_scatter": "sort", "scatter": "sort", "sorted_scan": "sort", "scan_leaves": "sort",
    "bucket_scatter": "locate+sort", "bucket_eval": "eval", "sorted_eval": "eval", "eval_tensor_tiles": "eval", "eval_tiles": "eval", "mesh_tensor_mma": "eval", "mesh_tensor_tiles": "eval",
    "mesh_gather": "eval", "weights": "eval", "eval_scalar_direct": "all", "eval_tensor_direct": "all", "jitter_points": "generate",
    "error_reduce": "reduce", "mesh_nan_rows": "other", "checksum": "other",
}


def role_bytes_per_query(role, N, depth, T, K, D):
    locate = 8 * N + 16 * depth + 5
    evalb = T * (16 * N + 8 + 4) + K * D * 8 + 8 * D
    return {"locate": locate, "sort": 2 * (8 * N + 8), "locate+sort": locate + 2 * (8 * N + 8), "eval": evalb + 8 * N, "all": locate + evalb, "generate": 8 * N + 12, "reduce": 8 * D * 2 + 5}.get(role, 0)


def measure(name, w, args, rank, world, local, config, full):
    """one workload on this rank's GPU: timed steps (max over ranks), per-kernel CUDA-event times, roofline, e2e (full only)"""
    import torch
    import torch.distributed as dist
    from locointerpolate_b200 import capi
    steps = args.steps if full else max(2, min(args.steps, 3))
    warmup = args.warmup if full else 3
    m, amp, phase = build_model(capi, w, local)
    m.set_option("path", args.path)
    m.set_option("chunk", args.chunk)
    if args.ecell_poly:
        m.set_option("ecell_poly", args.ecell_poly)
    if args.overlap:
        m.set_option("overlap", args.overlap)
    info = m.info
    N, D, Q = w["N"], w["D"], w["Q"]
    gen = torch.Generator("cuda").manual_seed(1234 + rank)
    q = torch.rand((Q, N), dtype=torch.float64, device="cuda", generator=gen)
    stream = torch.cuda.current_stream().cuda_stream
    leaf = torch.empty(Q, dtype=torch.int32, device="cuda")
    status = torch.empty(Q, dtype=torch.uint8, device="cuda")
    flush = None
    if Q * (N + D) * 8 <= 126e6:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    if name == "c5":
        err = torch.empty(info["n_leaves"], dtype=torch.float64, device="cuda")
        split = torch.empty(info["n_leaves"], dtype=torch.uint8, device="cuda")
        d_amp, d_phase = torch.from_numpy(amp).cuda(), torch.from_numpy(phase).cuda()
        ppl = Q // info["n_leaves"]
        Q = ppl * info["n_leaves"]

        def step(i):
            m.error_check_generated_device(ppl, 1000 + i, d_amp.data_ptr(), d_phase.data_ptr(), 10, 0.05, err.data_ptr(), split.data_ptr(), stream)
    elif D == 1:
        NB = 3 if world > 1 else 2  # result buffers in rotation (with several GPUs: gathers of two earlier batches may be in flight)
        outs = [torch.empty(Q, dtype=torch.float64, device="cuda") for _ in range(NB)]
        # with several GPUs the step is issued in sub-batches so that the gather of a finished part overlaps the rest
        n_sub = 4 if world > 1 and Q % 4 == 0 and Q >= (1 << 22) else 1
        Qs = Q // n_sub

        def step(i, gather=None):
            for j in range(n_sub):
                m.eval_device(q.data_ptr() + j * Qs * N * 8, Qs, outs[i % NB].data_ptr() + j * Qs * 8, leaf.data_ptr() + j * Qs * 4,
                              status.data_ptr() + j * Qs, stream)
                if gather is not None:
                    gather(i, j)
    elif w.get("ring"):
        # mesh-valued, streamed: results pass through a ring in leaf order and are reduced to one checksum per query
        ring = torch.empty((w["ring"], D), dtype=torch.float64, device="cuda")
        chk = torch.empty(Q, dtype=torch.float64, device="cuda")

        def step(i):
            m.eval_ring_device(q.data_ptr(), Q, ring.data_ptr(), w["ring"], chk.data_ptr(), leaf.data_ptr(), status.data_ptr(), stream)
    else:
        # mesh-valued: the whole batch of results stays in HBM ([Q x D] doubles; sized for the 180 GB part)
        assert D % 2 == 0, "bench assumes an even value dimension"
        mesh_out = torch.empty((Q, D), dtype=torch.float64, device="cuda")

        def step(i):
            m.eval_device(q.data_ptr(), Q, mesh_out.data_ptr(), leaf.data_ptr(), status.data_ptr(), stream)

    gathered, handles = None, []
    if world > 1 and D == 1 and name != "c5":
        gathered = [torch.empty((n_sub, world, Qs), dtype=torch.float64, device="cuda") for _ in range(NB)]  # [sub-batch][rank][query]

    # The path's one exchange step: the per-shard results of batch i are gathered on rank (i mod world) -- the consumer of
    # that batch -- asynchronously on NCCL's stream, so that it overlaps the next batches.  (Gathering EVERY batch on ONE
    # GPU, or on all of them, cannot scale: 8 GPUs produce 8 B x 27 G results/s each = 1.7 TB/s, twice what one GPU's
    # NVLink ingress takes; with rotating consumers each GPU ingests its own production rate.)  --gather all: all_gather.
    per_step = {}
    # one communicator per consumer rank: NCCL runs the operations of ONE communicator in order, so gathers towards
    # different consumers could not overlap on a single one (a gather lasts as long as its root's ingress takes)
    root_groups = None
    if gathered is not None and args.gather == "rotate":
        root_groups = []
        for _ in range(world):
            try:
                root_groups.append(dist.new_group(list(range(world)), pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)))
            except Exception:
                root_groups.append(dist.new_group(list(range(world))))

    def gather_part(i, j):
        src = outs[i % NB][j * Qs:(j + 1) * Qs]
        if args.gather == "all":
            h = dist.all_gather_into_tensor(gathered[i % NB][j], src, async_op=True)
        else:
            root = i % world
            h = dist.gather(src, list(gathered[i % NB][j].unbind(0)) if rank == root else None, dst=root, group=root_groups[root], async_op=True)
        handles.append(h)
        per_step.setdefault(i, []).append(h)

    def run(i):
        if gathered is not None:
            for h in per_step.pop(i - NB, []):  # outs[i % NB] / gathered[i % NB] are reused: batch i-NB must have left
                h.wait()
            step(i, gather_part)
        else:
            step(i)
        if gathered is None and world > 1 and name == "c5":  # error maxima: all_reduce(MAX) of the bit patterns + split flags (shard.py)
            from locointerpolate_b200 import shard
            shard.reduce_error_maxima(err, split)

    sampler = ClockSampler(local)
    if rank == 0 and full:
        sampler.start()  # from the warm-up on: short timed regions would otherwise fall between two 100 ms samples
    for i in range(warmup):
        if flush is not None:
            flush.zero_()
        run(i)
    for h in handles:
        h.wait()
    handles.clear()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = m.kernel_launches
    t_wall = time.perf_counter()
    if flush is None:
        # inputs + outputs of a step exceed the L2: time the K steps back to back
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            run(i)
        for h in handles:
            h.wait()
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
            run(i)
            for h in handles:
                h.wait()
            handles.clear()
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 and full else None
    launches = m.kernel_launches - launches0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / steps
    value = world * Q / (ms_per_step * 1e-3)
    bad = int((status != 0).sum().item()) if name != "c5" else 0

    # ---- every kernel of the step, CUDA events around each launch on the launching stream (no flush, no gather)
    #      (measured with the chunks of the sorted path on ONE stream: with the default two-stream overlap the events of
    #      a kernel would also cover the kernels running beside it)
    m.set_option("overlap", 1)
    m.profile_begin()
    for i in range(steps):
        step(i)
    prof = m.profile_end()
    m.set_option("overlap", args.overlap or 3)
    tot_ms = sum(v["total_ms"] for v in prof.values()) or 1.0
    kernels = {k: {"launches_per_step": v["launches"] / steps, "ms_per_launch": v["total_ms"] / v["launches"], "share": v["total_ms"] / tot_ms}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["total_ms"])}
    dom = next(iter(kernels))
    dom_ms = kernels[dom]["ms_per_launch"]
    q_per_launch = Q / kernels[dom]["launches_per_step"]
    T = info["n_terms"] / info["n_leaves"]
    K = info["n_support_entries"] / info["n_leaves"]
    Kf = info["n_fold_entries"] / info["n_leaves"] if info.get("n_fold_entries") else K
    role = KERNEL_ROLE.get(dom, "all")
    algo_q = role_bytes_per_query(role, N, info["depth"], T, K, D)
    algo = algo_q * q_per_launch
    path_q = algorithmic_bytes_per_query(N, info["depth"], T, K, D)
    compulsory = Q * (8 * N + 8 * D + 5) + info["device_bytes"]
    peak, peak_src = measured_peak()
    step_kernel_ms = tot_ms / steps
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get(name, {})
        if tj.get("kernel") == dom:
            traffic = tj["dram_bytes_per_query"] * q_per_launch
    roofline = {"bound": "hbm", "kernel": dom, "kernel_role": role, "achieved": algo / (dom_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": algo / (dom_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms_per_launch": dom_ms, "queries_per_launch": q_per_launch, "algorithmic_bytes_per_query": algo_q,
                "whole_path": {"algorithmic_bytes_per_query": path_q, "kernels_ms_per_step": step_kernel_ms,
                               "achieved_gbs": path_q * Q / (step_kernel_ms * 1e-3) / 1e9},
                "compulsory": {"bytes_per_step": compulsory, "achieved_gbs": compulsory / (step_kernel_ms * 1e-3) / 1e9,
                               "frac": compulsory / (step_kernel_ms * 1e-3) / 1e9 / peak},
                "fp64": {"rows_per_leaf": K, "rows_per_leaf_folded": Kf, "flops_per_query": 2.0 * Kf * D,
                         "achieved_tflops": 2.0 * Kf * D * Q / (step_kernel_ms * 1e-3) / 1e12, "peak_tflops": 36.4,
                         "peak_source": "measured DFMA loop, tools/micro/dmma_rate.cu",
                         "note": "flops counted on the folded rows actually multiplied (samples sharing one value object are summed first)"}
                if D > 1 else None,
                "note": "algorithmic bytes = SURVEY.md 8(d): the reference's per-query data flow without caches, restricted to the part the dominant "
                        "kernel covers (kernel_role); leaf data is served from L1/L2/shared memory, so 'achieved' may exceed the HBM peak -- "
                        "'compulsory' is what must cross HBM (queries in, results out, model once)"}

    result = {"value": value, "unit": "queries/s", "ms_per_step": ms_per_step, "steps": steps, "queries_per_step_per_gpu": Q,
              "gpu_launches": int(launches), "kernels": kernels, "roofline": roofline, "bad_status": bad,
              "model": {k: info[k] for k in ("n_leaves", "n_samples", "n_terms", "max_support", "depth", "device_bytes")}}
    if not full:
        result["workload"] = f"{name}: {w['desc']}"
        del q, leaf, status
        m.free()
        return result

    # ---- end to end through the host-buffer C-ABI call (pinned host memory; H2D + kernels + D2H inside)
    e2e = None
    if name != "c5" and not args.no_e2e:
        Qe = Q if D == 1 else min(Q, max(1, (4 << 30) // (D * 8)))
        hq = torch.empty((Qe, N), dtype=torch.float64).pin_memory()
        hq.copy_(q[:Qe].cpu())
        ho = torch.empty((Qe, D), dtype=torch.float64).pin_memory()
        hl = torch.empty(Qe, dtype=torch.int32).pin_memory()
        hs = torch.empty(Qe, dtype=torch.uint8).pin_memory()
        n_e2e = max(2, min(steps, 3))
        m.eval_raw(hq.data_ptr(), Qe, ho.data_ptr(), hl.data_ptr(), hs.data_ptr())
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            m.eval_raw(hq.data_ptr(), Qe, ho.data_ptr(), hl.data_ptr(), hs.data_ptr())
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * Qe / dt, "unit": "queries/s", "h2d_bytes_per_step": Qe * N * 8, "d2h_bytes_per_step": Qe * (D * 8 + 5),
               "queries_per_step_per_gpu": Qe, "ms_per_step": dt * 1e3}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_default = {"c1": 2_000_000, "c2": 200_000, "c3": 500_000, "c4": 2048, "c4b": 512, "c5": 100_000, "a1": 600_000}[name]
        try:
            cpu, _ = cpu_reference_arm(w, args.cpu_sample or cpu_default, 1, 0, name)
        except SystemExit as e:
            cpu = {"value": None, "unit": "queries/s", "cores": 0, "kind": "port", "sample": str(e)}

    line = {"metric": "interpolated_queries_per_sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "kernels": kernels, "wall_s": wall, "bad_status": bad, "model": result["model"]}
    del q, leaf, status
    m.free()
    return line


if __name__ == "__main__":
    main()
