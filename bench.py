#!/usr/bin/env python
"""bench.py -- headline benchmark of the hybrid SpMM hot path (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  (N > 1: launched by torchrun, one rank per GPU)

A "step" is one pass of the hot path over one batch of synthetic input: Y = A * X on the
Reddit-shape power-law graph (232 965 vertices, ~114.6 M stored entries, dim 256, FP32).
At N > 1 the adjacency is row-window partitioned (nnz-balanced) and every step first
all-gathers the row-sharded X over NCCL, then runs the local SpMM -- the per-layer exchange of
a row-partitioned GCN (strong scaling: the total work is fixed).

Prints ONE JSON line (rank 0).  Keys follow the driver's contract; see DESIGN.md "Measurement".
Nothing here reads /root/reference.  The CPU oracle / torch.sparse.mm is executed only in the
`cpu_baseline` leg and under `--impl reference`.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "hc-spmm_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="reddit", choices=["reddit", "products", "proteins", "envelope"])
    ap.add_argument("--ref-kernel", action="store_true",
                    help="with --impl reference --shape envelope: time the recompiled reference CUDA extension "
                         "(oracle/_ref/HCSPMM_ref.so) instead of the CPU baseline")
    ap.add_argument("--dim", type=int, default=0, help="feature width (0 = the shape's own)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink N and nnz (smoke runs only)")
    ap.add_argument("--classifier", default="shipped",
                    help="core selector: shipped (reference, all CUDA-core) | intended | b200 | all_tc")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "tf32x2", "fp32", "bf16"])
    ap.add_argument("--slab", type=int, default=-1, help="feature-slab width (-1 = library default)")
    ap.add_argument("--long-row", type=int, default=-1)
    ap.add_argument("--vec8", type=int, default=-1, help="256-bit gathers: 1 on, 0 off (-1 = library default)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "gather", "slabs", "halo", "peer"],
                    help="N > 1: all-gather of X, all-gather pipelined in feature slabs, halo rows only (NCCL all-to-all), "
                         "peer = halo rows pulled over NVLink peer memory by our own kernel (auto picks this)")
    ap.add_argument("--exchange-passes", type=int, default=1, choices=[1, 2],
                    help="peer exchange: 2 = shard cut by source into two accumulating SpMM passes, second half of the pull overlapped")
    ap.add_argument("--overlap-ctas", type=int, default=64)
    ap.add_argument("--exchange-slabs", type=int, default=1,
                    help="N > 1: all-gather X in this many feature slabs, slab k+1 in flight while the SpMM of slab k runs "
                         "(1 = one all-gather, then one SpMM)")
    ap.add_argument("--dense", action="store_true", help="tcgen05 dense super-window plan (with --classifier b200|all_tc)")
    ap.add_argument("--tune", action="append", default=[], help="library tuning knob key=value (repeatable)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline budget per timed run")
    return ap.parse_args()


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l.split(", ") for t, l in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or \
               [l.split(", ") for _, l in self.lines]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_spmm_baseline(rp_cpu, ci_cpu, x_cpu, seconds: float):
    """torch.sparse.mm (CSR, FP32) on the host cores -- the paper's 'PyTorch CPU SpMM' baseline and the
    north star's stated reference -- on a BOUNDED sample: the first n_s rows of the same graph against
    the full X, n_s sized so one run takes about `seconds`."""
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n, dim = rp_cpu.numel() - 1, x_cpu.shape[1]

    def run(n_s):
        e = int(rp_cpu[n_s])
        a = torch.sparse_csr_tensor(rp_cpu[: n_s + 1].to(torch.int64), ci_cpu[:e].to(torch.int64),
                                    torch.ones(e, dtype=torch.float32), size=(n_s, x_cpu.shape[0]))
        t = time.perf_counter()
        y = torch.sparse.mm(a, x_cpu)
        dt = time.perf_counter() - t
        return dt, e, y

    probe_rows = max(16, n // 64)
    run(probe_rows)                                  # warm-up (thread pool, page-in)
    dt, e, _ = run(probe_rows)
    rate = max(e, 1) / max(dt, 1e-6)                 # entries / s
    want_e = min(int(rp_cpu[-1]), int(rate * seconds))
    n_s = int(torch.searchsorted(rp_cpu.to(torch.int64), torch.tensor(want_e)).clamp(16, n))
    times = []
    for _ in range(2):
        dt, e, _ = run(n_s)
        times.append(dt)
    dt = statistics.median(times)
    return {"value": 2.0 * e * dim / dt / 1e9, "unit": "GFLOP/s", "cores": threads, "kind": "port",
            "sample": f"torch.sparse.mm CSR FP32 on CPU, rows [0,{n_s}) of {n} ({e} of {int(rp_cpu[-1])} "
                      f"stored entries) x full X[{x_cpu.shape[0]},{dim}], median of 2 runs, {dt:.2f} s each",
            "seconds": dt, "entries": e, "rows": n_s}


def reference_kernel_arm(args):
    """The UNMODIFIED reference extension recompiled for sm_100 (oracle/_ref/HCSPMM_ref.so), timed on the
    only kind of input it can run: <= 62 edges per 16-row window, dim == 32 (forward_fixed32)."""
    import importlib.machinery
    import importlib.util
    from hcspmm import graphs
    so = os.path.join(ROOT, "oracle", "_ref", "HCSPMM_ref.so")
    if not os.path.exists(so):
        return {"impl": "reference", "unavailable": "oracle/_ref/HCSPMM_ref.so not built (reference absent at build time)"}
    loader = importlib.machinery.ExtensionFileLoader("HCSPMM_ref", so)
    ref = importlib.util.module_from_spec(importlib.util.spec_from_loader("HCSPMM_ref", loader))
    loader.exec_module(ref)
    assert args.shape == "envelope", "the reference kernels corrupt shared memory outside --shape envelope"
    dev = torch.device("cuda", 0)
    rp, ci, info = graphs.named("envelope", device=dev)
    n, nnz, dim = info["n"], info["nnz"], 32
    pre = ref.preprocess(ci, rp, n, nnz, n // 16)
    x = torch.randn(n, dim, device=dev, generator=torch.Generator(device=dev).manual_seed(1234))
    for _ in range(args.warmup):
        ref.forward_fixed32(x, rp, ci, *pre)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        y = ref.forward_fixed32(x, rp, ci, *pre)[0]
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    val = 2.0 * nnz * dim / (ms * 1e-3) / 1e9
    return {"impl": "reference", "metric": "spmm_gflops", "value": val, "unit": "GFLOP/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "envelope: ring + perfect matching, N=%d nnz=%d dim=32, reference "
                                   "forward_fixed32 kernel recompiled for sm_100" % (n, nnz)},
            "gpu_launches": args.steps}


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from hcspmm import graphs, partition

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference" and args.ref_kernel:
        if rank == 0:
            print(json.dumps(reference_kernel_arm(args)))
        return 0
    if args.impl == "reference":
        if rank != 0:
            return 0
        shape = graphs.SHAPES[args.shape]
        dim = args.dim or shape["dim"]
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        rp, ci, info = graphs.named(args.shape, device=dev, scale=args.scale)
        rp_c, ci_c = rp.cpu(), ci.cpu()
        x = torch.randn(info["n"], dim, generator=torch.Generator().manual_seed(0))
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        budget = min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup))
        base = cpu_spmm_baseline(rp_c, ci_c, x, budget)
        n_s, e = base["rows"], base["entries"]
        a = torch.sparse_csr_tensor(rp_c[: n_s + 1].to(torch.int64), ci_c[:e].to(torch.int64),
                                    torch.ones(e, dtype=torch.float32), size=(n_s, info["n"]))
        for _ in range(args.warmup):
            torch.sparse.mm(a, x)
        t = time.perf_counter()
        for _ in range(args.steps):
            torch.sparse.mm(a, x)
        dt = (time.perf_counter() - t) / max(1, args.steps)
        val = 2.0 * e * dim / dt / 1e9
        line = {"impl": "reference", "metric": "spmm_gflops", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{args.shape}-shape SpMM N={info['n']} nnz={info['nnz']} dim={dim}",
                           "note": "the reference's SpMM is a CUDA kernel that overflows its shared-memory "
                                   "tables above 62 edges / 16-row window (hybrid_all_kernel.cu:26,964-967), so it "
                                   "cannot run this shape; this arm times the north star's CPU reference, "
                                   "torch.sparse.mm, on all host threads"},
                "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": threads, "kind": "port",
                                 "sample": base["sample"]},
                "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    import HCSPMM
    from hcspmm import capi

    shape = graphs.SHAPES[args.shape]
    dim = args.dim or shape["dim"]
    t_gen = time.perf_counter()
    rp, ci, info = graphs.named(args.shape, device=dev, scale=args.scale)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    n, nnz = info["n"], info["nnz"]

    if args.slab >= 0:
        HCSPMM.set_tuning("slab", args.slab)
    if args.long_row >= 0:
        HCSPMM.set_tuning("long_row", args.long_row)
    if args.vec8 >= 0:
        HCSPMM.set_tuning("vec8", args.vec8)
    for kv in args.tune:
        k, v = kv.split("=")
        HCSPMM.set_tuning(k, int(v))
    HCSPMM.set_dense(bool(args.dense))
    HCSPMM.set_classifier(args.classifier)
    HCSPMM.set_precision(args.precision)

    # row-window partition (nnz-balanced) + the exchange plan of hcspmm.dist; world == 1: the whole graph
    from hcspmm import dist as hd
    sg = hd.ShardedGraph(rp, ci, schedule=args.exchange, n_slabs=max(1, args.exchange_slabs), n_passes=args.exchange_passes)
    sg.overlap_ctas = args.overlap_ctas
    cuts, r0, r1 = sg.cuts, sg.r0, sg.r1
    rp_l, ci_run, pre = sg.rowptr, sg.colidx, sg.pre
    n_l, nnz_l, x_rows_run = sg.n_local, sg.nnz_local, sg.x_rows

    # preprocessing (reported separately, like the paper: "x one SpMM")
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    pre = HCSPMM.preprocess(ci_run, rp_l, n_l, nnz_l, (n_l + 15) // 16)
    ev1.record()
    torch.cuda.synchronize()
    prep_ms = ev0.elapsed_time(ev1)
    sg.pre = pre
    tc_windows = int((pre[3] != 0).sum())
    dense_groups = int(pre[4][1]) if (pre[4].device.type == "cpu" and pre[4].numel() >= 4) else 0

    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    x_full = torch.randn(n, dim, device=dev, generator=g)       # same on every rank (same seed)
    x_loc = x_full[r0:r1].contiguous()
    n_slabs = sg.n_slabs if (world > 1 and sg.schedule in ("slabs", "halo", "peer")) else 1

    def step():
        # N > 1: exchange of the row shards of X (all-gather or halo all-to-all), then the local SpMM
        return sg.aggregate(x_loc if world > 1 else x_full)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        y = step()
    barrier()

    # timed region: exactly K steps, CUDA events on the launching (current) stream
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t0 = time.time()
    ev0.record()
    for i in range(args.steps):
        kern_ev[i][0].record()
        y = step()
        kern_ev[i][1].record()
    ev1.record()
    barrier()
    t1 = time.time()
    total_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in kern_ev]
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t)
    ms_per_step = total_ms / args.steps
    flops = 2.0 * nnz * dim
    value = flops / (ms_per_step * 1e-3) / 1e9

    phases = None
    if world > 1:
        def tm(fn, k=5):
            fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(k):
                fn()
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b) / k], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        operand = sg.exchange(x_loc)
        phases = {"exchange_only_ms": tm(lambda: sg.exchange(x_loc)),
                  "kernel_only_ms": tm(lambda: HCSPMM.forward(operand, rp_l, ci_run, *pre)),
                  "exchange_bytes_per_rank": int(sg.exchange_rows() * dim * 4),
                  "exchange_rows_vs_allgather": sg.exchange_rows() / max(1, (world - 1) * sg.max_rows)}

    sg.check()          # a peer barrier that timed out would have left stale rows in the operand
    # quick full-size sanity inside the bench (not timed): X = 1 gives the row degrees exactly
    ones = torch.ones(x_rows_run, 8, device=dev)
    deg = HCSPMM.forward(ones, rp_l, ci_run, *pre)[0][:, 0]
    want = (rp_l[1:] - rp_l[:-1]).float()
    assert torch.equal(deg, want) or float((deg - want).abs().max()) <= 1e-3 * float(want.max()), \
        "degree check failed: kernel output is wrong"

    # our kernels per step: the balanced path is merge_path_splits + spmm_balanced + fixup (library rule:
    # mean row >= 8 entries, knob "balance"), else the one hybrid kernel; BF16 adds the X conversion kernel
    bal_knob = HCSPMM.set_tuning("balance", 1)
    HCSPMM.set_tuning("balance", bal_knob)
    balanced = bal_knob >= 2 or (bal_knob == 1 and nnz_l >= 8 * n_l)
    launches_per_step = (3 if balanced else 1) * n_slabs + (1 if args.precision == "bf16" else 0) * n_slabs
    if world > 1 and sg.schedule == "peer":
        launches_per_step += 1 + n_slabs          # hcspmm_peer_barrier + hcspmm_halo_pull per slab
    # roofline of the dominant kernel (the SpMM launch of a step); single-GPU figures
    peak, peak_src = measured_peak_gbs()
    bytes_alg = nnz_l * (4 + 4 * dim) + n_l * (4 * dim + 4)
    bytes_min = 4 * nnz_l + 4 * (n_l + 1) + 4 * (n + n_l) * dim
    # N = 1: the step IS the SpMM; N > 1: this rank's local SpMM timed on its own (phases, max over ranks)
    kern_ms = statistics.mean(step_ms) if world == 1 else (phases or {}).get("kernel_only_ms")
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(f"{args.shape}_dim{dim}_{args.classifier}") if world == 1 else None
    except Exception:
        pass
    roofline = None
    if kern_ms:
        ach = bytes_alg / (kern_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "peak_source": peak_src,
                    "kernel": "spmm_balanced_kernel (+ merge_path_splits_kernel, spmm_balanced_fixup_kernel: the step)"
                              if balanced else "spmm_hybrid_kernel",
                    "kernel_ms": kern_ms, "algorithmic_bytes": bytes_alg,
                    "compulsory_bytes": bytes_min, "compulsory_frac": bytes_min / (kern_ms * 1e-3) / 1e9 / peak,
                    "frac_of_nominal_8TBs": ach / 8000.0,
                    "scope": "whole graph, one GPU" if world == 1 else
                             "rank 0's row shard on one GPU (local SpMM timed alone, max over ranks)"}

    # e2e: the reference-facing module with HOST buffers, H2D of X and D2H of Y inside the timed region
    e2e = None
    if not args.no_e2e:
        rows_in = n if world == 1 else n_l
        xh = torch.empty(rows_in, dim, pin_memory=True)
        xh.copy_(x_full if world == 1 else x_loc)
        yh = torch.empty(n_l, dim, pin_memory=True)

        # Three streams, double-buffered device X / Y: H2D(i+1) | kernel(i) | D2H(i-1) overlap; every
        # step's copies are inside the timed region.
        cur = torch.cuda.current_stream(dev)
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        xd = [torch.empty(rows_in, dim, device=dev) for _ in range(2)]
        outs = [None, None]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_k = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]

        def e2e_run(k):
            for i in range(k):
                b = i & 1
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_k[b])              # kernel i-2 has consumed xd[b]
                    xd[b].copy_(xh, non_blocking=True)
                    ev_in[b].record(s_in)
                cur.wait_event(ev_in[b])
                cur.wait_event(ev_out[b])                 # D2H of step i-2 has drained outs[b]
                if world > 1:
                    outs[b] = sg.aggregate(xd[b])
                else:
                    outs[b] = HCSPMM.forward(xd[b], rp_l, ci_run, *pre)[0]
                ev_k[b].record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_k[b])
                    yh.copy_(outs[b], non_blocking=True)
                    ev_out[b].record(s_out)
            cur.wait_stream(s_out)
            cur.wait_stream(s_in)

        e2e_run(2)
        barrier()
        k2 = max(4, min(args.steps, 10))
        ev0.record()
        e2e_run(k2)
        ev1.record()
        barrier()
        e_ms = ev0.elapsed_time(ev1) / k2
        if world > 1:
            t = torch.tensor([e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t)
        e2e = {"value": flops / (e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": e_ms, "steps": k2,
               "pipeline": "pinned host X -> H2D stream | HCSPMM.forward | D2H stream -> pinned host Y, double buffered",
               "h2d_bytes_per_step": rows_in * dim * 4 * world if world > 1 else rows_in * dim * 4,
               "d2h_bytes_per_step": n * dim * 4}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_spmm_baseline(rp.cpu(), ci.cpu(), x_full.cpu(), args.cpu_seconds)
        for k in ("seconds", "entries", "rows"):
            cpu_base.pop(k, None)

    if rank == 0:
        line = {"metric": "spmm_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{args.shape}-shape power-law graph, single-kernel SpMM Y=A*X "
                                       f"(BASELINE.json configs[1])" if args.shape == "reddit" else f"{args.shape}-shape SpMM",
                           "nodes": n, "stored_entries": nnz, "dim": dim, "classifier": args.classifier,
                           "precision_tc_windows": args.precision, "tc_windows": tc_windows, "dense_groups_tcgen05": dense_groups,
                           "windows": (n_l + 15) // 16, "partition": f"row windows, nnz-balanced, {world} shard(s)",
                           "exchange": ({"gather": "NCCL all_gather_into_tensor of the row shards of X per step",
                                         "slabs": "NCCL all_gather_into_tensor in %d feature slabs pipelined with the SpMM" % n_slabs,
                                         "halo": "halo rows only: pack + NCCL all_to_all_single per step, %d feature slab(s)" % n_slabs,
                                         "peer": "halo rows only, pulled from the owners' memory over NVLink by hcspmm_halo_pull "
                                                 "after hcspmm_peer_barrier, %d feature slab(s), %d source pass(es)"
                                                 % (n_slabs, 2 if sg.passes is not None else 1)}
                                        [sg.schedule]) if world > 1 else "none",
                           "phases": phases,
                           "l2": "inputs larger than L2 (X %.0f MB + CSR %.0f MB vs 126 MB), no flush" %
                                 (n * dim * 4 / 1e6, nnz * 4 / 1e6),
                           "preprocess_ms": prep_ms, "graph_gen_s": t_gen,
                           "generator": "R-MAT(0.57,0.19,0.19,0.05) folded mod N, symmetrised, de-duplicated, "
                                        "ids permuted, seed %d" % shape["seed"]},
                "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "clocks": clocks,
                "gpu_launches": args.steps * launches_per_step, "launches_per_step": launches_per_step,
                "step_ms": {"min": min(step_ms), "median": statistics.median(step_ms), "max": max(step_ms)}}
        print(json.dumps(line))
    if world > 1:
        sg.close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
