#!/usr/bin/env python
"""bench.py -- headline benchmark of the hybrid SpMM hot path (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  (N > 1: launched by torchrun, one rank per GPU)

A "step" is one pass of the hot path over one batch of synthetic input: Y = A * X on the
Reddit-shape power-law graph (232 965 vertices, ~114.6 M stored entries, dim 256, FP32).
At N > 1 the adjacency is row-window partitioned (nnz-balanced) and every step first exchanges
the row-sharded X -- by default only the HALO rows a shard references, pulled over NVLink peer
memory by our own kernels (hcspmm_peer_barrier + hcspmm_halo_pull); `--exchange gather` is the
north star's NCCL all-gather, kept as the comparison -- then runs the local SpMM: the per-layer
exchange of a row-partitioned GCN (strong scaling: the total work is fixed).  Every line carries a
`parity` record: the timed path's result against FP32 torch.sparse.mm on the host and, at N > 1,
against rows [r0, r1) of the single-GPU aggregation of the unpartitioned graph.
`extra.sweep` (N = 1): the other widths / shapes of the metric (dim 64 / 128 / 512, the products
and proteins shapes), each with its own roofline fraction; `extra.products` (N > 1): the
ogbn-products-shape scaling record of SURVEY 8(e).

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
    ap.add_argument("--exchange", default="auto", choices=["auto", "gather", "slabs", "halo", "peer", "push"],
                    help="N > 1: all-gather of X, all-gather pipelined in feature slabs, halo rows only (NCCL all-to-all), "
                         "peer = halo rows pulled over NVLink peer memory by our own kernel (auto picks this)")
    ap.add_argument("--exchange-passes", type=int, default=1, choices=[1, 2],
                    help="peer exchange: 2 = shard cut by source into two accumulating SpMM passes, second half of the pull overlapped")
    ap.add_argument("--overlap-ctas", type=int, default=64)
    ap.add_argument("--stage-copy", action="store_true",
                    help="N > 1: keep the rank's X shard in ordinary memory and copy it into the exchange operand every step "
                         "(default: X lives in the operand's own-rows segment)")
    ap.add_argument("--direct-refs", type=int, default=0,
                    help="peer exchange: remote rows of X referenced at most this many times by a shard are read in place by "
                         "the SpMM instead of being pulled (0 = off, default; -1 = auto)")
    ap.add_argument("--row-blocks", type=int, default=0,
                    help="peer exchange: row-block pipeline -- the halo travels in the order the shard's row blocks need it, "
                         "block b's SpMM starts when its part has landed (0 / 1 = off, default)")
    ap.add_argument("--exchange-slabs", type=int, default=1,
                    help="N > 1: all-gather X in this many feature slabs, slab k+1 in flight while the SpMM of slab k runs "
                         "(1 = one all-gather, then one SpMM)")
    ap.add_argument("--dense", action="store_true", help="tcgen05 dense super-window plan (with --classifier b200|all_tc)")
    ap.add_argument("--tune", action="append", default=[], help="library tuning knob key=value (repeatable)")
    ap.add_argument("--no-row-sort", action="store_true",
                    help="preprocess() does not keep the row-sorted copy of low-degree CSRs (A/B of csrc/rowsort.cu)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline only: skip extra.sweep / extra.products")
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


def cpu_spmm_baseline(rp_cpu, ci_cpu, x_cpu, seconds: float, keep_y: bool = False):
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
        dt, e, y = run(n_s)
        times.append(dt)
    dt = statistics.median(times)
    res = {"value": 2.0 * e * dim / dt / 1e9, "unit": "GFLOP/s", "cores": threads, "kind": "port",
           "sample": f"torch.sparse.mm CSR FP32 on CPU, rows [0,{n_s}) of {n} ({e} of {int(rp_cpu[-1])} "
                     f"stored entries) x full X[{x_cpu.shape[0]},{dim}], median of 2 runs, {dt:.2f} s each",
           "seconds": dt, "entries": e, "rows": n_s}
    if keep_y:          # the bench's parity leg compares the GPU result with this Y (rows [0, n_s))
        pub = {k: v for k, v in res.items() if k not in ("seconds", "entries", "rows")}
        return pub, y, n_s
    return res


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

    if args.slab >= 0:
        HCSPMM.set_tuning("slab", args.slab)
    if args.long_row >= 0:
        HCSPMM.set_tuning("long_row", args.long_row)
    if args.vec8 >= 0:
        HCSPMM.set_tuning("vec8", args.vec8)
    for kv in args.tune:
        k, v = kv.split("=")
        HCSPMM.set_tuning(k, int(v))
    if args.no_row_sort:
        HCSPMM.set_row_sort(False)

    ctx = Ctx(args, world, rank, dev)
    # the measured L2 -> SM gather roof (hcspmm_debug_l2_gather: random 1 KB rows of an L2-resident 32 MB buffer, the
    # SpMM's own 256-bit loads) -- the honest roof of a gather that L2 serves; HBM figures are reported beside it
    ctx.l2_peak = capi.l2_gather_bandwidth(dev, 256, 32)
    ctx.hbm_peak, ctx.hbm_src = measured_peak_gbs()

    head = run_workload(ctx, args.shape, args.dim or graphs.SHAPES[args.shape]["dim"], args.steps, args.warmup,
                        classifier=args.classifier, dense=args.dense, precision=args.precision, headline=True)
    extra = {}
    if not args.no_extra and args.shape == "reddit":
        k, w = max(3, min(args.steps, 10)), max(3, min(args.warmup, 3))
        if world > 1:
            # SURVEY 8(e) / north star: scaling on the ogbn-products shape (dim 128), same exchange as the headline
            extra["products"] = slim(run_workload(ctx, "products", 128, k, w, classifier=args.classifier, dense=False,
                                                  precision=args.precision))
        else:
            sweep = []
            for shape, dim, cls, dense in (("reddit", 64, "shipped", False), ("reddit", 128, "shipped", False),
                                           ("reddit", 512, "shipped", False), ("products", 128, "shipped", False),
                                           ("proteins", 256, "shipped", False), ("proteins", 256, "b200", True)):
                sweep.append(slim(run_workload(ctx, shape, dim, k, w, classifier=cls, dense=dense, precision="tf32")))
            extra["sweep"] = sweep
    if rank == 0:
        line = head
        line["extra"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


class Ctx:
    def __init__(self, args, world, rank, dev):
        self.args, self.world, self.rank, self.dev = args, world, rank, dev
        self.l2_peak = self.hbm_peak = None
        self.hbm_src = ""
        self._graphs = {}

    def graph(self, shape):
        from hcspmm import graphs
        if shape not in self._graphs:
            self._graphs.clear()                      # one generated graph resident at a time
            t = time.perf_counter()
            rp, ci, info = graphs.named(shape, device=self.dev, scale=self.args.scale)
            torch.cuda.synchronize()
            self._graphs[shape] = (rp, ci, info, time.perf_counter() - t)
        return self._graphs[shape]


def slim(rec):
    """A sub-record of the line: the numbers, without the prose."""
    keep = ("workload", "value", "unit", "ms_per_step", "steps", "dim", "nodes", "stored_entries", "classifier", "dense_groups_tcgen05",
            "precision", "roofline", "parity", "phases", "launches_per_step", "preprocess_ms", "step_ms")
    out = {k: rec[k] for k in keep if k in rec}
    cfg = rec.get("config", {})
    for k in ("workload", "dim", "nodes", "stored_entries", "classifier", "dense_groups_tcgen05", "phases", "preprocess_ms", "exchange"):
        if k in cfg:
            out[k] = cfg[k]
    return out


def rel_fro(a, b):
    den = float(torch.linalg.vector_norm(b.double()))
    return float(torch.linalg.vector_norm(a.double() - b.double())) / (den if den > 0 else 1.0)


def run_workload(ctx, shape_name, dim, steps, warmup, classifier, dense, precision, headline=False):
    """One workload = K timed aggregations Y = A X of one synthetic shape at one width, on `world` GPUs (row-window
    partition + exchange per step at world > 1), with its parity check and roofline."""
    import HCSPMM
    from hcspmm import dist as hd, graphs
    args, world, rank, dev = ctx.args, ctx.world, ctx.rank, ctx.dev
    if world > 1:
        import torch.distributed as dist
    shape = graphs.SHAPES[shape_name]
    rp, ci, info, t_gen = ctx.graph(shape_name)
    n, nnz = info["n"], info["nnz"]

    HCSPMM.set_dense(bool(dense))               # defaults recorded per graph by the preprocess calls below
    HCSPMM.set_classifier(classifier)
    HCSPMM.set_precision(precision)
    operand = "bf16" if (precision == "bf16" and world > 1) else "fp32"
    sg = hd.ShardedGraph(rp, ci, schedule=args.exchange, n_slabs=max(1, args.exchange_slabs), n_passes=args.exchange_passes,
                         operand=operand, direct_refs=None if args.direct_refs < 0 else args.direct_refs,
                         row_blocks=None if args.row_blocks <= 0 else args.row_blocks)
    sg.overlap_ctas = args.overlap_ctas
    r0, r1 = sg.r0, sg.r1
    rp_l, ci_run = sg.rowptr, sg.colidx
    n_l, nnz_l = sg.n_local, sg.nnz_local

    # preprocessing (reported separately, like the paper: "x one SpMM")
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    pre = HCSPMM.preprocess(ci_run, rp_l, n_l, nnz_l, (n_l + 15) // 16)
    ev1.record()
    torch.cuda.synchronize()
    prep_ms = ev0.elapsed_time(ev1)
    sg.pre = pre
    hdr = pre[4]
    tc_windows = int(hdr[7])
    dense_groups = int(hdr[1])

    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    x_full = torch.randn(n, dim, device=dev, generator=g)       # same on every rank (same seed)
    x_loc = x_full[r0:r1].contiguous()
    x_in_operand = False
    if world > 1 and not args.stage_copy:
        # the rank's shard of X lives where the exchange reads it: in its rows of the peer-visible operand buffer (what a
        # GCN layer's Update GEMM writes there directly, hcspmm.dist.ShardedGCNLayer) -- no per-step staging copy
        xo = sg.own_rows(dim)
        if xo is not None:
            xo.copy_(x_loc)
            x_loc, x_in_operand = xo, True
    n_slabs = sg.n_slabs if (world > 1 and sg.schedule in ("slabs", "halo", "peer")) else 1

    def step():
        # N > 1: exchange of the row shards of X (halo rows pulled over NVLink, or an NCCL schedule), then the local SpMM
        return sg.aggregate(x_loc if world > 1 else x_full)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        y = step()
    barrier()

    # timed region: exactly K steps, CUDA events on the launching (current) stream
    sampler = ClockSampler(dev.index)
    if rank == 0 and headline:
        sampler.start()
        time.sleep(0.25)
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t0 = time.time()
    ev0.record()
    for i in range(steps):
        kern_ev[i][0].record()
        y = step()
        kern_ev[i][1].record()
    ev1.record()
    barrier()
    t1 = time.time()
    total_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t0, t1) if (rank == 0 and headline) else None
    step_ms = [a.elapsed_time(b) for a, b in kern_ev]
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t)
    ms_per_step = total_ms / steps
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
        operand_t = sg.exchange(x_loc)
        esz = 2 if operand_t.dtype == torch.bfloat16 else 4
        phases = {"exchange_only_ms": tm(lambda: sg.exchange(x_loc)),
                  "kernel_only_ms": tm(lambda: sg.local_spmm(operand_t)),
                  "exchange_bytes_per_rank": int(sg.exchange_rows() * dim * esz),
                  "exchange_rows_vs_allgather": sg.exchange_rows() / max(1, (world - 1) * sg.max_rows)}
        if phases["exchange_only_ms"] > 0:
            phases["exchange_gbs_per_rank"] = phases["exchange_bytes_per_rank"] / phases["exchange_only_ms"] / 1e6
        if sg.blocks is not None:
            phases["row_blocks"] = {"blocks": sg.blocks["B"], "halo_fraction_first_block": round(sg.blocks["first_fraction"], 4),
                                    "halo_rows_per_block": [b_["rows"] for b_ in sg.blocks["list"]],
                                    "note": "kernel_only_ms = the blocks' SpMMs back to back on one stream; in a step they run on "
                                            "their own streams behind the pull of their part of the halo"}
        if sg.direct is not None:
            # segment mode: rows referenced <= T times are read in place by the SpMM (inside kernel_only_ms); only the
            # pulled rows travel in exchange_only_ms
            d_ = sg.direct
            phases["in_place"] = {"max_refs": d_["T"], "rows_in_place": d_["rows"], "references_in_place": d_["refs"],
                                  "rows_pulled": d_["pulled_rows"], "halo_rows": d_["halo_rows"]}
            pulled_bytes = d_["pulled_rows"] * dim * esz
            if phases["exchange_only_ms"] > 0:
                phases["exchange_gbs_per_rank"] = pulled_bytes / phases["exchange_only_ms"] / 1e6
    sg.check()          # a peer barrier that timed out would have left stale rows in the operand

    # ---- parity, outside the timed region, on the result of a real step ------------------------------------------
    tol = 1e-2 if precision == "bf16" else (1e-3 if (tc_windows or dense_groups) else 1e-5)
    y = step()
    parity = {"tol": tol}
    if world > 1:
        # every rank holds the full graph and the full X: the partitioned aggregate (exchange included) must equal rows
        # [r0, r1) of the SINGLE-GPU aggregation of the unpartitioned graph
        HCSPMM.set_precision("tf32")
        pre_full = HCSPMM.preprocess(ci, rp, n, nnz, (n + 15) // 16)
        y_full = HCSPMM.forward(x_full, rp, ci, *pre_full)[0]
        HCSPMM.set_precision(precision)
        e = torch.tensor([rel_fro(y, y_full[r0:r1])], device=dev, dtype=torch.float64)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        parity.update(rel_fro=float(e), vs="rows [r0, r1) of the single-GPU HCSPMM.forward of the unpartitioned graph, "
                                          "max over ranks")
        del pre_full, y_full
    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline:
        # torch.sparse.mm FP32 on the host cores: the north star's oracle at full size.  Headline at N = 1: the bounded
        # sample that is also the reported CPU baseline; otherwise a small row sample of this rank's rows.
        if world == 1 and headline:
            cpu_base, y_cpu, rows = cpu_spmm_baseline(rp.cpu(), ci.cpu(), x_full.cpu(), args.cpu_seconds, keep_y=True)
        else:
            rows = min(n_l, 8192)
            rp_c = rp[r0:r0 + rows + 1].cpu()
            e0, e1 = int(rp_c[0]), int(rp_c[-1])
            a = torch.sparse_csr_tensor((rp_c - e0).to(torch.int64), ci[e0:e1].cpu().to(torch.int64),
                                        torch.ones(e1 - e0, dtype=torch.float32), size=(rows, n))
            torch.set_num_threads(os.cpu_count() or 1)
            y_cpu = torch.sparse.mm(a, x_full.cpu())
        parity.update(cpu_rel_fro=rel_fro(y[:rows].cpu(), y_cpu), cpu_rows=int(rows),
                      cpu_vs="torch.sparse.mm CSR FP32 on the host, first %d of rank 0's rows" % rows)
        del y_cpu
    ok = all(v <= tol for k_, v in parity.items() if k_ in ("rel_fro", "cpu_rel_fro"))
    parity["ok"] = bool(ok)
    if not ok:      # the line still goes out, with parity.ok = false: a wrong result must be visible, not a missing line
        print(f"[bench] PARITY CHECK FAILED on {shape_name} dim {dim}: {parity}", file=sys.stderr, flush=True)

    # ---- launches of OUR kernels per step (library rules; per-graph products from preprocess) -------------------
    bal_knob = HCSPMM.set_tuning("balance", 1)
    HCSPMM.set_tuning("balance", bal_knob)
    balanced = bal_knob >= 2 or (bal_knob == 1 and nnz_l >= 8 * n_l)
    spmm_launches = (2 if balanced else 1) + (1 if (balanced and tc_windows > 0 and precision in ("tf32", "tf32x2")) else 0)
    if dense_groups > 0 and precision == "tf32":
        spmm_launches += 2                         # tf32_round_rows + spmm_dense_ws
    launches_per_step = spmm_launches * n_slabs + (1 if (precision == "bf16" and (world == 1 or dim % 8)) else 0) * n_slabs
    if world > 1 and sg.schedule in ("peer", "push"):
        launches_per_step += 1 + n_slabs + (1 if operand == "bf16" and dim % 8 == 0 else 0)   # barrier + pull(s) / push (+ f32->bf16 of own rows)
    if world > 1 and sg.blocks is not None:     # row-block pipeline: per block one pull (if it brings rows) + the SpMM launches
        launches_per_step = 1 + (1 if operand == "bf16" and dim % 8 == 0 else 0) + sum(
            (1 if b_["rows"] > 0 else 0) + spmm_launches for b_ in sg.blocks["list"])

    # ---- roofline of the dominant kernel (the SpMM launch of a step) ------------------------------------------------
    esz_x = 2 if precision == "bf16" else 4
    bytes_alg = nnz_l * (4 + esz_x * dim) + n_l * (4 * dim + 4)
    bytes_min = 4 * nnz_l + 4 * (n_l + 1) + esz_x * min(n, sg.x_rows) * dim + 4 * n_l * dim
    kern_ms = statistics.mean(step_ms) if world == 1 else (phases or {}).get("kernel_only_ms")
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(f"{shape_name}_dim{dim}_{classifier}{'_dense' if dense_groups else ''}") if world == 1 else None
        if isinstance(ent, dict):
            traffic, traffic_src = ent.get("bytes"), ent.get("source")
        elif ent is not None:
            traffic = ent
    except Exception:
        pass
    roofline = None
    if kern_ms and dense_groups > 0 and world == 1:
        # dense super-windows: the tensor pipe is the roof.  EXECUTED flops = 2 * 128 * (condensed columns, padded to
        # 32) * dim per super-window (useful = 2 * nnz * dim is `value`); peak = half the measured cuBLAS BF16 burst
        # figure (TF32 issues at half the BF16 rate on tcgen05; MEASURED_PEAKS.json holds BF16 only)
        total_cols = int(hdr[2])
        exec_tflops = 2.0 * 128 * total_cols * dim / (kern_ms * 1e-3) / 1e12
        try:
            tf32_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]) / 2
            psrc = "measured cuBLAS BF16 burst (MEASURED_PEAKS.json) / 2: TF32 issues at half the BF16 rate"
        except Exception:
            tf32_peak, psrc = 1590.0 / 2, "fallback BF16 figure (B200_PROFILING.md) / 2"
        roofline = {"bound": "tensor", "achieved": exec_tflops, "peak": tf32_peak, "unit": "TFLOP/s", "frac": exec_tflops / tf32_peak,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": psrc,
                    "kernel": "spmm_dense_ws_kernel (tcgen05.mma kind::tf32, M128 N=dim K8), whole step incl. tf32_round_rows",
                    "kernel_ms": kern_ms, "executed_flops": 2.0 * 128 * total_cols * dim, "useful_flops": flops,
                    "executed_over_useful": 128.0 * total_cols / max(1, nnz),
                    "gathered_bytes": total_cols * dim * 4, "gather_gbs": total_cols * dim * 4 / (kern_ms * 1e-3) / 1e9,
                    "hbm_peak": ctx.hbm_peak, "compulsory_bytes": bytes_min,
                    "compulsory_frac": bytes_min / (kern_ms * 1e-3) / 1e9 / ctx.hbm_peak, "scope": "whole graph, one GPU"}
    elif kern_ms:
        ach = bytes_alg / (kern_ms * 1e-3) / 1e9
        kernel = ("spmm_dense_ws_kernel (tcgen05) + spmm_balanced_kernel for the remaining windows" if dense_groups else
                  "spmm_balanced_kernel (+ spmm_balanced_fixup_kernel: the step)" if balanced else "spmm_hybrid_kernel")
        roofline = {"bound": "l2", "achieved": ach, "peak": ctx.l2_peak, "unit": "GB/s", "frac": ach / ctx.l2_peak,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": "measured in this run: hcspmm_debug_l2_gather, random 1 KB rows of an L2-resident 32 MB "
                                   "buffer with the SpMM's 256-bit evict_last loads",
                    "why_l2": "the gather is served by L2 (ncu: 82 % sector hit rate at the Reddit shape, XBAR->L1 bytes = "
                              "algorithmic bytes); against the HBM copy peak the same number reads hbm_alg_frac > 1",
                    "kernel": kernel, "kernel_ms": kern_ms, "algorithmic_bytes": bytes_alg,
                    "hbm_peak": ctx.hbm_peak, "hbm_peak_source": ctx.hbm_src,
                    "hbm_alg_frac": ach / ctx.hbm_peak,
                    "hbm_actual_frac": (traffic / (kern_ms * 1e-3) / 1e9 / ctx.hbm_peak) if traffic else None,
                    "compulsory_bytes": bytes_min, "compulsory_frac": bytes_min / (kern_ms * 1e-3) / 1e9 / ctx.hbm_peak,
                    "scope": "whole graph, one GPU" if world == 1 else
                             "this rank's row shard on one GPU (local SpMM timed alone, max over ranks)"}

    # ---- e2e (headline only): the reference-facing module with HOST buffers ------------------------------------------
    e2e = None
    if headline and not args.no_e2e:
        rows_in = n if world == 1 else n_l
        xh = torch.empty(rows_in, dim, pin_memory=True)
        xh.copy_(x_full if world == 1 else x_loc)
        yh = torch.empty(n_l, dim, pin_memory=True)
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
        k2 = max(4, steps)
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
               "kind": "THROUGHPUT of a double-buffered pipeline (H2D of step i+1 | kernels of step i | D2H of step i-1 on three "
                       "streams), every step's copies inside the timed region; the latency of one isolated step is "
                       "H2D + kernel + D2H in series",
               "pipeline": "pinned host X -> H2D stream | HCSPMM.forward | D2H stream -> pinned host Y",
               "h2d_bytes_per_step": rows_in * dim * 4 * world if world > 1 else rows_in * dim * 4,
               "d2h_bytes_per_step": n * dim * 4}

    if world > 1:
        sg.close()
    workload = (f"{shape_name}-shape graph, single-kernel SpMM Y=A*X, dim {dim}" +
                (" (BASELINE.json configs[1])" if shape_name == "reddit" and dim == 256 else ""))
    gens = {"rmat": "R-MAT(0.57,0.19,0.19,0.05) folded mod N, symmetrised, de-duplicated, ids permuted, seed %d",
            "sbm": "stochastic block model: 16-aligned communities of 512 vertices (p_in 0.55) + 15 random neighbours per "
                   "vertex, ids NOT permuted, seed %d",
            "ring": "ring + one random perfect matching (every degree 3), seed %d"}
    exchange_desc = {"gather": "NCCL all_gather_into_tensor of the row shards of X per step (the north star's schedule; kept as "
                               "the comparison: --exchange gather)",
                     "slabs": "NCCL all_gather_into_tensor in %d feature slabs pipelined with the SpMM" % n_slabs,
                     "halo": "halo rows only: pack + NCCL all_to_all_single per step, %d feature slab(s)" % n_slabs,
                     "peer": "halo rows only (%s), pulled from the owners' memory over NVLink by hcspmm_halo_pull after "
                             "hcspmm_peer_barrier, %d feature slab(s), %d source pass(es) -- moves fewer bytes than the north "
                             "star's all-gather (exchange_rows_vs_allgather)" % (operand, n_slabs, 2 if sg.passes is not None else 1),
                     "push": "halo rows only (%s), written into the consumers' operand buffers over NVLink by their owner "
                             "(hcspmm_halo_push), then hcspmm_peer_barrier" % operand}
    line = {"metric": "spmm_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16-stored X, f32 accumulate" if precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": workload, "nodes": n, "stored_entries": nnz, "dim": dim, "classifier": classifier,
                       "precision": precision, "tc_windows_label1": tc_windows, "dense_groups_tcgen05": dense_groups,
                       "windows": (n_l + 15) // 16, "partition": f"row windows, nnz-balanced, {world} shard(s)",
                       "exchange": exchange_desc[sg.schedule] if world > 1 else "none",
                       "phases": phases,
                       "x_placement": None if world == 1 else (
                           "each rank's row shard of X lives in its rows of the peer-visible exchange operand (no staging copy; "
                           "--stage-copy adds the copy)" if x_in_operand else "ordinary device memory, copied into the exchange operand every step"),
                       "l2": "inputs larger than L2 (X %.0f MB + CSR %.0f MB vs 126 MB), no flush" %
                             (n * dim * 4 / 1e6, nnz * 4 / 1e6),
                       "preprocess_ms": prep_ms, "graph_gen_s": t_gen,
                       "generator": gens[shape["kind"]] % shape["seed"]},
            "roofline": roofline, "parity": parity, "cpu_baseline": cpu_base, "e2e": e2e, "clocks": clocks,
            "gpu_launches": steps * launches_per_step, "launches_per_step": launches_per_step,
            "step_ms": {"min": min(step_ms), "median": statistics.median(step_ms), "max": max(step_ms)}}
    del sg, x_full, x_loc, y
    torch.cuda.empty_cache()
    return line


if __name__ == "__main__":
    sys.exit(main())
