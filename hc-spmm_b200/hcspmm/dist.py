"""Multi-GPU aggregation: A partitioned by row windows, X exchanged per layer.

One process per GPU (torchrun), `torch.distributed` for the plumbing.  The reference has no
multi-GPU code at all; this layer sits above the HCSPMM entry points and changes nothing below.

Rank r owns the 16-row windows [cuts[r], cuts[r+1]) of A (nnz-balanced, hcspmm.partition), the
matching rows of X / Y, and the preprocessing of its shard.  One aggregation Y_r = A_r * X is
    exchange:  the rows of X the shard references reach the rank (default: pulled over NVLink peer
               memory by our own kernels; NCCL collectives as alternatives / fallback)
    compute:   local hybrid SpMM on the exchanged operand (rectangular: n_r x operand rows)
Schedules:
  * "gather":   one all_gather_into_tensor, then one SpMM launch;
  * "slabs":    X is exchanged in feature slabs; the SpMM of slab k (a strided view of the gathered
                buffer, the kernel takes ldx/ldy) runs on the compute stream while slab k+1 is in
                flight on the communication stream -- the exchange hides behind the gather-bound
                kernel instead of preceding it.
  * "halo":     only the rows of X a shard actually references travel: at set-up every rank tells
                every owner which of its rows it needs (the distinct remote column ids of the
                shard); per aggregation the owners pack those rows and one all_to_all_single moves
                them.  On power-law graphs most vertices have few neighbours, so a 1/8 row shard of
                the products shape references 0.31 N remote rows where the all-gather moves 0.875 N.
                `n_slabs` > 1 pipelines the halo exchange in feature slabs like "slabs".
  * "peer":     the halo rows again, but PULLED: every rank keeps its shard of X in a CUDA-IPC buffer
                the others have mapped, and one kernel of ours (hcspmm_halo_pull, csrc/peer.cu) copies
                the referenced rows straight out of the owners' memory with NVLink loads, after a
                flag barrier of ours (hcspmm_peer_barrier).  No collective, no packing pass on the
                owner, nothing sent that is not needed.  `n_slabs` > 1 pulls feature slab k+1 on a
                high-priority stream while the SpMM of slab k runs.
  * "push":     the same halo rows over the same peer-mapped buffers, but WRITTEN by their owner (hcspmm_halo_push)
                before the barrier instead of read by the consumer after it: NVLink stores are posted, loads wait
                for a response.
  * "auto":     CUDA: "peer".  Otherwise "halo" when it moves <= 0.6 of the all-gather's rows
                (decided once, collectively), else "gather".
gather / slabs: shards are padded to the largest shard so the collective is a plain equal-size
all-gather; shard s occupies rows [s*max_rows, s*max_rows + n_s) of the gathered buffer.  halo:
the operand is [halo rows of rank 0 | ... | own rows | ... | halo rows of rank P-1], ascending
global ids inside each part.  Either way the local column ids are remapped once and the remap is
monotone, so the window-local column ranks (edgeToColumn), block counts and labels are exactly
those of the unpartitioned graph.

The aggregation operator is injectable (`spmm=`) so that the exchange / remap / autograd logic is
testable on CPU with gloo; the default is the CUDA path (HCSPMM), which has no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import partition


def _default_spmm():
    import HCSPMM

    def run(x, rowptr, colidx, pre, out=None, accumulate=False):
        if x.dtype == torch.bfloat16:            # BF16-stored exchange operand
            if out is None:
                out = torch.empty(rowptr.numel() - 1, x.shape[1], device=x.device)
            return HCSPMM.spmm_bf16(x, rowptr, colidx, *pre[:4], out, accumulate, *pre[4:6])
        if out is None:
            return HCSPMM.forward(x, rowptr, colidx, *pre)[0]
        return HCSPMM.spmm_strided(x, rowptr, colidx, *pre[:4], out, accumulate, *pre[4:6])

    def prep(colidx, rowptr):
        n = rowptr.numel() - 1
        return HCSPMM.preprocess(colidx, rowptr, n, colidx.numel(), (n + 15) // 16)

    return run, prep


class ShardedGraph:
    def __init__(self, rowptr: torch.Tensor, colidx: torch.Tensor, group=None, schedule: str = "gather",
                 n_slabs: int = 1, spmm=None, preprocess=None, cuts=None, n_passes: int = 1,
                 operand: str = "fp32", single: bool = False, direct_refs: int | None = 0,
                 row_blocks: int | None = None):
        """rowptr / colidx: the FULL graph's CSR on this rank's device (identical on every rank).
        cuts: reuse another ShardedGraph's row cuts (the transposed graph for backward must be
        partitioned like the forward one).
        n_passes = 2 ("peer" schedule only): the shard is cut by SOURCE into two CSRs -- own rows + the nearer half
        of the owners, and the farther half -- and aggregated in two accumulating SpMM passes, the second
        half of the pull travelling (few CTAs, high-priority stream) while the first pass computes.
        operand = "bf16" ("peer" schedule, widths that are multiples of 8): the exchange operand is stored as
        bfloat16 -- every rank rounds its own rows once (RNE), the halo travels at half the bytes and the local
        SpMM gathers bfloat16 rows with FP32 accumulation: the BF16 precision mode of the single-GPU path
        (north star: 1e-2 against FP32), end to end.  "fp32" (default) is exact.
        direct_refs = T ("peer" schedule): remote rows of X that this shard references at most T times are NOT pulled
        into the local operand; the SpMM's gather reads them in place from their owner's peer-mapped operand
        (segment-tagged column ids, HCSPMM.spmm_segments) -- on a power-law graph 40 % of a 1/8 shard's halo rows
        are referenced exactly once (4.5 % of its remote references), so the pull shrinks by that much and those
        rows are neither written to nor re-read from local memory.  Measured (profiles/README.md): the in-order
        gather ring stalls on the NVLink latency of those loads, so the SpMM loses about what the pull saves -- T = 1
        gains 7 % on the products shape at 8 GPUs, larger T and denser graphs lose.  0 (default) = off; None = 2 when at
        least a quarter of the halo rows qualify and the shard is all CUDA-core.
        row_blocks = B ("peer" schedule): the ROW-BLOCK PIPELINE.  The shard's rows are cut into B blocks of equal stored
        entries and the halo is pulled in the order the blocks need it: block b's SpMM (own stream) starts as soon as
        the rows first referenced by blocks <= b have landed, while the rest still travels on few CTAs -- on low-degree
        graphs the first of 8 blocks needs only a third of the halo.  None / 1 = off (default: the measured gain, 5 % with
        two blocks on the products shape at 8 GPUs, is inside the run-to-run spread)."""
        assert operand in ("fp32", "bf16")
        self.operand = operand
        self.group = group
        # single = True: the whole graph on this process even inside an initialised process group (the
        # single-GPU reference a multi-GPU run compares its losses with)
        self.world = dist.get_world_size(group) if (dist.is_initialized() and not single) else 1
        self.rank = dist.get_rank(group) if (dist.is_initialized() and not single) else 0
        self.n = rowptr.numel() - 1
        self.cuts = list(cuts) if cuts is not None else partition.window_cuts(rowptr, self.world)
        self.r0, self.r1 = self.cuts[self.rank], self.cuts[self.rank + 1]
        self.n_local = self.r1 - self.r0
        self.max_rows = max(self.cuts[i + 1] - self.cuts[i] for i in range(self.world))
        self.schedule, self.n_slabs = schedule, max(1, n_slabs)
        rp_l, ci_l = partition.local_shard(rowptr, colidx, self.r0, self.r1)
        dev = colidx.device
        bounds = torch.tensor(self.cuts, device=dev, dtype=torch.int64)
        ci64 = ci_l.to(torch.int64)
        self.rowptr = rp_l
        self.nnz_local = ci_l.numel()
        self.halo = None
        self.peer = None
        self.push = schedule == "push"
        if self.push:
            schedule = "peer"
        if self.world > 1 and schedule in ("auto", "peer") and dev.type == "cuda":
            # NVLink peer memory (CUDA IPC between the ranks' processes); "auto" falls back to the NCCL
            # schedules, on every rank together, where the platform refuses it
            from . import peer
            try:
                self.peer = peer.PeerMemory(dev, group)
                schedule = "peer"
            except peer.capi.HcspmmError:
                if schedule == "peer":
                    raise
        if self.world > 1 and schedule in ("halo", "auto", "peer"):
            self._setup_halo(ci64, bounds, dev, lists_only=schedule == "peer" and not self.push)
            if schedule == "auto" and self.halo["ratio"] > 0.6:
                self.halo = None
            self.schedule = schedule if schedule == "peer" else (
                "halo" if self.halo is not None else ("slabs" if self.n_slabs > 1 else "gather"))
        if self.peer is not None:
            self._pull = peer.halo_pull
            self._hi_stream = torch.cuda.Stream(device=dev, priority=-1)
        if self.halo is not None:
            self.colidx = torch.searchsorted(self.halo["ids"], ci64).to(torch.int32).contiguous()
        else:
            owner = torch.bucketize(ci64, bounds[1:-1], right=True)
            self.colidx = (ci64 - bounds[owner] + owner * self.max_rows).to(torch.int32).contiguous()
        self._fused, self._gemm = None, torch.mm
        native = spmm is None
        if spmm is None:
            spmm, preprocess = _default_spmm()
            import HCSPMM
            self._fused = lambda x, rp_, ci_, pre_, w: HCSPMM.forward_fixed32_fused(x, rp_, ci_, *pre_, w)
            self._gemm = lambda a, b: HCSPMM.gemm_tf32(a.contiguous(), b.contiguous())
        self._spmm = spmm
        self.pre = preprocess(self.colidx, self.rowptr) if preprocess is not None else ()
        self.direct = None
        if (native and self.peer is not None and not self.push and self.halo is not None and direct_refs != 0
                and self.world <= 8 and n_passes == 1 and self.n_slabs == 1):
            self._setup_direct(ci64, bounds, direct_refs)
        del ci64
        self.blocks = None
        if (native and self.peer is not None and not self.push and self.halo is not None and row_blocks not in (None, 1)
                and n_passes == 1 and self.n_slabs == 1):
            self._setup_row_blocks(row_blocks, preprocess)
        self.passes = None
        self.overlap_ctas = 64
        if self.peer is not None and n_passes == 2 and self.world > 2:
            self._setup_passes(preprocess)
        self._bufs = {}
        self._comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        if self.push and self.peer is not None:
            self.schedule = "push"

    def _setup_halo(self, ci64, bounds, dev, lists_only=False):
        """Which rows of X this shard needs from every owner, and which of its own rows every peer needs."""
        uniq, refs = torch.unique(ci64, return_counts=True)                   # ascending global ids
        own = torch.bucketize(uniq, bounds[1:-1], right=True)
        self._uniq_refs = (uniq, refs, own) if lists_only else None           # _setup_direct
        mine = torch.arange(self.r0, self.r1, device=dev, dtype=torch.int64)  # own rows: all of them
        ids = torch.cat([uniq[own < self.rank], mine, uniq[own > self.rank]])
        own = torch.bucketize(ids, bounds[1:-1], right=True)
        recv_counts = torch.bincount(own, minlength=self.world)               # rows arriving from each owner
        want = (ids - bounds[own]).contiguous()                               # row index at its owner
        moved = torch.tensor([ids.numel() - self.n_local], device=dev, dtype=torch.int64)
        dist.all_reduce(moved, op=dist.ReduceOp.MAX, group=self.group)
        ratio = float(moved) / max(1, (self.world - 1) * self.max_rows)
        if lists_only:      # pull model: the owners need no send lists
            seg = torch.zeros(self.world + 1, dtype=torch.int32, device=dev)
            seg[1:] = torch.cumsum(recv_counts, 0).to(torch.int32)
            self.halo = dict(ids=ids, rows=int(ids.numel()), recv=recv_counts.tolist(), ratio=ratio,
                             src_row=want.to(torch.int32), seg=seg)
            return
        send_counts = torch.empty_like(recv_counts)
        dist.all_to_all_single(send_counts, recv_counts, group=self.group)    # rows every peer wants from me
        rc, sc = recv_counts.tolist(), send_counts.tolist()
        send_idx = torch.empty(sum(sc), dtype=torch.int64, device=dev)
        dist.all_to_all_single(send_idx, want, output_split_sizes=sc, input_split_sizes=rc, group=self.group)
        self.halo = dict(ids=ids, rows=int(ids.numel()), recv=rc, send=sc, send_idx=send_idx, ratio=ratio)
        if self.push:       # owner-side lists of the push schedule + the consumer-side layout of the pull model
            seg = torch.zeros(self.world + 1, dtype=torch.int32, device=dev)
            seg[1:] = torch.cumsum(recv_counts, 0).to(torch.int32)
            send_seg = torch.zeros(self.world + 1, dtype=torch.int32, device=dev)
            send_seg[1:] = torch.cumsum(send_counts, 0).to(torch.int32)
            self.halo.update(seg=seg, src_row=want.to(torch.int32), send_seg=send_seg,
                             send_row=send_idx.to(torch.int32).contiguous())

    def _setup_direct(self, ci64, bounds, direct_refs):
        """Segment mode (see __init__, direct_refs): split the halo into rows that are pulled (referenced more than T
        times) and rows the SpMM reads in place from their owner's operand.  Rebuilds the pull lists for the pulled
        rows only and writes `colidx_seg`: entries of pulled / own rows address the local operand (segment 0), entries
        of in-place rows carry segment (owner - rank) mod P in bits 29..31 and the row inside the OWNER's operand."""
        uniq, refs, own = self._uniq_refs
        self._uniq_refs = None
        hdr = self.pre[4] if len(self.pre) >= 6 and self.pre[4].device.type == "cpu" else None
        all_cuda = hdr is not None and int(hdr[1]) == 0 and int(hdr[7]) == 0        # no dense plan, no label-1 windows
        T = 2 if direct_refs is None else int(direct_refs)
        remote = own != self.rank
        cold = remote & (refs <= T)
        n_remote = int(remote.sum())
        want_it = all_cuda and n_remote > 0 and (direct_refs is not None or 4 * int(cold.sum()) >= n_remote)
        flag = torch.tensor([1 if want_it else 0], device=uniq.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)             # every owner's layout changes: together
        if int(flag) == 0:
            return
        dev = uniq.device
        ids, _ = partition.pulled_layout(ci64, bounds, self.rank, T)
        own_ids = torch.bucketize(ids, bounds[1:-1], right=True)
        recv_counts = torch.bincount(own_ids, minlength=self.world)
        seg = torch.zeros(self.world + 1, dtype=torch.int32, device=dev)
        seg[1:] = torch.cumsum(recv_counts, 0).to(torch.int32)
        firsts = [None] * self.world            # where every rank's own rows start in ITS operand
        rows_all = [None] * self.world          # and how many rows that operand has
        dist.all_gather_object(firsts, int(seg[self.rank]), group=self.group)
        dist.all_gather_object(rows_all, int(ids.numel()), group=self.group)
        self.colidx_seg = partition.tag_segments(ci64, ids, bounds, self.rank, self.world, firsts)
        pulled_before = self.halo["rows"] - self.n_local
        self.halo = dict(ids=ids, rows=int(ids.numel()), recv=recv_counts.tolist(), ratio=self.halo["ratio"],
                         src_row=(ids - bounds[own_ids]).to(torch.int32).contiguous(), seg=seg)
        self.direct = dict(T=T, rows=int(cold.sum()), refs=int(refs[cold].sum()), pulled_rows=self.halo["rows"] - self.n_local,
                           halo_rows=pulled_before, x_rows=max(rows_all), firsts=firsts)

    def _setup_row_blocks(self, row_blocks, preprocess):
        """Row-block pipeline (see __init__, row_blocks): per block its rebased CSR + preprocessing products, and the
        part of the halo it is the first to reference, as a pull list with explicit destination rows."""
        # measured on the products shape at 8 GPUs: 2 blocks 1.18 -> 1.125 ms in the sweep and 1.194 ms in the bench line,
        # 4 blocks 1.17, 8 blocks 1.24 (the blocks' SpMMs fill the machine worse than one launch); BF16 operand: loses
        # (0.90 -> 0.98 ms).  A gain inside the run-to-run spread is not a default: opt-in.
        B = int(row_blocks)
        hdr = self.pre[4] if len(self.pre) >= 6 and self.pre[4].device.type == "cpu" else None
        all_cuda = hdr is not None and int(hdr[1]) == 0 and int(hdr[7]) == 0        # no dense plan, no label-1 windows
        if B < 2 or not all_cuda or self.n_local < 16 * B or self.nnz_local < 64 * B:
            return
        h = self.halo
        col = self.colidx_seg if self.direct is not None else self.colidx        # tagged ids: only segment 0 is local
        cuts, first = partition.row_block_order(self.rowptr, col, h["rows"], B)
        own0 = int(h["seg"][self.rank])
        first[own0: own0 + self.n_local] = B + 1                                  # own rows are in place already
        n_halo = h["rows"] - self.n_local
        f1 = float((first == 0).sum()) / max(1, n_halo)
        dev = col.device
        rp64 = self.rowptr.to(torch.int64)
        seg64 = h["seg"].to(torch.int64)
        blocks = []
        for b in range(B):
            c0, c1 = cuts[b], cuts[b + 1]
            if c1 <= c0:
                continue
            e0, e1 = int(rp64[c0]), int(rp64[c1])
            rp_b = (rp64[c0:c1 + 1] - e0).to(torch.int32).contiguous()
            ci_b = self.colidx[e0:e1].clone()
            blk = dict(r0=c0, r1=c1, rowptr=rp_b, colidx=ci_b, pre=preprocess(ci_b, rp_b),
                       colidx_seg=self.colidx_seg[e0:e1].clone() if self.direct is not None else None)
            dst = torch.nonzero(first == b).flatten()                             # ascending operand rows = owner order
            seg_b = torch.searchsorted(dst, seg64).to(torch.int32).contiguous()   # entries of owner o: [seg_b[o], seg_b[o+1])
            blk.update(dst_row=dst.to(torch.int32).contiguous(), src_row=h["src_row"][dst].contiguous(), seg=seg_b,
                       rows=int(dst.numel()))
            blocks.append(blk)
        self.blocks = dict(B=len(blocks), list=blocks, first_fraction=f1,
                           streams=[torch.cuda.Stream(device=dev) for _ in blocks])

    # ------------------------------------------------------------------------------------------
    def shard_rows(self, x_full: torch.Tensor) -> torch.Tensor:
        return x_full[self.r0:self.r1].contiguous()

    def _buffers(self, dim, device, dtype):
        key = (dim, device, dtype)
        if key not in self._bufs:
            self._bufs[key] = (torch.zeros(self.max_rows, dim, device=device, dtype=dtype),
                               torch.empty(self.world * self.max_rows, dim, device=device, dtype=dtype))
        return self._bufs[key]

    def aggregate(self, x_local: torch.Tensor) -> torch.Tensor:
        """Y_r = A_r * X from the row shard X_r of every rank."""
        assert x_local.shape[0] == self.n_local
        dim = x_local.shape[1]
        if self.world == 1:
            return self._spmm(x_local.contiguous(), self.rowptr, self.colidx, self.pre)
        if self.peer is not None:
            self.peer.check()            # pinned-host flag, no synchronisation: a barrier of an EARLIER step timed out
            return self._aggregate_peer(x_local)
        if self.halo is not None:
            return self._aggregate_halo(x_local)
        if self.schedule == "slabs" and self.n_slabs > 1 and dim >= 8 * self.n_slabs and self._comm_stream is not None:
            return self._aggregate_slabs(x_local)
        return self._spmm(self.exchange(x_local), self.rowptr, self.colidx, self.pre)

    def aggregate_fused(self, x_local: torch.Tensor, weights: torch.Tensor):
        """(out, Z) = ((A_r X) W, A_r X): the exchange, then the reference's fused Aggregation + Update entry point
        (HCSPMM.forward_fixed32_fused: one tcgen05 kernel when the shard's dense plan covers it, else aggregation +
        TMA Update GEMM) on the exchanged operand.  Injected operators (CPU tests) and BF16 operands take the two
        steps separately."""
        if self._fused is None or self.direct is not None:     # segment mode: the operand alone does not hold every row
            z = self.aggregate(x_local)
            return self._gemm(z, weights), z
        if self.world == 1:
            operand = x_local.contiguous()
        else:
            if self.peer is not None:
                self.peer.check()
            operand = self.exchange(x_local)
            if operand.dtype != torch.float32 or not operand.is_contiguous():
                z = self.aggregate(x_local)
                return self._gemm(z, weights), z
        out, z = self._fused(operand, self.rowptr, self.colidx, self.pre, weights)
        return out, z

    def update(self, h: torch.Tensor, weights: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """The row-local Update product H W on the library's TF32 GEMM (TMA + tcgen05); `out`: e.g. own_rows(width),
        so that the product lands in the exchange operand of the aggregation that follows."""
        if out is None or self._fused is None:
            return self._gemm(h, weights)
        import HCSPMM
        return HCSPMM.gemm_tf32(h.contiguous(), weights.contiguous(), out)

    @property
    def x_rows(self) -> int:
        """Rows of the exchanged operand the local SpMM reads (column ids address this layout)."""
        if self.world == 1:
            return self.n
        return self.halo["rows"] if self.halo is not None else self.world * self.max_rows

    def exchange(self, x_local: torch.Tensor) -> torch.Tensor:
        """The exchange step on its own (one piece, not pipelined): the operand of the local SpMM.  The result is
        one of the graph's reusable exchange buffers: consume it before the next exchange of the same width."""
        dim, dev, dt = x_local.shape[1], x_local.device, x_local.dtype
        if self.world == 1:
            return x_local
        if self.peer is not None:
            cat, dpad = self._peer_stage(x_local)
            self._pull_halo(cat, dpad)
            return cat if dpad == dim else cat[:, :dim]   # bfloat16 rows when operand = "bf16"
        if self.halo is not None:
            h = self.halo
            key = ("halo", dim, dev, dt)
            if key not in self._bufs:
                self._bufs[key] = torch.empty(h["rows"], dim, device=dev, dtype=dt)
            cat = self._bufs[key]
            packed = x_local.index_select(0, h["send_idx"])
            dist.all_to_all_single(cat, packed, output_split_sizes=h["recv"], input_split_sizes=h["send"],
                                   group=self.group)
            return cat
        pad, gathered = self._buffers(dim, dev, dt)
        pad[: self.n_local].copy_(x_local)
        dist.all_gather_into_tensor(gathered, pad, group=self.group)
        return gathered

    def exchange_rows(self) -> int:
        """Rows of X this rank receives from its peers per aggregation."""
        if self.world == 1:
            return 0
        if self.direct is not None:       # pulled rows + one row per in-place reference: what crosses NVLink
            return self.direct["pulled_rows"] + self.direct["refs"]
        return self.halo["rows"] - self.n_local if self.halo is not None else (self.world - 1) * self.max_rows

    def check(self):
        """Raise if a peer barrier ever timed out (a rank did not reach an exchange): results would be stale."""
        if self.peer is not None:                                # collective: every rank raises, or none does
            torch.cuda.synchronize(self.peer.device)
            flag = self.peer.err.to(self.peer.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
            if int(flag.item()) != 0:
                raise RuntimeError("peer barrier timed out on some rank: a rank did not reach an exchange")

    def close(self):
        if self.peer is not None:
            self._bufs.clear()
            self.peer.close()
            self.peer = None

    def _setup_passes(self, preprocess):
        """Cut the shard by source: pass 0 = own rows + the owners rank+1 .. rank+h (mod P), pass 1 = the rest."""
        P, r, seg = self.world, self.rank, self.halo["seg"].to(torch.int64)
        h = (P - 1 + 1) // 2
        near = [(r + i) % P for i in range(1, h + 1)]
        far = [o for o in range(P) if o != r and o not in near]
        owner = torch.bucketize(self.colidx.to(torch.int64), seg[1:-1], right=True)
        in0 = torch.zeros(P, dtype=torch.bool, device=owner.device)
        in0[[r] + near] = True
        m0 = in0[owner]
        rows = torch.repeat_interleave(torch.arange(self.n_local, device=owner.device),
                                       (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64))
        self.passes = []
        for m, owners in ((m0, near), (~m0, far)):
            rp = torch.zeros(self.n_local + 1, dtype=torch.int64, device=owner.device)
            rp[1:] = torch.cumsum(torch.bincount(rows[m], minlength=self.n_local), 0)
            rp, ci = rp.to(torch.int32), self.colidx[m].contiguous()
            self.passes.append(dict(rowptr=rp, colidx=ci, pre=preprocess(ci, rp), mask=sum(1 << o for o in owners)))

    def _peer_bufs(self, dim: int):
        """The two peer-visible operand buffers of this aggregation width (created collectively on first use)."""
        b16 = self.operand == "bf16" and dim % 8 == 0
        dpad = dim if b16 else (dim + 7) // 8 * 8                 # 32-byte FP32 rows: the 256-bit gather path
        esz = 2 if b16 else 4
        key = ("peer", dpad, b16)
        h = self.halo
        if key not in self._bufs:
            dev = self.peer.device
            pm, slots = self.peer, []
            own0 = int(h["seg"][self.rank])
            firsts = [None] * self.world                         # every rank's own-segment offset in ITS operand
            dist.all_gather_object(firsts, own0, group=self.group)
            segs = [None] * self.world                           # push: where MY rows start in every peer's operand
            dist.all_gather_object(segs, h["seg"].tolist(), group=self.group)
            for _ in range(2):                                   # alternate: see csrc/peer.cu header
                ptr, ptrs = pm.shared(h["rows"] * dpad * esz)
                if self.push:
                    table = torch.tensor([p + segs[s][self.rank] * dpad * esz for s, p in enumerate(ptrs)],
                                         dtype=torch.int64, device=dev)
                else:
                    table = torch.tensor([p + f * dpad * esz for p, f in zip(ptrs, firsts)], dtype=torch.int64, device=dev)
                t = pm.tensor(ptr, (h["rows"], dpad), torch.int16).view(torch.bfloat16) if b16 else \
                    pm.tensor(ptr, (h["rows"], dpad))
                # segment s = the operand of rank (rank + s) mod P, from its first row (segment mode)
                segs_x = [0] + [ptrs[(self.rank + t_) % self.world] for t_ in range(1, self.world)]
                slots.append((t, table, segs_x))
            self._bufs[key] = dict(slots=slots, turn=0, own0=own0)
        return self._bufs[key], dpad, b16

    def own_rows(self, dim: int):
        """Where this rank's rows of the NEXT exchange operand of width `dim` live: a [n_local, dim] FP32 view of
        peer-visible memory.  A producer that writes X there (the Update GEMM of a GCN layer: `update(h, w, out=...)`,
        or a caller that keeps a static X there) hands `aggregate` that view and no staging copy is made.  Stream
        ordered: write it on the stream the previous aggregation ran on.  None when the exchange does not go through
        peer memory (one rank, NCCL schedules) or the operand is stored as bfloat16."""
        if self.peer is None or self.world == 1:
            return None
        b, dpad, b16 = self._peer_bufs(dim)
        if b16:
            return None
        cat = b["slots"][b["turn"]][0]
        return cat[b["own0"]: b["own0"] + self.n_local, :dim]

    def _peer_stage(self, x_local: torch.Tensor):
        """Write the shard into the own-rows segment of this width's next operand buffer (peer-visible) and pass
        the barrier: afterwards every rank's shard of this aggregation can be pulled.  -> (operand, padded width).
        With operand = "bf16" (and a width that is a multiple of 8) the buffer holds bfloat16 rows.  An x_local that
        already IS the own-rows segment of one of the two buffers (own_rows) is used where it lies."""
        dim = x_local.shape[1]
        b, dpad, b16 = self._peer_bufs(dim)
        h = self.halo
        in_place = None
        if not b16 and x_local.dtype == torch.float32 and x_local.stride(0) == dpad:
            for k_, slot in enumerate(b["slots"]):
                if x_local.data_ptr() == slot[0].data_ptr() + b["own0"] * dpad * 4:
                    in_place = k_
        if in_place is not None:
            b["turn"] = in_place
        cat, self._peer_tab, self._peer_segs = b["slots"][b["turn"]]
        b["turn"] ^= 1
        own = cat[b["own0"]: b["own0"] + self.n_local]
        if in_place is not None:
            pass                                                 # the producer wrote the rows where they are read
        elif b16:
            import HCSPMM
            HCSPMM.f32_to_bf16_into(x_local, own)                # one rounding per row, on its owner
        else:
            own[:, :dim].copy_(x_local)                          # pad columns stay zero
        if self.push:                                            # owner writes its rows into the peers' operands
            from . import peer as _peer
            src = own.view(torch.float32) if b16 else own
            mask = ((1 << self.world) - 1) & ~(1 << self.rank)
            _peer.halo_push(src, h["send_row"], h["send_seg"], self._peer_tab, src.stride(0), self.world, mask, self.rank + 1)
        self.peer.barrier()
        return cat, dpad

    def _pull_halo(self, cat, dpad, col0=0, width=None, mask=None):
        if self.push:                                            # the owners already wrote every row (before the barrier)
            return
        h = self.halo                                            # own rows are already in place: own bit clear
        mask = (((1 << self.world) - 1) & ~(1 << self.rank)) if mask is None else mask
        if cat.dtype == torch.bfloat16:                          # the pull moves bytes: count a row in float units
            assert col0 % 2 == 0 and (width is None or width % 2 == 0)
            raw = cat.view(torch.float32)
            self._pull(self._peer_tab, dpad // 2, h["src_row"], h["seg"], self.world, raw, col0 // 2,
                       None if width is None else width // 2, mask, self.rank + 1)
            return
        self._pull(self._peer_tab, dpad, h["src_row"], h["seg"], self.world, cat, col0, width, mask, self.rank + 1)

    def _spmm_segments(self, cat: torch.Tensor) -> torch.Tensor:
        """Segment mode: Y_r from the local operand `cat` (own + pulled rows) and the peers' operands of the same
        exchange, read in place (the operand returned by the latest _peer_stage of this width)."""
        import HCSPMM
        dpad = cat.shape[1]
        y = torch.empty(self.n_local, dpad, device=cat.device)
        wmax = 256                                                   # the kernel's widest row: column blocks beyond
        esz = cat.element_size()
        for c0 in range(0, dpad, wmax):
            c1 = min(dpad, c0 + wmax)
            segs = [0] + [p_ + c0 * esz for p_ in self._peer_segs[1:]]
            HCSPMM.spmm_segments(cat[:, c0:c1], self.rowptr, self.colidx_seg, segs, self.direct["x_rows"], y[:, c0:c1],
                                 False, self.pre[4], self.pre[5])
        return y

    def _spmm_block(self, blk, cat: torch.Tensor, y: torch.Tensor):
        """Rows [r0, r1) of the shard on the operand `cat` -> y[r0:r1]."""
        out = y[blk["r0"]:blk["r1"]]
        if self.direct is None:
            self._spmm(cat, blk["rowptr"], blk["colidx"], blk["pre"], out=out)
            return
        import HCSPMM
        esz = cat.element_size()
        for c0 in range(0, cat.shape[1], 256):
            c1 = min(cat.shape[1], c0 + 256)
            segs = [0] + [p_ + c0 * esz for p_ in self._peer_segs[1:]]
            HCSPMM.spmm_segments(cat[:, c0:c1], blk["rowptr"], blk["colidx_seg"], segs, self.direct["x_rows"], out[:, c0:c1],
                                 False, blk["pre"][4], blk["pre"][5])

    def _aggregate_row_blocks(self, cat: torch.Tensor, dpad: int, dim: int) -> torch.Tensor:
        """The row-block pipeline: the halo travels in the order the row blocks need it (communication stream; the
        first part on the full grid, the later parts on few CTAs because they share the SMs with the SpMMs), block b's
        SpMM starts on its own stream when its part has landed.  The blocks write disjoint rows of Y."""
        from . import capi
        dev = cat.device
        cur, comm = torch.cuda.current_stream(dev), self._hi_stream
        y = torch.empty(self.n_local, dpad, device=dev)
        staged = torch.cuda.Event()
        staged.record(cur)                                   # own rows written, barrier passed (and y allocated)
        raw = cat.view(torch.float32) if cat.dtype == torch.bfloat16 else cat
        lds = raw.shape[1]
        landed = []
        with torch.cuda.stream(comm):
            comm.wait_event(staged)
            for b, blk in enumerate(self.blocks["list"]):
                if blk["rows"] > 0:
                    old = capi.set_tuning("pull_ctas", self.overlap_ctas) if b > 0 else None
                    mask = ((1 << self.world) - 1) & ~(1 << self.rank)
                    self._pull(self._peer_tab, lds, blk["src_row"], blk["seg"], self.world, raw, 0, lds, mask, self.rank + 1,
                               dst_row=blk["dst_row"])
                    if old is not None:
                        capi.set_tuning("pull_ctas", old)
                ev = torch.cuda.Event()
                ev.record(comm)
                landed.append(ev)
        for blk, ev, st in zip(self.blocks["list"], landed, self.blocks["streams"]):
            with torch.cuda.stream(st):
                st.wait_event(staged)
                st.wait_event(ev)
                self._spmm_block(blk, cat, y)
            cur.wait_stream(st)
        y.record_stream(cur)
        return y if dpad == dim else y[:, :dim]

    def local_spmm(self, operand: torch.Tensor) -> torch.Tensor:
        """The compute step alone, on the operand the latest `exchange` returned (phase timing)."""
        if self.blocks is not None:                         # the blocks back to back on the current stream
            y = torch.empty(self.n_local, operand.shape[1], device=operand.device)
            for blk in self.blocks["list"]:
                self._spmm_block(blk, operand, y)
            return y
        if self.direct is not None:
            return self._spmm_segments(operand)
        return self._spmm(operand, self.rowptr, self.colidx, self.pre)

    def _aggregate_peer(self, x_local: torch.Tensor) -> torch.Tensor:
        dim, dev = x_local.shape[1], x_local.device
        cat, dpad = self._peer_stage(x_local if x_local.dtype == torch.float32 else x_local.float())
        n_slabs = self.n_slabs if (dpad >= 32 * self.n_slabs and cat.dtype == torch.float32 and not self.push) else 1
        if self.passes is not None and not self.push:
            p0, p1 = self.passes
            y = torch.empty(self.n_local, dpad, device=dev)
            cur, comm = torch.cuda.current_stream(dev), self._hi_stream
            self._pull_halo(cat, dpad, mask=p0["mask"])
            ev0 = torch.cuda.Event()
            ev0.record(cur)
            with torch.cuda.stream(comm):
                comm.wait_event(ev0)                              # first half has landed
                from . import capi
                old = capi.set_tuning("pull_ctas", self.overlap_ctas)   # NVLink-bound: leave the SMs to pass 0
                self._pull_halo(cat, dpad, mask=p1["mask"])
                capi.set_tuning("pull_ctas", old)
                ev1 = torch.cuda.Event()
                ev1.record(comm)
            self._spmm(cat, p0["rowptr"], p0["colidx"], p0["pre"], out=y)
            cur.wait_event(ev1)
            self._spmm(cat, p1["rowptr"], p1["colidx"], p1["pre"], out=y, accumulate=True)
            return y if dpad == dim else y[:, :dim]
        if self.blocks is not None:
            return self._aggregate_row_blocks(cat, dpad, dim)
        if self.direct is not None:
            # pulled rows first (they are gathered many times), then ONE kernel that sums local operand rows and rows
            # read in place from the peers' operands over NVLink
            self._pull_halo(cat, dpad)
            y = self._spmm_segments(cat)
            return y if dpad == dim else y[:, :dim]
        if n_slabs == 1:
            self._pull_halo(cat, dpad)
            y = self._spmm(cat, self.rowptr, self.colidx, self.pre)
            return y if dpad == dim else y[:, :dim]
        step = (dpad // n_slabs + 31) // 32 * 32
        edges = list(range(0, dpad, step)) + [dpad]
        y = torch.empty(self.n_local, dpad, device=dev)
        cur, comm = torch.cuda.current_stream(dev), self._hi_stream
        comm.wait_stream(cur)
        events = []
        from . import capi
        with torch.cuda.stream(comm):
            for k in range(len(edges) - 1):
                # slab 0 has nothing to hide behind: full grid.  Later slabs travel while the SpMM of the previous slab
                # runs: NVLink-bound, a few CTAs keep the links busy and leave the SMs to the SpMM
                old = capi.set_tuning("pull_ctas", self.overlap_ctas) if k > 0 else None
                self._pull_halo(cat, dpad, edges[k], edges[k + 1] - edges[k])
                if old is not None:
                    capi.set_tuning("pull_ctas", old)
                ev = torch.cuda.Event()
                ev.record(comm)
                events.append(ev)
        for k, ev in enumerate(events):
            cur.wait_event(ev)
            self._spmm(cat[:, edges[k]:edges[k + 1]], self.rowptr, self.colidx, self.pre, out=y[:, edges[k]:edges[k + 1]])
        comm.wait_stream(cur)
        return y if dpad == dim else y[:, :dim]

    def _aggregate_halo(self, x_local: torch.Tensor) -> torch.Tensor:
        h = self.halo
        dim, dev, dt = x_local.shape[1], x_local.device, x_local.dtype
        pipelined = self.n_slabs > 1 and dim >= 8 * self.n_slabs and self._comm_stream is not None
        if not pipelined:
            return self._spmm(self.exchange(x_local), self.rowptr, self.colidx, self.pre)
        step = (dim // self.n_slabs + 7) // 8 * 8
        edges = list(range(0, dim, step)) + [dim]
        y = torch.empty(self.n_local, dim, device=dev, dtype=dt)
        cur, comm = torch.cuda.current_stream(dev), self._comm_stream
        comm.wait_stream(cur)
        events, cats = [], []
        for k in range(len(edges) - 1):
            w = edges[k + 1] - edges[k]
            key = ("halo_slab", k, w, dev, dt)
            if key not in self._bufs:
                self._bufs[key] = torch.empty(h["rows"], w, device=dev, dtype=dt)
            cat = self._bufs[key]
            with torch.cuda.stream(comm):
                packed = x_local[:, edges[k]:edges[k + 1]].index_select(0, h["send_idx"])
                dist.all_to_all_single(cat, packed, output_split_sizes=h["recv"], input_split_sizes=h["send"],
                                       group=self.group)
                packed.record_stream(comm)
                ev = torch.cuda.Event()
                ev.record(comm)
            events.append(ev)
            cats.append(cat)
        for k, (ev, cat) in enumerate(zip(events, cats)):
            cur.wait_event(ev)
            self._spmm(cat, self.rowptr, self.colidx, self.pre, out=y[:, edges[k]:edges[k + 1]])
        comm.wait_stream(cur)
        return y

    def _aggregate_slabs(self, x_local: torch.Tensor) -> torch.Tensor:
        dim = x_local.shape[1]
        step = (dim // self.n_slabs + 3) // 4 * 4
        edges = list(range(0, dim, step)) + [dim]
        dev, dt = x_local.device, x_local.dtype
        y = torch.empty(self.n_local, dim, device=dev, dtype=dt)
        cur = torch.cuda.current_stream(dev)
        comm = self._comm_stream
        comm.wait_stream(cur)                       # x_local is produced on the current stream
        events, gathered = [], []
        for k in range(len(edges) - 1):
            w = edges[k + 1] - edges[k]
            key = ("slab", k, w, dev, dt)
            if key not in self._bufs:
                self._bufs[key] = (torch.zeros(self.max_rows, w, device=dev, dtype=dt),
                                   torch.empty(self.world * self.max_rows, w, device=dev, dtype=dt))
            pad, gat = self._bufs[key]
            with torch.cuda.stream(comm):
                pad[: self.n_local].copy_(x_local[:, edges[k]:edges[k + 1]])
                dist.all_gather_into_tensor(gat, pad, group=self.group)
                ev = torch.cuda.Event()
                ev.record(comm)
            events.append(ev)
            gathered.append(gat)
        for k, (ev, gat) in enumerate(zip(events, gathered)):
            cur.wait_event(ev)                      # slab k has landed; slab k+1.. still in flight
            self._spmm(gat, self.rowptr, self.colidx, self.pre, out=y[:, edges[k]:edges[k + 1]])
        comm.wait_stream(cur)                       # buffers are reused by the next call
        return y


class ShardedAggregate(torch.autograd.Function):
    """Autograd wrapper: forward Y_r = A_r X; backward dX_r = A_r dY (A symmetric, as the reference
    assumes -- GNN_model.py:77) or (A^T)_r dY when a transposed ShardedGraph is supplied."""

    @staticmethod
    def forward(ctx, x_local, graph: ShardedGraph, graph_t):
        ctx.graph_t = graph_t if graph_t is not None else graph
        return graph.aggregate(x_local)

    @staticmethod
    def backward(ctx, d_y):
        return ctx.graph_t.aggregate(d_y.contiguous()), None, None


class ShardedGCNLayer(torch.autograd.Function):
    """X' = A (X W) on a row-partitioned graph -- the routing of the reference's GCN Functions
    (GNN_model.py:61-162; hcspmm/gnn.py "pre"): forward = Update GEMM, then the exchanged aggregation; backward =
    the FUSED entry point (dX, d(XW)) = ((A^T dY) W^T, A^T dY) on the exchanged dY, then dW = X^T d(XW)."""

    @staticmethod
    def forward(ctx, x_local, weights, graph: ShardedGraph, graph_t):
        ctx.graph_t = graph_t if graph_t is not None else graph
        ctx.save_for_backward(x_local, weights)
        # the Update product is written straight into the rank's rows of the exchange operand (no staging copy)
        return graph.aggregate(graph.update(x_local, weights, out=graph.own_rows(weights.shape[1])))

    @staticmethod
    def backward(ctx, d_y):
        x_local, weights = ctx.saved_tensors
        d_x, d_xw = ctx.graph_t.aggregate_fused(d_y.contiguous(), weights.t())
        return d_x, torch.mm(x_local.t(), d_xw), None, None


class ShardedGINLayer(torch.autograd.Function):
    """X' = (A X) W (GNN_model.py:166-232; hcspmm/gnn.py "post"): forward = the fused entry point on the exchanged
    X; backward = dAgg = dY W^T (Update GEMM), dW = Agg^T dY, dX = A^T dAgg (exchanged aggregation)."""

    @staticmethod
    def forward(ctx, x_local, weights, graph: ShardedGraph, graph_t):
        ctx.graph, ctx.graph_t = graph, graph_t if graph_t is not None else graph
        out, agg = graph.aggregate_fused(x_local, weights)
        ctx.save_for_backward(agg, weights)
        return out

    @staticmethod
    def backward(ctx, d_y):
        agg, weights = ctx.saved_tensors
        d_y = d_y.contiguous()
        d_agg = ctx.graph.update(d_y, weights.t())
        return ctx.graph_t.aggregate(d_agg), torch.mm(agg.t(), d_y), None, None


class DistGCN(torch.nn.Module):
    """2+-layer GCN / GIN of HC-SpMM_main.py:67-87 on a row-partitioned graph: every rank holds its rows of X and of
    the labels and a replica of the weights; Update GEMMs are row-local (the library's TMA + tcgen05 GEMM), every
    Aggregation goes through the exchange and the HCSPMM entry points -- the fused ones where the reference uses them
    (ShardedGCNLayer / ShardedGINLayer) -- and weight gradients are summed with one all-reduce per step."""

    def __init__(self, graph: ShardedGraph, in_dim, hidden, classes, num_layers=2, seed=0, graph_t=None,
                 order: str = "auto", pad_to: int = 8):
        """order: "update_first" = A (H W) (GCN), "aggregate_first" = (A H) W (GIN), "auto" = the cheaper per layer.
        pad_to: feature / class widths are laid out as multiples of this (zero columns): 100 input features travel
        and aggregate as 104, 47 classes as 48, so every aggregation takes the 256-bit gather path and every Update
        GEMM the TMA path without per-call pad / unpad passes.  The padding weights start at zero, receive zero
        gradients (the padded logits are sliced away before the loss) and so stay zero: the model is unchanged."""
        super().__init__()
        self.order = order
        g = torch.Generator().manual_seed(seed)            # identical replicas on every rank
        dims = [in_dim] + [hidden] * (num_layers - 1) + [classes]
        pad = lambda d: (d + pad_to - 1) // pad_to * pad_to if pad_to > 1 else d
        self.in_dim, self.classes = in_dim, classes
        ws = []
        for i in range(num_layers):
            w = torch.randn(dims[i], dims[i + 1], generator=g) / dims[i] ** 0.5
            wp = torch.zeros(pad(dims[i]), pad(dims[i + 1]))
            wp[:dims[i], :dims[i + 1]] = w
            ws.append(torch.nn.Parameter(wp))
        self.weights = torch.nn.ParameterList(ws)
        self.graph, self.graph_t = graph, graph_t

    def pad_features(self, x_local: torch.Tensor) -> torch.Tensor:
        """Lay the input features out at the padded width once (a data-layout step, outside the epoch)."""
        wp = self.weights[0].shape[0]
        if x_local.shape[1] == wp:
            return x_local
        out = torch.zeros(x_local.shape[0], wp, device=x_local.device, dtype=x_local.dtype)
        out[:, :x_local.shape[1]] = x_local
        return out

    def forward(self, x_local):
        h = self.pad_features(x_local)
        for i, w in enumerate(self.weights):
            # A (H W) = (A H) W: exchange and aggregate at the narrower of the two widths -- the exchange moves
            # rows * width * 4 bytes per rank and the SpMM gathers nnz * width * 4
            update_first = {"update_first": True, "aggregate_first": False}.get(self.order, w.shape[1] <= w.shape[0])
            layer = ShardedGCNLayer if update_first else ShardedGINLayer
            h = layer.apply(h, w, self.graph, self.graph_t)
            if i + 1 < len(self.weights):
                h = torch.relu(h)
        return torch.nn.functional.log_softmax(h[:, :self.classes], dim=1)

    def loss(self, x_local, y_local):
        """nll_loss averaged over ALL vertices (sum over local rows / N) -- as a gather + sum: torch's nll_loss with
        reduction="sum" reduces 2.4 M rows in one block (2.6 ms forward + 1.4 ms backward at the products shape)."""
        logp = self.forward(x_local)
        return -logp.gather(1, y_local.view(-1, 1)).sum() / self.graph.n

    def sync_grads(self):
        if self.graph.world > 1:
            flat = torch.cat([p.grad.reshape(-1) for p in self.parameters()])
            dist.all_reduce(flat, group=self.graph.group)
            o = 0
            for p in self.parameters():
                p.grad.copy_(flat[o:o + p.numel()].view_as(p))
                o += p.numel()
