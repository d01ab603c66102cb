// spmm.cu -- the hybrid SpMM kernel  Y = A * X  (binary CSR A, FP32 X/Y) for sm_100a.
//
// Replaces the nine spmm_forward_cuda_kernel_arbi_warps_hybrid_* kernels of the reference
// (/root/reference/hybrid_kernel/hybrid_all_kernel.cu:919-2770) for the plain-SpMM part:
// one launch, one CTA per 16-row window, the window's hybrid_type picks the path --
//
//  CUDA-core path (hybrid_type == 0; reference :960-1037, :1371-1382)
//    HBM/L2-gather bound.  A warp walks a row: 32 column ids are fetched with one
//    coalesced load and broadcast by shuffle; each edge's X row is read with 128-bit
//    loads by a group of LPE lanes (NV float4 per lane), 8/NV*... independent loads in
//    flight per lane; FP32 register accumulators; groups are reduced by shuffles.  Rows
//    with >= long_row non-zeros (power-law hubs) are split over all eight warps of the
//    CTA and reduced through shared memory in a fixed order (deterministic, no atomics).
//
//  tensor-core path (hybrid_type != 0; reference :1039-1121)
//    The window is condensed to its U distinct columns (edgeToColumn).  The CTA rebuilds,
//    in shared memory, the column list and one 16-bit row mask per condensed column from
//    edgeToColumn / edgeToRow / edgeList -- the same three arrays the reference scatters
//    into its sparse_A tile (:1072-1079) -- so each distinct X row is gathered ONCE per
//    window (cp.async 16-byte copies into a padded, bank-conflict-free tile, double
//    buffered) and multiplied on the tensor cores: mma.sync.m16n8k8 TF32, A fragments
//    built straight from the bit masks (no A tile in memory), FP32 accumulate.  Unlike the
//    reference there is no cap on U (MAX_BLK, :23) or on dim (:1098).
//
// Feature slabs: blockIdx.y selects a slab of `slab` features.  CTAs are scheduled
// x-fastest, so all windows of one slab run before the next slab starts and the
// N x slab slice of X they gather from can stay L2-resident.
#include "common.cuh"

namespace hcspmm {

#ifndef HCSPMM_MIN_CTAS
#define HCSPMM_MIN_CTAS 2
#endif
constexpr int KC = 16;      // condensed columns staged per pipeline step (two k=8 MMA steps)
#ifndef HCSPMM_TC_STAGES
#define HCSPMM_TC_STAGES 4
#endif
constexpr int NSTAGE = HCSPMM_TC_STAGES;  // cp.async pipeline depth of the tensor-core path
constexpr int UCAP = 1024;  // condensed columns whose (col, mask) are resident at once

struct SpmmParams {
  const float *x;
  long long ldx;
  int x_rows;
  const int *rowptr, *colidx, *bp, *etc, *etr, *ht;
  int n_rows, dim, slab;
  int precision, accumulate;
  float *y;
  long long ldy;
  int long_row;   // rows with >= long_row entries: all warps of the CTA share the row
  int short_row;  // rows with <  short_row entries: one lane GROUP per row (G rows per warp at once)
  int wpc;        // 16-row windows per CTA (1..MAX_WPC), > 1 on low-degree graphs
  int n_windows;
  int cuda_elsewhere;  // 1: CUDA-core windows (label 0) are computed by spmm_balanced_kernel, skip them here
};

// Work-balanced CUDA-core kernel (below): merge-path items over (rows + stored entries)
struct BalParams {
  SpmmParams s;
  long long nnz;
  int chunk;         // rows + entries per item (one CTA per item and feature slab)
  int n_items;
  float *partial;    // [n_items][2][dim] FP32 partial sums of rows that straddle item boundaries
  int *split_row;    // [n_items][2]: row whose sum ENDS in this item but began earlier (slot 0) /
                     //               row whose sum continues in the next item (slot 1); -1 = none
  const int *splits; // rows consumed before diagonal i * (chunk / split_stride), i = 0 .. split_max
                     // (merge_path_splits_kernel; precomputed once per graph at a base chunk when the caller
                     // passes hcspmm_aux_t.d_splits -- item k then reads entries k * split_stride, clamped)
  int split_stride, split_max;
  int low_degree;    // mean row length < 64: launch the three-CTAs-per-SM build
  int warp_split;    // > 0: items with a mean row length >= warp_split give every warp an equal run of entries;
                     // other items (and 0) use the warp-per-row / CTA-per-long-row phases
  int hint_cls_min;  // >= 0: s.colidx is TAGGED (hcspmm_tag_columns); rows of class >= this are loaded evict_last,
                     // the others evict_first.  -1: plain column ids, every row evict_last
  const int *row_id;      // row-sorted CSR (csrc/rowsort.cu): s.rowptr / s.colidx are the sorted copy and sorted row i is
                          // row row_id[i] of Y (and of the window labels); nullptr = rows in place
  int seg_mode;           // 1: s.colidx carries segment tags, seg_x[] is valid
  const float *seg_x[8];  // segment mode (hcspmm_aux_t.d_colidx_segments): base of the X segment a column id's bits
                          // 29..31 name; seg_x[0] = s.x
};
constexpr int MAX_WPC = 8;

// ---------------------------------------------------------------------------------------
// Per-lane feature vector: VW = 4 floats (LDG.128) or 8 floats (LDG.256, new on sm_100).
// The 256-bit form also carries an L2 eviction priority: X rows are loaded evict_last so the
// streamed column ids / Y writes do not push the gathered matrix out of the 126 MB L2.
// ---------------------------------------------------------------------------------------
template <int VW, bool B16 = false> struct Vec;
template <> struct Vec<4, false> {
  float4 a;
  __device__ __forceinline__ void zero() { a = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void load(const float *p) { a = ldg_f4(p); }
  __device__ __forceinline__ void load_hint(const float *p, uint64_t pol) {
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p), "l"(pol));
  }
  __device__ __forceinline__ void add(const Vec &o) { add4(a, o.a); }
  __device__ __forceinline__ void xor_reduce(int off) {
    a.x += __shfl_xor_sync(0xffffffffu, a.x, off);
    a.y += __shfl_xor_sync(0xffffffffu, a.y, off);
    a.z += __shfl_xor_sync(0xffffffffu, a.z, off);
    a.w += __shfl_xor_sync(0xffffffffu, a.w, off);
  }
  __device__ __forceinline__ void load_plain(const float *p) { a = *reinterpret_cast<const float4 *>(p); }
  __device__ __forceinline__ void store(float *p) const { *reinterpret_cast<float4 *>(p) = a; }
};
template <> struct Vec<8, false> {
  float4 a, b;
  __device__ __forceinline__ void zero() { a = b = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void load(const float *p) {
    asm volatile("ld.global.nc.L2::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
  }
  // the same 256-bit load with a per-load L2 eviction policy (createpolicy): hot rows evict_last, cold rows evict_first
  __device__ __forceinline__ void load_hint(const float *p, uint64_t pol) {
    asm volatile("ld.global.nc.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p), "l"(pol));
  }
  __device__ __forceinline__ void add(const Vec &o) { add4(a, o.a); add4(b, o.b); }
  __device__ __forceinline__ void xor_reduce(int off) {
    a.x += __shfl_xor_sync(0xffffffffu, a.x, off);
    a.y += __shfl_xor_sync(0xffffffffu, a.y, off);
    a.z += __shfl_xor_sync(0xffffffffu, a.z, off);
    a.w += __shfl_xor_sync(0xffffffffu, a.w, off);
    b.x += __shfl_xor_sync(0xffffffffu, b.x, off);
    b.y += __shfl_xor_sync(0xffffffffu, b.y, off);
    b.z += __shfl_xor_sync(0xffffffffu, b.z, off);
    b.w += __shfl_xor_sync(0xffffffffu, b.w, off);
  }
  __device__ __forceinline__ void load_plain(const float *p) {
    a = *reinterpret_cast<const float4 *>(p);
    b = *reinterpret_cast<const float4 *>(p + 4);
  }
  __device__ __forceinline__ void store(float *p) const {
    *reinterpret_cast<float4 *>(p) = a;
    *reinterpret_cast<float4 *>(p + 4) = b;
  }
};

// BF16-stored X (precision = BF16): 8 features arrive as one 128-bit load of eight bfloat16 and are
// widened to FP32 (a shift / mask each) before the FP32 accumulation.  Pointers and offsets into X
// are kept in float units, i.e. halved (XDIV = 2).
template <> struct Vec<8, true> {
  float4 a, b;
  __device__ __forceinline__ void zero() { a = b = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void widen(const float4 raw) {
    const uint32_t w0 = __float_as_uint(raw.x), w1 = __float_as_uint(raw.y), w2 = __float_as_uint(raw.z),
                   w3 = __float_as_uint(raw.w);
    a = make_float4(__uint_as_float(w0 << 16), __uint_as_float(w0 & 0xffff0000u), __uint_as_float(w1 << 16),
                    __uint_as_float(w1 & 0xffff0000u));
    b = make_float4(__uint_as_float(w2 << 16), __uint_as_float(w2 & 0xffff0000u), __uint_as_float(w3 << 16),
                    __uint_as_float(w3 & 0xffff0000u));
  }
  __device__ __forceinline__ void load(const float *p) { widen(ldg_f4(p)); }
  __device__ __forceinline__ void load_hint(const float *p, uint64_t pol) {
    float4 raw;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(raw.x), "=f"(raw.y), "=f"(raw.z), "=f"(raw.w) : "l"(p), "l"(pol));
    widen(raw);
  }
  __device__ __forceinline__ void add(const Vec &o) { add4(a, o.a); add4(b, o.b); }
  __device__ __forceinline__ void xor_reduce(int off) {
    a.x += __shfl_xor_sync(0xffffffffu, a.x, off);
    a.y += __shfl_xor_sync(0xffffffffu, a.y, off);
    a.z += __shfl_xor_sync(0xffffffffu, a.z, off);
    a.w += __shfl_xor_sync(0xffffffffu, a.w, off);
    b.x += __shfl_xor_sync(0xffffffffu, b.x, off);
    b.y += __shfl_xor_sync(0xffffffffu, b.y, off);
    b.z += __shfl_xor_sync(0xffffffffu, b.z, off);
    b.w += __shfl_xor_sync(0xffffffffu, b.w, off);
  }
  __device__ __forceinline__ void load_plain(const float *p) {   // FP32 side (Y)
    a = *reinterpret_cast<const float4 *>(p);
    b = *reinterpret_cast<const float4 *>(p + 4);
  }
  __device__ __forceinline__ void store(float *p) const {
    *reinterpret_cast<float4 *>(p) = a;
    *reinterpret_cast<float4 *>(p + 4) = b;
  }
};

// Per-load L2 residency control of the balanced kernel.  The gather of a graph whose X does not fit L2 is served by
// whatever L2 happens to keep; with every row loaded evict_last (round 1) nothing discriminates, and the products
// shape hits 51 % although 5 % of its rows take 75 % of the references.  Here column ids arrive TAGGED with the
// hotness class of their column (3 bits above bit 28, hcspmm_tag_columns: class by rank in descending reference
// count) and each row is loaded with the policy of its class: rows among the hottest `budget / row bytes` are
// evict_last, all others evict_first -- same sums, same order, only the cache hints differ.
struct GatherHint {
  unsigned mask;      // column id = tagged & mask
  unsigned cls_min;   // hot iff (tagged >> 29) >= cls_min
  uint64_t pol_hot, pol_cold;
  const float *const *seg;   // MODE 2: bases of the X segments (shared memory), indexed by tagged >> 29
  int lane_off;              // MODE 2: this lane's float offset inside a row (feature slab + lane vector)
};
__device__ __forceinline__ GatherHint no_hint() { return GatherHint{0xffffffffu, 0u, 0ull, 0ull, nullptr, 0}; }
__device__ __forceinline__ GatherHint make_hint(int cls_min) {
  GatherHint h;
  h.mask = 0x1fffffffu;
  h.cls_min = (unsigned)cls_min;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(h.pol_hot));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(h.pol_cold));
  h.seg = nullptr;
  h.lane_off = 0;
  return h;
}
// MODE 2 -- X in SEGMENTS: bits 29..31 of a column id select the buffer its row lives in.  Segment 0 is the rank's own
// exchange operand; segments 1..7 are the operands of its PEERS, mapped over NVLink (CUDA IPC): rows of X that the
// shard references only once or twice are not copied into the local operand first (halo pull) but read in place by
// the gather itself -- the transfer of those rows overlaps the sums, and they are neither written to nor re-read
// from local memory.  Same loads, same sum order: only the address of a row depends on its segment.
__device__ __forceinline__ GatherHint make_segments(const float *const *seg, int lane_off) {
  return GatherHint{0x1fffffffu, 0u, 0ull, 0ull, seg, lane_off};
}
// the X row of a (possibly tagged) column id, at this lane's offset
template <int MODE>
__device__ __forceinline__ const float *row_ptr(const float *__restrict__ xlane, long long ldx, unsigned ct,
                                                const GatherHint &h) {
  if constexpr (MODE == 2) return h.seg[ct >> 29] + h.lane_off + (long long)(ct & h.mask) * ldx;
  else if constexpr (MODE == 1) return xlane + (long long)(ct & h.mask) * ldx;
  else return xlane + (long long)(int)ct * ldx;
}

#ifndef HCSPMM_INFLIGHT_BYTES
#define HCSPMM_INFLIGHT_BYTES 128  // gathered bytes kept in flight per lane (ring depth x vector bytes)
#endif

// ---------------------------------------------------------------------------------------
// CUDA-core gather: accumulate X rows of edges [eb, ee), visiting 32-edge chunks
// chunk0, chunk0 + chunk_stride, ...   A group of LPE lanes reads one X row, lane g of the group
// owning vectors g, g + LPE, ... (NV of them, VW floats each).
// ---------------------------------------------------------------------------------------
template <int LPE, int NV, int VW, bool B16, int MODE = 0>
__device__ __forceinline__ void gather_accumulate(Vec<VW, B16> (&acc)[NV], const float *__restrict__ xlane,
                                                  long long ldx, int x_rows,
                                                  const int *__restrict__ colidx, int eb, int ee,
                                                  int chunk0, int chunk_stride, int lane, int q,
                                                  const bool (&active)[NV], bool all_active,
                                                  const GatherHint h = no_hint()) {
  constexpr int G = 32 / LPE;         // edges handled concurrently by one warp
  constexpr int STEPS = 32 / G;       // gather steps per full 32-edge chunk
  constexpr int XDIV = B16 ? 2 : 1;      // X offsets in float units: a BF16 row is half as long
  constexpr int U0 = HCSPMM_INFLIGHT_BYTES * XDIV / (NV * VW * 4);
  constexpr int U = U0 < 1 ? 1 : (U0 > STEPS ? STEPS : U0);   // ring depth
  int voff[NV];   // float offset of vector i from xlane; inactive lanes point at the slab's first vector
#pragma unroll
  for (int i = 0; i < NV; ++i) voff[i] = (active[i] ? i * LPE * VW : -(lane % LPE) * VW) / XDIV;
  int base = eb + chunk0 * 32;
  int c_next = (base + lane < ee) ? __ldg(colidx + base + lane) : -1;
  for (; base < ee; base += chunk_stride * 32) {
    const int n = min(32, ee - base);
    const int c = c_next;
    const int nb = base + chunk_stride * 32;
    c_next = (nb + lane < ee) ? __ldg(colidx + nb + lane) : -1;  // next chunk's ids, early
    (void)all_active;
    const bool fast = n == 32 && __all_sync(0xffffffffu, ((unsigned)c & ((MODE == 1 || MODE == 2) ? h.mask : 0xffffffffu)) < (unsigned)x_rows);
    if (fast) {
      // full chunk, every id valid: unpredicated ring of U loads in flight -- slot s % U is consumed
      // and immediately refilled with step s + U.  Lanes beyond the slab width (voff < 0) re-read
      // the row's first vector (same cache line as lane 0) and never store their sums.
      Vec<VW, B16> v[U][NV];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned ct = (unsigned)__shfl_sync(0xffffffffu, c, u * G + q);
        const float *src = row_ptr<MODE>(xlane, ldx, ct, h);
        if constexpr (MODE == 1) {
          const uint64_t pol = (ct >> 29) >= h.cls_min ? h.pol_hot : h.pol_cold;
#pragma unroll
          for (int i = 0; i < NV; ++i) v[u][i].load_hint(src + voff[i], pol);
        } else {
#pragma unroll
          for (int i = 0; i < NV; ++i) v[u][i].load(src + voff[i]);
        }
      }
#pragma unroll
      for (int s = 0; s < STEPS; ++s) {
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i].add(v[s % U][i]);
        if (s + U < STEPS) {
          const unsigned ct = (unsigned)__shfl_sync(0xffffffffu, c, (s + U) * G + q);
          const float *src = row_ptr<MODE>(xlane, ldx, ct, h);
          if constexpr (MODE == 1) {
            const uint64_t pol = (ct >> 29) >= h.cls_min ? h.pol_hot : h.pol_cold;
#pragma unroll
            for (int i = 0; i < NV; ++i) v[s % U][i].load_hint(src + voff[i], pol);
          } else {
#pragma unroll
            for (int i = 0; i < NV; ++i) v[s % U][i].load(src + voff[i]);
          }
        }
      }
    } else {
#pragma unroll 1
      for (int t = 0; t * G < n; t += U) {
        Vec<VW, B16> v[U][NV];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = (t + u) * G + q;
          const unsigned ct = (unsigned)__shfl_sync(0xffffffffu, c, j & 31);
          const unsigned cu = (MODE == 1 || MODE == 2) ? (ct & h.mask) : ct;
          const bool ok = (j < n) && (cu < (unsigned)x_rows);
          const float *src = row_ptr<MODE>(xlane, ldx, ct, h);
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            if (ok && active[i]) {
              if constexpr (MODE == 1) v[u][i].load_hint(src + i * LPE * VW / XDIV, (ct >> 29) >= h.cls_min ? h.pol_hot : h.pol_cold);
              else v[u][i].load(src + i * LPE * VW / XDIV);
            } else v[u][i].zero();
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int i = 0; i < NV; ++i) acc[i].add(v[u][i]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Low-degree rows: one group of LPE lanes owns a whole row, so a warp advances G = 32 / LPE rows at
// once and nothing is reduced across lanes.  The group fetches LPE column ids with one load and
// broadcasts them inside the group (sub-warp shuffle masks: groups may run different trip counts).
// ---------------------------------------------------------------------------------------
template <int LPE, int NV, int VW, bool B16, int MODE = 0>
__device__ __forceinline__ void gather_group_row(Vec<VW, B16> (&acc)[NV], const float *__restrict__ xlane,
                                                 long long ldx, int x_rows,
                                                 const int *__restrict__ colidx, int eb, int ee, int lane,
                                                 int q, int g, const bool (&active)[NV],
                                                 const GatherHint h = no_hint()) {
  constexpr int XDIV = B16 ? 2 : 1;
  constexpr int UB0 = HCSPMM_INFLIGHT_BYTES * XDIV / (NV * VW * 4);
  constexpr int UB = UB0 < 1 ? 1 : (UB0 > LPE ? LPE : UB0);
  const unsigned gmask = LPE == 32 ? 0xffffffffu : (((1u << LPE) - 1u) << (q * LPE));
  (void)lane;
  for (int e = eb; e < ee; e += LPE) {
    const int my = (e + g < ee) ? __ldg(colidx + e + g) : -1;
    const int cnt = min(LPE, ee - e);
#pragma unroll 1
    for (int j0 = 0; j0 < cnt; j0 += UB) {
      Vec<VW, B16> v[UB][NV];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const unsigned ct = (unsigned)__shfl_sync(gmask, my, q * LPE + ((j0 + u) & (LPE - 1)));
        const unsigned cu = (MODE == 1 || MODE == 2) ? (ct & h.mask) : ct;
        const bool ok = (j0 + u < cnt) && (cu < (unsigned)x_rows);
        const float *src = row_ptr<MODE>(xlane, ldx, ct, h);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if (ok && active[i]) {
            if constexpr (MODE == 1) v[u][i].load_hint(src + i * LPE * VW / XDIV, (ct >> 29) >= h.cls_min ? h.pol_hot : h.pol_cold);
            else v[u][i].load(src + i * LPE * VW / XDIV);
          } else v[u][i].zero();
        }
      }
#pragma unroll
      for (int u = 0; u < UB; ++u)
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i].add(v[u][i]);
    }
  }
}

template <int LPE, int NV, int VW, bool B16>
__device__ __forceinline__ void group_reduce(Vec<VW, B16> (&acc)[NV]) {
#pragma unroll
  for (int o = LPE; o < 32; o <<= 1)
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i].xor_reduce(o);
}

// ---------------------------------------------------------------------------------------
// Tensor-core window
// ---------------------------------------------------------------------------------------
template <int MAXNT>
__device__ __forceinline__ void tc_window(const SpmmParams &p, int w, int e0, int e1, int feat0,
                                          int S, float *smem) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const int stride = S + 8;
  int *ucol = reinterpret_cast<int *>(smem);
  unsigned *umask = reinterpret_cast<unsigned *>(smem) + (UCAP + KC);
  float *xs = smem + 2 * (UCAP + KC);
  const int NT = S >> 3;
  const int nvec = S >> 2;
  const int upad = p.bp[w] * BLK_W;
  const float *xbase = p.x + feat0;

  float acc[MAXNT][4];
#pragma unroll
  for (int k = 0; k < MAXNT; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;

  for (int u0 = 0; u0 < upad; u0 += UCAP) {
    const int ucnt = min(UCAP, upad - u0);
    const int ucnt16 = (ucnt + KC - 1) / KC * KC;
    __syncthreads();  // previous chunk fully consumed
    for (int i = tid; i < ucnt16; i += CTA_THREADS) { ucol[i] = -1; umask[i] = 0u; }
    __syncthreads();
    for (int e = e0 + tid; e < e1; e += CTA_THREADS) {
      const int r = __ldg(p.etc + e) - u0;
      if (r >= 0 && r < ucnt) {
        ucol[r] = __ldg(p.colidx + e);
        atomicOr(&umask[r], 1u << (__ldg(p.etr + e) & (BLK_H - 1)));
      }
    }
    __syncthreads();
    const int nk = ucnt16 / KC;
    auto stage = [&](int kc, int buf) {
      float *dst = xs + buf * KC * stride;
      for (int rr = wid; rr < KC; rr += CTA_WARPS) {
        const int col = ucol[kc * KC + rr];
        const bool ok = (unsigned)col < (unsigned)p.x_rows;
        const float *src = ok ? xbase + (long long)col * p.ldx : p.x;
        for (int v = lane; v < nvec; v += 32)
          cp_async_16(dst + rr * stride + v * 4, ok ? src + v * 4 : src, ok ? 16 : 0);
      }
      cp_async_commit();
    };
    // NSTAGE-deep cp.async pipeline: NSTAGE-1 chunks of KC gathered rows are in flight while one
    // is multiplied; one barrier per chunk
#pragma unroll
    for (int s0 = 0; s0 < NSTAGE - 1; ++s0) {
      if (s0 < nk) stage(s0, s0);
      else cp_async_commit();
    }
    for (int kc = 0; kc < nk; ++kc) {
      const int buf = kc % NSTAGE;
      cp_async_wait<NSTAGE - 2>();
      __syncthreads();  // chunk kc visible to all; buffer (kc-1) % NSTAGE free for refill
      if (kc + NSTAGE - 1 < nk) stage(kc + NSTAGE - 1, (kc + NSTAGE - 1) % NSTAGE);
      else cp_async_commit();
      const float *xt = xs + buf * KC * stride;
#pragma unroll
      for (int ks = 0; ks < KC / 8; ++ks) {
        const int k0 = kc * KC + ks * 8;
        const unsigned mlo = umask[k0 + tig], mhi = umask[k0 + tig + 4];
        if (__any_sync(0xffffffffu, (mlo | mhi) != 0u)) {
          uint32_t a[4];
          a[0] = ((mlo >> g) & 1u) ? 0x3f800000u : 0u;
          a[1] = ((mlo >> (g + 8)) & 1u) ? 0x3f800000u : 0u;
          a[2] = ((mhi >> g) & 1u) ? 0x3f800000u : 0u;
          a[3] = ((mhi >> (g + 8)) & 1u) ? 0x3f800000u : 0u;
          const float *r0p = xt + (ks * 8 + tig) * stride + g;
          const float *r1p = r0p + 4 * stride;
#pragma unroll
          for (int k = 0; k < MAXNT; ++k) {
            const int nt = wid + k * CTA_WARPS;
            if (nt < NT) {
              const float x0 = r0p[nt * 8], x1 = r1p[nt * 8];
              const uint32_t b0 = f32_to_tf32(x0), b1 = f32_to_tf32(x1);
              mma_m16n8k8_tf32(acc[k], a, b0, b1);
              if (p.precision == HCSPMM_PRECISION_TF32X2) {
                const uint32_t c0 = f32_to_tf32(x0 - __uint_as_float(b0));
                const uint32_t c1 = f32_to_tf32(x1 - __uint_as_float(b1));
                mma_m16n8k8_tf32(acc[k], a, c0, c1);
              }
            }
          }
        }
      }
    }
    cp_async_wait<0>();
  }
  const int row0 = w * BLK_H + g, row1 = row0 + 8;
#pragma unroll
  for (int k = 0; k < MAXNT; ++k) {
    const int nt = wid + k * CTA_WARPS;
    if (nt < NT) {
      const int col = feat0 + nt * 8 + tig * 2;
      if (row0 < p.n_rows) {
        float2 *dst = reinterpret_cast<float2 *>(p.y + (long long)row0 * p.ldy + col);
        float2 v = make_float2(acc[k][0], acc[k][1]);
        if (p.accumulate) { float2 o = *dst; v.x += o.x; v.y += o.y; }
        *dst = v;
      }
      if (row1 < p.n_rows) {
        float2 *dst = reinterpret_cast<float2 *>(p.y + (long long)row1 * p.ldy + col);
        float2 v = make_float2(acc[k][2], acc[k][3]);
        if (p.accumulate) { float2 o = *dst; v.x += o.x; v.y += o.y; }
        *dst = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// The hybrid kernel.  Slab width S <= LPE * NV * 4 floats; X/Y 16-byte aligned, ldx/ldy % 4 == 0.
// ---------------------------------------------------------------------------------------
// MINB = minimum resident CTAs per SM the register allocation must allow: 2 for high-degree graphs
// (the gather ring wants registers), 3 for low-degree graphs (latency hiding wants warps).
template <int LPE, int NV, int VW, int MINB, bool B16 = false>
__global__ void __launch_bounds__(CTA_THREADS, MINB) spmm_hybrid_kernel(const SpmmParams p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int rp[BLK_H * MAX_WPC + 1];
  __shared__ int s_next;
  __shared__ unsigned s_tcmask, s_skipmask;
  constexpr int G = 32 / LPE;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int w0 = blockIdx.x * p.wpc;
  const int nwin = min(p.wpc, p.n_windows - w0);
  const int r0 = w0 * BLK_H;
  const int rows_here = min(BLK_H * nwin, p.n_rows - r0);
  const int feat0 = blockIdx.y * p.slab;
  const int S = min(p.slab, p.dim - feat0);
  const int nvec = S / VW;
  for (int i = tid; i <= rows_here; i += CTA_THREADS) rp[i] = __ldg(p.rowptr + r0 + i);
  if (tid == 0) {
    s_next = 0;
    // label 1: tensor-core window (mma.sync path below); label 2: the window belongs to a dense
    // super-window that the tcgen05 kernel (dense.cu) computes -- nothing to do here
    unsigned m = 0, sk = 0;
    if (p.ht != nullptr)
      for (int i = 0; i < nwin; ++i) {
        const int l = __ldg(p.ht + w0 + i);
        // 1: tensor-core window (mma.sync path below); 2: part of a dense super-window (tcgen05 kernel);
        // 0 / 3 (candidate left without a dense plan): CUDA cores, here or in the balanced kernel
        if (l == 2 || ((l == 0 || l == 3) && p.cuda_elsewhere)) sk |= 1u << i;
        else if (l == 1 && p.precision != HCSPMM_PRECISION_FP32 && (S & 7) == 0) m |= 1u << i;
      }
    s_tcmask = m;
    s_skipmask = sk;
  }
  __syncthreads();
  const unsigned tcmask0 = s_tcmask;
  const unsigned skipmask = s_skipmask;
  if (skipmask == (nwin >= 32 ? 0xffffffffu : ((1u << nwin) - 1u))) return;

  // tensor-core windows of this CTA, one after another (CTA-wide)
  if (tcmask0 != 0u) {
    for (int i = 0; i < nwin; ++i) {
      if (!((tcmask0 >> i) & 1u)) continue;
      const int e0 = rp[i * BLK_H], e1 = rp[min((i + 1) * BLK_H, rows_here)];
      if (e1 > e0) {
        tc_window<(LPE * NV * VW + 63) / 64>(p, w0 + i, e0, e1, feat0, S, smem);
      } else {
        // empty window labelled TC: rows are zero
        for (int t = tid; t < BLK_H * (S >> 2); t += CTA_THREADS) {
          const int r = i * BLK_H + t / (S >> 2), v = t % (S >> 2);
          if (r < rows_here && !p.accumulate)
            *reinterpret_cast<float4 *>(p.y + (long long)(r0 + r) * p.ldy + feat0 + v * 4) =
                make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      __syncthreads();
    }
    if ((tcmask0 | skipmask) == (nwin >= 32 ? 0xffffffffu : ((1u << nwin) - 1u))) return;
  }
  const unsigned tcmask = tcmask0 | skipmask;   // windows the CUDA-core phases must not touch

  const int q = lane / LPE, g = lane % LPE;
  bool active[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) active[i] = (g + i * LPE) < nvec;
  const bool all_active = nvec == LPE * NV;
  const float *xlane = p.x + (feat0 + g * VW) / (B16 ? 2 : 1);   // BF16 rows: float units are halved
  // phase A is for CTAs made of short rows only (mean population below p.short_row = knob * G);
  // in a mixed CTA the idle lane groups would cost more than the one-warp-per-row bookkeeping
  const int e_cta = rp[rows_here] - rp[0];
  const int short_row = (G > 1 && p.short_row > 0 && e_cta < rows_here * p.short_row &&
                         2 * rows_here >= CTA_WARPS * G)   // enough rows to occupy the lane groups
                            ? 4 * p.short_row : 0;

  // phase A: short rows, one lane group per row, static round-robin over the CTA's rows
  if (short_row > 0) {
    for (int r = wid * G + q; r < rows_here; r += CTA_WARPS * G) {
      const int eb = rp[r], ee = rp[r + 1];
      if (ee - eb >= short_row || ((tcmask >> (r / BLK_H)) & 1u)) continue;
      Vec<VW, B16> acc[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i].zero();
      gather_group_row<LPE, NV, VW, B16>(acc, xlane, p.ldx, p.x_rows, p.colidx, eb, ee, lane, q, g, active);
      float *yrow = p.y + (long long)(r0 + r) * p.ldy + feat0 + g * VW;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (active[i]) {
          if (p.accumulate) {
            Vec<VW, B16> o;
            o.load_plain(yrow + i * LPE * VW);
            acc[i].add(o);
          }
          acc[i].store(yrow + i * LPE * VW);
        }
    }
    __syncwarp();
  }

  // which of the remaining phases does this CTA need?  (one barrier-reduction each)
  int has_medium = 0, has_long = 0;
  for (int r = tid; r < rows_here; r += CTA_THREADS) {
    const int d = rp[r + 1] - rp[r];
    if (!((tcmask >> (r / BLK_H)) & 1u)) {
      has_long |= d >= p.long_row;
      has_medium |= d >= short_row && d < p.long_row && d > 0;
    }
    if (d == 0 && short_row == 0 && !p.accumulate && !((tcmask >> (r / BLK_H)) & 1u)) {
      // empty rows are written here when phase A is off (phase B skips them)
      for (int v = 0; v < (S >> 2); ++v)
        *reinterpret_cast<float4 *>(p.y + (long long)(r0 + r) * p.ldy + feat0 + v * 4) =
            make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  has_medium = __syncthreads_or(has_medium);
  has_long = __syncthreads_or(has_long);

  // phase B: medium rows, one warp per row, rows claimed dynamically
  while (has_medium) {
    int r = 0;
    if (lane == 0) r = atomicAdd(&s_next, 1);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= rows_here) break;
    const int eb = rp[r], ee = rp[r + 1];
    if (ee - eb >= p.long_row || ee - eb < short_row || ee == eb || ((tcmask >> (r / BLK_H)) & 1u)) continue;
    Vec<VW, B16> acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i].zero();
    gather_accumulate<LPE, NV, VW, B16>(acc, xlane, p.ldx, p.x_rows, p.colidx, eb, ee, 0, 1, lane, q, active,
                                   all_active);
    group_reduce<LPE, NV, VW, B16>(acc);
    if (q == 0) {
      float *yrow = p.y + (long long)(r0 + r) * p.ldy + feat0 + g * VW;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (active[i]) {
          if (p.accumulate) {
            Vec<VW, B16> o;
            o.load_plain(yrow + i * LPE * VW);
            acc[i].add(o);
          }
          acc[i].store(yrow + i * LPE * VW);
        }
    }
  }

  // phase C: long rows, all warps of the CTA share one row (CTA-uniform control flow)
  if (!has_long) return;
  const int nvec4 = S >> 2;
  for (int r = 0; r < rows_here; ++r) {
    const int eb = rp[r], ee = rp[r + 1];
    if (ee - eb < p.long_row || ((tcmask >> (r / BLK_H)) & 1u)) continue;
    Vec<VW, B16> acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i].zero();
    gather_accumulate<LPE, NV, VW, B16>(acc, xlane, p.ldx, p.x_rows, p.colidx, eb, ee, wid, CTA_WARPS, lane, q,
                                   active, all_active);
    group_reduce<LPE, NV, VW, B16>(acc);
    if (q == 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (active[i]) acc[i].store(smem + wid * S + (g + i * LPE) * VW);
    }
    __syncthreads();
    for (int v = tid; v < nvec4; v += CTA_THREADS) {
      float4 s = *reinterpret_cast<const float4 *>(smem + v * 4);
#pragma unroll
      for (int ww = 1; ww < CTA_WARPS; ++ww)
        add4(s, *reinterpret_cast<const float4 *>(smem + ww * S + v * 4));
      float4 *dst = reinterpret_cast<float4 *>(p.y + (long long)(r0 + r) * p.ldy + feat0 + v * 4);
      if (p.accumulate) add4(s, *dst);
      *dst = s;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// Work-balanced CUDA-core kernel.  Power-law graphs put 10^5 stored entries into single rows (and
// so into single 16-row windows): with one CTA per window the longest window bounds the kernel --
// at the Reddit shape it is 0.29 of a CTA slot's fair share on one GPU and 2.3 x the share of a
// 1/8 row shard.  Here the CUDA-core work is cut merge-path style instead: item k covers the
// k-th run of `chunk` steps through the merged sequence (row ends, stored entries), so every CTA
// gets the same rows + entries whatever the degree distribution.  A row that straddles item
// boundaries is summed in pieces: every item writes the piece of its first / last row into
// partial[k][0|1], and spmm_balanced_fixup_kernel adds the pieces of a row in item order -- a
// fixed order, so results do not depend on scheduling (no atomics).
// Rows of windows labelled tensor-core / dense are skipped (the hybrid kernel or the tcgen05
// kernel computes them); their entries still count as item steps.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int merge_path_rows(const int *__restrict__ rowptr, int n_rows, long long nnz,
                                               long long diag) {
  long long lo = diag > nnz ? diag - nnz : 0;
  long long hi = diag < n_rows ? diag : n_rows;
  while (lo < hi) {   // first row whose end marker lies beyond the diagonal
    const long long mid = (lo + hi) >> 1;
    if ((long long)__ldg(rowptr + mid + 1) <= diag - 1 - mid) lo = mid + 1;
    else hi = mid;
  }
  return (int)lo;
}

// All split points at once (one thread per diagonal): ~20 dependent loads each, but in parallel, so
// the item CTAs start with two loads instead of two binary searches.
__global__ void merge_path_splits_kernel(const int *__restrict__ rowptr, int n_rows, long long nnz, int chunk,
                                         int n_items, int *__restrict__ splits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_items) return;
  const long long total = (long long)n_rows + nnz;
  long long diag = (long long)i * chunk;
  if (diag > total) diag = total;
  splits[i] = merge_path_rows(rowptr, n_rows, nnz, diag);
}

template <int LPE, int NV, int VW, int MINB, bool B16 = false, int MODE = 0>
__global__ void __launch_bounds__(CTA_THREADS, MINB) spmm_balanced_kernel(const BalParams bp) {
  const SpmmParams &p = bp.s;
  extern __shared__ __align__(16) float smem[];   // [2 * CTA_WARPS * slab] row pieces | uint16 row offsets [chunk + 2]
  __shared__ int s_next;
  __shared__ int s_prow[CTA_WARPS][2];
  __shared__ const float *s_seg[8];
  constexpr int G = 32 / LPE;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int k = blockIdx.x;
  const int feat0 = blockIdx.y * p.slab;
  const int S = min(p.slab, p.dim - feat0);
  const int nvec = S / VW;
  const unsigned short *rp16 = reinterpret_cast<const unsigned short *>(smem + 2 * CTA_WARPS * p.slab);

  // the item's two diagonals: rows consumed (precomputed) and entries consumed
  const long long total = (long long)p.n_rows + bp.nnz;
  const long long d0 = min(total, (long long)k * bp.chunk), d1 = min(total, (long long)(k + 1) * bp.chunk);
  const int x0 = __ldg(bp.splits + min(k * bp.split_stride, bp.split_max));
  const int x1 = __ldg(bp.splits + min((k + 1) * bp.split_stride, bp.split_max));
  const int y0 = (int)(d0 - x0), y1 = (int)(d1 - x1);
  if (tid == 0) s_next = 0;
  if constexpr (MODE == 2) {
    if (tid < 8) s_seg[tid] = bp.seg_x[tid];   // published by the __syncthreads() below, before any gather
  }
  // entries [y0, y1); rows x0 .. x1-1 end here, row x1 takes part through its entries below y1
  const bool last_in = x1 < p.n_rows && __ldg(p.rowptr + x1) < y1;
  const int rows_here = x1 - x0 + (last_in ? 1 : 0);
  if (rows_here <= 0) {
    if (tid == 0 && blockIdx.y == 0) bp.split_row[2 * k] = bp.split_row[2 * k + 1] = -1;
    return;
  }
  // row ranges clipped to the item, as 16-bit offsets from y0 (chunk <= 16384): a small footprint
  // leaves the shared-memory / L1 carve-out to the gathered rows
  {
    unsigned short *w16 = reinterpret_cast<unsigned short *>(smem + 2 * CTA_WARPS * p.slab);
    for (int i = tid; i <= rows_here; i += CTA_THREADS)
      w16[i] = (unsigned short)(min(max(__ldg(p.rowptr + x0 + i), y0), y1) - y0);
  }
  auto rp = [&](int i) -> int { return y0 + (int)rp16[i]; };
  const int L = rows_here - 1;
  const int s0 = __ldg(p.rowptr + x0), t0 = __ldg(p.rowptr + x0 + 1);
  auto orig = [&](int row) -> int {   // row of Y / of the labels: MODE 3 walks the row-sorted copy (csrc/rowsort.cu)
    // (written with the run-time test on purpose: in the 80-register build ptxas spills 16 bytes per thread with it
    //  and 32 without -- products shape 2.93 vs 3.04 ms)
    if constexpr (MODE == 3) return bp.row_id ? __ldg(bp.row_id + row) : row;
    else return row;
  };
  auto elsewhere = [&](int row) -> bool {   // window labels 1 (mma.sync) / 2 (tcgen05 dense): not ours
    if (p.ht == nullptr) return false;
    const int l = __ldg(p.ht + (orig(row) >> 4));
    return l == 1 || l == 2;
  };
  const bool lab0 = elsewhere(x0);
  const bool lab1 = last_in && elsewhere(x1);
  const bool is_tail = last_in && !lab1 && __ldg(p.rowptr + x1 + 1) > y1;     // row x1 continues in item k+1
  const bool head_skip = s0 < y0 && t0 <= y0;                                 // row x0 was finished by item k-1
  const bool is_head = s0 < y0 && t0 > y0 && !lab0 && !(is_tail && L == 0);   // row x0 began earlier, ends here
  if (tid == 0 && blockIdx.y == 0) {
    bp.split_row[2 * k] = is_head ? x0 : -1;
    bp.split_row[2 * k + 1] = is_tail ? x1 : -1;
  }
  __syncthreads();

  auto skip = [&](int i) -> bool {
    return (i == 0 && head_skip) || elsewhere(x0 + i);
  };
  auto out_row = [&](int i, int &acc_flag) -> float * {
    if (i == 0 && is_head) { acc_flag = 0; return bp.partial + (size_t)(2 * k) * p.dim + feat0; }
    if (i == L && is_tail) { acc_flag = 0; return bp.partial + (size_t)(2 * k + 1) * p.dim + feat0; }
    acc_flag = p.accumulate;
    return p.y + (long long)orig(x0 + i) * p.ldy + feat0;
  };

  const int q = lane / LPE, g = lane % LPE;
  bool active[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) active[i] = (g + i * LPE) < nvec;
  const bool all_active = nvec == LPE * NV;
  const float *xlane = p.x + (feat0 + g * VW) / (B16 ? 2 : 1);
  const GatherHint hint = MODE == 2 ? make_segments(s_seg, (feat0 + g * VW) / (B16 ? 2 : 1))
                                    : (MODE == 1 ? make_hint(bp.hint_cls_min) : no_hint());
  const int e_cta = y1 - y0;
  const int short_row = (G > 1 && p.short_row > 0 && e_cta < rows_here * p.short_row &&
                         2 * rows_here >= CTA_WARPS * G)
                            ? 4 * p.short_row : 0;

  // warp-split mode (items that are not made of short rows): every warp takes an equal, contiguous
  // run of the item's entries and walks the rows it covers, so the eight warps finish together
  // whatever the row lengths.  A row cut by a warp boundary is summed in pieces through shared
  // memory, pieces added in warp order (same scheme as the item-level fix-up).
  if (short_row == 0 && bp.warp_split > 0 && e_cta >= rows_here * bp.warp_split) {
    float *pieces = smem;   // [CTA_WARPS][2][S]
    const int ew = ((e_cta + CTA_WARPS - 1) / CTA_WARPS + 31) & ~31;
    // first row with entries beyond position b (y0 <= b < y1)
    auto row_of = [&](int b) -> int {
      int lo = 0, hi = rows_here - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (rp(mid + 1) > b) hi = mid;
        else lo = mid + 1;
      }
      return lo;
    };
    // run boundary w: the equal cut, moved to a row boundary when one lies within a quarter run --
    // rows of moderate length then stay whole (no pieces, no pipeline restart inside them)
    auto boundary = [&](int w) -> int {
      const int b = y0 + w * ew;
      if (w <= 0) return y0;
      if (b >= y1) return y1;
      const int r = row_of(b);
      const int lo = rp(r), hi = rp(r + 1);
      const int cand = (b - lo <= hi - b) ? lo : hi;
      return (abs(cand - b) <= (ew >> 2)) ? cand : b;
    };
    const int wb = boundary(wid), we = boundary(wid + 1);
    int head_row = -1, tail_row = -1;
    if (wb < we) {
      const int lo = row_of(wb);
      for (int r = lo; r < rows_here && rp(r) < we; ++r) {
        const int eb = max(rp(r), wb), ee = min(rp(r + 1), we);
        if (ee <= eb || skip(r)) continue;
        Vec<VW, B16> acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i].zero();
        gather_accumulate<LPE, NV, VW, B16, MODE>(acc, xlane, p.ldx, p.x_rows, p.colidx, eb, ee, 0, 1, lane, q, active,
                                            all_active, hint);
        group_reduce<LPE, NV, VW, B16>(acc);
        int accf = 0;
        float *dst;
        if (rp(r + 1) > we) { dst = pieces + (wid * 2 + 1) * S; tail_row = r; }       // continues in the next warp
        else if (rp(r) < wb) { dst = pieces + (wid * 2) * S; head_row = r; }          // began in an earlier warp
        else dst = out_row(r, accf);
        if (q == 0) {
          dst += g * VW;
#pragma unroll
          for (int i = 0; i < NV; ++i)
            if (active[i]) {
              if (accf) {
                Vec<VW, B16> o;
                o.load_plain(dst + i * LPE * VW);
                acc[i].add(o);
              }
              acc[i].store(dst + i * LPE * VW);
            }
        }
      }
    }
    if (lane == 0) { s_prow[wid][0] = head_row; s_prow[wid][1] = tail_row; }
    // empty rows of the item (never pieces)
    if (!p.accumulate)
      for (int r = tid; r < rows_here; r += CTA_THREADS)
        if (rp(r + 1) == rp(r) && !skip(r))
          for (int v = 0; v < (S >> 2); ++v)
            *reinterpret_cast<float4 *>(p.y + (long long)orig(x0 + r) * p.ldy + feat0 + v * 4) =
                make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int b = 1; b < CTA_WARPS; ++b) {
      const int r = s_prow[b][0];
      if (r < 0) continue;
      int a = b;
      while (a > 0 && s_prow[a - 1][1] == r) --a;
      int accf;
      float *yrow = out_row(r, accf);
      for (int v = tid; v < (S >> 2); v += CTA_THREADS) {
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int w = a; w < b; ++w) add4(sum, *reinterpret_cast<const float4 *>(pieces + (w * 2 + 1) * S + v * 4));
        add4(sum, *reinterpret_cast<const float4 *>(pieces + (b * 2) * S + v * 4));
        float4 *dst = reinterpret_cast<float4 *>(yrow + v * 4);
        if (accf) add4(sum, *dst);
        *dst = sum;
      }
    }
    return;
  }

  // phase A: short rows, one lane group per row
  if (short_row > 0) {
    for (int r = wid * G + q; r < rows_here; r += CTA_WARPS * G) {
      const int eb = rp(r), ee = rp(r + 1);
      if (ee - eb >= short_row || skip(r)) continue;
      Vec<VW, B16> acc[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i].zero();
      gather_group_row<LPE, NV, VW, B16, MODE>(acc, xlane, p.ldx, p.x_rows, p.colidx, eb, ee, lane, q, g, active, hint);
      int accf;
      float *yrow = out_row(r, accf) + g * VW;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (active[i]) {
          if (accf) {
            Vec<VW, B16> o;
            o.load_plain(yrow + i * LPE * VW);
            acc[i].add(o);
          }
          acc[i].store(yrow + i * LPE * VW);
        }
    }
    __syncwarp();
  }

  int has_medium = 0, has_long = 0;
  for (int r = tid; r < rows_here; r += CTA_THREADS) {
    const int d = rp(r + 1) - rp(r);
    const bool sk = skip(r);
    if (!sk) {
      has_long |= d >= p.long_row;
      has_medium |= d >= short_row && d < p.long_row && d > 0;
    }
    if (d == 0 && short_row == 0 && !p.accumulate && !sk) {   // empty rows (never partial)
      for (int v = 0; v < (S >> 2); ++v)
        *reinterpret_cast<float4 *>(p.y + (long long)orig(x0 + r) * p.ldy + feat0 + v * 4) =
            make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  has_medium = __syncthreads_or(has_medium);
  has_long = __syncthreads_or(has_long);

  // phase B: medium rows, one warp per row, claimed dynamically
  while (has_medium) {
    int r = 0;
    if (lane == 0) r = atomicAdd(&s_next, 1);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= rows_here) break;
    const int eb = rp(r), ee = rp(r + 1);
    if (ee - eb >= p.long_row || ee - eb < short_row || ee == eb || skip(r)) continue;
    Vec<VW, B16> acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i].zero();
    gather_accumulate<LPE, NV, VW, B16, MODE>(acc, xlane, p.ldx, p.x_rows, p.colidx, eb, ee, 0, 1, lane, q, active,
                                        all_active, hint);
    group_reduce<LPE, NV, VW, B16>(acc);
    if (q == 0) {
      int accf;
      float *yrow = out_row(r, accf) + g * VW;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (active[i]) {
          if (accf) {
            Vec<VW, B16> o;
            o.load_plain(yrow + i * LPE * VW);
            acc[i].add(o);
          }
          acc[i].store(yrow + i * LPE * VW);
        }
    }
  }

  // phase C: long rows (or long pieces of a hub row), all warps share the row
  if (!has_long) return;
  const int nvec4 = S >> 2;
  for (int r = 0; r < rows_here; ++r) {
    const int eb = rp(r), ee = rp(r + 1);
    if (ee - eb < p.long_row || skip(r)) continue;
    Vec<VW, B16> acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i].zero();
    gather_accumulate<LPE, NV, VW, B16, MODE>(acc, xlane, p.ldx, p.x_rows, p.colidx, eb, ee, wid, CTA_WARPS, lane, q,
                                        active, all_active, hint);
    group_reduce<LPE, NV, VW, B16>(acc);
    if (q == 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (active[i]) acc[i].store(smem + wid * S + (g + i * LPE) * VW);
    }
    __syncthreads();
    int accf;
    float *yrow = out_row(r, accf);
    for (int v = tid; v < nvec4; v += CTA_THREADS) {
      float4 s = *reinterpret_cast<const float4 *>(smem + v * 4);
#pragma unroll
      for (int ww = 1; ww < CTA_WARPS; ++ww)
        add4(s, *reinterpret_cast<const float4 *>(smem + ww * S + v * 4));
      float4 *dst = reinterpret_cast<float4 *>(yrow + v * 4);
      if (accf) add4(s, *dst);
      *dst = s;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// Staged gather for low-degree graphs (rows of 260..512 bytes, every window CUDA-core).
//
// The balanced kernel above keeps gathered rows in REGISTERS while they are in flight: 4 ring slots per lane, 8 rows
// per warp, 24 warps per SM = 192 rows of X on their way per SM, and that is all the 80-register build can hold.  On
// the products shape the kernel then waits: long_scoreboard 12 warps per issue, DRAM at 0.66 of its copy bandwidth,
// issue slots 32 % busy -- it is short of memory-level parallelism, not of bandwidth.  Here the rows in flight live
// in SHARED MEMORY instead: every lane copies "its" 16 bytes of a row with cp.async (LDGSTS: no register is held
// while the copy is outstanding, no depth limit), STG_S groups of STG_K rows deep, and later adds the same 16 bytes
// back from shared memory -- lane L writes and reads only column L of the ring, so the pipeline needs no warp
// synchronisation at all, only cp.async.wait_group.  20 rows in flight per warp, 16 warps per SM = 320 rows per SM.
// A lane owns 4 features of the row (32 lanes x 16 bytes = 512 bytes), so a row sum is complete in its lane: no
// cross-lane reduction at a row end, which on short rows is most of the non-gather work.
//
// Work decomposition, item boundaries, partial rows and the fix-up pass are those of the balanced kernel (same
// BalParams, same split points): an item's entries are cut into eight equal warp runs; a warp streams its run
// straight through the row boundaries (a row end only flushes the lane's sum), so the pipeline never drains on short
// rows.  Sum order inside a row is CSR order, pieces are added in warp / item order: deterministic.
// ---------------------------------------------------------------------------------------
constexpr int STG_K = 4;                 // rows per cp.async group
constexpr int STG_S = 5;                 // groups in flight per warp
constexpr int STG_R = STG_K * STG_S;     // ring slots per warp
constexpr int STG_ROW = 128;             // floats per ring slot

static size_t staged_smem_bytes(int chunk) {
  const size_t rp_words = (((size_t)chunk + 4) / 2 + 3) & ~(size_t)3;
  return sizeof(float) * (2 * CTA_WARPS * STG_ROW + rp_words + (size_t)CTA_WARPS * STG_R * STG_ROW);
}

__global__ void __launch_bounds__(CTA_THREADS, 2) spmm_staged_kernel(const BalParams bp) {
  const SpmmParams &p = bp.s;
  extern __shared__ __align__(16) float smem[];   // [2 * CTA_WARPS * 128] row pieces | uint16 row offsets | rings
  __shared__ int s_prow[CTA_WARPS][2];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int k = blockIdx.x;
  const int S = p.dim;                            // 64 < S <= 128, S % 4 == 0, one feature slab
  const bool act = lane * 4 < S;
  const int rp_words = (((bp.chunk + 4) / 2) + 3) & ~3;
  const unsigned short *rp16 = reinterpret_cast<const unsigned short *>(smem + 2 * CTA_WARPS * STG_ROW);
  float *ring = smem + 2 * CTA_WARPS * STG_ROW + rp_words + (size_t)wid * STG_R * STG_ROW + lane * 4;

  const long long total = (long long)p.n_rows + bp.nnz;
  const long long d0 = min(total, (long long)k * bp.chunk), d1 = min(total, (long long)(k + 1) * bp.chunk);
  const int x0 = __ldg(bp.splits + min(k * bp.split_stride, bp.split_max));
  const int x1 = __ldg(bp.splits + min((k + 1) * bp.split_stride, bp.split_max));
  const int y0 = (int)(d0 - x0), y1 = (int)(d1 - x1);
  const bool last_in = x1 < p.n_rows && __ldg(p.rowptr + x1) < y1;
  const int rows_here = x1 - x0 + (last_in ? 1 : 0);
  if (rows_here <= 0) {
    if (tid == 0) bp.split_row[2 * k] = bp.split_row[2 * k + 1] = -1;
    return;
  }
  {
    unsigned short *w16 = reinterpret_cast<unsigned short *>(smem + 2 * CTA_WARPS * STG_ROW);
    for (int i = tid; i <= rows_here; i += CTA_THREADS)
      w16[i] = (unsigned short)(min(max(__ldg(p.rowptr + x0 + i), y0), y1) - y0);
  }
  auto rp = [&](int i) -> int { return y0 + (int)rp16[i]; };
  auto orig = [&](int row) -> int { return bp.row_id ? __ldg(bp.row_id + row) : row; };
  const int L = rows_here - 1;
  const int s0 = __ldg(p.rowptr + x0), t0 = __ldg(p.rowptr + x0 + 1);
  const bool is_tail = last_in && __ldg(p.rowptr + x1 + 1) > y1;          // row x1 continues in item k+1
  const bool head_skip = s0 < y0 && t0 <= y0;                              // row x0 was finished by item k-1
  const bool is_head = s0 < y0 && t0 > y0 && !(is_tail && L == 0);         // row x0 began earlier, ends here
  if (tid == 0) {
    bp.split_row[2 * k] = is_head ? x0 : -1;
    bp.split_row[2 * k + 1] = is_tail ? x1 : -1;
  }
  __syncthreads();
  auto out_row = [&](int i, int &acc_flag) -> float * {
    if (i == 0 && is_head) { acc_flag = 0; return bp.partial + (size_t)(2 * k) * p.dim; }
    if (i == L && is_tail) { acc_flag = 0; return bp.partial + (size_t)(2 * k + 1) * p.dim; }
    acc_flag = p.accumulate;
    return p.y + (long long)orig(x0 + i) * p.ldy;
  };

  float *pieces = smem;   // [CTA_WARPS][2][S]
  const int e_cta = y1 - y0;
  const int ew = ((e_cta + CTA_WARPS - 1) / CTA_WARPS + 31) & ~31;
  auto row_of = [&](int b) -> int {   // first row with entries beyond position b (y0 <= b < y1)
    int lo = 0, hi = rows_here - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (rp(mid + 1) > b) hi = mid;
      else lo = mid + 1;
    }
    return lo;
  };
  auto boundary = [&](int w) -> int {   // the equal cut, moved to a row boundary when one lies within a quarter run
    const int b = y0 + w * ew;
    if (w <= 0) return y0;
    if (b >= y1) return y1;
    const int r = row_of(b);
    const int lo = rp(r), hi = rp(r + 1);
    const int cand = (b - lo <= hi - b) ? lo : hi;
    return (abs(cand - b) <= (ew >> 2)) ? cand : b;
  };
  const int wb = boundary(wid), we = boundary(wid + 1);
  int head_row = -1, tail_row = -1;
  if (wb < we) {
    const int n_b = (we - wb + STG_K - 1) / STG_K;
    const float *xl = p.x + lane * 4;
    // column ids of the run, 32 at a time, one chunk ahead of the copies
    int jc = 0;
    int cid = (wb + lane < we) ? __ldg(p.colidx + wb + lane) : -1;
    int cid_next = (wb + 32 + lane < we) ? __ldg(p.colidx + wb + 32 + lane) : -1;
    int si = 0, sc = 0;                           // ring slot of the next batch to issue / to consume
    // Every lane always copies: lanes beyond the row width, entries beyond the run and invalid ids copy zero bytes, which
    // zero-fills their 16 bytes -- no branch around the shuffles, and the sums below need no predicate.  Row addresses
    // are base + id * pitch with a 32-bit byte pitch (one IMAD.WIDE; the launcher guarantees x_rows * pitch < 4 GB).
    const unsigned pitch = (unsigned)p.ldx * 4u;
    const char *xb = reinterpret_cast<const char *>(p.x) + lane * 16;
    const unsigned xr = act ? (unsigned)p.x_rows : 0u;
    (void)xl;
    auto issue = [&](int b) {
      const int o0 = b * STG_K;
      const int j = o0 >> 5;
      if (j != jc) {
        jc = j;
        cid = cid_next;
        const int nb = wb + (j + 1) * 32 + lane;
        cid_next = nb < we ? __ldg(p.colidx + nb) : -1;
      }
      float *slot = ring + si * STG_ROW;
      if (wb + o0 + STG_K <= we) {               // the whole batch lies inside the run (uniform)
#pragma unroll
        for (int kk = 0; kk < STG_K; ++kk) {
          const unsigned c = (unsigned)__shfl_sync(0xffffffffu, cid, (o0 + kk) & 31);
          const bool ok = c < xr;
          cp_async_16_ca(slot + kk * STG_ROW, xb + (size_t)(ok ? c : 0u) * pitch, ok ? 16 : 0);
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < STG_K; ++kk) {
          const unsigned c = (unsigned)__shfl_sync(0xffffffffu, cid, (o0 + kk) & 31);
          const bool ok = (wb + o0 + kk < we) && (c < xr);
          cp_async_16_ca(slot + kk * STG_ROW, xb + (size_t)(ok ? c : 0u) * pitch, ok ? 16 : 0);
        }
      }
      si = si + STG_K == STG_R ? 0 : si + STG_K;
      cp_async_commit();
    };
#pragma unroll
    for (int b = 0; b < STG_S - 1; ++b) issue(b);
    int r = row_of(wb);
    int cur_end = min(rp(r + 1), we);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < n_b; ++b) {
      issue(b + STG_S - 1);                      // refills the slots of batch b - 1, consumed last iteration
      cp_async_wait<STG_S - 1>();                // this lane's copies of batch b have landed
      const float *slot = ring + sc * STG_ROW;
      const int e0 = wb + b * STG_K;
      if (cur_end > e0 + STG_K) {                // no row ends inside this batch (uniform): four plain sums
#pragma unroll
        for (int kk = 0; kk < STG_K; ++kk) add4(acc, *reinterpret_cast<const float4 *>(slot + kk * STG_ROW));
      } else {
#pragma unroll
        for (int kk = 0; kk < STG_K; ++kk) {
          const int e = e0 + kk;
          add4(acc, *reinterpret_cast<const float4 *>(slot + kk * STG_ROW));   // zero beyond the run / the width
          if (e + 1 == cur_end) {                // the row's last entry inside this run: flush the lane's sum
            int accf = 0;
            float *dst;
            if (rp(r + 1) > we) { dst = pieces + (wid * 2 + 1) * S; tail_row = r; }       // continues in the next warp
            else if (rp(r) < wb) { dst = pieces + (wid * 2) * S; head_row = r; }          // began in an earlier warp
            else dst = out_row(r, accf);
            if (act) {
              float4 *d4 = reinterpret_cast<float4 *>(dst + lane * 4);
              if (accf) add4(acc, *d4);
              *d4 = acc;
            }
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
            do { ++r; } while (r < rows_here && rp(r + 1) <= e + 1);   // rows without entries are zero-filled below
            cur_end = r < rows_here ? min(rp(r + 1), we) : 0x7fffffff;
          }
        }
      }
      sc = sc + STG_K == STG_R ? 0 : sc + STG_K;
    }
    cp_async_wait<0>();
  }
  if (lane == 0) { s_prow[wid][0] = head_row; s_prow[wid][1] = tail_row; }
  if (!p.accumulate)
    for (int r = tid; r < rows_here; r += CTA_THREADS)
      if (rp(r + 1) == rp(r) && !(r == 0 && head_skip))
        for (int v = 0; v < (S >> 2); ++v)
          *reinterpret_cast<float4 *>(p.y + (long long)orig(x0 + r) * p.ldy + v * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  for (int b = 1; b < CTA_WARPS; ++b) {          // rows cut by warp boundaries: pieces added in warp order
    const int r = s_prow[b][0];
    if (r < 0) continue;
    int a = b;
    while (a > 0 && s_prow[a - 1][1] == r) --a;
    int accf;
    float *yrow = out_row(r, accf);
    for (int v = tid; v < (S >> 2); v += CTA_THREADS) {
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int w = a; w < b; ++w) add4(sum, *reinterpret_cast<const float4 *>(pieces + (w * 2 + 1) * S + v * 4));
      add4(sum, *reinterpret_cast<const float4 *>(pieces + (b * 2) * S + v * 4));
      float4 *dst = reinterpret_cast<float4 *>(yrow + v * 4);
      if (accf) add4(sum, *dst);
      *dst = sum;
    }
  }
}

// Y[r] (+)= sum of the pieces of every row that was cut by item boundaries, pieces in item order.
__global__ void __launch_bounds__(64) spmm_balanced_fixup_kernel(const BalParams bp) {
  const SpmmParams &p = bp.s;
  for (int k = blockIdx.x; k < bp.n_items; k += gridDim.x) {
    const int r = bp.split_row[2 * k];
    if (r < 0) continue;
    int k1 = k;
    while (k1 > 0 && bp.split_row[2 * (k1 - 1) + 1] == r) --k1;
    for (int f = threadIdx.x * 4; f < p.dim; f += 64 * 4) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int kk = k1; kk < k; ++kk)
        add4(s, *reinterpret_cast<const float4 *>(bp.partial + (size_t)(2 * kk + 1) * p.dim + f));
      add4(s, *reinterpret_cast<const float4 *>(bp.partial + (size_t)(2 * k) * p.dim + f));
      float4 *dst = reinterpret_cast<float4 *>(p.y + (long long)(bp.row_id ? bp.row_id[r] : r) * p.ldy + f);
      if (p.accumulate) add4(s, *dst);
      *dst = s;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Scalar kernel: any dim / alignment.  One warp per row, lane = feature (mod 32), FP32 sum in
// CSR order -- bit-identical to the reference's CUDA-core arithmetic (:1003-1010).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CTA_THREADS) spmm_scalar_kernel(const SpmmParams p) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * CTA_WARPS + (threadIdx.x >> 5);
  if (row >= p.n_rows) return;
  const int eb = __ldg(p.rowptr + row), ee = __ldg(p.rowptr + row + 1);
  for (int f0 = 0; f0 < p.dim; f0 += 128) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int base = eb; base < ee; base += 32) {
      const int n = min(32, ee - base);
      const int c = (lane < n) ? __ldg(p.colidx + base + lane) : -1;
#pragma unroll 4
      for (int j = 0; j < n; ++j) {
        const int cu = __shfl_sync(0xffffffffu, c, j);
        if ((unsigned)cu < (unsigned)p.x_rows) {
          const float *src = p.x + (long long)cu * p.ldx + f0 + lane;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (f0 + lane + 32 * i < p.dim) acc[i] += __ldg(src + 32 * i);
        }
      }
    }
    float *dst = p.y + row * p.ldy + f0 + lane;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (f0 + lane + 32 * i < p.dim) dst[32 * i] = p.accumulate ? dst[32 * i] + acc[i] : acc[i];
  }
}

// dst[r, 0..dpad) = src[r, 0..dim) zero-extended (dst dense with leading dim dpad)
__global__ void pad_rows_kernel(const float *__restrict__ src, long long lds, int dim, float *__restrict__ dst,
                                int dpad, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long r = i / dpad;
  const int c = (int)(i - r * dpad);
  dst[i] = c < dim ? __ldg(src + r * lds + c) : 0.f;
}
__global__ void unpad_rows_kernel(const float *__restrict__ src, int dpad, int dim, float *__restrict__ dst,
                                  long long ldd, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long r = i / dim;
  const int c = (int)(i - r * dim);
  dst[r * ldd + c] = src[r * dpad + c];
}

void launch_pad_rows(const float *src, int64_t lds, int32_t rows, int32_t dim, float *dst, int32_t dpad, cudaStream_t stream) {
  const long long total = (long long)rows * dpad;
  if (total > 0) pad_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, lds, dim, dst, dpad, total);
}
void launch_unpad_rows(const float *src, int32_t dpad, int32_t rows, int32_t dim, float *dst, int64_t ldd, cudaStream_t stream) {
  const long long total = (long long)rows * dim;
  if (total > 0) unpad_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, dpad, dim, dst, ldd, total);
}

// X (FP32, leading dim ldx) -> dense BF16 copy [rows, dim], round to nearest even
__global__ void f32_to_bf16_rows_kernel(const float *__restrict__ x, long long ldx, int dim2, uint32_t *__restrict__ xb,
                                        long long ldb2, long long total2) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total2; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / dim2;
    const int c = (int)(i - r * dim2) * 2;
    const float lo = __ldg(x + r * ldx + c), hi = __ldg(x + r * ldx + c + 1);
    const uint32_t ulo = __float_as_uint(lo), uhi = __float_as_uint(hi);
    // RNE on the upper 16 bits (NaN / Inf pass through the truncation unchanged enough for a gather-sum)
    const uint32_t rlo = ((ulo & 0x7f800000u) == 0x7f800000u) ? ulo : ulo + 0x7fffu + ((ulo >> 16) & 1u);
    const uint32_t rhi = ((uhi & 0x7f800000u) == 0x7f800000u) ? uhi : uhi + 0x7fffu + ((uhi >> 16) & 1u);
    xb[r * ldb2 + (c >> 1)] = (rlo >> 16) | (rhi & 0xffff0000u);
  }
}

static size_t hybrid_smem_bytes(int S, bool tc) {
  size_t cuda_path = (size_t)CTA_WARPS * S * sizeof(float);
  size_t tc_path = tc ? (size_t)(2 * (UCAP + KC) + NSTAGE * KC * (S + 8)) * sizeof(float) : 0;
  return cuda_path > tc_path ? cuda_path : tc_path;
}

template <int LPE, int NV, int VW, int MINB>
static cudaError_t launch_hybrid_b(const SpmmParams &p, dim3 grid, size_t smem, cudaStream_t stream) {
  cudaError_t err = cudaFuncSetAttribute(spmm_hybrid_kernel<LPE, NV, VW, MINB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  SpmmParams q = p;
  q.short_row = p.short_row * (32 / LPE);   // "fewer than short_row steps of a whole warp"
  spmm_hybrid_kernel<LPE, NV, VW, MINB><<<grid, CTA_THREADS, smem, stream>>>(q);
  return cudaGetLastError();
}
template <int LPE, int NV>
static cudaError_t launch_hybrid_bf16(const SpmmParams &p, dim3 grid, size_t smem, cudaStream_t stream) {
  cudaError_t err = cudaFuncSetAttribute(spmm_hybrid_kernel<LPE, NV, 8, HCSPMM_MIN_CTAS, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  SpmmParams q = p;
  q.short_row = p.short_row * (32 / LPE);
  spmm_hybrid_kernel<LPE, NV, 8, HCSPMM_MIN_CTAS, true><<<grid, CTA_THREADS, smem, stream>>>(q);
  return cudaGetLastError();
}

template <int LPE, int NV, int VW>
static cudaError_t launch_hybrid(const SpmmParams &p, dim3 grid, size_t smem, cudaStream_t stream) {
  // p.wpc > 1 <=> low-degree graph (few hundred entries per window)
  if (p.wpc > 1 && tuning().occupancy3) return launch_hybrid_b<LPE, NV, VW, 3>(p, grid, smem, stream);
  return launch_hybrid_b<LPE, NV, VW, HCSPMM_MIN_CTAS>(p, grid, smem, stream);
}

template <int LPE, int NV, int VW, bool B16, int MINB, int MODE = 0>
static cudaError_t launch_balanced_b(const BalParams &bp, dim3 grid, size_t smem, cudaStream_t stream) {
  auto kern = spmm_balanced_kernel<LPE, NV, VW, MINB, B16, MODE>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  BalParams q = bp;
  q.s.short_row = bp.s.short_row * (32 / LPE);
  kern<<<grid, CTA_THREADS, smem, stream>>>(q);
  return cudaGetLastError();
}
template <int LPE, int NV, int VW, bool B16>
static cudaError_t launch_balanced_t(const BalParams &bp, dim3 grid, size_t smem, cudaStream_t stream) {
  // low-degree graphs (rows of a few dozen entries) are latency-bound: three CTAs per SM hide more of it than
  // the deeper gather ring of the two-CTA build does (FP32, one vector per lane: dim <= 256 / 128)
  if (bp.row_id != nullptr) {   // the row-sorted copy of the CSR: its own build, so the in-place walk pays nothing for it
    if constexpr (!B16 && NV == 1)
      if (bp.low_degree && tuning().occupancy3) return launch_balanced_b<LPE, NV, VW, B16, 3, 3>(bp, grid, smem, stream);
    return launch_balanced_b<LPE, NV, VW, B16, HCSPMM_MIN_CTAS, 3>(bp, grid, smem, stream);
  }
  if (bp.seg_mode) {   // X in segments (peer-mapped operands read in place): 256-bit FP32 / BF16 rows up to 1 KB
    if constexpr (VW == 8 && NV == 1) {
      if constexpr (!B16)
        if (bp.low_degree && tuning().occupancy3) return launch_balanced_b<LPE, NV, VW, B16, 3, 2>(bp, grid, smem, stream);
      return launch_balanced_b<LPE, NV, VW, B16, HCSPMM_MIN_CTAS, 2>(bp, grid, smem, stream);
    } else {
      return cudaErrorInvalidValue;
    }
  }
  if constexpr (!B16 && VW == 8) {
    // tagged column ids + L2 residency hints (run_balanced decides; only these variants read tagged ids)
    if (bp.hint_cls_min >= 0) {
      if constexpr (NV == 1)
        if (bp.low_degree && tuning().occupancy3) return launch_balanced_b<LPE, NV, VW, B16, 3, true>(bp, grid, smem, stream);
      return launch_balanced_b<LPE, NV, VW, B16, HCSPMM_MIN_CTAS, true>(bp, grid, smem, stream);
    }
  }
  if constexpr (!B16 && NV == 1) {
    if (bp.low_degree && tuning().occupancy3 >= 2) return launch_balanced_b<LPE, NV, VW, B16, 4>(bp, grid, smem, stream);
    if (bp.low_degree && tuning().occupancy3) return launch_balanced_b<LPE, NV, VW, B16, 3>(bp, grid, smem, stream);
  }
  return launch_balanced_b<LPE, NV, VW, B16, HCSPMM_MIN_CTAS>(bp, grid, smem, stream);
}

// "balance" knob: 0 off, 2 always, 1 (default) when the mean row holds >= 8 entries -- below that the
// windows-per-CTA / lane-group-per-row machinery of the hybrid kernel is faster (envelope shape: 0.113 vs 0.172 ms)
static bool use_balanced(int n_rows, long long nnz) {
  const int b = tuning().balance;
  return b >= 2 || (b == 1 && nnz >= 8LL * n_rows);
}

// Caller-provided per-graph products (hcspmm_aux_t): precomputed merge-path split points and a workspace for the
// partial sums, so that a step of a static graph neither recomputes the splits nor allocates.
struct BalAux {
  const int *splits = nullptr;   // split points at diagonals i * splits_chunk
  int splits_chunk = 0, n_splits = 0;
  void *ws = nullptr;
  size_t ws_bytes = 0;
  const int *colidx_tagged = nullptr;   // hcspmm_tag_columns: column ids with their hotness class in bits 29..31
  const int *sorted_rowptr = nullptr, *sorted_colidx = nullptr, *sorted_row_id = nullptr;   // hcspmm_row_sort
  const int *colidx_seg = nullptr;      // column ids with the SEGMENT of their X row in bits 29..31 (segment mode)
  const float *seg_x[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

static int default_chunk(int slab, bool b16) {
  // item size: about 4 MB of gathered rows -- 4096 steps for rows of >= 1 KB, 8192 below (measured: Reddit
  // shape dim 256 FP32 best at 4096, dim 64 / BF16 / products dim 128 at 8192)
  const long long row_bytes = (long long)slab * (b16 ? 2 : 4);
  return row_bytes >= 1024 ? 4096 : 8192;
}

size_t balanced_workspace_bytes(int32_t n_rows, int64_t nnz, int32_t dim) {
  const long long n_items = ((long long)n_rows + nnz + 4095) / 4096 + 1;   // the smallest default item size
  return sizeof(float) * 2 * (size_t)n_items * (size_t)dim + sizeof(int) * (3 * (size_t)n_items + 2) + 256;
}

// CUDA-core windows of p (all of them when p.ht == nullptr) on the work-balanced kernel.
static cudaError_t run_balanced(const SpmmParams &p, long long nnz, bool v8, bool b16, const BalAux &aux,
                                cudaStream_t stream) {
  BalParams bp;
  bp.s = p;
  bp.nnz = nnz;
  bp.row_id = nullptr;
  if (aux.sorted_rowptr != nullptr) {      // the row-sorted copy (aux.splits belong to it)
    bp.s.rowptr = aux.sorted_rowptr;
    bp.s.colidx = aux.sorted_colidx;
    bp.row_id = aux.sorted_row_id;
  }
  const SpmmParams &ps = bp.s;
  int chunk = tuning().chunk > 0 ? tuning().chunk : default_chunk(p.slab, b16);
  if (chunk < 64) chunk = 64;
  if (chunk > 16384) chunk = 16384;
  const long long total = (long long)p.n_rows + nnz;
  // Items are equal pieces of work and one CTA takes one item, so a launch runs in whole "rounds" of the resident CTA
  // slots (148 SMs x 2 or 3): a 1/8 shard of the products shape is 980 items of 8192 steps on 444 slots = 2.2 rounds
  // paid as 3.  Small operands therefore take the smaller item size (the split points serve both), knob "chunk" unset.
  if (tuning().chunk <= 0 && chunk > 4096) {
    const long long slots = 148LL * ((nnz < 64LL * p.n_rows && tuning().occupancy3) ? 3 : 2);
    if (total / chunk < 6 * slots) chunk = 4096;
  }
  const long long n_items = total > 0 ? (total + chunk - 1) / chunk : 1;
  if (n_items > 0x7fffffffLL / 2) return cudaErrorInvalidValue;
  bp.chunk = chunk;
  bp.n_items = (int)n_items;
  const bool have_splits = aux.splits != nullptr && aux.splits_chunk > 0 && chunk % aux.splits_chunk == 0 &&
                           (long long)aux.n_splits == (total + aux.splits_chunk - 1) / aux.splits_chunk;
  void *ws = nullptr;
  const size_t part_bytes = sizeof(float) * 2 * (size_t)n_items * p.dim;
  const size_t need = part_bytes + sizeof(int) * (2 * (size_t)n_items + (have_splits ? 0 : (size_t)n_items + 1));
  const bool own_ws = !(aux.ws != nullptr && aux.ws_bytes >= need && (reinterpret_cast<uintptr_t>(aux.ws) & 15) == 0);
  cudaError_t err = cudaSuccess;
  if (own_ws) err = scratch_alloc(&ws, need, stream);
  else ws = aux.ws;
  if (err != cudaSuccess) return err;
  bp.partial = reinterpret_cast<float *>(ws);
  bp.split_row = reinterpret_cast<int *>(reinterpret_cast<char *>(ws) + part_bytes);
  if (have_splits) {
    bp.splits = aux.splits;
    bp.split_stride = chunk / aux.splits_chunk;
    bp.split_max = aux.n_splits;
  } else {
    int *splits = bp.split_row + 2 * (size_t)n_items;
    bp.splits = splits;
    bp.split_stride = 1;
    bp.split_max = (int)n_items;
    merge_path_splits_kernel<<<(unsigned)((n_items + 1 + 127) / 128), 128, 0, stream>>>(ps.rowptr, p.n_rows, nnz, chunk,
                                                                                      (int)n_items, splits);
  }
  bp.warp_split = tuning().warp_split;   // mean row length from which an item is cut by warp runs
  bp.low_degree = nnz < 64LL * p.n_rows;
  // L2 residency hints: only when X does not fit the budget anyway (otherwise every row is evict_last, as before)
  bp.hint_cls_min = -1;
  // Measured (scripts/r2/hint_probe.py): the hints pay only for rows of >= 2 KB (Reddit shape dim 512: 15.5 -> 14.4 ms);
  // at 512-byte / 1 KB rows they LOSE 5-15 % (products dim 128: 3.26 -> 3.74 ms, Reddit dim 256: 6.37 -> 6.73 ms) -- LRU
  // already keeps the hub rows that matter -- so they apply from knob "l2_hot_min_row" bytes per row upwards.
  if (aux.colidx_tagged != nullptr && aux.sorted_rowptr == nullptr && tuning().l2_hot_mb > 0 && v8 && !b16 &&
      (long long)p.dim * 4 >= tuning().l2_hot_min_row) {
    const long long row_bytes = (long long)p.dim * (b16 ? 2 : 4);
    const long long budget_rows = ((long long)tuning().l2_hot_mb << 20) / (row_bytes > 0 ? row_bytes : 1);
    if (budget_rows < (long long)p.x_rows) {
      int c = 1;
      while (c <= 7 && ((16384LL << (7 - c)) > budget_rows)) ++c;
      bp.hint_cls_min = c;               // 8: nothing fits -> every row evict_first
      bp.s.colidx = aux.colidx_tagged;
    }
  }
  bp.seg_mode = 0;
  for (int i = 0; i < 8; ++i) bp.seg_x[i] = nullptr;
  if (aux.colidx_seg != nullptr) {
    bp.seg_mode = 1;
    bp.hint_cls_min = -1;
    bp.s.colidx = aux.colidx_seg;
    bp.seg_x[0] = p.x;
    for (int i = 1; i < 8; ++i) bp.seg_x[i] = aux.seg_x[i] != nullptr ? aux.seg_x[i] : p.x;
  }
  const int slab = p.slab;
  dim3 grid((unsigned)n_items, (p.dim + slab - 1) / slab, 1);
  const size_t smem = (size_t)2 * CTA_WARPS * slab * sizeof(float) + (size_t)(chunk + 4) / 2 * sizeof(int);
  // low-degree graphs with rows of 260..512 bytes, every window CUDA-core: rows in flight staged in shared memory
  const bool staged = tuning().staged != 0 && !b16 && v8 && bp.low_degree && p.ht == nullptr && !bp.seg_mode &&
                      bp.hint_cls_min < 0 && slab == p.dim && p.dim > 64 && p.dim <= STG_ROW &&
                      (long long)p.x_rows * p.ldx * 4 < (1LL << 32);
  if (staged) {
    const size_t ssm = staged_smem_bytes(chunk);
    err = cudaFuncSetAttribute(spmm_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm);
    if (err == cudaSuccess) {
      spmm_staged_kernel<<<(unsigned)n_items, CTA_THREADS, ssm, stream>>>(bp);
      err = cudaGetLastError();
    }
  } else if (b16) {
    if (slab <= 32) err = launch_balanced_t<4, 1, 8, true>(bp, grid, smem, stream);
    else if (slab <= 64) err = launch_balanced_t<8, 1, 8, true>(bp, grid, smem, stream);
    else if (slab <= 128) err = launch_balanced_t<16, 1, 8, true>(bp, grid, smem, stream);
    else if (slab <= 256) err = launch_balanced_t<32, 1, 8, true>(bp, grid, smem, stream);
    else err = launch_balanced_t<32, 2, 8, true>(bp, grid, smem, stream);
  } else if (v8) {
    if (slab <= 32) err = launch_balanced_t<4, 1, 8, false>(bp, grid, smem, stream);
    else if (slab <= 64) err = launch_balanced_t<8, 1, 8, false>(bp, grid, smem, stream);
    else if (slab <= 128) err = launch_balanced_t<16, 1, 8, false>(bp, grid, smem, stream);
    else if (slab <= 256) err = launch_balanced_t<32, 1, 8, false>(bp, grid, smem, stream);
    else err = launch_balanced_t<32, 2, 8, false>(bp, grid, smem, stream);
  } else {
    if (slab <= 32) err = launch_balanced_t<8, 1, 4, false>(bp, grid, smem, stream);
    else if (slab <= 64) err = launch_balanced_t<16, 1, 4, false>(bp, grid, smem, stream);
    else if (slab <= 128) err = launch_balanced_t<32, 1, 4, false>(bp, grid, smem, stream);
    else if (slab <= 256) err = launch_balanced_t<32, 2, 4, false>(bp, grid, smem, stream);
    else err = launch_balanced_t<32, 4, 4, false>(bp, grid, smem, stream);
  }
  if (err == cudaSuccess && n_items > 1) {
    const int fgrid = (int)(n_items < 148 * 32 ? n_items : 148 * 32);
    spmm_balanced_fixup_kernel<<<fgrid, 64, 0, stream>>>(bp);
    err = cudaGetLastError();
  }
  if (own_ws) scratch_free(ws, stream);
  return err;
}

// out[r, 0..dim) (bfloat16, row pitch ld_out elements) = RNE(x[r, 0..dim)): how a rank writes its own rows into a
// BF16 exchange operand
int launch_f32_to_bf16(const float *x, int64_t ldx, int32_t rows, int32_t dim, void *out, int64_t ld_out,
                       cudaStream_t stream) {
  if (rows <= 0 || dim <= 0) return 0;
  if (!x || !out || (dim & 1) || (ld_out & 1) || ld_out < dim || ldx < dim || (reinterpret_cast<uintptr_t>(out) & 3)) {
    set_error("f32_to_bf16: dim and ld_out must be even, out 4-byte aligned");
    return HCSPMM_E_INVALID;
  }
  const long long total2 = (long long)rows * (dim / 2);
  f32_to_bf16_rows_kernel<<<1184, 256, 0, stream>>>(x, ldx, dim / 2, reinterpret_cast<uint32_t *>(out), ld_out / 2, total2);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("f32_to_bf16: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

int launch_merge_path_splits(const int32_t *rowptr, int32_t n_rows, int64_t nnz, int32_t chunk, int32_t *splits,
                             cudaStream_t stream) {
  if (!rowptr || !splits || n_rows < 0 || nnz < 0 || chunk < 64) { set_error("merge_path_splits: bad argument"); return HCSPMM_E_INVALID; }
  const long long total = (long long)n_rows + nnz;
  const long long n_items = (total + chunk - 1) / chunk;
  merge_path_splits_kernel<<<(unsigned)((n_items + 1 + 127) / 128), 128, 0, stream>>>(rowptr, n_rows, nnz, chunk, (int)n_items,
                                                                                    splits);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("merge_path_splits: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

int launch_spmm(const float *x, int64_t ldx, int32_t x_rows, const int32_t *rowptr,
                const int32_t *colidx, const int32_t *bp, const int32_t *etc, const int32_t *etr,
                const int32_t *ht, int32_t n_rows, int64_t nnz, int32_t dim, int precision,
                int accumulate, float *y, int64_t ldy, const hcspmm_aux_t *aux_in, cudaStream_t stream) {
  BalAux aux;
  int n_tc_windows = -1;   // windows labelled 1 (mma.sync path); -1 = unknown
  if (aux_in) {
    aux.splits = aux_in->d_splits; aux.splits_chunk = aux_in->splits_chunk; aux.n_splits = aux_in->n_splits;
    aux.ws = aux_in->d_workspace; aux.ws_bytes = aux_in->workspace_bytes;
    aux.colidx_tagged = aux_in->d_colidx_tagged;
    aux.colidx_seg = aux_in->d_colidx_segments;
    if (aux_in->d_sorted_rowptr && aux_in->d_sorted_row_id && (aux_in->d_sorted_colidx || nnz == 0) && !aux_in->d_colidx_segments) {
      aux.sorted_rowptr = aux_in->d_sorted_rowptr; aux.sorted_colidx = aux_in->d_sorted_colidx;
      aux.sorted_row_id = aux_in->d_sorted_row_id;
    } else if (aux_in->d_sorted_rowptr) {
      aux.splits = nullptr;   // the split points belong to the sorted copy, which this call does not use: recompute
    }
    for (int i = 0; i < 8; ++i) aux.seg_x[i] = reinterpret_cast<const float *>(aux_in->segment_x[i]);
    n_tc_windows = aux_in->n_tc_windows;
  }
  const bool seg_mode = aux.colidx_seg != nullptr;
  if (n_rows < 0 || dim < 0 || nnz < 0 || x_rows < 0) {
    set_error("spmm: negative size");
    return HCSPMM_E_INVALID;
  }
  if (n_rows == 0 || dim == 0) return 0;
  if (!x || !rowptr || !y || (!colidx && nnz > 0)) {
    set_error("spmm: null pointer argument");
    return HCSPMM_E_INVALID;
  }
  if (ldx < dim || ldy < dim) {
    set_error("spmm: leading dimension smaller than dim");
    return HCSPMM_E_INVALID;
  }
  if (precision < 0 || precision > HCSPMM_PRECISION_BF16_STORED) {
    set_error("spmm: unknown precision %d", precision);
    return HCSPMM_E_INVALID;
  }
  // BF16 mode: X is stored as bfloat16 for the gather (half the L2 / HBM traffic of the dominant
  // stream), FP32 accumulate, every window on the CUDA-core path.  Needs dim % 8 == 0 and an aligned
  // Y; anything else is computed in FP32 (which is within the BF16 tolerance a fortiori).
  const bool y_vec = (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (ldy & 3) == 0;
  if (precision == HCSPMM_PRECISION_BF16 && !((dim & 7) == 0 && y_vec)) precision = HCSPMM_PRECISION_FP32;
  // BF16_STORED: d_x already holds bfloat16 rows (ldx counted in bfloat16 elements) -- the multi-GPU operand the
  // halo exchange moved at half the bytes; no conversion pass, the same gather as BF16
  const bool stored16 = precision == HCSPMM_PRECISION_BF16_STORED;
  if (stored16) {
    if (!((dim & 7) == 0 && y_vec && (ldx & 7) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0)) {
      set_error("spmm: BF16-stored X needs dim and ldx multiples of 8 and 16-byte aligned X / Y");
      return HCSPMM_E_ALIGN;
    }
    precision = HCSPMM_PRECISION_BF16;
  }
  bool labels = ht != nullptr && precision != HCSPMM_PRECISION_FP32 && precision != HCSPMM_PRECISION_BF16;
  if (seg_mode) {
    // rows of X read in place from several buffers: only the balanced CUDA-core kernel resolves segment tags
    if (labels && n_tc_windows != 0) {
      set_error("spmm: segment-tagged column ids need a graph without tensor-core windows (n_tc_windows = 0)");
      return HCSPMM_E_INVALID;
    }
    labels = false;
  }
  if (labels && (!bp || (nnz > 0 && (!etc || !etr)))) {
    set_error("spmm: hybrid_type given without blockPartition/edgeToColumn/edgeToRow");
    return HCSPMM_E_INVALID;
  }
  SpmmParams p;
  p.x = x; p.ldx = ldx; p.x_rows = x_rows;
  p.rowptr = rowptr; p.colidx = colidx; p.bp = bp; p.etc = etc; p.etr = etr;
  p.ht = labels ? ht : nullptr;
  p.n_rows = n_rows; p.dim = dim;
  p.precision = precision; p.accumulate = accumulate;
  p.y = y; p.ldy = ldy;
  p.long_row = tuning().long_row > 0 ? tuning().long_row : 0x7fffffff;
  p.short_row = tuning().short_row;   // multiplied by G inside the launcher below
  p.wpc = 1;
  p.n_windows = (n_rows + BLK_H - 1) / BLK_H;
  p.cuda_elsewhere = 0;

  const int n_windows = (n_rows + BLK_H - 1) / BLK_H;
  const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 &&
                   (ldx & 3) == 0 && (ldy & 3) == 0 && (dim & 3) == 0;
  cudaError_t err;
  if (precision == HCSPMM_PRECISION_BF16) {
    uint32_t *xb = nullptr;
    if (stored16) {
      p.ldx = ldx / 2;   // float units
    } else {
      err = scratch_alloc((void **)&xb, sizeof(uint16_t) * (size_t)x_rows * dim, stream);
      if (err != cudaSuccess) { set_error("spmm bf16: scratch: %s", cudaGetErrorString(err)); return (int)err; }
      const long long total2 = (long long)x_rows * (dim / 2);
      f32_to_bf16_rows_kernel<<<1184, 256, 0, stream>>>(x, ldx, dim / 2, xb, dim / 2, total2);
      p.x = reinterpret_cast<const float *>(xb);
      p.ldx = dim / 2;
    }
    int slab = tuning().slab > 0 ? (tuning().slab + 31) / 32 * 32 : 512;
    if (slab > 512) slab = 512;
    if (slab > dim) slab = dim;
    p.slab = slab;
    const long long per_window = n_windows > 0 ? (long long)(nnz / n_windows) : 0;
    int wpc = tuning().wpc;
    if (wpc <= 0) { wpc = 1; while (wpc < MAX_WPC && per_window * wpc < 4096) wpc <<= 1; }
    if (wpc > MAX_WPC) wpc = MAX_WPC;
    p.wpc = wpc;
    p.n_windows = n_windows;
    dim3 grid((n_windows + wpc - 1) / wpc, (dim + slab - 1) / slab, 1);
    const size_t smem = hybrid_smem_bytes(slab, false);
    if (seg_mode && (slab < dim || dim > 256)) {
      if (xb) scratch_free(xb, stream);
      set_error("spmm: segment mode serves rows of at most 256 features (column blocks beyond)");
      return HCSPMM_E_INVALID;
    }
    if (use_balanced(n_rows, nnz) || seg_mode) err = run_balanced(p, nnz, true, true, aux, stream);
    else if (slab <= 32) err = launch_hybrid_bf16<4, 1>(p, grid, smem, stream);
    else if (slab <= 64) err = launch_hybrid_bf16<8, 1>(p, grid, smem, stream);
    else if (slab <= 128) err = launch_hybrid_bf16<16, 1>(p, grid, smem, stream);
    else if (slab <= 256) err = launch_hybrid_bf16<32, 1>(p, grid, smem, stream);
    else err = launch_hybrid_bf16<32, 2>(p, grid, smem, stream);
    if (xb) scratch_free(xb, stream);
    if (err != cudaSuccess) { set_error("spmm bf16 launch: %s", cudaGetErrorString(err)); return (int)err; }
    return 0;
  }
  if (seg_mode && !(vec && tuning().vec8 != 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 31) == 0 &&
                    (ldx & 7) == 0 && (ldy & 7) == 0 && (dim & 7) == 0 && dim <= 256)) {
    set_error("spmm: segment mode needs 32-byte aligned rows of at most 256 features (FP32)");
    return HCSPMM_E_ALIGN;
  }
  if (!vec && (long long)n_rows * dim >= (1 << 20) && tuning().pad_odd) {
    // Large operand with an odd width / unaligned rows (e.g. dim = 47 classes): run the vector
    // kernel on 32-byte-aligned padded copies instead of the one-warp-per-row scalar kernel.
    // Two streaming copies cost 2 * (x_rows + n_rows) * dim * 4 bytes -- small next to the gather.
    const int dpad = (dim + 7) / 8 * 8;
    float *xp = nullptr, *yp = nullptr;
      err = scratch_alloc((void **)&xp, sizeof(float) * (size_t)x_rows * dpad, stream);
    if (err == cudaSuccess) err = scratch_alloc((void **)&yp, sizeof(float) * (size_t)n_rows * dpad, stream);
    if (err != cudaSuccess) {
      if (xp) scratch_free(xp, stream);
      set_error("spmm: cudaMallocAsync: %s", cudaGetErrorString(err));
      return (int)err;
    }
    const long long nx = (long long)x_rows * dpad, ny = (long long)n_rows * dpad;
    pad_rows_kernel<<<(unsigned)((nx + 255) / 256), 256, 0, stream>>>(x, ldx, dim, xp, dpad, nx);
    if (accumulate) pad_rows_kernel<<<(unsigned)((ny + 255) / 256), 256, 0, stream>>>(y, ldy, dim, yp, dpad, ny);
    int rc = launch_spmm(xp, dpad, x_rows, rowptr, colidx, bp, etc, etr, ht, n_rows, nnz, dpad, precision,
                         accumulate, yp, dpad, aux_in, stream);
    if (rc == 0) {
      unpad_rows_kernel<<<(unsigned)(((long long)n_rows * dim + 255) / 256), 256, 0, stream>>>(
          yp, dpad, dim, y, ldy, (long long)n_rows * dim);
      err = cudaGetLastError();
      if (err != cudaSuccess) { set_error("spmm pad path: %s", cudaGetErrorString(err)); rc = (int)err; }
    }
    scratch_free(xp, stream);
    scratch_free(yp, stream);
    return rc;
  }
  if (!vec) {
    p.slab = dim;
    spmm_scalar_kernel<<<(n_rows + CTA_WARPS - 1) / CTA_WARPS, CTA_THREADS, 0, stream>>>(p);
    err = cudaGetLastError();
  } else {
    int slab = tuning().slab > 0 ? (tuning().slab + 31) / 32 * 32 : 512;
    if (slab > 512) slab = 512;
    if (slab > dim) slab = dim;
    p.slab = slab;
    // low-degree graphs: several windows per CTA so that a CTA has a few thousand entries to chew on
    // (measured on the products shape, 405 entries / window: 1 -> 4 windows per CTA = -28 % time)
    const long long per_window = n_windows > 0 ? (long long)(nnz / n_windows) : 0;
    int wpc = tuning().wpc;
    if (wpc <= 0) {
      wpc = 1;
      while (wpc < MAX_WPC && per_window * wpc < 4096) wpc <<= 1;
    }
    if (wpc > MAX_WPC) wpc = MAX_WPC;
    p.wpc = wpc;
    p.n_windows = n_windows;
    dim3 grid((n_windows + wpc - 1) / wpc, (dim + slab - 1) / slab, 1);
    const bool tc = labels && (slab % 8 == 0);
    const size_t smem = hybrid_smem_bytes(slab, tc);
    // 256-bit loads need 32-byte aligned rows in every slab
    const bool v8 = tuning().vec8 != 0 &&
                    ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 31) == 0 &&
                    (ldx & 7) == 0 && (ldy & 7) == 0 && (dim & 7) == 0 && (slab & 7) == 0;
    if (use_balanced(n_rows, nnz) || seg_mode) {
      // tensor-core windows (if any are labelled) on the per-window kernel, everything else balanced
      err = cudaSuccess;
      if (labels && n_tc_windows != 0) {   // no window labelled 1 (e.g. the shipped selector): nothing to launch
        SpmmParams pt = p;
        pt.cuda_elsewhere = 1;
        if (v8) {
          if (slab <= 32) err = launch_hybrid<4, 1, 8>(pt, grid, smem, stream);
          else if (slab <= 64) err = launch_hybrid<8, 1, 8>(pt, grid, smem, stream);
          else if (slab <= 128) err = launch_hybrid<16, 1, 8>(pt, grid, smem, stream);
          else if (slab <= 256) err = launch_hybrid<32, 1, 8>(pt, grid, smem, stream);
          else err = launch_hybrid<32, 2, 8>(pt, grid, smem, stream);
        } else {
          if (slab <= 32) err = launch_hybrid<8, 1, 4>(pt, grid, smem, stream);
          else if (slab <= 64) err = launch_hybrid<16, 1, 4>(pt, grid, smem, stream);
          else if (slab <= 128) err = launch_hybrid<32, 1, 4>(pt, grid, smem, stream);
          else if (slab <= 256) err = launch_hybrid<32, 2, 4>(pt, grid, smem, stream);
          else err = launch_hybrid<32, 4, 4>(pt, grid, smem, stream);
        }
      }
      if (err == cudaSuccess) {
        // no window is labelled 1 and no dense plan relabelled any (its labels arrive with n_tc_windows = -1): every
        // window is the balanced kernel's, which then needs no label look-up per row
        SpmmParams pb = p;
        if (n_tc_windows == 0) pb.ht = nullptr;
        err = run_balanced(pb, nnz, v8, false, aux, stream);
      }
    } else if (v8) {
      if (slab <= 32) err = launch_hybrid<4, 1, 8>(p, grid, smem, stream);
      else if (slab <= 64) err = launch_hybrid<8, 1, 8>(p, grid, smem, stream);
      else if (slab <= 128) err = launch_hybrid<16, 1, 8>(p, grid, smem, stream);
      else if (slab <= 256) err = launch_hybrid<32, 1, 8>(p, grid, smem, stream);
      else err = launch_hybrid<32, 2, 8>(p, grid, smem, stream);
    } else {
      if (slab <= 32) err = launch_hybrid<8, 1, 4>(p, grid, smem, stream);
      else if (slab <= 64) err = launch_hybrid<16, 1, 4>(p, grid, smem, stream);
      else if (slab <= 128) err = launch_hybrid<32, 1, 4>(p, grid, smem, stream);
      else if (slab <= 256) err = launch_hybrid<32, 2, 4>(p, grid, smem, stream);
      else err = launch_hybrid<32, 4, 4>(p, grid, smem, stream);
    }
  }
  if (err != cudaSuccess) {
    set_error("spmm launch: %s", cudaGetErrorString(err));
    return (int)err;
  }
  return 0;
}

}  // namespace hcspmm
