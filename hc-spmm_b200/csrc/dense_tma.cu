// dense_tma.cu -- the dense super-window path on TMA + tcgen05, with the Update product FUSED onto the
// TMEM-resident aggregate (reference: the fused Aggregation+Update kernels, hybrid_all_kernel.cu:1639-1848,
// 2067-2317, 2572-2770; `out = (A X) W` per row window with Z never leaving the SM before it is multiplied).
//
//   Z[128, D] = Abits[128, U] * X[cols[0..U), D]          phase 1, per 128-row super-window (plan of dense.cu)
//   out[128, H] = rna(Z)[128, D] * rna(W)[D, H]           phase 2 (fused entry points only)
//
// One persistent CTA per SM, fifteen warps in five roles, no block barrier in the steady state:
//   warps 0-1   TMA PRODUCERS  X rows are gathered by cp.async.bulk.tensor ... tile::gather4 (four row indices per
//               instruction) straight into the MN-major SWIZZLE_128B_BASE32B operand layout; the tensor map is typed
//               TFLOAT32, so the TMA unit rounds FP32 -> TF32 on the way in and neither a rounded copy of X nor a
//               register pass exists.  Column ids / row masks of the plan arrive by 1-D bulk copies, double buffered.
//               In phase 2 warp 0 streams 32-wide k-blocks of W^T (K-major SWIZZLE_128B) through the SAME ring.
//   warp 2      MMA ISSUER     phase 1: 4 x tcgen05.mma (M128, N = D, K8) per stage, A = the 0/1 tile in shared memory;
//               phase 2: tcgen05.mma with the A operand IN TENSOR MEMORY -- the aggregate's own accumulator columns,
//               rounded in place by the epilogue warps -- against the W^T k-block; both accumulators live in TMEM.
//   warps 3-10  EXPANDERS      the 128 x 32 0/1 operand tile of a stage from the plan's bit masks (K-major SWIZZLE_128B).
//   warps 11-14 EPILOGUE       tcgen05.ld Z -> global Z (exact FP32) and, fused, cvt.rna -> tcgen05.st back into TMEM;
//               then out from its accumulator -> global.
// Both accumulators are double buffered when 2 (D + H) <= 512 TMEM columns, so the next super-window's MMAs overlap
// this one's epilogue.  SASS: UTMALDG (gather4 + tiled), UBLKCP, UTCHMMA (SS and TS forms), UTCBAR, LDTM, STTM.
#include <cuda.h>

#include "common.cuh"
#include "umma.cuh"

namespace hcspmm {

bool make_tensor_map_2d(CUtensorMap *map, const void *base, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_cols,
                        uint32_t box_rows, int swizzle, int tf32_type);

constexpr int DT_SW_H = 128;
constexpr int DT_KC = 32;                       // condensed columns per stage
// GATHER = 0: X rows by TMA gather4 (warps 0-1), 8 expander warps.  GATHER = 1: X rows by 16-byte cp.async from 16
// expander warps (each also builds its share of the 0/1 tile), from a TF32-rounded copy of X; warp 0 still moves the
// plan's index chunks and the W^T k-blocks with TMA.  Measured (proteins shape, D = 256): gather4 moves 512 bytes per
// TMA operation and the TMA unit serves one operation per ~46 cycles per SM, so 64 operations per 32 KB stage bound the
// kernel at 1.24 ms, where 2048 cp.async pieces from 512 threads reach 0.64 ms -- GATHER = 1 is the default.
constexpr int DT_PROD_WARPS = 2, DT_EPI_WARPS = 4;
constexpr int DT_WARP_MMA = DT_PROD_WARPS;      // 2
constexpr int DT_WARP_EXP = DT_WARP_MMA + 1;    // 3
__host__ __device__ constexpr int dt_exp_warps(int gather) { return gather ? 16 : 8; }
__host__ __device__ constexpr int dt_warp_epi(int gather) { return DT_WARP_EXP + dt_exp_warps(gather); }
__host__ __device__ constexpr int dt_threads(int gather) { return 32 * (dt_warp_epi(gather) + DT_EPI_WARPS); }
__host__ __device__ constexpr uint32_t dt_full_count(int gather) {
  return gather ? 1u + 2u * dt_exp_warps(1) * 32u : (uint32_t)DT_PROD_WARPS + dt_exp_warps(0) * 32u;
}
constexpr int DT_IDX = 512;                     // condensed columns per index chunk (2 KB ids + 8 KB masks)
constexpr int DT_MAX_STAGES = 6;
constexpr uint32_t DT_A_BYTES = DT_SW_H * 128;  // 16 KB

struct DenseTmaParams {
  const float *xr;       // GATHER = 1: TF32-rounded dense copy of X [x_rows, dim]
  int x_rows, dim, n_rows, n_dense, accumulate;
  const int *sw_ids, *sw_off, *cols;
  const unsigned *masks;
  float *z;
  long long ldz;
  int hidden, hp;        // fused Update: hidden > 0; hp = hidden rounded up to 16 (N of the second product)
  float *out;
  long long ldo;
  int stages, nbuf;      // ring depth; accumulator sets in TMEM (1 or 2)
  int acc_stride, o_off, tmem_cols;
  int *err;
};

// ---- small PTX wrappers (this file only) --------------------------------------------------------------------
namespace dt {
__device__ __forceinline__ void expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3,
                                        uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(umma::smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ void load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(umma::smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void bulk_1d(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(umma::smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// the sequence of (super-window, index chunk) pairs a CTA walks -- every role iterates it identically
struct Walk {
  int ti, ch, c0, ucols, nchunks, sw;
  bool valid;
};
__device__ __forceinline__ void load_tile(Walk &w, const DenseTmaParams &p) {
  w.valid = w.ti < p.n_dense;
  if (!w.valid) return;
  w.sw = __ldg(p.sw_ids + w.ti);
  w.c0 = __ldg(p.sw_off + w.ti);
  w.ucols = __ldg(p.sw_off + w.ti + 1) - w.c0;
  w.nchunks = (w.ucols + DT_IDX - 1) / DT_IDX;
  w.ch = 0;
}
__device__ __forceinline__ Walk walk_first(const DenseTmaParams &p) {
  Walk w;
  w.ti = blockIdx.x;
  load_tile(w, p);
  return w;
}
__device__ __forceinline__ Walk walk_next(Walk w, const DenseTmaParams &p) {
  if (++w.ch < w.nchunks) return w;
  w.ti += gridDim.x;
  load_tile(w, p);
  return w;
}
}  // namespace dt

__device__ __forceinline__ void dt_cp_async_arrive_noinc(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

template <int GATHER>
__global__ void __launch_bounds__(dt_threads(GATHER), 1)
spmm_dense_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                      const DenseTmaParams p) {
  constexpr int DT_EXP_WARPS = dt_exp_warps(GATHER);
  constexpr int DT_WARP_EPI = dt_warp_epi(GATHER);
  constexpr uint32_t DT_FULL_COUNT = dt_full_count(GATHER);
  constexpr uint32_t DT_IEMPTY_COUNT = DT_PROD_WARPS + DT_EXP_WARPS * 32;
  extern __shared__ __align__(1024) uint8_t dt_smem[];
  __shared__ __align__(8) uint64_t bar_full[DT_MAX_STAGES], bar_empty[DT_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_ifull[2], bar_iempty[2];
  __shared__ __align__(8) uint64_t bar_zfull[2], bar_zready[2], bar_ofull[2], bar_free[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int D = p.dim;
  const int natoms = (D + 31) / 32;
  const uint32_t b_sbo = (uint32_t)natoms * 512u;             // between 4-row k-atoms of the gathered tile
  const uint32_t b_bytes = (DT_KC / 4) * b_sbo;
  const uint32_t w_bytes = (uint32_t)p.hp * 128u;             // one 32-wide k-block of W^T (phase 2)
  const uint32_t bw_bytes = b_bytes > w_bytes ? b_bytes : w_bytes;
  const uint32_t stage_bytes = DT_A_BYTES + ((bw_bytes + 1023u) & ~1023u);
  const uint32_t smem_base = (umma::smem_u32(dt_smem) + 1023u) & ~1023u;
  uint8_t *gen = dt_smem + (smem_base - umma::smem_u32(dt_smem));
  const int S = p.stages;
  const uint32_t idx_base = smem_base + (uint32_t)S * stage_bytes;   // [2][DT_IDX] ids | [2][DT_IDX][4] masks
  const int *cols_s = reinterpret_cast<const int *>(gen + (uint32_t)S * stage_bytes);
  const unsigned *masks_s = reinterpret_cast<const unsigned *>(cols_s + 2 * DT_IDX);
  const bool fused = p.hidden > 0;
  const int nkb2 = fused ? (D + DT_KC - 1) / DT_KC : 0;       // W^T k-blocks of phase 2

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { umma::mbar_init(&bar_full[s], DT_FULL_COUNT); umma::mbar_init(&bar_empty[s], 1); }
    for (int a = 0; a < 2; ++a) {
      umma::mbar_init(&bar_ifull[a], 1);
      umma::mbar_init(&bar_iempty[a], DT_IEMPTY_COUNT);
      umma::mbar_init(&bar_zfull[a], 1);
      umma::mbar_init(&bar_zready[a], DT_EPI_WARPS * 32);
      umma::mbar_init(&bar_ofull[a], 1);
      umma::mbar_init(&bar_free[a], DT_EPI_WARPS * 32);
    }
    umma::fence_barrier_init();
    if (GATHER == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
    if (fused) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
  }
  if (wid == DT_WARP_MMA) umma::tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  bool ok = true;

  if (wid < DT_PROD_WARPS) {
    // ===================== TMA producers =====================
    if (lane == 0) {
      auto issue_idx = [&](const dt::Walk &w, uint32_t q) {
        const uint32_t slot = q & 1u;
        const int cbase = w.c0 + w.ch * DT_IDX, ncols = min(DT_IDX, w.ucols - w.ch * DT_IDX);
        dt::expect_tx(&bar_ifull[slot], (uint32_t)ncols * 20u);
        dt::bulk_1d(idx_base + slot * DT_IDX * 4u, p.cols + cbase, (uint32_t)ncols * 4u, &bar_ifull[slot]);
        dt::bulk_1d(idx_base + 2u * DT_IDX * 4u + slot * DT_IDX * 16u, p.masks + 4ll * cbase, (uint32_t)ncols * 16u,
                    &bar_ifull[slot]);
      };
      dt::Walk w = dt::walk_first(p);
      uint32_t g = 0, q = 0;
      if (wid == 0 && w.valid) issue_idx(w, 0);
      while (w.valid) {
        const dt::Walk nx = dt::walk_next(w, p);
        ok = umma::mbar_wait(&bar_ifull[q & 1u], (q >> 1) & 1u) && ok;
        const int nst = w.ucols / DT_KC;
        const int f0 = w.ch * (DT_IDX / DT_KC), f1 = min(nst, f0 + DT_IDX / DT_KC);
        const int *cs = cols_s + (q & 1u) * DT_IDX;
        bool prefetched = !(wid == 0 && nx.valid);
        for (int f = f0; f < f1; ++f, ++g) {
          const uint32_t s = g % S;
          ok = umma::mbar_wait(&bar_empty[s], ((g / S) & 1u) ^ 1u) && ok;
          if (!prefetched && f - f0 > S) {
            // the next index chunk goes into the slot chunk q-1 used: its consumers are at most S stages behind
            // this thread, so by now the wait does not block and the ring never drains at a chunk boundary
            ok = umma::mbar_wait(&bar_iempty[(q + 1) & 1u], (((q + 1) >> 1) & 1u) ^ 1u) && ok;
            issue_idx(nx, q + 1);
            prefetched = true;
          }
          if (GATHER == 1) {
            if (wid == 0) dt::arrive(&bar_full[s]);       // the expanders gather; this arrival keeps the count uniform
            continue;
          }
          const uint32_t sb = smem_base + s * stage_bytes + DT_A_BYTES;
          dt::expect_tx(&bar_full[s], b_bytes / DT_PROD_WARPS);
          const int kb = (f - f0) * DT_KC;
#pragma unroll
          for (int kq = 0; kq < DT_KC / 4 / DT_PROD_WARPS; ++kq) {
            const int ka = wid * (DT_KC / 4 / DT_PROD_WARPS) + kq;       // this warp's 4-row k-atoms
            const int4 c4 = *reinterpret_cast<const int4 *>(cs + kb + ka * 4);
            // padding (-1) and ids outside the operand (rectangular shards) read row x_rows: out of bounds,
            // the TMA unit fills zeros
            const int r0 = (unsigned)c4.x < (unsigned)p.x_rows ? c4.x : p.x_rows;
            const int r1 = (unsigned)c4.y < (unsigned)p.x_rows ? c4.y : p.x_rows;
            const int r2 = (unsigned)c4.z < (unsigned)p.x_rows ? c4.z : p.x_rows;
            const int r3 = (unsigned)c4.w < (unsigned)p.x_rows ? c4.w : p.x_rows;
            for (int na = 0; na < natoms; ++na)
              dt::gather4(sb + ka * b_sbo + na * 512, &tm_x, na * 32, r0, r1, r2, r3, &bar_full[s]);
          }
        }
        if (!prefetched) {
          ok = umma::mbar_wait(&bar_iempty[(q + 1) & 1u], (((q + 1) >> 1) & 1u) ^ 1u) && ok;
          issue_idx(nx, q + 1);
        }
        dt::arrive(&bar_iempty[q & 1u]);
        ++q;
        if (fused && w.ch == w.nchunks - 1) {
          // phase 2 operand: W^T k-blocks through the same ring (warp 0 loads, warp 1 only keeps the count)
          for (int kb = 0; kb < nkb2; ++kb, ++g) {
            const uint32_t s = g % S;
            ok = umma::mbar_wait(&bar_empty[s], ((g / S) & 1u) ^ 1u) && ok;
            if (wid == 0) {
              dt::expect_tx(&bar_full[s], w_bytes);
              dt::load_2d(smem_base + s * stage_bytes + DT_A_BYTES, &tm_w, kb * DT_KC, 0, &bar_full[s]);
            } else if (GATHER == 0) {
              dt::arrive(&bar_full[s]);
            }
          }
        }
        w = nx;
      }
    }
    __syncwarp();
  } else if (wid == DT_WARP_MMA) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc1 = umma::make_idesc_tf32(DT_SW_H, D, /*A K-major*/ 0, /*B MN-major*/ 1);
      const uint32_t idesc2 = umma::make_idesc_tf32(DT_SW_H, p.hp, 0, /*B K-major*/ 0);
      uint32_t g = 0, it = 0;
      for (int ti = blockIdx.x; ti < p.n_dense; ti += gridDim.x, ++it) {
        const int nst = (__ldg(p.sw_off + ti + 1) - __ldg(p.sw_off + ti)) / DT_KC;
        const uint32_t a = p.nbuf == 2 ? (it & 1u) : 0u, use = p.nbuf == 2 ? (it >> 1) : it;
        ok = umma::mbar_wait(&bar_free[a], (use & 1u) ^ 1u) && ok;       // both accumulators of set a drained
        umma::tc_fence_after_sync();
        const uint32_t tz = tmem_base + a * (uint32_t)p.acc_stride, to = tz + (uint32_t)p.o_off;
        for (int s0 = 0; s0 < nst; ++s0, ++g) {
          const uint32_t s = g % S;
          ok = umma::mbar_wait(&bar_full[s], (g / S) & 1u) && ok;
          // (GATHER = 1: the expanders' cp.async / st.shared writes are generic-proxy writes: order them before the
          // tensor core's async-proxy reads on the consumer side too)
          umma::fence_proxy_async_smem();
          umma::tc_fence_after_sync();
          const uint32_t sa = smem_base + s * stage_bytes, sb = sa + DT_A_BYTES;
#pragma unroll
          for (int j = 0; j < DT_KC / 8; ++j) {
            const uint64_t da = umma::make_desc_sw128(sa + j * 32, 16, 1024);
            const uint64_t db = umma::make_desc(sb + 2 * j * b_sbo, 512, b_sbo, umma::LAYOUT_SW128_BASE32B);
            umma::mma_tf32_ss(tz, da, db, idesc1, (s0 > 0 || j > 0) ? 1u : 0u);
          }
          umma::mma_commit(&bar_empty[s]);
        }
        umma::mma_commit(&bar_zfull[a]);
        if (fused) {
          ok = umma::mbar_wait(&bar_zready[a], use & 1u) && ok;          // Z rounded in place by the epilogue warps
          umma::tc_fence_after_sync();
          for (int kb = 0; kb < nkb2; ++kb, ++g) {
            const uint32_t s = g % S;
            ok = umma::mbar_wait(&bar_full[s], (g / S) & 1u) && ok;
            umma::tc_fence_after_sync();
            const uint32_t sb = smem_base + s * stage_bytes + DT_A_BYTES;
#pragma unroll
            for (int j = 0; j < DT_KC / 8; ++j) {
              if (kb * DT_KC + j * 8 < D) {
                const uint64_t db = umma::make_desc_sw128(sb + j * 32, 16, 1024);
                dt::mma_tf32_ts(to, tz + (uint32_t)(kb * DT_KC + j * 8), db, idesc2, (kb > 0 || j > 0) ? 1u : 0u);
              }
            }
            umma::mma_commit(&bar_empty[s]);
          }
          umma::mma_commit(&bar_ofull[a]);
        }
      }
    }
    __syncwarp();
  } else if (wid < DT_WARP_EPI) {
    // ===================== expanders: the 0/1 operand tile of every stage (GATHER = 1: and the X rows) ==========
    constexpr int NE = DT_EXP_WARPS * 32;                 // 256 or 512 threads
    constexpr int KS = DT_KC * DT_SW_H / NE;              // condensed columns of its row a thread expands: 16 or 8
    const int et = tid - DT_WARP_EXP * 32;
    const int row = et & (DT_SW_H - 1), part = et >> 7;
    const int word = row >> 5, bit = row & 31;
    uint32_t a_off[KS / 4];
#pragma unroll
    for (int c = 0; c < KS / 4; ++c) a_off[c] = umma::kmajor_off(row, part * KS + c * 4);
    // GATHER = 1: the 16-byte pieces of the stage's 32 gathered rows this thread copies
    constexpr int B_PER = GATHER ? DT_KC * 64 / NE : 1;
    const int row_pieces = D / 4;
    int b_kr[B_PER], b_ch[B_PER];
    uint32_t b_off[B_PER];
    if (GATHER == 1) {
#pragma unroll
      for (int i = 0; i < B_PER; ++i) {
        const int pid = et + i * NE;
        if (pid < DT_KC * row_pieces) {
          b_kr[i] = pid / row_pieces;
          b_ch[i] = pid - b_kr[i] * row_pieces;
          b_off[i] = umma::mnmajor_chunk_off(b_kr[i], b_ch[i], 512, b_sbo);
        } else {
          b_kr[i] = -1; b_ch[i] = 0; b_off[i] = 0;
        }
      }
    }
    dt::Walk w = dt::walk_first(p);
    uint32_t g = 0, q = 0;
    while (w.valid) {
      ok = umma::mbar_wait(&bar_ifull[q & 1u], (q >> 1) & 1u) && ok;
      const int nst = w.ucols / DT_KC;
      const int f0 = w.ch * (DT_IDX / DT_KC), f1 = min(nst, f0 + DT_IDX / DT_KC);
      const int *cs = cols_s + (q & 1u) * DT_IDX;
      const unsigned *ms = masks_s + (size_t)(q & 1u) * DT_IDX * 4;
      for (int f = f0; f < f1; ++f, ++g) {
        const uint32_t s = g % S;
        const int kb0 = (f - f0) * DT_KC, kb = kb0 + part * KS;
        // the stage's masks and ids first (shared-memory broadcasts), then the ring slot
        float v[KS];
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          // (ids outside the operand need no masking here: their gathered rows are zero-filled)
          v[k] = ((ms[(kb + k) * 4 + word] >> bit) & 1u) ? 1.f : 0.f;
        }
        ok = umma::mbar_wait(&bar_empty[s], ((g / S) & 1u) ^ 1u) && ok;
        uint8_t *sa = gen + s * stage_bytes;
#pragma unroll
        for (int c = 0; c < KS / 4; ++c)
          *reinterpret_cast<float4 *>(sa + a_off[c]) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        umma::fence_proxy_async_smem();
        dt::arrive(&bar_full[s]);
        if (GATHER == 1) {
          uint8_t *sb = sa + DT_A_BYTES;
#pragma unroll
          for (int i = 0; i < B_PER; ++i) {
            if (b_kr[i] >= 0) {
              const int col = cs[kb0 + b_kr[i]];
              const bool valid = (unsigned)col < (unsigned)p.x_rows;
              const float *src = valid ? p.xr + (long long)col * D + b_ch[i] * 4 : p.xr;
              cp_async_16(sb + b_off[i], src, valid ? 16 : 0);
            }
          }
          dt_cp_async_arrive_noinc(&bar_full[s]);          // second arrival: when this thread's copies have landed
        }
      }
      dt::arrive(&bar_iempty[q & 1u]);
      ++q;
      if (fused && w.ch == w.nchunks - 1) {
        for (int kb = 0; kb < nkb2; ++kb, ++g) {           // phase 2 stages carry no 0/1 tile: keep the count only
          const uint32_t s = g % S;
          ok = umma::mbar_wait(&bar_empty[s], ((g / S) & 1u) ^ 1u) && ok;
          dt::arrive(&bar_full[s]);
          if (GATHER == 1) dt::arrive(&bar_full[s]);
        }
      }
      w = dt::walk_next(w, p);
    }
    if (GATHER == 1) asm volatile("cp.async.wait_all;" ::: "memory");
  } else {
    // ===================== epilogue =====================
    const int lq = wid & 3;                             // the TMEM lane quarter this warp may access
    uint32_t it = 0;
    for (int ti = blockIdx.x; ti < p.n_dense; ti += gridDim.x, ++it) {
      const int sw = __ldg(p.sw_ids + ti);
      const uint32_t a = p.nbuf == 2 ? (it & 1u) : 0u, use = p.nbuf == 2 ? (it >> 1) : it;
      const uint32_t tz = tmem_base + a * (uint32_t)p.acc_stride + ((uint32_t)(lq * 32) << 16);
      const uint32_t to = tz + (uint32_t)p.o_off;
      const int row = sw * DT_SW_H + lq * 32 + lane;
      ok = umma::mbar_wait(&bar_zfull[a], use & 1u) && ok;
      umma::tc_fence_after_sync();
      for (int cc = 0; cc < D; cc += 32) {
        uint32_t v[32];
        umma::tmem_ld_32x32(tz + (uint32_t)cc, v);
        umma::tmem_ld_wait();
        if (row < p.n_rows) {
          float *dst = p.z + (long long)row * p.ldz + cc;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (cc + j < D) {
              float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                     __uint_as_float(v[j + 3]));
              float4 *d4 = reinterpret_cast<float4 *>(dst + j);
              if (p.accumulate) add4(o, *d4);
              *d4 = o;
            }
          }
        }
        if (fused) {
          // the aggregate becomes the A operand of the Update product where it lies: round (cvt.rna, the
          // reference's :1809-1837) and write back to the same TMEM columns
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = f32_to_tf32(__uint_as_float(v[j]));
          dt::tmem_st_32x32(tz + (uint32_t)cc, v);
        }
      }
      if (fused) {
        dt::tmem_st_wait();
        umma::tc_fence_before_sync();
        dt::arrive(&bar_zready[a]);
        ok = umma::mbar_wait(&bar_ofull[a], use & 1u) && ok;
        umma::tc_fence_after_sync();
        for (int cc = 0; cc < p.hidden; cc += 32) {
          uint32_t v[32];
          umma::tmem_ld_32x32(to + (uint32_t)cc, v);
          umma::tmem_ld_wait();
          if (row < p.n_rows) {
            float *dst = p.out + (long long)row * p.ldo + cc;
            if ((p.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (cc + j + 3 < p.hidden) {
                  *reinterpret_cast<float4 *>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                } else {
#pragma unroll
                  for (int t = 0; t < 4; ++t)
                    if (cc + j + t < p.hidden) dst[j + t] = __uint_as_float(v[j + t]);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cc + j < p.hidden) dst[j] = __uint_as_float(v[j]);
            }
          }
        }
      }
      umma::tc_fence_before_sync();
      dt::arrive(&bar_free[a]);
    }
  }
  if (!ok && p.err) atomicExch(p.err, 1);
  umma::tc_fence_before_sync();
  __syncthreads();
  if (wid == DT_WARP_MMA) umma::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// W [k, n] row-major -> W^T rounded to TF32, [n_pad][k_pad] K-major (update_gemm.cu)
void launch_update_gemm_wt(const float *w, int64_t ldw, int k, int n, int k_pad, int n_pad, float *wt, cudaStream_t stream);

bool dense_tma_supported(const float *x, int64_t ldx, const float *z, int64_t ldz, int32_t dim) {
  return dim >= 16 && dim <= 256 && (dim % 16) == 0 && (ldx & 3) == 0 && (ldz & 3) == 0 &&
         ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(z)) & 15) == 0;
}

size_t dense_tma_scratch_floats(int32_t dim, int32_t hidden) {
  if (hidden <= 0) return 0;
  const size_t k_pad = ((size_t)dim + DT_KC - 1) / DT_KC * DT_KC, hp = ((size_t)hidden + 15) / 16 * 16;
  return k_pad * hp;
}

// Z rows of the plan's super-windows (+)= A X; with hidden > 0 also out = rna(Z) rna(W) for those rows (hidden <= 256).
// xr_scratch != nullptr selects GATHER = 1: X is first rounded to TF32 into xr_scratch ([x_rows * dim] floats) and
// gathered with cp.async; nullptr selects the TMA gather4 variant (no copy of X).
void launch_tf32_round_rows(const float *x, int64_t ldx, int32_t x_rows, int32_t dim, float *xr, cudaStream_t stream);

int launch_spmm_dense_tma(const float *x, int64_t ldx, int32_t x_rows, int32_t n_rows, int32_t dim, const int *sw_ids,
                          const int *sw_off, const int *cols, const unsigned *masks, int32_t n_dense, int accumulate,
                          float *z, int64_t ldz, const float *w, int64_t ldw, int32_t hidden, float *out, int64_t ldo,
                          float *wt_scratch, float *xr_scratch, int *d_err, cudaStream_t stream) {
  if (n_dense <= 0) return 0;
  DenseTmaParams p;
  p.xr = xr_scratch;
  const int gather = xr_scratch != nullptr ? 1 : 0;
  if (gather) launch_tf32_round_rows(x, ldx, x_rows, dim, xr_scratch, stream);
  p.x_rows = x_rows; p.dim = dim; p.n_rows = n_rows; p.n_dense = n_dense; p.accumulate = accumulate;
  p.sw_ids = sw_ids; p.sw_off = sw_off; p.cols = cols; p.masks = masks; p.z = z; p.ldz = ldz;
  p.hidden = hidden > 0 ? hidden : 0;
  p.hp = hidden > 0 ? (hidden + 15) / 16 * 16 : 16;
  p.out = out; p.ldo = ldo; p.err = d_err;
  if (p.hidden > 256 || (p.hidden > 0 && (accumulate || !w || !out || !wt_scratch))) {
    set_error("spmm_dense_tma: fused update needs hidden <= 256, no accumulate, W / out / scratch");
    return HCSPMM_E_UNSUPPORTED;
  }
  const int dz = (dim + 31) / 32 * 32, dh = p.hidden > 0 ? (p.hp + 31) / 32 * 32 : 0;
  p.o_off = dz;
  p.acc_stride = dz + dh;
  p.nbuf = 2 * p.acc_stride <= 512 ? 2 : 1;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.nbuf * p.acc_stride) p.tmem_cols <<= 1;
  const int natoms = (dim + 31) / 32;
  const uint32_t b_bytes = (DT_KC / 4) * natoms * 512u, w_bytes = p.hidden > 0 ? (uint32_t)p.hp * 128u : 0u;
  const uint32_t bw = b_bytes > w_bytes ? b_bytes : w_bytes;
  const uint32_t stage_bytes = DT_A_BYTES + ((bw + 1023u) & ~1023u);
  const uint32_t idx_bytes = 2 * DT_IDX * 20;
  int stages = (int)((232448u - 2048u - 1024u - idx_bytes) / stage_bytes);
  if (stages > DT_MAX_STAGES) stages = DT_MAX_STAGES;
  if (tuning().gemm_stages > 0 && tuning().gemm_stages < stages) stages = tuning().gemm_stages;
  if (stages < 2) { set_error("spmm_dense_tma: stage does not fit shared memory"); return HCSPMM_E_UNSUPPORTED; }
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + idx_bytes + 1024;
  CUtensorMap tm_x, tm_w;
  // X: TFLOAT32-typed (the TMA unit rounds on load), one row x 32 columns per gathered row, 32-byte-atom swizzle
  if (!make_tensor_map_2d(&tm_x, x, (uint64_t)dim, (uint64_t)x_rows, (uint64_t)ldx, 32, 1, 2, 1)) {
    set_error("spmm_dense_tma: cuTensorMapEncodeTiled (X) failed");
    return HCSPMM_E_UNSUPPORTED;
  }
  if (p.hidden > 0) {
    const int k_pad = (dim + DT_KC - 1) / DT_KC * DT_KC;
    launch_update_gemm_wt(w, ldw, dim, hidden, k_pad, p.hp, wt_scratch, stream);
    if (!make_tensor_map_2d(&tm_w, wt_scratch, (uint64_t)k_pad, (uint64_t)p.hp, (uint64_t)k_pad, DT_KC, (uint32_t)p.hp, 1, 0)) {
      set_error("spmm_dense_tma: cuTensorMapEncodeTiled (W) failed");
      return HCSPMM_E_UNSUPPORTED;
    }
  } else {
    tm_w = tm_x;
  }
  cudaError_t err = gather ? cudaFuncSetAttribute(spmm_dense_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                           : cudaFuncSetAttribute(spmm_dense_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) { set_error("spmm_dense_tma attr: %s", cudaGetErrorString(err)); return (int)err; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = n_dense < sms ? n_dense : sms;
  if (gather) spmm_dense_tma_kernel<1><<<grid, dt_threads(1), smem, stream>>>(tm_x, tm_w, p);
  else spmm_dense_tma_kernel<0><<<grid, dt_threads(0), smem, stream>>>(tm_x, tm_w, p);
  err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("spmm_dense_tma launch: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

}  // namespace hcspmm
