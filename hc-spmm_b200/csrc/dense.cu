// dense.cu -- the tcgen05 / TMEM dense path: "super-windows" of 128 rows whose 16-row windows are all
// labelled tensor-core are multiplied as ONE dense contraction per super-window,
//
//     Y[128 rows, D] = Abits[128, U] * X[cols[0..U), D]          U = distinct columns of the 128 rows
//
// with tcgen05.mma (M = 128, N = D <= 256, K = 8, TF32, FP32 accumulate in TMEM).  Stacking eight
// windows is what makes the tensor core worth feeding on B200: every distinct X row is gathered once
// per 128 rows instead of once per 16 (community-structured graphs share most columns across
// neighbouring windows), and the 0/1 operand is free.
//
//   plan (built once, at preprocess time, from the reference's own arrays + one 128-row rank pass):
//       sw_ids[n_dense]        super-window ids that take this path
//       sw_off[n_dense + 1]    offsets into cols / masks (each super-window padded to 32 columns)
//       cols[total]            condensed column -> X row (-1 = padding)
//       masks[total][4]        128-bit row mask of the condensed column
//       ht2[W]                 window labels for the hybrid kernel: 2 = "handled here, skip"
//   kernel (persistent CTAs, 256 threads):
//       B tile  32 gathered X rows x D   cp.async 16-byte copies, MN-major SWIZZLE_128B_BASE32B
//       A tile  128 rows x 32 columns    expanded from the bit masks, K-major SWIZZLE_128B
//       NST-stage ring; one thread issues 4 MMAs per stage and commits to the stage's mbarrier, which
//       gates the refill; epilogue tcgen05.ld -> global.
//   X is first rounded to TF32 (cvt.rna) into a scratch copy so that the tensor core's truncation
//   reproduces the reference's rounding (hybrid_all_kernel.cu:1102-1109); the copy costs
//   2 * x_rows * D * 4 bytes of streaming traffic and is only made when the plan is non-empty.
#include "common.cuh"
#include "umma.cuh"

namespace hcspmm {

constexpr int SW_H = 128;        // rows per super-window
constexpr int DN_THREADS = 256;
constexpr int DN_KC = 32;        // condensed columns per pipeline stage
constexpr int DN_STAGES = 4;
constexpr int PLAN_HEADER = 16;
constexpr int PLAN_MAGIC = 0x48435044;  // "HCPD"

int launch_preprocess_super(const int32_t *colidx, const int32_t *rowptr, int32_t n_rows, int32_t n_super,
                            int32_t *bp128, int32_t *etc128, int32_t *ht_scratch, void *ws, cudaStream_t stream);
size_t preprocess_workspace_bytes(int32_t n_rows, int64_t nnz);

static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

// ---- plan: selection ---------------------------------------------------------------------------
struct PlanScratch {
  int32_t *etc128;   // [nnz] rank of each entry's column inside its super-window
  int32_t *bp128;    // [SW] ceil(U / 8)
  int32_t *ht_tmp;   // [SW]
  int32_t *sw_slot;  // [SW] dense index or -1
  int32_t *sw_ids;   // [SW]
  int32_t *sw_off;   // [SW + 1]
  int32_t *ht2;      // [W]
  int32_t *counts;   // [2]
  void *pre_ws;
};

static PlanScratch carve(void *ws, int32_t n_rows, int64_t nnz) {
  const size_t sw = ((size_t)n_rows + SW_H - 1) / SW_H, w = ((size_t)n_rows + BLK_H - 1) / BLK_H;
  char *p = reinterpret_cast<char *>(ws);
  PlanScratch s;
  s.etc128 = reinterpret_cast<int32_t *>(p); p += al((size_t)(nnz > 0 ? nnz : 1) * 4);
  s.bp128 = reinterpret_cast<int32_t *>(p); p += al(sw * 4);
  s.ht_tmp = reinterpret_cast<int32_t *>(p); p += al(sw * 4);
  s.sw_slot = reinterpret_cast<int32_t *>(p); p += al(sw * 4);
  s.sw_ids = reinterpret_cast<int32_t *>(p); p += al(sw * 4);
  s.sw_off = reinterpret_cast<int32_t *>(p); p += al((sw + 1) * 4);
  s.ht2 = reinterpret_cast<int32_t *>(p); p += al(w * 4);
  s.counts = reinterpret_cast<int32_t *>(p); p += 256;
  s.pre_ws = p;
  return s;
}

size_t dense_plan_workspace_bytes(int32_t n_rows, int64_t nnz) {
  const size_t sw = ((size_t)n_rows + SW_H - 1) / SW_H, w = ((size_t)n_rows + BLK_H - 1) / BLK_H;
  return al((size_t)(nnz > 0 ? nnz : 1) * 4) + 4 * al(sw * 4) + al((sw + 1) * 4) + al(w * 4) + 256 +
         preprocess_workspace_bytes(n_rows, nnz) + 256;
}

// one CTA: decide per super-window, then scan
__global__ void __launch_bounds__(1024) dense_select_kernel(const int *__restrict__ rowptr, const int *__restrict__ ht,
                                                            const int *__restrict__ bp128, int n_rows, int n_windows,
                                                            int n_super, int min_reuse_x2, int min_rowlen, int *sw_slot, int *sw_ids,
                                                            int *sw_off, int *ht2, int *counts) {
  __shared__ int wbuf[32];
  __shared__ int s_carry_n, s_carry_c;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) { s_carry_n = 0; s_carry_c = 0; }
  __syncthreads();
  for (int base = 0; base < n_super; base += 1024) {
    const int sw = base + tid;
    int dense = 0, ucols = 0;
    if (sw < n_super) {
      const int r0 = sw * SW_H, r1 = min(r0 + SW_H, n_rows);
      const int e = rowptr[r1] - rowptr[r0];
      bool all_tc = e > 0;
      for (int w = sw * 8; w < min(sw * 8 + 8, n_windows); ++w) {
        const int wr0 = w * BLK_H, wr1 = min(wr0 + BLK_H, n_rows);
        const bool empty = rowptr[wr1] == rowptr[wr0];
        all_tc = all_tc && (empty || ht[w] != 0);
      }
      ucols = (bp128[sw] * BLK_W + DN_KC - 1) / DN_KC * DN_KC;
      // B200 re-fit of the selector at super-window granularity (benchmarks/selector_fit.py): the tcgen05 path wins
      // when the rows are long enough on average (>= 8 entries: 97 % of the measured cells) -- the CUDA-core path
      // is latency-bound on short rows, the dense contraction is indifferent to them -- and the columns are shared
      // (reuse >= min_reuse: executed flops grow with the distinct columns, useful ones with the entries)
      dense = all_tc && ucols > 0 && (2LL * e >= (long long)min_reuse_x2 * ucols) &&
              (min_reuse_x2 == 0 || e >= (long long)min_rowlen * (r1 - r0));
    }
    // two block scans: dense index, column offset
    int vals[2] = {dense, dense ? ucols : 0}, exc[2], tot[2];
    for (int k = 0; k < 2; ++k) {
      int inc = vals[k];
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      __syncthreads();
      if (lane == 31) wbuf[wid] = inc;
      __syncthreads();
      int ws = wbuf[lane], wi = ws;
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
      tot[k] = __shfl_sync(0xffffffffu, wi, 31);
      exc[k] = __shfl_sync(0xffffffffu, wi - ws, wid) + inc - vals[k];
    }
    const int cn = s_carry_n, cc = s_carry_c;
    if (sw < n_super) {
      sw_slot[sw] = dense ? cn + exc[0] : -1;
      if (dense) { sw_ids[cn + exc[0]] = sw; sw_off[cn + exc[0]] = cc + exc[1]; }
      for (int w = sw * 8; w < min(sw * 8 + 8, n_windows); ++w) ht2[w] = dense ? 2 : ht[w];
    }
    __syncthreads();
    if (tid == 0) { s_carry_n = cn + tot[0]; s_carry_c = cc + tot[1]; }
    __syncthreads();
  }
  if (tid == 0) { sw_off[s_carry_n] = s_carry_c; counts[0] = s_carry_n; counts[1] = s_carry_c; }
}

__global__ void dense_fill_kernel(const int *__restrict__ colidx, const int *__restrict__ etr,
                                  const int *__restrict__ etc128, const int *__restrict__ sw_slot,
                                  const int *__restrict__ sw_off, long long nnz, int *__restrict__ cols,
                                  unsigned *__restrict__ masks) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (long long)gridDim.x * blockDim.x) {
    const int r = __ldg(etr + e);
    const int slot = __ldg(sw_slot + r / SW_H);
    if (slot < 0) continue;
    const int pos = __ldg(sw_off + slot) + __ldg(etc128 + e);
    cols[pos] = __ldg(colidx + e);
    const int rl = r % SW_H;
    atomicOr(&masks[4 * (long long)pos + (rl >> 5)], 1u << (rl & 31));
  }
}

size_t dense_plan_words(int32_t n_rows, int32_t n_dense, int64_t total_cols) {
  const size_t w = ((size_t)n_rows + BLK_H - 1) / BLK_H;
  size_t words = PLAN_HEADER + (size_t)n_dense + (size_t)n_dense + 1 + w;
  words = (words + 3) & ~(size_t)3;
  return words + (size_t)total_cols * 5;
}

struct PlanView {
  const int *sw_ids, *sw_off, *ht2, *cols;
  const unsigned *masks;
};
static PlanView view_plan(const int32_t *plan, int32_t n_rows, int32_t n_dense, int64_t total_cols) {
  const size_t w = ((size_t)n_rows + BLK_H - 1) / BLK_H;
  PlanView v;
  v.sw_ids = plan + PLAN_HEADER;
  v.sw_off = v.sw_ids + n_dense;
  v.ht2 = v.sw_off + n_dense + 1;
  size_t words = PLAN_HEADER + (size_t)n_dense + (size_t)n_dense + 1 + w;
  words = (words + 3) & ~(size_t)3;
  v.cols = plan + words;
  v.masks = reinterpret_cast<const unsigned *>(v.cols + total_cols);
  return v;
}

int dense_plan_count(const int32_t *colidx, const int32_t *rowptr, const int32_t *ht, int32_t n_rows, int64_t nnz,
                     int min_reuse_x2, void *ws, size_t ws_bytes, int32_t *h_counts, cudaStream_t stream) {
  h_counts[0] = h_counts[1] = 0;
  if (n_rows <= 0 || nnz <= 0) return 0;
  if (!colidx || !rowptr || !ht || !ws) { set_error("dense_plan: null pointer argument"); return HCSPMM_E_INVALID; }
  if (ws_bytes < dense_plan_workspace_bytes(n_rows, nnz)) { set_error("dense_plan: workspace too small"); return HCSPMM_E_WORKSPACE; }
  const int n_super = (n_rows + SW_H - 1) / SW_H, n_windows = (n_rows + BLK_H - 1) / BLK_H;
  PlanScratch s = carve(ws, n_rows, nnz);
  int rc = launch_preprocess_super(colidx, rowptr, n_rows, n_super, s.bp128, s.etc128, s.ht_tmp, s.pre_ws, stream);
  if (rc) return rc;
  dense_select_kernel<<<1, 1024, 0, stream>>>(rowptr, ht, s.bp128, n_rows, n_windows, n_super, min_reuse_x2,
                                              tuning().dense_min_rowlen, s.sw_slot,
                                              s.sw_ids, s.sw_off, s.ht2, s.counts);
  cudaError_t err = cudaGetLastError();
  if (err == cudaSuccess) err = cudaMemcpyAsync(h_counts, s.counts, 8, cudaMemcpyDeviceToHost, stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(stream);
  if (err != cudaSuccess) { set_error("dense_plan_count: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

int dense_plan_fill(const int32_t *colidx, const int32_t *etr, int32_t n_rows, int64_t nnz, void *ws, int32_t n_dense,
                    int64_t total_cols, int32_t *plan, size_t plan_words, cudaStream_t stream) {
  if (plan_words < dense_plan_words(n_rows, n_dense, total_cols)) { set_error("dense_plan_fill: plan buffer too small"); return HCSPMM_E_WORKSPACE; }
  PlanScratch s = carve(ws, n_rows, nnz);
  const size_t w = ((size_t)n_rows + BLK_H - 1) / BLK_H;
  const int n_super = (n_rows + SW_H - 1) / SW_H;
  int32_t header[PLAN_HEADER] = {PLAN_MAGIC, n_dense, (int32_t)total_cols, n_super, (int32_t)w, n_rows};
  PlanView v = view_plan(plan, n_rows, n_dense, total_cols);
  cudaError_t err = cudaMemcpyAsync(plan, header, sizeof(header), cudaMemcpyHostToDevice, stream);
  if (err == cudaSuccess && n_dense > 0)
    err = cudaMemcpyAsync(const_cast<int *>(v.sw_ids), s.sw_ids, (size_t)n_dense * 4, cudaMemcpyDeviceToDevice, stream);
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(const_cast<int *>(v.sw_off), s.sw_off, ((size_t)n_dense + 1) * 4, cudaMemcpyDeviceToDevice, stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(const_cast<int *>(v.ht2), s.ht2, w * 4, cudaMemcpyDeviceToDevice, stream);
  if (err == cudaSuccess && total_cols > 0) err = cudaMemsetAsync(const_cast<int *>(v.cols), 0xff, (size_t)total_cols * 4, stream);
  if (err == cudaSuccess && total_cols > 0) err = cudaMemsetAsync(const_cast<unsigned *>(v.masks), 0, (size_t)total_cols * 16, stream);
  if (err != cudaSuccess) { set_error("dense_plan_fill: %s", cudaGetErrorString(err)); return (int)err; }
  if (n_dense > 0) {
    dense_fill_kernel<<<1184, 256, 0, stream>>>(colidx, etr, s.etc128, s.sw_slot, s.sw_off, nnz, const_cast<int *>(v.cols),
                                                const_cast<unsigned *>(v.masks));
    err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaStreamSynchronize(stream);   // the header array lives on this stack frame
    if (err != cudaSuccess) { set_error("dense_fill: %s", cudaGetErrorString(err)); return (int)err; }
  } else {
    cudaStreamSynchronize(stream);
  }
  return 0;
}

// ---- X -> TF32 (cvt.rna) scratch copy ----------------------------------------------------------------
__global__ void tf32_round_rows_kernel(const float *__restrict__ x, long long ldx, int dim4, float *__restrict__ xr,
                                       long long total4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / dim4;
    const int c = (int)(i - r * dim4);
    float4 v = ldg_f4(x + r * ldx + (long long)c * 4);
    v.x = __uint_as_float(f32_to_tf32(v.x));
    v.y = __uint_as_float(f32_to_tf32(v.y));
    v.z = __uint_as_float(f32_to_tf32(v.z));
    v.w = __uint_as_float(f32_to_tf32(v.w));
    *reinterpret_cast<float4 *>(xr + i * 4) = v;
  }
}

void launch_tf32_round_rows(const float *x, int64_t ldx, int32_t x_rows, int32_t dim, float *xr, cudaStream_t stream) {
  const long long total4 = (long long)x_rows * (dim / 4);
  tf32_round_rows_kernel<<<1184, 256, 0, stream>>>(x, ldx, dim / 4, xr, total4);
}

// ---- the dense kernel ------------------------------------------------------------------------------
struct DenseParams {
  const float *xr;   // TF32-rounded X, dense [x_rows, dim]
  int x_rows, dim, n_rows, n_dense, accumulate;
  const int *sw_ids, *sw_off, *cols;
  const unsigned *masks;
  float *y;
  long long ldy;
  int *err;
};

__global__ void __launch_bounds__(DN_THREADS, 1) spmm_dense_kernel(const DenseParams p) {
  extern __shared__ __align__(1024) uint8_t dn_smem[];
  __shared__ __align__(8) uint64_t bar_empty[DN_STAGES];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int D = p.dim;                       // multiple of 16, <= 256
  const int natoms = (D + 31) / 32;
  const uint32_t a_bytes = SW_H * 128;       // 16 KB: 128 rows x 32 tf32
  const uint32_t b_lbo = 512, b_sbo = natoms * 512;
  const uint32_t b_bytes = (DN_KC / 4) * b_sbo;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t smem_base = (umma::smem_u32(dn_smem) + 1023u) & ~1023u;
  uint8_t *gen = dn_smem + (smem_base - umma::smem_u32(dn_smem));

  if (tid == 0) {
    for (int s = 0; s < DN_STAGES; ++s) umma::mbar_init(&bar_empty[s], 1);
    umma::mbar_init(&bar_done, 1);
    umma::fence_barrier_init();
  }
  if (wid == 0) umma::tmem_alloc(&tmem_slot, 256);
  // zero the B buffers once: tails of a partial last n-atom are never written again
  for (uint32_t o = tid * 16; o < DN_STAGES * stage_bytes; o += DN_THREADS * 16)
    *reinterpret_cast<float4 *>(gen + o) = make_float4(0.f, 0.f, 0.f, 0.f);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_slot;
  const uint32_t idesc = umma::make_idesc_tf32(SW_H, D, 0, 1);
  const int row_pieces = D / 4;              // 16-byte pieces per X row
  bool ok = true;
  // per-thread piece map (depends on tid and D only): B pieces tid + i * 256 -> (k-row, 16-byte chunk)
  constexpr int B_PER = DN_KC * 64 / DN_THREADS;   // 8 pieces per thread at D = 256
  int b_kr[B_PER], b_ch[B_PER], rc[B_PER];
  uint32_t b_off[B_PER], a_off[4], rm[16];
#pragma unroll
  for (int i = 0; i < B_PER; ++i) {
    const int pid = tid + i * DN_THREADS;
    if (pid < DN_KC * row_pieces) {
      b_kr[i] = pid / row_pieces;
      b_ch[i] = pid - b_kr[i] * row_pieces;
      b_off[i] = umma::mnmajor_chunk_off(b_kr[i], b_ch[i], b_lbo, b_sbo);
    } else {
      b_kr[i] = -1; b_ch[i] = 0; b_off[i] = 0;
    }
    rc[i] = -1;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) a_off[c] = umma::kmajor_off(tid & (SW_H - 1), (tid >> 7) * 16 + c * 4);
#pragma unroll
  for (int k = 0; k < 16; ++k) rm[k] = 0u;
  uint32_t g = 0;                            // global stage counter (drives buffer index and barrier parity)
  uint32_t tile_iter = 0;

  for (int ti = blockIdx.x; ti < p.n_dense; ti += gridDim.x, ++tile_iter) {
    const int sw = __ldg(p.sw_ids + ti);
    const int c0 = __ldg(p.sw_off + ti);
    const int nst = (__ldg(p.sw_off + ti + 1) - c0) / DN_KC;
    const uint32_t g0 = g;

    // Index prefetch: the column ids and row masks of the NEXT stage to be filled are loaded into
    // registers one pipeline step ahead, so their L2 latency is off the critical path.
    auto load_idx = [&](int s) {
      const int cbase = c0 + s * DN_KC;
#pragma unroll
      for (int i = 0; i < B_PER; ++i) rc[i] = (b_kr[i] >= 0) ? __ldg(p.cols + cbase + b_kr[i]) : -1;
      const int word = (tid & (SW_H - 1)) >> 5, kh = (tid >> 7) * 16;
#pragma unroll
      for (int k = 0; k < 16; ++k) rm[k] = __ldg(p.masks + 4 * (long long)(cbase + kh + k) + word);
    };
    auto fill = [&](uint32_t buf) {
      uint8_t *sa = gen + buf * stage_bytes, *sb = sa + a_bytes;
      // B: gathered rows, 16-byte cp.async pieces into the swizzled MN-major tile
#pragma unroll
      for (int i = 0; i < B_PER; ++i) {
        if (b_kr[i] >= 0) {
          const bool valid = (unsigned)rc[i] < (unsigned)p.x_rows;
          const float *src = valid ? p.xr + (long long)rc[i] * D + b_ch[i] * 4 : p.xr;
          cp_async_16(sb + b_off[i], src, valid ? 16 : 0);
        }
      }
      // A: thread (row r, half h) expands 16 condensed columns of its row from the bit masks
      const int bit = tid & 31;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 v = make_float4((float)((rm[4 * c] >> bit) & 1u), (float)((rm[4 * c + 1] >> bit) & 1u),
                                     (float)((rm[4 * c + 2] >> bit) & 1u), (float)((rm[4 * c + 3] >> bit) & 1u));
        *reinterpret_cast<float4 *>(sa + a_off[c]) = v;
      }
    };

    load_idx(0);
#pragma unroll
    for (int s0 = 0; s0 < DN_STAGES - 1; ++s0) {
      if (s0 < nst) {
        fill((g0 + s0) % DN_STAGES);
        if (s0 + 1 < nst) load_idx(s0 + 1);
      }
      cp_async_commit();
    }
    for (int s = 0; s < nst; ++s) {
      const uint32_t gs = g0 + s, buf = gs % DN_STAGES;
      cp_async_wait<DN_STAGES - 2>();
      umma::fence_proxy_async_smem();
      __syncthreads();                        // stage s complete in shared memory for every thread
      if (tid == 0) {
        umma::tc_fence_after_sync();
        const uint32_t sa = smem_base + buf * stage_bytes, sb = sa + a_bytes;
#pragma unroll
        for (int j = 0; j < DN_KC / 8; ++j) {
          const uint64_t da = umma::make_desc_sw128(sa + j * 32, 16, 1024);
          const uint64_t db = umma::make_desc(sb + 2 * j * b_sbo, b_lbo, b_sbo, umma::LAYOUT_SW128_BASE32B);
          umma::mma_tf32_ss(tmem_d, da, db, idesc, (s > 0 || j > 0) ? 1u : 0u);
        }
        umma::mma_commit(&bar_empty[buf]);
        if (s == nst - 1) umma::mma_commit(&bar_done);
      }
      // refill the buffer the previous stage used, once its MMAs have drained
      const int nxt = s + DN_STAGES - 1;
      if (nxt < nst) {
        if (s >= 1) {
          const uint32_t gp = gs - 1;
          ok = umma::mbar_wait(&bar_empty[gp % DN_STAGES], (gp / DN_STAGES) & 1) && ok;
        }
        fill((g0 + nxt) % DN_STAGES);
        if (nxt + 1 < nst) load_idx(nxt + 1);
      }
      cp_async_commit();
    }
    g = g0 + nst;
    cp_async_wait<0>();

    // epilogue: accumulator -> registers -> Y
    ok = umma::mbar_wait(&bar_done, tile_iter & 1) && ok;
    umma::tc_fence_after_sync();
    {
      const int lq = wid & 3, half = wid >> 2;
      const int row = sw * SW_H + lq * 32 + lane;
      for (int cc = half * 128; cc < half * 128 + 128 && cc < D; cc += 32) {
        uint32_t v[32];
        umma::tmem_ld_32x32(tmem_d + ((uint32_t)(lq * 32) << 16) + (uint32_t)cc, v);
        umma::tmem_ld_wait();
        if (row < p.n_rows) {
          float *dst = p.y + (long long)row * p.ldy + cc;
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
      }
    }
    // every warp has drained TMEM before the next tile's first MMA overwrites it; every stage's
    // MMAs of this tile are complete (bar_done), so all buffers are free for the next prologue
    umma::tc_fence_before_sync();
    __syncthreads();
    umma::tc_fence_after_sync();
  }
  if (!ok && p.err) atomicExch(p.err, 1);
  umma::tc_fence_before_sync();
  __syncthreads();
  if (wid == 0) umma::tmem_dealloc(tmem_d, 256);
}

// ---- warp-specialised version ---------------------------------------------------------------------
// 9 warps: warps 0-7 (256 threads) PRODUCE stages (cp.async gathers + mask expansion) and drain the
// accumulator; warp 8 ISSUES the MMAs.  No block barrier in the steady state -- four kinds of mbarrier:
//   full[s]       256 producer arrivals (cp.async.mbarrier.arrive.noinc: fires when the thread's copies
//                 have landed)                                   producer -> issuer
//   empty[s]      tcgen05.commit of the stage's MMAs            issuer  -> producer
//   tmem_full[a]  tcgen05.commit after a super-window's last MMA issuer  -> epilogue
//   tmem_empty[a] 256 arrivals after the accumulator was read    epilogue -> issuer
// The accumulator is double buffered in TMEM (2 x 256 columns), so the issuer starts the next
// super-window while the previous one is being written out.
#ifndef HCSPMM_DW_PRODUCERS
#define HCSPMM_DW_PRODUCERS 512
#endif
constexpr int DW_PRODUCERS = HCSPMM_DW_PRODUCERS;
constexpr int DW_THREADS = DW_PRODUCERS + 32;
constexpr int DW_STAGES = 3;
constexpr int DW_IDX = 1024;   // condensed columns per shared-memory index chunk (20 KB, double buffered)

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(DW_THREADS, 1) spmm_dense_ws_kernel(const DenseParams p) {
  extern __shared__ __align__(1024) uint8_t dn_smem[];
  __shared__ __align__(8) uint64_t bar_full[DW_STAGES], bar_empty[DW_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int D = p.dim;
  const int natoms = (D + 31) / 32;
  const uint32_t a_bytes = SW_H * 128;
  const uint32_t b_lbo = 512, b_sbo = natoms * 512;
  const uint32_t b_bytes = (DN_KC / 4) * b_sbo;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t smem_base = (umma::smem_u32(dn_smem) + 1023u) & ~1023u;
  uint8_t *gen = dn_smem + (smem_base - umma::smem_u32(dn_smem));

  if (tid == 0) {
    for (int s = 0; s < DW_STAGES; ++s) { umma::mbar_init(&bar_full[s], 2 * DW_PRODUCERS); umma::mbar_init(&bar_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { umma::mbar_init(&bar_tfull[a], 1); umma::mbar_init(&bar_tempty[a], DW_PRODUCERS); }
    umma::fence_barrier_init();
  }
  if (wid == 0) umma::tmem_alloc(&tmem_slot, 512);
  for (uint32_t o = tid * 16; o < DW_STAGES * stage_bytes; o += DW_THREADS * 16)
    *reinterpret_cast<float4 *>(gen + o) = make_float4(0.f, 0.f, 0.f, 0.f);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  bool ok = true;

  if (wid == DW_PRODUCERS / 32) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      const uint32_t idesc = umma::make_idesc_tf32(SW_H, D, 0, 1);
      uint32_t g = 0, it = 0;
      for (int ti = blockIdx.x; ti < p.n_dense; ti += gridDim.x, ++it) {
        const int nst = (__ldg(p.sw_off + ti + 1) - __ldg(p.sw_off + ti)) / DN_KC;
        const uint32_t acc = it & 1, use = it >> 1;
        ok = umma::mbar_wait(&bar_tempty[acc], (use & 1) ^ 1) && ok;      // accumulator drained (first use: free)
        umma::tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * 256;
        for (int s = 0; s < nst; ++s) {
          const uint32_t gs = g + s, buf = gs % DW_STAGES;
          ok = umma::mbar_wait(&bar_full[buf], (gs / DW_STAGES) & 1) && ok;
          // consumer-side proxy fence: the producers' generic-proxy writes (st.shared, cp.async), made
          // visible to this thread by the mbarrier, are ordered before the tensor core's async-proxy reads
          umma::fence_proxy_async_smem();
          umma::tc_fence_after_sync();
          const uint32_t sa = smem_base + buf * stage_bytes, sb = sa + a_bytes;
#pragma unroll
          for (int j = 0; j < DN_KC / 8; ++j) {
            const uint64_t da = umma::make_desc_sw128(sa + j * 32, 16, 1024);
            const uint64_t db = umma::make_desc(sb + 2 * j * b_sbo, b_lbo, b_sbo, umma::LAYOUT_SW128_BASE32B);
            umma::mma_tf32_ss(tmem_d, da, db, idesc, (s > 0 || j > 0) ? 1u : 0u);
          }
          umma::mma_commit(&bar_empty[buf]);
        }
        umma::mma_commit(&bar_tfull[acc]);
        g += nst;
      }
    }
    __syncwarp();   // lanes 1..31 wait here, so the whole warp reaches the final block barrier together
  } else {
    // ================= producers + epilogue (256 threads) =================
    const int row_pieces = D / 4;
    constexpr int B_PER = DN_KC * 64 / DW_PRODUCERS;       // 16-byte B pieces per thread at D = 256
    constexpr int A_KS = DN_KC * SW_H / DW_PRODUCERS;      // condensed columns of its row a thread expands
    constexpr int A_PER = A_KS / 4;
    int b_kr[B_PER], b_ch[B_PER];
    uint32_t b_off[B_PER], a_off[A_PER];
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int pid = tid + i * DW_PRODUCERS;
      if (pid < DN_KC * row_pieces) {
        b_kr[i] = pid / row_pieces;
        b_ch[i] = pid - b_kr[i] * row_pieces;
        b_off[i] = umma::mnmajor_chunk_off(b_kr[i], b_ch[i], b_lbo, b_sbo);
      } else {
        b_kr[i] = -1; b_ch[i] = 0; b_off[i] = 0;
      }
    }
#pragma unroll
    for (int c = 0; c < A_PER; ++c) a_off[c] = umma::kmajor_off(tid & (SW_H - 1), (tid >> 7) * A_KS + c * 4);
    const int word = (tid & (SW_H - 1)) >> 5, kh = (tid >> 7) * A_KS, bit = tid & 31;

    // Column ids and row masks are staged in shared memory in chunks of DW_IDX condensed columns
    // (double buffered, cp.async): the stage loop below never waits on a global index load.
    int *cols_s = reinterpret_cast<int *>(gen + DW_STAGES * stage_bytes);
    unsigned *masks_s = reinterpret_cast<unsigned *>(cols_s + 2 * DW_IDX);
    auto load_chunk = [&](int cbase, int ncols, int slot) {
      // ncols is a multiple of 32: cols = ncols / 4 pieces of 16 B, masks = ncols pieces of 16 B
      for (int i = tid; i < ncols / 4; i += DW_PRODUCERS) cp_async_16(cols_s + slot * DW_IDX + i * 4, p.cols + cbase + i * 4, 16);
      for (int i = tid; i < ncols; i += DW_PRODUCERS)
        cp_async_16(masks_s + (size_t)(slot * DW_IDX + i) * 4, p.masks + 4 * (long long)(cbase + i), 16);
    };
    auto producers_sync = [&]() {
      asm volatile("cp.async.wait_all;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(DW_PRODUCERS) : "memory");
    };

    uint32_t g = 0, it = 0;
    for (int ti = blockIdx.x; ti < p.n_dense; ti += gridDim.x, ++it) {
      const int sw = __ldg(p.sw_ids + ti);
      const int c0 = __ldg(p.sw_off + ti);
      const int ucols = __ldg(p.sw_off + ti + 1) - c0;
      const int nst = ucols / DN_KC;
      const int nchunks = (ucols + DW_IDX - 1) / DW_IDX;
      producers_sync();                       // previous super-window's index buffers are no longer read
      load_chunk(c0, min(DW_IDX, ucols), 0);
      for (int ch = 0; ch < nchunks; ++ch) {
        const int slot = ch & 1;
        producers_sync();                     // chunk ch has landed (and chunk ch-1 is fully consumed)
        if (ch + 1 < nchunks) load_chunk(c0 + (ch + 1) * DW_IDX, min(DW_IDX, ucols - (ch + 1) * DW_IDX), slot ^ 1);
        const int f0 = ch * (DW_IDX / DN_KC), f1 = min(nst, f0 + DW_IDX / DN_KC);
        for (int f = f0; f < f1; ++f) {
          const uint32_t gf = g + f, buf = gf % DW_STAGES;
          const int kb = slot * DW_IDX + (f - f0) * DN_KC;
          ok = umma::mbar_wait(&bar_empty[buf], ((gf / DW_STAGES) & 1) ^ 1) && ok;   // first round: buffers are free
          uint8_t *sa = gen + buf * stage_bytes, *sb = sa + a_bytes;
#pragma unroll
          for (int c = 0; c < A_PER; ++c) {
            const unsigned m0 = masks_s[(kb + kh + 4 * c) * 4 + word], m1 = masks_s[(kb + kh + 4 * c + 1) * 4 + word];
            const unsigned m2 = masks_s[(kb + kh + 4 * c + 2) * 4 + word], m3 = masks_s[(kb + kh + 4 * c + 3) * 4 + word];
            const float4 v = make_float4(((m0 >> bit) & 1u) ? 1.f : 0.f, ((m1 >> bit) & 1u) ? 1.f : 0.f,
                                         ((m2 >> bit) & 1u) ? 1.f : 0.f, ((m3 >> bit) & 1u) ? 1.f : 0.f);
            *reinterpret_cast<float4 *>(sa + a_off[c]) = v;
          }
          mbar_arrive(&bar_full[buf]);                   // release: this thread's part of the A tile is written
#pragma unroll
          for (int i = 0; i < B_PER; ++i) {
            if (b_kr[i] >= 0) {
              const int col = cols_s[kb + b_kr[i]];
              const bool valid = (unsigned)col < (unsigned)p.x_rows;
              const float *src = valid ? p.xr + (long long)col * D + b_ch[i] * 4 : p.xr;
              cp_async_16(sb + b_off[i], src, valid ? 16 : 0);
            }
          }
          cp_async_mbar_arrive_noinc(&bar_full[buf]);    // second arrival: when this thread's copies have landed
        }
      }
      g += nst;

      // ---- epilogue of this super-window
      const uint32_t acc = it & 1, use = it >> 1;
      ok = umma::mbar_wait(&bar_tfull[acc], use & 1) && ok;
      umma::tc_fence_after_sync();
      {
        constexpr int CW = 256 / (DW_PRODUCERS / 128);     // accumulator columns per warp group of four
        const int lq = wid & 3, part = wid >> 2;
        const int row = sw * SW_H + lq * 32 + lane;
        for (int cc = part * CW; cc < part * CW + CW && cc < D; cc += 32) {
          uint32_t v[32];
          umma::tmem_ld_32x32(tmem_base + acc * 256 + ((uint32_t)(lq * 32) << 16) + (uint32_t)cc, v);
          umma::tmem_ld_wait();
          if (row < p.n_rows) {
            float *dst = p.y + (long long)row * p.ldy + cc;
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
        }
      }
      umma::tc_fence_before_sync();
      mbar_arrive(&bar_tempty[acc]);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  if (!ok && p.err) atomicExch(p.err, 1);
  umma::tc_fence_before_sync();
  __syncthreads();
  if (wid == 0) umma::tmem_dealloc(tmem_base, 512);
}

bool dense_supported(const float *x, const float *y, int64_t ldy, int32_t dim) {
  return dim >= 16 && dim <= 256 && (dim % 16) == 0 && (ldy % 4) == 0 &&
         ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
}

// Y rows of the plan's super-windows (+)= A * X.  xr_scratch: [x_rows * dim] floats.
int launch_spmm_dense(const float *x, int64_t ldx, int32_t x_rows, int32_t n_rows, int32_t dim, const int32_t *plan,
                      int32_t n_dense, int64_t total_cols, int accumulate, float *y, int64_t ldy, float *xr_scratch,
                      int *d_err, cudaStream_t stream) {
  if (n_dense <= 0) return 0;
  PlanView v = view_plan(plan, n_rows, n_dense, total_cols);
  const long long total4 = (long long)x_rows * (dim / 4);
  tf32_round_rows_kernel<<<1184, 256, 0, stream>>>(x, ldx, dim / 4, xr_scratch, total4);
  DenseParams p;
  p.xr = xr_scratch; p.x_rows = x_rows; p.dim = dim; p.n_rows = n_rows; p.n_dense = n_dense; p.accumulate = accumulate;
  p.sw_ids = v.sw_ids; p.sw_off = v.sw_off; p.cols = v.cols; p.masks = v.masks; p.y = y; p.ldy = ldy; p.err = d_err;
  const int natoms = (dim + 31) / 32;
  const size_t smem = (size_t)DN_STAGES * (SW_H * 128 + (DN_KC / 4) * natoms * 512) + 1024;
  cudaError_t err = cudaFuncSetAttribute(spmm_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) { set_error("spmm_dense attr: %s", cudaGetErrorString(err)); return (int)err; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = n_dense < sms ? n_dense : sms;
  if (tuning().dense_ws) {
    const size_t smem_ws = (size_t)DW_STAGES * (SW_H * 128 + (DN_KC / 4) * natoms * 512) + 2 * DW_IDX * 20 + 1024;
    err = cudaFuncSetAttribute(spmm_dense_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws);
    if (err != cudaSuccess) { set_error("spmm_dense_ws attr: %s", cudaGetErrorString(err)); return (int)err; }
    spmm_dense_ws_kernel<<<grid, DW_THREADS, smem_ws, stream>>>(p);
  } else {
    spmm_dense_kernel<<<grid, DN_THREADS, smem, stream>>>(p);
  }
  err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("spmm_dense launch: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

void dense_plan_arrays(const int32_t *plan, int32_t n_rows, int32_t n_dense, int64_t total_cols, const int **sw_ids,
                       const int **sw_off, const int **cols, const unsigned **masks) {
  PlanView v = view_plan(plan, n_rows, n_dense, total_cols);
  *sw_ids = v.sw_ids; *sw_off = v.sw_off; *cols = v.cols; *masks = v.masks;
}

const int32_t *dense_plan_labels(const int32_t *plan, int32_t n_rows, int32_t n_dense, int64_t total_cols) {
  return view_plan(plan, n_rows, n_dense, total_cols).ht2;
}

}  // namespace hcspmm
