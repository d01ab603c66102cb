// loa.cu -- LOA (layout-optimisation) vertex reordering on the GPU, bit-exact with the reference's
// reorder_plus_new_direct (/root/reference/LOI.cpp:660-805) and main's output order (:873-891).
//
// The algorithm is a sequential greedy: blocks of <= 16 vertices are grown one after another, each
// block by up to 15 dependent "pick the best candidate" steps, and block k+1 depends on the visit
// set left by block k.  Speculating across blocks would change the result, so the sequential
// skeleton is kept and everything INSIDE a step is made data-parallel on one persistent CTA of
// 1024 threads (state in global memory / L2, a few block barriers per step, no host round trips):
//
//   scan     for every new block column c and every row r of c's CSC list that is still unvisited (work is dealt
//            by flat position over all lists, so hub columns are shared by the whole CTA):
//            cns[r] += 1 (atomicAdd) and r is discovered.  The reference discovers candidates in
//            scan order and breaks profit ties by "first discovered" (strict '>' at :731,779), so
//            each touch carries its position in the reference's nested-loop order and the vertex
//            keeps the minimum (atomicMin on a 64-bit key): the discovery order is reproduced
//            without executing the scan serially.
//   select   arg-max over the discovered, unvisited vertices of (float)ones / rows with
//            ones = old_ones + deg, rows = old_rows + deg - cns (:768-775); ties -> smallest
//            discovery key.  Block-wide reduction.
//   residual the winner's columns not yet in the block's sorted column set (binary search, :60-73),
//            compacted IN ROW ORDER (their order drives the next scan), then merged into the set.
//   reset    counters of every discovered vertex (:797-801).
#include "common.cuh"

namespace hcspmm {

constexpr int LOA_THREADS = 1024;
constexpr unsigned long long DISC_NONE = ~0ull;

struct LoaParams {
  const int *rowptr, *colidx, *rowptr_in, *colidx_in;
  int n;
  unsigned char *visit;        // [n]
  int *cns;                    // [n]
  unsigned long long *disc;    // [n] discovery key, DISC_NONE = not discovered in this block
  int *pro;                    // [n] discovered vertices of the current block (any order)
  int *cols_a, *cols_b;        // [cap] sorted column set of the block (ping-pong)
  int *resi;                   // [cap] new columns contributed by the last winner, row order
  long long *pref;             // [cap+1] exclusive prefix of CSC list lengths over the scan list
  int *blk_vert;               // [n] vertices in block creation order
  int *blk_start;              // [n+1]
  int *n_blocks;               // [1]
  int *perm;                   // [n] output
  int *n_full;                 // [1]
};

template <typename T>
__device__ __forceinline__ T block_scan_excl(T v, T *warp_buf, T &total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  T inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) warp_buf[wid] = inc;
  __syncthreads();
  T wsum = warp_buf[lane];  // LOA_THREADS / 32 == 32 warps
  T winc = wsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, winc, o);
    if (lane >= o) winc += t;
  }
  total = __shfl_sync(0xffffffffu, winc, 31);
  T wexc = __shfl_sync(0xffffffffu, winc - wsum, wid);
  return wexc + inc - v;
}

__device__ __forceinline__ int lower_bound_i(const int *a, int n, int x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(LOA_THREADS, 1) loa_kernel(const LoaParams p) {
  __shared__ long long sbuf_ll[32];
  __shared__ int sbuf_i[32];
  __shared__ float red_p[32];
  __shared__ unsigned long long red_k[32];
  __shared__ int red_i[32];
  __shared__ int s_cur, s_npro, s_winner;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = p.n;
  int *cols = p.cols_a, *cols_next = p.cols_b;
  int nblk = 0, nvert = 0;
  int cur = 0;
  if (tid == 0) p.blk_start[0] = 0;

  for (;;) {
    // ---- seed: first vertex >= cur with degree > 0 that is unvisited (LOI.cpp:664-669, 698-704)
    for (;;) {
      if (tid == 0) s_cur = 0x7fffffff;
      __syncthreads();
      const int v = cur + tid;
      if (v < n && p.rowptr[v + 1] - p.rowptr[v] > 0 && !__ldcg(p.visit + v)) atomicMin(&s_cur, v);
      __syncthreads();
      const int found = s_cur;
      __syncthreads();
      if (found != 0x7fffffff) { cur = found; break; }
      cur += LOA_THREADS;
      if (cur >= n) break;
    }
    if (cur >= n) break;
    const int v0 = cur;
    const int deg0 = p.rowptr[v0 + 1] - p.rowptr[v0];
    if (tid == 0) {
      __stcg(p.visit + v0, (unsigned char)1);
      p.blk_vert[nvert] = v0;
      s_npro = 0;
    }
    int bsz = 1;
    // block column set := cols(v0) (sorted: canonical CSR), :746-748
    for (int i = tid; i < deg0; i += LOA_THREADS) cols[i] = p.colidx[p.rowptr[v0] + i];
    int ncols = deg0;
    int old_ones = deg0, old_rows = deg0;
    const int *scan_list = p.colidx + p.rowptr[v0];
    int scan_len = deg0;
    unsigned long long key_base = 0;
    __syncthreads();

    for (int step = 0; step < 15; ++step) {
      // ---- scan the CSC lists of the new columns (:709-719, :760-770)
      long long carry = 0;
      for (int t0 = 0; t0 < scan_len; t0 += LOA_THREADS) {
        const int i = t0 + tid;
        long long len = 0;
        if (i < scan_len) {
          const int c = scan_list[i];
          len = p.rowptr_in[c + 1] - p.rowptr_in[c];
        }
        long long total;
        long long ex = block_scan_excl<long long>(len, sbuf_ll, total);
        if (i < scan_len) p.pref[i] = carry + ex;
        carry += total;
      }
      __syncthreads();
      // every touch (column i of the scan list, position j in its CSC list) is one unit of work, dealt to the
      // threads by its FLAT position t in the reference's nested-loop order -- which is also its discovery key --
      // so a hub column with 10^5 entries is spread over the whole CTA instead of being walked by one warp
      // (one warp per column: 564 s on the products shape).  The column of t is found in the prefix sums.
      for (long long t = tid; t < carry; t += LOA_THREADS) {
        int lo = 0, hi = scan_len - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (p.pref[mid] <= t) lo = mid; else hi = mid - 1;
        }
        const int c = scan_list[lo];
        const int r = p.colidx_in[p.rowptr_in[c] + (int)(t - p.pref[lo])];
        if (!__ldcg(p.visit + r)) {
          atomicAdd(&p.cns[r], 1);
          const unsigned long long old = atomicMin(&p.disc[r], key_base + (unsigned long long)t);
          if (old == DISC_NONE) p.pro[atomicAdd(&s_npro, 1)] = r;
        }
      }
      key_base += (unsigned long long)carry;
      __syncthreads();
      const int npro = s_npro;

      // ---- select (:724-736, :772-787): max profit, ties -> first discovered
      float best_p = 0.0f;
      unsigned long long best_k = DISC_NONE;
      int best_i = -1;
      for (int k = tid; k < npro; k += LOA_THREADS) {
        const int i = p.pro[k];
        if (!__ldcg(p.visit + i)) {
          const int deg = p.rowptr[i + 1] - p.rowptr[i];
          const int rows = old_rows + deg - __ldcg(p.cns + i);
          const int ones = old_ones + deg;
          const float pr = __fdiv_rn((float)ones, (float)rows);
          const unsigned long long dk = __ldcg(p.disc + i);
          if (pr > best_p || (pr == best_p && dk < best_k)) { best_p = pr; best_k = dk; best_i = i; }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float op = __shfl_xor_sync(0xffffffffu, best_p, o);
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, best_k, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (op > best_p || (op == best_p && ok < best_k)) { best_p = op; best_k = ok; best_i = oi; }
      }
      if (lane == 0) { red_p[wid] = best_p; red_k[wid] = best_k; red_i[wid] = best_i; }
      __syncthreads();
      if (wid == 0) {
        best_p = red_p[lane]; best_k = red_k[lane]; best_i = red_i[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float op = __shfl_xor_sync(0xffffffffu, best_p, o);
          const unsigned long long ok = __shfl_xor_sync(0xffffffffu, best_k, o);
          const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
          if (op > best_p || (op == best_p && ok < best_k)) { best_p = op; best_k = ok; best_i = oi; }
        }
        if (lane == 0) s_winner = best_i;
      }
      __syncthreads();
      const int win = s_winner;
      if (win < 0) break;                       // :737-740 / :788-790
      const int wdeg = p.rowptr[win + 1] - p.rowptr[win];
      if (tid == 0) {
        __stcg(p.visit + win, (unsigned char)1);
        p.blk_vert[nvert + bsz] = win;
      }
      ++bsz;
      if (bsz == 16) break;                     // 1 seed + 15 picks; no further scan is needed
      __syncthreads();

      // ---- residual columns of the winner, in row order (cal_resi_elements, :60-73)
      int rbase = 0;
      for (int t0 = 0; t0 < wdeg; t0 += LOA_THREADS) {
        const int i = t0 + tid;
        int c = 0, isnew = 0;
        if (i < wdeg) {
          c = p.colidx[p.rowptr[win] + i];
          const int pos = lower_bound_i(cols, ncols, c);
          isnew = !(pos < ncols && cols[pos] == c);
        }
        int total;
        const int ex = block_scan_excl<int>(isnew, sbuf_i, total);
        if (isnew) p.resi[rbase + ex] = c;
        rbase += total;
      }
      __syncthreads();
      const int nresi = rbase;
      // merge the sorted set with the (sorted, disjoint) residual
      for (int i = tid; i < ncols; i += LOA_THREADS) {
        const int x = cols[i];
        cols_next[i + lower_bound_i(p.resi, nresi, x)] = x;
      }
      for (int i = tid; i < nresi; i += LOA_THREADS) {
        const int x = p.resi[i];
        cols_next[i + lower_bound_i(cols, ncols, x)] = x;
      }
      __syncthreads();
      { int *t = cols; cols = cols_next; cols_next = t; }
      ncols += nresi;
      old_ones += wdeg;
      old_rows = ncols;
      scan_list = p.resi;
      scan_len = nresi;
    }
    __syncthreads();
    // ---- reset the counters of every discovered vertex (:797-801) and emit the block
    const int npro = s_npro;
    for (int k = tid; k < npro; k += LOA_THREADS) {
      const int i = p.pro[k];
      p.cns[i] = 0;
      p.disc[i] = DISC_NONE;
    }
    nvert += bsz;
    ++nblk;
    if (tid == 0) p.blk_start[nblk] = nvert;
    __syncthreads();
  }

  // ---- main's file order (:873-891): full blocks, then partial blocks, then unvisited ascending
  __syncthreads();
  int full_total = 0;
  {
    int carry_f = 0, carry_p = 0;
    // pass 1: number of vertices in full blocks
    for (int t0 = 0; t0 < nblk; t0 += LOA_THREADS) {
      const int b = t0 + tid;
      int isfull = (b < nblk) && (p.blk_start[b + 1] - p.blk_start[b] == 16);
      int total;
      block_scan_excl<int>(isfull, sbuf_i, total);
      full_total += total;
    }
    const int part_base = full_total * 16;
    for (int t0 = 0; t0 < nblk; t0 += LOA_THREADS) {
      const int b = t0 + tid;
      int sz = 0;
      if (b < nblk) sz = p.blk_start[b + 1] - p.blk_start[b];
      int tf, tp;
      const int exf = block_scan_excl<int>(sz == 16 ? 16 : 0, sbuf_i, tf);
      const int exp_ = block_scan_excl<int>(sz < 16 ? sz : 0, sbuf_i, tp);
      if (b < nblk) {
        const int dst = sz == 16 ? carry_f + exf : part_base + carry_p + exp_;
        for (int k = 0; k < sz; ++k) p.perm[dst + k] = p.blk_vert[p.blk_start[b] + k];
      }
      carry_f += tf;
      carry_p += tp;
    }
    int obase = nvert;
    for (int t0 = 0; t0 < n; t0 += LOA_THREADS) {
      const int v = t0 + tid;
      const int un = (v < n) && !__ldcg(p.visit + v);
      int total;
      const int ex = block_scan_excl<int>(un, sbuf_i, total);
      if (un) p.perm[obase + ex] = v;
      obase += total;
    }
  }
  if (tid == 0) { *p.n_blocks = nblk; *p.n_full = full_total; }
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

size_t loa_workspace_bytes(int32_t n, int64_t nnz, int32_t max_degree) {
  size_t cap = (size_t)16 * (size_t)(max_degree > 0 ? max_degree : 1) + 16;
  if ((int64_t)cap > nnz + 16) cap = (size_t)nnz + 16;
  size_t nn = (size_t)(n > 0 ? n : 1);
  return align_up(nn) + align_up(nn * 4) + align_up(nn * 8) + align_up(nn * 4) + 3 * align_up(cap * 4) +
         align_up((cap + 1) * 8) + align_up(nn * 4) + align_up((nn + 1) * 4) + 512;
}

int launch_loa(const int32_t *rowptr, const int32_t *colidx, const int32_t *rowptr_in,
               const int32_t *colidx_in, int32_t n, int64_t nnz, int32_t max_degree, int32_t *perm,
               int32_t *block_sizes, int32_t *counts /* [2]: n_blocks, n_full */, void *ws,
               size_t ws_bytes, cudaStream_t stream) {
  if (n < 0 || nnz < 0) { set_error("loa: negative size"); return HCSPMM_E_INVALID; }
  if (n == 0) return 0;
  if (!rowptr || !rowptr_in || !perm || !counts || !ws || (nnz > 0 && (!colidx || !colidx_in))) {
    set_error("loa: null pointer argument");
    return HCSPMM_E_INVALID;
  }
  if (ws_bytes < loa_workspace_bytes(n, nnz, max_degree)) {
    set_error("loa: workspace too small");
    return HCSPMM_E_WORKSPACE;
  }
  size_t cap = (size_t)16 * (size_t)(max_degree > 0 ? max_degree : 1) + 16;
  if ((int64_t)cap > nnz + 16) cap = (size_t)nnz + 16;
  char *w = reinterpret_cast<char *>(ws);
  size_t nn = (size_t)n;
  LoaParams p;
  p.rowptr = rowptr; p.colidx = colidx; p.rowptr_in = rowptr_in; p.colidx_in = colidx_in; p.n = n;
  p.visit = reinterpret_cast<unsigned char *>(w); w += align_up(nn);
  p.cns = reinterpret_cast<int *>(w); w += align_up(nn * 4);
  p.disc = reinterpret_cast<unsigned long long *>(w); w += align_up(nn * 8);
  p.pro = reinterpret_cast<int *>(w); w += align_up(nn * 4);
  p.cols_a = reinterpret_cast<int *>(w); w += align_up(cap * 4);
  p.cols_b = reinterpret_cast<int *>(w); w += align_up(cap * 4);
  p.resi = reinterpret_cast<int *>(w); w += align_up(cap * 4);
  p.pref = reinterpret_cast<long long *>(w); w += align_up((cap + 1) * 8);
  p.blk_vert = reinterpret_cast<int *>(w); w += align_up(nn * 4);
  p.blk_start = reinterpret_cast<int *>(w); w += align_up((nn + 1) * 4);
  p.n_blocks = counts; p.n_full = counts + 1;
  p.perm = perm;
  cudaError_t err = cudaMemsetAsync(p.visit, 0, align_up(nn) + align_up(nn * 4), stream);  // visit + cns
  if (err == cudaSuccess) err = cudaMemsetAsync(p.disc, 0xff, nn * 8, stream);
  if (err != cudaSuccess) { set_error("loa: memset: %s", cudaGetErrorString(err)); return (int)err; }
  loa_kernel<<<1, LOA_THREADS, 0, stream>>>(p);
  err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("loa launch: %s", cudaGetErrorString(err)); return (int)err; }
  if (block_sizes) {
    // block sizes = adjacent differences of blk_start; the caller reads counts[0] of them
    err = cudaMemcpyAsync(block_sizes, p.blk_start, (nn + 1) * 4, cudaMemcpyDeviceToDevice, stream);
    if (err != cudaSuccess) { set_error("loa: copy: %s", cudaGetErrorString(err)); return (int)err; }
  }
  return 0;
}

}  // namespace hcspmm
