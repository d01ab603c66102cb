// gemm.cu -- the Update product out = Z * W of the fused entry points.
//
// The reference multiplies the aggregated window tile by W with wmma TF32 inside its fused
// kernels (/root/reference/hybrid_kernel/hybrid_all_kernel.cu:1809-1837, 2285-2316,
// 2741-2768): both operands rounded with cvt.rna.tf32, FP32 accumulate, hidden == 32 only.
// This is the same arithmetic for any (m, k, n): a tiled mma.sync.m16n8k8 TF32 kernel,
// 128 x 64 x 16 CTA tile, eight warps as 4 x 2, register-prefetched global loads, padded
// bank-conflict-free shared tiles.
#include "common.cuh"

namespace hcspmm {

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int AS = BK + 4;  // A tile row stride (floats): rows g*20 + k hit distinct banks
constexpr int BS = BN + 8;  // B tile row stride: k*72 + n hit distinct banks

struct GemmParams {
  const float *a, *b;
  float *out;
  long long lda, ldb, ldo;
  int m, k, n;
  int a_vec, b_vec, o_vec;
};

__global__ void __launch_bounds__(CTA_THREADS) gemm_tf32_kernel(const GemmParams p) {
  __shared__ __align__(16) float As[BM * AS];
  __shared__ __align__(16) float Bs[BK * BS];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const int wm = wid >> 1, wn = wid & 1;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // global -> register staging: A: 128 rows x 16 k, thread = (row t/2, 8 k's); B: 16 k x 64 n,
  // thread = (k t/16, 4 n's)
  const int ar = tid >> 1, ak = (tid & 1) * 8;
  const int bk = tid >> 4, bn = (tid & 15) * 4;
  float ra[8], rb[4];

  auto load_tiles = [&](int k0) {
    const int row = m0 + ar;
    const float *ap = p.a + (long long)row * p.lda + k0 + ak;
    if (row < p.m && p.a_vec && k0 + ak + 8 <= p.k) {
      float4 v0 = ldg_f4(ap), v1 = ldg_f4(ap + 4);
      ra[0] = v0.x; ra[1] = v0.y; ra[2] = v0.z; ra[3] = v0.w;
      ra[4] = v1.x; ra[5] = v1.y; ra[6] = v1.z; ra[7] = v1.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) ra[i] = (row < p.m && k0 + ak + i < p.k) ? __ldg(ap + i) : 0.f;
    }
    const int kk = k0 + bk, col = n0 + bn;
    const float *bp = p.b + (long long)kk * p.ldb + col;
    if (kk < p.k && p.b_vec && col + 4 <= p.n) {
      float4 v = ldg_f4(bp);
      rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) rb[i] = (kk < p.k && col + i < p.n) ? __ldg(bp + i) : 0.f;
    }
  };
  auto store_tiles = [&]() {
    float *as = As + ar * AS + ak;
    *reinterpret_cast<float4 *>(as) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    *reinterpret_cast<float4 *>(as + 4) = make_float4(ra[4], ra[5], ra[6], ra[7]);
    *reinterpret_cast<float4 *>(Bs + bk * BS + bn) = make_float4(rb[0], rb[1], rb[2], rb[3]);
  };

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

  load_tiles(0);
  for (int k0 = 0; k0 < p.k; k0 += BK) {
    __syncthreads();
    store_tiles();
    __syncthreads();
    if (k0 + BK < p.k) load_tiles(k0 + BK);
#pragma unroll
    for (int ks = 0; ks < BK / 8; ++ks) {
      uint32_t af[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const float *ap = As + (wm * 32 + mi * 16 + g) * AS + ks * 8 + tig;
        af[mi][0] = f32_to_tf32(ap[0]);
        af[mi][1] = f32_to_tf32(ap[8 * AS]);
        af[mi][2] = f32_to_tf32(ap[4]);
        af[mi][3] = f32_to_tf32(ap[8 * AS + 4]);
      }
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const float *bp = Bs + (ks * 8 + tig) * BS + wn * 32 + ni * 8 + g;
        const uint32_t b0 = f32_to_tf32(bp[0]), b1 = f32_to_tf32(bp[4 * BS]);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) mma_m16n8k8_tf32(acc[mi][ni], af[mi], b0, b1);
      }
    }
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const int col = n0 + wn * 32 + ni * 8 + tig * 2;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = m0 + wm * 32 + mi * 16 + g + h * 8;
        if (row >= p.m) continue;
        float *dst = p.out + (long long)row * p.ldo + col;
        const float v0 = acc[mi][ni][2 * h], v1 = acc[mi][ni][2 * h + 1];
        if (p.o_vec && col + 2 <= p.n) {
          *reinterpret_cast<float2 *>(dst) = make_float2(v0, v1);
        } else {
          if (col < p.n) dst[0] = v0;
          if (col + 1 < p.n) dst[1] = v1;
        }
      }
    }
}

int launch_gemm_tf32(const float *a, int64_t lda, const float *b, int64_t ldb, int32_t m, int32_t k,
                     int32_t n, float *out, int64_t ldo, cudaStream_t stream) {
  if (m < 0 || k < 0 || n < 0) { set_error("gemm: negative size"); return HCSPMM_E_INVALID; }
  if (m == 0 || n == 0) return 0;
  if (!a || !b || !out) { set_error("gemm: null pointer argument"); return HCSPMM_E_INVALID; }
  if (lda < k || ldb < n || ldo < n) { set_error("gemm: leading dimension too small"); return HCSPMM_E_INVALID; }
  GemmParams p;
  p.a = a; p.b = b; p.out = out; p.lda = lda; p.ldb = ldb; p.ldo = ldo; p.m = m; p.k = k; p.n = n;
  p.a_vec = ((reinterpret_cast<uintptr_t>(a) & 15) == 0) && (lda % 4 == 0);
  p.b_vec = ((reinterpret_cast<uintptr_t>(b) & 15) == 0) && (ldb % 4 == 0);
  p.o_vec = ((reinterpret_cast<uintptr_t>(out) & 7) == 0) && (ldo % 2 == 0);
  dim3 grid((m + BM - 1) / BM, (n + BN - 1) / BN, 1);
  gemm_tf32_kernel<<<grid, CTA_THREADS, 0, stream>>>(p);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("gemm launch: %s", cudaGetErrorString(err)); return (int)err; }
  return 0;
}

}  // namespace hcspmm
