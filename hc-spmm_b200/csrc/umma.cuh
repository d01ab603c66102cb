// umma.cuh -- hand-written sm_100a tensor-core plumbing: tcgen05.mma (TF32, operands in shared
// memory, accumulator in TMEM), TMEM allocation / loads, mbarriers, and the 128-byte-swizzled
// shared-memory operand layouts with their matrix descriptors.
//
// Field encodings follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor"
// tables (cross-checked against the comments of cute/arch/mma_sm100_desc.hpp in the CUTLASS
// headers shipped in this image; no CUTLASS code is used).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hcspmm {
namespace umma {

// ---- shared-memory address / mbarrier -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Bounded wait: returns false if the phase did not complete within `max_spins` polls, so that a
// protocol bug shows up as an error code instead of a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, uint32_t max_spins = 1u << 26) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t i = 0; i < max_spins; ++i) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return true;
  }
  return false;
}

// generic-proxy writes (st.shared, cp.async) -> visible to the async proxy (tcgen05.mma reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- TMEM -----------------------------------------------------------------------------------
// one full warp; ncols power of two in [32, 512]; the base address lands in *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp receives lane (row) t
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64 bit):
//   [0,14)  start address >> 4        [16,30) leading-dimension byte offset >> 4
//   [32,46) stride-dimension byte offset >> 4        [46,48) version = 1 (Blackwell)
//   [49,52) base offset = 0 (atoms are 1024-byte aligned)   [61,64) layout: 2 = SWIZZLE_128B
//           1 = SWIZZLE_128B_BASE32B -- the ONLY layout the hardware accepts for an MN-major
//           operand with 32-bit (TF32) elements (measured: a plain SWIZZLE_128B MN-major TF32
//           operand multiplies as all-zero; scripts/debug/umma_probe.cu)
constexpr uint32_t LAYOUT_SW128 = 2, LAYOUT_SW128_BASE32B = 1;
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return make_desc(smem_addr, lbo_bytes, sbo_bytes, LAYOUT_SW128);
}
// Instruction descriptor, kind::tf32, FP32 accumulate:
//   [4,6) D format: 1 = F32   [7,10) A format: 2 = TF32   [10,13) B format: 2 = TF32
//   [15] A major (0 = K)      [16] B major (1 = MN)       [17,23) N >> 3     [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; one thread issues on behalf of the CTA
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread -> one arrival on the mbarrier when they have completed
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- 128-byte swizzled operand tiles (TF32 = 4-byte elements, 32 per 128-byte row) -------------
// Swizzle<3,4,3>: within a 1024-byte atom (8 rows x 128 B) the 16-byte chunk index is XORed with
// the row index.  Atoms must start on 1024-byte boundaries.
//
// K-major operand (A): element (row r, k): atom = r / 8, stride 1024 B between atoms (SBO); one
// 32-element k-block per tile.  Byte offset inside the k-block tile:
__device__ __forceinline__ uint32_t kmajor_off(int r, int k /* 0..31 */) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + ((k & 3) << 2));
}
// MN-major TF32 operand (B: rows = k, contiguous along n -- how X / W rows lie in memory), layout
// SWIZZLE_128B_BASE32B: atom = 4 k-rows x 128 B (32 floats along n), Swizzle<2,5,2>: the 32-byte
// chunk index inside a row is XORed with the row index (k % 4).  n-atoms at stride LBO, k-atoms
// (4 rows) at stride SBO; one K = 8 MMA consumes two k-atoms.  Atoms start on 512-byte boundaries.
// Offset of the 16-byte piece holding features [4*chunk, 4*chunk + 4) of row k:
__device__ __forceinline__ uint32_t mnmajor_chunk_off(int k, int chunk /* 16-byte chunk along n */,
                                                      uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const int natom = chunk >> 3, kr = k & 3;
  const uint32_t b = (uint32_t)(chunk & 7) << 4;          // byte offset inside the 128-byte row
  return (uint32_t)(k >> 2) * sbo_bytes + (uint32_t)natom * lbo_bytes + (uint32_t)(kr * 128) +
         ((((b >> 5) ^ (uint32_t)kr) & 3u) << 5) + (b & 31u);
}

}  // namespace umma
}  // namespace hcspmm
