// Minimal tcgen05 / TMEM helpers (sm_100a): single-CTA (cta_group::1) UMMA with bf16 operands in
// shared memory (K-major, 128 B swizzle) and fp32 accumulators in tensor memory.
//
// Operand tile convention used by every kernel in this library ("K64 tile"):
//   a tile holds R rows (M or N index) x 64 bf16 (128 B) along K; rows are grouped in atoms of 8 rows
//   (1024 B, 1024 B-aligned); inside an atom the 16-byte chunk c (0..7) of row r (0..7) is stored at chunk
//   position c ^ r (Swizzle<3,4,3>, the layout TMA's SWIZZLE_128B and UMMA's LayoutType::SWIZZLE_128B share).
//   Element (row m, k):  byte offset = (m / 8) * 1024 + (m % 8) * 128 + (((k / 8) ^ (m % 8)) << 4) + (k % 8) * 2.
//   A K extent of 128 is two K64 tiles.  One tcgen05.mma consumes K = 16 (32 B): k-step j of a tile is
//   addressed by adding j * 32 B to the descriptor's start address.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace dpt {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t k64_offset(int m, int k) {
  return (uint32_t)((m >> 3) * 1024 + (m & 7) * 128 + ((((k >> 3) ^ (m & 7)) & 7) << 4) + (k & 7) * 2);
}

// 64-bit shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row atoms 1024 B apart.
__device__ __forceinline__ uint64_t make_desc_k64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // [0,14)  start address >> 4
  d |= (uint64_t)1 << 16;                        // [16,30) leading-dim byte offset >> 4 (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // [32,46) stride-dim byte offset >> 4: next 8-row atom
  d |= (uint64_t)1 << 46;                        // [46,48) descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                        // [61,64) layout type SWIZZLE_128B
  return d;
}

// 32-bit instruction descriptor, kind::f16: D = fp32, A = B = bf16, both K-major, M x N.
__device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                    // [4,6)   D format: F32
  d |= 1u << 7;                    // [7,10)  A format: BF16
  d |= 1u << 10;                   // [10,13) B format: BF16
  d |= (uint32_t)(N >> 3) << 17;   // [17,23) N >> 3
  d |= (uint32_t)(M >> 4) << 24;   // [24,29) M >> 4
  return d;
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, one K = 16 step; issued by ONE thread.
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// all previously issued MMAs of this thread -> arrive on the mbarrier when they have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane (row) -> registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// TMEM address of (this warp's 32-lane slice, column c) relative to the allocation base
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, int warp, int col) {
  return base + ((uint32_t)(warp * 32) << 16) + (uint32_t)col;
}

// store 8 consecutive-k bf16 values (one 16-byte chunk) of row m, chunk c (k = 8c..8c+7) into a K64 tile
__device__ __forceinline__ void st_chunk(unsigned char* tile, int m, int c, const float* v8) {
  uint4 u;
  __nv_bfloat162 p0 = __floats2bfloat162_rn(v8[0], v8[1]), p1 = __floats2bfloat162_rn(v8[2], v8[3]);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(v8[4], v8[5]), p3 = __floats2bfloat162_rn(v8[6], v8[7]);
  u.x = *reinterpret_cast<uint32_t*>(&p0), u.y = *reinterpret_cast<uint32_t*>(&p1);
  u.z = *reinterpret_cast<uint32_t*>(&p2), u.w = *reinterpret_cast<uint32_t*>(&p3);
  *reinterpret_cast<uint4*>(tile + (m >> 3) * 1024 + (m & 7) * 128 + (((c ^ (m & 7)) & 7) << 4)) = u;
}

}  // namespace umma
}  // namespace dpt
