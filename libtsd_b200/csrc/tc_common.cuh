// tcgen05 / TMEM building blocks shared by the tensor-core kernels (fir_tc.cu, resamp_tc.cu): tf32 split, UMMA
// shared-memory descriptors (K-major, 128-byte swizzle), MMA issue with the A operand in tensor memory, TMEM stores,
// commits, fences.  Bit layouts follow cute::UMMA::SmemDescriptor / InstrDescriptor (CUTLASS 4.x, mma_sm100_desc.hpp).
#pragma once
#include "common.cuh"

namespace tsdgpu {
namespace tc {

__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }   // Swizzle<3,4,3>
// round to the nearest tf32 (10-bit mantissa), ties away from zero like cvt.rna.tf32.f32, with two full-rate integer
// instructions (the conversion instruction itself issues at a small fraction of the FP32 rate: it made the producers
// the bottleneck of the kernel)
__device__ __forceinline__ float to_tf32(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
// shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr)
{
  return (uint64_t) ((saddr >> 4) & 0x3FFFu) | ((uint64_t) 1 << 16) /* LBO (unused for swizzled K-major) */ |
         ((uint64_t) (1024 >> 4) << 32) /* SBO */ | ((uint64_t) 1 << 46) /* version */ | ((uint64_t) 2 << 61) /* SWIZZLE_128B */;
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128; N (bits 17..22, N >> 3) is added per MMA
constexpr uint32_t IDESC_M128 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (128 >> 4) << 24);

// D[128][N] += A[128][8] * B[N][8]^T, A read from tensor memory (lanes 0..127, 8 columns), B from shared memory
// (always accumulating: the epilogue leaves every region zeroed)
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc)
{
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "setp.ne.b32 p, 1, 0;\n\t"
    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
    ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc)
    : "memory");
}
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// 16 consecutive TMEM columns of this thread's lane <- 16 registers
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16])
{
  asm volatile(
    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
    "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
    "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])),
    "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
    "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
    : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }


} // namespace tc
} // namespace tsdgpu
