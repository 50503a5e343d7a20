// Overlap-save block filter with the whole transform resident on ONE SM (sm_100a).
//
// Serves every filtre_fft / filtre_rif_fft object whose spectral gains are the transform of K taps
// (FiltreFFTRIF convention, reference fourier.cc:946-990; C ABI: fir_len = K > 0).  For such an H the samples the
// reference's overlap-add emits (fourier.cc:837-882) are y_out[t] = sum_m h[m] x[t - (Ne-K) - m] whatever its block
// length Ne and transform size N are, so the device is free to pick the transform size that fits the
// machine: M = 16384 points = 128 KiB of cf32, the largest block that stays inside one SM's shared memory.  The
// reference's bookkeeping (Ne, N, N_zeros, re-blocking residual, Ne samples per completed block, delay Ne-K) is kept
// on the host (ola.cu); this file only produces the samples.
//
// One persistent CTA per SM walks consecutive internal blocks of L = M - O outputs (O = 4096 or 8192 >= K-1 samples of
// overlap).  Per block, with n = 512 n1 + 32 n2 + n3 and k = k1 + 32 k2 + 512 k3 (radices 32 x 16 x 32):
//   P1  window (shared memory, filled by the TMA unit)  -> radix-32 over n1 -> exchange buffer E          [warp = n2, lane = n3]
//   P2  E -> W512 twiddle -> radix-16 over n2 -> W_M twiddle -> warp-local transpose -> radix-32 over n3 -> x H/M ->
//       inverse radix-32 -> warp-local transpose -> conj W_M twiddle -> inverse radix-16 -> conj W512 twiddle -> E
//       (a warp owns k1 = 2w, 2w+1: everything between the two CTA barriers happens in the warp's own 8 KiB of E)
//   P3  E -> inverse radix-32 over k1 -> the L valid outputs, plain coalesced stores                        [warp = n2, lane = n3]
// A sample crosses HBM once in and once out (16 B per sample; the O overlap samples are re-read from L2), nothing
// else leaves the SM: no scratch in L2, no inter-CTA hand-over.
//
// What is Blackwell-specific here:
//   * the 256 KiB of tensor memory hold the per-thread constants — the 32 gains H[k]/M and the 32 W_M twiddles each
//     thread needs in every block (tcgen05.st once per launch, tcgen05.ld in the loop): 16 B per point that neither
//     shared memory (full: 128 KiB exchange + 96 KiB staging) nor L2 has to deliver;
//   * the next block's input is prefetched by 1-D bulk copies (cp.async.bulk, SASS UBLKCP) issued by a producer warp:
//     the L new samples as soon as every math warp has consumed the current window, the O overlap samples (an L2
//     hit) into the exchange buffer as soon as the current block has left it; completion on mbarriers, the math
//     warps never wait on a global load in steady state.
#include "common.cuh"
#include "tc_common.cuh"
#include "ols16k.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace tsdgpu {

constexpr int OLS_M = 16384;
constexpr int OLS_MATH_WARPS = 16;
constexpr int OLS_MATH_THREADS = OLS_MATH_WARPS * 32;
constexpr int OLS_THREADS = OLS_MATH_THREADS + 128;    // + producer warpgroup (one lane works; a whole group so that setmaxnreg applies)
constexpr int OLS_MATH_REGS = 112, OLS_PROD_REGS = 24;  // 512 * 112 + 128 * 24 <= 640 * 96, the CTA's register allocation at launch
constexpr int OLS_E_BYTES = OLS_M * 8;                 // exchange buffer
constexpr int OLS_PIECE = 16384;                       // bytes per bulk copy

// OLS_TW1_OUTER = 1 (experiment, measured -1 %: 164.8 vs 166.4 Gsamples/s): the W512^(n2*k1) products sit in P1 / P3, where
// n2 is the same for the whole warp and the twiddle a constant-bank operand, instead of P2 (the FMA-bound phase).  P2 gets
// ~1 200 cycles shorter, but the hand-over phase is bound by the serial chain of each role warp (load -> butterflies ->
// wait -> store), and the extra products sit on that chain.  Default 0: products in P2.
#ifndef OLS_TW1_OUTER
#define OLS_TW1_OUTER 0
#endif
// W512^(n2*k1), k1 < 32, n2 < 16 (host-built in double): at [n2*32 + k1] (OLS_TW1_OUTER) or [k1*16 + n2]
__constant__ float2 c_ols_tw1[512];

struct OlsParams
{
  const float2 *x;
  float2 *y;
  const float2 *carry;        // [nchan][carry_len], samples preceding x[0]
  const float4 *tmem_init;    // [512 threads][32 float4]: 16 float4 of gains, 16 of twiddles
  long long x_stride, y_stride, out_count, total;
  int carry_len, n, base, jblocks, aligned;
  int offset_roles;
  unsigned *prof;             // optional per-phase clock trace of CTA 0 (TSDGPU_OLS_PROF), else null
};

// ---- radix-4 / 16 / 32 butterflies on registers: v[i*S], all indices static ---------------------------------
#define OLS_C1 0.92387953251128674f   // cos(pi/8)
#define OLS_S1 0.38268343236508977f   // sin(pi/8)
#define OLS_R2 0.70710678118654752f   // sqrt(1/2)

template<bool INV> __device__ __forceinline__ void bf4(float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
  const float2 s0 = add2(a0, a2), s1 = sub2(a0, a2), s2 = add2(a1, a3), d = sub2(a1, a3);
  const float2 ds = make_float2(d.y, d.x);
  const float2 j = INV ? make_float2(-1.f, 1.f) : make_float2(1.f, -1.f);
  const float2 nj = INV ? make_float2(1.f, -1.f) : make_float2(-1.f, 1.f);
  a0 = add2(s0, s2);
  a2 = sub2(s0, s2);
  a1 = fma2(ds, j, s1);
  a3 = fma2(ds, nj, s1);
}
// cos / sin of 2 pi k / 32 as compile-time constants
__host__ __device__ constexpr float ols_cos32(int k)
{
  k &= 31;
  if(k > 16) k = 32 - k;
  return k == 0 ? 1.f : k == 1 ? 0.98078528040323044913f : k == 2 ? 0.92387953251128675613f : k == 3 ? 0.83146961230254523708f
       : k == 4 ? 0.70710678118654752440f : k == 5 ? 0.55557023301960222474f : k == 6 ? 0.38268343236508977173f
       : k == 7 ? 0.19509032201612826785f : k == 8 ? 0.f : k == 9 ? -0.19509032201612826785f : k == 10 ? -0.38268343236508977173f
       : k == 11 ? -0.55557023301960222474f : k == 12 ? -0.70710678118654752440f : k == 13 ? -0.83146961230254523708f
       : k == 14 ? -0.92387953251128675613f : k == 15 ? -0.98078528040323044913f : -1.f;
}
__host__ __device__ constexpr float ols_sin32(int k) { return ols_cos32(k - 8); }
// OLS_CONST_TW = 1 (default): the radix-32 / 16 butterfly twiddles (w, i w) = (wr, wi, -wi, wr) come from a __constant__
// table: the packed FFMA2 / FMUL2 take them as ONE uniform-register operand pair fetched by LDCU.128, where the
// compile-time literals of OLS_CONST_TW = 0 are materialised with up to four UMOV / HFMA2 / MOV per twiddle (HFMA2 sits on
// the FMA pipe; ~13 % of the executed instructions of the kernel were such moves).
#ifndef OLS_CONST_TW
#define OLS_CONST_TW 1
#endif
// Where the table form is used: P3 only (butterflies fft16s and the last radix-2 stage comb32).  Measured at the BASELINE
// size on one box: literals everywhere 167.3 Gsamples/s, table in P3 168.9, table in P1 + P3 161.2 (ptxas then hoists the 32
// swizzled transpose addresses of P2 out of the block loop and spills them: 152 B of stack), table everywhere 169.0 (128 B
// of stack), table in the fft16s of P1 + P3 only 166.0.
#ifndef OLS_CT_P1F
#define OLS_CT_P1F 0
#endif
#ifndef OLS_CT_P1C
#define OLS_CT_P1C 0
#endif
#ifndef OLS_CT_P3F
#define OLS_CT_P3F 1
#endif
#ifndef OLS_CT_P3C
#define OLS_CT_P3C 1
#endif
__constant__ float4 c_ols_w32[2][32];   // [inverse][k] = (cos, -+sin, +-sin, cos) of 2 pi k / 32
// v * W32^K (forward, W = exp(-2 pi i / 32)) or its conjugate (inverse): the twiddle and its rotation are constants
// CT: table form (used in the hand-over phases P1 / P3; in P2, where every register is taken, the table operands cost spills)
template<bool INV, int K, bool CT> __device__ __forceinline__ float2 mul_w32(float2 v)
{
  if(K == 0) return v;
  if(K == 8) return mul2(make_float2(v.y, v.x), INV ? make_float2(-1.f, 1.f) : make_float2(1.f, -1.f));   // -+ i
  if(CT && OLS_CONST_TW)
  {
    const float4 t = c_ols_w32[INV ? 1 : 0][K & 31];
    return cmul_rot(v, make_float2(t.x, t.y), make_float2(t.z, t.w));
  }
  constexpr float wr = ols_cos32(K), wi = INV ? ols_sin32(K) : -ols_sin32(K);
  return cmul_rot(v, make_float2(wr, wi), make_float2(-wi, wr));
}
// 16-point DFT of v[0], v[S], ..., v[15 S]; natural order in and out
template<bool INV, int S, bool CT = false> __device__ __forceinline__ void fft16s(float2 *v)
{
#pragma unroll
  for(int b = 0; b < 4; b++) bf4<INV>(v[b * S], v[(4 + b) * S], v[(8 + b) * S], v[(12 + b) * S]);
  v[5 * S] = mul_w32<INV, 2, CT>(v[5 * S]);
  v[6 * S] = mul_w32<INV, 4, CT>(v[6 * S]);
  v[7 * S] = mul_w32<INV, 6, CT>(v[7 * S]);
  v[9 * S] = mul_w32<INV, 4, CT>(v[9 * S]);
  v[10 * S] = mul_w32<INV, 8, CT>(v[10 * S]);
  v[11 * S] = mul_w32<INV, 12, CT>(v[11 * S]);
  v[13 * S] = mul_w32<INV, 6, CT>(v[13 * S]);
  v[14 * S] = mul_w32<INV, 12, CT>(v[14 * S]);
  v[15 * S] = mul_w32<INV, 18, CT>(v[15 * S]);
#pragma unroll
  for(int k0 = 0; k0 < 4; k0++) bf4<INV>(v[(4 * k0) * S], v[(4 * k0 + 1) * S], v[(4 * k0 + 2) * S], v[(4 * k0 + 3) * S]);
  float2 t;
#define OLS_SWAP(i, j) t = v[(i) * S]; v[(i) * S] = v[(j) * S]; v[(j) * S] = t;
  OLS_SWAP(1, 4) OLS_SWAP(2, 8) OLS_SWAP(3, 12) OLS_SWAP(6, 9) OLS_SWAP(7, 13) OLS_SWAP(11, 14)
#undef OLS_SWAP
}
// X[k] = E[k] + W32^k O[k], X[k+16] = E[k] - W32^k O[k] as four packed FMAs
template<bool INV, int K, bool CT = false> __device__ __forceinline__ void comb32(float2 e, float2 o, float2 &lo, float2 &hi)
{
  if(K == 0)
  {
    lo = add2(e, o);
    hi = sub2(e, o);
    return;
  }
  if(K == 8)
  {
    const float2 os = make_float2(o.y, o.x);
    lo = fma2(os, INV ? make_float2(-1.f, 1.f) : make_float2(1.f, -1.f), e);
    hi = fma2(os, INV ? make_float2(1.f, -1.f) : make_float2(-1.f, 1.f), e);
    return;
  }
  const float2 ox = bcast2(o.x), oy = bcast2(o.y);
  if(CT && OLS_CONST_TW)
  {
    const float4 t = c_ols_w32[INV ? 1 : 0][K];
    const float2 w = make_float2(t.x, t.y), rw = make_float2(t.z, t.w);
    lo = fma2(oy, rw, fma2(ox, w, e));
    hi = fma2(oy, make_float2(-rw.x, -rw.y), fma2(ox, make_float2(-w.x, -w.y), e));
    return;
  }
  constexpr float wr = ols_cos32(K), wi = INV ? ols_sin32(K) : -ols_sin32(K);
  lo = fma2(oy, make_float2(-wi, wr), fma2(ox, make_float2(wr, wi), e));
  hi = fma2(oy, make_float2(wi, -wr), fma2(ox, make_float2(-wr, -wi), e));
}
// 32-point DFT of v[0..32), natural order in and out
template<bool INV> __device__ __forceinline__ void fft32(float2 (&v)[32])
{
  fft16s<INV, 2>(&v[0]);   // E[k] at v[2k]
  fft16s<INV, 2>(&v[1]);   // O[k] at v[2k+1]
  float2 r[32];
#define OLS_CB(K) comb32<INV, K>(v[2 * K], v[2 * K + 1], r[K], r[K + 16]);
  OLS_CB(0) OLS_CB(1) OLS_CB(2) OLS_CB(3) OLS_CB(4) OLS_CB(5) OLS_CB(6) OLS_CB(7)
  OLS_CB(8) OLS_CB(9) OLS_CB(10) OLS_CB(11) OLS_CB(12) OLS_CB(13) OLS_CB(14) OLS_CB(15)
#undef OLS_CB
#pragma unroll
  for(int i = 0; i < 32; i++) v[i] = r[i];
}

// ---- complex products with run-time twiddles held as plain (re, im): four scalar FMA-pipe instructions ----------
__device__ __forceinline__ float2 cmul_s(float2 a, float wr, float wi)
{
  return make_float2(fmaf(-a.y, wi, a.x * wr), fmaf(a.y, wr, a.x * wi));
}
__device__ __forceinline__ float2 cmulc_s(float2 a, float wr, float wi)
{
  return make_float2(fmaf(a.y, wi, a.x * wr), fmaf(-a.x, wi, a.y * wr));
}

// 16 consecutive TMEM columns of this thread's lane -> 16 registers
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16])
{
  uint32_t u[16];
  asm volatile(
    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
      "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
    : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for(int i = 0; i < 16; i++) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// v[i] = v[i] (*) w[i] for the 32 constants of this thread at TMEM columns [col, col + 64)
template<bool CONJ> __device__ __forceinline__ void mul_tmem32(float2 (&v)[32], uint32_t taddr)
{
#pragma unroll
  for(int q = 0; q < 4; q++)
  {
    float w[16];
    tmem_ld16(taddr + 16 * q, w);
#pragma unroll
    for(int i = 0; i < 8; i++)
      v[8 * q + i] = CONJ ? cmulc_s(v[8 * q + i], w[2 * i], w[2 * i + 1]) : cmul_s(v[8 * q + i], w[2 * i], w[2 * i + 1]);
  }
}

// the same for 16 values at 32 consecutive TMEM columns
template<bool CONJ> __device__ __forceinline__ void mul_tmem16(float2 *v, uint32_t taddr)
{
#pragma unroll
  for(int q = 0; q < 2; q++)
  {
    float w[16];
    tmem_ld16(taddr + 16 * q, w);
#pragma unroll
    for(int i = 0; i < 8; i++)
      v[8 * q + i] = CONJ ? cmulc_s(v[8 * q + i], w[2 * i], w[2 * i + 1]) : cmul_s(v[8 * q + i], w[2 * i], w[2 * i + 1]);
  }
}
__device__ __forceinline__ float2 lds64(uint32_t addr)
{
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, float2 v)
{
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
// swizzled accesses of the warp-local transposes: address = (base ^ X) + OFF with compile-time X, OFF.  The XOR sits
// inside the asm so that the compiler recomputes it (one ALU-pipe LOP3) instead of keeping 64 addresses alive.
template<int X, int OFF> __device__ __forceinline__ float2 lds64x(uint32_t base)
{
  float2 v;
  asm volatile("{\n\t.reg .u32 t;\n\txor.b32 t, %2, %3;\n\tld.shared.v2.f32 {%0, %1}, [t+%4];\n\t}"
               : "=f"(v.x), "=f"(v.y) : "r"(base), "n"(X), "n"(OFF));
  return v;
}
template<int X, int OFF> __device__ __forceinline__ void sts64x(uint32_t base, float2 v)
{
  asm volatile("{\n\t.reg .u32 t;\n\txor.b32 t, %0, %1;\n\tst.shared.v2.f32 [t+%2], {%3, %4};\n\t}"
               ::"r"(base), "n"(X), "n"(OFF), "f"(v.x), "f"(v.y) : "memory");
}
template<int R, int END> struct OlsX
{
  // rows R..END-1 of the row-static pattern (row r at r*256, column lane ^ r) and of the lane = row pattern
  static __device__ __forceinline__ void st_rows(uint32_t bl, const float2 (&v)[32])
  {
    sts64x<R * 8, R * 256>(bl, v[R]);
    OlsX<R + 1, END>::st_rows(bl, v);
  }
  static __device__ __forceinline__ void ld_rows(uint32_t bl, float2 (&v)[32])
  {
    v[R] = lds64x<R * 8, R * 256>(bl);
    OlsX<R + 1, END>::ld_rows(bl, v);
  }
  static __device__ __forceinline__ void ld_cols_even(uint32_t al, float2 (&v)[32])
  {
    if((R & 1) == 0) v[R] = lds64x<R * 8, 0>(al);
    OlsX<R + 1, END>::ld_cols_even(al, v);
  }
  static __device__ __forceinline__ void ld_cols_odd(uint32_t al, float2 (&v)[32])
  {
    if(R & 1) v[R] = lds64x<R * 8, 0>(al);
    OlsX<R + 1, END>::ld_cols_odd(al, v);
  }
};
template<int END> struct OlsX<END, END>
{
  static __device__ __forceinline__ void st_rows(uint32_t, const float2 (&)[32]) {}
  static __device__ __forceinline__ void ld_rows(uint32_t, float2 (&)[32]) {}
  static __device__ __forceinline__ void ld_cols_even(uint32_t, float2 (&)[32]) {}
  static __device__ __forceinline__ void ld_cols_odd(uint32_t, float2 (&)[32]) {}
};
// mbarrier wait that lets the hardware suspend the warp until the phase completes (or the hint expires) instead of
// re-issuing try_wait every few cycles: the spinning producer lane and the warps parked on row / column barriers were
// taking ~12 % of the issue slots of the kernel (ncu: 146 M TRYWAIT executions per 4 ms launch)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, unsigned parity)
{
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "WAIT_%=:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
    "@p bra DONE_%=;\n\t"
    "bra WAIT_%=;\n\t"
    "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// OLS_L1PF = 1 (experiment): the O overlap samples of a window are L2 hits whose latency (~700 cycles) sits at the head of each
// P1 round; prefetch.global.L1 brings the NEXT round's lines into the small L1 that is left beside the shared memory while
// the current round computes
#ifndef OLS_L1PF
#define OLS_L1PF 0
#endif
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// profiling aid: lane 0 of every math warp of CTA 0 records the SM clock at phase boundaries of a few blocks
constexpr int OLS_PROF_IT0 = 8, OLS_PROF_NIT = 6, OLS_PROF_PTS = 14;
__device__ __forceinline__ void ols_stamp(const OlsParams &p, unsigned it, int w, int l, int pt)
{
  if(p.prof && blockIdx.x == 0 && l == 0 && it - OLS_PROF_IT0 < (unsigned) OLS_PROF_NIT)
  {
    unsigned c;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
    p.prof[((it - OLS_PROF_IT0) * OLS_MATH_WARPS + w) * OLS_PROF_PTS + pt] = c;
  }
}

// output stores: streaming (L1 no-allocate) by default; OLS_STG = 1 plain, 2 evict-first (.cs) for A/B runs
#ifndef OLS_STG
#define OLS_STG 0
#endif
__device__ __forceinline__ void ols_stg(float2 *p, float2 v)
{
#if OLS_STG == 0
  stg_stream(p, v);
#elif OLS_STG == 1
  *p = v;
#else
  __stcs(p, v);
#endif
}

// Shared-memory map (dynamic): [E : 128 KiB][S : L*8 + 16][barriers]
//   E  exchange buffer, element (k1, n2, n3) at sample index (k1*16 + n2)*32 + n3; its first O*8 + 16 bytes double as
//      the landing zone X of the next window's overlap samples while the block is not in E
//   S  the L new samples of the next window (+ 2 samples of alignment slack)
template<int OQ>   // overlap in quarters of M: O = 4096 * OQ
__global__ void __launch_bounds__(OLS_THREADS, 1) ols16k_kernel(OlsParams p)
{
  constexpr int O = 4096 * OQ, L = OLS_M - O, NX = 8 * OQ;   // window samples n1 < NX live in X
  constexpr uint32_t S_BYTES = L * 8 + 16;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sE = smem_u32(smem), sS = sE + OLS_E_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OLS_E_BYTES + S_BYTES);
  uint64_t *s_full = bars, *w_free = bars + 2, *e_free = bars + 1;   // e_free[r] at bars + 1 + 2 r (r = 0, 1): columns 8r..8r+7 of E read
  uint64_t *row_full = bars + 4;   // [16]: rows 2w', 2w'+1 of E written by every warp -> their owner may start P2
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4 + OLS_MATH_WARPS);
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;

  if(tid == 0)
  {
    mbar_init(s_full, 1);
    mbar_init(w_free, OLS_MATH_WARPS);
    mbar_init(e_free, OLS_MATH_WARPS / 2);
    mbar_init(e_free + 2, OLS_MATH_WARPS / 2);
    for(int i = 0; i < OLS_MATH_WARPS; i++) mbar_init(row_full + i, OLS_MATH_WARPS);
    mbar_fence_init();
  }
  if(w == 0)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if((sE & 255u) != 0) __trap();   // the swizzled addressing relies on a 256-byte aligned base

  // contiguous share of the (channel-major) internal blocks
  const long long first = p.total * blockIdx.x / gridDim.x, last = p.total * (blockIdx.x + 1) / gridDim.x;

  if(w >= OLS_MATH_WARPS)
  {
    // ===== producer: one lane feeds the window of every block through the TMA unit =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(OLS_PROD_REGS));
    if(tid == OLS_MATH_THREADS)
    {
      int chan = (int) (first / p.jblocks), j = (int) (first - (long long) chan * p.jblocks) - 1;
      for(long long b = first; b < last; b++)
      {
        const unsigned it = (unsigned) (b - first);
        if(++j == p.jblocks) { j = 0; chan++; }
        const long long pos0 = (long long) j * L + p.base;
        const long long a0 = pos0 & ~1LL;                          // 16-byte aligned start
        const bool fast = p.aligned && a0 >= 0 && a0 + OLS_M + 2 <= p.n;
        const float2 *src = p.x + (long long) chan * p.x_stride + a0;
        if(it > 0) mbar_wait_sleep(w_free, (it - 1) & 1);                // S consumed by the input warps (both rounds)
        if(fast)
        {
          mbar_expect_tx(s_full, S_BYTES);
#pragma unroll 1
          for(uint32_t off = 0; off < S_BYTES; off += OLS_PIECE)
          {
            const uint32_t nb = (S_BYTES - off < OLS_PIECE + 4096) ? S_BYTES - off : OLS_PIECE;
            bulk_g2s(smem + OLS_E_BYTES + off, reinterpret_cast<const unsigned char *>(src + O) + off, nb, s_full);
            if(nb != OLS_PIECE) break;
          }
        }
        else mbar_arrive_cta(s_full);
      }
    }
  }
  else
  {
    // ===== math warps =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(OLS_MATH_REGS));
    // per-thread constants -> tensor memory: columns [128 (w>>2), +64) gains, [+64, +128) W_M twiddles
    const uint32_t tm = tmem_base + ((uint32_t) (32 * (w & 3)) << 16) + 128u * (uint32_t) (w >> 2);
    {
      const float4 *src = p.tmem_init + (size_t) tid * 32;
#pragma unroll
      for(int q = 0; q < 8; q++)
      {
        float r[16];
#pragma unroll
        for(int i = 0; i < 4; i++)
        {
          const float4 t = __ldg(src + 4 * q + i);
          r[4 * i] = t.x;
          r[4 * i + 1] = t.y;
          r[4 * i + 2] = t.z;
          r[4 * i + 3] = t.w;
        }
        tc::tmem_st16(tm + 16 * q, r);
      }
      tmem_wait_st();
    }
    const uint32_t sEw = sE + (uint32_t) w * 8192u;   // this warp's rows of E (k1 = 2w, 2w+1)
    const uint32_t lx = (uint32_t) l * 8u;
    // W512^(n2 * k1), k1 = 2w + c, is the same for every lane: lane r = c*16 + n2 keeps entry r in two registers and the
    // warp spreads the 32 entries through the first 256 bytes of its own rows of E (free while the data sits in
    // registers) right before each use: 1 STS + 16 broadcast LDS.128 instead of 30 constant-bank loads
#if !OLS_TW1_OUTER
    const float2 tw1c = c_ols_tw1[(2 * w) * 16 + l];
#endif

    // The sixteen math warps share P2 (a warp owns two rows of E), but the two phases that touch all of E are split between
    // two ROLES so that they run at the same time instead of one after the other: the output warps (0-3, 8-11: two per
    // scheduler) read block b out of E, finish its inverse transform and store it (P3), while the input warps (4-7, 12-15)
    // already load the window of block b+1 and run its first pass (P1); each role takes the sixteen n2 columns in two
    // rounds.  The shared-memory / store traffic of one role overlaps the butterflies of the other.
    const bool out_role = ((w >> 2) & 1) == 0;
    const int ri = (w & 3) | ((w >> 3) << 2);          // 0..7 within the role: columns n2 = ri, ri + 8

    // P1 of block (chan1, j1), iteration number it1: window -> radix-32 over n1 -> rows of E
    auto phase1 = [&](int chan1, int j1, unsigned it1) {
      const long long pos0 = (long long) j1 * L + p.base;
      const long long a0 = pos0 & ~1LL;
      const bool fast = p.aligned && a0 >= 0 && a0 + OLS_M + 2 <= p.n;
      const float2 *xc = p.x + (long long) chan1 * p.x_stride;
      const float2 *cr = p.carry + (long long) chan1 * p.carry_len + p.carry_len;
      mbar_wait_sleep(s_full, it1 & 1);
      // EXPERIMENT (TSDGPU_OLS_OFFSET=1, off by default: measured -2 %): start one step behind the output warps
      if(it1 > 0 && p.offset_roles) mbar_wait_sleep(e_free, (it1 - 1) & 1);
#pragma unroll 1
      for(int r = 0; r < 2; r++)
      {
        const int n2 = ri + 8 * r;
        float2 v[32];
        if(fast)
        {
          // the O overlap samples were loaded (as new samples) one block ago: straight from L2 into registers, issued
          // first so that they travel while the shared-memory loads run; the L new samples come from the staging
          const float2 *xo = xc + pos0 + 32 * n2 + l;
#pragma unroll
          for(int h = 0; h < 2; h++)
#pragma unroll
            for(int n1 = h; n1 < NX; n1 += 2) v[n1] = __ldg(xo + 512 * n1);
          if(OLS_L1PF && r == 0 && (l & 15) == 0)
          {
            // round 1 of this warp: column n2 + 8, one prefetch per 128-byte line (lanes 0 and 16)
#pragma unroll
            for(int n1 = 0; n1 < NX; n1++) prefetch_l1(xo + 256 + 512 * n1);
          }
          const uint32_t sh = (uint32_t) (pos0 & 1) * 8u + (uint32_t) (32 * n2) * 8u + lx;
#pragma unroll
          for(int h = 0; h < 2; h++)      // even n1 first: the first radix-16 starts while the odd half is still arriving
#pragma unroll
            for(int n1 = NX + h; n1 < 32; n1 += 2) v[n1] = lds64(sS + sh + (uint32_t) (512 * (n1 - NX)) * 8u);
        }
        else
        {
          // edge block (start / end of the call, or unaligned rows): bounds-checked loads straight into registers
#pragma unroll
          for(int n1 = 0; n1 < 32; n1++)
          {
            const long long pos = pos0 + 512 * n1 + 32 * n2 + l;
            float2 val = make_float2(0.f, 0.f);
            if(pos >= 0) { if(pos < p.n) val = ldg_stream(xc + pos); }
            else if(pos >= -(long long) p.carry_len) val = __ldg(cr + pos);
            v[n1] = val;
          }
        }
        __syncwarp();
        if(l == 0) mbar_arrive_cta(w_free);            // staging consumed (16 arrivals: 8 warps x 2 rounds)
        if(r == 0 && it1 > 0) ols_stamp(p, it1 - 1, w, l, 1);
        fft16s<false, 2, OLS_CT_P1F>(&v[0]);
        fft16s<false, 2, OLS_CT_P1F>(&v[1]);
        if(r == 0 && it1 > 0) ols_stamp(p, it1 - 1, w, l, 2);
        if(it1 > 0) mbar_wait_sleep(e_free + 2 * r, (it1 - 1) & 1);  // the output warps have read these eight columns of the previous block
        if(r == 0 && it1 > 0) ols_stamp(p, it1 - 1, w, l, 3);
        // last radix-2 stage: rows k1 = K and K + 16 are stored as soon as they exist.  One arrival per row pair and
        // round: the owner of rows (2w', 2w'+1) starts P2 when all sixteen columns have been written.
        const uint32_t a = sE + (uint32_t) (32 * n2) * 8u + lx;
        const float2 *tw1 = c_ols_tw1 + 32 * n2;   // W512^(n2*k1), the same for every lane
#define OLS_P1(K)                                                                              \
        {                                                                                          \
          float2 lo, hi;                                                                           \
          comb32<false, K, OLS_CT_P1C>(v[2 * K], v[2 * K + 1], lo, hi);                            \
          if(OLS_TW1_OUTER)                                                                        \
          {                                                                                        \
            const float2 ta = tw1[K], tb = tw1[K + 16];                                            \
            if(K) lo = cmul_s(lo, ta.x, ta.y);                                                     \
            hi = cmul_s(hi, tb.x, tb.y);                                                           \
          }                                                                                        \
          sts64(a + (uint32_t) (K * 512) * 8u, lo);                                                \
          sts64(a + (uint32_t) ((K + 16) * 512) * 8u, hi);                                         \
          if(K & 1)                                                                                \
          {                                                                                        \
            __syncwarp();                                                                          \
            if(l == 0)                                                                             \
            {                                                                                      \
              mbar_arrive_cta(row_full + (K >> 1));                                                \
              mbar_arrive_cta(row_full + 8 + (K >> 1));                                            \
            }                                                                                      \
          }                                                                                        \
        }
        OLS_P1(0) OLS_P1(1) OLS_P1(2) OLS_P1(3) OLS_P1(4) OLS_P1(5) OLS_P1(6) OLS_P1(7)
        OLS_P1(8) OLS_P1(9) OLS_P1(10) OLS_P1(11) OLS_P1(12) OLS_P1(13) OLS_P1(14) OLS_P1(15)
#undef OLS_P1
        if(r == 0 && it1 > 0) ols_stamp(p, it1 - 1, w, l, 11);
      }
    };

    // P3 of block (chan3, j3): rows of E -> inverse radix-32 over k1 -> the L valid outputs
    auto phase3 = [&](int chan3, int j3, unsigned it3) {
#pragma unroll 1
      for(int r = 0; r < 2; r++)
      {
        const int n2 = ri + 8 * r;
        float2 v[32];
        const uint32_t a = sE + (uint32_t) (32 * n2) * 8u + lx;
#pragma unroll
        for(int h = 0; h < 2; h++)
#pragma unroll
          for(int k1 = h; k1 < 32; k1 += 2) v[k1] = lds64(a + (uint32_t) (k1 * 512) * 8u);
        __syncwarp();
        if(l == 0) mbar_arrive_cta(e_free + 2 * r);    // 8 arrivals per round: columns 8r..8r+7 may be overwritten
        if(OLS_TW1_OUTER)
        {
          const float2 *tw1 = c_ols_tw1 + 32 * n2;     // conj W512^(n2*k1)
#pragma unroll
          for(int h = 0; h < 2; h++)
#pragma unroll
            for(int k1 = h ? 1 : 2; k1 < 32; k1 += 2)
            {
              const float2 t = tw1[k1];
              v[k1] = cmulc_s(v[k1], t.x, t.y);
            }
        }
        if(r == 0) ols_stamp(p, it3, w, l, 1);
        const long long i0 = (long long) j3 * L + 32 * n2 + l;   // output index of window sample n = O + 32 n2 + l
        float2 *yc = p.y + (long long) chan3 * p.y_stride + i0;
        const int rem = (int) min(p.out_count - i0, (long long) L);   // outputs of this lane's column still inside the call
        // inverse radix-32 whose last stage hands every output pair (n1 = K, K + 16) to the store as soon as it exists
        fft16s<true, 2, OLS_CT_P3F>(&v[0]);
        fft16s<true, 2, OLS_CT_P3F>(&v[1]);
        if(r == 0) ols_stamp(p, it3, w, l, 2);
#define OLS_OUT(K)                                                                                   \
        {                                                                                                \
          float2 lo, hi;                                                                                 \
          comb32<true, K, OLS_CT_P3C>(v[2 * K], v[2 * K + 1], lo, hi);                                   \
          if(K >= NX && 512 * (K - NX) < rem) ols_stg(yc + 512 * (K - NX), lo);                          \
          if(K + 16 >= NX && 512 * (K + 16 - NX) < rem) ols_stg(yc + 512 * (K + 16 - NX), hi);           \
        }
        OLS_OUT(0) OLS_OUT(1) OLS_OUT(2) OLS_OUT(3) OLS_OUT(4) OLS_OUT(5) OLS_OUT(6) OLS_OUT(7)
        OLS_OUT(8) OLS_OUT(9) OLS_OUT(10) OLS_OUT(11) OLS_OUT(12) OLS_OUT(13) OLS_OUT(14) OLS_OUT(15)
#undef OLS_OUT
        if(r == 0) ols_stamp(p, it3, w, l, 11);
      }
    };

    int chan = (int) (first / p.jblocks), j = (int) (first - (long long) chan * p.jblocks);
    if(!out_role && first < last) phase1(chan, j, 0u);
    for(long long b = first; b < last; b++)
    {
      const unsigned it = (unsigned) (b - first);
      float2 v[32];
      ols_stamp(p, it, w, l, 0);
      mbar_wait_sleep(row_full + w, it & 1);
      ols_stamp(p, it, w, l, 4);

      // ---- P2: everything between the two exchanges, inside the warp's own 8 KiB ----
      // W512 twiddle table of the warp: row 31 of its region, each lane overwrites the element it has just read; the
      // row is rewritten last (by the final store of row 31), after the last table read
#pragma unroll
      for(int r = 0; r < 32; r++) v[r] = lds64(sEw + (uint32_t) (r * 32) * 8u + lx);
#if !OLS_TW1_OUTER
      const uint32_t tab = sEw + 31u * 256u;
      sts64(tab + lx, tw1c);
      __syncwarp();
#endif
#pragma unroll
      for(int c = 0; c < 2; c++)
      {
        // k1 = 2w + c: (W512 twiddle,) radix-16 over n2, W_M twiddle, rows c*16 + k2 into the transpose
#if !OLS_TW1_OUTER
#pragma unroll
        for(int i = 8 * c; i < 8 * c + 8; i++)
        {
          const float4 t = lds128(tab + 16u * i);
          if(i != 8 * c) v[2 * i] = cmul_s(v[2 * i], t.x, t.y);
          v[2 * i + 1] = cmul_s(v[2 * i + 1], t.z, t.w);
        }
#endif
        fft16s<false, 1>(&v[16 * c]);
        mul_tmem16<false>(&v[16 * c], tm + 64 + 32 * c);
        if(c == 0) OlsX<0, 16>::st_rows(sEw + lx, v);   // row r, column n3 = lane, at r*32 + (lane ^ r)
        else OlsX<16, 32>::st_rows(sEw + lx, v);
      }
      ols_stamp(p, it, w, l, 5);
      __syncwarp();
      {
        const uint32_t al = sEw + (uint32_t) l * 256u + lx;
        OlsX<0, 32>::ld_cols_even(al, v);      // row = lane, columns n3 = 0, 2, ..., 30
        OlsX<0, 32>::ld_cols_odd(al, v);
        ols_stamp(p, it, w, l, 6);
        fft32<false>(v);
        mul_tmem32<false>(v, tm);
        fft16s<true, 2>(&v[0]);
        fft16s<true, 2>(&v[1]);
        ols_stamp(p, it, w, l, 7);
#define OLS_T2(K)                                                                              \
        {                                                                                          \
          float2 lo, hi;                                                                           \
          comb32<true, K>(v[2 * K], v[2 * K + 1], lo, hi);                                         \
          sts64x<K * 8, 0>(al, lo);                                                                \
          sts64x<(K + 16) * 8, 0>(al, hi);                                                         \
        }
        OLS_T2(0) OLS_T2(1) OLS_T2(2) OLS_T2(3) OLS_T2(4) OLS_T2(5) OLS_T2(6) OLS_T2(7)
        OLS_T2(8) OLS_T2(9) OLS_T2(10) OLS_T2(11) OLS_T2(12) OLS_T2(13) OLS_T2(14) OLS_T2(15)
#undef OLS_T2
      }
      __syncwarp();
      OlsX<0, 32>::ld_rows(sEw + lx, v);
#if !OLS_TW1_OUTER
      sts64(tab + lx, tw1c);
      __syncwarp();
#endif
      ols_stamp(p, it, w, l, 8);
#pragma unroll
      for(int c = 0; c < 2; c++)
      {
        mul_tmem16<true>(&v[16 * c], tm + 64 + 32 * c);
        fft16s<true, 1>(&v[16 * c]);
#if !OLS_TW1_OUTER
#pragma unroll
        for(int i = 8 * c; i < 8 * c + 8; i++)
        {
          const float4 t = lds128(tab + 16u * i);
          if(i != 8 * c) v[2 * i] = cmulc_s(v[2 * i], t.x, t.y);
          v[2 * i + 1] = cmulc_s(v[2 * i + 1], t.z, t.w);
        }
#endif
#pragma unroll
        for(int r = 16 * c; r < 16 * c + 16; r++) sts64(sEw + (uint32_t) (r * 32) * 8u + lx, v[r]);
      }
      ols_stamp(p, it, w, l, 9);
      // the output warps need every row of the block; the input warps only announce theirs and move on to the next window
      if(out_role) tc::named_bar(2, OLS_MATH_THREADS);
      else asm volatile("bar.arrive 2, %0;" ::"n"(OLS_MATH_THREADS) : "memory");
      ols_stamp(p, it, w, l, 10);

      // ---- P3 of this block (output warps) alongside P1 of the next one (input warps) ----
      int chan_n = chan, j_n = j + 1;
      if(j_n == p.jblocks) { j_n = 0; chan_n++; }
      if(out_role) phase3(chan, j, it);
      else if(b + 1 < last) phase1(chan_n, j_n, it + 1);
      chan = chan_n;
      j = j_n;
      ols_stamp(p, it, w, l, 12);
    }
  }

  tc::fence_before();
  __syncthreads();
  if(w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// ---- host side ------------------------------------------------------------------------------------------------
static void fft_double(std::vector<std::complex<double>> &a, bool inverse)
{
  const size_t n = a.size();
  for(size_t i = 1, j = 0; i < n; i++)
  {
    size_t bit = n >> 1;
    for(; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if(i < j) std::swap(a[i], a[j]);
  }
  for(size_t len = 2; len <= n; len <<= 1)
  {
    const double ang = 2.0 * M_PI / (double) len * (inverse ? 1.0 : -1.0);
    std::vector<std::complex<double>> w(len / 2);
    for(size_t k = 0; k < len / 2; k++) w[k] = std::polar(1.0, ang * (double) k);
    for(size_t i = 0; i < n; i += len)
      for(size_t k = 0; k < len / 2; k++)
      {
        const std::complex<double> u = a[i + k], t = a[i + k + len / 2] * w[k];
        a[i + k] = u + t;
        a[i + k + len / 2] = u - t;
      }
  }
}


int ols16k_create(const float *H, int N, int K, Ols16k **out)
{
  *out = nullptr;
  if(K < 1 || K - 1 > 8192 || N < K) return 0;   // not served here: the caller keeps its N-point path
  // taps back from the gains: H = DFT(h2), h2 = [0^(N-K), h]  (fourier.cc:962-965)
  std::vector<std::complex<double>> a((size_t) N);
  for(int i = 0; i < N; i++) a[i] = std::complex<double>(H[2 * i], H[2 * i + 1]);
  fft_double(a, true);
  double e_in = 0, e_out = 0;
  for(int i = 0; i < N; i++)
  {
    a[i] /= (double) N;
    (i >= N - K ? e_in : e_out) += std::norm(a[i]);
  }
  // the caller's promise (fir_len) is checked: energy outside the K-tap support must be rounding noise
  if(!(e_out <= 1e-9 * e_in)) return 0;
  std::vector<std::complex<double>> taps((size_t) K);
  for(int m = 0; m < K; m++) taps[m] = a[(size_t) (N - K + m)];
  return ols16k_create_taps(taps.data(), K, out);
}

int ols16k_create_taps(const std::complex<double> *taps, int K, Ols16k **out)
{
  *out = nullptr;
  if(K < 1 || K - 1 > 8192) return 0;
  const int O = (K - 1 <= 4096) ? 4096 : 8192;
  std::vector<std::complex<double>> hm((size_t) OLS_M);
  for(int m = 0; m < K; m++) hm[m] = taps[m];
  fft_double(hm, false);
  // thread-major constants: thread t = 32 w + l owns gains k = (2w + (l>>4)) + 32 (l&15) + 512 k3 and the twiddles
  // W_M^(l * ((2w + c) + 32 k2)) at [c*16 + k2]
  std::vector<float> init((size_t) 512 * 128);
  for(int t = 0; t < 512; t++)
  {
    const int w = t >> 5, l = t & 31;
    float *dst = init.data() + (size_t) t * 128;
    for(int k3 = 0; k3 < 32; k3++)
    {
      const std::complex<double> g = hm[(size_t) ((2 * w + (l >> 4)) + 32 * (l & 15) + 512 * k3)] / (double) OLS_M;
      dst[2 * k3] = (float) g.real();
      dst[2 * k3 + 1] = (float) g.imag();
    }
    for(int c = 0; c < 2; c++)
      for(int k2 = 0; k2 < 16; k2++)
      {
        const long long e = ((long long) l * ((2 * w + c) + 32 * k2)) % OLS_M;
        const double ang = -2.0 * M_PI * (double) e / (double) OLS_M;
        dst[64 + 2 * (c * 16 + k2)] = (float) cos(ang);
        dst[64 + 2 * (c * 16 + k2) + 1] = (float) sin(ang);
      }
  }
  if(!rt().ols_ready)
  {
    std::vector<float2> t1(512);
    for(int k1 = 0; k1 < 32; k1++)
      for(int n2 = 0; n2 < 16; n2++)
      {
        const double ang = -2.0 * M_PI * (double) (k1 * n2) / 512.0;
        t1[OLS_TW1_OUTER ? n2 * 32 + k1 : k1 * 16 + n2] = make_float2((float) cos(ang), (float) sin(ang));
      }
    TSD_CUDA(cudaMemcpyToSymbol(c_ols_tw1, t1.data(), 512 * sizeof(float2)));
    std::vector<float4> w32(64);
    for(int k = 0; k < 32; k++)
    {
      const float wr = ols_cos32(k), ws = ols_sin32(k);
      w32[k] = make_float4(wr, -ws, ws, wr);        // forward: w = (cos, -sin), i w = (sin, cos)
      w32[32 + k] = make_float4(wr, ws, -ws, wr);   // inverse: conj
    }
    TSD_CUDA(cudaMemcpyToSymbol(c_ols_w32, w32.data(), 64 * sizeof(float4)));
    TSD_CUDA(cudaFuncSetAttribute(ols16k_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ols16k_smem_bytes(4096)));
    TSD_CUDA(cudaFuncSetAttribute(ols16k_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ols16k_smem_bytes(8192)));
    rt().ols_ready = true;
  }
  auto *o = new Ols16k;
  o->O = O;
  o->L = OLS_M - O;
  o->K = K;
  cudaError_t e = cudaMalloc(&o->d_init, init.size() * sizeof(float));
  if(e == cudaSuccess) e = cudaMemcpy(o->d_init, init.data(), init.size() * sizeof(float), cudaMemcpyHostToDevice);
  if(e != cudaSuccess)
  {
    ols16k_destroy(o);
    return fail(std::string("ols16k_create: ") + cudaGetErrorString(e));
  }
  *out = o;
  return 0;
}

void ols16k_destroy(Ols16k *o)
{
  if(!o) return;
  if(o->d_init) cudaFree(o->d_init);
  delete o;
}

int ols16k_smem_bytes(int O) { return OLS_E_BYTES + (OLS_M - O) * 8 + 16 + 32 + 8 * OLS_MATH_WARPS + 16; }

// y[c][i] = y_fir[t0 + i - delay], i in [0, out_count): window positions are relative to x[0] of this call, whose
// stream index is t0 + residual
int ols16k_run(Ols16k *o, const float2 *x, long long xs, int n, const float2 *carry, int carry_len, float2 *y, long long ys,
               long long out_count, int delay, int residual, int nchan)
{
  if(out_count <= 0) return 0;
  Runtime &r = rt();
  OlsParams p;
  p.x = x;
  p.y = y;
  p.carry = carry;
  p.tmem_init = reinterpret_cast<const float4 *>(o->d_init);
  p.x_stride = xs;
  p.y_stride = ys;
  p.out_count = out_count;
  p.carry_len = carry_len;
  p.n = n;
  p.base = -(delay + o->O + residual);
  p.jblocks = (int) ((out_count + o->L - 1) / o->L);
  p.total = (long long) nchan * p.jblocks;
  p.aligned = ((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (xs & 1) == 0) ? 1 : 0;
  const int grid = (int) std::min<long long>(r.num_sms, p.total);
  const int smem = ols16k_smem_bytes(o->O);
  p.prof = nullptr;
  p.offset_roles = getenv("TSDGPU_OLS_OFFSET") ? atoi(getenv("TSDGPU_OLS_OFFSET")) : 0;
  const char *prof_path = getenv("TSDGPU_OLS_PROF");   // profiling aid: clock trace of CTA 0 written to this file
  const size_t prof_words = (size_t) OLS_PROF_NIT * OLS_MATH_WARPS * OLS_PROF_PTS;
  if(prof_path)
  {
    TSD_CUDA(cudaMalloc(&p.prof, prof_words * sizeof(unsigned)));
    TSD_CUDA(cudaMemsetAsync(p.prof, 0, prof_words * sizeof(unsigned), r.stream));
  }
  {
  KernelTimer timer;
  if(o->O == 4096) ols16k_kernel<1><<<grid, OLS_THREADS, smem, r.stream>>>(p);
  else ols16k_kernel<2><<<grid, OLS_THREADS, smem, r.stream>>>(p);
  }
  TSD_LAUNCH_CHECK();
  if(prof_path)
  {
    std::vector<unsigned> h(prof_words);
    TSD_CUDA(cudaStreamSynchronize(r.stream));
    TSD_CUDA(cudaMemcpy(h.data(), p.prof, prof_words * sizeof(unsigned), cudaMemcpyDeviceToHost));
    cudaFree(p.prof);
    if(FILE *fp = fopen(prof_path, "wb"))
    {
      fwrite(h.data(), sizeof(unsigned), prof_words, fp);
      fclose(fp);
    }
  }
  return 0;
}

} // namespace tsdgpu
