// Register/shared-memory building blocks of the sm_100a FFT kernels.
//
// A "tile" is 16 independent 256-point transforms handled by one 256-thread CTA (4096 points,
// 32 KB of shared memory).  Each thread runs two radix-16 butterflies in registers per
// 256-point transform, with ONE shared-memory exchange in between.  Two access patterns:
//   COLS  the 16 transforms are the 16 lanes `lo = tid & 15` (strided in memory: element n of
//         lane lo is at n*pitch + lo), used for the column passes of the four-step FFT;
//   ROWS  the 16 transforms are 16 contiguous rows, used for the row passes.
// Both patterns keep global accesses in full 128-byte segments and shared accesses
// conflict-free (skewed layout for ROWS).
//
// Arithmetic: the reference computes a unitary radix-2 DFT in float32 with double-generated
// twiddles (fourier.cc:32-46,61-121).  Here the same DFT is evaluated with radix-16 butterflies;
// local twiddles come from a 256-entry shared table (sincospif on exact dyadic angles), the W_N
// four-step twiddles from sincospif + a power tree of depth <= 4, so the result differs from the
// reference by float rounding only (tests: <= 1e-5 of signal RMS).
#pragma once
#include "common.cuh"

namespace tsdgpu {

#define TSD_C1 0.92387953251128674f   // cos(pi/8)
#define TSD_S1 0.38268343236508977f   // sin(pi/8)
#define TSD_R2 0.70710678118654752f   // sqrt(1/2)

// 4-point DFT, natural order in and out: 6 packed add/sub + 2 packed FMA.
// (-i)*d = (d.y, -d.x) is folded into an FFMA2 on the half-swapped operand: s1 + swap(d) * (1, -1).
template<bool INV> __device__ __forceinline__ void fft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3)
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
// multiply by -i (forward) / +i (inverse)
template<bool INV> __device__ __forceinline__ float2 mul_mi(float2 a)
{
  return mul2(make_float2(a.y, a.x), INV ? make_float2(-1.f, 1.f) : make_float2(1.f, -1.f));
}

// v *= W16^m (forward) or conj (inverse), m a compile-time constant in {1,2,3,6,9}: the twiddle and its
// 90-degree rotation are immediates, the product is FMUL2 + FFMA2
template<bool INV, int M> __device__ __forceinline__ float2 mul_w16(float2 v)
{
  float wr, wi;   // forward value of W16^M = exp(-2 pi i M / 16)
  if(M == 1) { wr = TSD_C1; wi = -TSD_S1; }
  else if(M == 2) { wr = TSD_R2; wi = -TSD_R2; }
  else if(M == 3) { wr = TSD_S1; wi = -TSD_C1; }
  else if(M == 6) { wr = -TSD_R2; wi = -TSD_R2; }
  else { wr = -TSD_C1; wi = TSD_S1; }   // M == 9
  if(INV) wi = -wi;
  return cmul_rot(v, make_float2(wr, wi), make_float2(-wi, wr));
}

// 16-point DFT of v[0..15], natural order in and out, all indices static
template<bool INV> __device__ __forceinline__ void fft16(float2 (&v)[16])
{
  // radix-4 over a (n = 4a + b): Y[b][k0] lands in v[4*k0 + b]
#pragma unroll
  for(int b = 0; b < 4; b++)
  {
    float2 a0 = v[b], a1 = v[4 + b], a2 = v[8 + b], a3 = v[12 + b];
    fft4<INV>(a0, a1, a2, a3);
    v[b] = a0;
    v[4 + b] = a1;
    v[8 + b] = a2;
    v[12 + b] = a3;
  }
  // twiddles W16^(b*k0) on v[4*k0 + b]
  v[5] = mul_w16<INV, 1>(v[5]);
  v[6] = mul_w16<INV, 2>(v[6]);
  v[7] = mul_w16<INV, 3>(v[7]);
  v[9] = mul_w16<INV, 2>(v[9]);
  v[10] = mul_mi<INV>(v[10]);
  v[11] = mul_w16<INV, 6>(v[11]);
  v[13] = mul_w16<INV, 3>(v[13]);
  v[14] = mul_w16<INV, 6>(v[14]);
  v[15] = mul_w16<INV, 9>(v[15]);
  // radix-4 over b: X[k0 + 4*k1] lands in v[4*k0 + k1]
#pragma unroll
  for(int k0 = 0; k0 < 4; k0++) fft4<INV>(v[4 * k0], v[4 * k0 + 1], v[4 * k0 + 2], v[4 * k0 + 3]);
  // digit reversal back to natural order (register renaming only)
  float2 t;
#define TSD_SWAP(i, j) t = v[i]; v[i] = v[j]; v[j] = t;
  TSD_SWAP(1, 4) TSD_SWAP(2, 8) TSD_SWAP(3, 12) TSD_SWAP(6, 9) TSD_SWAP(7, 13) TSD_SWAP(11, 14)
#undef TSD_SWAP
}

// v[k] *= base * step^k for k = 0..15.  Power tree of depth <= 4 arranged so that only a handful of
// temporaries are live at a time (register pressure decides the CTAs per SM).
__device__ __forceinline__ void mul_geometric(float2 (&v)[16], float2 base, float2 step)
{
  const float2 s2 = cmul(step, step), s4 = cmul(s2, s2), s8 = cmul(s4, s4);
  float2 t[4];
  t[0] = base;
  t[1] = cmul(base, step);
  t[2] = cmul(base, s2);
  t[3] = cmul(t[1], s2);
#pragma unroll
  for(int i = 0; i < 4; i++)
  {
    v[i] = cmul(v[i], t[i]);
    v[4 + i] = cmul(v[4 + i], cmul(t[i], s4));
    const float2 u = cmul(t[i], s8);
    v[8 + i] = cmul(v[8 + i], u);
    v[12 + i] = cmul(v[12 + i], cmul(u, s4));
  }
}

// Four-step twiddles from the host-built tables (Runtime::tw4): stage A, thread (hi, column n2): v[p2] *= W^(n2*(hi + 16 p2));
// row stage of the inverse, thread (row k1, lo): v[pp] *= conj(W)^(k1*(lo + 16 pp)).  Two coalesced 8-byte loads
// replace two sincospif evaluations.
template<bool INV> __device__ __forceinline__ float2 tw4_load(const float2 *p)
{
  const float2 w = __ldg(p);
  return INV ? make_float2(w.x, -w.y) : w;
}
template<bool INV> __device__ __forceinline__ void mul_fourstep_cols(float2 (&v)[16], const float2 *tw4, int n2, int hi)
{
  mul_geometric(v, tw4_load<INV>(tw4 + hi * 256 + n2), tw4_load<INV>(tw4 + 8192 + n2));
}
template<bool INV> __device__ __forceinline__ void mul_fourstep_rows(float2 (&v)[16], const float2 *tw4, int k1, int lo)
{
  mul_geometric(v, tw4_load<INV>(tw4 + 4096 + k1 * 16 + lo), tw4_load<INV>(tw4 + 8192 + k1));
}

// Local twiddles of a 256-point transform from the shared table tw[k*16 + i] = {w, i*w} with
// w = exp(-2 pi i * i*k / 256), i, k in [0,16): v[k] *= w (conjugated for the inverse), one LDS.128 +
// FMUL2 + FFMA2 each.  For a fixed k the 16 lanes of a half-warp read either one entry (idx = hi:
// broadcast) or 16 consecutive entries (idx = lo).
// With TBL = true the table already holds the twiddles of this direction (conjugated for the inverse,
// fill_tw256_from), so that {w, i*w} are two aligned register pairs of the LDS.128 and no MOVs are needed.
template<bool INV, bool TBL = false> __device__ __forceinline__ void mul_table(float2 (&v)[16], const float4 *tw, int idx)
{
#pragma unroll
  for(int k = 1; k < 16; k++)
  {
    const float4 t = tw[k * 16 + idx];   // {w.x, w.y, -w.y, w.x}
    v[k] = (INV && !TBL) ? cmul_rot(v[k], make_float2(t.x, t.z), make_float2(t.y, t.x)) : cmul_rot(v[k], make_float2(t.x, t.y), make_float2(t.z, t.w));
  }
}
// host-built tables (tw256_host_table): [0..256) forward, [256..512) conjugated
__device__ __forceinline__ void fill_tw256_from(float4 *tw, const float4 *global_table, int tid, bool inverse)
{
  tw[tid] = __ldg(global_table + (inverse ? 256 : 0) + tid);
}
__device__ __forceinline__ void fill_tw256(float4 *tw, int tid)
{
  // 256 threads, one entry each
  const float2 w = twiddle<false>((unsigned) ((tid >> 4) * (tid & 15)), 2.0f / 256.0f);
  tw[tid] = make_float4(w.x, w.y, -w.y, w.x);
}

// ---- 256-point transform, COLS pattern -------------------------------------------------------
// in : thread (hi = tid>>4, lo = tid&15) holds v[j] = x_lo[16*j + hi]
// out: thread (hi, lo) holds v[k2] = X_lo[hi + 16*k2]
// The *_tail variants expect the first radix-16 pass (fft16 over j) to have been done by the caller, so
// that the software-pipelined kernels can issue the next item's global loads right after it.
template<bool INV, bool TBL = false> __device__ __forceinline__ void fft256_cols_tail(float2 (&v)[16], float2 *sm, const float4 *tw, int hi, int lo)
{
  mul_table<INV, TBL>(v, tw, hi);   // W256^(hi*k1)
#pragma unroll
  for(int k1 = 0; k1 < 16; k1++) sm[(hi * 16 + k1) * 16 + lo] = v[k1];
  __syncthreads();
#pragma unroll
  for(int a = 0; a < 16; a++) v[a] = sm[(a * 16 + hi) * 16 + lo];
  fft16<INV>(v);
}
template<bool INV, bool TBL = false> __device__ __forceinline__ void fft256_cols(float2 (&v)[16], float2 *sm, const float4 *tw, int hi, int lo)
{
  fft16<INV>(v);
  fft256_cols_tail<INV, TBL>(v, sm, tw, hi, lo);
}

// ---- 256-point transform, ROWS pattern -------------------------------------------------------
// in : thread (hi = row r, lo = b) holds v[j] = x_r[16*j + b]
// out: thread (hi = k1, lo = row r) holds v[k2] = X_r[k1 + 16*k2]
template<bool INV, bool TBL = false> __device__ __forceinline__ void fft256_rows_a_tail(float2 (&v)[16], float2 *sm, const float4 *tw, int hi, int lo)
{
  mul_table<INV, TBL>(v, tw, lo);   // W256^(b*k1)
#pragma unroll
  for(int k1 = 0; k1 < 16; k1++) sm[k1 * 256 + lo * 16 + ((hi + lo) & 15)] = v[k1];
  __syncthreads();
#pragma unroll
  for(int b = 0; b < 16; b++) v[b] = sm[hi * 256 + b * 16 + ((lo + b) & 15)];
  fft16<INV>(v);
}
template<bool INV, bool TBL = false> __device__ __forceinline__ void fft256_rows_a(float2 (&v)[16], float2 *sm, const float4 *tw, int hi, int lo)
{
  fft16<INV>(v);
  fft256_rows_a_tail<INV, TBL>(v, sm, tw, hi, lo);
}
// in : thread (hi = k1, lo = row r) holds v[k2] = X_r[k1 + 16*k2]
// out: thread (hi = row r, lo = q) holds v[p] = x_r[16*p + q]
template<bool INV, bool TBL = false> __device__ __forceinline__ void fft256_rows_b(float2 (&v)[16], float2 *sm, const float4 *tw, int hi, int lo)
{
  fft16<INV>(v);               // over k2 -> q
  mul_table<INV, TBL>(v, tw, hi);   // W256^(k1*q)
#pragma unroll
  for(int q = 0; q < 16; q++) sm[hi * 256 + q * 16 + ((lo + q) & 15)] = v[q];
  __syncthreads();
#pragma unroll
  for(int k1 = 0; k1 < 16; k1++) v[k1] = sm[k1 * 256 + lo * 16 + ((hi + lo) & 15)];
  fft16<INV>(v);               // over k1 -> p
}

// ---- persistent-kernel plumbing ---------------------------------------------------------------
// Every warp publishes its own completion (its stores -> __syncwarp -> fence -> red.release), so no
// CTA-wide barrier is needed at the end of an item; a finished item counts ITEM_WARPS per tile.
constexpr unsigned ITEM_WARPS = 8;
__device__ __forceinline__ void warp_release(unsigned *flag)
{
  __syncwarp();
  // red.release.gpu orders every store that happens-before it (the warp's, via __syncwarp) — no extra
  // __threadfence, which would add a second MEMBAR and an L1 invalidate (CCTL.IVALL) per item
  if((threadIdx.x & 31) == 0) red_release_add(flag, 1u);
}
__device__ __forceinline__ void spin_until(const unsigned *flag, unsigned target)
{
  while(ld_acquire(flag) < target) __nanosleep(32);
}
// global load that bypasses L1 (data written by other SMs in the same launch, or touched once)
__device__ __forceinline__ float2 ld_cg(const float2 *p)
{
  float2 v;
  asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}

} // namespace tsdgpu
