// FFT-domain block filter on sm_100a.  Replaces OLA<cfloat> / filtre_fft (reference
// fourier.cc:737-882,935-940) in plain mode with the spectral callback "X *= H"
// (fourier.cc:956-959) given as data.
//
// Bookkeeping is the reference's, computed on the host with the same expressions:
// Ne, N = p2(Ne + nb_zeros_min), N_zeros = N - Ne (fourier.cc:764-776), re-blocking of arbitrary
// input chunks into Ne-blocks with a residual (TamponNv2, tsd.cc:332-370), Ne samples out per
// completed block (fourier.cc:813-833).
//
// N = 65536 runs as ONE persistent kernel per step(): every Ne-block goes through three tile
// stages (fft_tiles.cuh) that hand over through an L2-resident scratch ring,
//   A  gather the N-point window, 256 column DFTs, W_N twiddles
//   B  256 row DFTs, multiply by H, 256 inverse row DFTs, conj(W_N) twiddles   (in place)
//   C  256 inverse column DFTs, emit the Ne output samples
// so a sample crosses HBM once on the way in and once on the way out.  Two output forms:
//   overlap-save  (fir_len = K > 0): window = the N inputs ending K-1 samples after the block
//                 start, plain stores — same samples as the reference's overlap-add for an H
//                 that is the transform of K taps placed at the tail (fourier.cc:962-965);
//   overlap-add   (fir_len = 0, any H): zero-padded block like the reference (fourier.cc:850),
//                 tail of block b and head of block b+1 meet through two-addend red.add on a
//                 zeroed region; the last block's tail is carried in `svg` (fourier.cc:870-872).
// Other N use an unfused path built on the generic FFT plan (gather, FFT, *H, IFFT, overlap-add).
#include "fft_tiles.cuh"
#include "fft_plan.h"
#include "host_pipe.cuh"
#include "ols16k.h"
#include "tsdgpu.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace tsdgpu {

struct OlaParams
{
  const float2 *x;        // this call's input, [nchan][x_stride]
  float2 *y;              // this call's output, [nchan][y_stride]
  const float2 *carry;    // [nchan][carry_len] samples preceding x[0]
  float2 *svg;            // [nchan][Ne] overlap-add carry (OLA form)
  const float2 *H;        // [65536], pre-scaled by 1/N
  float2 *scratch;        // [ring][65536]
  unsigned *done_a, *done_b, *done_c, *ticket;
  long long x_stride, y_stride;
  int carry_len;
  int nblocks;            // Ne-blocks per channel in this call
  int Q;                  // nchan * nblocks
  int Ne, Nz;
  int base_off;           // window position of element 0 relative to (block start - residual)
  int zero_below;         // window elements n < zero_below are zero (OLA padding)
  int residual;           // samples already buffered before x[0]
  int out_shift;          // OLS: output i = j - out_shift
  int ola_form;           // 0 overlap-save, 1 overlap-add
  int ring, lag;
  int n;                  // samples per channel in this call
  const float2 *tw4;      // rt().tw4: four-step twiddle tables
};

constexpr int OLA_NT = 256;

__device__ __forceinline__ void red_add_f2(float2 *p, float2 v)
{
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(&p->x), "f"(v.x) : "memory");
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(&p->y), "f"(v.y) : "memory");
}

__global__ void __launch_bounds__(OLA_NT, 3) ola64k_kernel(OlaParams p)
{
  __shared__ float2 sm[4096];
  __shared__ float4 tw[256];
  __shared__ unsigned s_ticket[2];
  const int tid = threadIdx.x, hi = tid >> 4, lo = tid & 15;
  const unsigned total = (unsigned) (p.Q + 2 * p.lag) * 48u;
  const unsigned full = 16u * ITEM_WARPS;
  fill_tw256(tw, tid);
  if(tid == 0) s_ticket[0] = atomicAdd(p.ticket, 1u);
  __syncthreads();

  for(int it = 0;; it ^= 1)
  {
    const unsigned ticket = s_ticket[it];
    if(ticket >= total) break;
    const int s = (int) (ticket / 48u), sub = (int) (ticket - (unsigned) s * 48u);
    const int g = sub & 15, stage = sub >> 4;
    const int q = s - stage * p.lag;
    const bool valid = q >= 0 && q < p.Q;
    if(tid == 0)
    {
      s_ticket[it ^ 1] = atomicAdd(p.ticket, 1u);   // next item, fetched early
      if(valid)
      {
        if(stage == 1) spin_until(p.done_a + q, full);
        else if(stage == 2) spin_until(p.done_b + q, full);
        else if(q >= p.ring) spin_until(p.done_c + (q - p.ring), full);   // ring slot free again
      }
    }
    __syncthreads();
    if(!valid) continue;
    const int chan = q / p.nblocks, blk = q - chan * p.nblocks;
    float2 *sc = p.scratch + (long long) (q % p.ring) * 65536;
    float2 v[16];

    if(stage == 0)
    {
      // window element n sits at position pos0 + n relative to x[0]
      const long long pos0 = (long long) blk * p.Ne - p.residual + p.base_off;
      const float2 *x = p.x + (long long) chan * p.x_stride;
      const int n0 = hi * 256 + 16 * g + lo;
      if(pos0 >= 0 && p.zero_below == 0)
      {
        const float2 *xw = x + pos0 + n0;
#pragma unroll
        for(int j = 0; j < 16; j++) v[j] = ldg_stream(xw + j * 4096);
      }
      else
      {
        // block at the start of the call: part of the window is carried history or zero padding
        const float2 *cr = p.carry + (long long) chan * p.carry_len + p.carry_len;
#pragma unroll
        for(int j = 0; j < 16; j++)
        {
          const int n = n0 + j * 4096;
          const long long pos = pos0 + n;
          float2 val = make_float2(0.f, 0.f);
          if(n >= p.zero_below) val = (pos >= 0) ? ldg_stream(x + pos) : __ldg(cr + pos);
          v[j] = val;
        }
      }
      fft256_cols<false>(v, sm, tw, hi, lo);
      mul_fourstep_cols<false>(v, p.tw4, 16 * g + lo, hi);
      float2 *dst = sc + n0;
#pragma unroll
      for(int p2 = 0; p2 < 16; p2++) dst[p2 * 4096] = v[p2];
      warp_release(p.done_a + q);
    }
    else if(stage == 1)
    {
      float2 *row = sc + (16 * g + hi) * 256 + lo;
#pragma unroll
      for(int j = 0; j < 16; j++) v[j] = __ldcg(row + 16 * j);
      fft256_rows_a<false>(v, sm, tw, hi, lo);
      // thread (hi = k', lo = r): v[k2] = X[k], k = (16g + r) + 256*(k' + 16*k2)
      const float2 *H = p.H + hi * 256 + 16 * g + lo;
#ifndef TSD_EXPERIMENT_NO_H
#pragma unroll
      for(int k2 = 0; k2 < 16; k2++) v[k2] = cmul(v[k2], __ldg(H + k2 * 4096));
#else
      (void) H;   // TIMING EXPERIMENT ONLY: no spectral gain
#endif
      __syncthreads();   // exchange buffer is reused
      fft256_rows_b<true>(v, sm, tw, hi, lo);
      // thread (hi = r, lo = q'): v[pp] = b[k1 = 16g + r][n2 = 16*pp + q']; conj four-step twiddle
      mul_fourstep_rows<true>(v, p.tw4, 16 * g + hi, lo);
#pragma unroll
      for(int pp = 0; pp < 16; pp++) row[16 * pp] = v[pp];
      warp_release(p.done_b + q);
    }
    else
    {
      const float2 *col = sc + hi * 256 + 16 * g + lo;
#pragma unroll
      for(int j = 0; j < 16; j++) v[j] = __ldcg(col + j * 4096);
      fft256_cols<true>(v, sm, tw, hi, lo);
      // thread (hi = p1, lo): v[p2] = x2[256*(p1 + 16*p2) + 16g + lo]
      float2 *y = p.y + (long long) chan * p.y_stride;
      const int j0 = hi * 256 + 16 * g + lo;
      if(!p.ola_form)
      {
        const int i0 = j0 - p.out_shift;
        float2 *yb = y + (long long) blk * p.Ne + i0;
#pragma unroll
        for(int p2 = 0; p2 < 16; p2++)
          if((unsigned) (i0 + p2 * 4096) < (unsigned) p.Ne) stg_stream(yb + p2 * 4096, v[p2]);
      }
      else
      {
        const bool last = (blk + 1 == p.nblocks);
        float2 *svg = p.svg + (long long) chan * p.Ne;
#pragma unroll
        for(int p2 = 0; p2 < 16; p2++)
        {
          const int j = j0 + p2 * 4096;
          if(j < p.Nz)
            red_add_f2(y + (long long) blk * p.Ne + (p.Ne - p.Nz) + j, v[p2]);   // tail of block blk
          else if(last)
            svg[j - p.Nz] = v[p2];                                                 // carried (fourier.cc:872)
          else if(j < p.Ne)
            stg_stream(y + (long long) (blk + 1) * p.Ne + (j - p.Nz), v[p2]);
          else
            red_add_f2(y + (long long) (blk + 1) * p.Ne + (j - p.Nz), v[p2]);     // meets head of block blk+1
        }
      }
      warp_release(p.done_c + q);
    }
  }
}

// ---- staged form: the hand-over between stages is a kernel boundary ---------------------------------
// Same three tile stages as the persistent kernel, but without tickets, flags, spins or fences on the
// math warps: a chunk of `chunk` blocks runs one kernel per stage, back to back on one auxiliary stream,
// with its own L2-resident scratch (chunk x 512 KiB); consecutive chunks go round-robin to `nstreams`
// streams so that the kernels of different chunks overlap and fill each other's tails.  Stages A and C
// run at 64 registers (4 CTAs per SM), stage B at 80 (3 CTAs per SM).
// Measured alternatives: A/B/C of different chunks chained by events on three prioritised
// streams (CPU-bound on the event calls), the same captured in a CUDA graph (node-to-node dependency latency
// eats the gain), one launch per pipeline slot with interleaved A/B/C CTAs on one stream (tail at every
// launch boundary), L2 bulk prefetch of the next chunks' input from stage B (no gain).
struct OlaRole
{
  float2 *scratch;        // this chunk's scratch, [nb][65536]
  int q0, nb;             // first block (channel-major block index) and number of blocks of the chunk
};
struct OlaStageParams
{
  OlaParams o;
  const float4 *tw;       // rt().tw256
  OlaRole role[1];
};

template<int STAGE>
__device__ __forceinline__ void ola_stage_body(const OlaParams &p, const float4 *gtw, float2 *sc, int q, int g, float2 *sm, float4 *tw)
{
  const int tid = threadIdx.x, hi = tid >> 4, lo = tid & 15;
  const int chan = q / p.nblocks, blk = q - chan * p.nblocks;
  float2 v[16];
  if(STAGE == 0)
  {
    // window element n sits at position pos0 + n relative to x[0]
    const long long pos0 = (long long) blk * p.Ne - p.residual + p.base_off;
    const float2 *x = p.x + (long long) chan * p.x_stride;
    const int n0 = hi * 256 + 16 * g + lo;
    if(pos0 >= 0 && p.zero_below == 0)
    {
      const float2 *xw = x + pos0 + n0;
#pragma unroll
      for(int j = 0; j < 16; j++) v[j] = ldg_stream(xw + j * 4096);
    }
    else
    {
      // block at the start of the call: part of the window is carried history or zero padding
      const float2 *cr = p.carry + (long long) chan * p.carry_len + p.carry_len;
#pragma unroll
      for(int j = 0; j < 16; j++)
      {
        const int n = n0 + j * 4096;
        const long long pos = pos0 + n;
        float2 val = make_float2(0.f, 0.f);
        if(n >= p.zero_below) val = (pos >= 0) ? ldg_stream(x + pos) : __ldg(cr + pos);
        v[j] = val;
      }
    }
    fill_tw256_from(tw, gtw, tid, false);
    const float2 tb = tw4_load<false>(p.tw4 + hi * 256 + 16 * g + lo), ts = tw4_load<false>(p.tw4 + 8192 + 16 * g + lo);
    __syncthreads();
    fft256_cols<false, true>(v, sm, tw, hi, lo);
    mul_geometric(v, tb, ts);   // four-step twiddle W_N^(n2*k1)
    float2 *dst = sc + n0;
#pragma unroll
    for(int p2 = 0; p2 < 16; p2++) dst[p2 * 4096] = v[p2];
  }
  else if(STAGE == 1)
  {
    float2 *row = sc + (16 * g + hi) * 256 + lo;
#pragma unroll
    for(int j = 0; j < 16; j++) v[j] = __ldcg(row + 16 * j);
    fill_tw256_from(tw, gtw, tid, false);
    fill_tw256_from(tw + 256, gtw, tid, true);
    const float2 tb = tw4_load<true>(p.tw4 + 4096 + (16 * g + hi) * 16 + lo), ts = tw4_load<true>(p.tw4 + 8192 + 16 * g + hi);
    __syncthreads();
    fft256_rows_a<false, true>(v, sm, tw, hi, lo);
    // thread (hi = k', lo = r): v[k2] = X[k], k = (16g + r) + 256*(k' + 16*k2)
    const float2 *H = p.H + hi * 256 + 16 * g + lo;
#pragma unroll
    for(int k2 = 0; k2 < 16; k2++) v[k2] = cmul(v[k2], __ldg(H + k2 * 4096));
    __syncthreads();   // exchange buffer is reused
    fft256_rows_b<true, true>(v, sm, tw + 256, hi, lo);
    // thread (hi = r, lo = q'): v[pp] = b[k1 = 16g + r][n2 = 16*pp + q']; conj four-step twiddle
    mul_geometric(v, tb, ts);
#pragma unroll
    for(int pp = 0; pp < 16; pp++) row[16 * pp] = v[pp];
  }
  else
  {
    const float2 *col = sc + hi * 256 + 16 * g + lo;
#pragma unroll
    for(int j = 0; j < 16; j++) v[j] = __ldcg(col + j * 4096);
    fill_tw256_from(tw, gtw, tid, true);
    __syncthreads();
    fft256_cols<true, true>(v, sm, tw, hi, lo);
    // thread (hi = p1, lo): v[p2] = x2[256*(p1 + 16*p2) + 16g + lo]
    float2 *y = p.y + (long long) chan * p.y_stride;
    const int j0 = hi * 256 + 16 * g + lo;
    if(!p.ola_form)
    {
      const int i0 = j0 - p.out_shift;
      float2 *yb = y + (long long) blk * p.Ne + i0;
#pragma unroll
      for(int p2 = 0; p2 < 16; p2++)
        if((unsigned) (i0 + p2 * 4096) < (unsigned) p.Ne) stg_stream(yb + p2 * 4096, v[p2]);
    }
    else
    {
      const bool last = (blk + 1 == p.nblocks);
      float2 *svg = p.svg + (long long) chan * p.Ne;
#pragma unroll
      for(int p2 = 0; p2 < 16; p2++)
      {
        const int j = j0 + p2 * 4096;
        if(j < p.Nz)
          red_add_f2(y + (long long) blk * p.Ne + (p.Ne - p.Nz) + j, v[p2]);   // tail of block blk
        else if(last)
          svg[j - p.Nz] = v[p2];                                                 // carried (fourier.cc:872)
        else if(j < p.Ne)
          stg_stream(y + (long long) (blk + 1) * p.Ne + (j - p.Nz), v[p2]);
        else
          red_add_f2(y + (long long) (blk + 1) * p.Ne + (j - p.Nz), v[p2]);     // meets head of block blk+1
      }
    }
  }
}

template<int STAGE> __global__ void __launch_bounds__(OLA_NT, STAGE == 1 ? 3 : 4) ola64k_stage(OlaStageParams sp)
{
  __shared__ float2 sm[4096];
  __shared__ float4 tw[STAGE == 1 ? 512 : 256];
  const OlaRole &r = sp.role[0];
  ola_stage_body<STAGE>(sp.o, sp.tw, r.scratch + (long long) blockIdx.y * 65536, r.q0 + blockIdx.y, blockIdx.x, sm, tw);
}

// OLA form: y[0..Ne) of every channel starts as the carried svg; the other two-addend regions
// [b*Ne + Ne - Nz, (b+1)*Ne), b >= 1, start at zero.
__global__ void ola_prepare_kernel(float2 *y, long long y_stride, const float2 *svg, int Ne, int Nz, int nblocks)
{
  const int chan = blockIdx.y;
  float2 *yc = y + (long long) chan * y_stride;
  const float2 *sv = svg + (long long) chan * Ne;
  const long long total = (long long) Ne + (long long) (nblocks - 1) * Nz;
  for(long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x)
  {
    if(i < Ne) yc[i] = sv[i];
    else
    {
      const long long r = i - Ne;
      const int b = (int) (r / Nz) + 1, j = (int) (r % Nz);
      yc[(long long) b * Ne + (Ne - Nz) + j] = make_float2(0.f, 0.f);
    }
  }
}

// new_carry = last L samples of (old_carry ++ x[0..n))
__global__ void carry_update_kernel(const float2 *x, long long x_stride, int n, const float2 *o, float2 *d, int L)
{
  const int chan = blockIdx.y;
  for(int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x)
  {
    const long long pos = (long long) n - L + j;
    d[(long long) chan * L + j] = (pos >= 0) ? x[(long long) chan * x_stride + pos] : o[(long long) chan * L + L + pos];
  }
}

// ---- unfused path (N != 65536) ----------------------------------------------------------------
// padded[q][n] = n < Nz ? 0 : stream[(b0+blk)*Ne - residual + n - Nz]     (fourier.cc:850)
__global__ void ola_gather_kernel(const float2 *x, long long x_stride, const float2 *carry, int carry_len, float2 *work,
                                  int N, int Ne, int Nz, int nb, int b0, int residual)
{
  const int q = blockIdx.y, chan = q / nb, blk = q - chan * nb;
  const float2 *xc = x + (long long) chan * x_stride;
  const float2 *cr = carry + (long long) chan * carry_len + carry_len;
  for(int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x)
  {
    float2 v = make_float2(0.f, 0.f);
    if(n >= Nz)
    {
      const long long pos = (long long) (b0 + blk) * Ne - residual + n - Nz;
      v = (pos >= 0) ? xc[pos] : cr[pos];
    }
    work[(long long) q * N + n] = v;
  }
}
// ---- fused frame kernel of the batch pipeline (16 <= N <= 16384): gather (+ window) -> forward transform -> x H -> inverse
// transform, the frame resident in shared memory in between (one read of the Ne samples, one write of the N-point result,
// one launch instead of four).  Same Stockham passes and unitary scalings as fft_smem_kernel (fft.cu), parametrised by where
// a pass loads from / stores to.  SYNC_FIRST: the first pass reads shared memory as well (inverse leg) -> barrier before
// anything is stored; STORE_SM: the last pass stores into shared memory (forward leg) -> barrier in the radix-2 tail too.
template<bool INV, bool SYNC_FIRST, bool STORE_SM, class LoadF, class StoreF>
__device__ __forceinline__ void ola_smem_fft(LoadF load, StoreF store, float2 *sm, int N, int j, bool active)
{
  const int T = N >> 4;
  float2 v[16];
  int Ns = 1, rem = N;
  bool first = true;
  while(rem >= 16)
  {
    rem >>= 4;
    const bool last = rem == 1;
    if(active)
    {
#pragma unroll
      for(int t = 0; t < 16; t++) v[t] = first ? load(j + t * T) : sm[j + t * T];
    }
    if(!first || SYNC_FIRST) __syncthreads();
    const int k = j & (Ns - 1);
    if(Ns > 1) mul_geometric(v, make_float2(1.f, 0.f), twiddle<INV>((unsigned) k, 2.0f / (float) (Ns * 16)));
    fft16<INV>(v);
    const int j0 = (j - k) * 16 + k;
    if(active)
    {
#pragma unroll
      for(int t = 0; t < 16; t++)
      {
        if(last) store(j0 + t * Ns, v[t]);
        else sm[j0 + t * Ns] = v[t];
      }
    }
    if(!last) __syncthreads();
    Ns <<= 4;
    first = false;
  }
  if(rem >= 4)
  {
    rem >>= 2;
    const bool last = rem == 1;
    const int Q = N >> 2;
    if(active)
    {
#pragma unroll
      for(int m = 0; m < 4; m++)
#pragma unroll
        for(int t = 0; t < 4; t++) v[4 * m + t] = first ? load(j + m * T + t * Q) : sm[j + m * T + t * Q];
    }
    if(!first || SYNC_FIRST) __syncthreads();
#pragma unroll
    for(int m = 0; m < 4; m++)
    {
      const int jj = j + m * T, k = jj & (Ns - 1);
      if(Ns > 1)
      {
        const float2 w = twiddle<INV>((unsigned) k, 2.0f / (float) (Ns * 4)), w2 = cmul(w, w);
        v[4 * m + 1] = cmul(v[4 * m + 1], w);
        v[4 * m + 2] = cmul(v[4 * m + 2], w2);
        v[4 * m + 3] = cmul(v[4 * m + 3], cmul(w2, w));
      }
      fft4<INV>(v[4 * m], v[4 * m + 1], v[4 * m + 2], v[4 * m + 3]);
      const int j0 = (jj - k) * 4 + k;
      if(active)
      {
#pragma unroll
        for(int t = 0; t < 4; t++)
        {
          if(last) store(j0 + t * Ns, v[4 * m + t]);
          else sm[j0 + t * Ns] = v[4 * m + t];
        }
      }
    }
    if(!last) __syncthreads();
    Ns <<= 2;
    first = false;
  }
  if(rem == 2)
  {
    const int Hh = N >> 1;
    if(active)
    {
#pragma unroll
      for(int m = 0; m < 8; m++)
      {
        v[2 * m] = first ? load(j + m * T) : sm[j + m * T];
        v[2 * m + 1] = first ? load(j + m * T + Hh) : sm[j + m * T + Hh];
      }
    }
    if(STORE_SM || (first && SYNC_FIRST)) __syncthreads();
#pragma unroll
    for(int m = 0; m < 8; m++)
    {
      const int jj = j + m * T, k = jj & (Ns - 1);
      const float2 wb = Ns > 1 ? cmul(v[2 * m + 1], twiddle<INV>((unsigned) k, 2.0f / (float) (Ns * 2))) : v[2 * m + 1];
      const float2 a = cadd(v[2 * m], wb), d = csub(v[2 * m], wb);
      const int j0 = (jj - k) * 2 + k;
      if(active)
      {
        store(j0, a);
        store(j0 + Ns, d);
      }
    }
  }
}

// frame q of the chunk: plain mode q = chan * nb + blk; windowed mode q = (chan * nb + blk) * 2 + ph (the two frames of a block)
template<bool FEN>
__global__ void __launch_bounds__(1024, 1) ola_sandwich_kernel(const float2 *x, long long x_stride, const float2 *carry, int carry_len, float2 *work,
                                                               const float2 *H, const float *fen, int N, int Ne, int Nz, int nb, int b0, int residual,
                                                               int batch, float scale)
{
  extern __shared__ float2 ola_sm[];
  const int T = N >> 4, local = threadIdx.x / T, j = threadIdx.x - local * T;
  const long long q = (long long) blockIdx.x * (blockDim.x / T) + local;
  const bool active = q < batch;
  float2 *sm = ola_sm + (size_t) local * N;
  const int per = FEN ? 2 * nb : nb;
  const int chan = active ? (int) (q / per) : 0, r = (int) (q - (long long) chan * per), blk = FEN ? r >> 1 : r, ph = FEN ? r & 1 : 1;
  const float2 *xc = x + (long long) chan * x_stride;
  const float2 *cr = carry + (long long) chan * carry_len + carry_len;
  const long long first_pos = (long long) (b0 + blk) * Ne - residual - ((FEN && ph == 0) ? Ne / 2 : 0);
  auto gather = [&](int n) -> float2 {
    float2 v = make_float2(0.f, 0.f);
    if(n >= Nz)
    {
      const long long pos = first_pos + n - Nz;
      v = (pos >= 0) ? xc[pos] : cr[pos];
      if(FEN)
      {
        const float w = fen[n - Nz];
        v.x *= w;
        v.y *= w;
      }
    }
    return v;
  };
  auto to_sm = [&](int idx, float2 v) { sm[idx] = make_float2(v.x * scale, v.y * scale); };
  ola_smem_fft<false, false, true>(gather, to_sm, sm, N, j, active);
  __syncthreads();
  auto from_sm = [&](int idx) -> float2 { const float2 v = sm[idx]; return H ? cmul(v, __ldg(H + idx)) : v; };
  float2 *dst = work + q * N;
  auto to_work = [&](int idx, float2 v) { dst[idx] = make_float2(v.x * scale, v.y * scale); };
  ola_smem_fft<true, true, false>(from_sm, to_work, sm, N, j, active);
}

__global__ void ola_mulH_kernel(float2 *work, const float2 *H, int N, long long total)
{
  for(long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x)
    work[i] = cmul(work[i], H[i % N]);
}
// y_b[i] = prev[Nz + i] + (i >= Ne-Nz ? x2_b[i-(Ne-Nz)] : 0), prev = x2_{b-1} or svg   (fourier.cc:870-872)
__global__ void ola_scatter_kernel(const float2 *work, float2 *svg, float2 *y, long long y_stride, int N, int Ne, int Nz,
                                   int nb, int b0)
{
  const int q = blockIdx.y, chan = q / nb, blk = q - chan * nb;
  const float2 *cur = work + (long long) q * N;
  float2 *yb = y + (long long) chan * y_stride + (long long) (b0 + blk) * Ne;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < Ne; i += gridDim.x * blockDim.x)
  {
    float2 a = (blk == 0) ? svg[(long long) chan * Ne + i] : cur[i + Nz - N];   // previous block, offset Nz
    if(i >= Ne - Nz) a = cadd(a, cur[i - (Ne - Nz)]);
    yb[i] = a;
  }
}
__global__ void ola_svg_kernel(const float2 *work, float2 *svg, int N, int Ne, int Nz, int nb)
{
  const int chan = blockIdx.y;
  const float2 *lastb = work + ((long long) chan * nb + (nb - 1)) * N;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < Ne; i += gridDim.x * blockDim.x)
    svg[(long long) chan * Ne + i] = lastb[Nz + i];
}

// ---- Hann-window, 50 % overlap mode (fourier.cc:884-930), even Ne ------------------------------
// Two frames per block g: phase 0 = stream window starting Ne/2 before the block, phase 1 = the block itself, both
// times the window, zero-padded in front.  Z1_g, Z2_g = the filtered frames.  The reference's svg becomes a VIEW of
// its x2 buffer at the end of the first block (fourier.cc:923 moves a temporary view into svg, tableau.hpp:545-566),
// so from block 1 on every frame is folded onto ITSELF:
//   F(Z)[i] = Z[Nz+i] + (i >= Ne-Nz ? Z[i-(Ne-Nz)] : 0)
//   S1_g = F(Z1_g), S2_g = F(Z2_g)                                            (g >= 1)
//   S1_0[i] = i >= Ne-Nz ? Z1_0[i-(Ne-Nz)] : 0 ;  S2_0[i] = Z1_0[Nz+i] + (i >= Ne-Nz ? Z2_0[i-(Ne-Nz)] : 0)
//   last_g  = [ S1_g.tail(h)/2 + S2_g.head(h)/2 , S2_g.tail(h)/2 ]            (h = Ne/2)
//   y_g     = [ last_{g-1}.head(h) , last_{g-1}.tail(h) + S1_g.head(h)/2 ]    (g >= 1; block 0 emits nothing)
// which makes all blocks of a call independent given `last` of the previous call.
__global__ void ola_gather_fen_kernel(const float2 *x, long long x_stride, const float2 *carry, int carry_len, float2 *work,
                                      const float *fen, int N, int Ne, int Nz, int nb, int b0, int residual)
{
  const int q = blockIdx.y, chan = q / (2 * nb), r = q - chan * 2 * nb, blk = r >> 1, ph = r & 1;
  const float2 *xc = x + (long long) chan * x_stride;
  const float2 *cr = carry + (long long) chan * carry_len + carry_len;
  const long long first = (long long) (b0 + blk) * Ne - residual - (ph == 0 ? Ne / 2 : 0);
  for(int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x)
  {
    float2 v = make_float2(0.f, 0.f);
    if(n >= Nz)
    {
      const long long pos = first + n - Nz;
      v = (pos >= 0) ? xc[pos] : cr[pos];
      const float w = fen[n - Nz];
      v.x *= w;
      v.y *= w;
    }
    work[(long long) q * N + n] = v;
  }
}
__device__ __forceinline__ float2 fen_half(float2 a) { return make_float2(a.x * 0.5f, a.y * 0.5f); }
__device__ __forceinline__ float2 fen_s1(const float2 *Z1, bool first, int i, int Ne, int Nz)
{
  float2 a = first ? make_float2(0.f, 0.f) : Z1[Nz + i];
  if(i >= Ne - Nz) a = cadd(a, Z1[i - (Ne - Nz)]);
  return a;
}
__device__ __forceinline__ float2 fen_s2(const float2 *Z1, const float2 *Z2, bool first, int i, int Ne, int Nz)
{
  float2 a = first ? Z1[Nz + i] : Z2[Nz + i];
  if(i >= Ne - Nz) a = cadd(a, Z2[i - (Ne - Nz)]);
  return a;
}
// y of the blocks g = g0 + blk >= 1 of this chunk; e0 = emitted-block index of g0 within the call's output
__global__ void ola_scatter_fen_kernel(const float2 *work, const float2 *last, float2 *y, long long y_stride, int N, int Ne,
                                       int Nz, int nb, long long g0, long long e0)
{
  const int q = blockIdx.y, chan = q / nb, blk = q - chan * nb, h = Ne / 2;
  const long long g = g0 + blk;
  if(g < 1) return;
  const float2 *Z1 = work + ((long long) q * 2) * N;                 // this block, first frame
  const float2 *P1 = Z1 - 2 * (long long) N, *P2 = P1 + N;           // previous block (blk >= 1)
  float2 *yb = y + (long long) chan * y_stride + (e0 + blk) * Ne;
  const float2 *lc = last + (long long) chan * Ne;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < Ne; i += gridDim.x * blockDim.x)
  {
    float2 a;
    if(blk == 0) a = lc[i];
    else
    {
      a = fen_half(fen_s2(P1, P2, g - 1 == 0, i, Ne, Nz));
      if(i < h) a = cadd(fen_half(fen_s1(P1, g - 1 == 0, h + i, Ne, Nz)), a);
    }
    if(i >= h) a = cadd(a, fen_half(fen_s1(Z1, false, i - h, Ne, Nz)));
    yb[i] = a;
  }
}
// periodogramme_tfd: out[chan][frame][k] = 10 log10(|X[k]|^2 + 1e-20), k < N/2 (fourier.cc:1467-1468)
__global__ void periodo_logmag_kernel(const float2 *work, float *out, long long out_stride, int N, int frames_chunk, int frame0,
                                      int frames_per_chan)
{
  const int q = blockIdx.y, chan = q / frames_chunk, fr = q - chan * frames_chunk, nb = N / 2;
  const float2 *X = work + (long long) q * N;
  float *o = out + (long long) chan * out_stride + (long long) (frame0 + fr) * nb;
  (void) frames_per_chan;
  for(int k = blockIdx.x * blockDim.x + threadIdx.x; k < nb; k += gridDim.x * blockDim.x)
  {
    const float2 v = X[k];
    o[k] = 10.0f * log10f(v.x * v.x + v.y * v.y + 1e-20f);
  }
}
// last = last_g of the chunk's final block
__global__ void ola_last_fen_kernel(const float2 *work, float2 *last, int N, int Ne, int Nz, int nb, long long g_last)
{
  const int chan = blockIdx.y, h = Ne / 2;
  const float2 *Z1 = work + (((long long) chan * nb + (nb - 1)) * 2) * N, *Z2 = Z1 + N;
  const bool first = g_last == 0;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < Ne; i += gridDim.x * blockDim.x)
  {
    float2 a = fen_half(fen_s2(Z1, Z2, first, i, Ne, Nz));
    if(i < h) a = cadd(fen_half(fen_s1(Z1, first, h + i, Ne, Nz)), a);
    last[(long long) chan * Ne + i] = a;
  }
}


// ---- normalised-correlation detector (reference src/fourier/detection.cc:204-260) ---------------------------------------
// score[i] = ratio * sqrt(|corr[i]|^2 / (en[i] + 1e-20)), en[i] = mean of |stream|^2 over the M samples the correlation
// at output i covers, i.e. stream positions [i - Ne, i - Ne + M) (filtre_mg<float,double>(M) followed by the delay line
// of Ne - M + 1 samples, detection.cc:132,165,213-217).  The moving sum is a difference of a double-precision running sum,
// like the reference's double accumulator (filtre-rt.cc:633-667).
// Pass 1: P[c][j] = sum of |s|^2 over stream positions [-Ne - 1, -Ne - 1 + j), j in [0, Ne + n + 1], one CTA per channel.
__global__ void __launch_bounds__(1024) detect_energy_scan_kernel(const float2 *x, long long x_stride, const float2 *carry, int carry_len,
                                                                  int Ne, int n, double *P)
{
  __shared__ double part[1024];
  const int chan = blockIdx.x, tid = threadIdx.x;
  const long long L = (long long) Ne + 1 + n;                 // elements, positions -Ne-1 .. n-1
  const long long per = (L + 1023) / 1024;
  const float2 *xc = x + (long long) chan * x_stride;
  const float2 *cr = carry + (long long) chan * carry_len + carry_len;
  double *Pc = P + (long long) chan * (L + 1);
  const long long j0 = tid * per, j1 = min(L, j0 + per);
  double s = 0;
  for(long long j = j0; j < j1; j++)
  {
    const long long pos = j - Ne - 1;
    const float2 v = pos >= 0 ? xc[pos] : cr[pos];
    s += (double) (v.x * v.x + v.y * v.y);                   // abs2 in float like the reference, accumulated in double
  }
  part[tid] = s;
  __syncthreads();
  for(int d = 1; d < 1024; d <<= 1)
  {
    const double t = tid >= d ? part[tid - d] : 0.0;
    __syncthreads();
    part[tid] += t;
    __syncthreads();
  }
  double run = tid ? part[tid - 1] : 0.0;
  if(tid == 0) Pc[0] = 0.0;
  for(long long j = j0; j < j1; j++)
  {
    const long long pos = j - Ne - 1;
    const float2 v = pos >= 0 ? xc[pos] : cr[pos];
    run += (double) (v.x * v.x + v.y * v.y);
    Pc[j + 1] = run;
  }
}
// Pass 2: the score, and the reference's clean-up of tiny correlations (detection.cc:240-244) applied to corr in place
__global__ void detect_score_kernel(float2 *corr, long long corr_stride, const double *P, float *score, long long score_stride,
                                    int Ne, int M, int n, float ratio, float K_inv)
{
  const int chan = blockIdx.y;
  const long long L1 = (long long) Ne + 2 + n;
  const double *Pc = P + (long long) chan * L1;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    // positions [i - Ne, i - Ne + M): prefix indices (pos + Ne + 1) .. (+ M)
    const double acc = Pc[i + 1 + M] - Pc[i + 1];
    const float en = ((float) acc) * K_inv;
    float2 c = corr[(long long) chan * corr_stride + i];
    if(hypotf(c.x, c.y) <= 1e-6f)                       // abs(corr) <= sqrt(1e-12f)
    {
      c = make_float2(0.f, 0.f);
      corr[(long long) chan * corr_stride + i] = c;
    }
    const float a2 = c.x * c.x + c.y * c.y;
    score[(long long) chan * score_stride + i] = ratio * sqrtf(a2 / (en + 1e-20f));
  }
}

} // namespace tsdgpu

using namespace tsdgpu;

struct tsdgpu_ola_s
{
  int device = 0;              // CUDA device the object lives on
  int Ne = 0, N = 0, Nz = 0, K = 0, nchan = 0;
  int residual = 0;            // TamponNv2 windex (tsd.cc:310)
  long long blocks_done = 0;
  bool fused = false;
  // FIR-derived gains (fir_len > 0): single-SM overlap-save kernel with its own transform size (ols16k.cu)
  Ols16k *ols = nullptr;
  int delay = 0;               // output t = FIR output t - delay: Ne - K (FiltreFFTRIF convention), Ne - M + 1 for the detector's correlator
  float2 *d_H = nullptr;       // fused: H/N ; unfused: raw H (nullptr = identity)
  float2 *d_carry[2] = {nullptr, nullptr};
  int cur = 0, carry_len = 0;
  float2 *d_svg = nullptr;
  // fused
  float2 *scratch = nullptr;
  unsigned *flags = nullptr;
  size_t flags_cap = 0;
  int ring = 80, lag = 24, ctas = 0;
  // staged form (default): blocks per stage kernel, number of auxiliary streams
  // 1 = staged kernels over auxiliary streams (default), 0 = single persistent kernel
  int staged = 1, chunk = 32, nslots = 8;   // 8 streams: 112 vs 108 Gsamples/s with 4 (profiles/sweep_ola_streams.sh)
  // windowed mode (always unfused): window [Ne], `last` of the reference [nchan][Ne]
  bool fen = false;
  float *d_fen = nullptr;
  float2 *d_last = nullptr;
  // generic spectral callback (FiltreFFTConfig::traitement_freq, fourier.hpp:319): host function called once per
  // transformed block, in stream order; the block spectra make a round trip through pinned host memory
  tsdgpu_spectral_cb cb = nullptr;
  void *cb_user = nullptr;
  float2 *h_spec = nullptr;    // pinned, [batch][N]
  size_t h_spec_cap = 0;
  // unfused
  tsdgpu_fft_s *plan = nullptr;
  float2 *work = nullptr;
  int plan_batch = 0;
};

static int ola_run_fused(tsdgpu_ola_s *f, const float2 *x, long long xs, int n, float2 *y, long long ys, int B)
{
  Runtime &r = rt();
  const long long Q = (long long) f->nchan * B;
  if(Q > (1LL << 25)) return fail("tsdgpu_ola_step: too many blocks in one call (split the call)");
  const size_t need = (size_t) 3 * Q + 1;
  if(!f->staged && need > f->flags_cap)
  {
    if(f->flags) cudaFree(f->flags);
    f->flags = nullptr;
    f->flags_cap = 0;
    TSD_CUDA(cudaMalloc(&f->flags, need * sizeof(unsigned)));
    f->flags_cap = need;
  }
  if(!f->staged) TSD_CUDA(cudaMemsetAsync(f->flags, 0, need * sizeof(unsigned), r.stream));
  OlaParams p;
  p.x = x;
  p.y = y;
  p.carry = f->d_carry[f->cur];
  p.svg = f->d_svg;
  p.H = f->d_H;
  p.scratch = f->scratch;
  p.done_a = f->flags;
  p.done_b = f->flags + Q;
  p.done_c = f->flags + 2 * Q;
  p.ticket = f->flags + 3 * Q;
  p.x_stride = xs;
  p.y_stride = ys;
  p.carry_len = f->carry_len;
  p.nblocks = B;
  p.Q = (int) Q;
  p.Ne = f->Ne;
  p.Nz = f->Nz;
  p.residual = f->residual;
  p.ring = f->ring;
  p.lag = f->lag;
  p.n = n;
  p.tw4 = r.tw4;
  if(f->K > 0)
  {
    p.ola_form = 0;
    p.base_off = f->K - f->N;     // window = N inputs ending K-1 samples after the block start
    p.zero_below = 0;
    p.out_shift = f->Nz - f->K;
  }
  else
  {
    p.ola_form = 1;
    p.base_off = -f->Nz;          // padded.tail(Ne) = x (fourier.cc:850)
    p.zero_below = f->Nz;
    p.out_shift = 0;
    dim3 grid(std::min(1024LL, ((long long) f->Ne + (long long) (B - 1) * f->Nz + 255) / 256), f->nchan);
    ola_prepare_kernel<<<grid, 256, 0, r.stream>>>(y, ys, f->d_svg, f->Ne, f->Nz, B);
    TSD_LAUNCH_CHECK();
  }
  if(f->staged)
  {
    KernelTimer timer;
    OlaStageParams sp;
    sp.o = p;
    sp.tw = r.tw256;
    const int C = f->chunk;
    // TSDGPU_OLA_SKEW=d gives stream s chunks of C + d*(s - (S-1)/2) blocks (<= C) so that the streams drift out of
    // lock-step: +1..2 % (measured; marking the scratch as an L2 persisting window instead costs 15 %)
    static const int skew = getenv("TSDGPU_OLA_SKEW") ? atoi(getenv("TSDGPU_OLA_SKEW")) : 0;
    const int S = f->nslots;
    if(aux_fork(S)) return 1;
    long long q0 = 0;
    for(int c = 0; q0 < Q; c++)
    {
      const int s = c % S;
      int cs = C + (skew * (2 * s - (S - 1))) / 2;
      cs = std::max(1, std::min(cs, C));
      OlaRole &ro = sp.role[0];
      ro.q0 = (int) q0;
      ro.nb = (int) std::min<long long>(cs, Q - q0);
      ro.scratch = f->scratch + (size_t) s * C * 65536;
      q0 += ro.nb;
      const dim3 grid(16, ro.nb);
      ola64k_stage<0><<<grid, OLA_NT, 0, r.aux[s]>>>(sp);
      TSD_LAUNCH_CHECK();
      ola64k_stage<1><<<grid, OLA_NT, 0, r.aux[s]>>>(sp);
      TSD_LAUNCH_CHECK();
      ola64k_stage<2><<<grid, OLA_NT, 0, r.aux[s]>>>(sp);
      TSD_LAUNCH_CHECK();
    }
    if(aux_join(f->nslots)) return 1;
    return 0;
  }
  const long long tickets = (Q + 2LL * f->lag) * 48;
  const int grid = (int) std::min<long long>(f->ctas, tickets);
  {
    KernelTimer timer;
    ola64k_kernel<<<grid, OLA_NT, 0, r.stream>>>(p);
    TSD_LAUNCH_CHECK();
  }
  return 0;
}

// The unfused / windowed paths keep ONE plan and work buffer sized for the largest batch seen; smaller chunks (the
// ragged last chunk, or calls that alternate between B and B+1 blocks) run the same plan with the live batch, so no
// allocation or device synchronisation happens on the data path after the first call of that size.
static int ola_reserve_plan(tsdgpu_ola_s *f, int batch)
{
  if(f->plan && f->plan_batch >= batch)
  {
    f->plan->batch = batch;
    return 0;
  }
  if(f->plan)
  {
    TSD_CUDA(cudaStreamSynchronize(rt().stream));
    fft_plan_destroy(f->plan);
  }
  f->plan = nullptr;
  if(f->work) cudaFree(f->work);
  f->work = nullptr;
  if(fft_plan_create(f->N, batch, &f->plan)) return 1;
  f->plan_batch = batch;
  TSD_CUDA(cudaMalloc(&f->work, (size_t) batch * f->N * sizeof(float2)));
  return 0;
}

// The spectral step between the two transforms: multiply by H on the device, or hand every spectrum to the caller's
// host callback (FFT -> D2H -> callback -> H2D -> IFFT, SURVEY 7 "generic callbacks").  Spectra are unitary-scaled like
// the reference's X (fourier.cc:855-865); batch entries are in the reference's call order for every channel.
static int ola_spectral_step(tsdgpu_ola_s *f, int batch, int per_chan)
{
  Runtime &r = rt();
  const int N = f->N;
  if(f->cb)
  {
    const size_t need = (size_t) batch * N;
    if(need > f->h_spec_cap)
    {
      if(f->h_spec) cudaFreeHost(f->h_spec);
      f->h_spec = nullptr;
      f->h_spec_cap = 0;
      TSD_CUDA(cudaMallocHost(&f->h_spec, need * sizeof(float2)));
      f->h_spec_cap = need;
    }
    TSD_CUDA(cudaMemcpyAsync(f->h_spec, f->work, need * sizeof(float2), cudaMemcpyDeviceToHost, r.stream));
    TSD_CUDA(cudaStreamSynchronize(r.stream));
    for(int q = 0; q < batch; q++) f->cb(f->cb_user, q / per_chan, reinterpret_cast<float *>(f->h_spec + (size_t) q * N), N);
    TSD_CUDA(cudaMemcpyAsync(f->work, f->h_spec, need * sizeof(float2), cudaMemcpyHostToDevice, r.stream));
    return 0;
  }
  if(f->d_H)
  {
    const long long total = (long long) batch * N;
    ola_mulH_kernel<<<(int) std::min<long long>((total + 255) / 256, r.num_sms * 16), 256, 0, r.stream>>>(f->work, f->d_H, N, total);
    TSD_LAUNCH_CHECK();
  }
  return 0;
}

// bytes of frames in flight per chunk of the batch pipeline (TSDGPU_OLA_WORK_MB, default 256).  Measured: smaller chunks that
// would stay in the 126 MB L2 are SLOWER (windowed mode 25.7 / 23.1 / 18.1 Gsamples/s at 256 / 64 / 16 MiB): the per-chunk
// launch sequence, not the HBM round trip of the frames, bounds this path.
static long long ola_work_budget()
{
  static const long long mb = [] { const char *e = getenv("TSDGPU_OLA_WORK_MB"); const int v = e ? atoi(e) : 256; return (long long) (v >= 1 && v <= 4096 ? v : 256); }();
  return mb << 20;
}
// gather -> FFT -> x H -> IFFT of a chunk in ONE launch when the frame fits shared memory (power of two, 16 ... 16384) and the
// spectral step is a multiplication (no host callback); TSDGPU_OLA_SANDWICH=0 keeps the four launches
static bool ola_sandwich_ok(const tsdgpu_ola_s *f)
{
  const int N = f->N;
  return !f->cb && N >= 16 && N <= 16384 && (N & (N - 1)) == 0 && !(getenv("TSDGPU_OLA_SANDWICH") && atoi(getenv("TSDGPU_OLA_SANDWICH")) == 0);
}
template<bool FEN>
static int ola_sandwich(tsdgpu_ola_s *f, const float2 *x, long long xs, int batch, int nb, int b0)
{
  Runtime &r = rt();
  const int N = f->N, T = N / 16, threads = std::max(256, T), per_cta = threads / T;
  const size_t smem = (size_t) per_cta * N * sizeof(float2);
  if(!r.ola_sandwich_ready)
  {
    TSD_CUDA(cudaFuncSetAttribute(ola_sandwich_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    TSD_CUDA(cudaFuncSetAttribute(ola_sandwich_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    r.ola_sandwich_ready = true;
  }
  const unsigned grid = (unsigned) ((batch + per_cta - 1) / per_cta);
  KernelTimer timer;
  ola_sandwich_kernel<FEN><<<grid, threads, smem, r.stream>>>(x, xs, f->d_carry[f->cur], f->carry_len, f->work, f->d_H, f->d_fen, N, f->Ne, f->Nz, nb,
                                                              b0, f->residual, batch, 1.0f / sqrtf((float) N));
  TSD_LAUNCH_CHECK();
  return 0;
}

static int ola_run_unfused(tsdgpu_ola_s *f, const float2 *x, long long xs, float2 *y, long long ys, int B)
{
  Runtime &r = rt();
  const int N = f->N, Ne = f->Ne, Nz = f->Nz;
  // blocks per chunk: bounded work buffer (<= 256 MiB)
  // blocks per chunk: bounded work buffer (<= 256 MiB) and grid (<= 65535 rows per launch)
  if(f->nchan > 65535) return fail("tsdgpu_ola_step: more than 65535 channels on the unfused path");
  long long per_block = (long long) f->nchan * N * (long long) sizeof(float2);
  const int nb_max = (int) std::max(1LL, std::min(ola_work_budget() / per_block, 65535LL / f->nchan));
  for(int b0 = 0; b0 < B; b0 += nb_max)
  {
    const int nb = std::min(nb_max, B - b0);
    const int batch = f->nchan * nb;
    if(ola_reserve_plan(f, batch)) return 1;
    if(ola_sandwich_ok(f))
    {
      if(ola_sandwich<false>(f, x, xs, batch, nb, b0)) return 1;
    }
    else
    {
      dim3 gg((N + 255) / 256, batch);
      ola_gather_kernel<<<gg, 256, 0, r.stream>>>(x, xs, f->d_carry[f->cur], f->carry_len, f->work, N, Ne, Nz, nb, b0,
                                                 f->residual);
      TSD_LAUNCH_CHECK();
      if(fft_exec_device(f->plan, f->work, N, f->work, N, true)) return 1;
      if(ola_spectral_step(f, batch, nb)) return 1;
      if(fft_exec_device(f->plan, f->work, N, f->work, N, false)) return 1;
    }
    dim3 gs((Ne + 255) / 256, batch);
    ola_scatter_kernel<<<gs, 256, 0, r.stream>>>(f->work, f->d_svg, y, ys, N, Ne, Nz, nb, b0);
    TSD_LAUNCH_CHECK();
    dim3 gv((Ne + 255) / 256, f->nchan);
    ola_svg_kernel<<<gv, 256, 0, r.stream>>>(f->work, f->d_svg, N, Ne, Nz, nb);
    TSD_LAUNCH_CHECK();
  }
  return 0;
}

static int ola_run_fen(tsdgpu_ola_s *f, const float2 *x, long long xs, float2 *y, long long ys, int B)
{
  Runtime &r = rt();
  const int N = f->N, Ne = f->Ne, Nz = f->Nz;
  if(2LL * f->nchan > 65535) return fail("tsdgpu_ola_step: more than 32767 channels in the windowed mode");
  const long long per_block = 2LL * f->nchan * N * (long long) sizeof(float2);
  const int nb_max = (int) std::max(1LL, std::min(ola_work_budget() / per_block, 65535LL / (2LL * f->nchan)));
  const long long g0 = f->blocks_done, e_first = std::max(g0, 1LL);
  for(int b0 = 0; b0 < B; b0 += nb_max)
  {
    const int nb = std::min(nb_max, B - b0);
    const int batch = f->nchan * nb * 2;
    if(ola_reserve_plan(f, batch)) return 1;
    if(ola_sandwich_ok(f))
    {
      if(ola_sandwich<true>(f, x, xs, batch, nb, b0)) return 1;
    }
    else
    {
      dim3 gg((N + 255) / 256, batch);
      ola_gather_fen_kernel<<<gg, 256, 0, r.stream>>>(x, xs, f->d_carry[f->cur], f->carry_len, f->work, f->d_fen, N, Ne, Nz, nb,
                                                     b0, f->residual);
      TSD_LAUNCH_CHECK();
      if(fft_exec_device(f->plan, f->work, N, f->work, N, true)) return 1;
      if(ola_spectral_step(f, batch, 2 * nb)) return 1;
      if(fft_exec_device(f->plan, f->work, N, f->work, N, false)) return 1;
    }
    dim3 gs((Ne + 255) / 256, f->nchan * nb);
    ola_scatter_fen_kernel<<<gs, 256, 0, r.stream>>>(f->work, f->d_last, y, ys, N, Ne, Nz, nb, g0 + b0, g0 + b0 - e_first);
    TSD_LAUNCH_CHECK();
    dim3 gv((Ne + 255) / 256, f->nchan);
    ola_last_fen_kernel<<<gv, 256, 0, r.stream>>>(f->work, f->d_last, N, Ne, Nz, nb, g0 + b0 + nb - 1);
    TSD_LAUNCH_CHECK();
  }
  return 0;
}

static int ola_run_device(tsdgpu_ola_s *f, const float2 *x, long long xs, int n, float2 *y, long long ys, long long *n_out)
{
  Runtime &r = rt();
  const long long tot = (long long) f->residual + n;
  const int B = (int) (tot / f->Ne);
  // windowed mode: the first block of the stream emits nothing (cnt_ech < 0, fourier.cc:900-903)
  *n_out = (long long) (B - ((f->fen && f->blocks_done == 0 && B > 0) ? 1 : 0)) * f->Ne;
  if(n <= 0) return 0;
  if(B > 0)
  {
    if(ys < *n_out) return fail("tsdgpu_ola_step: output stride smaller than the emitted count");
    int rc = f->fen   ? ola_run_fen(f, x, xs, y, ys, B)
             : f->ols ? ols16k_run(f->ols, x, xs, n, f->d_carry[f->cur], f->carry_len, y, ys, *n_out, f->delay, f->residual, f->nchan)
             : f->fused ? ola_run_fused(f, x, xs, n, y, ys, B)
                        : ola_run_unfused(f, x, xs, y, ys, B);
    if(rc) return rc;
  }
  dim3 grid((f->carry_len + 255) / 256, f->nchan);
  carry_update_kernel<<<grid, 256, 0, r.stream>>>(x, xs, n, f->d_carry[f->cur], f->d_carry[f->cur ^ 1], f->carry_len);
  TSD_LAUNCH_CHECK();
  f->cur ^= 1;
  f->residual = (int) (tot % f->Ne);
  f->blocks_done += B;
  return 0;
}

extern "C" {

static int ola_create(int dim_blocs_temporel, int nb_zeros_min, const float *H, int fir_len, const float *fenetre, int nchan,
                      tsdgpu_ola_t *out, tsdgpu_spectral_cb cb = nullptr, void *cb_user = nullptr);

int tsdgpu_ola_create_cb(int dim_blocs_temporel, int nb_zeros_min, tsdgpu_spectral_cb cb, void *user, const float *fenetre,
                         int nchan, tsdgpu_ola_t *out)
{
  if(!cb) return fail("tsdgpu_ola_create_cb: null callback");
  return ola_create(dim_blocs_temporel, nb_zeros_min, nullptr, 0, fenetre, nchan, out, cb, user);
}

int tsdgpu_ola_create(int dim_blocs_temporel, int nb_zeros_min, const float *H, int fir_len, int nchan, tsdgpu_ola_t *out)
{
  return ola_create(dim_blocs_temporel, nb_zeros_min, H, fir_len, nullptr, nchan, out);
}

int tsdgpu_ola_create_fen(int dim_blocs_temporel, int nb_zeros_min, const float *H, const float *fenetre, int nchan,
                          tsdgpu_ola_t *out)
{
  if(!fenetre) return fail("tsdgpu_ola_create_fen: null window");
  return ola_create(dim_blocs_temporel, nb_zeros_min, H, 0, fenetre, nchan, out);
}

} // extern "C"

static int ola_create(int dim_blocs_temporel, int nb_zeros_min, const float *H, int fir_len, const float *fenetre, int nchan,
                      tsdgpu_ola_t *out, tsdgpu_spectral_cb cb, void *cb_user)
{
  TSD_ENTER(-1);
  if(!out) return fail("tsdgpu_ola_create: null argument");
  if(nchan <= 0) return fail("tsdgpu_ola_create: nchan must be > 0");
  if(nb_zeros_min < 0) return fail("tsdgpu_ola_create: nb_zeros_min < 0");
  int Ne = dim_blocs_temporel;
  if(Ne <= 0) Ne = 512;                                   // fourier.cc:769-770
  if((long long) Ne + nb_zeros_min > (1 << 24)) return fail("tsdgpu_ola_create: block too large");
  const int N = tsdgpu_p2(Ne + nb_zeros_min);             // fourier.cc:775
  const int Nz = N - Ne;
  if(Nz > Ne)
    return fail("tsdgpu_ola_create: N_zeros > Ne — the reference indexes before its svg buffer here "
                "(svg.tail(N_zeros), fourier.cc:870); choose a larger dim_blocs_temporel");
  // K taps at the tail of the N-vector convolve without wrap-around only if K <= N_zeros
  if(fir_len < 0 || fir_len > Nz)
    return fail("tsdgpu_ola_create: fir_len must be in [1, N_zeros] (or 0 for an arbitrary H)");
  if(fir_len > 0 && !H) return fail("tsdgpu_ola_create: fir_len > 0 needs H");
  if(fenetre && (Ne & 1))
    return fail("tsdgpu_ola_create_fen: odd dim_blocs_temporel is not supported in the windowed mode (the reference never "
                "resets the centre sample of its `last` buffer there, fourier.cc:899-906)");
  auto *f = new tsdgpu_ola_s;
  f->device = rt().device;
  f->Ne = Ne;
  f->N = N;
  f->Nz = Nz;
  f->nchan = nchan;
  f->fen = fenetre != nullptr;
  f->cb = cb;
  f->cb_user = cb_user;
  f->fused = (N == 65536) && !f->fen && !cb;   // a host callback needs the spectra on the host: N-point unfused path
  f->K = f->fused ? fir_len : 0;   // the unfused path always runs the overlap-add form
  f->carry_len = N + Ne;
  // gains that are the transform of K taps: the samples do not depend on the block structure, the single-SM
  // overlap-save kernel produces them with its own 16384-point transforms (TSDGPU_OLA_OLS=0 keeps the N-point paths)
  {
    const char *v = getenv("TSDGPU_OLA_OLS");
    if(fir_len > 0 && !f->fen && !(v && v[0] == '0'))
    {
      if(ols16k_create(H, N, fir_len, &f->ols))
      {
        delete f;
        return 1;
      }
      if(f->ols)
      {
        f->delay = Ne - fir_len;
        f->K = fir_len;
        f->fused = false;
        f->carry_len = std::max(N + Ne, 2 * Ne + f->ols->O + 2);   // look-back of a window: delay + overlap + residual
      }
    }
  }
  cudaError_t e = cudaSuccess;
  for(int i = 0; i < 2 && e == cudaSuccess; i++)
  {
    e = cudaMalloc(&f->d_carry[i], (size_t) nchan * f->carry_len * sizeof(float2));
    if(e == cudaSuccess) e = cudaMemsetAsync(f->d_carry[i], 0, (size_t) nchan * f->carry_len * sizeof(float2), rt().stream);
  }
  if(e == cudaSuccess) e = cudaMalloc(&f->d_svg, (size_t) nchan * Ne * sizeof(float2));
  if(e == cudaSuccess) e = cudaMemsetAsync(f->d_svg, 0, (size_t) nchan * Ne * sizeof(float2), rt().stream);   // svg.setZero(Ne), fourier.cc:785
  if(e == cudaSuccess && f->fen)
  {
    e = cudaMalloc(&f->d_fen, (size_t) Ne * sizeof(float));
    if(e == cudaSuccess) e = cudaMemcpy(f->d_fen, fenetre, (size_t) Ne * sizeof(float), cudaMemcpyHostToDevice);
    if(e == cudaSuccess) e = cudaMalloc(&f->d_last, (size_t) nchan * Ne * sizeof(float2));
    if(e == cudaSuccess) e = cudaMemsetAsync(f->d_last, 0, (size_t) nchan * Ne * sizeof(float2), rt().stream);   // fourier.cc:784
  }
  if(e == cudaSuccess && (H || f->fused))
  {
    std::vector<float2> h((size_t) N);
    const float sc = f->fused ? 1.0f / (float) N : 1.0f;   // both unitary scalings folded into H (exact: N = 2^16)
    for(int i = 0; i < N; i++)
      h[i] = H ? make_float2(H[2 * i] * sc, H[2 * i + 1] * sc) : make_float2(sc, 0.f);
    e = cudaMalloc(&f->d_H, (size_t) N * sizeof(float2));
    if(e == cudaSuccess) e = cudaMemcpy(f->d_H, h.data(), (size_t) N * sizeof(float2), cudaMemcpyHostToDevice);
  }
  if(e == cudaSuccess && f->fused)
  {
    // pipeline depth knobs (tuning experiments): lag = slots between dependent stages, ring = scratch slots
    if(const char *v = getenv("TSDGPU_OLA_LAG")) f->lag = std::max(1, atoi(v));
    if(const char *v = getenv("TSDGPU_OLA_RING")) f->ring = atoi(v);
    if(f->ring <= 2 * f->lag) f->ring = 2 * f->lag + 16;
    // TSDGPU_OLA_MODE=persistent selects the single persistent kernel; default = staged kernels
    if(const char *v = getenv("TSDGPU_OLA_MODE")) f->staged = v[0] == 'p' ? 0 : 1;
    // blocks per stage kernel: one full wave of the 4-CTA/SM stages (16 tiles per block): 37 on 148 SMs, +3.5 % over 32
    f->chunk = std::max(8, rt().num_sms * 4 / 16);
    if(const char *v = getenv("TSDGPU_OLA_CHUNK")) f->chunk = std::max(1, atoi(v));
    if(const char *v = getenv("TSDGPU_OLA_STREAMS")) f->nslots = std::min((int) Runtime::MAX_AUX, std::max(1, atoi(v)));
    if(aux_init()) e = cudaErrorUnknown;   // twiddle tables (and the auxiliary streams of the staged form)
    const size_t slots = f->staged ? (size_t) f->chunk * f->nslots : (size_t) f->ring;
    if(e == cudaSuccess) e = cudaMalloc(&f->scratch, slots * 65536 * sizeof(float2));
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ola64k_kernel, OLA_NT, 0);
    f->ctas = rt().num_sms * std::max(1, occ);
  }
  if(e == cudaSuccess) e = cudaStreamSynchronize(rt().stream);
  if(e != cudaSuccess)
  {
    tsdgpu_ola_destroy(f);
    return fail(std::string("tsdgpu_ola_create: ") + cudaGetErrorString(e));
  }
  *out = f;
  return 0;
}

extern "C" {

int tsdgpu_ola_dims(tsdgpu_ola_t f, int *Ne, int *N, int *Nz, int *residual)
{
  if(!f) return fail("tsdgpu_ola_dims: null handle");
  if(Ne) *Ne = f->Ne;
  if(N) *N = f->N;
  if(Nz) *Nz = f->Nz;
  if(residual) *residual = f->residual;
  return 0;
}

long long tsdgpu_ola_out_count(tsdgpu_ola_t f, int n)
{
  if(!f || n < 0) return 0;
  const long long B = ((long long) f->residual + n) / f->Ne;
  return (B - ((f->fen && f->blocks_done == 0 && B > 0) ? 1 : 0)) * f->Ne;
}

int tsdgpu_ola_step(tsdgpu_ola_t f, const void *x, long long xs, int n, void *y, long long ys, long long *n_out, int mem)
{
  TSD_ENTER(f ? f->device : -1);
  if(!f || !n_out) return fail("tsdgpu_ola_step: null argument");
  *n_out = 0;
  if(n < 0) return fail("tsdgpu_ola_step: n < 0");
  if(n == 0) return 0;
  if(!x) return fail("tsdgpu_ola_step: null input");
  if(xs < n) return fail("tsdgpu_ola_step: channel stride smaller than n");
  const long long cnt = tsdgpu_ola_out_count(f, n);
  if(cnt > 0 && !y) return fail("tsdgpu_ola_step: null output");
  if(mem == TSDGPU_DEVICE)
  {
    if(x == y) return fail("tsdgpu_ola_step: in-place operation is not supported (neither does the reference, fourier.cc:871)");
    return ola_run_device(f, (const float2 *) x, xs, n, (float2 *) y, ys, n_out);
  }
  (void) cnt;
  const long long chunk = host_chunk_len(f->nchan, 8, n, 2);
  const long long out_cap = (chunk / f->Ne + 2) * (long long) f->Ne;
  if(host_stage_reserve((size_t) f->nchan * chunk * 8, (size_t) f->nchan * out_cap * 8)) return 1;
  HostStage &hs = host_stage();
  const float2 *xh = (const float2 *) x;
  float2 *yh = (float2 *) y;
  return host_pipeline(
    n, chunk,
    [&](int slot, long long first, long long count) -> int {
      if(stage_in(slot, hs.in[slot], (size_t) chunk * 8, xh + first, (size_t) xs * 8, (size_t) count * 8, f->nchan)) return 1;
      return 0;
    },
    [&](long long count) { return tsdgpu_ola_out_count(f, (int) count); },
    [&](int slot, long long count, long long *got) -> int {
      return ola_run_device(f, (const float2 *) hs.in[slot], chunk, (int) count, (float2 *) hs.out[slot], out_cap, got);
    },
    [&](int slot, long long out_first, long long count) -> int {
      if(stage_out(slot, yh + out_first, (size_t) ys * 8, hs.out[slot], (size_t) out_cap * 8, (size_t) count * 8, f->nchan)) return 1;
      return 0;
    },
    n_out);
}


int tsdgpu_periodogramme_tfd(const void *x, long long x_stride, int n, int nchan, int N, const float *fenetre, float *out,
                             long long out_stride, int *n_frames, int *n_bins, int mem)
{
  TSD_ENTER(-1);
  if(!x || !fenetre || !out || !n_frames || !n_bins) return fail("tsdgpu_periodogramme_tfd: null argument");
  if(n < 0 || nchan <= 0 || N < 2 || N > (1 << 24)) return fail("tsdgpu_periodogramme_tfd: invalid size");
  if(N & 1) return fail("tsdgpu_periodogramme_tfd: odd N is not supported (see tsdgpu_ola_create_fen)");
  if(x_stride < n) return fail("tsdgpu_periodogramme_tfd: channel stride smaller than n");
  const int N2 = tsdgpu_p2(N), Nz = N2 - N, B = n / N, frames = 2 * B, nbins = N2 / 2;
  *n_frames = frames;
  *n_bins = nbins;
  if(frames == 0) return 0;
  if(out_stride < (long long) frames * nbins) return fail("tsdgpu_periodogramme_tfd: output stride too small");
  Runtime &r = rt();
  const float2 *dx = (const float2 *) x;
  float *dout = out;
  float2 *tmp_x = nullptr, *zeros = nullptr, *work = nullptr;
  float *tmp_out = nullptr, *d_fen = nullptr;
  tsdgpu_fft_s *plan = nullptr;
  long long dxs = x_stride, dos = out_stride;
  int rc = 0;
  auto cleanup = [&]() {
    cudaStreamSynchronize(r.stream);
    if(tmp_x) cudaFree(tmp_x);
    if(tmp_out) cudaFree(tmp_out);
    if(zeros) cudaFree(zeros);
    if(work) cudaFree(work);
    if(d_fen) cudaFree(d_fen);
    if(plan) fft_plan_destroy(plan);
  };
#define PG_CUDA(call)                                                                              \
  do                                                                                               \
  {                                                                                                \
    cudaError_t e_ = (call);                                                                       \
    if(e_ != cudaSuccess)                                                                          \
    {                                                                                              \
      cleanup();                                                                                   \
      return fail(std::string("tsdgpu_periodogramme_tfd: ") + cudaGetErrorString(e_));             \
    }                                                                                              \
  } while(0)
  if(mem != TSDGPU_DEVICE)
  {
    PG_CUDA(cudaMalloc(&tmp_x, (size_t) nchan * n * sizeof(float2)));
    PG_CUDA(cudaMemcpy2DAsync(tmp_x, (size_t) n * 8, x, (size_t) x_stride * 8, (size_t) n * 8, nchan, cudaMemcpyHostToDevice, r.stream));
    PG_CUDA(cudaMalloc(&tmp_out, (size_t) nchan * frames * nbins * sizeof(float)));
    dx = tmp_x;
    dxs = n;
    dout = tmp_out;
    dos = (long long) frames * nbins;
  }
  const int h = N / 2;
  PG_CUDA(cudaMalloc(&zeros, (size_t) nchan * h * sizeof(float2)));
  PG_CUDA(cudaMemsetAsync(zeros, 0, (size_t) nchan * h * sizeof(float2), r.stream));   // padded starts as zeros (fourier.cc:783)
  PG_CUDA(cudaMalloc(&d_fen, (size_t) N * sizeof(float)));
  PG_CUDA(cudaMemcpyAsync(d_fen, fenetre, (size_t) N * sizeof(float), cudaMemcpyHostToDevice, r.stream));
  // blocks per pass: bounded work buffer (<= 256 MiB), at most 65535 frames per launch
  const long long per_block = 2LL * nchan * N2 * (long long) sizeof(float2);
  int nb_max = (int) std::max(1LL, std::min((256LL << 20) / per_block, 65535LL / (2LL * nchan)));
  if(2LL * nchan > 65535) { cleanup(); return fail("tsdgpu_periodogramme_tfd: too many channels"); }
  nb_max = std::min(nb_max, B);
  PG_CUDA(cudaMalloc(&work, (size_t) nchan * nb_max * 2 * N2 * sizeof(float2)));
  int plan_batch = 0;
  for(int b0 = 0; b0 < B && !rc; b0 += nb_max)
  {
    const int nb = std::min(nb_max, B - b0), batch = nchan * nb * 2;
    if(!plan || plan_batch != batch)
    {
      if(plan) { cudaStreamSynchronize(r.stream); fft_plan_destroy(plan); plan = nullptr; }
      if(fft_plan_create(N2, batch, &plan)) { cleanup(); return 1; }
      plan_batch = batch;
    }
    dim3 gg((N2 + 255) / 256, batch);
    ola_gather_fen_kernel<<<gg, 256, 0, r.stream>>>(dx, dxs, zeros, h, work, d_fen, N2, N, Nz, nb, b0, 0);
    rc = fft_exec_device(plan, work, N2, work, N2, true);
    if(rc) break;
    // frames of this pass: per channel [nb * 2] consecutive rows starting at frame 2 * b0
    dim3 gl((nbins + 255) / 256, batch);
    periodo_logmag_kernel<<<gl, 256, 0, r.stream>>>(work, dout, dos, N2, 2 * nb, 2 * b0, frames);
  }
  if(!rc && cudaGetLastError() != cudaSuccess) rc = fail("tsdgpu_periodogramme_tfd: kernel launch failed");
  if(!rc && mem != TSDGPU_DEVICE)
    PG_CUDA(cudaMemcpy2DAsync(out, (size_t) out_stride * 4, tmp_out, (size_t) frames * nbins * 4, (size_t) frames * nbins * 4, nchan,
                              cudaMemcpyDeviceToHost, r.stream));
#undef PG_CUDA
  cleanup();
  return rc;
}

int tsdgpu_ola_state_dims(tsdgpu_ola_t f, int *carry_len, int *svg_len, int *last_len)
{
  if(!f) return fail("tsdgpu_ola_state_dims: null handle");
  if(carry_len) *carry_len = f->carry_len;
  if(svg_len) *svg_len = (f->K > 0) ? 0 : f->Ne;      // overlap-save forms carry no partial sums
  if(last_len) *last_len = f->fen ? f->Ne : 0;
  return 0;
}

int tsdgpu_ola_get_state(tsdgpu_ola_t f, int *residual, long long *blocks_done, void *carry_host, void *svg_host, void *last_host)
{
  TSD_ENTER(f ? f->device : -1);
  if(!f) return fail("tsdgpu_ola_get_state: null handle");
  if(residual) *residual = f->residual;
  if(blocks_done) *blocks_done = f->blocks_done;
  TSD_CUDA(cudaStreamSynchronize(rt().stream));
  if(carry_host) TSD_CUDA(cudaMemcpy(carry_host, f->d_carry[f->cur], (size_t) f->nchan * f->carry_len * sizeof(float2), cudaMemcpyDeviceToHost));
  if(svg_host && f->K == 0) TSD_CUDA(cudaMemcpy(svg_host, f->d_svg, (size_t) f->nchan * f->Ne * sizeof(float2), cudaMemcpyDeviceToHost));
  if(last_host && f->fen) TSD_CUDA(cudaMemcpy(last_host, f->d_last, (size_t) f->nchan * f->Ne * sizeof(float2), cudaMemcpyDeviceToHost));
  return 0;
}

int tsdgpu_ola_set_state(tsdgpu_ola_t f, int residual, long long blocks_done, const void *carry_host, const void *svg_host,
                         const void *last_host)
{
  TSD_ENTER(f ? f->device : -1);
  if(!f || !carry_host) return fail("tsdgpu_ola_set_state: null argument");
  if(residual < 0 || residual >= f->Ne || blocks_done < 0) return fail("tsdgpu_ola_set_state: invalid counters");
  if(f->K == 0 && !svg_host && blocks_done > 0) return fail("tsdgpu_ola_set_state: the overlap-add form needs the carried partial sums (svg)");
  if(f->fen && !last_host && blocks_done > 0) return fail("tsdgpu_ola_set_state: the windowed mode needs its `last` buffer");
  TSD_CUDA(cudaStreamSynchronize(rt().stream));
  TSD_CUDA(cudaMemcpy(f->d_carry[f->cur], carry_host, (size_t) f->nchan * f->carry_len * sizeof(float2), cudaMemcpyHostToDevice));
  if(svg_host && f->K == 0) TSD_CUDA(cudaMemcpy(f->d_svg, svg_host, (size_t) f->nchan * f->Ne * sizeof(float2), cudaMemcpyHostToDevice));
  if(last_host && f->fen) TSD_CUDA(cudaMemcpy(f->d_last, last_host, (size_t) f->nchan * f->Ne * sizeof(float2), cudaMemcpyHostToDevice));
  f->residual = residual;
  f->blocks_done = blocks_done;
  return 0;
}

int tsdgpu_ola_destroy(tsdgpu_ola_t f)
{
  if(!f) return 0;
  TSD_ENTER(f->device);
  cudaStreamSynchronize(rt().stream);
  cudaFree(f->d_H);
  cudaFree(f->d_carry[0]);
  cudaFree(f->d_carry[1]);
  cudaFree(f->d_svg);
  if(f->d_fen) cudaFree(f->d_fen);
  if(f->d_last) cudaFree(f->d_last);
  if(f->scratch) cudaFree(f->scratch);
  if(f->flags) cudaFree(f->flags);
  if(f->plan) fft_plan_destroy(f->plan);
  if(f->work) cudaFree(f->work);
  ols16k_destroy(f->ols);
  if(f->h_spec) cudaFreeHost(f->h_spec);
  delete f;
  return 0;
}

} // extern "C"

// ---- detector object ---------------------------------------------------------------------------------------------------------
struct tsdgpu_detect_s
{
  int device = 0;
  tsdgpu_ola_s *ola = nullptr;   // correlator: block filter whose taps are the reversed, conjugated, normalised motif / sqrt(N)
  int M = 0, Ne = 0, N = 0, nchan = 0;
  float ratio = 1, K_inv = 1, norme_motif = 1;
  double *P = nullptr;           // [nchan][Ne + n + 2] running sums
  size_t P_cap = 0;
  float2 *d_x = nullptr, *d_corr = nullptr;   // staging of the host-memory entry
  float *d_score = nullptr;
  size_t stage_cap = 0;          // samples per channel
};

extern "C" {

int tsdgpu_detect_create(const float *motif, int M, int Ne, int nchan, tsdgpu_detect_t *out)
{
  TSD_ENTER(-1);
  if(!motif || !out) return fail("tsdgpu_detect_create: null argument");
  if(M < 2 || nchan <= 0) return fail("tsdgpu_detect_create: M must be >= 2 and nchan > 0");
  if(Ne <= 0 && tsdgpu_ola_complexite_optimise(M, nullptr, nullptr, nullptr, &Ne)) return 1;   // detection.cc:136-144
  // norme_motif = ||motif||_2 in float like Tab::norme; motif normalised (detection.cc:124-127)
  double e = 0;
  for(int i = 0; i < M; i++) e += (double) motif[2 * i] * motif[2 * i] + (double) motif[2 * i + 1] * motif[2 * i + 1];
  if(!(e > 0)) return fail("tsdgpu_detect_create: null motif");
  const double nrm = std::sqrt(e);
  const int N = tsdgpu_p2(Ne + M - 1);                   // nb_zeros_min = M - 1 (detection.cc:149-150, fourier.cc:775)
  if(N - Ne > Ne) return fail("tsdgpu_detect_create: N_zeros > Ne (choose a larger Ne)");
  if(2 * M > N) return fail("tsdgpu_detect_create: the motif must fit twice in the transform (detection.cc:166)");
  auto *d = new tsdgpu_detect_s;
  d->device = rt().device;
  d->M = M;
  d->Ne = Ne;
  d->N = N;
  d->nchan = nchan;
  d->norme_motif = (float) nrm;
  d->ratio = std::sqrt(1.0f * N) / std::sqrt(1.0f * M);   // detection.cc:231
  d->K_inv = (float) (1.0 / (double) M);                  // filtre-rt.cc:646
  // corr_ref[t] = (1/sqrt N) sum_q conj(m[q]) stream[t - Ne + q]  (unitary transforms both ways, X *= conj(fft(motif))):
  // as a causal filter, taps h[k] = conj(m[M-1-k]) / sqrt(N) and delay Ne - M + 1
  std::vector<std::complex<double>> taps((size_t) M);
  const double sc = 1.0 / (nrm * std::sqrt((double) N));
  for(int k = 0; k < M; k++)
    taps[k] = std::complex<double>(motif[2 * (M - 1 - k)] * sc, -motif[2 * (M - 1 - k) + 1] * sc);
  // the block-filter object carries the re-blocking state and the input history; gains as data for the N-point path
  std::vector<float> H((size_t) 2 * N);
  {
    std::vector<std::complex<double>> g((size_t) N);
    for(int q = 0; q < M; q++) g[(size_t) ((N - q) % N)] = std::conj(std::complex<double>(motif[2 * q], motif[2 * q + 1])) / (nrm * std::sqrt((double) N));
    // plain DFT by the definition is O(N^2): use the radix-2 recursion on the host (N is a power of two)
    std::vector<std::complex<double>> a = g;
    for(size_t i2 = 1, j = 0; i2 < a.size(); i2++)
    {
      size_t bit = a.size() >> 1;
      for(; j & bit; bit >>= 1) j ^= bit;
      j ^= bit;
      if(i2 < j) std::swap(a[i2], a[j]);
    }
    for(size_t len = 2; len <= a.size(); len <<= 1)
    {
      const double ang = -2.0 * M_PI / (double) len;
      for(size_t i2 = 0; i2 < a.size(); i2 += len)
        for(size_t k = 0; k < len / 2; k++)
        {
          const std::complex<double> w = std::polar(1.0, ang * (double) k), u = a[i2 + k], t = a[i2 + k + len / 2] * w;
          a[i2 + k] = u + t;
          a[i2 + k + len / 2] = u - t;
        }
    }
    for(int i2 = 0; i2 < N; i2++) { H[2 * i2] = (float) a[i2].real(); H[2 * i2 + 1] = (float) a[i2].imag(); }
  }
  tsdgpu_ola_t o = nullptr;
  if(ola_create(Ne, M - 1, H.data(), 0, nullptr, nchan, &o)) { delete d; return 1; }
  d->ola = o;
  // single-SM overlap-save kernel when the motif is short enough for it
  const char *v = getenv("TSDGPU_OLA_OLS");
  if(!(v && v[0] == '0'))
  {
    Ols16k *k16 = nullptr;
    if(ols16k_create_taps(taps.data(), M, &k16)) { tsdgpu_ola_destroy(o); delete d; return 1; }
    if(k16)
    {
      o->ols = k16;
      o->K = M;
      o->delay = Ne - M + 1;
      o->fused = false;
      // the single-SM kernel looks back delay + overlap + residual samples: grow the history the object was created with
      const int need = 2 * Ne + k16->O + 2;
      if(o->carry_len < need)
      {
        cudaError_t e2 = cudaSuccess;
        for(int i = 0; i < 2 && e2 == cudaSuccess; i++)
        {
          cudaFree(o->d_carry[i]);
          o->d_carry[i] = nullptr;
          e2 = cudaMalloc(&o->d_carry[i], (size_t) nchan * need * sizeof(float2));
          if(e2 == cudaSuccess) e2 = cudaMemset(o->d_carry[i], 0, (size_t) nchan * need * sizeof(float2));
        }
        if(e2 != cudaSuccess)
        {
          tsdgpu_ola_destroy(o);
          delete d;
          return fail(std::string("tsdgpu_detect_create: ") + cudaGetErrorString(e2));
        }
        o->carry_len = need;
      }
    }
  }
  *out = d;
  return 0;
}

int tsdgpu_detect_dims(tsdgpu_detect_t d, int *Ne, int *N, int *M, int *delais_corr, float *norme_motif)
{
  if(!d) return fail("tsdgpu_detect_dims: null handle");
  if(Ne) *Ne = d->Ne;
  if(N) *N = d->N;
  if(M) *M = d->M;
  if(delais_corr) *delais_corr = d->Ne;                    // detection.cc:170
  if(norme_motif) *norme_motif = d->norme_motif;
  return 0;
}

static int detect_run_device(tsdgpu_detect_s *d, const float2 *x, long long xs, int n, float *score, long long ss, float2 *corr,
                             long long cs)
{
  Runtime &r = rt();
  tsdgpu_ola_s *o = d->ola;
  // running sums over [carry tail, x] BEFORE the block filter advances its history
  const size_t needP = (size_t) d->nchan * ((size_t) d->Ne + n + 2);
  if(needP > d->P_cap)
  {
    if(d->P) { TSD_CUDA(cudaStreamSynchronize(r.stream)); cudaFree(d->P); d->P = nullptr; d->P_cap = 0; }
    TSD_CUDA(cudaMalloc(&d->P, needP * sizeof(double)));
    d->P_cap = needP;
  }
  detect_energy_scan_kernel<<<d->nchan, 1024, 0, r.stream>>>(x, xs, o->d_carry[o->cur], o->carry_len, d->Ne, n, d->P);
  TSD_LAUNCH_CHECK();
  long long n_out = 0;
  if(ola_run_device(o, x, xs, n, corr, cs, &n_out)) return 1;
  if(n_out != n) return fail("tsdgpu_detect_step: internal: correlator emitted a different count");
  dim3 grid((unsigned) std::min(1024, (n + 255) / 256), d->nchan);
  detect_score_kernel<<<grid, 256, 0, r.stream>>>(corr, cs, d->P, score, ss, d->Ne, d->M, n, d->ratio, d->K_inv);
  TSD_LAUNCH_CHECK();
  return 0;
}

int tsdgpu_detect_step(tsdgpu_detect_t d, const void *x, long long xs, int n, float *score, long long ss, void *corr, long long cs,
                       int mem)
{
  TSD_ENTER(d ? d->device : -1);
  if(!d || !x || !score || !corr) return fail("tsdgpu_detect_step: null argument");
  if(n <= 0 || n % d->Ne) return fail("tsdgpu_detect_step: n must be a positive multiple of Ne (the reference asserts corr.rows() == n, detection.cc:227)");
  if(xs < n || ss < n || cs < n) return fail("tsdgpu_detect_step: stride smaller than n");
  if(mem == TSDGPU_DEVICE) return detect_run_device(d, (const float2 *) x, xs, n, score, ss, (float2 *) corr, cs);
  Runtime &r = rt();
  if((size_t) n > d->stage_cap)
  {
    TSD_CUDA(cudaStreamSynchronize(r.stream));
    if(d->d_x) cudaFree(d->d_x);
    if(d->d_corr) cudaFree(d->d_corr);
    if(d->d_score) cudaFree(d->d_score);
    d->d_x = d->d_corr = nullptr;
    d->d_score = nullptr;
    d->stage_cap = 0;
    TSD_CUDA(cudaMalloc(&d->d_x, (size_t) d->nchan * n * sizeof(float2)));
    TSD_CUDA(cudaMalloc(&d->d_corr, (size_t) d->nchan * n * sizeof(float2)));
    TSD_CUDA(cudaMalloc(&d->d_score, (size_t) d->nchan * n * sizeof(float)));
    d->stage_cap = (size_t) n;
  }
  TSD_CUDA(cudaMemcpy2DAsync(d->d_x, (size_t) n * 8, x, (size_t) xs * 8, (size_t) n * 8, d->nchan, cudaMemcpyHostToDevice, r.stream));
  if(detect_run_device(d, d->d_x, n, n, d->d_score, n, d->d_corr, n)) return 1;
  TSD_CUDA(cudaMemcpy2DAsync(score, (size_t) ss * 4, d->d_score, (size_t) n * 4, (size_t) n * 4, d->nchan, cudaMemcpyDeviceToHost, r.stream));
  TSD_CUDA(cudaMemcpy2DAsync(corr, (size_t) cs * 8, d->d_corr, (size_t) n * 8, (size_t) n * 8, d->nchan, cudaMemcpyDeviceToHost, r.stream));
  TSD_CUDA(cudaStreamSynchronize(r.stream));
  return 0;
}

int tsdgpu_detect_destroy(tsdgpu_detect_t d)
{
  if(!d) return 0;
  TSD_ENTER(d->device);
  cudaStreamSynchronize(rt().stream);
  tsdgpu_ola_destroy(d->ola);
  if(d->P) cudaFree(d->P);
  if(d->d_x) cudaFree(d->d_x);
  if(d->d_corr) cudaFree(d->d_corr);
  if(d->d_score) cudaFree(d->d_score);
  delete d;
  return 0;
}

} // extern "C"
