// Batched complex FFT plans on sm_100a.  Replaces TFRPlanDefaut / tfr_radix2 for power-of-two
// sizes (reference fourier.cc:61-121,360-467): unitary DFT in both directions
// (X /= sqrt(N), fourier.cc:119-120).
//
//  * N = 65536: four-step 256 x 256 transform in ONE persistent kernel.  Work items are
//    4096-point tiles (fft_tiles.cuh); stage A does the 256 column transforms of a tile group
//    + the W_N twiddles and writes an intermediate that lives in a small ring of L2-resident
//    scratch slots, stage B does the row transforms and writes the result in natural order.
//    CTAs draw items from a global ticket counter; a B item spins (acquire) on the per-transform
//    counter that the A items bump (release).  Items only ever wait on smaller tickets, so the
//    schedule cannot deadlock whatever the number of resident CTAs.  HBM traffic is one read and
//    one write of the data (16 B per point per transform).
//  * other power-of-two N: Stockham radix-2, one launch per pass (correct for every size, not
//    tuned; the reference's own structure, fourier.cc:86-117).
#include "fft_tiles.cuh"
#include "fft_plan.h"
#include "host_pipe.cuh"
#include "tsdgpu.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

namespace tsdgpu {

// ------------------------------------------------------------------ generic radix-2 passes
// out[k*pas + m] = e + w*g ; out[N/2 + k*pas + m] = e - w*g with e = in[2*k*pas + m], g = in[2*k*pas + m + pas]
template<bool INV>
__global__ void fft_radix2_pass(const float2 *in, long long in_stride, float2 *out, long long out_stride, int N, int n,
                                int batch, float scale)
{
  const int half = N >> 1;
  const long long total = (long long) half * batch;
  const int pas = N / (2 * n);
  const float two_over_N = 2.0f / (float) N;
  for(long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long) gridDim.x * blockDim.x)
  {
    const int b = (int) (idx / half), i = (int) (idx - (long long) b * half);
    const int k = i / pas, m = i - k * pas;
    const float2 *src = in + (long long) b * in_stride + (long long) 2 * k * pas + m;
    const float2 e = src[0], g = src[pas];
    const float2 w = twiddle<INV>((unsigned) (k * pas), two_over_N);
    const float2 p = cmul(w, g);
    float2 *dst = out + (long long) b * out_stride + (long long) k * pas + m;
    dst[0] = make_float2((e.x + p.x) * scale, (e.y + p.y) * scale);
    dst[half] = make_float2((e.x - p.x) * scale, (e.y - p.y) * scale);
  }
}

__global__ void copy_strided(const float2 *in, long long in_stride, float2 *out, long long out_stride, int n, int batch)
{
  const long long total = (long long) n * batch;
  for(long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long) gridDim.x * blockDim.x)
  {
    const int b = (int) (idx / n), i = (int) (idx - (long long) b * n);
    out[(long long) b * out_stride + i] = in[(long long) b * in_stride + i];
  }
}

// ------------------------------------------------------------------ n not a power of two
// even n (fourier.cc:438-462): xe(i) = x(2i), xo(i) = x(2i+1); E, O = transforms of n/2;
// y.head = E + rot.head * O, y.tail = E + rot.tail * O (conj(rot) for the inverse); y *= 1/sqrt(2)
__global__ void fft_even_split(const float2 *x, long long x_stride, float2 *w, int m, int batch)
{
  const long long total = (long long) 2 * m * batch;
  for(long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long) gridDim.x * blockDim.x)
  {
    const int b = (int) (idx / (2 * m)), i = (int) (idx - (long long) b * 2 * m);
    w[((long long) 2 * b + (i & 1)) * m + (i >> 1)] = x[(long long) b * x_stride + i];
  }
}
__global__ void fft_even_combine(const float2 *w, const float2 *rot, float2 *y, long long y_stride, int m, int batch, int inverse)
{
  const long long total = (long long) 2 * m * batch;
  const float s = 1.0f / sqrtf(2.0f);
  for(long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long) gridDim.x * blockDim.x)
  {
    const int b = (int) (idx / (2 * m)), k = (int) (idx - (long long) b * 2 * m), kk = k < m ? k : k - m;
    const float2 E = w[(long long) 2 * b * m + kk], O = w[((long long) 2 * b + 1) * m + kk];
    float2 r = rot[k];
    if(inverse) r.y = -r.y;
    const float2 v = cadd(E, cmul(r, O));
    y[(long long) b * y_stride + k] = make_float2(v.x * s, v.y * s);
  }
}
// odd n (tfr_czt_impl, fourier.cc:237-255): xp.head(n) = x * chirp.tail(n), zero-padded to n2
__global__ void fft_czt_pre(const float2 *x, long long x_stride, const float2 *chirp_tail, float2 *w, int n, int n2, int batch)
{
  const long long total = (long long) n2 * batch;
  for(long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long) gridDim.x * blockDim.x)
  {
    const int b = (int) (idx / n2), i = (int) (idx - (long long) b * n2);
    w[idx] = i < n ? cmul(x[(long long) b * x_stride + i], chirp_tail[i]) : make_float2(0.f, 0.f);
  }
}
__global__ void fft_mul_vec(float2 *w, const float2 *v, int n2, long long total)
{
  for(long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long) gridDim.x * blockDim.x)
    w[idx] = cmul(w[idx], v[idx % n2]);
}
// y = y2.segment(n-1, n) * chirp.tail(n) * (sqrt(n2) / sqrt(n)); inverse: Y(0) = X(0), Y(k) = X(n-k) (tfr2itfr, fourier.cc:258-277)
__global__ void fft_czt_post(const float2 *w, const float2 *chirp_tail, float2 *y, long long y_stride, int n, int n2, int batch, float scale,
                             int inverse)
{
  const long long total = (long long) n * batch;
  for(long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long) gridDim.x * blockDim.x)
  {
    const int b = (int) (idx / n), k = (int) (idx - (long long) b * n);
    const float2 v = cmul(w[(long long) b * n2 + n - 1 + k], chirp_tail[k]);
    const int ko = (inverse && k) ? n - k : k;
    y[(long long) b * y_stride + ko] = make_float2(v.x * scale, v.y * scale);
  }
}

// ------------------------------------------------------------------ 16 <= N <= 16384: one launch, shared memory
// Stockham autosort with 16 points per thread: radix-16 passes (fft16 in registers), then one radix-4 and/or one
// radix-2 pass for the remaining factor.  Pass with sub-transform length Ns and radix R, butterfly j:
//   k = j mod Ns;  v[t] = in[j + t*N/R] * W_(Ns*R)^(k*t);  DFT_R;  out[(j - k)*R + k + t*Ns] = v[t]
// All inputs of a pass are read into registers before any output is written (barrier in between), so one buffer
// of N points per transform is enough; the first pass reads global memory and the last one writes it, both fully
// coalesced.  N/16 threads per transform; CTAs of 256 threads pack 4096/N transforms when N < 4096.
template<bool INV>
__global__ void __launch_bounds__(1024, 1) fft_smem_kernel(const float2 *x, long long x_stride, float2 *y, long long y_stride, int N, int batch,
                                                           float scale, int R = 1)
{
  // R > 1 (decimation in time of a larger transform, fft_split_*): transform b = t * R + r reads x[t][r + R m], m < N
  extern __shared__ float2 fft_sm[];
  const int T = N >> 4;
  const int local = threadIdx.x / T, j = threadIdx.x - local * T;
  const long long b = (long long) blockIdx.x * (blockDim.x / T) + local;
  const bool active = b < batch;
  float2 *sm = fft_sm + (size_t) local * N;
  const long long es = R;   // element stride of the input
  const float2 *in = R == 1 ? x + b * x_stride : x + (b / R) * x_stride + (b % R);
  float2 *out = y + b * y_stride;
  float2 v[16];
  int Ns = 1, rem = N;
  bool first = true;
  // ---- radix-16 passes
  while(rem >= 16)
  {
    rem >>= 4;
    const bool last = rem == 1;
    if(active)
    {
#pragma unroll
      for(int t = 0; t < 16; t++) v[t] = first ? in[(j + t * T) * es] : sm[j + t * T];
    }
    if(!first) __syncthreads();
    const int k = j & (Ns - 1);
    if(Ns > 1) mul_geometric(v, make_float2(1.f, 0.f), twiddle<INV>((unsigned) k, 2.0f / (float) (Ns * 16)));
    fft16<INV>(v);
    const int j0 = (j - k) * 16 + k;
    if(active)
    {
      if(last)
      {
#pragma unroll
        for(int t = 0; t < 16; t++) out[j0 + t * Ns] = make_float2(v[t].x * scale, v[t].y * scale);
      }
      else
      {
#pragma unroll
        for(int t = 0; t < 16; t++) sm[j0 + t * Ns] = v[t];
      }
    }
    if(!last) __syncthreads();
    Ns <<= 4;
    first = false;
  }
  // ---- one radix-4 pass (remaining factor 4 or 8): 4 butterflies per thread
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
        for(int t = 0; t < 4; t++) v[4 * m + t] = first ? in[(j + m * T + t * Q) * es] : sm[j + m * T + t * Q];
    }
    if(!first) __syncthreads();
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
          if(last) out[j0 + t * Ns] = make_float2(v[4 * m + t].x * scale, v[4 * m + t].y * scale);
          else sm[j0 + t * Ns] = v[4 * m + t];
        }
      }
    }
    if(!last) __syncthreads();
    Ns <<= 2;
    first = false;
  }
  // ---- one radix-2 pass (remaining factor 2): 8 butterflies per thread, always the last pass
  if(rem == 2)
  {
    const int Hh = N >> 1;
    if(active)
    {
#pragma unroll
      for(int m = 0; m < 8; m++)
      {
        v[2 * m] = first ? in[(j + m * T) * es] : sm[j + m * T];
        v[2 * m + 1] = first ? in[(j + m * T + Hh) * es] : sm[j + m * T + Hh];
      }
    }
    // no barrier needed: this pass writes global memory only
#pragma unroll
    for(int m = 0; m < 8; m++)
    {
      const int jj = j + m * T, k = jj & (Ns - 1);
      const float2 wb = Ns > 1 ? cmul(v[2 * m + 1], twiddle<INV>((unsigned) k, 2.0f / (float) (Ns * 2))) : v[2 * m + 1];
      const float2 a = cadd(v[2 * m], wb), d = csub(v[2 * m], wb);
      const int j0 = (jj - k) * 2 + k;
      if(active)
      {
        out[j0] = make_float2(a.x * scale, a.y * scale);
        out[j0 + Ns] = make_float2(d.x * scale, d.y * scale);
      }
    }
  }
}

// ------------------------------------------------------------------ 2^k > 16384 (other than 65536): N = R x 16384
// Decimation in time: F_r = FFT_M(x[r + R m]) (fft_smem_kernel with element stride R -> work[t][r][k], contiguous), then
//   X[k + M q] = sum_r W_R^(r q) * (W_N^(r k) F_r[k]),  q < R
// one thread per (t, k): R coalesced loads, twiddles on exact dyadic angles, an R-point DFT in registers, R coalesced
// stores.  Two passes over the data instead of log2(N) (the reference's own radix-2 structure, fourier.cc:86-117).
// G > 1 (two-level plans, N >= 2^19): the R inputs of item (t, g) are interleaved with those of the other G - 1 items of
// transform t, i.e. sub-transform r of item g sits at row g + G r of work[t][G R][M]
template<bool INV, int R>
__global__ void fft_split_combine_kernel(const float2 *w, float2 *y, long long y_stride, int M, int batch, float scale, int G = 1)
{
  // grid-stride loop (the launches size their grids with grid_for, which caps the number of blocks)
  for(long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x; idx < (long long) M * batch; idx += (long long) gridDim.x * blockDim.x)
  {
  const int t = (int) (idx / M), k = (int) (idx - (long long) t * M);
  const float2 *src = w + ((long long) (t / G) * G * R + (t % G)) * M + k;
  const long long rs = (long long) G * M;
  float2 v[R];
#pragma unroll
  for(int r = 0; r < R; r++) v[r] = src[r * rs];
  const float two_over_n = 2.0f / (float) ((long long) R * M);
#pragma unroll
  for(int r = 1; r < R; r++) v[r] = cmul(v[r], twiddle<INV>((unsigned) ((long long) r * k), two_over_n));
  if(R == 2)
  {
    const float2 a = cadd(v[0], v[1]), d = csub(v[0], v[1]);
    v[0] = a;
    v[1] = d;
  }
  else if(R == 4) fft4<INV>(v[0], v[1], v[2], v[3]);
  else if(R == 8)
  {
    // two 4-point transforms of the even / odd inputs, then X[q] = E[q] + W8^q O[q], X[q + 4] = E[q] - W8^q O[q]
    float2 e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
    fft4<INV>(e[0], e[1], e[2], e[3]);
    fft4<INV>(o[0], o[1], o[2], o[3]);
    const float h = 0.70710678118654752f, sg = INV ? 1.f : -1.f;
    const float2 w8[4] = {make_float2(1.f, 0.f), make_float2(h, sg * h), make_float2(0.f, sg), make_float2(-h, sg * h)};
#pragma unroll
    for(int q = 0; q < 4; q++)
    {
      const float2 ow = q == 0 ? o[0] : cmul(o[q], w8[q]);
      v[q] = cadd(e[q], ow);
      v[q + 4] = csub(e[q], ow);
    }
  }
  else
  {
    float2 u[16];
#pragma unroll
    for(int r = 0; r < 16; r++) u[r] = v[r % R];
    fft16<INV>(u);
#pragma unroll
    for(int r = 0; r < 16; r++) v[r % R] = u[r];
  }
  float2 *dst = y + (long long) t * y_stride + k;
#pragma unroll
  for(int q = 0; q < R; q++) dst[(long long) q * M] = make_float2(v[q].x * scale, v[q].y * scale);
  }
}

// ------------------------------------------------------------------ N = 65536 pipeline
struct Fft64kParams
{
  const float2 *x;
  float2 *y;
  long long x_stride, y_stride;
  float2 *scratch;        // [ring][65536]
  unsigned *done_a;       // [batch] number of finished A items (16 per transform)
  unsigned *done_b;       // [batch]
  unsigned *ticket;
  int batch, ring, lag;
  const float2 *tw4;      // rt().tw4: four-step twiddle tables
};

constexpr int FFT_NT = 256;

template<bool INV>
__global__ void __launch_bounds__(FFT_NT, 4) fft64k_kernel(Fft64kParams p)
{
  __shared__ float2 sm[4096];
  __shared__ float4 tw[256];
  __shared__ unsigned s_ticket[2];
  const int tid = threadIdx.x, hi = tid >> 4, lo = tid & 15;
  const float inv256 = 1.0f / 256.0f;
  const unsigned total = (unsigned) (p.batch + p.lag) * 32u;
  const unsigned full = 16u * ITEM_WARPS;
  fill_tw256(tw, tid);
  if(tid == 0) s_ticket[0] = atomicAdd(p.ticket, 1u);
  __syncthreads();

  for(int it = 0;; it ^= 1)
  {
    const unsigned ticket = s_ticket[it];
    if(ticket >= total) break;
    const int s = (int) (ticket >> 5), sub = (int) (ticket & 31u), g = sub & 15;
    const bool is_a = sub < 16;
    const int t = is_a ? s : s - p.lag;
    const bool valid = t >= 0 && t < p.batch;
    if(tid == 0)
    {
      s_ticket[it ^ 1] = atomicAdd(p.ticket, 1u);   // next item, fetched early
      if(valid)
      {
        if(!is_a) spin_until(p.done_a + t, full);                       // all 16 column tiles written
        else if(t >= p.ring) spin_until(p.done_b + (t - p.ring), full); // ring slot free again
      }
    }
    __syncthreads();
    if(!valid) continue;
    float2 v[16];
    if(is_a)
    {
      // ---- stage A: columns n2 in [16g, 16g+16), transform over n1
      const float2 *x = p.x + (long long) t * p.x_stride + hi * 256 + 16 * g + lo;
#pragma unroll
      for(int j = 0; j < 16; j++) v[j] = ldg_stream(x + j * 4096);
      fft256_cols<INV>(v, sm, tw, hi, lo);
      // v[p2] = Y[k1 = hi + 16 p2][n2]; four-step twiddle W_N^(n2*k1)
      mul_fourstep_cols<INV>(v, p.tw4, 16 * g + lo, hi);
      float2 *sc = p.scratch + (long long) (t % p.ring) * 65536 + hi * 256 + 16 * g + lo;
#pragma unroll
      for(int p2 = 0; p2 < 16; p2++) sc[p2 * 4096] = v[p2];
      warp_release(p.done_a + t);
    }
    else
    {
      // ---- stage B: rows k1 in [16g, 16g+16), transform over n2, natural-order output
      const float2 *sc = p.scratch + (long long) (t % p.ring) * 65536 + (16 * g + hi) * 256 + lo;
#pragma unroll
      for(int j = 0; j < 16; j++) v[j] = __ldcg(sc + 16 * j);
      fft256_rows_a<INV>(v, sm, tw, hi, lo);
      // thread (hi = k', lo = r): v[k2] = X[(16g + r) + 256*(k' + 16*k2)]
      float2 *y = p.y + (long long) t * p.y_stride + hi * 256 + 16 * g + lo;
#pragma unroll
      for(int k2 = 0; k2 < 16; k2++) stg_stream(y + k2 * 4096, make_float2(v[k2].x * inv256, v[k2].y * inv256));
      warp_release(p.done_b + t);
    }
  }
}

// ---- staged form (default): one kernel per stage and chunk of transforms, chunks round-robin over the
// auxiliary streams; the hand-over between the stages is the kernel boundary (see ola.cu).
struct Fft64kStageParams
{
  const float2 *x;
  float2 *y;
  long long x_stride, y_stride;
  float2 *scratch;        // this chunk's scratch, [gridDim.y][65536]
  const float4 *tw;       // rt().tw256
  const float2 *tw4;      // rt().tw4
  int t0;                 // first transform of the chunk
};

template<bool INV, int STAGE> __global__ void __launch_bounds__(FFT_NT, 4) fft64k_stage(Fft64kStageParams p)
{
  __shared__ float2 sm[4096];
  __shared__ float4 tw[256];
  const int tid = threadIdx.x, hi = tid >> 4, lo = tid & 15;
  const int g = blockIdx.x, t = p.t0 + blockIdx.y;
  float2 *sc = p.scratch + (long long) blockIdx.y * 65536;
  float2 v[16];
  if(STAGE == 0)
  {
    // columns n2 in [16g, 16g+16), transform over n1, four-step twiddle W_N^(n2*k1)
    const float2 *x = p.x + (long long) t * p.x_stride + hi * 256 + 16 * g + lo;
#pragma unroll
    for(int j = 0; j < 16; j++) v[j] = ldg_stream(x + j * 4096);
    fill_tw256_from(tw, p.tw, tid, INV);
    const float2 tb = tw4_load<INV>(p.tw4 + hi * 256 + 16 * g + lo), ts = tw4_load<INV>(p.tw4 + 8192 + 16 * g + lo);
    __syncthreads();
    fft256_cols<INV, true>(v, sm, tw, hi, lo);
    mul_geometric(v, tb, ts);   // four-step twiddle W_N^(n2*k1)
    float2 *dst = sc + hi * 256 + 16 * g + lo;
#pragma unroll
    for(int p2 = 0; p2 < 16; p2++) dst[p2 * 4096] = v[p2];
  }
  else
  {
    // rows k1 in [16g, 16g+16), transform over n2, natural-order output, unitary scaling
    const float2 *row = sc + (16 * g + hi) * 256 + lo;
#pragma unroll
    for(int j = 0; j < 16; j++) v[j] = __ldcg(row + 16 * j);
    fill_tw256_from(tw, p.tw, tid, INV);
    __syncthreads();
    fft256_rows_a<INV, true>(v, sm, tw, hi, lo);
    // thread (hi = k', lo = r): v[k2] = X[(16g + r) + 256*(k' + 16*k2)]
    const float inv256 = 1.0f / 256.0f;
    float2 *y = p.y + (long long) t * p.y_stride + hi * 256 + 16 * g + lo;
#pragma unroll
    for(int k2 = 0; k2 < 16; k2++) stg_stream(y + k2 * 4096, make_float2(v[k2].x * inv256, v[k2].y * inv256));
  }
}

} // namespace tsdgpu

using namespace tsdgpu;

static int grid_for(long long work, int threads)
{
  long long blocks = (work + threads - 1) / threads;
  long long cap = (long long) rt().num_sms * 16;
  return (int) std::max(1LL, std::min(blocks, cap));
}

namespace tsdgpu {

int fft_plan_create(int n, int batch, tsdgpu_fft_s **out)
{
  if(n <= 0) return fail("tsdgpu_fft_plan: n must be >= 1");
  if(batch <= 0) return fail("tsdgpu_fft_plan: batch must be > 0");
  if(n > (1 << 24)) return fail("tsdgpu_fft_plan: n > 2^24 not supported");
  auto *p = new tsdgpu_fft_s;
  p->device = rt().device;
  p->n = n;
  p->batch = batch;
  p->batch_created = batch;
  if(n & (n - 1))
  {
    cudaError_t e = cudaSuccess;
    if((n & 1) == 0)
    {
      // even: decompose while even (fourier.cc:385-389); rotations = tfr_rotation<float>(n): cdouble recurrence (fourier.cc:32-46)
      const int m = n / 2;
      if((long long) 2 * batch > (1LL << 30) || fft_plan_create(m, 2 * batch, &p->sub)) { fft_plan_destroy(p); return 1; }
      std::vector<float2> rot((size_t) n);
      double rr = 1.0, ri = 0.0;
      const double wr = cos(-2.0 * M_PI / n), wi = sin(-2.0 * M_PI / n);
      for(int i = 0; i < n; i++)
      {
        rot[i] = make_float2((float) rr, (float) ri);
        const double t = rr * wr - ri * wi;
        ri = rr * wi + ri * wr;
        rr = t;
      }
      e = cudaMalloc(&p->d_rot, (size_t) n * sizeof(float2));
      if(e == cudaSuccess) e = cudaMemcpy(p->d_rot, rot.data(), (size_t) n * sizeof(float2), cudaMemcpyHostToDevice);
      if(e == cudaSuccess) e = cudaMalloc(&p->nwork, (size_t) 2 * batch * m * sizeof(float2));
    }
    else
    {
      // odd: chirp-z.  n2 = p2(2n-1); t = square(linspace(-(n-1), n-1, 2n-1)) / 2; t *= -2*pi/n; chirp = polar(t)
      // (fourier.cc:392-398) -- float arithmetic reproduced operation by operation: for large n the reference's own
      // chirp carries float phase errors far above 1e-5, and parity is against the reference.
      if(2LL * n - 1 > (1 << 24)) { fft_plan_destroy(p); return fail("tsdgpu_fft_plan: odd n too large for the chirp-z plan"); }
      p->n2 = tsdgpu_p2(2 * n - 1);
      if(fft_plan_create(p->n2, batch, &p->sub)) { fft_plan_destroy(p); return 1; }
      const float c = (float) (-2.0 * M_PI / n);
      std::vector<float2> chirp((size_t) 2 * n - 1), icp((size_t) p->n2, make_float2(0.f, 0.f));
      for(int i = 0; i < 2 * n - 1; i++)
      {
        const float lin = (float) (-(double) (n - 1) + (double) i);
        float t = (lin * lin) / 2.0f;
        t = t * c;
        chirp[i] = make_float2(1.0f * cosf(t), 1.0f * sinf(t));      // std::polar((float) 1, t)
        icp[i] = make_float2(chirp[i].x, -chirp[i].y);               // icp.head(2n-1) = chirp.conjugate() (fourier.cc:246-247)
      }
      e = cudaMalloc(&p->d_chirp, (size_t) n * sizeof(float2));
      if(e == cudaSuccess) e = cudaMemcpy(p->d_chirp, chirp.data() + (n - 1), (size_t) n * sizeof(float2), cudaMemcpyHostToDevice);
      if(e == cudaSuccess) e = cudaMalloc(&p->d_Xc, (size_t) p->n2 * sizeof(float2));
      if(e == cudaSuccess) e = cudaMemcpy(p->d_Xc, icp.data(), (size_t) p->n2 * sizeof(float2), cudaMemcpyHostToDevice);
      if(e == cudaSuccess) e = cudaMalloc(&p->nwork, (size_t) batch * p->n2 * sizeof(float2));
      if(e == cudaSuccess)
      {
        // Xc = unitary transform of icp, computed once (the reference redoes it at every call, fourier.cc:252)
        tsdgpu_fft_s *one = nullptr;
        if(fft_plan_create(p->n2, 1, &one) || fft_exec_device(one, p->d_Xc, p->n2, p->d_Xc, p->n2, true)) { fft_plan_destroy(one); fft_plan_destroy(p); return 1; }
        e = cudaStreamSynchronize(rt().stream);
        fft_plan_destroy(one);
      }
    }
    if(e != cudaSuccess)
    {
      fft_plan_destroy(p);
      return fail(std::string("tsdgpu_fft_plan: ") + cudaGetErrorString(e));
    }
    *out = p;
    return 0;
  }
  if(n == 65536)
  {
    p->ring = 64;
    p->lag = 32;
    p->staged = 1;
    if(const char *v = getenv("TSDGPU_FFT_MODE")) { p->staged = v[0] == 'p' ? 0 : 1; p->pipe = v[0] == 't' ? 1 : 0; }   // tma (default) | staged | persistent
    if(const char *v = getenv("TSDGPU_FFT_PRING")) p->pipe_ring = std::max(2, atoi(v));
    if(const char *v = getenv("TSDGPU_FFT_CHUNK")) p->chunk = std::max(1, atoi(v));
    if(const char *v = getenv("TSDGPU_FFT_STREAMS")) p->nstreams = std::min((int) Runtime::MAX_AUX, std::max(1, atoi(v)));
    if(aux_init()) { fft_plan_destroy(p); return 1; }   // twiddle tables (and the auxiliary streams of the staged form)
    if(const char *v = getenv("TSDGPU_FFT_LAG")) p->lag = std::max(1, atoi(v));
    if(const char *v = getenv("TSDGPU_FFT_RING")) p->ring = atoi(v);
    if(p->ring <= p->lag) p->ring = p->lag + 16;
    size_t slots = p->staged ? (size_t) p->chunk * p->nstreams : (size_t) p->ring;
    if(p->pipe) slots = std::max(slots, (size_t) p->pipe_ring * (size_t) std::max(1, rt().num_sms / 16));
    if(cudaMalloc(&p->scratch, slots * 65536 * sizeof(float2)) != cudaSuccess ||
       cudaMalloc(&p->flags, ((size_t) 2 * batch + 1) * sizeof(unsigned)) != cudaSuccess)
    {
      fft_plan_destroy(p);
      return fail("tsdgpu_fft_plan: out of device memory");
    }
    int occ_f = 0, occ_i = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, fft64k_kernel<false>, FFT_NT, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_i, fft64k_kernel<true>, FFT_NT, 0);
    p->ctas = rt().num_sms * std::max(1, std::min(occ_f, occ_i));
  }
  *out = p;
  return 0;
}

void fft_plan_destroy(tsdgpu_fft_s *p)
{
  if(!p) return;
  if(p->sub) fft_plan_destroy(p->sub);
  if(p->d_rot) cudaFree(p->d_rot);
  if(p->d_chirp) cudaFree(p->d_chirp);
  if(p->d_Xc) cudaFree(p->d_Xc);
  if(p->nwork) cudaFree(p->nwork);
  if(p->scratch) cudaFree(p->scratch);
  if(p->flags) cudaFree(p->flags);
  if(p->work[0]) cudaFree(p->work[0]);
  if(p->work[1]) cudaFree(p->work[1]);
  delete p;
}

// work buffers of the radix-2 path: sized for the live batch (the host entry temporarily runs the plan with a
// smaller batch, fft_exec below), regrown when a later call needs more
static int ensure_work(tsdgpu_fft_s *p, int which)
{
  const size_t need = (size_t) p->n * p->batch;
  if(p->work[which] && p->work_cap[which] >= need) return 0;
  if(p->work[which])
  {
    TSD_CUDA(cudaStreamSynchronize(rt().stream));
    cudaFree(p->work[which]);
    p->work[which] = nullptr;
    p->work_cap[which] = 0;
  }
  const size_t cap = std::max(need, (size_t) p->n * p->batch_created);
  TSD_CUDA(cudaMalloc(&p->work[which], cap * sizeof(float2)));
  p->work_cap[which] = cap;
  return 0;
}

int fft_exec_device(tsdgpu_fft_s *p, const float2 *x, long long xs, float2 *y, long long ys, bool forward)
{
  Runtime &r = rt();
  const int N = p->n, batch = p->batch;
  if(p->sub && p->n2 == 0)
  {
    const int m = N / 2;
    const int grid = grid_for((long long) N * batch, 256);
    fft_even_split<<<grid, 256, 0, r.stream>>>(x, xs, p->nwork, m, batch);
    TSD_LAUNCH_CHECK();
    if(fft_exec_device(p->sub, p->nwork, m, p->nwork, m, forward)) return 1;
    fft_even_combine<<<grid, 256, 0, r.stream>>>(p->nwork, p->d_rot, y, ys, m, batch, forward ? 0 : 1);
    TSD_LAUNCH_CHECK();
    return 0;
  }
  if(p->sub)
  {
    const int n2 = p->n2;
    const long long total = (long long) n2 * batch;
    fft_czt_pre<<<grid_for(total, 256), 256, 0, r.stream>>>(x, xs, p->d_chirp, p->nwork, N, n2, batch);
    TSD_LAUNCH_CHECK();
    if(fft_exec_device(p->sub, p->nwork, n2, p->nwork, n2, true)) return 1;
    fft_mul_vec<<<grid_for(total, 256), 256, 0, r.stream>>>(p->nwork, p->d_Xc, n2, total);
    TSD_LAUNCH_CHECK();
    if(fft_exec_device(p->sub, p->nwork, n2, p->nwork, n2, false)) return 1;
    fft_czt_post<<<grid_for((long long) N * batch, 256), 256, 0, r.stream>>>(p->nwork, p->d_chirp, y, ys, N, n2, batch,
                                                                              sqrtf((float) n2) / sqrtf((float) N), forward ? 0 : 1);
    TSD_LAUNCH_CHECK();
    return 0;
  }
  if(N == 1)
  {
    if(x != y || xs != ys)
    {
      copy_strided<<<grid_for(batch, 256), 256, 0, r.stream>>>(x, xs, y, ys, 1, batch);
      TSD_LAUNCH_CHECK();
    }
    return 0;
  }
  if(N == 65536)
  {
    const float2 *src = x;
    long long src_stride = xs;
    if(x == y)
    {
      // stage B of transform t overwrites y[t] while stage A items of the same transform have
      // long finished, but A reads x[t] == y[t] only before B(t) starts: in place is safe.
    }
    if(p->pipe && fft64k_pipe_usable(x, xs, y, ys)) return fft64k_pipe_run(p, x, xs, y, ys, batch, forward);
    if(p->staged)
    {
      KernelTimer timer;
      if(aux_fork(p->nstreams)) return 1;
      Fft64kStageParams sp;
      sp.x = x;
      sp.y = y;
      sp.x_stride = xs;
      sp.y_stride = ys;
      sp.tw = r.tw256;
      sp.tw4 = r.tw4;
      int c = 0;
      for(int t0 = 0; t0 < batch; t0 += p->chunk, c++)
      {
        const int s = c % p->nstreams;
        const dim3 grid(16, std::min(p->chunk, batch - t0));
        sp.t0 = t0;
        sp.scratch = p->scratch + (size_t) s * p->chunk * 65536;
        if(forward)
        {
          fft64k_stage<false, 0><<<grid, FFT_NT, 0, r.aux[s]>>>(sp);
          TSD_LAUNCH_CHECK();
          fft64k_stage<false, 1><<<grid, FFT_NT, 0, r.aux[s]>>>(sp);
          TSD_LAUNCH_CHECK();
        }
        else
        {
          fft64k_stage<true, 0><<<grid, FFT_NT, 0, r.aux[s]>>>(sp);
          TSD_LAUNCH_CHECK();
          fft64k_stage<true, 1><<<grid, FFT_NT, 0, r.aux[s]>>>(sp);
          TSD_LAUNCH_CHECK();
        }
      }
      if(aux_join(p->nstreams)) return 1;
      return 0;
    }
    TSD_CUDA(cudaMemsetAsync(p->flags, 0, ((size_t) 2 * batch + 1) * sizeof(unsigned), r.stream));
    Fft64kParams q;
    q.x = src;
    q.y = y;
    q.x_stride = src_stride;
    q.y_stride = ys;
    q.scratch = p->scratch;
    q.done_a = p->flags;
    q.done_b = p->flags + batch;
    q.ticket = p->flags + 2 * batch;
    q.tw4 = r.tw4;
    q.batch = batch;
    q.ring = p->ring;
    q.lag = p->lag;
    const int grid = std::min(p->ctas, (batch + p->lag) * 32);
    {
      KernelTimer timer;
      if(forward) fft64k_kernel<false><<<grid, FFT_NT, 0, r.stream>>>(q);
      else fft64k_kernel<true><<<grid, FFT_NT, 0, r.stream>>>(q);
      TSD_LAUNCH_CHECK();
    }
    return 0;
  }
  if(N >= 16 && N <= 16384)
  {
    const int T = N / 16, threads = std::max(256, T), per_cta = threads / T;
    const size_t smem = (size_t) per_cta * N * sizeof(float2);
    const float scale = 1.0f / sqrtf((float) N);
    const unsigned grid = (unsigned) ((batch + per_cta - 1) / per_cta);
    if(smem > 48 * 1024 && !p->smem_optin)
    {
      TSD_CUDA(cudaFuncSetAttribute(fft_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
      TSD_CUDA(cudaFuncSetAttribute(fft_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
      p->smem_optin = true;
    }
    KernelTimer timer;
    if(forward) fft_smem_kernel<false><<<grid, threads, smem, r.stream>>>(x, xs, y, ys, N, batch, scale);
    else fft_smem_kernel<true><<<grid, threads, smem, r.stream>>>(x, xs, y, ys, N, batch, scale);
    TSD_LAUNCH_CHECK();
    return 0;
  }
  // ---- 32768, 131072, 262144 (and 65536 when the pipeline cannot serve the buffers): R strided 16384-point transforms + one
  // combine pass.  TSDGPU_FFT_SPLIT=0 keeps the radix-2 passes.
  if(N > 16384 && N <= 16 * 16384 && !(getenv("TSDGPU_FFT_SPLIT") && atoi(getenv("TSDGPU_FFT_SPLIT")) == 0))
  {
    const int M = 16384, R = N / M;
    if(ensure_work(p, 0)) return 1;
    if(!p->smem_optin)
    {
      TSD_CUDA(cudaFuncSetAttribute(fft_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
      TSD_CUDA(cudaFuncSetAttribute(fft_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
      p->smem_optin = true;
    }
    const long long subs = (long long) batch * R;
    const float s1 = 1.0f / sqrtf((float) M), s2 = 1.0f / sqrtf((float) R);
    KernelTimer timer;
    if(forward) fft_smem_kernel<false><<<(unsigned) subs, M / 16, (size_t) M * sizeof(float2), r.stream>>>(x, xs, p->work[0], M, M, (int) subs, s1, R);
    else fft_smem_kernel<true><<<(unsigned) subs, M / 16, (size_t) M * sizeof(float2), r.stream>>>(x, xs, p->work[0], M, M, (int) subs, s1, R);
    TSD_LAUNCH_CHECK();
    const int grid = grid_for((long long) M * batch, 256);
#define SPLIT_GO(RR)                                                                                                        \
    if(forward) fft_split_combine_kernel<false, RR><<<grid, 256, 0, r.stream>>>(p->work[0], y, ys, M, batch, s2);             \
    else fft_split_combine_kernel<true, RR><<<grid, 256, 0, r.stream>>>(p->work[0], y, ys, M, batch, s2);
    if(R == 2) { SPLIT_GO(2) }
    else if(R == 4) { SPLIT_GO(4) }
    else if(R == 8) { SPLIT_GO(8) }
    else { SPLIT_GO(16) }
#undef SPLIT_GO
    TSD_LAUNCH_CHECK();
    return 0;
  }
  // ---- 2^19 ... 2^22: two levels, N = 16 x R2 x 16384: strided 16384-point transforms, R2-point combines into sixteen
  // transforms of N / 16 points, 16-point combine: three passes over the data
  if(N > 16 * 16384 && N <= 256 * 16384 && !(getenv("TSDGPU_FFT_SPLIT") && atoi(getenv("TSDGPU_FFT_SPLIT")) == 0))
  {
    const int M = 16384, M2 = N / 16, R2 = M2 / M, RT = 16 * R2;
    if(ensure_work(p, 0) || ensure_work(p, 1)) return 1;
    if(!p->smem_optin)
    {
      TSD_CUDA(cudaFuncSetAttribute(fft_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
      TSD_CUDA(cudaFuncSetAttribute(fft_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
      p->smem_optin = true;
    }
    const long long subs = (long long) batch * RT;
    const float s1 = 1.0f / sqrtf((float) M), s2 = 1.0f / sqrtf((float) R2), s3 = 0.25f;
    KernelTimer timer;
    if(forward) fft_smem_kernel<false><<<(unsigned) subs, M / 16, (size_t) M * sizeof(float2), r.stream>>>(x, xs, p->work[0], M, M, (int) subs, s1, RT);
    else fft_smem_kernel<true><<<(unsigned) subs, M / 16, (size_t) M * sizeof(float2), r.stream>>>(x, xs, p->work[0], M, M, (int) subs, s1, RT);
    TSD_LAUNCH_CHECK();
    const int g2 = grid_for((long long) M * batch * 16, 256);
#define SPLIT_GO2(RR)                                                                                                               \
    if(forward) fft_split_combine_kernel<false, RR><<<g2, 256, 0, r.stream>>>(p->work[0], p->work[1], M2, M, batch * 16, s2, 16);       \
    else fft_split_combine_kernel<true, RR><<<g2, 256, 0, r.stream>>>(p->work[0], p->work[1], M2, M, batch * 16, s2, 16);
    if(R2 == 2) { SPLIT_GO2(2) }
    else if(R2 == 4) { SPLIT_GO2(4) }
    else if(R2 == 8) { SPLIT_GO2(8) }
    else { SPLIT_GO2(16) }
#undef SPLIT_GO2
    TSD_LAUNCH_CHECK();
    const int g3 = grid_for((long long) M2 * batch, 256);
    if(forward) fft_split_combine_kernel<false, 16><<<g3, 256, 0, r.stream>>>(p->work[1], y, ys, M2, batch, s3);
    else fft_split_combine_kernel<true, 16><<<g3, 256, 0, r.stream>>>(p->work[1], y, ys, M2, batch, s3);
    TSD_LAUNCH_CHECK();
    return 0;
  }
  // ---- generic: log2(N) radix-2 passes, ping-pong so that the last pass lands in y
  int L = 0;
  while((1 << L) < N) L++;
  if(ensure_work(p, 0)) return 1;
  const float2 *in = x;
  long long in_stride = xs;
  if(x == y)
  {
    // a pass cannot run in place: read from a private copy
    if(ensure_work(p, 1)) return 1;
    copy_strided<<<grid_for((long long) N * batch, 256), 256, 0, r.stream>>>(x, xs, p->work[1], N, N, batch);
    TSD_LAUNCH_CHECK();
    in = p->work[1];
    in_stride = N;
  }
  const float scale = 1.0f / sqrtf((float) N);
  int pass = 0;
  for(int n = 1; n < N; n *= 2, pass++)
  {
    const bool last = (pass == L - 1);
    // pass i writes y when (L-1-i) is even, else the work buffer
    const bool to_y = ((L - 1 - pass) & 1) == 0;
    float2 *out;
    long long out_stride;
    if(to_y) { out = y; out_stride = ys; }
    else { out = p->work[0]; out_stride = N; }
    const int grid = grid_for((long long) (N / 2) * batch, 256);
    if(forward) fft_radix2_pass<false><<<grid, 256, 0, r.stream>>>(in, in_stride, out, out_stride, N, n, batch, last ? scale : 1.0f);
    else fft_radix2_pass<true><<<grid, 256, 0, r.stream>>>(in, in_stride, out, out_stride, N, n, batch, last ? scale : 1.0f);
    TSD_LAUNCH_CHECK();
    in = out;
    in_stride = out_stride;
  }
  return 0;
}

} // namespace tsdgpu

extern "C" {

int tsdgpu_fft_plan(int n, int batch, tsdgpu_fft_t *out)
{
  TSD_ENTER(-1);
  if(!out) return fail("tsdgpu_fft_plan: null argument");
  return fft_plan_create(n, batch, out);
}

int tsdgpu_fft_exec(tsdgpu_fft_t p, const void *x, long long xs, void *y, long long ys, int forward, int mem)
{
  TSD_ENTER(p ? p->device : -1);
  if(!p || !x || !y) return fail("tsdgpu_fft_exec: null argument");
  if(xs < p->n || ys < p->n) return fail("tsdgpu_fft_exec: stride smaller than n");
  if(mem == TSDGPU_DEVICE) return fft_exec_device(p, (const float2 *) x, xs, (float2 *) y, ys, forward != 0);
  // host memory: groups of transforms pipelined through the persistent device staging buffers (H2D of group k+1,
  // transforms of group k and D2H of group k-1 overlap on three streams)
  const size_t row = (size_t) p->n * sizeof(float2);
  // plans with sub-plans (n not a power of two) are sized for the whole batch: one group
  long long group = p->sub ? p->batch : std::max<long long>(1, std::min<long long>(p->batch, (long long) ((32ull << 20) / row)));
  if(host_stage_reserve(row * group, row * group)) return 1;
  HostStage &hs = host_stage();
  const float2 *xh = (const float2 *) x;
  float2 *yh = (float2 *) y;
  const int full_batch = p->batch;
  long long done = 0;
  const int rc = host_pipeline(
    full_batch, group,
    [&](int slot, long long first, long long count) -> int {
      if(stage_in(slot, hs.in[slot], row, xh + first * xs, (size_t) xs * 8, row, (size_t) count)) return 1;
      return 0;
    },
    [&](long long count) { return count; },
    [&](int slot, long long count, long long *got) -> int {
      p->batch = (int) count;   // the kernels take the number of transforms from the plan
      const int r2 = fft_exec_device(p, (const float2 *) hs.in[slot], p->n, (float2 *) hs.out[slot], p->n, forward != 0);
      p->batch = full_batch;
      *got = count;
      return r2;
    },
    [&](int slot, long long out_first, long long count) -> int {
      if(stage_out(slot, yh + out_first * ys, (size_t) ys * 8, hs.out[slot], row, row, (size_t) count)) return 1;
      return 0;
    },
    &done);
  p->batch = full_batch;
  return rc;
}

int tsdgpu_fft_destroy(tsdgpu_fft_t p)
{
  if(!p) return 0;
  TSD_ENTER(p->device);
  cudaStreamSynchronize(rt().stream);
  fft_plan_destroy(p);
  return 0;
}

} // extern "C"
