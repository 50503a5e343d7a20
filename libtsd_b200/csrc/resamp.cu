// Arbitrary-ratio LUT resampler on sm_100a.  Replaces AdaptationRythmeSimple::step
// (reference ra.cc:39-77) + InterpolateurRIF::step (filtrage.hpp:1873-1881) + the LUT lookup of
// InterpolateurSinc::coefs (itrp.cc:16-22).
//
// The reference advances a float32 phase per input sample (ra.cc:64-73); that recurrence does
// not depend on the data, so it is run ONCE on the host for all channels, bit for bit as the
// reference writes it, and yields for every output j the pair
//   (in_idx[j] = index of the newest input in the window, lut_idx[j] = (int)(phase*nphases)).
// The device then evaluates out[c][j] = sum_{i<K} lut[lut_idx[j]][i] * x[c][in_idx[j]-(K-1)+i]
// with i ascending, as InterpolateurRIF::step does.  State across calls: phase (host) and the
// last K-1 inputs of every channel (device).
// Kernels: resamp_tc.cu (tcgen05 3xTF32 banded filter-bank GEMM, default whenever resamp_tc_eligible() accepts the
// chunk's schedule and alignment), resamp_banded_kernel (FP32 FMA, 8 outputs x 64 channels per warp) and
// resamp_lut_kernel (one thread per output, any window) below.
#include "common.cuh"
#include "host_pipe.cuh"
#include "resamp_tc.h"
#include "tsdgpu.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace tsdgpu {

struct ResampParams
{
  const float2 *x;
  float2 *y;
  const float2 *hist;    // [nchan][K-1]
  const float *lut;      // [(nphases+1)][K]
  const int2 *sched;     // per output of this chunk: {in_idx (call-relative), lut_idx}
  long long x_stride, y_stride;
  long long out0;        // first output (call-relative) of this chunk
  int n_out_chunk, K, hist_len;
};

constexpr int RS_NT = 128;

// generic fallback: one thread per output, window and LUT column read through L1
__global__ void __launch_bounds__(RS_NT) resamp_lut_kernel(ResampParams p)
{
  const int j = blockIdx.x * RS_NT + threadIdx.x;
  if(j >= p.n_out_chunk) return;
  const int chan = blockIdx.y;
  const int2 sc = p.sched[j];
  const float2 *x = p.x + (long long) chan * p.x_stride;
  const float2 *hist = p.hist + (long long) chan * p.hist_len;
  const float *h = p.lut + (size_t) sc.y * p.K;
  const int first = sc.x - (p.K - 1);   // call-relative index of the oldest sample in the window
  float sr = 0.f, si = 0.f;
  if(first >= 0)
  {
    const float2 *w = x + first;
    for(int i = 0; i < p.K; i++)
    {
      const float2 v = __ldg(w + i);
      const float c = __ldg(h + i);
      sr = fmaf(v.x, c, sr);
      si = fmaf(v.y, c, si);
    }
  }
  else
  {
    for(int i = 0; i < p.K; i++)
    {
      const int idx = first + i;
      const float2 v = (idx >= 0) ? __ldg(x + idx) : hist[p.hist_len + idx];
      const float c = __ldg(h + i);
      sr = fmaf(v.x, c, sr);
      si = fmaf(v.y, c, si);
    }
  }
  p.y[(long long) chan * p.y_stride + p.out0 + j] = make_float2(sr, si);
}

// ---- generic kernel for everything the fast kernels do not take: real-valued data (filtre_itrp<float>, ra.cc:190-195)
// and the interpolators whose coefficients are evaluated at the exact delay instead of a LUT column (itrp.cc:82-127).
// One thread per output; MODE 0: LUT column sched.y; MODE 1: InterpolateurLineaire {1 - tau, tau}; MODE 2:
// InterpolateurLagrange of degree d = K - 1, t = (d - 1)/2 + tau, h_j = prod_{k != j} (t - k) / (j - k) with the reference's
// float32 operations in the reference's order; tau = the float32 phase of the output (bits in sched.y).
__device__ __forceinline__ float2 rs_load(const float2 *p) { return __ldg(p); }
__device__ __forceinline__ float rs_load(const float *p) { return __ldg(p); }
__device__ __forceinline__ void rs_acc(float2 &s, float2 v, float c) { s.x = fmaf(v.x, c, s.x); s.y = fmaf(v.y, c, s.y); }
__device__ __forceinline__ void rs_acc(float &s, float v, float c) { s = fmaf(v, c, s); }
__device__ __forceinline__ void rs_zero(float2 &s) { s = make_float2(0.f, 0.f); }
__device__ __forceinline__ void rs_zero(float &s) { s = 0.f; }

template<typename T> struct ResampGenParams
{
  const T *x;
  T *y;
  const T *hist;         // [nchan][K-1]
  const float *lut;      // MODE 0
  const int2 *sched;
  long long x_stride, y_stride, out0;
  int n_out_chunk, K, hist_len;
};

template<typename T, int MODE> __global__ void __launch_bounds__(RS_NT) resamp_gen_kernel(ResampGenParams<T> p)
{
  const int j = blockIdx.x * RS_NT + threadIdx.x;
  if(j >= p.n_out_chunk) return;
  const int chan = blockIdx.y;
  const int2 sc = p.sched[j];
  const T *x = p.x + (long long) chan * p.x_stride;
  const T *hist = p.hist + (long long) chan * p.hist_len;
  const int first = sc.x - (p.K - 1);   // call-relative index of the oldest sample in the window
  const float tau = __int_as_float(sc.y);
  const int d = p.K - 1;
  const float t = ((float) d - 1.0f) / 2 + tau;   // itrp.cc:118
  T acc;
  rs_zero(acc);
  for(int i = 0; i < p.K; i++)
  {
    float c;
    if(MODE == 0) c = __ldg(p.lut + (size_t) sc.y * p.K + i);
    else if(MODE == 1) c = (i == 0) ? 1 - tau : tau;            // itrp.cc:86
    else
    {
      c = 1.0f;
      for(int k = 0; k <= d; k++)
        if(k != i) c *= (t - k) / (i - k);                       // itrp.cc:120-126
    }
    const int idx = first + i;
    const T v = (idx >= 0) ? rs_load(x + idx) : hist[p.hist_len + idx];
    rs_acc(acc, v, c);
  }
  p.y[(long long) chan * p.y_stride + p.out0 + j] = acc;
}

// ---- main kernel: channel-batched banded product ----------------------------------------------
// All channels share the schedule, so a tile of 16 consecutive outputs is a small banded matrix
//   A[s][jj] = lut[lut_idx[j]][s - d_jj]  (0 <= s - d_jj < K, else 0),  d_jj = window start of output jj
// applied to every channel: out[c][j] = sum_s A[s][jj] * x[c][b + s].  One warp owns 8 outputs x 64
// channels (lane = channel, 2 channels per lane): per window sample it reads 8 coefficients with
// two broadcast LDS.128 and its two samples with two LDS.64, and issues 32 FFMA -> FP32-FMA bound.
// A CTA = 8 warps = 64 consecutive outputs x 64 channels; the input window is staged once in shared
// memory as [sample][channel] (row pitch 65 float2: conflict-free transposing stores).
// s ascending == tap index ascending, i.e. the reference's accumulation order
// (filtrage.hpp:1877-1879); the zero coefficients outside the band only add +0.
constexpr int RS2_RJ = 8, RS2_WARPS = 8, RS2_CH = 64, RS2_J = RS2_RJ * RS2_WARPS, RS2_PITCH = RS2_CH + 1;

struct Resamp2Params
{
  ResampParams b;
  int nchan, s_cta_max, s_w_max;
};

__global__ void __launch_bounds__(RS2_WARPS * 32) resamp_banded_kernel(Resamp2Params q)
{
  const ResampParams &p = q.b;
  extern __shared__ __align__(16) unsigned char rs_smem[];
  float2 *tile = reinterpret_cast<float2 *>(rs_smem);                                        // [s_cta_max][65]
  float *Aall = reinterpret_cast<float *>(rs_smem + (((size_t) q.s_cta_max * RS2_PITCH * sizeof(float2) + 15) & ~(size_t) 15));   // [4][s_w_max][16]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j0 = blockIdx.x * RS2_J;                       // first output of this CTA within the chunk
  const int c0 = blockIdx.y * RS2_CH;                      // first channel
  const int jlast = min(j0 + RS2_J, p.n_out_chunk) - 1;
  const int K = p.K;
  const int b_first = p.sched[j0].x - (K - 1);             // oldest input needed by the CTA
  const int s_cta = p.sched[jlast].x - b_first + 1;

  // ---- stage the window asynchronously: tile[s][c] = stream_c[b_first + s]   (LDGSTS, transposing)
  for(int c = warp; c < RS2_CH; c += RS2_WARPS)
  {
    const int chan = c0 + c;
    const bool ok = chan < q.nchan;
    const float2 *x = p.x + (long long) chan * p.x_stride;
    const float2 *hist = p.hist + (long long) chan * p.hist_len + p.hist_len;
    for(int sidx = lane; sidx < s_cta; sidx += 32)
    {
      const int t = b_first + sidx;
      float2 *dst = tile + sidx * RS2_PITCH + c;
      if(ok) cp_async8(dst, (t >= 0) ? (x + t) : (hist + t));
      else *dst = make_float2(0.f, 0.f);
    }
  }
  cp_async_commit();
  // ---- meanwhile: this warp's banded coefficient matrix.  Lane l fills output jj = l & 7 for the
  //      window samples s = (l >> 3), (l >> 3) + 4, ... : a strided walk down one LUT column.
  float *A = Aall + (size_t) warp * q.s_w_max * RS2_RJ;
  const int jw0 = j0 + warp * RS2_RJ;
  const bool warp_active = jw0 <= jlast;
  int b_w = 0, s_w = 0;
  if(warp_active)
  {
    const int jw_last = min(jw0 + RS2_RJ - 1, jlast);
    b_w = p.sched[jw0].x - (K - 1);
    s_w = p.sched[jw_last].x - b_w + 1;
    const int jj = lane & (RS2_RJ - 1);
    const bool have = jw0 + jj <= jlast;
    const int2 sc = have ? p.sched[jw0 + jj] : make_int2(0, 0);
    const int d = sc.x - (K - 1) - b_w;
    const float *colp = p.lut + (size_t) sc.y * K;
#pragma unroll 8
    for(int sidx = lane / RS2_RJ; sidx < s_w; sidx += 32 / RS2_RJ)
    {
      const int i = sidx - d;
      float a = 0.f;
      if(have && i >= 0 && i < K) a = __ldg(colp + i);
      A[sidx * RS2_RJ + jj] = a;
    }
  }
  cp_async_wait_all();
  __syncthreads();
  if(!warp_active) return;

  // ---- 16 outputs x 2 channels per lane
  float2 acc0[RS2_RJ], acc1[RS2_RJ];
#pragma unroll
  for(int jj = 0; jj < RS2_RJ; jj++) acc0[jj] = acc1[jj] = make_float2(0.f, 0.f);
  const float2 *col = tile + (size_t) (b_w - b_first) * RS2_PITCH + lane;
  for(int sidx = 0; sidx < s_w; sidx++)
  {
    const float4 *a4 = reinterpret_cast<const float4 *>(A + sidx * RS2_RJ);
    const float2 x0 = col[sidx * RS2_PITCH], x1 = col[sidx * RS2_PITCH + 32];
    float cf[RS2_RJ];
#pragma unroll
    for(int k = 0; k < RS2_RJ / 4; k++)
    {
      const float4 t = a4[k];
      cf[4 * k] = t.x; cf[4 * k + 1] = t.y; cf[4 * k + 2] = t.z; cf[4 * k + 3] = t.w;
    }
#pragma unroll
    for(int jj = 0; jj < RS2_RJ; jj++)
    {
      acc0[jj].x = fmaf(x0.x, cf[jj], acc0[jj].x);
      acc0[jj].y = fmaf(x0.y, cf[jj], acc0[jj].y);
      acc1[jj].x = fmaf(x1.x, cf[jj], acc1[jj].x);
      acc1[jj].y = fmaf(x1.y, cf[jj], acc1[jj].y);
    }
  }
  // ---- transpose through the warp's (now dead) coefficient area so that 8 lanes store the 8
  //      consecutive outputs (64 B) of one channel
  __syncwarp();
  float2 *patch = reinterpret_cast<float2 *>(A);   // [32 channels][RS2_RJ + 1]
  const int nout_w = min(RS2_RJ, jlast - jw0 + 1);
#pragma unroll
  for(int half = 0; half < 2; half++)
  {
#pragma unroll
    for(int jj = 0; jj < RS2_RJ; jj++) patch[lane * (RS2_RJ + 1) + jj] = half ? acc1[jj] : acc0[jj];
    __syncwarp();
    const int jj = lane & (RS2_RJ - 1);
    for(int cc = lane / RS2_RJ; cc < 32; cc += 32 / RS2_RJ)
    {
      const int chan = c0 + half * 32 + cc;
      if(chan < q.nchan && jj < nout_w) p.y[(long long) chan * p.y_stride + p.out0 + jw0 + jj] = patch[cc * (RS2_RJ + 1) + jj];
    }
    __syncwarp();
  }
}

// new_hist = last hist_len samples of (old_hist ++ x[0..n))
template<typename T> __global__ void resamp_hist_kernel(const T *x, long long x_stride, int n, const T *o, T *d, int hl)
{
  const int chan = blockIdx.y;
  for(int j = blockIdx.x * blockDim.x + threadIdx.x; j < hl; j += gridDim.x * blockDim.x)
  {
    int pos = n - hl + j;
    d[(long long) chan * hl + j] = (pos >= 0) ? x[(long long) chan * x_stride + pos] : o[(long long) chan * hl + hl + pos];
  }
}

} // namespace tsdgpu

using namespace tsdgpu;

struct tsdgpu_resamp_s
{
  int device = 0;              // CUDA device the object lives on
  float ratio = 1, increment = 1, phase = 0;
  int K = 0, nphases = 0, nchan = 0, hist_len = 0;
  int dc = 2;                  // floats per sample: 1 = real data (filtre_itrp<float>), 2 = cfloat
  int mode = 0;                // 0 LUT, 1 linear, 2 Lagrange (coefficients at the exact phase)
  float *d_lut = nullptr;
  float2 *d_hist[2] = {nullptr, nullptr};
  int cur = 0;
  // rotating pinned + device schedule buffers
  static constexpr int NBUF = 3;
  int2 *h_sched[NBUF] = {nullptr, nullptr, nullptr};
  int2 *d_sched[NBUF] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev[NBUF] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_sched[NBUF] = {nullptr, nullptr, nullptr};
  cudaStream_t sched_stream = nullptr;   // schedule uploads overlap the previous chunk's kernel
  size_t sched_cap = 0;
  int next_buf = 0;
  size_t smem_set = 0;
  // real-valued data with a LUT interpolator (filtre_itrp<float>, filtre_reechan<float>): channels 2p, 2p+1 ride a complex
  // object as the real and imaginary part of channel p (exact: the LUT is real); this object then only packs / unpacks, the
  // streaming state (history, phase) lives in `pair`
  tsdgpu_resamp_s *pair = nullptr;
  float2 *pair_x = nullptr, *pair_y = nullptr;
  size_t pair_x_cap = 0, pair_y_cap = 0;   // samples per row
};

// real rows 2p, 2p+1 -> one complex row p (a missing odd partner reads as zero); and back
__global__ void resamp_pair_pack_kernel(const float *x, long long xs, long long n, int nchan, float2 *xp, long long xps)
{
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  if(i >= n) return;
  const float a = x[(long long) (2 * p) * xs + i], b = (2 * p + 1 < nchan) ? x[(long long) (2 * p + 1) * xs + i] : 0.f;
  xp[(long long) p * xps + i] = make_float2(a, b);
}
__global__ void resamp_pair_unpack_kernel(const float2 *yp, long long yps, long long n, int nchan, float *y, long long ys)
{
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  if(i >= n) return;
  const float2 v = yp[(long long) p * yps + i];
  y[(long long) (2 * p) * ys + i] = v.x;
  if(2 * p + 1 < nchan) y[(long long) (2 * p + 1) * ys + i] = v.y;
}

// Integer-exact form of the recurrence for 1 < increment <= 1.125 (ratio in [8/9, 1)), in units of 2^-23 (P = phase *
// 2^23).  In this range every value of the float32 chain of ra.cc:64-73 is a multiple of 2^-23 below 4, so:
//   * phase - 1 is exact; phase + increment (phase < 1) lies in [1, 3): exact below 2, rounded to the grid 2^-22 with
//     ties to even from 2 on, i.e. S = (S + ((S >> 1) & 1)) & ~1;
//   * with d = increment - 1 the phase climbs by exactly d per input (P0, P0 + d, ... all below 1: m consecutive inputs
//     emit one output each), the m-th addition crosses 2 (the only rounding of the run), the next input emits nothing.
// A run of m >= 8 outputs then costs one multiplication for m instead of 2 m dependent float additions, and its outputs
// are independent of each other: 1.5 x faster than the float loop on the host (profiles/microbench/sched_bench.cc).
// The LUT index keeps the reference's own float operations, (int)(phase * nphases) with phase = P * 2^-23 (exact).
// Returns false (nothing written) when the preconditions do not hold; the caller then runs the float32 loop.
static bool resamp_schedule_runs(float *phase_io, float increment, int nphases, int i0, int i1, int2 *out, size_t *count)
{
  const float ph = *phase_io;
  if(!(increment > 1.0f && increment <= 1.125f) || !(ph >= 0.0f && ph < 2.0f)) return false;
  const float t = ph * 8388608.0f;                   // exact scaling
  if(t != std::floor(t)) return false;
  const uint32_t ONE = 1u << 23;
  uint32_t P = (uint32_t) t;
  const uint32_t I = (uint32_t) (increment * 8388608.0f), d = I - ONE;   // exact: floats in [1, 2) are multiples of 2^-23
  const double inv_d = 1.0 / (double) d;
  const float sc = 0x1p-23f, nph = (float) nphases;
  size_t j = 0;
  int i = i0;
  if(i < i1 && P >= ONE) { P -= ONE; i++; }          // carried phase in [1, 2): this input emits nothing
  while(i < i1)
  {
    // P < 1 here.  m = smallest count with P + m d >= 1 (estimate by the reciprocal, then made exact)
    uint32_t m = (uint32_t) ((double) (ONE - P) * inv_d);
    if(m < 1) m = 1;
    while(P + m * d < ONE) m++;
    while(m > 1 && P + (m - 1) * d >= ONE) m--;
    const uint32_t k_emit = std::min<uint32_t>(m, (uint32_t) (i1 - i));
    int2 *o = out + j;
    for(uint32_t k = 0; k < k_emit; k++)
    {
      o[k].x = i + (int) k;
      o[k].y = (int) ((float) (P + k * d) * sc * nph);
    }
    j += k_emit;
    if(k_emit < m) { P += k_emit * d; break; }       // the call ends inside the run: nothing has been rounded yet
    uint32_t S = P + m * d + ONE;                    // last addition of the run: >= 2, float32 grid 2^-22, ties to even
    S = (S + ((S >> 1) & 1u)) & ~1u;
    P = S - ONE;                                     // in [1, 2)
    i += (int) m;
    if(i < i1) { P -= ONE; i++; }                    // the input that emits nothing
  }
  *phase_io = (float) P * sc;                        // P < 2^24: exact
  *count = j;
  return true;
}

// The reference recurrence (ra.cc:58-73), verbatim in float32.  Processes inputs [i0, i1) of the
// call, appends (in_idx, lut_idx) pairs, returns the updated phase.
static float resamp_schedule(float phase, float increment, int nphases, int i0, int i1, int2 *out, size_t cap,
                             size_t *count, bool *overflow, bool exact = false)
{
  size_t j = 0;
  if(exact && out)
  {
    // exact-delay interpolators: the second word carries the float32 phase itself
    for(int i = i0; i < i1; i++)
    {
      while(phase < 1)
      {
        if(j >= cap) { *overflow = true; *count = j; return phase; }
        int bits;
        memcpy(&bits, &phase, sizeof bits);
        out[j] = make_int2(i, bits);
        j++;
        phase += increment;
      }
      phase--;
    }
    *count = j;
    return phase;
  }
  // This loop is the host-side critical path of a step() (one float add + one float subtract of dependent latency per
  // input): keep it free of bounds checks (the caller sizes `out` for ceil((i1 - i0) * max(1, ratio)) + 16 entries,
  // checked once here) and, for increment >= 1 (at most one output per input), free of the inner loop.
  const float nph = (float) nphases;
  if(out && (double) (i1 - i0) / (double) increment + 16.0 <= (double) cap)
  {
    if(!getenv("TSDGPU_RESAMP_SCHED_FLOAT") && resamp_schedule_runs(&phase, increment, nphases, i0, i1, out, count)) return phase;
    if(increment >= 1.0f)
    {
      for(int i = i0; i < i1; i++)
      {
        if(phase < 1)
        {
          out[j] = make_int2(i, (int) (phase * nph));
          j++;
          phase += increment;
        }
        phase--;
      }
    }
    else
    {
      for(int i = i0; i < i1; i++)
      {
        while(phase < 1)
        {
          out[j] = make_int2(i, (int) (phase * nph));
          j++;
          phase += increment;
        }
        phase--;
      }
    }
    *count = j;
    return phase;
  }
  for(int i = i0; i < i1; i++)
  {
    while(phase < 1)
    {
      if(out)
      {
        if(j >= cap) { *overflow = true; *count = j; return phase; }
        out[j].x = i;
        out[j].y = (int) (phase * nphases);
      }
      j++;
      phase += increment;
    }
    phase--;
  }
  *count = j;
  return phase;
}

template<typename T> static int resamp_launch_gen(tsdgpu_resamp_s *f, const void *x, long long xs, void *y, long long ys,
                                                  const void *hist, const int2 *sched, long long out0, size_t cnt)
{
  ResampGenParams<T> g;
  g.x = (const T *) x;
  g.y = (T *) y;
  g.hist = (const T *) hist;
  g.lut = f->d_lut;
  g.sched = sched;
  g.x_stride = xs;
  g.y_stride = ys;
  g.out0 = out0;
  g.n_out_chunk = (int) cnt;
  g.K = f->K;
  g.hist_len = f->hist_len;
  dim3 grid((unsigned) ((cnt + RS_NT - 1) / RS_NT), f->nchan);
  KernelTimer timer;
  if(f->mode == 0) resamp_gen_kernel<T, 0><<<grid, RS_NT, 0, rt().stream>>>(g);
  else if(f->mode == 1) resamp_gen_kernel<T, 1><<<grid, RS_NT, 0, rt().stream>>>(g);
  else resamp_gen_kernel<T, 2><<<grid, RS_NT, 0, rt().stream>>>(g);
  TSD_LAUNCH_CHECK();
  return 0;
}

static int resamp_run_device(tsdgpu_resamp_s *f, const float2 *x, long long xs, int n, float2 *y, long long ys,
                             long long ycap, long long *n_out)
{
  *n_out = 0;
  if(n <= 0) return 0;
  Runtime &r = rt();
  if(f->pair)
  {
    // real-valued data: pack channel pairs, run the complex object (tensor-core kernel when eligible), unpack
    const int pairs = f->pair->nchan;
    const long long cnt = tsdgpu_resamp_out_count(f, n);
    if(cnt > ycap) return fail("tsdgpu_resamp_step: output capacity too small");
    const size_t xcap = ((size_t) n + 1) & ~(size_t) 1, ycap2 = ((size_t) std::max<long long>(cnt, 1) + 1) & ~(size_t) 1;
    if(xcap > f->pair_x_cap || ycap2 > f->pair_y_cap)
    {
      TSD_CUDA(cudaStreamSynchronize(r.stream));
      if(xcap > f->pair_x_cap)
      {
        if(f->pair_x) cudaFree(f->pair_x);
        f->pair_x = nullptr;
        f->pair_x_cap = 0;
        TSD_CUDA(cudaMalloc(&f->pair_x, (size_t) pairs * xcap * sizeof(float2)));
        f->pair_x_cap = xcap;
      }
      if(ycap2 > f->pair_y_cap)
      {
        if(f->pair_y) cudaFree(f->pair_y);
        f->pair_y = nullptr;
        f->pair_y_cap = 0;
        TSD_CUDA(cudaMalloc(&f->pair_y, (size_t) pairs * ycap2 * sizeof(float2)));
        f->pair_y_cap = ycap2;
      }
    }
    resamp_pair_pack_kernel<<<dim3((unsigned) ((n + 255) / 256), pairs), 256, 0, r.stream>>>((const float *) x, xs, n, f->nchan, f->pair_x,
                                                                                               (long long) f->pair_x_cap);
    TSD_LAUNCH_CHECK();
    f->pair->phase = f->phase;
    long long got = 0;
    if(resamp_run_device(f->pair, f->pair_x, (long long) f->pair_x_cap, n, f->pair_y, (long long) f->pair_y_cap, (long long) f->pair_y_cap, &got)) return 1;
    if(got != cnt) return fail("tsdgpu_resamp_step: internal error (paired output count)");
    if(got > 0)
    {
      resamp_pair_unpack_kernel<<<dim3((unsigned) ((got + 255) / 256), pairs), 256, 0, r.stream>>>(f->pair_y, (long long) f->pair_y_cap, got, f->nchan,
                                                                                                     (float *) y, ys);
      TSD_LAUNCH_CHECK();
    }
    f->phase = f->pair->phase;
    *n_out = got;
    return 0;
  }
  float2 *hist_old = f->d_hist[f->cur], *hist_new = f->d_hist[f->cur ^ 1];
  const bool generic = f->dc == 1 || f->mode != 0;   // real data / exact-delay coefficients: one thread per output
  if(f->hist_len > 0)
  {
    dim3 grid((f->hist_len + 255) / 256, f->nchan);
    if(f->dc == 2) resamp_hist_kernel<float2><<<grid, 256, 0, r.stream>>>(x, xs, n, hist_old, hist_new, f->hist_len);
    else
      resamp_hist_kernel<float><<<grid, 256, 0, r.stream>>>((const float *) x, xs, n, (const float *) hist_old, (float *) hist_new, f->hist_len);
    TSD_LAUNCH_CHECK();
  }
  // chunks of inputs: the host computes chunk c+1 while the device works on chunk c
  const int chunk_in = 1 << 20;
  const size_t cap = (size_t) std::ceil((double) chunk_in * std::max(1.0f, f->ratio)) + 16;
  if(cap > f->sched_cap)
  {
    for(int b = 0; b < tsdgpu_resamp_s::NBUF; b++)
    {
      if(f->h_sched[b]) cudaFreeHost(f->h_sched[b]);
      if(f->d_sched[b]) cudaFree(f->d_sched[b]);
      TSD_CUDA(cudaMallocHost(&f->h_sched[b], cap * sizeof(int2)));
      TSD_CUDA(cudaMalloc(&f->d_sched[b], cap * sizeof(int2)));
      if(!f->ev[b]) TSD_CUDA(cudaEventCreateWithFlags(&f->ev[b], cudaEventDisableTiming));
      if(!f->ev_sched[b]) TSD_CUDA(cudaEventCreateWithFlags(&f->ev_sched[b], cudaEventDisableTiming));
    }
    f->sched_cap = cap;
  }
  float phase = f->phase;
  long long produced = 0;
  for(int i0 = 0; i0 < n; i0 += chunk_in)
  {
    const int i1 = std::min(n, i0 + chunk_in);
    const int b = f->next_buf;
    f->next_buf = (b + 1) % tsdgpu_resamp_s::NBUF;
    TSD_CUDA(cudaEventSynchronize(f->ev[b]));   // buffer b no longer in flight
    size_t cnt = 0;
    bool ovf = false;
    phase = resamp_schedule(phase, f->increment, f->nphases, i0, i1, f->h_sched[b], f->sched_cap, &cnt, &ovf, f->mode != 0);
    if(ovf) return fail("tsdgpu_resamp_step: schedule overflow");
    if(cnt == 0) continue;
    if(produced + (long long) cnt > ycap) return fail("tsdgpu_resamp_step: output capacity too small");
    // the schedule travels on its own stream so that it overlaps the previous chunk's kernel
    if(!f->sched_stream) TSD_CUDA(cudaStreamCreateWithFlags(&f->sched_stream, cudaStreamNonBlocking));
    TSD_CUDA(cudaMemcpyAsync(f->d_sched[b], f->h_sched[b], cnt * sizeof(int2), cudaMemcpyHostToDevice, f->sched_stream));
    TSD_CUDA(cudaEventRecord(f->ev_sched[b], f->sched_stream));
    TSD_CUDA(cudaStreamWaitEvent(r.stream, f->ev_sched[b], 0));
    if(generic)
    {
      const int rc = f->dc == 2 ? resamp_launch_gen<float2>(f, x, xs, y, ys, hist_old, f->d_sched[b], produced, cnt)
                                : resamp_launch_gen<float>(f, x, xs, y, ys, hist_old, f->d_sched[b], produced, cnt);
      if(rc) return rc;
      TSD_CUDA(cudaEventRecord(f->ev[b], r.stream));
      produced += (long long) cnt;
      continue;
    }
    ResampParams p;
    p.x = x;
    p.y = y;
    p.hist = hist_old;
    p.lut = f->d_lut;
    p.sched = f->d_sched[b];
    p.x_stride = xs;
    p.y_stride = ys;
    p.out0 = produced;
    p.n_out_chunk = (int) cnt;
    p.K = f->K;
    p.hist_len = f->hist_len;
    // window extents of the banded kernel's tiles (exact, from the schedule just computed)
    int s_cta_max = 0, s_w_max = 0;
    {
      const int2 *sc = f->h_sched[b];
      for(size_t j = 0; j < cnt; j += RS2_RJ)
      {
        const size_t jl = std::min(cnt - 1, j + RS2_RJ - 1);
        s_w_max = std::max(s_w_max, sc[jl].x - sc[j].x + f->K);
      }
      for(size_t j = 0; j < cnt; j += RS2_J)
      {
        const size_t jl = std::min(cnt - 1, j + RS2_J - 1);
        s_cta_max = std::max(s_cta_max, sc[jl].x - sc[j].x + f->K);
      }
      s_w_max = std::max(s_w_max, 72);   // the output transpose patch (32 x 9 float2 = 2304 B) reuses this area
    }
    const size_t smem2 = (((size_t) s_cta_max * RS2_PITCH * sizeof(float2) + 15) & ~(size_t) 15) + (size_t) RS2_WARPS * s_w_max * RS2_RJ * sizeof(float);
    // banded filter-bank GEMM on the tensor cores (3xTF32, resamp_tc.cu); TSDGPU_RESAMP_TC=0 keeps the FP32 FMA kernels
    const char *tc_env = getenv("TSDGPU_RESAMP_TC");
    const bool tc_on = !(tc_env && atoi(tc_env) == 0);
    int max_tile_chunks = 0;
    if(tc_on && resamp_tc_eligible(f->h_sched[b], (long long) cnt, f->K, f->nphases, x, xs, &max_tile_chunks))
    {
      ResampTcParams t;
      t.max_tile_chunks = max_tile_chunks;
      t.x = x;
      t.y = y;
      t.hist = hist_old;
      t.lut = f->d_lut;
      t.sched = f->d_sched[b];
      t.x_stride = xs;
      t.y_stride = ys;
      t.out0 = produced;
      t.n_out = (long long) cnt;
      t.n = n;
      t.K = f->K;
      t.hist_len = f->hist_len;
      t.nchan = f->nchan;
      t.lut_elems = f->K * (f->nphases + 1);
      KernelTimer timer;
      if(resamp_tc_launch(t)) return 1;
    }
    else if(smem2 <= 200 * 1024 && !getenv("TSDGPU_RESAMP_V1"))
    {
      Resamp2Params q;
      q.b = p;
      q.nchan = f->nchan;
      q.s_cta_max = s_cta_max;
      q.s_w_max = s_w_max;
      size_t &smem_set = r.resamp2_smem_set;   // the attribute is per function and device, not per filter object
      if(smem2 > smem_set)
      {
        TSD_CUDA(cudaFuncSetAttribute(resamp_banded_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem2, 64 * 1024)));
        smem_set = std::max<size_t>(smem2, 64 * 1024);
      }
      dim3 grid((unsigned) ((cnt + RS2_J - 1) / RS2_J), (unsigned) ((f->nchan + RS2_CH - 1) / RS2_CH));
      KernelTimer timer;
      resamp_banded_kernel<<<grid, RS2_WARPS * 32, smem2, r.stream>>>(q);
      TSD_LAUNCH_CHECK();
    }
    else
    {
      // generic fallback (very small ratios: the window of 64 outputs does not fit in shared memory)
      dim3 grid((unsigned) ((cnt + RS_NT - 1) / RS_NT), f->nchan);
      KernelTimer timer;
      resamp_lut_kernel<<<grid, RS_NT, 0, r.stream>>>(p);
      TSD_LAUNCH_CHECK();
    }
    TSD_CUDA(cudaEventRecord(f->ev[b], r.stream));
    produced += (long long) cnt;
  }
  f->phase = phase;
  f->cur ^= 1;
  *n_out = produced;
  return 0;
}

extern "C" {

static int resamp_create(float ratio, const float *lut, int K, int nphases, int mode, int data_complex, int nchan,
                         tsdgpu_resamp_t *out);

int tsdgpu_resamp_create(float ratio, const float *lut, int K, int nphases, int nchan, tsdgpu_resamp_t *out)
{
  return resamp_create(ratio, lut, K, nphases, 0, 1, nchan, out);
}

int tsdgpu_resamp_create_ex(float ratio, const float *lut, int K, int nphases, int data_complex, int nchan, tsdgpu_resamp_t *out)
{
  return resamp_create(ratio, lut, K, nphases, 0, data_complex, nchan, out);
}

int tsdgpu_resamp_create_exact(float ratio, int kind, int degree, int data_complex, int nchan, tsdgpu_resamp_t *out)
{
  if(kind == TSDGPU_ITRP_LINEAIRE) return resamp_create(ratio, nullptr, 2, 1, 1, data_complex, nchan, out);   // K = 2 (itrp.cc:92)
  if(kind == TSDGPU_ITRP_LAGRANGE)
  {
    if(degree < 1 || degree > 31) return fail("tsdgpu_resamp_create_exact: Lagrange degree must be in [1, 31]");
    return resamp_create(ratio, nullptr, degree + 1, 1, 2, data_complex, nchan, out);                          // K = d + 1 (itrp.cc:107)
  }
  return fail("tsdgpu_resamp_create_exact: unknown interpolator kind");
}

static int resamp_create(float ratio, const float *lut, int K, int nphases, int mode, int data_complex, int nchan,
                         tsdgpu_resamp_t *out)
{
  TSD_ENTER(-1);
  if(!out || (mode == 0 && !lut)) return fail("tsdgpu_resamp_create: null argument");
  if(K <= 0 || nphases <= 0) return fail("tsdgpu_resamp_create: K and nphases must be > 0");
  if(nchan <= 0 || nchan > 65535) return fail("tsdgpu_resamp_create: nchan must be in [1, 65535]");
  if(!(ratio > 0) || std::isinf(ratio)) return fail("tsdgpu_resamp_create: invalid ratio");
  auto *f = new tsdgpu_resamp_s;
  f->device = rt().device;
  f->ratio = ratio;
  f->increment = 1 / ratio;     // ra.cc:28
  f->phase = 0;
  f->K = K;
  f->nphases = nphases;
  f->nchan = nchan;
  f->hist_len = K - 1;
  f->mode = mode;
  f->dc = data_complex ? 2 : 1;
  if(mode == 0)
  {
    const size_t lut_n = (size_t) K * (nphases + 1);
    TSD_CUDA(cudaMalloc(&f->d_lut, lut_n * sizeof(float)));
    TSD_CUDA(cudaMemcpyAsync(f->d_lut, lut, lut_n * sizeof(float), cudaMemcpyHostToDevice, rt().stream));
  }
  for(int i = 0; i < 2; i++)
  {
    size_t bytes = std::max<size_t>(1, (size_t) nchan * f->hist_len) * sizeof(float) * f->dc;
    TSD_CUDA(cudaMalloc(&f->d_hist[i], bytes));
    TSD_CUDA(cudaMemsetAsync(f->d_hist[i], 0, bytes, rt().stream));
  }
  TSD_CUDA(cudaStreamSynchronize(rt().stream));
  // real-valued data + LUT: a complex object over channel pairs does the work (TSDGPU_RESAMP_PAIR=0: one thread per output,
  // resamp_gen_kernel<float>, 4 Gsamples/s where the pairs reach > 100)
  if(f->dc == 1 && mode == 0 && !(getenv("TSDGPU_RESAMP_PAIR") && atoi(getenv("TSDGPU_RESAMP_PAIR")) == 0))
  {
    if(resamp_create(ratio, lut, K, nphases, 0, 1, (nchan + 1) / 2, &f->pair))
    {
      tsdgpu_resamp_destroy(f);
      return 1;
    }
  }
  *out = f;
  return 0;
}

long long tsdgpu_resamp_out_count(tsdgpu_resamp_t f, int n)
{
  if(!f || n <= 0) return 0;
  size_t cnt = 0;
  bool ovf = false;
  resamp_schedule(f->phase, f->increment, f->nphases, 0, n, nullptr, 0, &cnt, &ovf);
  return (long long) cnt;
}

float tsdgpu_resamp_phase(tsdgpu_resamp_t f) { return f ? f->phase : 0.f; }

int tsdgpu_resamp_step(tsdgpu_resamp_t f, const void *x, long long xs, int n, void *y, long long ys, long long ycap,
                       long long *n_out, int mem)
{
  TSD_ENTER(f ? f->device : -1);
  if(!f || !n_out) return fail("tsdgpu_resamp_step: null argument");
  *n_out = 0;
  if(n < 0) return fail("tsdgpu_resamp_step: n < 0");
  if(n == 0) return 0;            // ra.cc:45-49
  if(!x) return fail("tsdgpu_resamp_step: null input");
  // a short call may emit nothing (n = 1 with a carried phase >= 1 at ratio < 1): the reference then returns an empty
  // vector (ra.cc:39-77) and the caller's y may be null; phase and history still advance
  if(!y && tsdgpu_resamp_out_count(f, n) > 0) return fail("tsdgpu_resamp_step: null output");
  if(xs < n) return fail("tsdgpu_resamp_step: channel stride smaller than n");
  if(mem == TSDGPU_DEVICE) return resamp_run_device(f, (const float2 *) x, xs, n, (float2 *) y, ys, ycap, n_out);
  const long long cnt = tsdgpu_resamp_out_count(f, n);
  if(cnt > ycap) return fail("tsdgpu_resamp_step: output capacity too small");
  const size_t es = sizeof(float) * f->dc;   // bytes per sample
  const long long chunk = host_chunk_len(f->nchan, es, n, 4);
  const long long out_cap = (((long long) std::ceil((double) chunk * std::max(1.0f, f->ratio)) + 16) + 3) & ~3LL;
  if(host_stage_reserve((size_t) f->nchan * chunk * es, (size_t) f->nchan * out_cap * es)) return 1;
  HostStage &hs = host_stage();
  const char *xh = (const char *) x;
  char *yh = (char *) y;
  return host_pipeline(
    n, chunk,
    [&](int slot, long long first, long long count) -> int {
      if(stage_in(slot, hs.in[slot], (size_t) chunk * es, xh + first * es, (size_t) xs * es, (size_t) count * es, f->nchan)) return 1;
      return 0;
    },
    [&](long long count) { return tsdgpu_resamp_out_count(f, (int) count); },
    [&](int slot, long long count, long long *got) -> int {
      return resamp_run_device(f, (const float2 *) hs.in[slot], chunk, (int) count, (float2 *) hs.out[slot], out_cap, out_cap, got);
    },
    [&](int slot, long long out_first, long long count) -> int {
      if(stage_out(slot, yh + out_first * es, (size_t) ys * es, hs.out[slot], (size_t) out_cap * es, (size_t) count * es, f->nchan)) return 1;
      return 0;
    },
    n_out);
}

int tsdgpu_resamp_schedule(float *phase, float ratio, int nphases, int n, int32_t *in_idx, int32_t *lut_idx,
                           long long capacity, long long *n_out)
{
  if(!phase || !n_out) return fail("tsdgpu_resamp_schedule: null argument");
  if(!(ratio > 0) || nphases <= 0 || n < 0) return fail("tsdgpu_resamp_schedule: invalid argument");
  const float inc = 1 / ratio;
  float ph = *phase;
  long long j = 0;
  if(in_idx && lut_idx && (double) n / (double) inc + 16.0 <= (double) capacity)
  {
    // same code path as step(): interleaved pairs, then split
    std::vector<int2> tmp((size_t) capacity);
    size_t cnt = 0;
    bool ovf = false;
    ph = resamp_schedule(ph, inc, nphases, 0, n, tmp.data(), (size_t) capacity, &cnt, &ovf);
    if(ovf) return fail("tsdgpu_resamp_schedule: capacity too small");
    for(size_t k = 0; k < cnt; k++) { in_idx[k] = tmp[k].x; lut_idx[k] = tmp[k].y; }
    *phase = ph;
    *n_out = (long long) cnt;
    return 0;
  }
  for(int i = 0; i < n; i++)
  {
    while(ph < 1)
    {
      if(in_idx || lut_idx)
      {
        if(j >= capacity) return fail("tsdgpu_resamp_schedule: capacity too small");
        if(in_idx) in_idx[j] = i;
        if(lut_idx) lut_idx[j] = (int) (ph * nphases);
      }
      j++;
      ph += inc;
    }
    ph--;
  }
  *phase = ph;
  *n_out = j;
  return 0;
}

int tsdgpu_resamp_get_state(tsdgpu_resamp_t f, float *phase, void *hist_host)
{
  TSD_ENTER(f ? f->device : -1);
  if(!f) return fail("tsdgpu_resamp_get_state: null handle");
  if(phase) *phase = f->phase;
  if(f->pair && hist_host && f->hist_len > 0)
  {
    // the history lives in the complex object: [pairs][hist_len] (re = channel 2p, im = channel 2p+1) -> real rows
    const int pairs = f->pair->nchan, hl = f->hist_len;
    std::vector<float2> h((size_t) pairs * hl);
    TSD_CUDA(cudaStreamSynchronize(rt().stream));
    TSD_CUDA(cudaMemcpy(h.data(), f->pair->d_hist[f->pair->cur], h.size() * sizeof(float2), cudaMemcpyDeviceToHost));
    float *o = (float *) hist_host;
    for(int c = 0; c < f->nchan; c++)
      for(int i = 0; i < hl; i++) o[(size_t) c * hl + i] = (c & 1) ? h[(size_t) (c >> 1) * hl + i].y : h[(size_t) (c >> 1) * hl + i].x;
    return 0;
  }
  if(hist_host && f->hist_len > 0)
  {
    TSD_CUDA(cudaStreamSynchronize(rt().stream));
    TSD_CUDA(cudaMemcpy(hist_host, f->d_hist[f->cur], (size_t) f->nchan * f->hist_len * sizeof(float) * f->dc, cudaMemcpyDeviceToHost));
  }
  return 0;
}

int tsdgpu_resamp_set_state(tsdgpu_resamp_t f, float phase, const void *hist_host)
{
  TSD_ENTER(f ? f->device : -1);
  if(!f) return fail("tsdgpu_resamp_set_state: null handle");
  if(!(phase >= 0.0f) || !(phase < 1e9f)) return fail("tsdgpu_resamp_set_state: invalid phase");
  TSD_CUDA(cudaStreamSynchronize(rt().stream));
  if(f->pair)
  {
    const int pairs = f->pair->nchan, hl = f->hist_len;
    if(hl > 0)
    {
      if(!hist_host) return fail("tsdgpu_resamp_set_state: null history");
      const float *in = (const float *) hist_host;
      std::vector<float2> h((size_t) pairs * hl, make_float2(0.f, 0.f));
      for(int c = 0; c < f->nchan; c++)
        for(int i = 0; i < hl; i++) ((c & 1) ? h[(size_t) (c >> 1) * hl + i].y : h[(size_t) (c >> 1) * hl + i].x) = in[(size_t) c * hl + i];
      TSD_CUDA(cudaMemcpy(f->pair->d_hist[f->pair->cur], h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
    }
    f->phase = f->pair->phase = phase;
    return 0;
  }
  if(f->hist_len > 0)
  {
    if(!hist_host) return fail("tsdgpu_resamp_set_state: null history");
    TSD_CUDA(cudaMemcpy(f->d_hist[f->cur], hist_host, (size_t) f->nchan * f->hist_len * sizeof(float) * f->dc, cudaMemcpyHostToDevice));
  }
  f->phase = phase;
  return 0;
}

int tsdgpu_resamp_destroy(tsdgpu_resamp_t f)
{
  if(!f) return 0;
  TSD_ENTER(f->device);
  cudaStreamSynchronize(rt().stream);
  cudaFree(f->d_lut);
  cudaFree(f->d_hist[0]);
  cudaFree(f->d_hist[1]);
  for(int b = 0; b < tsdgpu_resamp_s::NBUF; b++)
  {
    if(f->h_sched[b]) cudaFreeHost(f->h_sched[b]);
    if(f->d_sched[b]) cudaFree(f->d_sched[b]);
    if(f->ev[b]) cudaEventDestroy(f->ev[b]);
    if(f->ev_sched[b]) cudaEventDestroy(f->ev_sched[b]);
  }
  if(f->sched_stream) cudaStreamDestroy(f->sched_stream);
  if(f->pair) tsdgpu_resamp_destroy(f->pair);
  if(f->pair_x) cudaFree(f->pair_x);
  if(f->pair_y) cudaFree(f->pair_y);
  delete f;
  return 0;
}

} // extern "C"
