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
#include "common.cuh"
#include "tsdgpu.h"

#include <cmath>
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

// v1: one thread per output, window and LUT column read through L1
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

// new_hist = last hist_len samples of (old_hist ++ x[0..n))
__global__ void resamp_hist_kernel(const float2 *x, long long x_stride, int n, const float2 *o, float2 *d, int hl)
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
  float ratio = 1, increment = 1, phase = 0;
  int K = 0, nphases = 0, nchan = 0, hist_len = 0;
  float *d_lut = nullptr;
  float2 *d_hist[2] = {nullptr, nullptr};
  int cur = 0;
  // rotating pinned + device schedule buffers
  static constexpr int NBUF = 3;
  int2 *h_sched[NBUF] = {nullptr, nullptr, nullptr};
  int2 *d_sched[NBUF] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev[NBUF] = {nullptr, nullptr, nullptr};
  size_t sched_cap = 0;
  int next_buf = 0;
};

// The reference recurrence (ra.cc:58-73), verbatim in float32.  Processes inputs [i0, i1) of the
// call, appends (in_idx, lut_idx) pairs, returns the updated phase.
static float resamp_schedule(float phase, float increment, int nphases, int i0, int i1, int2 *out, size_t cap,
                             size_t *count, bool *overflow)
{
  size_t j = 0;
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

static int resamp_run_device(tsdgpu_resamp_s *f, const float2 *x, long long xs, int n, float2 *y, long long ys,
                             long long ycap, long long *n_out)
{
  *n_out = 0;
  if(n <= 0) return 0;
  Runtime &r = rt();
  float2 *hist_old = f->d_hist[f->cur], *hist_new = f->d_hist[f->cur ^ 1];
  if(f->hist_len > 0)
  {
    dim3 grid((f->hist_len + 255) / 256, f->nchan);
    resamp_hist_kernel<<<grid, 256, 0, r.stream>>>(x, xs, n, hist_old, hist_new, f->hist_len);
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
    phase = resamp_schedule(phase, f->increment, f->nphases, i0, i1, f->h_sched[b], f->sched_cap, &cnt, &ovf);
    if(ovf) return fail("tsdgpu_resamp_step: schedule overflow");
    if(cnt == 0) continue;
    if(produced + (long long) cnt > ycap) return fail("tsdgpu_resamp_step: output capacity too small");
    TSD_CUDA(cudaMemcpyAsync(f->d_sched[b], f->h_sched[b], cnt * sizeof(int2), cudaMemcpyHostToDevice, r.stream));
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
    dim3 grid((unsigned) ((cnt + RS_NT - 1) / RS_NT), f->nchan);
    {
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

int tsdgpu_resamp_create(float ratio, const float *lut, int K, int nphases, int nchan, tsdgpu_resamp_t *out)
{
  if(ensure_init()) return 1;
  if(!out || !lut) return fail("tsdgpu_resamp_create: null argument");
  if(K <= 0 || nphases <= 0) return fail("tsdgpu_resamp_create: K and nphases must be > 0");
  if(nchan <= 0 || nchan > 65535) return fail("tsdgpu_resamp_create: nchan must be in [1, 65535]");
  if(!(ratio > 0) || std::isinf(ratio)) return fail("tsdgpu_resamp_create: invalid ratio");
  auto *f = new tsdgpu_resamp_s;
  f->ratio = ratio;
  f->increment = 1 / ratio;     // ra.cc:28
  f->phase = 0;
  f->K = K;
  f->nphases = nphases;
  f->nchan = nchan;
  f->hist_len = K - 1;
  const size_t lut_n = (size_t) K * (nphases + 1);
  TSD_CUDA(cudaMalloc(&f->d_lut, lut_n * sizeof(float)));
  TSD_CUDA(cudaMemcpyAsync(f->d_lut, lut, lut_n * sizeof(float), cudaMemcpyHostToDevice, rt().stream));
  for(int i = 0; i < 2; i++)
  {
    size_t bytes = std::max<size_t>(1, (size_t) nchan * f->hist_len) * sizeof(float2);
    TSD_CUDA(cudaMalloc(&f->d_hist[i], bytes));
    TSD_CUDA(cudaMemsetAsync(f->d_hist[i], 0, bytes, rt().stream));
  }
  TSD_CUDA(cudaStreamSynchronize(rt().stream));
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
  if(ensure_init()) return 1;
  if(!f || !n_out) return fail("tsdgpu_resamp_step: null argument");
  *n_out = 0;
  if(n < 0) return fail("tsdgpu_resamp_step: n < 0");
  if(n == 0) return 0;            // ra.cc:45-49
  if(!x || !y) return fail("tsdgpu_resamp_step: null buffer");
  if(xs < n) return fail("tsdgpu_resamp_step: channel stride smaller than n");
  if(mem == TSDGPU_DEVICE) return resamp_run_device(f, (const float2 *) x, xs, n, (float2 *) y, ys, ycap, n_out);
  const long long cnt = tsdgpu_resamp_out_count(f, n);
  if(cnt > ycap) return fail("tsdgpu_resamp_step: output capacity too small");
  float2 *dx = nullptr, *dy = nullptr;
  TSD_CUDA(cudaMalloc(&dx, (size_t) f->nchan * n * sizeof(float2)));
  if(cudaMalloc(&dy, std::max<size_t>(1, (size_t) f->nchan * cnt) * sizeof(float2)) != cudaSuccess)
  {
    cudaFree(dx);
    return fail("tsdgpu_resamp_step: out of device memory");
  }
  int rc = 0;
  cudaError_t e = cudaMemcpy2DAsync(dx, (size_t) n * 8, x, (size_t) xs * 8, (size_t) n * 8, f->nchan,
                                    cudaMemcpyHostToDevice, rt().stream);
  if(e == cudaSuccess) rc = resamp_run_device(f, dx, n, n, dy, cnt, cnt, n_out);
  if(e == cudaSuccess && !rc && cnt > 0)
    e = cudaMemcpy2DAsync(y, (size_t) ys * 8, dy, (size_t) cnt * 8, (size_t) cnt * 8, f->nchan, cudaMemcpyDeviceToHost,
                          rt().stream);
  if(e == cudaSuccess) e = cudaStreamSynchronize(rt().stream);
  cudaFree(dx);
  cudaFree(dy);
  if(e != cudaSuccess) return fail(std::string("tsdgpu_resamp_step: ") + cudaGetErrorString(e));
  return rc;
}

int tsdgpu_resamp_schedule(float *phase, float ratio, int nphases, int n, int32_t *in_idx, int32_t *lut_idx,
                           long long capacity, long long *n_out)
{
  if(!phase || !n_out) return fail("tsdgpu_resamp_schedule: null argument");
  if(!(ratio > 0) || nphases <= 0 || n < 0) return fail("tsdgpu_resamp_schedule: invalid argument");
  const float inc = 1 / ratio;
  float ph = *phase;
  long long j = 0;
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

int tsdgpu_resamp_destroy(tsdgpu_resamp_t f)
{
  if(!f) return 0;
  cudaStreamSynchronize(rt().stream);
  cudaFree(f->d_lut);
  cudaFree(f->d_hist[0]);
  cudaFree(f->d_hist[1]);
  for(int b = 0; b < tsdgpu_resamp_s::NBUF; b++)
  {
    if(f->h_sched[b]) cudaFreeHost(f->h_sched[b]);
    if(f->d_sched[b]) cudaFree(f->d_sched[b]);
    if(f->ev[b]) cudaEventDestroy(f->ev[b]);
  }
  delete f;
  return 0;
}

} // extern "C"
