// rt_spectrum(SpectrumConfig) — averaged power spectrum in dB (reference fourier.hpp:909-952, src/fourier/fourier.cc:1162-1343).
//
// Per block of BS samples and channel: the nsubs sub-blocks of Nf = BS / nsubs samples are multiplied by the window
// (normalised to energy Nf, fourier.cc:1203), transformed (unitary plan, fourier.cc:1227-1228), |X|^2 is fft-shifted and
// added to the running sum mag_moy — in place for plain averaging, at offset i * sweep.step and times the edge / centre
// mask in sweep mode (fourier.cc:1246-1266).  The block that completes `nmeans` blocks returns
// 10 log10(mag_moy / (nmeans nsubs Nf) [/ mag_cnt] + FLT_MIN) and clears the sum (fourier.cc:1269-1275); the others return an
// empty vector.  Device work: one windowing kernel, the batched FFT plan of fft.cu (any Nf), one accumulation kernel (a thread
// per output bin walks the sub-blocks in the reference's order: no atomics, bit-deterministic), one final kernel.
#include "common.cuh"
#include "fft_plan.h"
#include "tsdgpu.h"

#include <cfloat>
#include <vector>

struct tsdgpu_spectrum_s
{
  int device = 0;
  int BS = 0, nmeans = 0, nsubs = 0, sweep = 0, step = 0, Nf = 0, Ns = 0, nchan = 0;
  int cntmag = 0;
  float *d_fen = nullptr, *d_masque = nullptr, *d_cnt = nullptr;   // [Nf], [Nf], [Ns]
  float *d_mag = nullptr;                                          // [nchan][Ns] running sums
  float2 *d_work = nullptr;                                        // [nchan * nsubs][Nf]
  float2 *d_x = nullptr;                                           // host calls: staged input [nchan][BS]
  float *d_y = nullptr;                                            // host calls: staged output [nchan][Ns]
  tsdgpu_fft_s *plan = nullptr;
};

namespace tsdgpu {

// work[(c * nsubs + i)][k] = x[c][i * Nf + k] * f[k]
__global__ void spectrum_window_kernel(const float2 *x, long long xs, const float *f, float2 *work, int Nf, int nsubs)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y, c = row / nsubs, i = row - c * nsubs;
  if(k >= Nf) return;
  const float2 v = x[(long long) c * xs + (long long) i * Nf + k];
  const float w = f[k];
  work[(long long) row * Nf + k] = make_float2(v.x * w, v.y * w);
}

// mag[c][p] += sum over the sub-blocks i that cover p of |X_i[(q + ceil(Nf / 2)) mod Nf]|^2 [* masque[q]], q = p - i * step
// (fftshift, fourier.hpp:233-248)
__global__ void spectrum_accum_kernel(const float2 *work, float *mag, const float *masque, int Nf, int Ns, int nsubs, int sweep, int step)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if(p >= Ns) return;
  const int sh = (Nf + 1) / 2;
  float acc = mag[(long long) c * Ns + p];
  if(!sweep)
  {
    int k = p + sh;
    if(k >= Nf) k -= Nf;
    for(int i = 0; i < nsubs; i++)
    {
      const float2 v = work[((long long) c * nsubs + i) * Nf + k];
      acc += v.x * v.x + v.y * v.y;
    }
  }
  else
  {
    for(int i = 0; i < nsubs; i++)
    {
      const int q = p - i * step;
      if(q < 0 || q >= Nf) continue;
      int k = q + sh;
      if(k >= Nf) k -= Nf;
      const float2 v = work[((long long) c * nsubs + i) * Nf + k];
      acc += (v.x * v.x + v.y * v.y) * masque[q];
    }
  }
  mag[(long long) c * Ns + p] = acc;
}

// y = 10 log10(mag / div [/ cnt] + FLT_MIN); mag = 0
__global__ void spectrum_final_kernel(float *mag, const float *cnt, float *y, long long ys, int Ns, float div, int sweep)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if(p >= Ns) return;
  float m = mag[(long long) c * Ns + p] / div;
  if(sweep) m = m / cnt[p];
  y[(long long) c * ys + p] = 10.f * log10f(m + FLT_MIN);
  mag[(long long) c * Ns + p] = 0.f;
}

} // namespace tsdgpu

using namespace tsdgpu;

extern "C" {

int tsdgpu_spectrum_destroy(tsdgpu_spectrum_t s)
{
  if(!s) return 0;
  TSD_ENTER(s->device);
  cudaStreamSynchronize(rt().stream);
  cudaFree(s->d_fen);
  cudaFree(s->d_masque);
  cudaFree(s->d_cnt);
  cudaFree(s->d_mag);
  cudaFree(s->d_work);
  cudaFree(s->d_x);
  cudaFree(s->d_y);
  if(s->plan) fft_plan_destroy(s->plan);
  delete s;
  return 0;
}

int tsdgpu_spectrum_create(int BS, int nmeans, int nsubs, int sweep_active, int sweep_step, int masque_bf, int masque_hf,
                           const float *fenetre, int nchan, tsdgpu_spectrum_t *out)
{
  TSD_ENTER(-1);
  if(!out || !fenetre) return fail("tsdgpu_spectrum_create: null argument");
  *out = nullptr;
  if(BS < 1 || nmeans < 1 || nsubs < 1 || nsubs > BS) return fail("tsdgpu_spectrum_create: invalid BS / nmeans / nsubs");
  if(nchan < 1 || (long long) nchan * nsubs > 65535) return fail("tsdgpu_spectrum_create: nchan * nsubs must be in [1, 65535]");
  const int Nf = BS / nsubs;
  if(sweep_active && sweep_step < 0) return fail("tsdgpu_spectrum_create: negative sweep step");
  if(masque_hf < 0 || masque_bf < 0 || masque_hf > Nf || Nf / 2 - masque_bf < 0 || Nf / 2 + masque_bf > Nf)
    return fail("tsdgpu_spectrum_create: mask wider than the spectrum");
  auto *s = new tsdgpu_spectrum_s;
  s->device = rt().device;
  s->BS = BS;
  s->nmeans = nmeans;
  s->nsubs = nsubs;
  s->sweep = sweep_active ? 1 : 0;
  s->step = sweep_step;
  s->Nf = Nf;
  s->Ns = sweep_active ? Nf + (nsubs - 1) * sweep_step : Nf;
  s->nchan = nchan;
  // masks as SpectrumConfig / configure_impl build them (fourier.cc:1180-1198)
  std::vector<float> masque((size_t) Nf, 1.f), cnt((size_t) s->Ns, 0.f);
  for(int i = 0; i < masque_hf; i++) masque[i] = masque[Nf - 1 - i] = 0.f;
  for(int i = 0; i < 2 * masque_bf; i++) masque[Nf / 2 - masque_bf + i] = 0.f;
  if(s->sweep)
  {
    for(int i = 0; i < nsubs; i++)
      for(int k = 0; k < Nf; k++) cnt[(size_t) i * sweep_step + k] += masque[k];
    for(auto &v : cnt) v = v > 1.f ? v : 1.f;
  }
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void **p, size_t bytes) { if(e == cudaSuccess) e = cudaMalloc(p, bytes); };
  alloc((void **) &s->d_fen, (size_t) Nf * 4);
  alloc((void **) &s->d_masque, (size_t) Nf * 4);
  alloc((void **) &s->d_cnt, (size_t) s->Ns * 4);
  alloc((void **) &s->d_mag, (size_t) nchan * s->Ns * 4);
  alloc((void **) &s->d_work, (size_t) nchan * nsubs * Nf * 8);
  if(e == cudaSuccess) e = cudaMemcpy(s->d_fen, fenetre, (size_t) Nf * 4, cudaMemcpyHostToDevice);
  if(e == cudaSuccess) e = cudaMemcpy(s->d_masque, masque.data(), (size_t) Nf * 4, cudaMemcpyHostToDevice);
  if(e == cudaSuccess) e = cudaMemcpy(s->d_cnt, cnt.data(), (size_t) s->Ns * 4, cudaMemcpyHostToDevice);
  if(e == cudaSuccess) e = cudaMemset(s->d_mag, 0, (size_t) nchan * s->Ns * 4);
  if(e != cudaSuccess)
  {
    tsdgpu_spectrum_destroy(s);
    return fail(std::string("tsdgpu_spectrum_create: ") + cudaGetErrorString(e));
  }
  if(fft_plan_create(Nf, nchan * nsubs, &s->plan))
  {
    tsdgpu_spectrum_destroy(s);
    return 1;
  }
  *out = s;
  return 0;
}

int tsdgpu_spectrum_dims(tsdgpu_spectrum_t s, int *Nf, int *Ns)
{
  if(!s) return fail("tsdgpu_spectrum_dims: null handle");
  if(Nf) *Nf = s->Nf;
  if(Ns) *Ns = s->Ns;
  return 0;
}

int tsdgpu_spectrum_step(tsdgpu_spectrum_t s, const void *x, long long x_stride, int n, float *y, long long y_stride, int *n_out, int mem)
{
  TSD_ENTER(s ? s->device : -1);
  if(!s || !x || !n_out) return fail("tsdgpu_spectrum_step: null argument");
  if(n != s->BS) return fail("tsdgpu_spectrum_step: a block must hold BS samples (\"Spectrum : dimension invalide\", fourier.cc:1236)");
  if(x_stride < n) return fail("tsdgpu_spectrum_step: channel stride smaller than n");
  const bool last = s->cntmag + 1 == s->nmeans;
  if(last && (!y || y_stride < s->Ns)) return fail("tsdgpu_spectrum_step: this block completes the average: y must hold Ns values per channel");
  Runtime &r = rt();
  const float2 *dx = (const float2 *) x;
  long long dxs = x_stride;
  if(mem != TSDGPU_DEVICE)
  {
    if(!s->d_x) TSD_CUDA(cudaMalloc(&s->d_x, (size_t) s->nchan * s->BS * 8));
    TSD_CUDA(cudaMemcpy2DAsync(s->d_x, (size_t) n * 8, x, (size_t) x_stride * 8, (size_t) n * 8, s->nchan, cudaMemcpyHostToDevice, r.stream));
    dx = s->d_x;
    dxs = n;
  }
  const int rows = s->nchan * s->nsubs;
  spectrum_window_kernel<<<dim3((s->Nf + 255) / 256, rows), 256, 0, r.stream>>>(dx, dxs, s->d_fen, s->d_work, s->Nf, s->nsubs);
  TSD_LAUNCH_CHECK();
  if(fft_exec_device(s->plan, s->d_work, s->Nf, s->d_work, s->Nf, true)) return 1;
  spectrum_accum_kernel<<<dim3((s->Ns + 255) / 256, s->nchan), 256, 0, r.stream>>>(s->d_work, s->d_mag, s->d_masque, s->Nf, s->Ns, s->nsubs,
                                                                                 s->sweep, s->step);
  TSD_LAUNCH_CHECK();
  s->cntmag++;
  *n_out = 0;
  if(last)
  {
    float *dy = y;
    long long dys = y_stride;
    if(mem != TSDGPU_DEVICE)
    {
      if(!s->d_y) TSD_CUDA(cudaMalloc(&s->d_y, (size_t) s->nchan * s->Ns * 4));
      dy = s->d_y;
      dys = s->Ns;
    }
    const float div = (float) (s->nmeans * s->nsubs * s->Nf);   // integer product, then float (fourier.cc:1273)
    spectrum_final_kernel<<<dim3((s->Ns + 255) / 256, s->nchan), 256, 0, r.stream>>>(s->d_mag, s->d_cnt, dy, dys, s->Ns, div, s->sweep);
    TSD_LAUNCH_CHECK();
    if(mem != TSDGPU_DEVICE)
    {
      TSD_CUDA(cudaMemcpy2DAsync(y, (size_t) y_stride * 4, s->d_y, (size_t) s->Ns * 4, (size_t) s->Ns * 4, s->nchan, cudaMemcpyDeviceToHost, r.stream));
      TSD_CUDA(cudaStreamSynchronize(r.stream));
    }
    s->cntmag = 0;
    *n_out = s->Ns;
  }
  else if(mem != TSDGPU_DEVICE) TSD_CUDA(cudaStreamSynchronize(r.stream));   // the caller may reuse x
  return 0;
}

} // extern "C"
