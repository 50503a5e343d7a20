// Runtime of libtsdgpu: device selection, the library stream, error reporting, launch counter.
#include "common.cuh"
#include "host_pipe.cuh"
#include "tsdgpu.h"

#include <algorithm>
#include <cmath>
#include <vector>

namespace tsdgpu {

static thread_local std::string g_error;

Runtime &rt()
{
  static Runtime r;
  return r;
}
void set_error(const std::string &s) { g_error = s; }
int fail(const std::string &s)
{
  g_error = s;
  return 1;
}

static int init_device(int device)
{
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if(e != cudaSuccess || count == 0)
    return fail(std::string("libtsdgpu: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
  if(device < 0 || device >= count) return fail("libtsdgpu: device index out of range");
  TSD_CUDA(cudaSetDevice(device));
  Runtime &r = rt();
  if(r.device == device && r.own_stream) return 0;
  cudaDeviceProp prop;
  TSD_CUDA(cudaGetDeviceProperties(&prop, device));
  if(prop.major < 10)
    return fail("libtsdgpu: built for sm_100a (Blackwell B200) only; found compute capability " +
                std::to_string(prop.major) + "." + std::to_string(prop.minor));
  r.device = device;
  r.tw256 = nullptr;   // auxiliary streams and tables belong to a device: rebuilt on demand (aux_init)
  r.num_sms = prop.multiProcessorCount;
  TSD_CUDA(cudaStreamCreateWithFlags(&r.own_stream, cudaStreamNonBlocking));
  TSD_CUDA(cudaStreamCreateWithFlags(&r.copy_in, cudaStreamNonBlocking));
  TSD_CUDA(cudaStreamCreateWithFlags(&r.copy_out, cudaStreamNonBlocking));
  r.stream = r.own_stream;
  return 0;
}

int ensure_init()
{
  if(rt().device >= 0)
  {
    // CUDA's current device is per host thread
    cudaSetDevice(rt().device);
    return 0;
  }
  return init_device(0);
}

int aux_init()
{
  Runtime &r = rt();
  if(r.tw256) return 0;
  // aux[i] has a higher scheduling priority than aux[i-1]: the staged pipelines put their LAST stage on the
  // highest-priority stream so that blocks are drained before new ones are started
  int least = 0, greatest = 0;
  TSD_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  for(int i = 0; i < Runtime::MAX_AUX; i++)
  {
    TSD_CUDA(cudaStreamCreateWithPriority(&r.aux[i], cudaStreamNonBlocking, std::max(greatest, least - i)));
    TSD_CUDA(cudaEventCreateWithFlags(&r.ev_join[i], cudaEventDisableTiming));
  }
  TSD_CUDA(cudaEventCreateWithFlags(&r.ev_fork, cudaEventDisableTiming));
  // tw[k*16 + i] = {w, i*w}, w = exp(-2 pi i * i*k / 256) (double on the host, like the reference's
  // twiddle generation fourier.cc:32-46), then the same for conj(w)
  std::vector<float4> h(512);
  for(int t = 0; t < 256; t++)
  {
    const double a = -2.0 * M_PI * (double) ((t >> 4) * (t & 15)) / 256.0;
    const float c = (float) cos(a), s = (float) sin(a);
    h[t] = make_float4(c, s, -s, c);
    h[256 + t] = make_float4(c, -s, s, c);
  }
  float4 *d = nullptr;
  TSD_CUDA(cudaMalloc(&d, 512 * sizeof(float4)));
  TSD_CUDA(cudaMemcpy(d, h.data(), 512 * sizeof(float4), cudaMemcpyHostToDevice));
  std::vector<float2> h4(8448);
  for(int a = 0; a < 16; a++)
    for(int n = 0; n < 256; n++)
    {
      const double ang = -2.0 * M_PI * (double) (a * n) / 65536.0;
      h4[a * 256 + n] = h4[4096 + n * 16 + a] = make_float2((float) cos(ang), (float) sin(ang));
    }
  for(int n = 0; n < 256; n++)
  {
    const double ang = -2.0 * M_PI * (double) (16 * n) / 65536.0;
    h4[8192 + n] = make_float2((float) cos(ang), (float) sin(ang));
  }
  float2 *d4 = nullptr;
  TSD_CUDA(cudaMalloc(&d4, h4.size() * sizeof(float2)));
  TSD_CUDA(cudaMemcpy(d4, h4.data(), h4.size() * sizeof(float2), cudaMemcpyHostToDevice));
  r.tw4 = d4;
  r.tw256 = d;
  return 0;
}
int aux_fork(int n)
{
  Runtime &r = rt();
  TSD_CUDA(cudaEventRecord(r.ev_fork, r.stream));
  for(int i = 0; i < n; i++) TSD_CUDA(cudaStreamWaitEvent(r.aux[i], r.ev_fork, 0));
  return 0;
}
int aux_join(int n)
{
  Runtime &r = rt();
  for(int i = 0; i < n; i++)
  {
    TSD_CUDA(cudaEventRecord(r.ev_join[i], r.aux[i]));
    TSD_CUDA(cudaStreamWaitEvent(r.stream, r.ev_join[i], 0));
  }
  return 0;
}

HostStage &host_stage()
{
  static HostStage hs;
  return hs;
}
int host_stage_reserve(size_t in_bytes, size_t out_bytes)
{
  HostStage &hs = host_stage();
  Runtime &r = rt();
  if(!hs.ev_in[0])
    for(int i = 0; i < 2; i++)
    {
      TSD_CUDA(cudaEventCreateWithFlags(&hs.ev_in[i], cudaEventDisableTiming));
      TSD_CUDA(cudaEventCreateWithFlags(&hs.ev_done[i], cudaEventDisableTiming));
      TSD_CUDA(cudaEventCreateWithFlags(&hs.ev_out[i], cudaEventDisableTiming));
    }
  if(in_bytes > hs.in_bytes || out_bytes > hs.out_bytes)
  {
    TSD_CUDA(cudaStreamSynchronize(r.stream));
    TSD_CUDA(cudaStreamSynchronize(r.copy_in));
    TSD_CUDA(cudaStreamSynchronize(r.copy_out));
  }
  if(in_bytes > hs.in_bytes)
  {
    for(int i = 0; i < 2; i++)
    {
      if(hs.in[i]) cudaFree(hs.in[i]);
      hs.in[i] = nullptr;
      TSD_CUDA(cudaMalloc(&hs.in[i], in_bytes));
    }
    hs.in_bytes = in_bytes;
  }
  if(out_bytes > hs.out_bytes)
  {
    for(int i = 0; i < 2; i++)
    {
      if(hs.out[i]) cudaFree(hs.out[i]);
      hs.out[i] = nullptr;
      TSD_CUDA(cudaMalloc(&hs.out[i], out_bytes));
    }
    hs.out_bytes = out_bytes;
  }
  return 0;
}

KernelTimer::KernelTimer()
{
  Runtime &r = rt();
  if(!r.timing) return;
  if(cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { a = b = nullptr; return; }
  cudaEventRecord(a, r.stream);
}
KernelTimer::~KernelTimer()
{
  if(!a) return;
  Runtime &r = rt();
  cudaEventRecord(b, r.stream);
  r.timed.emplace_back(a, b);
}

} // namespace tsdgpu

using namespace tsdgpu;

extern "C" {

int tsdgpu_init(int device) { return init_device(device); }

int tsdgpu_set_stream(void *s)
{
  if(ensure_init()) return 1;
  rt().stream = s ? (cudaStream_t) s : rt().own_stream;
  return 0;
}

int tsdgpu_synchronize(void)
{
  if(ensure_init()) return 1;
  TSD_CUDA(cudaStreamSynchronize(rt().stream));
  return 0;
}

const char *tsdgpu_last_error(void) { return g_error.c_str(); }

long long tsdgpu_launch_count(int reset)
{
  long long v = rt().launches;
  if(reset) rt().launches = 0;
  return v;
}

int tsdgpu_timing_enable(int on)
{
  if(ensure_init()) return 1;
  rt().timing = on != 0;
  return 0;
}

int tsdgpu_timing_read(double *total_ms, long long *launches)
{
  if(ensure_init()) return 1;
  Runtime &r = rt();
  TSD_CUDA(cudaStreamSynchronize(r.stream));
  double tot = 0;
  for(auto &pr : r.timed)
  {
    float ms = 0;
    cudaEventSynchronize(pr.second);
    cudaEventElapsedTime(&ms, pr.first, pr.second);
    tot += ms;
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  if(total_ms) *total_ms = tot;
  if(launches) *launches = (long long) r.timed.size();
  r.timed.clear();
  return 0;
}

// tsd.cc:287-291 — same float expression as the reference, evaluated on the host
int tsdgpu_p2(int i)
{
  int lg2 = (int) ceilf(logf((float) i) / logf(2.0f));
  return (int) (1l << lg2);
}

// fourier.cc:708-713 ola_complexité: same float expressions (log of a float is logf, evaluated left to right)
int tsdgpu_ola_complexite(int M, int Ne, float *C, int *Nf, int *Nz)
{
  if(Ne <= 0 || M <= 0) return fail("tsdgpu_ola_complexite: M and Ne must be > 0");
  const int nf = tsdgpu_p2(Ne + M - 1);
  if(Nf) *Nf = nf;
  if(Nz) *Nz = nf - Ne;
  if(C) *C = (1.0f / Ne) * 2 * 5 * nf * logf(1.0f * nf) / logf(2.0f);
  return 0;
}

// fourier.cc:715-735 ola_complexité_optimise: kmin from a DOUBLE log (integer argument), 20 candidates
int tsdgpu_ola_complexite_optimise(int M, float *C_, int *Nf_, int *Nz_, int *Ne_)
{
  if(M <= 0) return fail("tsdgpu_ola_complexite_optimise: M must be > 0");
  const int kmin = (int) ceil(log((double) M) / log(2.0));
  float best = 0;
  int bnf = 0, bne = 0;
  for(int k = kmin; (k < kmin + 20) && (k < 31); k++)
  {
    const int Ne = (1 << k) - (M - 1);
    float C;
    int Nf, Nz;
    tsdgpu_ola_complexite(M, Ne, &C, &Nf, &Nz);
    if((k == kmin) || (C < best))
    {
      bnf = Nf;
      bne = Ne;
      best = C;
    }
  }
  if(C_) *C_ = best;
  if(Nf_) *Nf_ = bnf;
  if(Nz_) *Nz_ = bnf - bne;
  if(Ne_) *Ne_ = bne;
  return 0;
}

} // extern "C"
