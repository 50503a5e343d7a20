// Runtime of libtsdgpu: device selection, the library stream, error reporting, launch counter.
#include "common.cuh"
#include "host_pipe.cuh"
#include "tsdgpu.h"

#include <algorithm>
#include <cmath>
#include <vector>

namespace tsdgpu {

static thread_local std::string g_error;

constexpr int MAX_DEVICES = 64;
static Runtime g_rt[MAX_DEVICES];
static std::mutex g_table_mu;          // guards creation / destruction of the per-device runtimes
static int g_default_device = -1;      // first device initialised in this process
static thread_local int t_device = -1; // device the calling thread works on

Runtime &rt() { return g_rt[t_device < 0 ? (g_default_device < 0 ? 0 : g_default_device) : t_device]; }
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
  if(device < 0 || device >= count || device >= MAX_DEVICES) return fail("libtsdgpu: device index out of range");
  TSD_CUDA(cudaSetDevice(device));
  t_device = device;
  std::lock_guard<std::mutex> table_guard(g_table_mu);
  if(g_default_device < 0) g_default_device = device;
  Runtime &r = g_rt[device];
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

int ensure_init() { return enter_device(-1); }

int enter_device(int device)
{
  if(device < 0) device = t_device >= 0 ? t_device : (g_default_device >= 0 ? g_default_device : 0);
  if(device < MAX_DEVICES && g_rt[device].device == device && g_rt[device].own_stream)
  {
    t_device = device;
    cudaSetDevice(device);   // CUDA's current device is per host thread
    return 0;
  }
  return init_device(device);
}

// releases everything the runtime of one device owns (objects must have been destroyed by their owners)
static void shutdown_device(Runtime &r)
{
  if(r.device < 0) return;
  cudaSetDevice(r.device);
  cudaDeviceSynchronize();
  for(int i = 0; i < Runtime::MAX_AUX; i++)
  {
    if(r.aux[i]) cudaStreamDestroy(r.aux[i]);
    if(r.ev_join[i]) cudaEventDestroy(r.ev_join[i]);
    r.aux[i] = nullptr;
    r.ev_join[i] = nullptr;
  }
  if(r.ev_fork) cudaEventDestroy(r.ev_fork);
  r.ev_fork = nullptr;
  if(r.tw256) cudaFree(r.tw256);
  if(r.tw4) cudaFree(r.tw4);
  r.tw256 = nullptr;
  r.tw4 = nullptr;
  HostStage &hs = r.hs;
  for(int i = 0; i < 2; i++)
  {
    if(hs.in[i]) cudaFree(hs.in[i]);
    if(hs.out[i]) cudaFree(hs.out[i]);
    if(hs.ev_in[i]) cudaEventDestroy(hs.ev_in[i]);
    if(hs.ev_done[i]) cudaEventDestroy(hs.ev_done[i]);
    if(hs.ev_out[i]) cudaEventDestroy(hs.ev_out[i]);
    hs.in[i] = hs.out[i] = nullptr;
    hs.ev_in[i] = hs.ev_done[i] = hs.ev_out[i] = nullptr;
  }
  hs.in_bytes = hs.out_bytes = 0;
  for(auto &pr : r.timed)
  {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  r.timed.clear();
  if(r.own_stream) cudaStreamDestroy(r.own_stream);
  if(r.copy_in) cudaStreamDestroy(r.copy_in);
  if(r.copy_out) cudaStreamDestroy(r.copy_out);
  r.own_stream = r.stream = r.copy_in = r.copy_out = nullptr;
  r.ols_ready = false;
  r.timing = false;
  r.launches = 0;
  r.device = -1;
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

HostStage &host_stage() { return rt().hs; }
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

int tsdgpu_init_devices(const int *devices, int n)
{
  if(!devices || n <= 0) return fail("tsdgpu_init_devices: empty device list");
  for(int i = n - 1; i >= 0; i--)
    if(init_device(devices[i])) return 1;   // the calling thread ends up on devices[0]
  return 0;
}

int tsdgpu_set_device(int device) { return enter_device(device); }

int tsdgpu_current_device(void) { return rt().device; }

int tsdgpu_shutdown(void)
{
  std::lock_guard<std::mutex> table_guard(g_table_mu);
  for(int d = 0; d < MAX_DEVICES; d++)
  {
    std::lock_guard<std::recursive_mutex> g(g_rt[d].mu);
    shutdown_device(g_rt[d]);
  }
  g_default_device = -1;
  t_device = -1;
  return 0;
}

int tsdgpu_set_stream(void *s)
{
  TSD_ENTER(-1);
  rt().stream = s ? (cudaStream_t) s : rt().own_stream;
  return 0;
}

int tsdgpu_synchronize(void)
{
  TSD_ENTER(-1);
  TSD_CUDA(cudaStreamSynchronize(rt().stream));
  return 0;
}

// Final gather of per-device result shards: shard i, `bytes[i]` bytes on device src_devices[i], lands at
// dst + dst_offsets[i] on dst_device.  Peer copies over NVLink on the source devices' library streams, so each shard leaves
// as soon as the kernels that produce it have finished; returns when every shard has arrived.
int tsdgpu_gather(void *dst, int dst_device, const long long *dst_offsets, const void *const *srcs, const int *src_devices,
                  const long long *bytes, int n)
{
  if(!dst || !dst_offsets || !srcs || !src_devices || !bytes || n <= 0) return fail("tsdgpu_gather: null argument");
  const int back = rt().device;
  for(int i = 0; i < n; i++)
  {
    if(bytes[i] <= 0) continue;
    if(enter_device(src_devices[i])) return 1;
    std::lock_guard<std::recursive_mutex> g(rt().mu);
    if(src_devices[i] != dst_device)
    {
      int can = 0;
      cudaDeviceCanAccessPeer(&can, src_devices[i], dst_device);
      if(can)
      {
        cudaError_t e = cudaDeviceEnablePeerAccess(dst_device, 0);
        if(e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(std::string("tsdgpu_gather: ") + cudaGetErrorString(e));
        cudaGetLastError();
      }
    }
    TSD_CUDA(cudaMemcpyPeerAsync((char *) dst + dst_offsets[i], dst_device, srcs[i], src_devices[i], (size_t) bytes[i], rt().stream));
  }
  for(int i = 0; i < n; i++)
  {
    if(bytes[i] <= 0) continue;
    if(enter_device(src_devices[i])) return 1;
    TSD_CUDA(cudaStreamSynchronize(rt().stream));
  }
  if(back >= 0) enter_device(back);
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
  TSD_ENTER(-1);
  rt().timing = on != 0;
  return 0;
}

int tsdgpu_timing_read(double *total_ms, long long *launches)
{
  TSD_ENTER(-1);
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
