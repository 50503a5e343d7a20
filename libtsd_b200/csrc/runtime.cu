// Runtime of libtsdgpu: device selection, the library stream, error reporting, launch counter.
#include "common.cuh"
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <thread>
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
    if(hs.pin_in[i]) cudaFreeHost(hs.pin_in[i]);
    if(hs.pin_out[i]) cudaFreeHost(hs.pin_out[i]);
    if(hs.ev_pin_in[i]) cudaEventDestroy(hs.ev_pin_in[i]);
    if(hs.ev_pin_out[i]) cudaEventDestroy(hs.ev_pin_out[i]);
    hs.pin_in[i] = hs.pin_out[i] = nullptr;
    hs.pin_in_bytes[i] = hs.pin_out_bytes[i] = 0;
    hs.ev_pin_in[i] = hs.ev_pin_out[i] = nullptr;
    hs.pend[i].active = false;
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

// ---- copy-thread pool: parallel memcpy between pageable caller memory and the pinned bounce buffers ------------------
namespace {
struct CopyPool
{
  std::vector<std::thread> th;
  std::mutex m, user;               // m: pool state; user: one parallel copy at a time
  std::condition_variable cv, cv_done;
  const std::function<void(int)> *job = nullptr;
  int ntasks = 0, done = 0, active = 0;   // active: workers inside pull() — run() does not return before it is 0 again
  std::atomic<int> next{0};
  unsigned long gen = 0;
  bool stop = false;
  int nthreads = 0;

  void start()
  {
    if(nthreads) return;
    const char *v = getenv("TSDGPU_COPY_THREADS");
    const unsigned hw = std::thread::hardware_concurrency();
    nthreads = v ? std::max(1, atoi(v)) : (int) std::max(1u, std::min(8u, hw ? hw / 2 : 4u));
    for(int i = 1; i < nthreads; i++) th.emplace_back([this] { worker(); });
  }
  void pull()
  {
    for(;;)
    {
      const int i = next.fetch_add(1);
      if(i >= ntasks) break;
      (*job)(i);
      std::lock_guard<std::mutex> g(m);
      if(++done == ntasks) cv_done.notify_all();
    }
  }
  void worker()
  {
    unsigned long seen = 0;
    for(;;)
    {
      {
        std::unique_lock<std::mutex> g(m);
        cv.wait(g, [&] { return stop || gen != seen; });
        if(stop) return;
        seen = gen;
        active++;
      }
      pull();
      std::lock_guard<std::mutex> g(m);
      if(--active == 0) cv_done.notify_all();
    }
  }
  void run(int n, const std::function<void(int)> &f)
  {
    std::lock_guard<std::mutex> u(user);
    start();
    {
      std::lock_guard<std::mutex> g(m);
      job = &f;
      ntasks = n;
      done = 0;
      next.store(0);
      gen++;
    }
    cv.notify_all();
    pull();
    std::unique_lock<std::mutex> g(m);
    cv_done.wait(g, [&] { return done >= ntasks && active == 0; });
    ntasks = 0;   // workers that wake up late for this generation find nothing
  }
  ~CopyPool()
  {
    {
      std::lock_guard<std::mutex> g(m);
      stop = true;
    }
    cv.notify_all();
    for(auto &t : th) t.join();
  }
};
CopyPool &copy_pool()
{
  static CopyPool p;
  return p;
}
// rows of `width` bytes, `height` of them; cut into pieces of >= 1 MiB, one task each
void parallel_copy2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height)
{
  const size_t total = width * height;
  if(total == 0) return;
  const size_t piece = std::max<size_t>(1u << 20, (total + 63) / 64);
  const int ntasks = (int) ((total + piece - 1) / piece);
  const std::function<void(int)> f = [&](int i) {
    size_t off = (size_t) i * piece;
    const size_t end = std::min(total, off + piece);
    while(off < end)
    {
      const size_t row = off / width, col = off % width, len = std::min(width - col, end - off);
      memcpy((char *) dst + row * dpitch + col, (const char *) src + row * spitch + col, len);
      off += len;
    }
  };
  if(ntasks == 1) f(0);
  else copy_pool().run(ntasks, f);
}
bool host_ptr_is_pinned(const void *p)
{
  cudaPointerAttributes a;
  if(cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
int pin_reserve(void **buf, size_t *have, size_t need, cudaEvent_t *ev, cudaStream_t s)
{
  if(!*ev) TSD_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
  if(need <= *have) return 0;
  TSD_CUDA(cudaStreamSynchronize(s));
  if(*buf) cudaFreeHost(*buf);
  *buf = nullptr;
  *have = 0;
  TSD_CUDA(cudaHostAlloc(buf, need, cudaHostAllocDefault));
  *have = need;
  return 0;
}
int stage_drain(HostStage &hs, int slot)
{
  HostStage::Pending &pd = hs.pend[slot];
  if(!pd.active) return 0;
  TSD_CUDA(cudaEventSynchronize(hs.ev_pin_out[slot]));
  parallel_copy2d(pd.dst, pd.dpitch, hs.pin_out[slot], pd.width, pd.width, pd.height);
  pd.active = false;
  return 0;
}
} // namespace

int stage_in(int slot, void *dst_dev, size_t dpitch, const void *src_host, size_t spitch, size_t width, size_t height)
{
  Runtime &r = rt();
  HostStage &hs = r.hs;
  static const bool off = getenv("TSDGPU_NO_BOUNCE") != nullptr;   // A/B: let the driver stage pageable memory itself
  if(off || width * height < (1u << 20) || host_ptr_is_pinned(src_host))
  {
    TSD_CUDA(cudaMemcpy2DAsync(dst_dev, dpitch, src_host, spitch, width, height, cudaMemcpyHostToDevice, r.copy_in));
    return 0;
  }
  if(pin_reserve(&hs.pin_in[slot], &hs.pin_in_bytes[slot], width * height, &hs.ev_pin_in[slot], r.copy_in)) return 1;
  TSD_CUDA(cudaEventSynchronize(hs.ev_pin_in[slot]));   // the previous upload out of this bounce buffer has left it
  parallel_copy2d(hs.pin_in[slot], width, src_host, spitch, width, height);
  TSD_CUDA(cudaMemcpy2DAsync(dst_dev, dpitch, hs.pin_in[slot], width, width, height, cudaMemcpyHostToDevice, r.copy_in));
  TSD_CUDA(cudaEventRecord(hs.ev_pin_in[slot], r.copy_in));
  return 0;
}

int stage_out(int slot, void *dst_host, size_t dpitch, const void *src_dev, size_t spitch, size_t width, size_t height)
{
  Runtime &r = rt();
  HostStage &hs = r.hs;
  static const bool off = getenv("TSDGPU_NO_BOUNCE") != nullptr;
  if(off || width * height < (1u << 20) || host_ptr_is_pinned(dst_host))
  {
    TSD_CUDA(cudaMemcpy2DAsync(dst_host, dpitch, src_dev, spitch, width, height, cudaMemcpyDeviceToHost, r.copy_out));
    return 0;
  }
  if(stage_drain(hs, slot)) return 1;   // the chunk that used this bounce buffer two steps ago
  if(pin_reserve(&hs.pin_out[slot], &hs.pin_out_bytes[slot], width * height, &hs.ev_pin_out[slot], r.copy_out)) return 1;
  TSD_CUDA(cudaMemcpy2DAsync(hs.pin_out[slot], width, src_dev, spitch, width, height, cudaMemcpyDeviceToHost, r.copy_out));
  TSD_CUDA(cudaEventRecord(hs.ev_pin_out[slot], r.copy_out));
  hs.pend[slot].active = true;
  hs.pend[slot].dst = dst_host;
  hs.pend[slot].dpitch = dpitch;
  hs.pend[slot].width = width;
  hs.pend[slot].height = height;
  return 0;
}

int stage_flush()
{
  HostStage &hs = rt().hs;
  for(int s = 0; s < 2; s++)
    if(stage_drain(hs, s)) return 1;
  return 0;
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
