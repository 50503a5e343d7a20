// Direct-form FIR on sm_100a.  Replaces FiltreRIF<T,Tc>::step (reference filtre-rt.cc:53-109):
//   y[n] = sum_{k<K} h[k] x[n-k], accumulated oldest sample first (h[K-1] first) like the
//   reference's loop (filtre-rt.cc:84-104), zeros before the first sample, state = last K-1 inputs.
//
// cf32 data with <= 127 real taps (BASELINE config 3) runs as a 3xTF32 banded Toeplitz GEMM on the tensor cores
// (fir_tc.cu); everything else (real data, complex taps, longer filters, unaligned rows) on the FP32 FMA kernel below.
//
// Kernel shape: one CTA = one tile of NT*R consecutive outputs of one channel.  The input tile
// (+ K-1 halo) is brought into shared memory by the TMA unit as 1-D bulk copies
// (cp.async.bulk, SASS UBLKCP) when the addresses allow it, the taps sit in shared memory in
// accumulation order, and each thread slides an R-sample register window over the tile:
// one LDS.64 + one broadcast LDS.32 per 2R FFMA.  R is odd so that the per-lane stride of the
// window loads (R*8 B) is bank-conflict free.  Results go back through shared memory so that the
// global stores are fully coalesced.
#include "common.cuh"
#include "fir_tc.h"
#include "ols16k.h"
#include "host_pipe.cuh"
#include "tsdgpu.h"

#include <complex>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace tsdgpu {

template<int DC> struct SampleT;
template<> struct SampleT<1> { using type = float; };
template<> struct SampleT<2> { using type = float2; };

template<int DC, int TC> struct Mac;
template<> struct Mac<1, 1>
{
  __device__ static __forceinline__ void run(float &acc, float w, float c) { acc = fmaf(w, c, acc); }
};
template<> struct Mac<2, 1>
{
  __device__ static __forceinline__ void run(float2 &acc, float2 w, float c)
  {
    acc.x = fmaf(w.x, c, acc.x);
    acc.y = fmaf(w.y, c, acc.y);
  }
};
template<> struct Mac<2, 2>
{
  // (a + ib)(c + id) as in std::complex operator* used by the reference
  __device__ static __forceinline__ void run(float2 &acc, float2 w, float2 c)
  {
    acc.x = fmaf(w.x, c.x, acc.x);
    acc.x = fmaf(-w.y, c.y, acc.x);
    acc.y = fmaf(w.x, c.y, acc.y);
    acc.y = fmaf(w.y, c.x, acc.y);
  }
};

struct FirParams
{
  const void *x;
  void *y;
  const void *hist;      // [nchan][halo] samples before x[0] (oldest first)
  const void *taps_rev;  // taps in accumulation order: taps_rev[m] = h[K-1-m]
  long long x_stride, y_stride;
  int n, K, halo, use_tma;
};

constexpr int FIR_NT = 256;

template<int DC, int TC, int R>
__global__ void __launch_bounds__(FIR_NT) fir_direct_kernel(FirParams p)
{
  using S = typename SampleT<DC>::type;
  using C = typename SampleT<TC>::type;
  constexpr int T = FIR_NT * R;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  // layout: [tile: halo + T + R samples][taps: K]
  S *tile = reinterpret_cast<S *>(smem_raw);
  const int tile_len = p.halo + T + R;
  C *taps = reinterpret_cast<C *>(smem_raw + (((size_t) tile_len * sizeof(S) + 15) & ~(size_t) 15));

  const int tid = threadIdx.x;
  const int chan = blockIdx.y;
  const int n0 = blockIdx.x * T;
  const S *x = reinterpret_cast<const S *>(p.x) + (long long) chan * p.x_stride;
  const S *hist = reinterpret_cast<const S *>(p.hist) + (long long) chan * p.halo;
  const int n_in_tile = min(T, p.n - n0);   // valid outputs of this tile

  // ---- stage the input tile: tile[j] = stream[n0 - halo + j]
  const int from_hist = max(0, p.halo - n0);              // leading samples taken from the history
  const int x_first = max(0, n0 - p.halo);                // first sample taken from x
  const int from_x = n0 + n_in_tile - x_first;
  if(p.use_tma)
  {
    if(tid == 0)
    {
      mbar_init(&bar, 1);
      mbar_fence_init();
      unsigned bytes = (unsigned) ((from_hist + from_x) * sizeof(S));
      mbar_expect_tx(&bar, bytes);
      if(from_hist > 0)
        bulk_g2s(tile, hist + (p.halo - from_hist), (unsigned) (from_hist * sizeof(S)), &bar);
      bulk_g2s(tile + from_hist, x + x_first, (unsigned) (from_x * sizeof(S)), &bar);
    }
  }
  else
  {
    for(int j = tid; j < from_hist; j += FIR_NT) tile[j] = hist[p.halo - from_hist + j];
    for(int j = tid; j < from_x; j += FIR_NT) tile[from_hist + j] = x[x_first + j];
  }
  for(int m = tid; m < p.K; m += FIR_NT) taps[m] = reinterpret_cast<const C *>(p.taps_rev)[m];
  // the window runs up to R samples past the last valid input: keep that slack finite
  for(int j = from_hist + from_x + tid; j < tile_len; j += FIR_NT)
  {
    S z;
    memset(&z, 0, sizeof(S));
    tile[j] = z;
  }
  __syncthreads();
  if(p.use_tma) mbar_wait(&bar, 0);

  // ---- register-blocked sliding window
  // output r of this thread at tap step m reads tile[base + r + m]
  const int base = tid * R + p.halo - (p.K - 1);
  S w[R], acc[R];
#pragma unroll
  for(int r = 0; r < R; r++)
  {
    w[r] = tile[base + r];
    memset(&acc[r], 0, sizeof(S));
  }
  const S *next = tile + base + R;
  int m0 = 0;
  for(; m0 + R <= p.K; m0 += R)
  {
#pragma unroll
    for(int j = 0; j < R; j++)
    {
      const C c = taps[m0 + j];
#pragma unroll
      for(int r = 0; r < R; r++) Mac<DC, TC>::run(acc[r], w[(r + j) % R], c);
      w[j] = next[m0 + j];
    }
  }
#pragma unroll
  for(int j = 0; j < R; j++)
  {
    if(m0 + j < p.K)
    {
      const C c = taps[m0 + j];
#pragma unroll
      for(int r = 0; r < R; r++) Mac<DC, TC>::run(acc[r], w[(r + j) % R], c);
      w[j] = next[m0 + j];
    }
  }

  // ---- coalesced write-back through shared memory (the tile is dead now)
  __syncthreads();
#pragma unroll
  for(int r = 0; r < R; r++) tile[tid * R + r] = acc[r];
  __syncthreads();
  S *y = reinterpret_cast<S *>(p.y) + (long long) chan * p.y_stride + n0;
  for(int j = tid; j < n_in_tile; j += FIR_NT) y[j] = tile[j];
}

// new_hist = last `halo` samples of (old_hist ++ x[0..n))
template<int DC>
__global__ void fir_hist_kernel(const void *x_, long long x_stride, int n, const void *old_, void *new_, int halo)
{
  using S = typename SampleT<DC>::type;
  const int chan = blockIdx.y;
  const S *x = reinterpret_cast<const S *>(x_) + (long long) chan * x_stride;
  const S *o = reinterpret_cast<const S *>(old_) + (long long) chan * halo;
  S *d = reinterpret_cast<S *>(new_) + (long long) chan * halo;
  for(int j = blockIdx.x * blockDim.x + threadIdx.x; j < halo; j += gridDim.x * blockDim.x)
  {
    int pos = n - halo + j;
    d[j] = (pos >= 0) ? x[pos] : o[halo + pos];
  }
}

} // namespace tsdgpu

using namespace tsdgpu;

struct tsdgpu_fir_s
{
  int device = 0;              // CUDA device the object lives on
  int kind = 0, K = 0, nchan = 0, halo = 0, DC = 1, TC = 1;
  long long total = 0;        // samples consumed per channel so far
  void *d_taps = nullptr;
  void *d_hist[2] = {nullptr, nullptr};
  int cur = 0;
  void *d_stage = nullptr;    // device staging for TSDGPU_HOST calls
  size_t stage_bytes = 0;
  Ols16k *ols = nullptr;      // cf32 data, K >= 128: the same filter on the single-SM overlap-save kernel (delay 0)
  // real-valued data, K >= 128: two channels ride the overlap-save kernel as the real and imaginary part of one complex channel
  float2 *d_pair[3] = {nullptr, nullptr, nullptr};   // packed input [pairs][n], packed output [pairs][n], packed history [pairs][halo]
  size_t pair_cap = 0;        // samples per pair row the buffers hold
};

template<int DC, int TC, int R>
static int fir_launch(tsdgpu_fir_s *f, const FirParams &p)
{
  constexpr int T = FIR_NT * R;
  size_t ssz = (DC == 1) ? 4 : 8, csz = (TC == 1) ? 4 : 8;
  size_t smem = ((((size_t) f->halo + T + R) * ssz + 15) & ~(size_t) 15) + (size_t) f->K * csz;
  auto kern = fir_direct_kernel<DC, TC, R>;
  if(smem > 48 * 1024)
    TSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  dim3 grid((p.n + T - 1) / T, f->nchan);
  {
    KernelTimer timer;
    kern<<<grid, FIR_NT, smem, rt().stream>>>(p);
    TSD_LAUNCH_CHECK();
  }
  return 0;
}

// real rows 2p, 2p+1 -> one complex row p (a missing odd partner reads as zero); and back
__global__ void fir_pair_pack_kernel(const float *x, long long xs, int n, int nchan, float2 *xp, long long xps)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
  if(i >= n) return;
  const float a = x[(long long) (2 * p) * xs + i], b = (2 * p + 1 < nchan) ? x[(long long) (2 * p + 1) * xs + i] : 0.f;
  xp[(long long) p * xps + i] = make_float2(a, b);
}
__global__ void fir_pair_unpack_kernel(const float2 *yp, long long yps, int n, int nchan, float *y, long long ys)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
  if(i >= n) return;
  const float2 v = yp[(long long) p * yps + i];
  y[(long long) (2 * p) * ys + i] = v.x;
  if(2 * p + 1 < nchan) y[(long long) (2 * p + 1) * ys + i] = v.y;
}

static int fir_run_device(tsdgpu_fir_s *f, const void *x, long long xs, int n, void *y, long long ys)
{
  if(n <= 0) return 0;
  const size_t ssz = (f->DC == 1) ? 4 : 8;
  // 1. next history (reads x before the main kernel may overwrite it when x == y)
  void *hist_old = f->d_hist[f->cur], *hist_new = f->d_hist[f->cur ^ 1];
  {
    dim3 grid((f->halo + 255) / 256, f->nchan);
    if(f->DC == 1) fir_hist_kernel<1><<<grid, 256, 0, rt().stream>>>(x, xs, n, hist_old, hist_new, f->halo);
    else fir_hist_kernel<2><<<grid, 256, 0, rt().stream>>>(x, xs, n, hist_old, hist_new, f->halo);
    TSD_LAUNCH_CHECK();
  }
  // real-valued data, 128 ... 8192 real taps, calls of >= 2048 samples: channels 2p, 2p+1 become the real and imaginary part of
  // one complex channel (exact for real taps), which the overlap-save kernel filters with delay 0; packing and unpacking are two
  // small streaming kernels (24 B of HBM traffic per real sample in all, whatever K; the direct form needs 2 K flop per sample)
  if(f->ols && f->DC == 1 && n >= 2048 && !(getenv("TSDGPU_FIR_OLS") && atoi(getenv("TSDGPU_FIR_OLS")) == 0))
  {
    const int pairs = (f->nchan + 1) / 2;
    if((size_t) n > f->pair_cap)
    {
      TSD_CUDA(cudaStreamSynchronize(rt().stream));
      for(int i = 0; i < 2; i++) { if(f->d_pair[i]) cudaFree(f->d_pair[i]); f->d_pair[i] = nullptr; }
      f->pair_cap = 0;
      const size_t cap = ((size_t) n + 1) & ~(size_t) 1;   // even row pitch: 16-byte aligned rows for the bulk copies
      for(int i = 0; i < 2; i++) TSD_CUDA(cudaMalloc(&f->d_pair[i], (size_t) pairs * cap * sizeof(float2)));
      if(!f->d_pair[2]) TSD_CUDA(cudaMalloc(&f->d_pair[2], (size_t) pairs * f->halo * sizeof(float2)));
      f->pair_cap = cap;
    }
    const long long ps = (long long) f->pair_cap;
    fir_pair_pack_kernel<<<dim3((n + 255) / 256, pairs), 256, 0, rt().stream>>>((const float *) x, xs, n, f->nchan, f->d_pair[0], ps);
    fir_pair_pack_kernel<<<dim3((f->halo + 255) / 256, pairs), 256, 0, rt().stream>>>((const float *) hist_old, f->halo, f->halo, f->nchan, f->d_pair[2],
                                                                                    f->halo);
    TSD_LAUNCH_CHECK();
    const int rc = ols16k_run(f->ols, f->d_pair[0], ps, n, f->d_pair[2], f->halo, f->d_pair[1], ps, n, 0, 0, pairs);
    if(rc) return rc;
    fir_pair_unpack_kernel<<<dim3((n + 255) / 256, pairs), 256, 0, rt().stream>>>(f->d_pair[1], ps, n, f->nchan, (float *) y, ys);
    TSD_LAUNCH_CHECK();
    f->cur ^= 1;
    f->total += n;
    return 0;
  }
  // 2. main kernel.  In-place operation needs a private copy of the input: tiles read the
  //    halo of their left neighbour, which that neighbour overwrites.
  const void *src = x;
  long long src_stride = xs;
  if(x == y)
  {
    size_t need = (size_t) f->nchan * n * ssz;
    if(need > f->stage_bytes)
    {
      if(f->d_stage) cudaFree(f->d_stage);
      f->d_stage = nullptr;
      f->stage_bytes = 0;
      TSD_CUDA(cudaMalloc(&f->d_stage, need));
      f->stage_bytes = need;
    }
    TSD_CUDA(cudaMemcpy2DAsync(f->d_stage, (size_t) n * ssz, x, (size_t) xs * ssz, (size_t) n * ssz, f->nchan,
                               cudaMemcpyDeviceToDevice, rt().stream));
    src = f->d_stage;
    src_stride = n;
  }
  FirParams p;
  p.x = src;
  p.y = y;
  p.hist = hist_old;
  p.taps_rev = f->d_taps;
  p.x_stride = src_stride;
  p.y_stride = ys;
  p.n = n;
  p.K = f->K;
  p.halo = f->halo;
  // bulk copies need 16-byte aligned addresses and sizes
  const int unit = (int) (16 / ssz);
  p.use_tma = (((uintptr_t) src & 15) == 0) && (src_stride % unit == 0) && (n % unit == 0) && (f->halo % unit == 0);
  int rc;
  // cf32 data, 128 ... 8192 taps (real or complex), calls of >= 2048 samples: y[n] = sum_k h[k] x[n-k] is what the single-SM
  // overlap-save kernel computes with delay 0 and the FIR history as its carry (ols16k.cu: 16 B per sample whatever K,
  // ~0.5e-6 of the RMS against the direct sum) -- the FP32 FMA kernel needs 4 K (8 K) flop per sample: 36 Gsamples/s at
  // K = 512.  TSDGPU_FIR_OLS=0 keeps the direct form.
  if(f->ols && f->DC == 2 && n >= 2048 && !(getenv("TSDGPU_FIR_OLS") && atoi(getenv("TSDGPU_FIR_OLS")) == 0))
  {
    rc = ols16k_run(f->ols, (const float2 *) src, src_stride, n, (const float2 *) hist_old, f->halo, (float2 *) y, ys, n, 0, 0, f->nchan);   // times itself
    if(rc) return rc;
    f->cur ^= 1;
    f->total += n;
    return 0;
  }
  // cf32 data, <= 127 real taps: banded Toeplitz GEMM on the tensor cores (3xTF32, fir_tc.cu); TSDGPU_FIR_TC=0 keeps FP32 FMA
  const char *tc_env = getenv("TSDGPU_FIR_TC");
  const bool tc_on = !(tc_env && atoi(tc_env) == 0);
  const bool tc_real = tc_on && f->kind == TSDGPU_FIR_F32_F32 && fir_tc_real_eligible(f->K, src, src_stride, y, ys, hist_old, f->halo);
  if(tc_real || (tc_on && fir_tc_eligible(f->kind == TSDGPU_FIR_CF32_F32, f->K, src, src_stride, y, ys, hist_old, f->halo)))
  {
    FirTcParams t;
    t.real = tc_real ? 1 : 0;
    t.x = (const float2 *) src;
    t.y = (float2 *) y;
    t.hist = (const float2 *) hist_old;
    t.taps_rev = (const float *) f->d_taps;
    t.x_stride = src_stride;
    t.y_stride = ys;
    t.n = n;
    t.K = f->K;
    t.halo = f->halo;
    t.nchan = f->nchan;
    {
      KernelTimer timer;
      rc = fir_tc_launch(t);
    }
  }
  else if(f->kind == TSDGPU_FIR_F32_F32) rc = fir_launch<1, 1, 9>(f, p);
  else if(f->kind == TSDGPU_FIR_CF32_F32) rc = fir_launch<2, 1, 9>(f, p);
  else rc = fir_launch<2, 2, 7>(f, p);
  if(rc) return rc;
  f->cur ^= 1;
  f->total += n;
  return 0;
}

extern "C" {

int tsdgpu_fir_create(int kind, const float *taps, int K, int nchan, tsdgpu_fir_t *out)
{
  TSD_ENTER(-1);
  if(!out || !taps) return fail("tsdgpu_fir_create: null argument");
  if(K <= 0) return fail("tsdgpu_fir_create: K must be > 0 (assertion K > 0, filtre-rt.cc:69)");
  if(nchan <= 0 || nchan > 65535) return fail("tsdgpu_fir_create: nchan must be in [1, 65535]");
  if(kind < 0 || kind > 2) return fail("tsdgpu_fir_create: unknown kind");
  if(K > 8192) return fail("tsdgpu_fir_create: K > 8192 is not supported by the direct form (use tsdgpu_ola_create)");
  auto *f = new tsdgpu_fir_s;
  f->device = rt().device;
  f->kind = kind;
  f->K = K;
  f->nchan = nchan;
  f->DC = (kind == TSDGPU_FIR_F32_F32) ? 1 : 2;
  f->TC = (kind == TSDGPU_FIR_CF32_CF32) ? 2 : 1;
  // history = last K samples (the reference ring holds K, filtre-rt.cc:56-64) rounded up to a
  // multiple of 4 samples so that bulk copies stay 16-byte aligned; the kernel needs K-1 of them
  f->halo = (K + 3) & ~3;
  const size_t ssz = (f->DC == 1) ? 4 : 8, csz = (f->TC == 1) ? 4 : 8;
  std::vector<float> rev((size_t) K * f->TC);
  for(int m = 0; m < K; m++)
    for(int c = 0; c < f->TC; c++) rev[(size_t) m * f->TC + c] = taps[(size_t) (K - 1 - m) * f->TC + c];
  TSD_CUDA(cudaMalloc(&f->d_taps, (size_t) K * csz));
  TSD_CUDA(cudaMemcpyAsync(f->d_taps, rev.data(), (size_t) K * csz, cudaMemcpyHostToDevice, rt().stream));
  for(int i = 0; i < 2; i++)
  {
    TSD_CUDA(cudaMalloc(&f->d_hist[i], (size_t) nchan * f->halo * ssz));
    TSD_CUDA(cudaMemsetAsync(f->d_hist[i], 0, (size_t) nchan * f->halo * ssz, rt().stream));
  }
  TSD_CUDA(cudaStreamSynchronize(rt().stream));
  // complex taps cost the direct form 8 K flop per sample and have no tensor-core kernel: the overlap-save kernel wins from ~32 taps on
  if((f->DC == 2 && (K >= 128 || (f->TC == 2 && K >= 32))) || (f->DC == 1 && K >= 128))
  {
    std::vector<std::complex<double>> ht((size_t) K);
    for(int m = 0; m < K; m++)
      ht[m] = f->TC == 2 ? std::complex<double>(taps[2 * m], taps[2 * m + 1]) : std::complex<double>(taps[m], 0.0);
    if(ols16k_create_taps(ht.data(), K, &f->ols)) { tsdgpu_fir_destroy(f); return 1; }
  }
  *out = f;
  return 0;
}

int tsdgpu_fir_step(tsdgpu_fir_t f, const void *x, long long xs, int n, void *y, long long ys, int mem)
{
  TSD_ENTER(f ? f->device : -1);
  if(!f) return fail("tsdgpu_fir_step: null handle");
  if(n < 0) return fail("tsdgpu_fir_step: n < 0");
  if(n == 0) return 0;
  if(!x || !y) return fail("tsdgpu_fir_step: null buffer");
  if(xs < n || ys < n) return fail("tsdgpu_fir_step: channel stride smaller than n");
  if(mem == TSDGPU_DEVICE) return fir_run_device(f, x, xs, n, y, ys);
  const size_t ssz = (f->DC == 1) ? 4 : 8;
  const long long chunk = host_chunk_len(f->nchan, ssz, n, 4);
  if(host_stage_reserve((size_t) f->nchan * chunk * ssz, (size_t) f->nchan * chunk * ssz)) return 1;
  HostStage &hs = host_stage();
  const char *xh = (const char *) x;
  char *yh = (char *) y;
  return host_pipeline(
    n, chunk,
    [&](int slot, long long first, long long count) -> int {
      if(stage_in(slot, hs.in[slot], (size_t) chunk * ssz, xh + (size_t) first * ssz, (size_t) xs * ssz,
                                 (size_t) count * ssz, f->nchan)) return 1;
      return 0;
    },
    [&](long long count) { return count; },
    [&](int slot, long long count, long long *got) -> int {
      *got = count;
      return fir_run_device(f, hs.in[slot], chunk, (int) count, hs.out[slot], chunk);
    },
    [&](int slot, long long out_first, long long count) -> int {
      if(stage_out(slot, yh + (size_t) out_first * ssz, (size_t) ys * ssz, hs.out[slot], (size_t) chunk * ssz,
                                 (size_t) count * ssz, f->nchan)) return 1;
      return 0;
    },
    nullptr);
}

int tsdgpu_fir_get_state(tsdgpu_fir_t f, void *fen_host, int *index)
{
  TSD_ENTER(f ? f->device : -1);
  if(!f) return fail("tsdgpu_fir_get_state: null handle");
  const size_t ssz = (f->DC == 1) ? 4 : 8;
  const int K = f->K, idx = (int) (f->total % K);
  if(index) *index = idx;
  if(fen_host)
  {
    std::vector<unsigned char> h((size_t) f->nchan * f->halo * ssz);
    TSD_CUDA(cudaStreamSynchronize(rt().stream));
    TSD_CUDA(cudaMemcpy(h.data(), f->d_hist[f->cur], h.size(), cudaMemcpyDeviceToHost));
    // reference ring: after t samples, fen[(t-1-j) mod K] = x[t-1-j] for j < K (filtre-rt.cc:88-90)
    unsigned char *out = (unsigned char *) fen_host;
    memset(out, 0, (size_t) f->nchan * K * ssz);
    for(int c = 0; c < f->nchan; c++)
      for(int j = 0; j < K; j++)
      {
        long long t = f->total - 1 - j;   // absolute sample number
        if(t < 0) break;
        int slot = (int) (t % K);
        memcpy(out + ((size_t) c * K + slot) * ssz, h.data() + ((size_t) c * f->halo + (f->halo - 1 - j)) * ssz, ssz);
      }
  }
  return 0;
}

int tsdgpu_fir_set_state(tsdgpu_fir_t f, const void *fen_host, int index)
{
  TSD_ENTER(f ? f->device : -1);
  if(!f || !fen_host) return fail("tsdgpu_fir_set_state: null argument");
  if(index < 0 || index >= f->K) return fail("tsdgpu_fir_set_state: index out of range");
  const size_t ssz = (f->DC == 1) ? 4 : 8;
  const int K = f->K;
  std::vector<unsigned char> h((size_t) f->nchan * f->halo * ssz, 0);
  const unsigned char *in = (const unsigned char *) fen_host;
  // newest sample sits at slot index-1, then backwards around the ring
  for(int c = 0; c < f->nchan; c++)
    for(int j = 0; j < K; j++)
    {
      int slot = ((index - 1 - j) % K + K) % K;
      memcpy(h.data() + ((size_t) c * f->halo + (f->halo - 1 - j)) * ssz, in + ((size_t) c * K + slot) * ssz, ssz);
    }
  TSD_CUDA(cudaStreamSynchronize(rt().stream));
  TSD_CUDA(cudaMemcpy(f->d_hist[f->cur], h.data(), h.size(), cudaMemcpyHostToDevice));
  // keep (total mod K) == index; history older than K samples is not observable
  f->total = (long long) K * 4 + index;
  return 0;
}

// Time-ordered history (oldest first): what a halo-split segment needs to start in the middle of a stream
int tsdgpu_fir_set_history(tsdgpu_fir_t f, const void *hist_host, long long samples_so_far)
{
  TSD_ENTER(f ? f->device : -1);
  if(!f || (!hist_host && f->K > 1)) return fail("tsdgpu_fir_set_history: null argument");
  if(samples_so_far < 0) return fail("tsdgpu_fir_set_history: negative sample count");
  const size_t ssz = (f->DC == 1) ? 4 : 8;
  const int L = f->K - 1;
  std::vector<unsigned char> h((size_t) f->nchan * f->halo * ssz, 0);
  for(int c = 0; c < f->nchan && L > 0; c++)
    memcpy(h.data() + ((size_t) c * f->halo + (f->halo - L)) * ssz, (const unsigned char *) hist_host + (size_t) c * L * ssz, (size_t) L * ssz);
  TSD_CUDA(cudaStreamSynchronize(rt().stream));
  TSD_CUDA(cudaMemcpy(f->d_hist[f->cur], h.data(), h.size(), cudaMemcpyHostToDevice));
  f->total = samples_so_far;
  return 0;
}

int tsdgpu_fir_destroy(tsdgpu_fir_t f)
{
  if(!f) return 0;
  TSD_ENTER(f->device);
  cudaFree(f->d_taps);
  cudaFree(f->d_hist[0]);
  cudaFree(f->d_hist[1]);
  if(f->d_stage) cudaFree(f->d_stage);
  ols16k_destroy(f->ols);
  for(int i = 0; i < 3; i++)
    if(f->d_pair[i]) cudaFree(f->d_pair[i]);
  delete f;
  return 0;
}

} // extern "C"
