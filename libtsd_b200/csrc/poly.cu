// Polyphase rate-change stages of the resampler chain on sm_100a.  Replaces (reference polyphase.cc):
//   filtre_rif_ups<float,T>(c, R)      FiltreRIFUps        :246-341   x R interpolator
//   filtre_rif_demi_bande<float,T>(c)  FiltreRIFDemiBande  :54-149    half-band decimator by 2
//   filtre_rif_decim<float,T>(c, R)    FiltreRIFDecim      :156-239   FIR + decimation by R
// which filtre_reechan chains in front of the LUT interpolator when the ratio leaves [0.5, 2) (ra.cc:104-177).
//
// All three are "window of the last L inputs times a coefficient row", evaluated for every output
// independently; what differs is the bookkeeping, kept exactly as in the reference:
//   ups    every input t yields R outputs i = 0..R-1: y[R t + i] = sum_j coefs'[R-1-i + j R] * w_t[j], L = K'/R,
//          coefs' = c * R zero-padded to a multiple of R (:259-269); w_t[j] = x[t - (L-1) + j] (oldest first);
//   decim  one output every R inputs, the first when the counter `cnt` (0 at start) has seen R-1 inputs (:213-218):
//          y = sum_j c[j] * w_t[j], L = K  -- the OLDEST sample meets c[0] (:229-235);
//   demi   same with R = 2, only even j, plus the literal 0.5f * w_t[K/2] whatever c[K/2] is (:120-139).
// Outputs per call: n*R (ups), (n + cnt) / R (decim, demi-bande); cnt <- (n + cnt) % R; ring index = samples mod L.
// The sums run oldest sample first like the reference loops; FMA contraction is the only arithmetic difference.
#include "common.cuh"
#include "host_pipe.cuh"
#include "tsdgpu.h"

#include <algorithm>
#include <vector>

namespace tsdgpu {

struct PolyParams
{
  const void *x;          // [nchan][x_stride]
  void *y;                // [nchan][y_stride]
  const void *hist;       // [nchan][L-1] inputs preceding x[0] (zeros at start)
  const float *rows;      // [nrows][L] coefficient rows, oldest sample first
  long long x_stride, y_stride;
  long long n_out;        // outputs per channel in this call
  int L, nrows;
  int up;                 // R for the interpolator (outputs per input), 1 otherwise
  int down;               // R for the decimators (inputs per output), 1 otherwise
  int first;              // decimators: input index (within this call) of the first output = R-1-cnt
};

template<int DC> struct PolySample;
template<> struct PolySample<1> { using type = float; };
template<> struct PolySample<2> { using type = float2; };

template<int DC> __global__ void __launch_bounds__(256) poly_kernel(PolyParams p)
{
  using S = typename PolySample<DC>::type;
  extern __shared__ float s_rows[];
  for(int i = threadIdx.x; i < p.nrows * p.L; i += blockDim.x) s_rows[i] = p.rows[i];
  __syncthreads();
  const int chan = blockIdx.y;
  const S *x = (const S *) p.x + (long long) chan * p.x_stride;
  const S *h = (const S *) p.hist + (long long) chan * (p.L - 1) + (p.L - 1);   // h[-1] = newest carried input
  S *y = (S *) p.y + (long long) chan * p.y_stride;
  for(long long o = (long long) blockIdx.x * blockDim.x + threadIdx.x; o < p.n_out; o += (long long) gridDim.x * blockDim.x)
  {
    long long t;
    int row;
    if(p.up > 1) { t = o / p.up; row = (int) (o - t * p.up); }
    else { t = o * p.down + p.first; row = 0; }
    const float *c = s_rows + row * p.L;
    const long long t0 = t - (p.L - 1);   // oldest sample of the window
    if(DC == 1)
    {
      float acc = 0.f;
      for(int j = 0; j < p.L; j++)
      {
        const long long pos = t0 + j;
        const float v = pos >= 0 ? __ldg((const float *) x + pos) : __ldg((const float *) h + pos);
        acc = fmaf(v, c[j], acc);
      }
      ((float *) y)[o] = acc;
    }
    else
    {
      float2 acc = make_float2(0.f, 0.f);
      for(int j = 0; j < p.L; j++)
      {
        const long long pos = t0 + j;
        const float2 v = pos >= 0 ? __ldg((const float2 *) x + pos) : __ldg((const float2 *) h + pos);
        acc.x = fmaf(v.x, c[j], acc.x);
        acc.y = fmaf(v.y, c[j], acc.y);
      }
      ((float2 *) y)[o] = acc;
    }
  }
}

// new history = last H samples of (old history ++ x[0..n))
template<int DC> __global__ void poly_hist_kernel(const void *x_, long long x_stride, int n, const void *old_, void *new_, int H)
{
  using S = typename PolySample<DC>::type;
  const int chan = blockIdx.y;
  const S *x = (const S *) x_ + (long long) chan * x_stride;
  const S *o = (const S *) old_ + (long long) chan * H;
  S *d = (S *) new_ + (long long) chan * H;
  for(int j = blockIdx.x * blockDim.x + threadIdx.x; j < H; j += gridDim.x * blockDim.x)
  {
    const long long pos = (long long) n - H + j;
    d[j] = pos >= 0 ? x[pos] : o[H + pos];
  }
}

} // namespace tsdgpu

using namespace tsdgpu;

struct tsdgpu_poly_s
{
  int device = 0;              // CUDA device the object lives on
  int kind = 0, K = 0, R = 1, L = 0, nrows = 1, DC = 2, nchan = 1;
  int cnt = 0;               // decimation counter of the reference object (`cnt` / `odd`)
  long long total = 0;       // samples fed so far (ring index = total mod L)
  float *d_rows = nullptr;
  void *d_hist[2] = {nullptr, nullptr};
  int cur = 0;
};

static long long poly_out_count(const tsdgpu_poly_s *f, long long n)
{
  if(n <= 0) return 0;
  return f->kind == TSDGPU_POLY_UPS ? n * f->R : (n + f->cnt) / f->R;
}

static int poly_run_device(tsdgpu_poly_s *f, const void *x, long long xs, int n, void *y, long long ys, long long *n_out)
{
  *n_out = poly_out_count(f, n);
  if(n <= 0) return 0;
  Runtime &r = rt();
  const size_t ssz = f->DC == 1 ? 4 : 8;
  if(*n_out > 0)
  {
    if(ys < *n_out) return fail("tsdgpu_poly_step: output stride smaller than the emitted count");
    PolyParams p;
    p.x = x;
    p.y = y;
    p.hist = f->d_hist[f->cur];
    p.rows = f->d_rows;
    p.x_stride = xs;
    p.y_stride = ys;
    p.n_out = *n_out;
    p.L = f->L;
    p.nrows = f->nrows;
    p.up = f->kind == TSDGPU_POLY_UPS ? f->R : 1;
    p.down = f->kind == TSDGPU_POLY_UPS ? 1 : f->R;
    p.first = f->R - 1 - f->cnt;
    const size_t smem = (size_t) f->nrows * f->L * sizeof(float);
    const long long blocks = std::min<long long>((*n_out + 255) / 256, (long long) r.num_sms * 32);
    dim3 grid((unsigned) std::max<long long>(1, blocks), f->nchan);
    KernelTimer timer;
    if(f->DC == 1) poly_kernel<1><<<grid, 256, smem, r.stream>>>(p);
    else poly_kernel<2><<<grid, 256, smem, r.stream>>>(p);
    TSD_LAUNCH_CHECK();
  }
  if(f->L > 1)
  {
    dim3 grid((f->L - 1 + 255) / 256, f->nchan);
    if(f->DC == 1) poly_hist_kernel<1><<<grid, 256, 0, r.stream>>>(x, xs, n, f->d_hist[f->cur], f->d_hist[f->cur ^ 1], f->L - 1);
    else poly_hist_kernel<2><<<grid, 256, 0, r.stream>>>(x, xs, n, f->d_hist[f->cur], f->d_hist[f->cur ^ 1], f->L - 1);
    TSD_LAUNCH_CHECK();
    f->cur ^= 1;
  }
  (void) ssz;
  if(f->kind != TSDGPU_POLY_UPS) f->cnt = (int) (((long long) n + f->cnt) % f->R);
  f->total += n;
  return 0;
}

extern "C" {

int tsdgpu_poly_create(int kind, const float *coefs, int K, int R, int data_complex, int nchan, tsdgpu_poly_t *out)
{
  TSD_ENTER(-1);
  if(!out || !coefs) return fail("tsdgpu_poly_create: null argument");
  if(K <= 0) return fail("tsdgpu_poly_create: K must be > 0 (assertion K > 0, polyphase.cc:69,172,283)");
  if(kind < 0 || kind > 2) return fail("tsdgpu_poly_create: unknown kind");
  if(kind == TSDGPU_POLY_DEMI_BANDE) R = 2;   // polyphase.cc:60
  if(R < 1) return fail("tsdgpu_poly_create: R must be >= 1");
  if(nchan <= 0 || nchan > 65535) return fail("tsdgpu_poly_create: nchan must be in [1, 65535]");
  auto *f = new tsdgpu_poly_s;
  f->device = rt().device;
  f->kind = kind;
  f->K = K;
  f->R = R;
  f->DC = data_complex ? 2 : 1;
  f->nchan = nchan;
  std::vector<float> rows;
  if(kind == TSDGPU_POLY_UPS)
  {
    // coefs' = c * R, zero-padded to a multiple of R (polyphase.cc:259-269); row i, tap j = coefs'[R-1-i + j*R] (:312-326)
    const int Kp = ((K + R - 1) / R) * R;
    std::vector<float> cp((size_t) Kp, 0.f);
    for(int k = 0; k < K; k++) cp[k] = coefs[k] * (float) R;
    f->L = Kp / R;
    f->nrows = R;
    rows.resize((size_t) R * f->L);
    for(int i = 0; i < R; i++)
      for(int j = 0; j < f->L; j++) rows[(size_t) i * f->L + j] = cp[R - 1 - i + j * R];
  }
  else
  {
    f->L = K;
    f->nrows = 1;
    rows.assign((size_t) K, 0.f);
    if(kind == TSDGPU_POLY_DECIM)
      for(int j = 0; j < K; j++) rows[j] = coefs[j];                     // oldest sample meets c[0] (:229-235)
    else
    {
      for(int j = 0; j < K; j += 2) rows[j] = coefs[j];                  // every other coefficient (:120-134)
      rows[K / 2] += 0.5f;                                               // literal centre tap (:137)
    }
  }
  if((size_t) f->nrows * f->L * sizeof(float) > 48 * 1024)
  {
    delete f;
    return fail("tsdgpu_poly_create: coefficient table larger than 48 KiB");
  }
  const size_t ssz = f->DC == 1 ? 4 : 8;
  const size_t hbytes = (size_t) nchan * std::max(1, f->L - 1) * ssz;
  cudaError_t e = cudaMalloc(&f->d_rows, rows.size() * sizeof(float));
  if(e == cudaSuccess) e = cudaMemcpy(f->d_rows, rows.data(), rows.size() * sizeof(float), cudaMemcpyHostToDevice);
  for(int i = 0; i < 2 && e == cudaSuccess; i++)
  {
    e = cudaMalloc(&f->d_hist[i], hbytes);
    if(e == cudaSuccess) e = cudaMemsetAsync(f->d_hist[i], 0, hbytes, rt().stream);   // fenêtre = zeros (:65,168,276)
  }
  if(e == cudaSuccess) e = cudaStreamSynchronize(rt().stream);
  if(e != cudaSuccess)
  {
    tsdgpu_poly_destroy(f);
    return fail(std::string("tsdgpu_poly_create: ") + cudaGetErrorString(e));
  }
  *out = f;
  return 0;
}

long long tsdgpu_poly_out_count(tsdgpu_poly_t f, int n) { return f ? poly_out_count(f, n) : 0; }

int tsdgpu_poly_state(tsdgpu_poly_t f, int *index, int *cnt)
{
  if(!f) return fail("tsdgpu_poly_state: null handle");
  if(index) *index = (int) (f->total % f->L);
  if(cnt) *cnt = f->cnt;
  return 0;
}

int tsdgpu_poly_step(tsdgpu_poly_t f, const void *x, long long xs, int n, void *y, long long ys, long long *n_out, int mem)
{
  TSD_ENTER(f ? f->device : -1);
  if(!f || !n_out) return fail("tsdgpu_poly_step: null argument");
  *n_out = 0;
  if(n < 0) return fail("tsdgpu_poly_step: n < 0");
  if(n == 0) return 0;
  if(!x) return fail("tsdgpu_poly_step: null input");
  if(xs < n) return fail("tsdgpu_poly_step: channel stride smaller than n");
  if(poly_out_count(f, n) > 0 && !y) return fail("tsdgpu_poly_step: null output");
  if(x == y) return fail("tsdgpu_poly_step: in-place operation is not supported");
  if(mem == TSDGPU_DEVICE) return poly_run_device(f, x, xs, n, y, ys, n_out);
  const size_t ssz = f->DC == 1 ? 4 : 8;
  const long long chunk = host_chunk_len(f->nchan, ssz * (f->kind == TSDGPU_POLY_UPS ? f->R : 1), n, 1);
  const long long out_cap = f->kind == TSDGPU_POLY_UPS ? chunk * f->R : chunk / f->R + 1;
  if(host_stage_reserve((size_t) f->nchan * chunk * ssz, (size_t) f->nchan * out_cap * ssz)) return 1;
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
    [&](long long count) { return poly_out_count(f, count); },
    [&](int slot, long long count, long long *got) -> int {
      return poly_run_device(f, hs.in[slot], chunk, (int) count, hs.out[slot], out_cap, got);
    },
    [&](int slot, long long out_first, long long count) -> int {
      if(stage_out(slot, yh + (size_t) out_first * ssz, (size_t) ys * ssz, hs.out[slot], (size_t) out_cap * ssz,
                                 (size_t) count * ssz, f->nchan)) return 1;
      return 0;
    },
    n_out);
}

int tsdgpu_poly_get_state(tsdgpu_poly_t f, long long *total, int *cnt, void *hist_host)
{
  TSD_ENTER(f ? f->device : -1);
  if(!f) return fail("tsdgpu_poly_get_state: null handle");
  if(total) *total = f->total;
  if(cnt) *cnt = f->cnt;
  if(hist_host && f->L > 1)
  {
    TSD_CUDA(cudaStreamSynchronize(rt().stream));
    TSD_CUDA(cudaMemcpy(hist_host, f->d_hist[f->cur], (size_t) f->nchan * (f->L - 1) * sizeof(float) * f->DC, cudaMemcpyDeviceToHost));
  }
  return 0;
}

int tsdgpu_poly_set_state(tsdgpu_poly_t f, long long total, int cnt, const void *hist_host)
{
  TSD_ENTER(f ? f->device : -1);
  if(!f) return fail("tsdgpu_poly_set_state: null handle");
  if(total < 0 || cnt < 0 || cnt >= std::max(1, f->R)) return fail("tsdgpu_poly_set_state: invalid counters");
  TSD_CUDA(cudaStreamSynchronize(rt().stream));
  if(f->L > 1)
  {
    if(!hist_host) return fail("tsdgpu_poly_set_state: null history");
    TSD_CUDA(cudaMemcpy(f->d_hist[f->cur], hist_host, (size_t) f->nchan * (f->L - 1) * sizeof(float) * f->DC, cudaMemcpyHostToDevice));
  }
  f->total = total;
  f->cnt = cnt;
  return 0;
}

int tsdgpu_poly_destroy(tsdgpu_poly_t f)
{
  if(!f) return 0;
  TSD_ENTER(f->device);
  cudaStreamSynchronize(rt().stream);
  cudaFree(f->d_rows);
  cudaFree(f->d_hist[0]);
  cudaFree(f->d_hist[1]);
  delete f;
  return 0;
}

} // extern "C"
