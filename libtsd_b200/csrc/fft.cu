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
#include "tsdgpu.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

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
      const unsigned n2 = (unsigned) (16 * g + lo);
      mul_geometric(v, twiddle<INV>(n2 * (unsigned) hi, 2.0f / 65536.0f), twiddle<INV>(16u * n2, 2.0f / 65536.0f));
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
    __syncthreads();
    fft256_cols<INV, true>(v, sm, tw, hi, lo);
    const unsigned n2 = (unsigned) (16 * g + lo);
    mul_geometric(v, twiddle<INV>(n2 * (unsigned) hi, 2.0f / 65536.0f), twiddle<INV>(16u * n2, 2.0f / 65536.0f));
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
  if(n <= 0 || (n & (n - 1))) return fail("tsdgpu_fft_plan: n must be a power of two >= 1 in this version");
  if(batch <= 0) return fail("tsdgpu_fft_plan: batch must be > 0");
  if(n > (1 << 24)) return fail("tsdgpu_fft_plan: n > 2^24 not supported");
  auto *p = new tsdgpu_fft_s;
  p->n = n;
  p->batch = batch;
  if(n == 65536)
  {
    p->ring = 64;
    p->lag = 32;
    p->staged = 1;
    if(const char *v = getenv("TSDGPU_FFT_MODE")) p->staged = v[0] == 'p' ? 0 : 1;
    if(const char *v = getenv("TSDGPU_FFT_CHUNK")) p->chunk = std::max(1, atoi(v));
    if(const char *v = getenv("TSDGPU_FFT_STREAMS")) p->nstreams = std::min((int) Runtime::MAX_AUX, std::max(1, atoi(v)));
    if(p->staged && aux_init()) { fft_plan_destroy(p); return 1; }
    if(const char *v = getenv("TSDGPU_FFT_LAG")) p->lag = std::max(1, atoi(v));
    if(const char *v = getenv("TSDGPU_FFT_RING")) p->ring = atoi(v);
    if(p->ring <= p->lag) p->ring = p->lag + 16;
    const size_t slots = p->staged ? (size_t) p->chunk * p->nstreams : (size_t) p->ring;
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
  if(p->scratch) cudaFree(p->scratch);
  if(p->flags) cudaFree(p->flags);
  if(p->work[0]) cudaFree(p->work[0]);
  if(p->work[1]) cudaFree(p->work[1]);
  delete p;
}

static int ensure_work(tsdgpu_fft_s *p, int which)
{
  if(p->work[which]) return 0;
  TSD_CUDA(cudaMalloc(&p->work[which], (size_t) p->n * p->batch * sizeof(float2)));
  return 0;
}

int fft_exec_device(tsdgpu_fft_s *p, const float2 *x, long long xs, float2 *y, long long ys, bool forward)
{
  Runtime &r = rt();
  const int N = p->n, batch = p->batch;
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
  if(ensure_init()) return 1;
  if(!out) return fail("tsdgpu_fft_plan: null argument");
  return fft_plan_create(n, batch, out);
}

int tsdgpu_fft_exec(tsdgpu_fft_t p, const void *x, long long xs, void *y, long long ys, int forward, int mem)
{
  if(ensure_init()) return 1;
  if(!p || !x || !y) return fail("tsdgpu_fft_exec: null argument");
  if(xs < p->n || ys < p->n) return fail("tsdgpu_fft_exec: stride smaller than n");
  if(mem == TSDGPU_DEVICE) return fft_exec_device(p, (const float2 *) x, xs, (float2 *) y, ys, forward != 0);
  float2 *d = nullptr;
  const size_t row = (size_t) p->n * sizeof(float2);
  TSD_CUDA(cudaMalloc(&d, row * p->batch));
  int rc = 0;
  cudaError_t e = cudaMemcpy2DAsync(d, row, x, (size_t) xs * 8, row, p->batch, cudaMemcpyHostToDevice, rt().stream);
  if(e == cudaSuccess) rc = fft_exec_device(p, d, p->n, d, p->n, forward != 0);
  if(e == cudaSuccess && !rc)
    e = cudaMemcpy2DAsync(y, (size_t) ys * 8, d, row, row, p->batch, cudaMemcpyDeviceToHost, rt().stream);
  if(e == cudaSuccess) e = cudaStreamSynchronize(rt().stream);
  cudaFree(d);
  if(e != cudaSuccess) return fail(std::string("tsdgpu_fft_exec: ") + cudaGetErrorString(e));
  return rc;
}

int tsdgpu_fft_destroy(tsdgpu_fft_t p)
{
  if(p) cudaStreamSynchronize(rt().stream);
  fft_plan_destroy(p);
  return 0;
}

} // extern "C"
