// Decides, by measurement, whether the radix-16 passes of the FFT-domain filter belong on the tensor cores
// (north_star: "tensor cores only if a 3xTF32 ... formulation is shown by ncu to beat FP32 FMA at fp32 accuracy";
// VERDICT r1, item 3).  Two kernels run the SAME unit of work, one radix-16 DFT pass over complex points that are
// resident on the SM, followed by the inter-pass twiddle multiplication, and nothing else (no global or shared-memory
// data traffic: both formulations would pay the same exchanges):
//
//   tc    the pass as a GEMM on tcgen05, exactly the recipe of the verdict: the data rows are M (128 independent
//         sub-transforms per MMA), the real 32 x 32 DFT-16 matrix is B (K = 32, N = 32; hi and lo tf32 parts resident in
//         shared memory, K-major, 128-byte swizzle), the data is the A operand in tensor memory.  Per tile of 128 x 16
//         complex points and per pass: split 32 floats into tf32 hi / lo (integer rounding, as in fir_tc.cu), tcgen05.st
//         (64 columns), 4 K-steps x 3 MMAs (hi*lo + lo*hi + hi*hi) of M128 N32 K8, tcgen05.ld of the 32 result columns,
//         twiddle multiplication on the CUDA cores.  Two groups of 4 warps alternate (two A stages, two accumulators)
//         so that the MMAs of one tile overlap the split / load of the other, one MMA-issuing warp.
//   simt  the pass as the register butterfly the library uses (fft16 of fft_tiles.cuh: packed FADD2 / FFMA2), followed
//         by the same twiddle multiplication; 768 threads per SM, 16 points per thread.
//
// Output: complex points per SM clock for both, the tensor-path error against a float64 DFT, and (under ncu) the pipe
// utilisations.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I../../libtsd_b200/csrc -o dft16_tc dft16_tc.cu
#include "fft_tiles.cuh"
#include "tc_common.cuh"

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace tsdgpu;
using namespace tsdgpu::tc;

#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while(0)

constexpr int NGROUP = 2;                 // worker groups (4 warps each: one warp per TMEM lane quadrant)
constexpr int MMA_WARP = 4 * NGROUP;
constexpr int NTHREADS = 32 * (MMA_WARP + 1);
constexpr int B_BYTES = 32 * 128;         // 32 rows (N) x 32 tf32 (K), 128 bytes per row: one swizzle atom group of 4 x 8 rows
// TMEM columns: accumulator of group g at [32 g, +32); A stage of group g at [64 + 64 g, +64) = hi | lo
constexpr int TMEM_COLS = 256;

// accumulate = 0 overwrites D
__device__ __forceinline__ void mma_tf32_acc(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "setp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
    ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
    : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32])
{
  uint32_t u[32];
  asm volatile(
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
    "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]),
      "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]),
      "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]),
      "=r"(u[31])
    : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for(int i = 0; i < 32; i++) r[i] = __uint_as_float(u[i]);
}

// One thread = one row of 16 complex points: v[2 j] = re, v[2 j + 1] = im.  out[2 k + c] = sum_{j,c'} v[2 j + c'] * Bm[2 k + c][2 j + c'],
// Bm = the real form of W16^(jk).  `tw` = 16 complex twiddles applied after every pass (|tw| = 1, so the data stay bounded).
__global__ void __launch_bounds__(NTHREADS, 1) dft16_tc_kernel(const float *Bhi_g, const float *Blo_g, const float2 *tw_g, float *io, int iters,
                                                               long long *cycles)
{
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  unsigned char *sm = raw + (base - smem_u32(raw));
  uint64_t *bars = reinterpret_cast<uint64_t *>(sm + 2 * B_BYTES);
  uint64_t *full = bars, *done = bars + NGROUP;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * NGROUP);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if(tid == 0)
  {
    for(int g = 0; g < NGROUP; g++) { mbar_init(full + g, 4); mbar_init(done + g, 1); }
    mbar_fence_init();
  }
  if(warp == MMA_WARP)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B matrix, hi and lo parts: row n (output component), 32 K values, 128-byte rows, swizzled
  for(int idx = tid; idx < 32 * 32; idx += NTHREADS)
  {
    const int row = idx >> 5, kk = idx & 31;
    const uint32_t off = swz((uint32_t) (row * 128 + kk * 4));
    *reinterpret_cast<float *>(sm + off) = Bhi_g[idx];
    *reinterpret_cast<float *>(sm + B_BYTES + off) = Blo_g[idx];
  }
  fence_proxy_async();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  const long long t0 = clock64();

  if(warp < MMA_WARP)
  {
    const int grp = warp >> 2, pw = warp & 3;
    const uint32_t lane_base = (uint32_t) (pw * 32) << 16;
    const uint32_t my_a = tmem + lane_base + (uint32_t) (64 + 64 * grp), my_d = tmem + lane_base + (uint32_t) (32 * grp);
    float v[32];
    const size_t row = ((size_t) blockIdx.x * NGROUP + grp) * 128 + pw * 32 + lane;
#pragma unroll
    for(int i = 0; i < 32; i++) v[i] = io[row * 32 + i];
    float2 tw[16];
#pragma unroll
    for(int k = 0; k < 16; k++) tw[k] = tw_g[k];
    for(int it = 0; it < iters; it++)
    {
      // split -> A stage
#pragma unroll
      for(int hq = 0; hq < 2; hq++)
      {
        float hi[16], lo[16];
#pragma unroll
        for(int m = 0; m < 16; m++)
        {
          hi[m] = to_tf32(v[16 * hq + m]);
          lo[m] = to_tf32(v[16 * hq + m] - hi[m]);
        }
        tmem_st16(my_a + hq * 16, hi);
        tmem_st16(my_a + 32 + hq * 16, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      fence_before();
      __syncwarp();
      if(lane == 0) mbar_arrive(full + grp);
      mbar_wait(done + grp, (unsigned) (it & 1));
      fence_after();
      tmem_ld32(my_d, v);
      // inter-pass twiddle on the CUDA cores
#pragma unroll
      for(int k = 0; k < 16; k++)
      {
        const float re = v[2 * k] * tw[k].x - v[2 * k + 1] * tw[k].y, im = v[2 * k] * tw[k].y + v[2 * k + 1] * tw[k].x;
        v[2 * k] = re;
        v[2 * k + 1] = im;
      }
    }
#pragma unroll
    for(int i = 0; i < 32; i++) io[row * 32 + i] = v[i];
  }
  else
  {
    const uint64_t dbase = smem_desc(0);
    const uint64_t bhi = dbase + (base >> 4), blo = dbase + ((base + B_BYTES) >> 4);
    const uint32_t idesc = IDESC_M128 | ((uint32_t) (32 >> 3) << 17);
    for(int it = 0; it < iters; it++)
      for(int grp = 0; grp < NGROUP; grp++)
      {
        mbar_wait(full + grp, (unsigned) (it & 1));
        fence_after();
        const uint32_t ah = tmem + (uint32_t) (64 + 64 * grp), al = ah + 32, d = tmem + (uint32_t) (32 * grp);
        if(elect_one())
        {
#pragma unroll
          for(int ks = 0; ks < 4; ks++)
          {
            mma_tf32_acc(d, ah + 8 * ks, blo + 2 * ks, idesc, ks ? 1u : 0u);
            mma_tf32_acc(d, al + 8 * ks, bhi + 2 * ks, idesc, 1u);
            mma_tf32_acc(d, ah + 8 * ks, bhi + 2 * ks, idesc, 1u);
          }
          mma_commit(done + grp);
        }
        __syncwarp();
      }
  }
  fence_before();
  __syncthreads();
  if(tid == 0) cycles[blockIdx.x] = clock64() - t0;
  if(warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

// SIMT: the library's register butterfly + the same twiddles
__global__ void __launch_bounds__(256, 3) dft16_simt_kernel(const float2 *tw_g, float2 *io, int iters, long long *cycles)
{
  const size_t row = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  float2 v[16], tw[16];
#pragma unroll
  for(int i = 0; i < 16; i++) { v[i] = io[row * 16 + i]; tw[i] = tw_g[i]; }
  __syncthreads();
  const long long t0 = clock64();
  for(int it = 0; it < iters; it++)
  {
    fft16<false>(v);
#pragma unroll
    for(int k = 0; k < 16; k++) v[k] = cmul(v[k], make_float2(0.25f * tw[k].x, 0.25f * tw[k].y));   // 1/4: keeps the magnitude (the GEMM form folds it into B)
  }
  __syncthreads();
  if(threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
#pragma unroll
  for(int i = 0; i < 16; i++) io[row * 16 + i] = v[i];
}

static float tf32_round(float v)
{
  uint32_t u;
  memcpy(&u, &v, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&v, &u, 4);
  return v;
}

int main(int argc, char **argv)
{
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  // real form of the unitary-free DFT-16: out[2k+c] = sum in[2j+c'] * Bm[2k+c][2j+c'], scaled by 1/4 per pass so that repeated
  // passes keep the magnitude (a power of two: exact)
  std::vector<float> Bhi(32 * 32), Blo(32 * 32);
  for(int k = 0; k < 16; k++)
    for(int j = 0; j < 16; j++)
    {
      const double a = -2.0 * M_PI * (double) (j * k % 16) / 16.0, c = cos(a) * 0.25, s = sin(a) * 0.25;
      const double m[2][2] = {{c, -s}, {s, c}};   // (re, im) out <- (re, im) in
      for(int co = 0; co < 2; co++)
        for(int ci = 0; ci < 2; ci++)
        {
          const float f = (float) m[co][ci], hi = tf32_round(f);
          Bhi[(2 * k + co) * 32 + 2 * j + ci] = hi;
          Blo[(2 * k + co) * 32 + 2 * j + ci] = tf32_round(f - hi);
        }
    }
  std::vector<float2> tw(16);
  for(int k = 0; k < 16; k++) tw[k] = make_float2((float) cos(-2 * M_PI * k * 3 / 256.0), (float) sin(-2 * M_PI * k * 3 / 256.0));
  float *dBhi, *dBlo, *dio;
  float2 *dtw;
  long long *dcyc;
  const size_t rows = (size_t) sms * NGROUP * 128;
  std::vector<float> h(rows * 32);
  srand(1);
  for(auto &x : h) x = (float) rand() / RAND_MAX - 0.5f;
  CK(cudaMalloc(&dBhi, Bhi.size() * 4)); CK(cudaMalloc(&dBlo, Blo.size() * 4)); CK(cudaMalloc(&dtw, 16 * 8));
  CK(cudaMalloc(&dio, h.size() * 4)); CK(cudaMalloc(&dcyc, 4096 * 8));
  CK(cudaMemcpy(dBhi, Bhi.data(), Bhi.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBlo, Blo.data(), Blo.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dtw, tw.data(), 16 * 8, cudaMemcpyHostToDevice));
  const int smem = 2 * B_BYTES + 1024 + 256;
  CK(cudaFuncSetAttribute(dft16_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

  // ---- accuracy of ONE tensor-core pass against float64
  CK(cudaMemcpy(dio, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  dft16_tc_kernel<<<sms, NTHREADS, smem>>>(dBhi, dBlo, dtw, dio, 1, dcyc);
  CK(cudaDeviceSynchronize());
  std::vector<float> out(h.size());
  CK(cudaMemcpy(out.data(), dio, h.size() * 4, cudaMemcpyDeviceToHost));
  double emax = 0, rms = 0;
  for(size_t r = 0; r < 512; r++)
    for(int k = 0; k < 16; k++)
    {
      std::complex<double> s = 0;
      for(int j = 0; j < 16; j++) s += std::complex<double>(h[r * 32 + 2 * j], h[r * 32 + 2 * j + 1]) * std::polar(0.25, -2.0 * M_PI * (j * k % 16) / 16.0);
      s *= std::complex<double>(tw[k].x, tw[k].y);
      const std::complex<double> g(out[r * 32 + 2 * k], out[r * 32 + 2 * k + 1]);
      emax = std::max(emax, std::abs(g - s));
      rms += std::norm(s);
    }
  rms = sqrt(rms / (512 * 16));
  printf("tensor-core pass (3xTF32): max error %.3e of the output RMS (%.3e abs)\n", emax / rms, emax);

  // ---- throughput
  for(int rep = 0; rep < 2; rep++)
  {
    CK(cudaMemcpy(dio, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaEventRecord(e0));
    dft16_tc_kernel<<<sms, NTHREADS, smem>>>(dBhi, dBlo, dtw, dio, iters, dcyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long cyc;
    CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    const double pts = (double) NGROUP * 128 * 16 * iters;   // complex points per SM
    if(rep) printf("tc   : %.3f ms, %lld cycles per CTA (1 CTA per SM) -> %.2f complex points per SM clock, %.1f G pass-points/s chip-wide\n", ms, cyc,
                   pts / cyc, pts * sms / (ms * 1e-3) / 1e9);
  }
  float2 *dio2;
  const size_t rows2 = (size_t) sms * 3 * 256;
  CK(cudaMalloc(&dio2, rows2 * 16 * 8));
  CK(cudaMemset(dio2, 0, rows2 * 16 * 8));
  for(int rep = 0; rep < 2; rep++)
  {
    CK(cudaEventRecord(e0));
    dft16_simt_kernel<<<sms * 3, 256>>>(dtw, dio2, iters, dcyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long cyc;
    CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    const double pts = 3.0 * 256 * 16 * iters;
    if(rep) printf("simt : %.3f ms, %lld cycles per CTA (3 CTAs per SM) -> %.2f complex points per SM clock, %.1f G pass-points/s chip-wide\n", ms, cyc,
                   pts / cyc, pts * sms / (ms * 1e-3) / 1e9);
  }
  printf("(SM clock attribute %.0f MHz, %d SMs, %d passes)\n", clk_khz / 1000.0, sms, iters);
  return 0;
}
