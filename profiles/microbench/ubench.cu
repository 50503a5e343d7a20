// Micro-benchmarks that decide the design of the FFT-domain filter kernels on B200 (sm_100a):
//   l2      bandwidth of global loads / stores that hit L2 (16 MiB working set) vs. streaming from HBM
//   dsmem   st.shared::cluster / ld.shared::cluster throughput per SM for clusters of 2/4/8 CTAs
//   cbar    barrier.cluster arrive+wait round trip
//   fp      issue rate of FFMA vs. packed FFMA2 / FADD2
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; run on one B200.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while(0)

// ------------------------------------------------------------------ L2 / HBM bandwidth
__global__ void rd_kernel(const float4 *p, size_t n_vec, int reps, float4 *sink)
{
  float4 acc = make_float4(0, 0, 0, 0);
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for(int r = 0; r < reps; r++)
    for(size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride)
    {
      float4 v;
      asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  if(acc.x == 123.456f) *sink = acc;
}
__global__ void cp_kernel(const float4 *p, float4 *q, size_t n_vec, int reps)
{
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for(int r = 0; r < reps; r++)
    for(size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride)
    {
      float4 v;
      asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i));
      asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(q + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

static void bench_l2(int sms)
{
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float4 *buf, *buf2, *sink;
  const size_t big = (size_t) 4 << 30;
  CK(cudaMalloc(&buf, big)); CK(cudaMalloc(&buf2, big)); CK(cudaMalloc(&sink, 64));
  CK(cudaMemset(buf, 0, big)); CK(cudaMemset(buf2, 0, big));
  for(size_t mb : {8, 16, 32, 64, 4096})
  {
    const size_t bytes = mb << 20, nv = bytes / 16;
    const int reps = (int) ((size_t) 16384 / mb < 1 ? 1 : (size_t) 16384 / mb);
    for(int bpsm : {4, 8})
    {
      rd_kernel<<<sms * bpsm, 256>>>(buf, nv, 2, sink);
      CK(cudaEventRecord(a));
      rd_kernel<<<sms * bpsm, 256>>>(buf, nv, reps, sink);
      CK(cudaEventRecord(b));
      float ms = time_ms(a, b);
      printf("l2 read   ws=%5zu MiB ctas/sm=%d : %8.1f GB/s\n", mb, bpsm, (double) bytes * reps / ms / 1e6);
      cp_kernel<<<sms * bpsm, 256>>>(buf, buf2, nv, 2);
      CK(cudaEventRecord(a));
      cp_kernel<<<sms * bpsm, 256>>>(buf, buf2, nv, reps);
      CK(cudaEventRecord(b));
      ms = time_ms(a, b);
      printf("l2 copy   ws=2x%5zu MiB ctas/sm=%d : %8.1f GB/s (read+write)\n", mb, bpsm, 2.0 * bytes * reps / ms / 1e6);
    }
  }
  CK(cudaFree(buf)); CK(cudaFree(buf2)); CK(cudaFree(sink));
}

// ------------------------------------------------------------------ DSMEM
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank)
{
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// mode 0: remote st v2 (8 B / lane), 1: remote st v4, 2: remote ld v2, 3: remote ld v4, 4: local st v2 + ld v2 (reference)
template<int MODE> __global__ void dsmem_kernel(int iters, unsigned long long *cycles, float *sink)
{
  extern __shared__ float4 sm4[];
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank(), csz = cl.num_blocks();
  const int tid = threadIdx.x;
  const int slots_v2 = 4096;   // 32 KB window per CTA
  for(int i = tid; i < slots_v2 / 2; i += blockDim.x) sm4[i] = make_float4(1.f, 2.f, 3.f, 4.f);
  cl.sync();
  const unsigned base = smem_u32(sm4);
  float acc = 0.f;
  const unsigned long long t0 = clock64();
  for(int it = 0; it < iters; it++)
  {
#pragma unroll
    for(int j = 0; j < 16; j++)
    {
      const unsigned dst = (MODE == 4) ? rank : (rank + 1 + ((it * 16 + j) % (csz - 1))) % csz;
      if(MODE == 0 || MODE == 4)
      {
        const unsigned addr = mapa(base + (unsigned) (((j * blockDim.x + tid) % slots_v2) * 8), dst);
        asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(acc), "f"(1.0f) : "memory");
      }
      else if(MODE == 1)
      {
        const unsigned addr = mapa(base + (unsigned) (((j * blockDim.x + tid) % (slots_v2 / 2)) * 16), dst);
        asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(acc), "f"(1.0f), "f"(2.0f), "f"(3.0f) : "memory");
      }
      else if(MODE == 2)
      {
        const unsigned addr = mapa(base + (unsigned) (((j * blockDim.x + tid) % slots_v2) * 8), dst);
        float x, y;
        asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(addr));
        acc += x + y;
      }
      else
      {
        const unsigned addr = mapa(base + (unsigned) (((j * blockDim.x + tid) % (slots_v2 / 2)) * 16), dst);
        float x, y, z, w;
        asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(addr));
        acc += x + y + z + w;
      }
    }
  }
  cl.sync();
  const unsigned long long t1 = clock64();
  if(tid == 0) cycles[blockIdx.x] = t1 - t0;
  if(acc == 123.456f) *sink = acc;
}

template<int MODE> static void run_dsmem(int csz, int ctas_per_sm_target, int sms, const char *name, int bytes_per_lane)
{
  const int threads = 256, iters = 2000;
  const size_t smem = 32 * 1024;
  CK(cudaFuncSetAttribute(dsmem_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.attrs = at; cfg.numAttrs = 1;
  cfg.gridDim = dim3(csz);
  int maxc = 0;
  CK(cudaOccupancyMaxActiveClusters(&maxc, dsmem_kernel<MODE>, &cfg));
  int nclusters = std::min(maxc, sms * ctas_per_sm_target / csz);
  cfg.gridDim = dim3(nclusters * csz);
  unsigned long long *cyc; float *sink;
  CK(cudaMalloc(&cyc, sizeof(unsigned long long) * nclusters * csz)); CK(cudaMalloc(&sink, 4));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  CK(cudaLaunchKernelEx(&cfg, dsmem_kernel<MODE>, 10, cyc, sink));
  CK(cudaEventRecord(a));
  CK(cudaLaunchKernelEx(&cfg, dsmem_kernel<MODE>, iters, cyc, sink));
  CK(cudaEventRecord(b));
  const float ms = time_ms(a, b);
  std::vector<unsigned long long> h(nclusters * csz);
  CK(cudaMemcpy(h.data(), cyc, h.size() * 8, cudaMemcpyDeviceToHost));
  double mean = 0; for(auto v : h) mean += (double) v; mean /= h.size();
  const double bytes_cta = (double) iters * 16 * threads * bytes_per_lane;
  const double ctas_per_sm = (double) nclusters * csz / sms;
  printf("dsmem %-12s cluster=%d maxActiveClusters=%3d launched=%3d (%.2f CTA/SM): %6.2f B/cyc/CTA, %7.2f B/cyc/SM, chip %8.1f GB/s\n", name, csz,
         maxc, nclusters, ctas_per_sm, bytes_cta / mean, bytes_cta / mean * ctas_per_sm, bytes_cta * nclusters * csz / ms / 1e6);
  CK(cudaFree(cyc)); CK(cudaFree(sink));
}

// ------------------------------------------------------------------ cluster barrier latency
__global__ void cbar_kernel(int iters, unsigned long long *cycles)
{
  cg::cluster_group cl = cg::this_cluster();
  cl.sync();
  const unsigned long long t0 = clock64();
  for(int i = 0; i < iters; i++)
  {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  const unsigned long long t1 = clock64();
  if(threadIdx.x == 0) cycles[blockIdx.x] = (t1 - t0) / iters;
}
__global__ void bar_kernel(int iters, unsigned long long *cycles)
{
  const unsigned long long t0 = clock64();
  for(int i = 0; i < iters; i++) __syncthreads();
  const unsigned long long t1 = clock64();
  if(threadIdx.x == 0) cycles[blockIdx.x] = (t1 - t0) / iters;
}
static void run_cbar(int csz, int threads)
{
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(threads); cfg.attrs = at; cfg.numAttrs = 1; cfg.gridDim = dim3(csz * 16);
  unsigned long long *cyc; CK(cudaMalloc(&cyc, 8 * csz * 16));
  CK(cudaLaunchKernelEx(&cfg, cbar_kernel, 1000, cyc));
  unsigned long long h[128]; CK(cudaMemcpy(h, cyc, 8 * csz * 16, cudaMemcpyDeviceToHost));
  printf("cluster barrier arrive+wait: cluster=%d threads=%d : %llu cycles\n", csz, threads, h[0]);
  CK(cudaFree(cyc));
}

// ------------------------------------------------------------------ FP32 issue rate
template<int MODE> __global__ void fp_kernel(int iters, float *out, unsigned long long *cycles)
{
  // 8 independent chains per thread
  float2 r[8];
  for(int i = 0; i < 8; i++) r[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.001f, -0.001f);
  const unsigned long long t0 = clock64();
  for(int it = 0; it < iters; it++)
  {
#pragma unroll
    for(int i = 0; i < 8; i++)
    {
      if(MODE == 0)
      {
        r[i].x = fmaf(r[i].x, m.x, c.x);
        r[i].y = fmaf(r[i].y, m.y, c.y);
      }
      else if(MODE == 1)
      {
        unsigned long long d, a = *reinterpret_cast<unsigned long long *>(&r[i]);
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(*reinterpret_cast<const unsigned long long *>(&m)), "l"(*reinterpret_cast<const unsigned long long *>(&c)));
        r[i] = *reinterpret_cast<float2 *>(&d);
      }
      else if(MODE == 2)
      {
        unsigned long long d, a = *reinterpret_cast<unsigned long long *>(&r[i]);
        asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(*reinterpret_cast<const unsigned long long *>(&c)));
        r[i] = *reinterpret_cast<float2 *>(&d);
      }
      else
      {
        r[i].x = r[i].x + c.x;
        r[i].y = r[i].y + c.y;
      }
    }
  }
  const unsigned long long t1 = clock64();
  float s = 0;
  for(int i = 0; i < 8; i++) s += r[i].x + r[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if(threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template<int MODE> static void run_fp(const char *name, int sms)
{
  const int iters = 4096, threads = 512;
  float *out; unsigned long long *cyc;
  CK(cudaMalloc(&out, 4 * threads * sms * 2)); CK(cudaMalloc(&cyc, 8 * sms * 2));
  fp_kernel<MODE><<<sms * 2, threads>>>(iters, out, cyc);
  fp_kernel<MODE><<<sms * 2, threads>>>(iters, out, cyc);
  unsigned long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  // per SM: 2 CTAs x 512 threads; scalar-float results per thread-iteration = 16
  const double flops_lane = 16.0 * iters;   // float results per thread
  printf("fp %-8s: %6.2f float results / cycle / SM (1024 threads/SM)\n", name, flops_lane * 1024 / (double) h);
  CK(cudaFree(out)); CK(cudaFree(cyc));
}

int main()
{
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, smem/SM %zu, smem/block optin %zu\n", prop.name, sms, prop.sharedMemPerMultiprocessor, prop.sharedMemPerBlockOptin);
  run_fp<0>("FFMA", sms);
  run_fp<1>("FFMA2", sms);
  run_fp<2>("FADD2", sms);
  run_fp<3>("FADD", sms);
  for(int csz : {2, 4, 8}) run_cbar(csz, 256);
  run_cbar(8, 512);
  for(int csz : {2, 4, 8})
  {
    for(int cps : {1, 2, 4})
    {
      run_dsmem<0>(csz, cps, sms, "st.v2 remote", 8);
      run_dsmem<1>(csz, cps, sms, "st.v4 remote", 16);
      run_dsmem<2>(csz, cps, sms, "ld.v2 remote", 8);
      run_dsmem<3>(csz, cps, sms, "ld.v4 remote", 16);
    }
  }
  run_dsmem<4>(2, 4, sms, "st.v2 local", 8);
  bench_l2(sms);
  return 0;
}
