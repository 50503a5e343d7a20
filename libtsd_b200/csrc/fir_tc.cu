// Direct FIR as a banded Toeplitz GEMM on the 5th-generation tensor cores (tcgen05 / TMEM), 3xTF32.
//
// Replaces the inner product of FiltreRIF<cfloat,float>::step (reference filtre-rt.cc:82-107) for K <= 127 real
// taps on cf32 data.  For a tile of 128 consecutive outputs t and a chunk of 32 consecutive inputs c,
//   y[128 t + j] += sum_kk h[(128 t + j) - (32 c + kk)] * x[32 c + kk]
// is D[j][n] += A[j][kk] * B[n][kk] with D in tensor memory (128 lanes x 128 columns, n = 2*channel + re/im of
// 64 channels), A[j][kk] = h[j - kk + 32 d], d = 4 t - c in [-3, 4], and B the de-interleaved input chunk.
// All eight A blocks are row-shifted views G[j + 32 d][kk] of ONE generator matrix G[r][kk] = h[r - kk]
// (352 rows x 32 columns), so the Toeplitz operand costs 44 KiB of shared memory per split part instead of 128 KiB.
// fp32 accuracy from tf32 inputs: x = x_hi + x_lo, h = h_hi + h_lo (each part rounded to tf32), three MMAs
// h_hi*x_hi + h_lo*x_hi + h_hi*x_lo accumulated in fp32 by the tensor core (error ~2^-21 per product).
//
// One CTA (416 threads, 1 per SM) = 64 channels x `span` tiles, input-stationary: every input chunk is loaded,
// split and stored to shared memory ONCE (K-major, 128-byte swizzle, the canonical UMMA layout) and feeds the
// two output tiles it overlaps, whose accumulators are live in TMEM at the same time (3 regions of 128 columns:
// two accumulating, one being drained).  Warp roles: warps 0-3 epilogue (tcgen05.ld -> coalesced float2 stores),
// warps 4-11 producers (two groups alternating chunks: LDG.128 one chunk ahead -> cvt.rna.tf32 split -> STS, 3-stage
// ring, mbarrier full/empty), warp 12 lane 0 issues the MMAs (24 per chunk: 2 tiles x 4 K-steps of 8 x 3 split terms) and the commits.
// Per chunk of 2048 complex samples: 24 MMAs x 64 cycles = 1536 cycles  =>  tensor bound 0.75 cycle per sample
// per SM (~375 Gsamples/s at 1.9 GHz) against 16 B/sample of HBM traffic (410 Gsamples/s): HBM / tensor balanced,
// where the FP32 FMA formulation (fir.cu) is capped at 146 Gsamples/s.
#include "common.cuh"
#include "fir_tc.h"

#include <cstdlib>

namespace tsdgpu {
namespace tc {

constexpr int TILE = 128;            // outputs per tile = UMMA M
constexpr int CH = 64;               // channels per CTA; UMMA N = 2 * CH
constexpr int NCOL = 2 * CH;
constexpr int CHUNK = 32;            // input samples per chunk = 4 UMMA K steps of 8 tf32
constexpr int NSTAGE = 3;
constexpr int GROWS = 352;           // generator rows r in [-96, 256)
constexpr int G_BYTES = GROWS * 128; // per split part (multiple of 1024)
constexpr int PART_BYTES = NCOL * 128;          // one split part of one chunk: 128 rows x 32 tf32
constexpr int STAGE_BYTES = 2 * PART_BYTES;     // hi + lo
constexpr int SMEM_BYTES = 2 * G_BYTES + NSTAGE * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr int NGROUP = 2;              // producer groups (4 warps each)
constexpr int MMA_WARP = 4 + 4 * NGROUP;
constexpr int NTHREADS = 32 * (MMA_WARP + 1);
constexpr int TMEM_COLS = 512;

__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }   // Swizzle<3,4,3>
__device__ __forceinline__ float to_tf32(float v)
{
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr)
{
  return (uint64_t) ((saddr >> 4) & 0x3FFFu) | ((uint64_t) 1 << 16) /* LBO (unused for swizzled K-major) */ |
         ((uint64_t) (1024 >> 4) << 32) /* SBO */ | ((uint64_t) 1 << 46) /* version */ | ((uint64_t) 2 << 61) /* SWIZZLE_128B */;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, both K-major, M = 128, N = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (NCOL >> 3) << 17) | ((uint32_t) (TILE >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate)
{
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "setp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
    ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
    : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__global__ void __launch_bounds__(NTHREADS, 1) fir_tc_kernel(FirTcParams p)
{
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  unsigned char *sm = raw + (base - smem_u32(raw));
  float *Ghi = reinterpret_cast<float *>(sm), *Glo = reinterpret_cast<float *>(sm + G_BYTES);
  unsigned char *stages = sm + 2 * G_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(stages + NSTAGE * STAGE_BYTES);
  uint64_t *full = bars, *empty = bars + NSTAGE, *tfull = bars + 2 * NSTAGE, *tempty = bars + 2 * NSTAGE + 3;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * NSTAGE + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ts = blockIdx.x * p.span, te = min(ts + p.span, p.ntiles);   // tiles [ts, te) of this CTA
  const int c0 = blockIdx.y * CH;                                          // first channel
  const int nchunks = 4 * (te - ts) + 4;                                   // chunks 4 ts - 4 ... 4 te - 1

  // ---- one-time set-up: barriers, tensor memory, generator matrix (hi / lo split, swizzled K-major)
  if(tid == 0)
  {
    for(int i = 0; i < NSTAGE; i++) { mbar_init(full + i, 4); mbar_init(empty + i, 1); }
    for(int i = 0; i < 3; i++) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 4); }
    mbar_fence_init();
  }
  if(warp == MMA_WARP)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for(int idx = tid; idx < GROWS * 32; idx += NTHREADS)
  {
    const int row = idx >> 5, kk = idx & 31, tap = (row - 96) - kk;
    const float v = (tap >= 0 && tap < p.K) ? __ldg(p.taps_rev + (p.K - 1 - tap)) : 0.f;
    const float hi = to_tf32(v), lo = to_tf32(v - hi);
    const uint32_t off = swz((uint32_t) (row * 128 + kk * 4));
    *reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(Ghi) + off) = hi;
    *reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(Glo) + off) = lo;
  }
  fence_proxy_async();   // generic-proxy writes of G -> visible to the tensor core (async proxy)
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;

  if(warp >= 4 && warp < 4 + 4 * NGROUP)
  {
    // ===== producers: NGROUP groups of 4 warps, group g takes chunks it = g, g + NGROUP, ...  Chunk it covers inputs
    // [32 c, 32 c + 32), c = 4 ts - 4 + it, of 64 channels.  The global loads of a group's NEXT chunk are issued
    // before the current one is converted (they do not depend on the ring), so that NGROUP + ... chunks are in flight.
    const int pw = (warp - 4) & 3, grp = (warp - 4) >> 2;
    const int sp = lane & 15, half = lane >> 4;
    auto load_chunk = [&](int it, float4 (&v)[8]) {
      const long long pos = (long long) (4 * ts - 4 + it) * CHUNK + 2 * sp;   // first of this lane's two samples
#pragma unroll
      for(int i = 0; i < 8; i++)
      {
        const int chan = c0 + pw * 16 + i * 2 + half;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if(chan < p.nchan)
        {
          if(pos >= 0)
          {
            const float2 *src = p.x + (long long) chan * p.x_stride + pos;
            if(pos + 1 < p.n) v[i] = __ldcs(reinterpret_cast<const float4 *>(src));
            else if(pos < p.n) { const float2 a = __ldcs(src); v[i] = make_float4(a.x, a.y, 0.f, 0.f); }
          }
          else if(pos >= -(long long) p.halo) v[i] = __ldg(reinterpret_cast<const float4 *>(p.hist + (long long) chan * p.halo + p.halo + pos));
        }
      }
    };
    auto store_chunk = [&](int it, const float4 (&v)[8]) {
      const int stage = it % NSTAGE;
      mbar_wait(empty + stage, (unsigned) (((it / NSTAGE) & 1) ^ 1));
      unsigned char *bhi = stages + stage * STAGE_BYTES, *blo = bhi + PART_BYTES;
#pragma unroll
      for(int i = 0; i < 8; i++)
      {
        const int cl = pw * 16 + i * 2 + half;            // local channel: rows 2 cl (re) and 2 cl + 1 (im)
        const float4 x = v[i];                            // (re0, im0, re1, im1)
        const float r0 = to_tf32(x.x), i0 = to_tf32(x.y), r1 = to_tf32(x.z), i1 = to_tf32(x.w);
        const uint32_t ore = swz((uint32_t) ((2 * cl) * 128 + sp * 8)), oim = swz((uint32_t) ((2 * cl + 1) * 128 + sp * 8));
        *reinterpret_cast<float2 *>(bhi + ore) = make_float2(r0, r1);
        *reinterpret_cast<float2 *>(bhi + oim) = make_float2(i0, i1);
        *reinterpret_cast<float2 *>(blo + ore) = make_float2(to_tf32(x.x - r0), to_tf32(x.z - r1));
        *reinterpret_cast<float2 *>(blo + oim) = make_float2(to_tf32(x.y - i0), to_tf32(x.w - i1));
      }
      fence_proxy_async();
      __syncwarp();
      if(lane == 0) mbar_arrive(full + stage);
    };
    // three register sets per thread: the loads of a group's next TWO chunks are in flight while one is converted
    float4 va[8], vb[8], vc[8];
    const int G = NGROUP;
    int it = grp;
    if(it < nchunks) load_chunk(it, va);
    if(it + G < nchunks) load_chunk(it + G, vb);
    for(; it < nchunks; it += 3 * G)
    {
      if(it + 2 * G < nchunks) load_chunk(it + 2 * G, vc);
      store_chunk(it, va);
      if(it + G >= nchunks) break;
      if(it + 3 * G < nchunks) load_chunk(it + 3 * G, va);
      store_chunk(it + G, vb);
      if(it + 2 * G >= nchunks) break;
      if(it + 4 * G < nchunks) load_chunk(it + 4 * G, vb);
      store_chunk(it + 2 * G, vc);
    }
  }
  else if(warp == MMA_WARP)
  {
    // ===== MMA issuer: the whole warp runs the loop on warp-uniform values (so that descriptors live in uniform
    // registers) and one elected lane issues the 24 MMAs + commits of a chunk in a single block.  The issue loop
    // is the critical path of the kernel: keep it free of per-MMA address arithmetic and branches.
    const uint32_t ghi = base, glo = base + G_BYTES, st0 = base + 2 * G_BYTES;
    const uint64_t dbase = smem_desc(0);
    const int ntl = te - ts;
    for(int it = 0; it < nchunks; it++)
    {
      const int stage = it % NSTAGE;
      mbar_wait(full + stage, (unsigned) ((it / NSTAGE) & 1));
      // chunk c = 4 ts - 4 + it feeds tiles ts + (it / 4) - 1 (d = 4 t - c = -(it % 4)) and ts + it / 4 (d = 4 - it % 4)
      const int q4 = it >> 2, r4 = it & 3;
      const int tl0 = q4 - 1, tl1 = q4;
      const bool on0 = tl0 >= 0 && tl0 < ntl, on1 = tl1 < ntl;
      if(on1 && r4 == 0) mbar_wait(tempty + tl1 % 3, (unsigned) (((tl1 / 3) & 1) ^ 1));   // accumulator region drained
      fence_after();
      const uint32_t bhi = st0 + stage * STAGE_BYTES;
      const uint64_t bh0 = dbase + (bhi >> 4), bl0 = bh0 + (PART_BYTES >> 4);
      const uint32_t arow0 = (uint32_t) ((96 - 32 * r4) * 128), arow1 = (uint32_t) ((224 - 32 * r4) * 128);
      const uint64_t ah0 = dbase + ((ghi + arow0) >> 4), al0 = dbase + ((glo + arow0) >> 4);
      const uint64_t ah1 = dbase + ((ghi + arow1) >> 4), al1 = dbase + ((glo + arow1) >> 4);
      const uint32_t d0 = tmem + (uint32_t) ((tl0 + 3) % 3 * NCOL), d1 = tmem + (uint32_t) (tl1 % 3 * NCOL);
      if(elect_one())
      {
        if(on0)
        {
#pragma unroll
          for(int ks = 0; ks < 4; ks++)
          {
            mma_tf32(d0, al0 + 2 * ks, bh0 + 2 * ks, 1u);
            mma_tf32(d0, ah0 + 2 * ks, bl0 + 2 * ks, 1u);
            mma_tf32(d0, ah0 + 2 * ks, bh0 + 2 * ks, 1u);
          }
          if(r4 == 3) mma_commit(tfull + tl0 % 3);          // d = -3: last chunk of tile tl0, accumulator complete
        }
        if(on1)
        {
#pragma unroll
          for(int ks = 0; ks < 4; ks++)
          {
            mma_tf32(d1, al1 + 2 * ks, bh0 + 2 * ks, (r4 == 0 && ks == 0) ? 0u : 1u);   // d = 4: first chunk of tile tl1
            mma_tf32(d1, ah1 + 2 * ks, bl0 + 2 * ks, 1u);
            mma_tf32(d1, ah1 + 2 * ks, bh0 + 2 * ks, 1u);
          }
        }
        mma_commit(empty + stage);                          // the stage may be refilled once these MMAs have read it
      }
      __syncwarp();
    }
  }
  else
  {
    // ===== epilogue: warp w owns TMEM lanes (= output rows) 32 w ... 32 w + 31
    for(int tl = 0; tl < te - ts; tl++)
    {
      const int region = tl % 3;
      mbar_wait(tfull + region, (unsigned) ((tl / 3) & 1));
      fence_after();
      const long long nabs = (long long) (ts + tl) * TILE + warp * 32 + lane;
#pragma unroll
      for(int q = 0; q < 4; q++)
      {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t) (warp * 32) << 16) + (uint32_t) (region * NCOL + q * 32);
        asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
            "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
            "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr)
          : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if(nabs < p.n)
        {
#pragma unroll
          for(int cp = 0; cp < 16; cp++)
          {
            const int chan = c0 + q * 16 + cp;
            if(chan < p.nchan)
              stg_stream(p.y + (long long) chan * p.y_stride + nabs, make_float2(__uint_as_float(r[2 * cp]), __uint_as_float(r[2 * cp + 1])));
          }
        }
      }
      fence_before();
      __syncwarp();
      if(lane == 0) mbar_arrive(tempty + region);
    }
  }
  // ---- teardown
  fence_before();
  __syncthreads();
  if(warp == MMA_WARP)
  {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

} // namespace tc

bool fir_tc_eligible(int kind_cf32_f32, int K, const void *x, long long x_stride, const void *hist, int halo)
{
  return kind_cf32_f32 && K >= 1 && K <= 127 && (((uintptr_t) x & 15) == 0) && (x_stride % 2 == 0) && (((uintptr_t) hist & 15) == 0) &&
         (halo % 2 == 0);
}

int fir_tc_launch(const FirTcParams &p0)
{
  FirTcParams p = p0;
  Runtime &r = rt();
  static bool attr_set = false;
  if(!attr_set)
  {
    TSD_CUDA(cudaFuncSetAttribute(tc::fir_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    attr_set = true;
  }
  p.ntiles = (p.n + tc::TILE - 1) / tc::TILE;
  const int groups = (p.nchan + tc::CH - 1) / tc::CH;
  // tiles per CTA: long spans amortise the 4 halo chunks and the set-up, short spans balance the 148 SMs
  // every CTA pays one extra tile's worth of halo chunks; pick the span in [4, 32] that minimises waves x (span + 1)
  int span = 1;
  long long best = -1;
  for(int sgs = 1; sgs <= 32; sgs++)
  {
    if(sgs < 4 && p.ntiles > 4) continue;
    const long long ctas = (long long) groups * ((p.ntiles + sgs - 1) / sgs);
    const long long cost = ((ctas + r.num_sms - 1) / r.num_sms) * (sgs + 1);
    if(best < 0 || cost < best) { best = cost; span = sgs; }
  }
  p.span = span;
  dim3 grid((p.ntiles + span - 1) / span, groups);
  tc::fir_tc_kernel<<<grid, tc::NTHREADS, tc::SMEM_BYTES, r.stream>>>(p);
  TSD_LAUNCH_CHECK();
  return 0;
}

} // namespace tsdgpu
