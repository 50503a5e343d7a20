// Direct FIR as a banded Toeplitz GEMM on the 5th-generation tensor cores (tcgen05 / TMEM), 3xTF32.
//
// Replaces the inner product of FiltreRIF<cfloat,float>::step (reference filtre-rt.cc:82-107) for K <= 127 real
// taps on cf32 data.  For a tile of 128 consecutive outputs t and a chunk of 32 consecutive inputs c,
//   y[128 t + j] += sum_kk h[(128 t + j) - (32 c + kk)] * x[32 c + kk]
// is D[n][j] += X[n][kk] * T[j][kk]: D in tensor memory (lane n = one of the 128 real rows {re, im} x 64 channels,
// column j = output), X the de-interleaved input chunk (A operand) and T[j][kk] = h[j - kk + 32 d], d = 4 t - c in
// [-3, 4], the Toeplitz block (B operand).  All eight blocks are row-shifted views G[j + 32 d][kk] of ONE generator
// matrix G[r][kk] = h[r - kk] (352 rows x 32 columns, 44 KiB per split part), and because the outputs are the N
// dimension each block only spends tensor time on the output columns its band touches (N = 32, 64, 96, 128, 128,
// 96, 64, 32 for K = 127: 62.5 % of the dense work).
// fp32 accuracy from tf32 inputs: x = x_hi + x_lo, h = h_hi + h_lo (each part rounded to tf32), three MMAs
// x_hi*h_lo + x_lo*h_hi + x_hi*h_hi accumulated in fp32 by the tensor core (measured error 3.7e-6 of the signal RMS).
//
// Two kernels share this formulation: fir_tc_kernel (round 1, described first: one CTA per span of tiles, LDGSTS loads, STG
// stores) and fir_tc2_kernel (round 2, the default, further down: persistent CTAs, tensor-map loads and stores, also for
// real-valued data).
// One CTA (448 threads, 1 per SM, all 512 TMEM columns) = 64 channels x `span` tiles, input-stationary: every input
// chunk crosses shared memory once as raw cf32 rows and tensor memory once as the split A operand, and feeds the two
// output tiles it overlaps, whose accumulators are live in TMEM together (3 regions of 128 columns: two accumulating,
// one being drained and re-zeroed; columns 384..511 hold two A stages of x_hi | x_lo).  Warp roles:
//   13      loader: 16-byte asynchronous copies (LDGSTS) of raw chunks into a 6-slot staging ring, two chunks of a
//           channel row back to back (512 contiguous bytes per DRAM page visit), completion on the slot's mbarrier
//   4-11    converters, two groups alternating chunks, one warp per TMEM lane quadrant: each thread reads the 32
//           samples of its row (channel, re|im) from the staging, splits them into tf32 hi / lo with two integer
//           instructions per value (cvt.rna.tf32 issues far too slowly) and writes them with tcgen05.st
//   12      MMA issuer: descriptors are warp-uniform values, one elected lane issues <= 24 tcgen05.mma.kind::tf32
//           (A from tensor memory, B = generator rows from shared memory) + tcgen05.commit per chunk
//   0-3     epilogue: tcgen05.ld.16x256b hands a thread (re, im) of two consecutive outputs of one channel -> one
//           16-byte store (a lane quad writes 64 contiguous bytes), then tcgen05.st zeros the region
// Tensor work per chunk of 2048 complex samples at K = 127: 2 tiles x 4 K-steps x 3 terms with N summing to 160 per
// term and K-step => 960 cycles at the nominal TF32 rate (64 cycles per 128x128x8); operand reads from shared memory
// 61 KiB per chunk.  Measured (-DTSD_TC_PROF, per-role clock64): MMA warp busy 950 cycles per chunk, compute-only
// bound 360 Gsamples/s, loads-only or stores-only 280, both 240: the kernel is bound by DRAM efficiency of 64-channel
// interleaved row pieces, not by the tensor cores; the FP32 FMA formulation (fir.cu) is capped at 146 Gsamples/s.
#include "tc_common.cuh"
#include "fir_tc.h"
#include "tma_host.h"

#include <algorithm>
#include <cstdlib>

namespace tsdgpu {
namespace tc {

constexpr int TILE = 128;            // outputs per tile = UMMA M
constexpr int CH = 64;               // channels per CTA; UMMA N = 2 * CH
constexpr int NCOL = 2 * CH;
constexpr int CHUNK = 32;            // input samples per chunk = 4 UMMA K steps of 8 tf32
constexpr int NRAW = 6;               // raw staging ring (chunks in flight from HBM)
constexpr int LGRP = 2;               // chunks the loader fetches together: 512 contiguous bytes per channel row (4: no further gain)
constexpr int NSTAGE = 2;             // A-operand stages in tensor memory = producer groups
constexpr int GROWS = 352;           // generator rows r in [-96, 256)
constexpr int G_BYTES = GROWS * 128; // per split part (multiple of 1024)
constexpr int RAW_PITCH = 272;                  // bytes per channel row of the raw staging (256 + 16: conflict-free LDS.128 down a column)
constexpr int RAW_BYTES = CH * RAW_PITCH;       // one group's staging buffer: 64 channels x 32 cf32 samples
constexpr int SMEM_BYTES = 2 * G_BYTES + NRAW * RAW_BYTES + 1024 /* alignment slack */ + 512 /* barriers */;
constexpr int ACOL = 3 * NCOL;                  // TMEM columns [ACOL + 64 s, +32) = x_hi, [+32, +64) = x_lo of stage s
constexpr int NGROUP = 2;              // converter groups (4 warps each, one warp per TMEM lane quadrant)
constexpr int MMA_WARP = 4 + 4 * NGROUP;
constexpr int LOAD_WARP = MMA_WARP + 1;
constexpr int NLOAD = 2;                         // loader warps: each takes half of the 64 channel rows of a chunk
constexpr int NTHREADS = 32 * (LOAD_WARP + NLOAD);
constexpr int TMEM_COLS = 512;

constexpr uint32_t IDESC = IDESC_M128;

// output columns [j0, j0 + nn) of a 128-output tile touched by Toeplitz block d (taps j - kk + 32 d in [0, K)), 16-aligned
__device__ __forceinline__ void band_cols(int d, int K, int &j0, int &nn)
{
  const int jlo = max(0, -32 * d), jhi = min(127, K + 30 - 32 * d);
  j0 = jlo & ~15;
  nn = jhi < jlo ? 0 : ((jhi + 16) & ~15) - j0;
}
#ifdef TSD_TC_PROF
__device__ long long g_tcprof[1024][24][4];
#define PROF_ARRAY g_tcprof
#endif
#include "tc_prof.cuh"

__global__ void __launch_bounds__(NTHREADS, 1) fir_tc_kernel(FirTcParams p)
{
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  unsigned char *sm = raw + (base - smem_u32(raw));
  float *Ghi = reinterpret_cast<float *>(sm), *Glo = reinterpret_cast<float *>(sm + G_BYTES);
  unsigned char *stages = sm + 2 * G_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(stages + NRAW * RAW_BYTES);
  uint64_t *full = bars, *empty = bars + NSTAGE, *tfull = bars + 2 * NSTAGE, *tempty = bars + 2 * NSTAGE + 3;
  uint64_t *rfull = bars + 2 * NSTAGE + 6, *rempty = rfull + NRAW;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(rempty + NRAW);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef TSD_TC_PROF
  const long long pf_entry = clock64();
#endif
  const int ts = blockIdx.x * p.span, te = min(ts + p.span, p.ntiles);   // tiles [ts, te) of this CTA
  const int c0 = blockIdx.y * CH;                                          // first channel
  const int nchunks = 4 * (te - ts) + 4;                                   // chunks 4 ts - 4 ... 4 te - 1

  // ---- one-time set-up: barriers, tensor memory, generator matrix (hi / lo split, swizzled K-major)
  if(tid == 0)
  {
    for(int i = 0; i < NSTAGE; i++) { mbar_init(full + i, 4); mbar_init(empty + i, 1); }
    for(int i = 0; i < 3; i++) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 4); }
    for(int i = 0; i < NRAW; i++) { mbar_init(rfull + i, 32 * NLOAD); mbar_init(rempty + i, 4); }
    mbar_fence_init();
  }
  __syncthreads();   // barriers initialised
  // The loader warps (the last NLOAD) go straight to their role: the first chunks cross HBM while the others build the
  // generator matrix, so the CTA starts with a full staging ring.  Named barrier 1 closes the set-up of the others.
  constexpr int NPRO = NTHREADS - 32 * NLOAD;
  if(warp < LOAD_WARP)
  {
    if(warp == MMA_WARP)
    {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for(int idx = tid; idx < GROWS * 32; idx += NPRO)
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
    named_bar(1, NPRO);
    fence_after();
  }
  const uint32_t tmem = warp < LOAD_WARP ? *tmem_slot : 0u;
#ifdef TSD_TC_PROF
  const long long pf_setup = clock64();
#endif

  if(warp >= LOAD_WARP)
  {
    const int j_lo = (warp - LOAD_WARP) * (32 / NLOAD), j_hi = j_lo + 32 / NLOAD;
    // ===== loader: raw chunk it (inputs [32 c, 32 c + 32), c = 4 ts - 4 + it, of 64 channels) -> staging slot it % NRAW,
    // one 256-byte row per channel, with 16-byte asynchronous copies (LDGSTS): no registers, NRAW chunks in flight,
    // completion counted on the slot's mbarrier (cp.async.mbarrier.arrive.noinc, one arrival per lane).  Lane l copies
    // pieces k = l + 32 j: channel (l / 16) + 2 j, sample pair l % 16 -> every warp instruction moves two 256-byte rows.
    // Chunks that touch the history, the end of the call or a ragged channel group use the zero-filling form
    // (src-size 0 / 8 / 16) on the same path.
    const int sp = lane & 15, clb = lane >> 4;
    auto one_chunk = [&](int it) {
      const int slot = it % NRAW;
      const uint32_t dst0 = smem_u32(stages + slot * RAW_BYTES + clb * RAW_PITCH + sp * 16);
      const long long pos = (long long) (4 * ts - 4 + it) * CHUNK + 2 * sp;
      for(int j = j_lo; j < j_hi; j++)
      {
        const int chan = c0 + clb + 2 * j;
        const float2 *src = p.x;   // any valid address when nothing is read
        unsigned bytes = 0;
        if(chan < p.nchan)
        {
          if(pos >= 0)
          {
            if(pos < p.n) { src = p.x + (long long) chan * p.x_stride + pos; bytes = pos + 1 < p.n ? 16u : 8u; }
          }
          else if(pos >= -(long long) p.halo) { src = p.hist + (long long) chan * p.halo + p.halo + pos; bytes = 16u; }
        }
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + j * 2 * RAW_PITCH), "l"(src), "r"(bytes) : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(rfull + slot)) : "memory");
    };
    for(int it = 0; it < nchunks;)
    {
      const long long pos0 = (long long) (4 * ts - 4 + it) * CHUNK;
      const bool group = it + LGRP <= nchunks && pos0 >= 0 && pos0 + LGRP * CHUNK <= p.n && c0 + CH <= p.nchan;
      mbar_wait(rempty + it % NRAW, (unsigned) (((it / NRAW) & 1) ^ 1));
      if(!group)
      {
        one_chunk(it);
        it += 1;
        continue;
      }
      // LGRP interior chunks at once: the LGRP 256-byte pieces of a channel row are adjacent in DRAM (one 1 KiB page visit)
#pragma unroll
      for(int g = 1; g < LGRP; g++) mbar_wait(rempty + (it + g) % NRAW, (unsigned) ((((it + g) / NRAW) & 1) ^ 1));
      uint32_t dst[LGRP];
#pragma unroll
      for(int g = 0; g < LGRP; g++) dst[g] = smem_u32(stages + ((it + g) % NRAW) * RAW_BYTES + clb * RAW_PITCH + sp * 16);
      const float2 *src = p.x + (long long) (c0 + clb) * p.x_stride + pos0 + 2 * sp;
#pragma unroll 4
      for(int j = j_lo; j < j_hi; j++)
      {
        const float2 *sj = src + (long long) j * 2 * p.x_stride;
#pragma unroll
        for(int g = 0; g < LGRP; g++)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst[g] + j * 2 * RAW_PITCH), "l"(sj + g * CHUNK) : "memory");
      }
#pragma unroll
      for(int g = 0; g < LGRP; g++)
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(rfull + (it + g) % NRAW)) : "memory");
      it += LGRP;
    }
  }
  else if(warp >= 4 && warp < 4 + 4 * NGROUP)
  {
    // ===== converters: group g takes chunks it = g, g + 2, ... and owns A stage g in tensor memory.  This thread's TMEM
    // lane 32 pw + lane is row (cl, ri): channel cl = 8 (lane / 16 + 2 pw) + lane % 8, ri = (lane / 8) % 2 (re rows and,
    // 8 lanes below, im rows: the pairing the 16x256b epilogue load wants).  It reads the 32 samples of its row out of the
    // staging (16 LDS.128), splits its component into tf32 hi / lo and writes both straight into tensor memory.
    const int pw = (warp - 4) & 3, grp = (warp - 4) >> 2;
    const int my_cl = 8 * (2 * pw + (lane >> 4)) + (lane & 7), my_ri = (lane >> 3) & 1;
    const uint32_t my_a = tmem + ((uint32_t) (pw * 32) << 16) + (uint32_t) (ACOL + 64 * grp);
    PROF_DECL
    for(int it = grp; it < nchunks; it += NGROUP)
    {
      const int slot = it % NRAW;
      PROF_BEGIN(t_w)
      mbar_wait(rfull + slot, (unsigned) ((it / NRAW) & 1));
      mbar_wait(empty + grp, (unsigned) (((it / NSTAGE) & 1) ^ 1));   // the MMAs of chunk it - 2 have read this A stage
      PROF_ADD(0, t_w)
      PROF_BEGIN(t_c)
      fence_after();
      const unsigned char *row = stages + slot * RAW_BYTES + my_cl * RAW_PITCH;
#pragma unroll
      for(int hq = 0; hq < 2; hq++)
      {
        float hi[16], lo[16];
#pragma unroll
        for(int m = 0; m < 8; m++)
        {
          const float4 x = *reinterpret_cast<const float4 *>(row + (hq * 8 + m) * 16);   // (re0, im0, re1, im1)
          const float a0 = my_ri ? x.y : x.x, a1 = my_ri ? x.w : x.z;
          hi[2 * m] = to_tf32(a0);
          hi[2 * m + 1] = to_tf32(a1);
          lo[2 * m] = to_tf32(a0 - hi[2 * m]);
          lo[2 * m + 1] = to_tf32(a1 - hi[2 * m + 1]);
        }
        tmem_st16(my_a + hq * 16, hi);
        tmem_st16(my_a + 32 + hq * 16, lo);
      }
      __syncwarp();
      if(lane == 0) mbar_arrive(rempty + slot);            // staging slot may be refilled
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      fence_before();
      __syncwarp();
      if(lane == 0) mbar_arrive(full + grp);
      PROF_ADD(2, t_c)
    }
    PROF_END
  }
  else if(warp == MMA_WARP)
  {
    // ===== MMA issuer: the whole warp runs the loop on warp-uniform values (so that descriptors live in uniform
    // registers) and one elected lane issues the 24 MMAs + commits of a chunk in a single block.  The issue loop
    // is the critical path of the kernel: keep it free of per-MMA address arithmetic and branches.
    const uint32_t ghi = base, glo = base + G_BYTES;
    const uint64_t dbase = smem_desc(0);
    const int ntl = te - ts;
    PROF_DECL
    for(int it = 0; it < nchunks; it++)
    {
      const int stage = it % NSTAGE;
      PROF_BEGIN(t_w)
      mbar_wait(full + stage, (unsigned) ((it / NSTAGE) & 1));
      PROF_ADD(0, t_w)
      // chunk c = 4 ts - 4 + it feeds tiles ts + (it / 4) - 1 (d = 4 t - c = -(it % 4)) and ts + it / 4 (d = 4 - it % 4)
      const int q4 = it >> 2, r4 = it & 3;
      const int tl0 = q4 - 1, tl1 = q4;
      const bool on0 = tl0 >= 0 && tl0 < ntl, on1 = tl1 < ntl;
      PROF_BEGIN(t_w2)
      if(on1 && r4 == 0) mbar_wait(tempty + tl1 % 3, (unsigned) ((tl1 / 3) & 1));   // accumulator region drained and zeroed
      PROF_ADD(1, t_w2)
      PROF_BEGIN(t_c)
      fence_after();
      const uint32_t xh0 = tmem + (uint32_t) (ACOL + 64 * stage), xl0 = xh0 + 32;   // A operand: data chunk in tensor memory, hi / lo
      int j0a, na, j0b, nb;
      band_cols(-r4, p.K, j0a, na);
      band_cols(4 - r4, p.K, j0b, nb);
      const uint32_t rowa = (uint32_t) ((96 - 32 * r4 + j0a) * 128), rowb = (uint32_t) ((224 - 32 * r4 + j0b) * 128);
      const uint64_t gha = dbase + ((ghi + rowa) >> 4), gla = dbase + ((glo + rowa) >> 4);   // B operand: Toeplitz rows
      const uint64_t ghb = dbase + ((ghi + rowb) >> 4), glb = dbase + ((glo + rowb) >> 4);
      const uint32_t da = tmem + (uint32_t) ((tl0 + 3) % 3 * NCOL + j0a), db = tmem + (uint32_t) (tl1 % 3 * NCOL + j0b);
      const uint32_t ida = IDESC | ((uint32_t) (na >> 3) << 17), idb = IDESC | ((uint32_t) (nb >> 3) << 17);
      if(elect_one())
      {
        if(on0 && na > 0)
        {
#pragma unroll
          for(int ks = 0; ks < 4; ks++)
          {
            mma_tf32(da, xh0 + 8 * ks, gla + 2 * ks, ida);
            mma_tf32(da, xl0 + 8 * ks, gha + 2 * ks, ida);
            mma_tf32(da, xh0 + 8 * ks, gha + 2 * ks, ida);
          }
        }
        if(on0 && r4 == 3) mma_commit(tfull + tl0 % 3);     // d = -3: last chunk of tile tl0, accumulator complete
        if(on1 && nb > 0)
        {
#pragma unroll
          for(int ks = 0; ks < 4; ks++)
          {
            mma_tf32(db, xh0 + 8 * ks, glb + 2 * ks, idb);
            mma_tf32(db, xl0 + 8 * ks, ghb + 2 * ks, idb);
            mma_tf32(db, xh0 + 8 * ks, ghb + 2 * ks, idb);
          }
        }
        mma_commit(empty + stage);                          // the stage may be refilled once these MMAs have read it
      }
      __syncwarp();
      PROF_ADD(2, t_c)
    }
    PROF_END
  }
  else
  {
    // ===== epilogue: warp w owns TMEM lanes 32 w ... 32 w + 31 = rows of channels 16 w ... 16 w + 15 (re rows and, 8 lanes
    // below, im rows).  tcgen05.ld.16x256b.x4 hands thread t, for column block i, registers {4i, 4i+1} = lane t/4, columns
    // 8 i + 2 (t % 4) + {0, 1} and {4i+2, 4i+3} = lane t/4 + 8, same columns: (re, im) of two consecutive outputs of one
    // channel -> one 16-byte store; a quad of lanes writes 64 contiguous bytes.
    auto zero_region = [&](int region) {
#pragma unroll
      for(int q = 0; q < 4; q++)
      {
        const uint32_t taddr = tmem + ((uint32_t) (warp * 32) << 16) + (uint32_t) (region * NCOL + q * 32);
        asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
          "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u) : "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    };
    for(int region = 0; region < 3; region++)
    {
      zero_region(region);
      fence_before();
      __syncwarp();
      if(lane == 0) mbar_arrive(tempty + region);
    }
    PROF_DECL
    for(int tl = 0; tl < te - ts; tl++)
    {
      const int region = tl % 3;
      PROF_BEGIN(t_w)
      mbar_wait(tfull + region, (unsigned) ((tl / 3) & 1));
      PROF_ADD(0, t_w)
      PROF_BEGIN(t_c)
      fence_after();
      const long long n0 = (long long) (ts + tl) * TILE + 2 * (lane & 3);
#pragma unroll
      for(int half = 0; half < 2; half++)
      {
        const int chan = c0 + 8 * (2 * warp + half) + (lane >> 2);
        float2 *yrow = p.y + (long long) chan * p.y_stride;
#pragma unroll
        for(int cb = 0; cb < 4; cb++)
        {
          uint32_t r[16];
          const uint32_t taddr = tmem + ((uint32_t) (warp * 32 + half * 16) << 16) + (uint32_t) (region * NCOL + cb * 32);
          asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr)
            : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if(chan < p.nchan)
          {
#pragma unroll
            for(int i = 0; i < 4; i++)
            {
              const long long nabs = n0 + cb * 32 + 8 * i;
              const float4 o = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 1]),
                                           __uint_as_float(r[4 * i + 3]));
              if(nabs + 1 < p.n) *reinterpret_cast<float4 *>(yrow + nabs) = o;
              else if(nabs < p.n) yrow[nabs] = make_float2(o.x, o.y);
            }
          }
        }
      }
      zero_region(region);
      fence_before();
      __syncwarp();
      if(lane == 0) mbar_arrive(tempty + region);
      PROF_ADD(2, t_c)
    }
    PROF_END
  }
  // ---- teardown
  fence_before();
  __syncthreads();
#ifdef TSD_TC_PROF
  if(tid == 0 && blockIdx.y * gridDim.x + blockIdx.x < 1024)
  {
    const int bi = blockIdx.y * gridDim.x + blockIdx.x;
    g_tcprof[bi][20][0] = pf_setup - pf_entry;
    g_tcprof[bi][20][1] = clock64() - pf_setup;
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_tcprof[bi][20][2] = (long long) gt;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_tcprof[bi][20][3] = smid;
  }
#endif
  if(warp == MMA_WARP)
  {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}


// =====================================================================================================================
// Round-2 form of the same GEMM (default): PERSISTENT CTAs, the TMA unit on both sides of the SM.
//   * grid = one CTA per SM; the (channel group, tile) space is cut into gridDim.x contiguous shares, a share that
//     crosses a channel-group boundary is walked as two items.  Barriers, tensor memory and the generator matrix are set
//     up once per CTA and launch; ring slots, A stages and accumulator regions are indexed by counters that run across
//     items, so the pipeline never drains between them.
//   * loads (TMAL): a chunk [64 channels][32 samples] = two 2-D tensor-map boxes {32 floats, 64 rows} with the 128-byte
//     swizzle (cp.async.bulk.tensor.2d, SASS UTMALDG), issued by ONE lane, completion by expect_tx on the slot's
//     mbarrier.  History chunks (positions < 0) come through a second map over the device history; the end of the call,
//     a short history and a ragged channel group are the tensor map's out-of-bounds zero fill: no special-case path.
//     The converters read 16-byte pieces at (piece ^ (row & 7)): conflict-free.
//   * stores: the epilogue warps write their (re, im) pairs into a warp-private staging [8 channel rows][128 outputs]
//     and hand it to the TMA unit (cp.async.bulk.tensor.2d.global.shared, SASS UTMASTG): 1 KiB contiguous per channel
//     row instead of 64-byte pieces per store instruction, no STG issue in the epilogue warps; the four tcgen05.ld of a
//     channel half are issued back to back and waited for once.  Rows / columns beyond nchan / n are clipped by the map.
constexpr int RAW2_BYTES = 16384;                 // TMAL: two swizzled boxes [64 rows][128 B]
constexpr int OUT2_BYTES = 8192;                  // per epilogue warp: [8 channel rows][128 outputs] cf32
constexpr int LOAD_WARP2 = MMA_WARP + 1;
template<bool TMAL> struct Cfg2
{
  static constexpr int NRAW = 6;
  static constexpr int NLD = TMAL ? 1 : NLOAD;    // loader warps
  static constexpr int SLOT = TMAL ? RAW2_BYTES : RAW_BYTES;
  static constexpr int NTHREADS = 32 * (LOAD_WARP2 + NLD);
  static constexpr int SMEM = 2 * G_BYTES + NRAW * SLOT + 4 * OUT2_BYTES + 1024 /* alignment slack */ + 512 /* barriers */;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
               "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, uint32_t src_smem)
{
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(src_smem)
               : "memory");
}

// REAL = true (FiltreRIF<float,float>, e.g. the README example / BASELINE config 1): the 128 GEMM rows are 128 real-valued
// channels instead of {re, im} x 64; a chunk is ONE box {32 floats, 128 rows}, the converters read four samples per 16-byte
// piece, the epilogue writes real outputs.  Generator matrix, MMA issue and accumulator handling are the same.
template<bool TMAL, bool REAL>
__global__ void __launch_bounds__(Cfg2<TMAL>::NTHREADS, 1)
fir_tc2_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap hmap, const __grid_constant__ CUtensorMap ymap, FirTcParams p)
{
  using C = Cfg2<TMAL>;
  constexpr int NR = C::NRAW;
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  unsigned char *sm = raw + (base - smem_u32(raw));
  float *Ghi = reinterpret_cast<float *>(sm), *Glo = reinterpret_cast<float *>(sm + G_BYTES);
  unsigned char *stages = sm + 2 * G_BYTES;
  unsigned char *outs = stages + NR * C::SLOT;
  uint64_t *bars = reinterpret_cast<uint64_t *>(outs + 4 * OUT2_BYTES);
  uint64_t *full = bars, *empty = bars + NSTAGE, *tfull = bars + 2 * NSTAGE, *tempty = bars + 2 * NSTAGE + 3;
  uint64_t *rfull = bars + 2 * NSTAGE + 6, *rempty = rfull + NR;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(rempty + NR);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // this CTA's contiguous share of the (channel group, tile) units
  static_assert(TMAL || !REAL, "real-valued data: tensor-map form only");
  constexpr int CHG = REAL ? 2 * CH : CH;   // channels per group
  const int groups = (p.nchan + CHG - 1) / CHG;
  const long long units = (long long) groups * p.ntiles;
  const long long u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1) / gridDim.x;
#define FIR2_ITEMS_BEGIN                                                                                   \
  for(long long u = u0; u < u1;)                                                                           \
  {                                                                                                        \
    const int g = (int) (u / p.ntiles), ts = (int) (u - (long long) g * p.ntiles);                         \
    const int ntl = (int) min((long long) (p.ntiles - ts), u1 - u), nchunks = 4 * ntl + 4, c0 = g * CHG;   \
    (void) c0; (void) nchunks;
#define FIR2_ITEMS_END(cnt_chunks, cnt_tiles)                                                              \
    u += ntl;                                                                                              \
    cnt_chunks += (unsigned) nchunks;                                                                      \
    cnt_tiles += (unsigned) ntl;                                                                           \
  }

  if(tid == 0)
  {
    for(int i = 0; i < NSTAGE; i++) { mbar_init(full + i, 4); mbar_init(empty + i, 1); }
    for(int i = 0; i < 3; i++) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 4); }
    for(int i = 0; i < NR; i++) { mbar_init(rfull + i, TMAL ? 1 : 32 * C::NLD); mbar_init(rempty + i, 4); }
    mbar_fence_init();
  }
  __syncthreads();
  constexpr int NPRO = 32 * LOAD_WARP2;   // everyone but the loaders builds the generator matrix
  if(warp < LOAD_WARP2)
  {
    if(warp == MMA_WARP)
    {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for(int idx = tid; idx < GROWS * 32; idx += NPRO)
    {
      const int row = idx >> 5, kk = idx & 31, tap = (row - 96) - kk;
      const float v = (tap >= 0 && tap < p.K) ? __ldg(p.taps_rev + (p.K - 1 - tap)) : 0.f;
      const float hi = to_tf32(v), lo = to_tf32(v - hi);
      const uint32_t off = swz((uint32_t) (row * 128 + kk * 4));
      *reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(Ghi) + off) = hi;
      *reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(Glo) + off) = lo;
    }
    fence_proxy_async();
    fence_before();
    named_bar(1, NPRO);
    fence_after();
  }
  const uint32_t tmem = warp < LOAD_WARP2 ? *tmem_slot : 0u;

  if(warp >= LOAD_WARP2)
  {
    if(TMAL)
    {
      // ===== loader (one lane): two chunks of a channel row back to back = four boxes, 512 contiguous bytes per row
      if(lane == 0)
      {
        unsigned gi = 0, gt = 0;
        FIR2_ITEMS_BEGIN
        for(int it = 0; it < nchunks; it += LGRP)
        {
#pragma unroll
          for(int k = 0; k < LGRP; k++) mbar_wait_long(rempty + (gi + it + k) % NR, (unsigned) ((((gi + it + k) / NR) & 1) ^ 1));
#pragma unroll
          for(int k = 0; k < LGRP; k++)
          {
            const unsigned slot = (gi + it + k) % NR;
            const int pos = (4 * ts - 4 + it + k) * CHUNK;
            const uint32_t dst = smem_u32(stages + slot * RAW2_BYTES);
            mbar_expect_tx(rfull + slot, RAW2_BYTES);
            if(REAL)
            {
              if(pos >= 0) tma_load_2d(dst, &xmap, pos, c0, rfull + slot);
              else tma_load_2d(dst, &hmap, p.halo + pos, c0, rfull + slot);
            }
            else if(pos >= 0)
            {
              tma_load_2d(dst, &xmap, 2 * pos, c0, rfull + slot);
              tma_load_2d(dst + 8192, &xmap, 2 * pos + 32, c0, rfull + slot);
            }
            else
            {
              tma_load_2d(dst, &hmap, 2 * (p.halo + pos), c0, rfull + slot);
              tma_load_2d(dst + 8192, &hmap, 2 * (p.halo + pos) + 32, c0, rfull + slot);
            }
          }
        }
        FIR2_ITEMS_END(gi, gt)
      }
    }
    else
    {
      // ===== loaders, LDGSTS form (see fir_tc_kernel): 16-byte copies into the pitch-272 staging, zero-filling at the edges
      const int j_lo = (warp - LOAD_WARP2) * (32 / C::NLD), j_hi = j_lo + 32 / C::NLD;
      const int sp = lane & 15, clb = lane >> 4;
      unsigned gi = 0, gt = 0;
      FIR2_ITEMS_BEGIN
      for(int it = 0; it < nchunks;)
      {
        const long long pos0 = (long long) (4 * ts - 4 + it) * CHUNK;
        const bool group = it + LGRP <= nchunks && pos0 >= 0 && pos0 + LGRP * CHUNK <= p.n && c0 + CH <= p.nchan;
        mbar_wait(rempty + (gi + it) % NR, (unsigned) ((((gi + it) / NR) & 1) ^ 1));
        if(!group)
        {
          const unsigned slot = (gi + it) % NR;
          const uint32_t dst0 = smem_u32(stages + slot * RAW_BYTES + clb * RAW_PITCH + sp * 16);
          const long long pos = pos0 + 2 * sp;
          for(int j = j_lo; j < j_hi; j++)
          {
            const int chan = c0 + clb + 2 * j;
            const float2 *src = p.x;
            unsigned bytes = 0;
            if(chan < p.nchan)
            {
              if(pos >= 0)
              {
                if(pos < p.n) { src = p.x + (long long) chan * p.x_stride + pos; bytes = pos + 1 < p.n ? 16u : 8u; }
              }
              else if(pos >= -(long long) p.halo) { src = p.hist + (long long) chan * p.halo + p.halo + pos; bytes = 16u; }
            }
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + j * 2 * RAW_PITCH), "l"(src), "r"(bytes) : "memory");
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(rfull + slot)) : "memory");
          it += 1;
          continue;
        }
#pragma unroll
        for(int k = 1; k < LGRP; k++) mbar_wait(rempty + (gi + it + k) % NR, (unsigned) ((((gi + it + k) / NR) & 1) ^ 1));
        uint32_t dst[LGRP];
#pragma unroll
        for(int k = 0; k < LGRP; k++) dst[k] = smem_u32(stages + ((gi + it + k) % NR) * RAW_BYTES + clb * RAW_PITCH + sp * 16);
        const float2 *src = p.x + (long long) (c0 + clb) * p.x_stride + pos0 + 2 * sp;
#pragma unroll 4
        for(int j = j_lo; j < j_hi; j++)
        {
          const float2 *sj = src + (long long) j * 2 * p.x_stride;
#pragma unroll
          for(int k = 0; k < LGRP; k++)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst[k] + j * 2 * RAW_PITCH), "l"(sj + k * CHUNK) : "memory");
        }
#pragma unroll
        for(int k = 0; k < LGRP; k++)
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(rfull + (gi + it + k) % NR)) : "memory");
        it += LGRP;
      }
      FIR2_ITEMS_END(gi, gt)
    }
  }
  else if(warp >= 4 && warp < 4 + 4 * NGROUP)
  {
    // ===== converters (see fir_tc_kernel); group = parity of the running chunk number
    const int pw = (warp - 4) & 3, grp = (warp - 4) >> 2;
    const int my_cl = REAL ? 32 * pw + lane : 8 * (2 * pw + (lane >> 4)) + (lane & 7), my_ri = (lane >> 3) & 1;
    const uint32_t my_a = tmem + ((uint32_t) (pw * 32) << 16) + (uint32_t) (ACOL + 64 * grp);
    const uint32_t sx = (uint32_t) (my_cl & 7);
    unsigned gi = 0, gt = 0;
    FIR2_ITEMS_BEGIN
    for(int it = (int) ((grp - gi) & 1u); it < nchunks; it += NGROUP)
    {
      const unsigned gc = gi + it, slot = gc % NR;
      mbar_wait(rfull + slot, (gc / NR) & 1);
      mbar_wait(empty + grp, ((gc / NSTAGE) & 1) ^ 1);
      fence_after();
      const unsigned char *row = TMAL ? stages + slot * RAW2_BYTES + my_cl * 128 : stages + slot * RAW_BYTES + my_cl * RAW_PITCH;
#pragma unroll
      for(int hq = 0; hq < 2; hq++)
      {
        float hi[16], lo[16];
        if(REAL)
        {
          // row = channel: 32 real samples = eight 16-byte pieces of ONE 128-byte swizzled row
#pragma unroll
          for(int m = 0; m < 4; m++)
          {
            const float4 x = *reinterpret_cast<const float4 *>(row + (((uint32_t) (hq * 4 + m) ^ sx) << 4));
            const float a[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for(int e = 0; e < 4; e++)
            {
              hi[4 * m + e] = to_tf32(a[e]);
              lo[4 * m + e] = to_tf32(a[e] - hi[4 * m + e]);
            }
          }
        }
        else
#pragma unroll
        for(int m = 0; m < 8; m++)
        {
          const float4 x = TMAL ? *reinterpret_cast<const float4 *>(row + hq * 8192 + (((uint32_t) m ^ sx) << 4))
                                : *reinterpret_cast<const float4 *>(row + (hq * 8 + m) * 16);
          const float a0 = my_ri ? x.y : x.x, a1 = my_ri ? x.w : x.z;
          hi[2 * m] = to_tf32(a0);
          hi[2 * m + 1] = to_tf32(a1);
          lo[2 * m] = to_tf32(a0 - hi[2 * m]);
          lo[2 * m + 1] = to_tf32(a1 - hi[2 * m + 1]);
        }
        tmem_st16(my_a + hq * 16, hi);
        tmem_st16(my_a + 32 + hq * 16, lo);
      }
      __syncwarp();
      if(lane == 0) mbar_arrive(rempty + slot);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      fence_before();
      __syncwarp();
      if(lane == 0) mbar_arrive(full + grp);
    }
    FIR2_ITEMS_END(gi, gt)
  }
  else if(warp == MMA_WARP)
  {
    // ===== MMA issuer (see fir_tc_kernel)
    const uint32_t ghi = base, glo = base + G_BYTES;
    const uint64_t dbase = smem_desc(0);
    unsigned gi = 0, gt = 0;
    FIR2_ITEMS_BEGIN
    for(int it = 0; it < nchunks; it++)
    {
      const unsigned gc = gi + it, stage = gc % NSTAGE;
      mbar_wait(full + stage, (gc / NSTAGE) & 1);
      const int q4 = it >> 2, r4 = it & 3;
      const int tl0 = q4 - 1, tl1 = q4;
      const bool on0 = tl0 >= 0, on1 = tl1 < ntl;
      const unsigned T0 = gt + (unsigned) tl0, T1 = gt + (unsigned) tl1;   // running tile numbers (T0 only used when on0)
      if(on1 && r4 == 0) mbar_wait(tempty + T1 % 3, (T1 / 3) & 1);
      fence_after();
      const uint32_t xh0 = tmem + (uint32_t) (ACOL + 64 * stage), xl0 = xh0 + 32;
      int j0a, na, j0b, nb;
      band_cols(-r4, p.K, j0a, na);
      band_cols(4 - r4, p.K, j0b, nb);
      const uint32_t rowa = (uint32_t) ((96 - 32 * r4 + j0a) * 128), rowb = (uint32_t) ((224 - 32 * r4 + j0b) * 128);
      const uint64_t gha = dbase + ((ghi + rowa) >> 4), gla = dbase + ((glo + rowa) >> 4);
      const uint64_t ghb = dbase + ((ghi + rowb) >> 4), glb = dbase + ((glo + rowb) >> 4);
      const uint32_t da = tmem + (uint32_t) ((on0 ? T0 % 3 : 0u) * NCOL + j0a), db = tmem + (uint32_t) (T1 % 3 * NCOL + j0b);
      const uint32_t ida = IDESC | ((uint32_t) (na >> 3) << 17), idb = IDESC | ((uint32_t) (nb >> 3) << 17);
      if(elect_one())
      {
        if(on0 && na > 0)
        {
#pragma unroll
          for(int ks = 0; ks < 4; ks++)
          {
            mma_tf32(da, xh0 + 8 * ks, gla + 2 * ks, ida);
            mma_tf32(da, xl0 + 8 * ks, gha + 2 * ks, ida);
            mma_tf32(da, xh0 + 8 * ks, gha + 2 * ks, ida);
          }
        }
        if(on0 && r4 == 3) mma_commit(tfull + T0 % 3);
        if(on1 && nb > 0)
        {
#pragma unroll
          for(int ks = 0; ks < 4; ks++)
          {
            mma_tf32(db, xh0 + 8 * ks, glb + 2 * ks, idb);
            mma_tf32(db, xl0 + 8 * ks, ghb + 2 * ks, idb);
            mma_tf32(db, xh0 + 8 * ks, ghb + 2 * ks, idb);
          }
        }
        mma_commit(empty + stage);
      }
      __syncwarp();
    }
    FIR2_ITEMS_END(gi, gt)
  }
  else
  {
    // ===== epilogue: warp w owns TMEM lanes 32 w ... 32 w + 31 = channels 16 w ... 16 w + 15; per channel half (8 channels)
    // four tcgen05.ld.16x256b.x4 (thread t: (re, im) of outputs 32 cb + 8 i + 2 (t % 4) + {0, 1} of channel t / 4) -> one
    // wait -> 16 STS.128 into the warp's staging [8 rows][128 outputs] -> one TMA store of the box {256 floats, 8 rows}
    auto zero_region = [&](unsigned region) {
#pragma unroll
      for(int q = 0; q < 4; q++)
      {
        const uint32_t taddr = tmem + ((uint32_t) (warp * 32) << 16) + (uint32_t) (region * NCOL + q * 32);
        asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
          "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u) : "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    };
    for(unsigned region = 0; region < 3; region++)
    {
      zero_region(region);
      fence_before();
      __syncwarp();
      if(lane == 0) mbar_arrive(tempty + region);
    }
    unsigned char *my_out = outs + warp * OUT2_BYTES;
    const uint32_t my_out_s = smem_u32(my_out);
    // cf32: staging [8 channel rows][128 outputs x 8 B]; real: [16 channel rows][128 outputs x 4 B] (TMEM lanes t/4 and t/4 + 8
    // are two different channels there)
    unsigned char *my_piece = REAL ? my_out + (lane >> 2) * 512 + (lane & 3) * 8 : my_out + (lane >> 2) * 1024 + (lane & 3) * 16;
    unsigned gi = 0, gt = 0;
    FIR2_ITEMS_BEGIN
    for(int tl = 0; tl < ntl; tl++)
    {
      const unsigned T = gt + (unsigned) tl, region = T % 3;
      mbar_wait_long(tfull + region, (T / 3) & 1);
      fence_after();
#pragma unroll
      for(int half = 0; half < 2; half++)
      {
        uint32_t r[4][16];
#pragma unroll
        for(int cb = 0; cb < 4; cb++)
        {
          const uint32_t taddr = tmem + ((uint32_t) (warp * 32 + half * 16) << 16) + (uint32_t) (region * NCOL + cb * 32);
          asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[cb][0]), "=r"(r[cb][1]), "=r"(r[cb][2]), "=r"(r[cb][3]), "=r"(r[cb][4]), "=r"(r[cb][5]), "=r"(r[cb][6]), "=r"(r[cb][7]),
              "=r"(r[cb][8]), "=r"(r[cb][9]), "=r"(r[cb][10]), "=r"(r[cb][11]), "=r"(r[cb][12]), "=r"(r[cb][13]), "=r"(r[cb][14]), "=r"(r[cb][15])
            : "r"(taddr)
            : "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if(lane == 0) bulk_wait_read0();      // the previous store of this warp has read the staging
        __syncwarp();
#pragma unroll
        for(int cb = 0; cb < 4; cb++)
#pragma unroll
          for(int i = 0; i < 4; i++)
          {
            if(REAL)
            {
              *reinterpret_cast<uint2 *>(my_piece + (cb * 32 + 8 * i) * 4) = make_uint2(r[cb][4 * i], r[cb][4 * i + 1]);
              *reinterpret_cast<uint2 *>(my_piece + 8 * 512 + (cb * 32 + 8 * i) * 4) = make_uint2(r[cb][4 * i + 2], r[cb][4 * i + 3]);
            }
            else
              *reinterpret_cast<uint4 *>(my_piece + (cb * 32 + 8 * i) * 8) = make_uint4(r[cb][4 * i], r[cb][4 * i + 2], r[cb][4 * i + 1], r[cb][4 * i + 3]);
          }
        fence_proxy_async();
        __syncwarp();
        if(lane == 0)
        {
          if(REAL) tma_store_2d(&ymap, TILE * (ts + tl), c0 + 32 * warp + 16 * half, my_out_s);
          else tma_store_2d(&ymap, 2 * TILE * (ts + tl), c0 + 16 * warp + 8 * half, my_out_s);
          bulk_commit();
        }
      }
      zero_region(region);
      fence_before();
      __syncwarp();
      if(lane == 0) mbar_arrive(tempty + region);
    }
    FIR2_ITEMS_END(gi, gt)
    if(lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  // ---- teardown
  fence_before();
  __syncthreads();
  if(warp == MMA_WARP)
  {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
#undef FIR2_ITEMS_BEGIN
#undef FIR2_ITEMS_END
}

} // namespace tc

// real-valued data (f32 x f32): tensor-map form only -> 16-byte aligned rows and pitches
bool fir_tc_real_eligible(int K, const void *x, long long x_stride, const void *y, long long y_stride, const void *hist, int halo)
{
  const char *e = getenv("TSDGPU_FIR_TC_VARIANT");
  if(e && atoi(e) >= 1 && atoi(e) <= 2) return false;
  return tma_encode_fn() && K >= 1 && K <= 127 && (((uintptr_t) x & 15) == 0) && (x_stride % 4 == 0) && (((uintptr_t) y & 15) == 0) &&
         (y_stride % 4 == 0) && (((uintptr_t) hist & 15) == 0) && (halo % 4 == 0);
}

bool fir_tc_eligible(int kind_cf32_f32, int K, const void *x, long long x_stride, const void *y, long long y_stride, const void *hist, int halo)
{
  return kind_cf32_f32 && K >= 1 && K <= 127 && (((uintptr_t) x & 15) == 0) && (x_stride % 2 == 0) && (((uintptr_t) y & 15) == 0) &&
         (y_stride % 2 == 0) && (((uintptr_t) hist & 15) == 0) && (halo % 2 == 0);
}

#ifdef TSD_TC_PROF
extern "C" int tsdgpu_debug_tcprof_dump(const char *path)
{
  cudaDeviceSynchronize();
  static long long h[1024][24][4];
  if(cudaMemcpyFromSymbol(h, tc::g_tcprof, sizeof(h)) != cudaSuccess) return 1;
  FILE *fp = fopen(path, "wb");
  if(!fp) return 1;
  fwrite(h, 1, sizeof(h), fp);
  fclose(fp);
  return 0;
}
#endif

// TSDGPU_FIR_TC_VARIANT: 1 = round-1 kernel (one CTA per span, LDGSTS loads, STG stores), 2 = persistent + TMA stores with
// the LDGSTS loader, 3 / unset = persistent + TMA loads + TMA stores (default).  Read at every call (tests switch it).
static int fir_tc_variant()
{
  const char *e = getenv("TSDGPU_FIR_TC_VARIANT");
  const int x = e ? atoi(e) : 3;
  return (x >= 1 && x <= 3) ? x : 3;
}

static int fir_tc_launch_v1(FirTcParams p)
{
  Runtime &r = rt();
  if(!r.fir_tc1_ready)
  {
    TSD_CUDA(cudaFuncSetAttribute(tc::fir_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    r.fir_tc1_ready = true;
  }
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

int fir_tc_launch(const FirTcParams &p0)
{
  FirTcParams p = p0;
  Runtime &r = rt();
  p.ntiles = (p.n + tc::TILE - 1) / tc::TILE;
  p.span = 0;
  const int variant = p.real ? 3 : fir_tc_variant();
  // tensor maps over float32 views of the rows: x [nchan][2 n], history [nchan][2 halo], y [nchan][2 n]
  CUtensorMap xmap, hmap, ymap;
  if(p.real)
  {
    // real-valued data: rows of n floats, 128 channel rows per box (x_stride / y_stride / halo count floats here)
    if(!(tma_map_rows(&ymap, p.y, (unsigned long long) p.n, p.nchan, (unsigned long long) p.y_stride * 4, 128, 16, false) &&
         tma_map_rows(&xmap, p.x, (unsigned long long) p.n, p.nchan, (unsigned long long) p.x_stride * 4, 32, 2 * tc::CH, true) &&
         tma_map_rows(&hmap, p.hist, (unsigned long long) p.halo, p.nchan, (unsigned long long) p.halo * 4, 32, 2 * tc::CH, true)))
      return fail("fir_tc_launch: cuTensorMapEncodeTiled refused the real-valued rows");
  }
  const bool maps = p.real || variant >= 2 && (unsigned long long) p.x_stride * 8 < (1ull << 40) && (unsigned long long) p.y_stride * 8 < (1ull << 40) &&
                    tma_map_rows(&ymap, p.y, 2ull * p.n, p.nchan, (unsigned long long) p.y_stride * 8, 256, 8, false) &&
                    tma_map_rows(&xmap, p.x, 2ull * p.n, p.nchan, (unsigned long long) p.x_stride * 8, 32, tc::CH, true) &&
                    tma_map_rows(&hmap, p.hist, 2ull * p.halo, p.nchan, (unsigned long long) p.halo * 8, 32, tc::CH, true);
  if(!maps) return fir_tc_launch_v1(p);
  if(!r.fir_tc2_ready)
  {
    TSD_CUDA(cudaFuncSetAttribute(tc::fir_tc2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg2<true>::SMEM));
    TSD_CUDA(cudaFuncSetAttribute(tc::fir_tc2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg2<false>::SMEM));
    TSD_CUDA(cudaFuncSetAttribute(tc::fir_tc2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg2<true>::SMEM));
    r.fir_tc2_ready = true;
  }
  const int chg = p.real ? 2 * tc::CH : tc::CH;
  const long long units = (long long) ((p.nchan + chg - 1) / chg) * p.ntiles;
  // one CTA per SM; small calls: at least 4 tiles per CTA (every item pays one tile's worth of halo chunks)
  const int grid = (int) std::max<long long>(1, std::min<long long>(r.num_sms, (units + 3) / 4));
  if(p.real) tc::fir_tc2_kernel<true, true><<<grid, tc::Cfg2<true>::NTHREADS, tc::Cfg2<true>::SMEM, r.stream>>>(xmap, hmap, ymap, p);
  else if(variant == 2) tc::fir_tc2_kernel<false, false><<<grid, tc::Cfg2<false>::NTHREADS, tc::Cfg2<false>::SMEM, r.stream>>>(xmap, hmap, ymap, p);
  else tc::fir_tc2_kernel<true, false><<<grid, tc::Cfg2<true>::NTHREADS, tc::Cfg2<true>::SMEM, r.stream>>>(xmap, hmap, ymap, p);
  TSD_LAUNCH_CHECK();
  return 0;
}

} // namespace tsdgpu
