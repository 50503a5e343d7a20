// Arbitrary-ratio LUT resampler as a banded filter-bank GEMM on the 5th-generation tensor cores (tcgen05 / TMEM), 3xTF32.
//
// Replaces the inner loop of AdaptationRythmeSimple::step + InterpolateurRIF::step (reference ra.cc:39-77,
// filtrage.hpp:1873-1881): out[c][j] = sum_i lut[p_j][i] * x[c][in_j - (K-1) + i], with (in_j, p_j) the host schedule of
// the reference's float32 phase recurrence (resamp.cu).  All channels share the schedule, so for a tile of 128
// consecutive outputs t and a chunk of 32 consecutive inputs c
//   D[n][j] += X[n][kk] * T[j][kk],   T[j][kk] = lut[p_j][32 c + kk - (in_j - (K-1))]  (0 outside the K taps)
// with D in tensor memory (lane n = one of the 128 real rows {re, im} x 64 channels, column j = output), X the
// de-interleaved input chunk (A operand, in tensor memory) and T the block of the banded coefficient matrix that
// eight generator warps build from the LUT and the schedule (B operand, shared memory, K-major, 128-byte swizzle).
// Same machine as fir_tc.cu (loader warp with LDGSTS staging ring, converter warps -> tcgen05.st, one elected MMA
// issuer, 16x256b epilogue, 3 accumulator regions); what differs is the B operand (generated per block instead of a
// view of one Toeplitz generator) and the irregular chunk <-> tile incidence: tile t is fed by the chunks
// floor((in_first - (K-1)) / 32) ... floor(in_last / 32), a chunk feeds one or two consecutive tiles (checked on the
// host: tile t+2 must start after tile t ends), every role walks the same (chunk, tile) block sequence.
// fp32 accuracy: x = x_hi + x_lo, T = T_hi + T_lo (tf32 parts), three MMAs per K-step, fp32 accumulation.
#include "tc_common.cuh"
#include "resamp_tc.h"

#include <algorithm>
#include <cstdlib>

namespace tsdgpu {
namespace rtc {
using namespace tc;

constexpr int TILE = 128, CH = 64, NCOL = 128, CHUNK = 32;
constexpr int NRAW = 4;                          // raw staging ring
constexpr int NSTAGE = 2;                        // A-operand stages in tensor memory
constexpr int NT = 2;                            // coefficient-block ring
constexpr int RAW_PITCH = 272, RAW_BYTES = CH * RAW_PITCH;
constexpr int TB_PART = TILE * 128, TB_BYTES = 2 * TB_PART;     // 128 rows x 32 tf32, hi + lo
constexpr int MAXSPAN = 16;                      // tiles per CTA: its slice of the schedule (16 KiB) sits in shared memory
constexpr int LUT_SMEM_MAX = 66 * 1024;          // the LUT too when it fits (64 taps x 257 phases = 64.25 KiB)
constexpr int SMEM_BYTES = NT * TB_BYTES + NRAW * RAW_BYTES + MAXSPAN * TILE * 8 + LUT_SMEM_MAX + 1024 + 512 + 2 * MAXSPAN * 4 + 64;
constexpr int ACOL = 3 * NCOL;
constexpr int CONV_WARP0 = 4, GEN_WARP0 = 8, NGEN = 16, MMA_WARP = GEN_WARP0 + NGEN, LOAD_WARP = MMA_WARP + 1;
constexpr int NTHREADS = 32 * (LOAD_WARP + 1);
constexpr int GROWS = (TILE + NGEN - 1) / NGEN;   // rows per generator warp and block: j = gw + NGEN * r
constexpr int TMEM_COLS = 512;

#ifdef TSD_TC_PROF
__device__ long long g_rtcprof[1024][24][4];
#define PROF_ARRAY g_rtcprof
#endif
#include "tc_prof.cuh"

__device__ __forceinline__ int floor_div32(int v) { return v >> 5; }   // arithmetic shift: floor for negatives too

// the (chunk, tile) block sequence: for chunk c, the tiles it feeds (at most two, consecutive)
struct Walk
{
  const int *cA, *cB;
  int T, tlo;
  __device__ __forceinline__ void feeds(int c, int &t0, int &t1)
  {
    while(tlo < T && cB[tlo] < c) tlo++;
    t0 = (tlo < T && cA[tlo] <= c) ? tlo : -1;
    t1 = (tlo + 1 < T && cA[tlo + 1] <= c) ? tlo + 1 : -1;
  }
};

template<bool LUTS> __global__ void __launch_bounds__(NTHREADS, 1) resamp_tc_kernel(ResampTcParams p)
{
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  unsigned char *sm = raw + (base - smem_u32(raw));
  unsigned char *tring = sm;                                  // [NT][hi 16 KiB | lo 16 KiB]
  unsigned char *stages = sm + NT * TB_BYTES;                 // [NRAW][64 rows x 272 B]
  int2 *sched_s = reinterpret_cast<int2 *>(stages + NRAW * RAW_BYTES);           // [T][128] schedule of this CTA's tiles
  float *lut_s = reinterpret_cast<float *>(sched_s + MAXSPAN * TILE);
  uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(lut_s) + LUT_SMEM_MAX);
  uint64_t *full = bars, *empty = full + NSTAGE, *tfull = empty + NSTAGE, *tempty = tfull + 3;
  uint64_t *rfull = tempty + 3, *rempty = rfull + NRAW, *bfull = rempty + NRAW, *bempty = bfull + NT;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bempty + NT);
  int *cA = reinterpret_cast<int *>(tmem_slot + 2), *cB = cA + MAXSPAN;
  int2 *bmeta = reinterpret_cast<int2 *>(cB + MAXSPAN);      // per coefficient-block slot: {first output column, columns}

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ts = blockIdx.x * p.span, te = min(ts + p.span, p.ntiles), T = te - ts;
  const int c0 = blockIdx.y * CH;
  const int K = p.K;

  if(tid == 0)
  {
    for(int i = 0; i < NSTAGE; i++) { mbar_init(full + i, 4); mbar_init(empty + i, 1); }
    for(int i = 0; i < 3; i++) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 4); }
    for(int i = 0; i < NRAW; i++) { mbar_init(rfull + i, 32); mbar_init(rempty + i, 4); }
    for(int i = 0; i < NT; i++) { mbar_init(bfull + i, NGEN); mbar_init(bempty + i, 1); }
    mbar_fence_init();
  }
  if(warp == MMA_WARP)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if(warp == 0 && lane < T)
  {
    // chunk range of tile ts + lane: first window sample ... newest input of its last output
    const int jf = (ts + lane) * TILE, jl = (int) min((long long) (jf + TILE), p.n_out) - 1;
    cA[lane] = floor_div32(p.sched[jf].x - (K - 1));
    cB[lane] = floor_div32(p.sched[jl].x);
  }
  for(int i = tid; i < T * TILE; i += NTHREADS)
  {
    const long long j = (long long) ts * TILE + i;
    // per output: {K-1 - in_j, p_j * K} (second word < 0 marks rows past the end): all the generators need per row
    int2 e = make_int2(0, -1);
    if(j < p.n_out) { const int2 q = __ldg(p.sched + j); e = make_int2(K - 1 - q.x, q.y * K); }
    sched_s[i] = e;
  }
  if(LUTS)
    for(int i = tid; i < p.lut_elems; i += NTHREADS) lut_s[i] = __ldg(p.lut + i);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  const int c_begin = cA[0], nchunks = cB[T - 1] - c_begin + 1;

  if(warp == LOAD_WARP)
  {
    // ===== loader: raw chunk (inputs [32 c, 32 c + 32) of 64 channels) -> staging slot, one 256-byte row per channel
    const int sp = lane & 15, clb = lane >> 4;
    for(int it = 0; it < nchunks;)
    {
      const long long pos0 = (long long) (c_begin + it) * CHUNK;
      const bool interior = pos0 >= 0 && c0 + CH <= p.nchan;
      const bool pair = interior && it + 1 < nchunks && pos0 + 2 * CHUNK <= p.n;
      mbar_wait(rempty + it % NRAW, (unsigned) (((it / NRAW) & 1) ^ 1));
      if(pair)
      {
        mbar_wait(rempty + (it + 1) % NRAW, (unsigned) ((((it + 1) / NRAW) & 1) ^ 1));
        const uint32_t da = smem_u32(stages + (it % NRAW) * RAW_BYTES + clb * RAW_PITCH + sp * 16);
        const uint32_t db = smem_u32(stages + ((it + 1) % NRAW) * RAW_BYTES + clb * RAW_PITCH + sp * 16);
        const float2 *src = p.x + (long long) (c0 + clb) * p.x_stride + pos0 + 2 * sp;
#pragma unroll 8
        for(int j = 0; j < 32; j++)
        {
          const float2 *sj = src + (long long) j * 2 * p.x_stride;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da + j * 2 * RAW_PITCH), "l"(sj) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(db + j * 2 * RAW_PITCH), "l"(sj + CHUNK) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(rfull + it % NRAW)) : "memory");
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(rfull + (it + 1) % NRAW)) : "memory");
        it += 2;
        continue;
      }
      // history / end of the call / ragged channel group: per-sample zero-filling copies (the history rows have odd length)
      const uint32_t dst0 = smem_u32(stages + (it % NRAW) * RAW_BYTES + clb * RAW_PITCH + sp * 16);
      for(int j = 0; j < 32; j++)
      {
        const int chan = c0 + clb + 2 * j;
#pragma unroll
        for(int e = 0; e < 2; e++)
        {
          const long long pos = pos0 + 2 * sp + e;
          const float2 *src = p.x;
          unsigned bytes = 0;
          if(chan < p.nchan)
          {
            if(pos >= 0) { if(pos < p.n) { src = p.x + (long long) chan * p.x_stride + pos; bytes = 8u; } }
            else if(pos >= -(long long) p.hist_len) { src = p.hist + (long long) chan * p.hist_len + p.hist_len + pos; bytes = 8u; }
          }
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst0 + j * 2 * RAW_PITCH + e * 8), "l"(src), "r"(bytes) : "memory");
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(rfull + it % NRAW)) : "memory");
      it += 1;
    }
  }
  else if(warp >= CONV_WARP0 && warp < CONV_WARP0 + 4)
  {
    // ===== converters (one warp per TMEM lane quadrant): raw row (channel, re|im) -> tf32 hi / lo -> tensor memory
    const int pw = warp - CONV_WARP0;
    const int my_cl = 8 * (2 * pw + (lane >> 4)) + (lane & 7), my_ri = (lane >> 3) & 1;
    PROF_DECL
    for(int it = 0; it < nchunks; it++)
    {
      const int slot = it % NRAW, stage = it & 1;
      PROF_BEGIN(t_w)
      mbar_wait(rfull + slot, (unsigned) ((it / NRAW) & 1));
      PROF_ADD(0, t_w)
      PROF_BEGIN(t_w2)
      mbar_wait(empty + stage, (unsigned) (((it >> 1) & 1) ^ 1));
      PROF_ADD(1, t_w2)
      PROF_BEGIN(t_c)
      fence_after();
      const uint32_t my_a = tmem + ((uint32_t) (pw * 32) << 16) + (uint32_t) (ACOL + 64 * stage);
      const unsigned char *row = stages + slot * RAW_BYTES + my_cl * RAW_PITCH;
#pragma unroll
      for(int hq = 0; hq < 2; hq++)
      {
        float hi[16], lo[16];
#pragma unroll
        for(int m = 0; m < 8; m++)
        {
          const float4 x = *reinterpret_cast<const float4 *>(row + (hq * 8 + m) * 16);
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
      if(lane == 0) mbar_arrive(full + stage);
      PROF_ADD(2, t_c)
    }
    PROF_END
  }
  else if(warp >= GEN_WARP0 && warp < GEN_WARP0 + NGEN)
  {
    // ===== coefficient-block generators: block (c, tile) row j = lut row p_j shifted to the chunk, lane = column kk
    const int gw = warp - GEN_WARP0;
    Walk wk{cA, cB, T, 0};
    int bseq = 0;
    PROF_DECL
    for(int it = 0; it < nchunks; it++)
    {
      const int c = c_begin + it;
      int tt[2];
      wk.feeds(c, tt[0], tt[1]);
#pragma unroll
      for(int w = 0; w < 2; w++)
      {
        if(tt[w] < 0) continue;
        const int slot = bseq % NT;
        PROF_BEGIN(t_w)
        mbar_wait(bempty + slot, (unsigned) (((bseq / NT) & 1) ^ 1));
        PROF_ADD(0, t_w)
        PROF_BEGIN(t_c)
        unsigned char *thi = tring + slot * TB_BYTES, *tlo_ = thi + TB_PART;
        // this warp's rows j = gw + 8 r: lane r < 16 fetches the schedule entry, broadcast row by row
        // band of this block: rows whose K taps overlap the chunk.  e.x = K-1 - in_j is non-increasing in j and the rows
        // past the end (e.y < 0) come last, so both conditions are prefix properties: count them with ballots.
        int jlo = 0, jend = 0;
#pragma unroll
        for(int q = 0; q < 4; q++)
        {
          const int2 e = sched_s[tt[w] * TILE + 32 * q + lane];
          jlo += __popc(__ballot_sync(0xffffffffu, e.y >= 0 && c * CHUNK + e.x > K - 1));     // window entirely after the chunk
          jend += __popc(__ballot_sync(0xffffffffu, e.y >= 0 && c * CHUNK + 31 + e.x >= 0));  // valid and not entirely before it
        }
        const int j0 = p.band ? (min(jlo, TILE - 16) & ~15) : 0;
        const int nn = p.band ? max(16, ((jend + 15) & ~15) - j0) : TILE;
        if(gw == 0 && lane == 0) bmeta[slot] = make_int2(j0, nn);
        // GROWS independent, branch-free rows per warp: clamped LUT index, value masked afterwards; lane = column kk
        const int2 *srow = sched_s + tt[w] * TILE + gw;
        const int tcol = c * CHUNK + lane;
        float v[GROWS];
#pragma unroll
        for(int r = 0; r < GROWS; r++)
        {
          const int2 e = srow[min(NGEN * r, TILE - 1 - gw)];           // broadcast read
          const int tap = tcol + e.x;
          const bool ok = (e.y >= 0) & ((unsigned) tap < (unsigned) K);
          const int idx = ok ? e.y + tap : 0;
          const float val = LUTS ? lut_s[idx] : __ldg(p.lut + idx);
          v[r] = ok ? val : 0.f;
        }
#pragma unroll
        for(int r = 0; r < GROWS; r++)
        {
          // rows outside the band are never read by the MMAs: skip their stores (warp-uniform predicate, no branch)
          const int j = gw + NGEN * r;
          const float hi = to_tf32(v[r]), lo = to_tf32(v[r] - hi);
          if(j < TILE && (unsigned) (j - j0) < (unsigned) nn)
          {
            const uint32_t off = swz((uint32_t) (j * 128 + lane * 4));
            *reinterpret_cast<float *>(thi + off) = hi;
            *reinterpret_cast<float *>(tlo_ + off) = lo;
          }
        }
        PROF_ADD(2, t_c)
        bseq++;
      }
      // one generic -> async proxy fence per chunk (it is the expensive part), then publish the chunk's blocks
      PROF_BEGIN(t_f)
      fence_proxy_async();
      __syncwarp();
      if(lane == 0)
      {
        const int nb = (tt[0] >= 0) + (tt[1] >= 0);
        for(int k = nb; k > 0; k--) mbar_arrive(bfull + (bseq - k) % NT);
      }
      PROF_ADD(1, t_f)
    }
    PROF_END
  }
  else if(warp == MMA_WARP)
  {
    // ===== MMA issuer
    const uint64_t dbase = smem_desc(0);
    Walk wk{cA, cB, T, 0};
    int bseq = 0;
    PROF_DECL
    for(int it = 0; it < nchunks; it++)
    {
      const int c = c_begin + it, stage = it & 1;
      int tt[2];
      wk.feeds(c, tt[0], tt[1]);
      PROF_BEGIN(t_w)
      mbar_wait(full + stage, (unsigned) ((it >> 1) & 1));
      PROF_ADD(0, t_w)
      const uint32_t xh0 = tmem + (uint32_t) (ACOL + 64 * stage), xl0 = xh0 + 32;
#pragma unroll
      for(int w = 0; w < 2; w++)
      {
        if(tt[w] < 0) continue;
        const int tl = tt[w], region = tl % 3, slot = bseq % NT;
        PROF_BEGIN(t_w2)
        mbar_wait(bfull + slot, (unsigned) ((bseq / NT) & 1));
        PROF_ADD(1, t_w2)
        PROF_BEGIN(t_c)
        if(c == cA[tl]) mbar_wait(tempty + region, (unsigned) ((tl / 3) & 1));   // first block of the tile: region drained and zeroed
        fence_after();
        const int2 meta = bmeta[slot];                    // {first output column, columns} of the block's band
        const uint32_t thi = base + slot * TB_BYTES + (uint32_t) meta.x * 128;
        const uint64_t bh0 = dbase + (thi >> 4), bl0 = bh0 + (TB_PART >> 4);
        const uint32_t dcol = tmem + (uint32_t) (region * NCOL + meta.x);
        const uint32_t idesc = IDESC_M128 | ((uint32_t) (meta.y >> 3) << 17);
        if(elect_one())
        {
#pragma unroll
          for(int ks = 0; ks < 4; ks++)
          {
            mma_tf32(dcol, xh0 + 8 * ks, bl0 + 2 * ks, idesc);
            mma_tf32(dcol, xl0 + 8 * ks, bh0 + 2 * ks, idesc);
            mma_tf32(dcol, xh0 + 8 * ks, bh0 + 2 * ks, idesc);
          }
          mma_commit(bempty + slot);
          if(c == cB[tl]) mma_commit(tfull + region);     // last block of the tile: accumulator complete
        }
        __syncwarp();
        PROF_ADD(2, t_c)
        bseq++;
      }
      if(elect_one()) mma_commit(empty + stage);
      __syncwarp();
    }
    PROF_END
  }
  else
  {
    // ===== epilogue (see fir_tc.cu): warp w owns TMEM lanes 32 w ... 32 w + 31 = channels 16 w ... 16 w + 15
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
    for(int tl = 0; tl < T; tl++)
    {
      const int region = tl % 3;
      mbar_wait(tfull + region, (unsigned) ((tl / 3) & 1));
      fence_after();
      const long long j0 = (long long) (ts + tl) * TILE + 2 * (lane & 3);
#pragma unroll
      for(int half = 0; half < 2; half++)
      {
        const int chan = c0 + 8 * (2 * warp + half) + (lane >> 2);
        float2 *yrow = p.y + (long long) chan * p.y_stride + p.out0;
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
              const long long jj = j0 + cb * 32 + 8 * i;
              const float2 o0 = make_float2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 2]));
              const float2 o1 = make_float2(__uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 3]));
              if(p.vec_store && jj + 1 < p.n_out) *reinterpret_cast<float4 *>(yrow + jj) = make_float4(o0.x, o0.y, o1.x, o1.y);
              else
              {
                if(jj < p.n_out) yrow[jj] = o0;
                if(jj + 1 < p.n_out) yrow[jj + 1] = o1;
              }
            }
          }
        }
      }
      zero_region(region);
      fence_before();
      __syncwarp();
      if(lane == 0) mbar_arrive(tempty + region);
    }
  }
  fence_before();
  __syncthreads();
  if(warp == MMA_WARP)
  {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

} // namespace rtc

// sched_host: the schedule of this chunk on the host.  Conditions: every chunk of 32 inputs feeds at most two tiles of
// 128 outputs (tile t + 2 starts after tile t ends), 16-byte aligned channel rows.
#ifdef TSD_TC_PROF
extern "C" int tsdgpu_debug_rtcprof_dump(const char *path)
{
  cudaDeviceSynchronize();
  static long long h[1024][24][4];
  if(cudaMemcpyFromSymbol(h, rtc::g_rtcprof, sizeof(h)) != cudaSuccess) return 1;
  FILE *fp = fopen(path, "wb");
  if(!fp) return 1;
  fwrite(h, 1, sizeof(h), fp);
  fclose(fp);
  return 0;
}
#endif

bool resamp_tc_eligible(const int2 *sched_host, long long n_out, int K, const void *x, long long x_stride)
{
  if(K < 1 || K > 4096 || n_out < 1) return false;
  if(((uintptr_t) x & 15) != 0 || (x_stride % 2) != 0) return false;
  const long long ntiles = (n_out + rtc::TILE - 1) / rtc::TILE;
  for(long long t = 0; t + 2 < ntiles; t++)
  {
    const int endc = sched_host[std::min<long long>(t * rtc::TILE + rtc::TILE, n_out) - 1].x >> 5;
    const int begc = (sched_host[(t + 2) * rtc::TILE].x - (K - 1)) >> 5;
    if(begc <= endc) return false;
  }
  return true;
}

int resamp_tc_launch(const ResampTcParams &p0)
{
  ResampTcParams p = p0;
  Runtime &r = rt();
  static bool attr_set = false;
  if(!attr_set)
  {
    TSD_CUDA(cudaFuncSetAttribute(rtc::resamp_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, rtc::SMEM_BYTES));
    TSD_CUDA(cudaFuncSetAttribute(rtc::resamp_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, rtc::SMEM_BYTES));
    attr_set = true;
  }
  p.ntiles = (int) ((p.n_out + rtc::TILE - 1) / rtc::TILE);
  p.band = !(getenv("TSDGPU_RESAMP_TC_BAND") && atoi(getenv("TSDGPU_RESAMP_TC_BAND")) == 0);
  const int groups = (p.nchan + rtc::CH - 1) / rtc::CH;
  int span = 1;
  long long best = -1;
  for(int s = 1; s <= rtc::MAXSPAN; s++)
  {
    if(s < 4 && p.ntiles > 4) continue;
    const long long ctas = (long long) groups * ((p.ntiles + s - 1) / s);
    const long long cost = ((ctas + r.num_sms - 1) / r.num_sms) * (s + 1);
    if(best < 0 || cost < best) { best = cost; span = s; }
  }
  p.span = span;
  p.vec_store = ((((uintptr_t) (p.y + p.out0)) & 15) == 0 && (p.y_stride % 2) == 0) ? 1 : 0;
  dim3 grid((p.ntiles + span - 1) / span, groups);
  if(p.lut_elems * 4 <= rtc::LUT_SMEM_MAX) rtc::resamp_tc_kernel<true><<<grid, rtc::NTHREADS, rtc::SMEM_BYTES, r.stream>>>(p);
  else rtc::resamp_tc_kernel<false><<<grid, rtc::NTHREADS, rtc::SMEM_BYTES, r.stream>>>(p);
  TSD_LAUNCH_CHECK();
  return 0;
}

} // namespace tsdgpu
