// Arbitrary-ratio LUT resampler as a banded filter-bank GEMM on the 5th-generation tensor cores (tcgen05 / TMEM), 3xTF32.
//
// Replaces the inner loop of AdaptationRythmeSimple::step + InterpolateurRIF::step (reference ra.cc:39-77,
// filtrage.hpp:1873-1881): out[c][j] = sum_i lut[p_j][i] * x[c][in_j - (K-1) + i], with (in_j, p_j) the host schedule of
// the reference's float32 phase recurrence (resamp.cu).  All channels share the schedule, so for a tile of 128
// consecutive outputs t and a chunk of 32 consecutive inputs c
//   D[n][j] += X[n][kk] * T[j][kk],   T[j][kk] = lut[p_j][32 c + kk - (in_j - (K-1))]  (0 outside the K taps)
// with D in tensor memory (lane n = one of the 128 real rows {re, im} x 64 channels, column j = output), X the
// de-interleaved input chunk (A operand, in tensor memory) and T the block of the banded coefficient matrix that
// sixteen generator warps build from the LUT and the schedule (B operand, shared memory, K-major, 128-byte swizzle).
// Same machine as fir_tc.cu (loader warp with LDGSTS staging ring, converter warps -> tcgen05.st, one elected MMA
// issuer, 16x256b epilogue, 3 accumulator regions); what differs is the B operand (generated per block instead of a
// view of one Toeplitz generator) and the irregular chunk <-> tile incidence: tile t is fed by the chunks
// floor((in_first - (K-1)) / 32) ... floor(in_last / 32), a chunk feeds one or two consecutive tiles (checked on the
// host: tile t+2 must start after tile t ends), every role walks the same (chunk, tile) block sequence, tabulated with
// the band of every block in the prologue.  Two CTAs of a cluster (adjacent 64-channel groups) run as a CTA pair
// (cta_group::2, M = 256): each builds half of the rows of every coefficient block.
// fp32 accuracy: x = x_hi + x_lo, T = T_hi + T_lo (tf32 parts), three MMAs per K-step, fp32 accumulation.
#include "tc_common.cuh"
#include "resamp_tc.h"
#include "tma_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace tsdgpu {
namespace rtc {
using namespace tc;

constexpr int TILE = 128, CH = 64, NCOL = 128, CHUNK = 32;
constexpr int NRAW = 4;                          // raw staging ring
constexpr int NSTAGE = 2;                        // A-operand stages in tensor memory
// coefficient-block ring: 2 blocks of 32 KiB (pair: 4 slots of half the rows) next to a LUT resident in shared memory;
// 4 blocks (pair: 8 half slots) when the LUT is read through L2 instead (LUTS = false: the ring takes the LUT's place)
constexpr int nt_of(bool luts) { return luts ? 2 : 4; }
constexpr int NT_MAX = 8;
constexpr int RAW_PITCH = 272, RAW_BYTES = CH * RAW_PITCH;
constexpr int TB_PART = TILE * 128, TB_BYTES = 2 * TB_PART;     // 128 rows x 32 tf32, hi + lo
#ifndef RTC_MAXSPAN
#define RTC_MAXSPAN 24                           // A/B builds: 12 = the span limit before the schedule slice was packed
#endif
constexpr int MAXSPAN = RTC_MAXSPAN;                    // tiles per CTA: its slice of the schedule (one packed word per output, 12 KiB) sits in shared memory
constexpr int LUT_SMEM_MAX = 64 * 1024 + 512;    // the LUT too when it fits (64 taps x 257 phases = 64.25 KiB)
constexpr int NB_MAX = 384;                       // (chunk, tile) blocks per CTA: their bands are tabulated in the prologue
// TMA form (round 2): raw slots are two swizzled tensor-map boxes [64 rows][128 B]; the epilogue warps own 4 KiB each of
// store staging (two buffers of [8 channel rows][32 outputs]) that they hand to the TMA unit
constexpr int RAW2_BYTES = 16384, OUT2_BYTES = 4096;
constexpr int smem_bytes(bool tma, bool luts)
{
  return nt_of(luts) * TB_BYTES + NRAW * (tma ? RAW2_BYTES : RAW_BYTES) + (tma ? 4 * OUT2_BYTES : 0) + MAXSPAN * TILE * 4 + (luts ? LUT_SMEM_MAX : 0) + 1024 + 512 +
         2 * MAXSPAN * 4 + 64 + NB_MAX * 8 + NB_MAX * 2;
}
static_assert(smem_bytes(true, true) <= 232448 - 1024 && smem_bytes(true, false) <= 232448 - 1024 && smem_bytes(false, false) <= 232448 - 1024, "shared memory");
constexpr int ACOL = 3 * NCOL;
#ifndef RTC_NCG
#define RTC_NCG 1                                 // converter groups of 4 warps; 2 (alternating chunks like fir_tc.cu, 30 warps at 64 registers) measured: no gain
#endif
constexpr int NCG = RTC_NCG;
#ifndef RTC_NGEN
#define RTC_NGEN 16                               // generator warps (20, at 64 registers: 177 vs 190 Gsamples/s; fewer than 24 prologue warps cannot tabulate 24-tile spans)
#endif
constexpr int CONV_WARP0 = 4, GEN_WARP0 = CONV_WARP0 + 4 * NCG, NGEN = RTC_NGEN, MMA_WARP = GEN_WARP0 + NGEN, LOAD_WARP = MMA_WARP + 1;
constexpr int NTHREADS = 32 * (LOAD_WARP + 1);
constexpr int NPRO = NTHREADS - 32;             // threads of the prologue: the loader (last warp) starts copying at once
constexpr int GROWS = (TILE + NGEN - 1) / NGEN;   // rows per generator warp and block: j = gw + NGEN * r
constexpr int TMEM_COLS = 512;

#ifdef TSD_TC_PROF
__device__ long long g_rtclife[8192][4];   // per CTA: start ns, end ns, SM id, cluster rank
__device__ long long g_rtcprof[1024][32][4];
#define PROF_ARRAY g_rtcprof
#endif
#include "tc_prof.cuh"

__device__ __forceinline__ int floor_div32(int v) { return v >> 5; }
// Schedule slice of a CTA in shared memory: one word per output, (in_j - in_0) << 10 | p_j (in_0 = newest input of the
// CTA's first output; resamp_tc_eligible checks the ranges), all ones for rows past the end.  Unpacked to what the
// generators need per row: {K-1 - in_j, p_j * K}, second word < 0 for rows past the end.
constexpr uint32_t SCHED_NONE = 0xFFFFFFFFu;
__device__ __forceinline__ int2 sched_unpack(uint32_t w, int base /* K-1 - in_0 */, int K)
{
  return w == SCHED_NONE ? make_int2(0, -1) : make_int2(base - (int) (w >> 10), (int) (w & 1023u) * K);
}

// ---- CTA pair (cta_group::2): two CTAs of a cluster = two groups of 64 channels over the same tiles.  The leader
// (cluster rank 0) issues M = 256 MMAs that read each CTA's own A rows from its tensor memory and half of the
// coefficient block from each CTA's shared memory, so every CTA generates only half of the block's rows.
constexpr uint32_t IDESC_M256 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (256 >> 4) << 24);
__device__ __forceinline__ uint32_t cluster_rank()
{
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the LEADER's copy of a barrier (same offset in the shared memory of cluster rank 0)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar)
{
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(ra) : "r"(smem_u32(bar)));
  // default semantics (release at CTA scope): what the peer's tensor core / the leader's MMA must see lives in THIS CTA's
  // shared / tensor memory and has been performed there before the arrive leaves; a cluster-scope release costs ~1 us
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, unsigned parity)
{
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "WAIT_%=:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
    "@p bra DONE_%=;\n\t"
    "bra WAIT_%=;\n\t"
    "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mma_tf32_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc)
{
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "setp.ne.b32 p, 1, 0;\n\t"
    "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
    ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc)
    : "memory");
}
// completion of all MMAs issued so far -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma_commit_pair(uint64_t *bar)
{
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((unsigned short) 3) : "memory");
}
template<bool PAIR> __device__ __forceinline__ void arrive_to_mma(uint64_t *bar)
{
  if(PAIR) mbar_arrive_leader(bar);
  else mbar_arrive(bar);
}
template<bool PAIR> __device__ __forceinline__ void commit_from_mma(uint64_t *bar)
{
  if(PAIR) mma_commit_pair(bar);
  else mma_commit(bar);
}
template<bool PAIR> __device__ __forceinline__ void wait_in_mma(uint64_t *bar, unsigned parity)
{
  if(PAIR) mbar_wait_cluster(bar, parity);
  else mbar_wait(bar, parity);
}   // arithmetic shift: floor for negatives too

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

// TMA = true (default when the output rows are 16-byte aligned): interior input chunks arrive as two 2-D tensor-map boxes
// {32 floats, 64 rows} with the 128-byte swizzle (UTMALDG; the end of the call and ragged channel groups are the map's zero
// fill; only the history chunks, whose rows have odd length, keep the LDGSTS path), and the epilogue warps store through
// the TMA unit (UTMASTG): 256 contiguous bytes per channel row and store instead of 64-byte pieces, no STG issue.
template<bool LUTS, bool PAIR, bool TMA> __global__ void __launch_bounds__(NTHREADS, 1)
resamp_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap ymap, ResampTcParams p)
{
  constexpr int RAWB = TMA ? RAW2_BYTES : RAW_BYTES;
  const uint32_t rank = PAIR ? cluster_rank() : 0u;
#ifdef TSD_TC_PROF
  const long long t_entry = clock64();
  unsigned long long gt_entry;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_entry));
#endif
  constexpr int NT = LUTS ? 2 : 4;   // = nt_of(LUTS)
  constexpr int NTR = PAIR ? 2 * NT : NT;                       // ring slots
  constexpr int TBP = PAIR ? TB_PART / 2 : TB_PART, TBB = 2 * TBP;   // bytes of the hi (= lo) part of a slot, of a slot
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  unsigned char *sm = raw + (base - smem_u32(raw));
  unsigned char *tring = sm;                                  // [NT][hi 16 KiB | lo 16 KiB]
  unsigned char *stages = sm + NT * TB_BYTES;                 // [NRAW][64 rows x 272 B]
  unsigned char *outs = stages + NRAW * RAWB;                 // TMA: [4 epilogue warps][2 buffers][8 rows x 256 B]
  uint32_t *sched_s = reinterpret_cast<uint32_t *>(outs + (TMA ? 4 * OUT2_BYTES : 0));   // [T][128] packed schedule of this CTA's tiles
  float *lut_s = reinterpret_cast<float *>(sched_s + MAXSPAN * TILE);
  uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(lut_s) + (LUTS ? LUT_SMEM_MAX : 0));
  uint64_t *full = bars, *empty = full + NSTAGE, *tfull = empty + NSTAGE, *tempty = tfull + 3;
  uint64_t *rfull = tempty + 3, *rempty = rfull + NRAW, *bfull = rempty + NRAW, *bempty = bfull + NT_MAX;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bempty + NT_MAX);
  int *cA = reinterpret_cast<int *>(tmem_slot + 2), *cB = cA + MAXSPAN;
  int2 *bandtab = reinterpret_cast<int2 *>(cB + MAXSPAN);    // per block of the walk: {first output column, columns} of its band
  unsigned short *ftab = reinterpret_cast<unsigned short *>(bandtab + NB_MAX);   // per chunk: tiles it feeds, t0 | t1 << 8 (0xff = none)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // pair: 1-D grid, consecutive CTAs (= the cluster) are the two channel groups 2g, 2g+1 of one tile span
  const int bspan = PAIR ? (int) blockIdx.x / p.groups : (int) blockIdx.x, bgroup = PAIR ? (int) blockIdx.x % p.groups : (int) blockIdx.y;
  const int ts = bspan * p.span, te = min(ts + p.span, p.ntiles), T = te - ts;
  const int c0 = bgroup * CH;
  const int K = p.K;

  if(tid == 0)
  {
    // pair: the barriers the MMA issuer waits on collect the arrivals of both CTAs (in the leader's shared memory)
    constexpr int NC = PAIR ? 2 : 1;
    for(int i = 0; i < NSTAGE; i++) { mbar_init(full + i, 4 * NC); mbar_init(empty + i, 1); }
    for(int i = 0; i < 3; i++) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 4 * NC); }
    for(int i = 0; i < NRAW; i++) { mbar_init(rfull + i, 32); mbar_init(rempty + i, 4); }   // a chunk is read by ONE converter group
    for(int i = 0; i < NTR; i++) { mbar_init(bfull + i, NGEN * NC); mbar_init(bempty + i, 1); }
    mbar_fence_init();
  }
  __syncthreads();   // barriers initialised (everybody is still at the top of the kernel)
  uint32_t tmem = 0;
  int c_begin = 0, nchunks = 0, in0 = 0, sbase = 0;
#ifdef TSD_TC_PROF
  long long t_pro = 0;
#endif
  if(warp == LOAD_WARP)
  {
    // The loader starts at once: its first chunks cross HBM while the other warps fill the tables, so every CTA begins
    // with a full staging ring.  It takes part in neither prologue barrier (named barrier 1 is for the other warps);
    // in a pair it arrives on the cluster barrier now and waits for it after its last copy.
    int v = 0;
    if(lane == 0) v = floor_div32(__ldg(&p.sched[(long long) ts * TILE].x) - (K - 1));
    if(lane == 1) v = floor_div32(__ldg(&p.sched[min((long long) te * TILE, p.n_out) - 1].x));
    c_begin = __shfl_sync(0xffffffffu, v, 0);
    nchunks = __shfl_sync(0xffffffffu, v, 1) - c_begin + 1;
    if(PAIR) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    // ===== loader: raw chunk (inputs [32 c, 32 c + 32) of 64 channels) -> staging slot, one 256-byte row per channel
    const int sp = lane & 15, clb = lane >> 4;
    for(int it = 0; it < nchunks;)
    {
      const long long pos0 = (long long) (c_begin + it) * CHUNK;
      const bool interior = pos0 >= 0 && c0 + CH <= p.nchan;
      const bool pair = interior && it + 1 < nchunks && pos0 + 2 * CHUNK <= p.n;
      mbar_wait_long(rempty + it % NRAW, (unsigned) (((it / NRAW) & 1) ^ 1));
      if(TMA && pos0 >= 0)
      {
        // one or two chunks through the TMA unit: lane 0 posts the byte count and issues the boxes, the others just arrive
        const int nk = it + 1 < nchunks ? 2 : 1;
        if(nk == 2) mbar_wait(rempty + (it + 1) % NRAW, (unsigned) ((((it + 1) / NRAW) & 1) ^ 1));
        for(int k = 0; k < nk; k++)
        {
          uint64_t *bar = rfull + (it + k) % NRAW;
          if(lane == 0)
          {
            const uint32_t dst = smem_u32(stages + ((it + k) % NRAW) * RAW2_BYTES);
            const int cx = 2 * ((int) pos0 + k * CHUNK);
            mbar_expect_tx(bar, RAW2_BYTES);
            tma_load_2d(dst, &xmap, cx, c0, bar);
            tma_load_2d(dst + 8192, &xmap, cx + 32, c0, bar);
          }
          else mbar_arrive(bar);
        }
        it += nk;
        continue;
      }
      if(!TMA && pair)
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
      const uint32_t dst0 = TMA ? smem_u32(stages + (it % NRAW) * RAW2_BYTES + (sp >> 3) * 8192)
                                : smem_u32(stages + (it % NRAW) * RAW_BYTES + clb * RAW_PITCH + sp * 16);
      for(int j = 0; j < 32; j++)
      {
        const int chan = c0 + clb + 2 * j;
        // TMA layout: row r of the half-chunk box at 128 r, 16-byte piece q at (q ^ (r & 7))
        const uint32_t dstj = TMA ? dst0 + (uint32_t) (clb + 2 * j) * 128u + ((uint32_t) ((sp & 7) ^ ((clb + 2 * j) & 7)) << 4) : dst0 + j * 2 * RAW_PITCH;
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
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dstj + e * 8), "l"(src), "r"(bytes) : "memory");
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(rfull + it % NRAW)) : "memory");
      it += 1;
    }
    if(PAIR) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    goto done;
  }
  if(warp == MMA_WARP)
  {
    if(PAIR)
    {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    else
    {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if(warp == 0 && lane < T)
  {
    // chunk range of tile ts + lane: first window sample ... newest input of its last output
    const int jf = (ts + lane) * TILE, jl = (int) min((long long) (jf + TILE), p.n_out) - 1;
    cA[lane] = floor_div32(p.sched[jf].x - (K - 1));
    cB[lane] = floor_div32(p.sched[jl].x);
  }
  in0 = __ldg(&p.sched[(long long) ts * TILE].x);
  sbase = K - 1 - in0;
  for(int i = tid; i < T * TILE; i += NPRO)
  {
    const long long j = (long long) ts * TILE + i;
    uint32_t e = SCHED_NONE;
    if(j < p.n_out) { const int2 q = __ldg(p.sched + j); e = ((uint32_t) (q.x - in0) << 10) | (uint32_t) q.y; }
    sched_s[i] = e;
  }
  if(LUTS)
  {
    // the LUT comes in with 16-byte loads when its size allows (cudaMalloc'ed: 256-byte aligned)
    const int n4 = (p.lut_elems % 4 == 0) ? p.lut_elems / 4 : 0;
    for(int i = tid; i < n4; i += NPRO) reinterpret_cast<float4 *>(lut_s)[i] = __ldg(reinterpret_cast<const float4 *>(p.lut) + i);
    for(int i = 4 * n4 + tid; i < p.lut_elems; i += NPRO) lut_s[i] = __ldg(p.lut + i);
  }
  named_bar(1, NPRO);
  {
    // tiles fed by every chunk of this CTA (at most two, consecutive): the first tile whose last chunk is >= c, and
    // its successor, when they have started
    const int cb0 = cA[0], nch = cB[T - 1] - cb0 + 1;
    for(int it = tid; it < nch; it += NPRO)
    {
      const int c = cb0 + it;
      int tlo = 0;
      while(tlo < T && cB[tlo] < c) tlo++;
      const int t0 = (tlo < T && cA[tlo] <= c) ? tlo : 0xff;
      const int t1 = (t0 != 0xff && tlo + 1 < T && cA[tlo + 1] <= c) ? tlo + 1 : 0xff;
      ftab[it] = (unsigned short) (t0 | (t1 << 8));
    }
  }
  if(warp < T)
  {
    // bands of the blocks of tile `warp`, one lane per chunk.  A block (c, tl) keeps the rows whose K taps overlap the
    // chunk: e.x = K-1 - in_j is non-increasing in j and the rows past the end (e.y < 0) come last, so "window entirely
    // after the chunk" and "valid and not entirely before it" are prefix properties -> two binary searches.
    // Position in the walk (chunks ascending, then tiles): blocks of earlier chunks + the other tile of this chunk.
    const int tl = warp;
    const uint32_t *srow = sched_s + tl * TILE;
    for(int c = cA[tl] + lane; c <= cB[tl]; c += 32)
    {
      int lo = 0, hi = TILE;
      while(lo < hi) { const int m = (lo + hi) >> 1; const int2 e = sched_unpack(srow[m], sbase, K); if(e.y >= 0 && c * CHUNK + e.x > K - 1) lo = m + 1; else hi = m; }
      const int jlo = lo;
      hi = TILE;
      while(lo < hi) { const int m = (lo + hi) >> 1; const int2 e = sched_unpack(srow[m], sbase, K); if(e.y >= 0 && c * CHUNK + 31 + e.x >= 0) lo = m + 1; else hi = m; }
      const int jend = lo;
      // band [j0, j0 + nn): multiple of 16 columns (32 for a pair: each CTA supplies nn / 2 rows of the block)
      constexpr int GRAN = PAIR ? 32 : 16;
      int j0 = p.band ? (min(jlo, TILE - GRAN) & ~15) : 0;
      const int nn = p.band ? max(GRAN, (jend - j0 + GRAN - 1) & ~(GRAN - 1)) : TILE;
      if(j0 + nn > TILE) j0 = TILE - nn;
      int seq = 0;
      for(int t = 0; t < T; t++) seq += min(max(c - cA[t], 0), cB[t] - cA[t] + 1);
      if(tl > 0 && cA[tl - 1] <= c && c <= cB[tl - 1]) seq++;
      bandtab[seq] = make_int2(j0, nn);
    }
  }
  fence_before();
  if(PAIR) cluster_sync_all();   // the peer's barriers are initialised before anyone arrives on them
  else named_bar(1, NPRO);
  fence_after();
  tmem = *tmem_slot;
  c_begin = cA[0];
  nchunks = cB[T - 1] - c_begin + 1;
#ifdef TSD_TC_PROF
  t_pro = clock64();
#endif

  if(warp >= CONV_WARP0 && warp < CONV_WARP0 + 4 * NCG)
  {
    // ===== converters (one warp per TMEM lane quadrant, NCG groups alternating chunks): raw row (channel, re|im) -> tf32
    // hi / lo -> tensor memory
    const int pw = (warp - CONV_WARP0) & 3, cgrp = (warp - CONV_WARP0) >> 2;
    const int my_cl = 8 * (2 * pw + (lane >> 4)) + (lane & 7), my_ri = (lane >> 3) & 1;
    PROF_DECL
    for(int it = cgrp; it < nchunks; it += NCG)
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
      const unsigned char *row = TMA ? stages + slot * RAW2_BYTES + my_cl * 128 : stages + slot * RAW_BYTES + my_cl * RAW_PITCH;
      const uint32_t sx = (uint32_t) (my_cl & 7);
#pragma unroll
      for(int hq = 0; hq < 2; hq++)
      {
        float hi[16], lo[16];
#pragma unroll
        for(int m = 0; m < 8; m++)
        {
          const float4 x = TMA ? *reinterpret_cast<const float4 *>(row + hq * 8192 + (((uint32_t) m ^ sx) << 4))
                               : *reinterpret_cast<const float4 *>(row + (hq * 8 + m) * 16);
          const float a0 = my_ri ? x.y : x.x, a1 = my_ri ? x.w : x.z;
          hi[2 * m] = to_tf32(a0);
          hi[2 * m + 1] = to_tf32(a1);
          lo[2 * m] = a0 - hi[2 * m];             // exact; the tensor core truncates it to tf32 (error <= 2^-21 |x|)
          lo[2 * m + 1] = a1 - hi[2 * m + 1];
        }
        tmem_st16(my_a + hq * 16, hi);
        tmem_st16(my_a + 32 + hq * 16, lo);
      }
      __syncwarp();
      if(lane == 0) mbar_arrive(rempty + slot);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      fence_before();
      __syncwarp();
      if(lane == 0) arrive_to_mma<PAIR>(full + stage);
      PROF_ADD(2, t_c)
    }
    PROF_END
  }
  else if(warp >= GEN_WARP0 && warp < GEN_WARP0 + NGEN)
  {
    // ===== coefficient-block generators: block (c, tile) row j = lut row p_j shifted to the chunk, lane = column kk
    const int gw = warp - GEN_WARP0;
    int bseq = 0;
    constexpr int GR = PAIR ? (TILE / 2 + NGEN - 1) / NGEN : GROWS;
    // pair: the two blocks of a chunk are built together (their loads interleave); single CTA: one after the other
    constexpr int NBK = PAIR ? 2 : 1;
    PROF_DECL
    for(int it = 0; it < nchunks; it++)
    {
      const int c = c_begin + it;
      const int f = ftab[it];
      const int tf[2] = {f & 0xff, (f >> 8) & 0xff};
      const int nb = (tf[0] != 0xff) + (tf[1] != 0xff);
      const int tcol = c * CHUNK + lane;
      for(int b0 = 0; b0 < nb; b0 += NBK)
      {
        unsigned char *thi[NBK];
        const uint32_t *srow[NBK];
        int nh[NBK], jb[NBK];
#pragma unroll
        for(int k = 0; k < NBK; k++)
        {
          const bool on = b0 + k < nb;                       // second block of the group absent: built nowhere (nh = 0)
          const int bs = bseq + b0 + (on ? k : 0), slot = bs % NTR;
          PROF_BEGIN(t_w)
          if(on) mbar_wait(bempty + slot, (unsigned) (((bs / NTR) & 1) ^ 1));
          PROF_ADD(0, t_w)
          const int2 meta = bandtab[bs];                     // {first output column, columns}, tabulated in the prologue
          // this CTA's rows of the block: nh rows from jb on, stored band-relative (row l at 128 l, swizzled); warp gw
          // builds rows l = gw + NGEN r
          const int nhk = PAIR ? meta.y >> 1 : meta.y;
          nh[k] = on ? nhk : 0;
          jb[k] = meta.x + (PAIR ? (int) rank * nhk : 0);
          thi[k] = tring + slot * TBB;
          srow[k] = sched_s + tf[on ? b0 + k : b0] * TILE;
        }
        PROF_BEGIN(t_c)
        // branch-free rows: clamped LUT index, value masked afterwards; lane = column kk
        float v[NBK][GR];
#pragma unroll
        for(int k = 0; k < NBK; k++)
#pragma unroll
          for(int r = 0; r < GR; r++)
          {
            const int2 e = sched_unpack(srow[k][min(jb[k] + gw + NGEN * r, TILE - 1)], sbase, K);      // broadcast read
            const int tap = tcol + e.x;
            const bool ok = (e.y >= 0) & ((unsigned) tap < (unsigned) K);
            const int idx = ok ? e.y + tap : 0;
            const float val = LUTS ? lut_s[idx] : __ldg(p.lut + idx);
            v[k][r] = ok ? val : 0.f;
          }
#pragma unroll
        for(int k = 0; k < NBK; k++)
#pragma unroll
          for(int r = 0; r < GR; r++)
          {
            const int l = gw + NGEN * r;
            const float hi = to_tf32(v[k][r]), lo = v[k][r] - hi;
            if(l < nh[k])                                                    // warp-uniform predicate, no branch
            {
              const uint32_t off = swz((uint32_t) (l * 128 + lane * 4));
              *reinterpret_cast<float *>(thi[k] + off) = hi;
              *reinterpret_cast<float *>(thi[k] + TBP + off) = lo;
            }
          }
        PROF_ADD(2, t_c)
      }
      bseq += nb;
      // one generic -> async proxy fence per chunk (it is the expensive part), then publish the chunk's blocks
      PROF_BEGIN(t_f)
      fence_proxy_async();
      __syncwarp();
      if(lane == 0)
        for(int k = nb; k > 0; k--) arrive_to_mma<PAIR>(bfull + (bseq - k) % NTR);
      PROF_ADD(1, t_f)
    }
    PROF_END
  }
  else if(warp == MMA_WARP)
  {
    // ===== MMA issuer (pair: the leader CTA only)
    if(PAIR && rank != 0) goto done;
    const uint64_t dbase = smem_desc(0);
    int bseq = 0;
    PROF_DECL
    for(int it = 0; it < nchunks; it++)
    {
      const int c = c_begin + it, stage = it & 1;
      const int f = ftab[it];
      const int tt[2] = {(f & 0xff) == 0xff ? -1 : (f & 0xff), ((f >> 8) & 0xff) == 0xff ? -1 : ((f >> 8) & 0xff)};
      PROF_BEGIN(t_w)
      wait_in_mma<PAIR>(full + stage, (unsigned) ((it >> 1) & 1));
      PROF_ADD(0, t_w)
      const uint32_t xh0 = tmem + (uint32_t) (ACOL + 64 * stage), xl0 = xh0 + 32;
#pragma unroll
      for(int w = 0; w < 2; w++)
      {
        if(tt[w] < 0) continue;
        const int tl = tt[w], region = tl % 3, slot = bseq % NTR;
        // descriptors first (shared-memory reads, uniform arithmetic), then the waits
        const int2 meta = bandtab[bseq];                  // {first output column, columns} of the block's band
        const bool first = c == cA[tl], last = c == cB[tl];
        const uint32_t thi = base + slot * TBB;           // rows are stored band-relative
        const uint64_t bh0 = dbase + (thi >> 4), bl0 = bh0 + (TBP >> 4);
        const uint32_t dcol = tmem + (uint32_t) (region * NCOL + meta.x);
        const uint32_t idesc = (PAIR ? IDESC_M256 : IDESC_M128) | ((uint32_t) (meta.y >> 3) << 17);
        PROF_BEGIN(t_w2)
        wait_in_mma<PAIR>(bfull + slot, (unsigned) ((bseq / NTR) & 1));
        PROF_ADD(1, t_w2)
        PROF_BEGIN(t_c)
        if(first) wait_in_mma<PAIR>(tempty + region, (unsigned) ((tl / 3) & 1));   // first block of the tile: region drained and zeroed
        fence_after();
        if(elect_one())
        {
#pragma unroll
          for(int ks = 0; ks < 4; ks++)
          {
            if(PAIR)
            {
              mma_tf32_pair(dcol, xh0 + 8 * ks, bl0 + 2 * ks, idesc);
              mma_tf32_pair(dcol, xl0 + 8 * ks, bh0 + 2 * ks, idesc);
              mma_tf32_pair(dcol, xh0 + 8 * ks, bh0 + 2 * ks, idesc);
            }
            else
            {
              mma_tf32(dcol, xh0 + 8 * ks, bl0 + 2 * ks, idesc);
              mma_tf32(dcol, xl0 + 8 * ks, bh0 + 2 * ks, idesc);
              mma_tf32(dcol, xh0 + 8 * ks, bh0 + 2 * ks, idesc);
            }
          }
          commit_from_mma<PAIR>(bempty + slot);
          if(last) commit_from_mma<PAIR>(tfull + region);     // last block of the tile: accumulator complete
        }
        __syncwarp();
        PROF_ADD(2, t_c)
        bseq++;
      }
      if(elect_one()) commit_from_mma<PAIR>(empty + stage);
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
      if(lane == 0) arrive_to_mma<PAIR>(tempty + region);
    }
    unsigned char *my_out = outs + warp * OUT2_BYTES;
    unsigned char *my_piece = my_out + (lane >> 2) * 256 + (lane & 3) * 16;
    unsigned nst = 0;   // stores issued by this warp: buffer = parity
    for(int tl = 0; tl < T; tl++)
    {
      const int region = tl % 3;
      mbar_wait_long(tfull + region, (unsigned) ((tl / 3) & 1));
      fence_after();
      const long long j0 = (long long) (ts + tl) * TILE + 2 * (lane & 3);
      if(TMA)
      {
        // two tcgen05.ld per wait; per 32 outputs x 8 channels: 4 STS.128 per thread into one of the warp's two staging
        // buffers, then one tensor-map store of the box {64 floats, 8 rows} (clipped at n_out / nchan by the map)
#pragma unroll
        for(int half = 0; half < 2; half++)
#pragma unroll
          for(int cp = 0; cp < 2; cp++)
          {
            uint32_t r[2][16];
#pragma unroll
            for(int q = 0; q < 2; q++)
            {
              const uint32_t taddr = tmem + ((uint32_t) (warp * 32 + half * 16) << 16) + (uint32_t) (region * NCOL + (2 * cp + q) * 32);
              asm volatile(
                "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[q][0]), "=r"(r[q][1]), "=r"(r[q][2]), "=r"(r[q][3]), "=r"(r[q][4]), "=r"(r[q][5]), "=r"(r[q][6]), "=r"(r[q][7]),
                  "=r"(r[q][8]), "=r"(r[q][9]), "=r"(r[q][10]), "=r"(r[q][11]), "=r"(r[q][12]), "=r"(r[q][13]), "=r"(r[q][14]), "=r"(r[q][15])
                : "r"(taddr)
                : "memory");
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for(int q = 0; q < 2; q++)
            {
              const uint32_t boff = (nst & 1u) * 2048u;
              if(lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store before last has read this buffer
              __syncwarp();
#pragma unroll
              for(int i = 0; i < 4; i++)
                *reinterpret_cast<uint4 *>(my_piece + boff + 64 * i) = make_uint4(r[q][4 * i], r[q][4 * i + 2], r[q][4 * i + 1], r[q][4 * i + 3]);
              fence_proxy_async();
              __syncwarp();
              if(lane == 0)
              {
                tma_store_2d(&ymap, 2 * ((ts + tl) * TILE + (2 * cp + q) * 32), c0 + 16 * warp + 8 * half, smem_u32(my_out + boff));
                bulk_commit();
              }
              nst++;
            }
          }
      }
      else
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
      if(lane == 0) arrive_to_mma<PAIR>(tempty + region);
    }
    if(TMA && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
done:
#ifdef TSD_TC_PROF
  const long long t_fin = clock64();
#endif
  fence_before();
  if(PAIR) cluster_sync_all();   // nobody leaves (or frees tensor memory) while the pair's MMAs and remote arrivals are in flight
  else __syncthreads();
  if(warp == MMA_WARP)
  {
    fence_after();
    if(PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
#ifdef TSD_TC_PROF
  if(tid == 0 && blockIdx.y == 0 && blockIdx.x < 8192)
  {
    unsigned long long gt_end;
    unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_end));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_rtclife[blockIdx.x][0] = (long long) gt_entry;
    g_rtclife[blockIdx.x][1] = (long long) gt_end;
    g_rtclife[blockIdx.x][2] = smid;
    g_rtclife[blockIdx.x][3] = rank;
  }
  if(tid == 0 && blockIdx.y == 0 && blockIdx.x < 1024)
  {
    g_rtcprof[blockIdx.x][31][0] = t_pro - t_entry;     // prologue
    g_rtcprof[blockIdx.x][31][1] = t_fin - t_pro;       // epilogue warp 0: its whole role
    g_rtcprof[blockIdx.x][31][2] = clock64() - t_fin;   // final sync
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_rtcprof[blockIdx.x][31][3] = (long long) gt;      // end time (ns)
  }
#endif
}

} // namespace rtc

// sched_host: the schedule of this chunk on the host.  Conditions: every chunk of 32 inputs feeds at most two tiles of
// 128 outputs (tile t + 2 starts after tile t ends), 16-byte aligned channel rows.
#ifdef TSD_TC_PROF
extern "C" int tsdgpu_debug_rtcprof_dump(const char *path)
{
  cudaDeviceSynchronize();
  static long long h[1024][32][4];
  if(cudaMemcpyFromSymbol(h, rtc::g_rtcprof, sizeof(h)) != cudaSuccess) return 1;
  FILE *fp = fopen(path, "wb");
  if(!fp) return 1;
  fwrite(h, 1, sizeof(h), fp);
  static long long l[8192][4];
  if(cudaMemcpyFromSymbol(l, rtc::g_rtclife, sizeof(l)) != cudaSuccess) return 1;
  fwrite(l, 1, sizeof(l), fp);
  fclose(fp);
  return 0;
}
#endif

bool resamp_tc_eligible(const int2 *sched_host, long long n_out, int K, int nphases, const void *x, long long x_stride, int *max_tile_chunks)
{
  if(K < 1 || K > 4096 || n_out < 1) return false;
  if(nphases + 1 > 1024) return false;   // packed schedule word: 10 bits of LUT row
  if(((uintptr_t) x & 15) != 0 || (x_stride % 2) != 0) return false;
  const long long ntiles = (n_out + rtc::TILE - 1) / rtc::TILE;
  // chunks per tile: the CTA tabulates the bands of its (chunk, tile) blocks, at most NB_MAX of them
  int mc = 1;
  for(long long t = 0; t < ntiles; t++)
  {
    const int endc = sched_host[std::min<long long>(t * rtc::TILE + rtc::TILE, n_out) - 1].x >> 5;
    const int begc = (sched_host[t * rtc::TILE].x - (K - 1)) >> 5;
    mc = std::max(mc, endc - begc + 1);
    // packed schedule word: 22 bits of input offset within a CTA's span of <= MAXSPAN tiles
    if((long long) sched_host[std::min<long long>(t * rtc::TILE + rtc::TILE, n_out) - 1].x - sched_host[t * rtc::TILE].x >= (1 << 21) / rtc::MAXSPAN) return false;
  }
  if(mc > rtc::NB_MAX) return false;
  *max_tile_chunks = mc;
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
  if(!r.resamp_tc_ready)
  {
#define RTC_ATTR(L, P, T) TSD_CUDA(cudaFuncSetAttribute(rtc::resamp_tc_kernel<L, P, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, rtc::smem_bytes(T, L)));
    RTC_ATTR(true, false, false) RTC_ATTR(false, false, false) RTC_ATTR(true, true, false) RTC_ATTR(false, true, false)
    RTC_ATTR(true, false, true) RTC_ATTR(false, false, true) RTC_ATTR(true, true, true) RTC_ATTR(false, true, true)
#undef RTC_ATTR
    r.resamp_tc_ready = true;
  }
  p.ntiles = (int) ((p.n_out + rtc::TILE - 1) / rtc::TILE);
  p.band = !(getenv("TSDGPU_RESAMP_TC_BAND") && atoi(getenv("TSDGPU_RESAMP_TC_BAND")) == 0);
  const int groups = (p.nchan + rtc::CH - 1) / rtc::CH;
  int span = 1;
  double best = -1;
  const int smax = std::max(1, std::min(rtc::MAXSPAN, rtc::NB_MAX / std::max(1, p.max_tile_chunks)));
  for(int s = 1; s <= smax; s++)
  {
    if(s < std::min(4, smax) && p.ntiles > 4) continue;
    const long long ctas = (long long) groups * ((p.ntiles + s - 1) / s);
    // CTAs per SM x (tiles + set-up / drain of a CTA, about 1.5 tile times: 9 us vs 6 us per tile measured).  With
    // many CTAs per SM the SMs drift apart (each takes the next CTA when it is free): no rounding up to whole waves;
    // with few, the launch does run in waves
    const double per_sm = (double) ctas / r.num_sms;
    const double cost = (per_sm >= 8.0 ? per_sm : std::ceil(per_sm)) * (2 * s + 3);
    if(best < 0 || cost <= best) { best = cost; span = s; }
  }
  p.span = span;
  p.groups = groups;
  p.vec_store = ((((uintptr_t) (p.y + p.out0)) & 15) == 0 && (p.y_stride % 2) == 0) ? 1 : 0;
  dim3 grid((p.ntiles + span - 1) / span, groups);
  // TSDGPU_RESAMP_TC_LUTS=0: LUT through L2 + the deeper coefficient ring even when the LUT would fit shared memory (A/B)
  const bool luts = p.lut_elems * 4 <= rtc::LUT_SMEM_MAX && !(getenv("TSDGPU_RESAMP_TC_LUTS") && atoi(getenv("TSDGPU_RESAMP_TC_LUTS")) == 0);
  // CTA pairs (cta_group::2, clusters of 2 along the channel groups) whenever the groups pair up
  const bool pair = (groups % 2 == 0) && !(getenv("TSDGPU_RESAMP_TC_PAIR") && atoi(getenv("TSDGPU_RESAMP_TC_PAIR")) == 0);
  // tensor maps (TSDGPU_RESAMP_TC_TMA=0 keeps the LDGSTS / STG form): x as float32 rows [nchan][2 n]; y from this chunk's
  // first output on, [nchan][2 n_out]
  CUtensorMap xmap, ymap;
  memset(&xmap, 0, sizeof(xmap));
  memset(&ymap, 0, sizeof(ymap));
  const bool tma = p.vec_store && !(getenv("TSDGPU_RESAMP_TC_TMA") && atoi(getenv("TSDGPU_RESAMP_TC_TMA")) == 0) &&
                   (unsigned long long) p.x_stride * 8 < (1ull << 40) && (unsigned long long) p.y_stride * 8 < (1ull << 40) &&
                   tma_map_rows(&xmap, p.x, 2ull * p.n, p.nchan, (unsigned long long) p.x_stride * 8, 32, rtc::CH, true) &&
                   tma_map_rows(&ymap, p.y + p.out0, 2ull * p.n_out, p.nchan, (unsigned long long) p.y_stride * 8, 64, 8, false);
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(rtc::NTHREADS);
  cfg.dynamicSmemBytes = rtc::smem_bytes(tma, luts);
  cfg.stream = r.stream;
  cudaLaunchAttribute at[1];
  if(pair)
  {
    cfg.gridDim = dim3(grid.x * grid.y);
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
  }
  else cfg.gridDim = grid;
#define RTC_GO(L, P, T) TSD_CUDA(cudaLaunchKernelEx(&cfg, rtc::resamp_tc_kernel<L, P, T>, xmap, ymap, p))
  if(tma)
  {
    if(pair) { if(luts) RTC_GO(true, true, true); else RTC_GO(false, true, true); }
    else { if(luts) RTC_GO(true, false, true); else RTC_GO(false, false, true); }
  }
  else
  {
    if(pair) { if(luts) RTC_GO(true, true, false); else RTC_GO(false, true, false); }
    else { if(luts) RTC_GO(true, false, false); else RTC_GO(false, false, false); }
  }
#undef RTC_GO
  TSD_LAUNCH_CHECK();
  return 0;
}

} // namespace tsdgpu
