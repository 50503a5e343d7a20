// 65536-point FFT plan, TMA-fed persistent form (sm_100a).
//
// Same four-step decomposition and the same tile arithmetic as fft.cu (65536 = 256 x 256; stage A = 16 columns x 256
// rows -> L2-resident scratch, stage B = 16 rows x 256 columns -> natural-order output; reference: tfr_radix2,
// fourier.cc:61-121), but the math warps never touch global memory on the load side:
//   * one persistent CTA per SM = 3 consumer groups of 256 threads + a producer warp + a scout warp, a ring of six 32 KiB
//     tile slots in shared memory;
//   * the producer's elected lane walks the CTA's (static, round-robin) item list and fills the ring ahead of the
//     consumers: a stage-A tile (256 row pieces of 128 bytes, 512 KiB apart in nothing but the row stride) is ONE
//     3-D tensor-map copy (cp.async.bulk.tensor, SASS UTMALDG), a stage-B tile (16 consecutive scratch rows) one 32 KiB
//     1-D bulk copy (UBLKCP); completion on the slot's mbarrier.  All inter-CTA waiting (ld.acquire on the
//     per-transform counters) is done by the scout lane, so that neither a math warp nor the copy-issuing lane ever spins;
//   * a consumer group runs both radix-16 register passes of its tile IN PLACE in the slot (stage A: a thread writes the
//     exchange values onto the very locations it read, one group barrier; stage B: skewed exchange, two group barriers),
//     releases the slot to the producer and streams the results out with plain coalesced 128-byte-row stores.
// The grid is cut into sets of 16 CTAs; set s owns the transforms t = s (mod nsets), CTA g of the set tile g of each of
// them, once as a stage-A and once as a stage-B item.  A transform therefore lives inside one set: its 16 column tiles are
// loaded together (the 128-byte row pieces of the 16 tiles make up full 2 KiB rows), its stage-B tiles become ready as soon
// as those 16 CTAs are through, and each set needs only a short private scratch ring (the whole scratch stays in L2; dealing
// tiles round-robin over all CTAs instead spread the transforms in flight over ~100 ring slots).  Each producer keeps one
// cursor per stage and prefers a ready stage-B tile; a not-yet-ready one never blocks the stage-A prefetch behind it.  Every
// CTA is resident, stage-B tiles only wait for stage-A tiles of the same transform and stage-A tiles only for stage-B tiles
// `ring` transforms back in the same set, both issued in increasing order per CTA: the schedule cannot deadlock.
#include "common.cuh"
#include "fft_tiles.cuh"
#include "fft_plan.h"
#include "tc_common.cuh"

#include <cuda.h>
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace tsdgpu {

constexpr int FP_GROUPS = 3, FP_SLOTS = 6;
constexpr int FP_THREADS = FP_GROUPS * 256 + 96;   // + producer warp + scout warp + releaser warp
constexpr int FP_HIST = 64;                          // descriptors kept for the releaser
constexpr int FP_SLOT_BYTES = 32768;
constexpr size_t FP_SMEM = (size_t) FP_SLOTS * FP_SLOT_BYTES + 256 * sizeof(float4) + 2 * FP_SLOTS * sizeof(uint64_t) + 128 + FP_HIST * sizeof(int);

struct FftPipeParams
{
  float2 *y;
  long long y_stride;
  float2 *scratch;        // [ring][65536]
  unsigned *done_a;       // [batch] warps that finished a stage-A tile (16 tiles x 8 warps per transform)
  unsigned *done_b;       // [batch] stage-B tiles whose scratch tile has been read (x 8)
  int batch, ring, hints;
  const float4 *tw;       // rt().tw256
  const float2 *tw4;      // rt().tw4
  unsigned long long *prof;   // optional per-CTA wait accounting (TSDGPU_FFT_PROF), else null: [cta][8]
  int *trace;                 // optional producer timeline of CTA 0: [item][6] = clock before / after the readiness wait, stage, ia, ib, ready_a | ready_b << 16
};

__device__ __forceinline__ void tma_load_3d(void *dst_smem, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                 smem_u32(dst_smem)),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void stg_hint(float2 *p, float2 v, uint64_t pol)
{
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(void *dst_smem, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar, uint64_t pol)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(
                 smem_u32(dst_smem)),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p)
{
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta(int *p, int v) { asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_volatile_s(const int *p)
{
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ int ld_acquire_cta_s(const int *p)
{
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void red_relaxed_add(unsigned *p, unsigned v) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// a consumer warp has issued all stores of its tile: CTA-scope release on the group's counter (the releaser lane turns it
// into the gpu-scope publication)
__device__ __forceinline__ void tile_done(int *cnt)
{
  __syncwarp();
  if((threadIdx.x & 31) == 0) asm volatile("red.release.cta.shared.add.s32 [%0], 1;" ::"r"(smem_u32(cnt)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, unsigned parity)
{
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_acquire_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

template<bool INV>
__global__ void __launch_bounds__(FP_THREADS, 1) fft64k_pipe_kernel(const __grid_constant__ CUtensorMap xmap, FftPipeParams p)
{
  extern __shared__ __align__(1024) unsigned char fp_smem[];
  float2 *slots = reinterpret_cast<float2 *>(fp_smem);
  float4 *tw = reinterpret_cast<float4 *>(fp_smem + (size_t) FP_SLOTS * FP_SLOT_BYTES);
  uint64_t *full = reinterpret_cast<uint64_t *>(tw + 256);
  uint64_t *empty = full + FP_SLOTS;
  int *desc = reinterpret_cast<int *>(empty + FP_SLOTS);   // per slot: (transform << 5) | (tile << 1) | stage
  int *ready = desc + FP_SLOTS;                            // items known ready: [0] stage A, [1] stage B (scout -> producer)
  int *done_cnt = ready + 2;                               // [group] warps that have finished storing a tile (consumers -> releaser)
  int *released = done_cnt + FP_GROUPS;                    // tiles published by the releaser
  int *hist = released + 1;                                // [FP_HIST] descriptor of tile j at j % FP_HIST (producer -> releaser)
  const int tid = threadIdx.x;
  const unsigned full_cnt = 16u * ITEM_WARPS;

  if(tid < 256) fill_tw256_from(tw, p.tw, tid, INV);
  if(tid == 0)
  {
    for(int s = 0; s < FP_SLOTS; s++)
    {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 256);
    }
    ready[0] = 0;
    ready[1] = 0;
    for(int i = 0; i < FP_GROUPS; i++) done_cnt[i] = 0;
    released[0] = 0;
    mbar_fence_init();
  }
  __syncthreads();

  // CTA (set, g): tile g of every transform t = set + n * nsets of its set, once as a stage-A and once as a stage-B item
  const int nsets = (int) gridDim.x >> 4, set = (int) blockIdx.x >> 4, g = (int) blockIdx.x & 15;
  const int my_tiles = (p.batch - set + nsets - 1) / nsets;
  if(tid >= FP_GROUPS * 256)
  {
    // ---------------------------------------------------------------- scout + producer
    // The SCOUT (first lane of the last warp) does all inter-CTA waiting: it polls the
    // per-transform counters with acquire loads and publishes, in shared memory, how many items of each list are ready.
    // The PRODUCER (first lane of the warp before) only reads those two words and issues copies — a thread with bulk
    // copies in flight must not execute gpu-scope acquires / fences itself: each one waited for the outstanding copies
    // (~2 500 cycles), which serialised the ring to one tile in flight.
    if(tid >= FP_GROUPS * 256 + 64)
    {
      // RELEASER (one lane): publishes finished tiles to the other CTAs.  The consumer warps only bump a CTA-scope counter
      // after their stores; the gpu-scope fence that makes those stores visible before the transform's counter moves costs
      // an L2 round trip (~2 500 cycles under load) — paid here, off the math warps, once per batch of finished tiles.
      if(tid == FP_GROUPS * 256 + 64)
      {
        const int total_items = 2 * my_tiles;
        int jr = 0;
        while(jr < total_items)
        {
          int cnt[FP_GROUPS];
#pragma unroll
          for(int i = 0; i < FP_GROUPS; i++) cnt[i] = ld_acquire_cta_s(done_cnt + i);
          int k = 0;
          while(jr + k < total_items && cnt[(jr + k) % FP_GROUPS] >= (int) ITEM_WARPS * ((jr + k) / FP_GROUPS + 1)) k++;
          if(k == 0) { __nanosleep(64); continue; }
          fence_acquire_gpu();
          for(int i = 0; i < k; i++)
          {
            const int d = ld_volatile_s(hist + (jr + i) % FP_HIST), t = set + (d >> 1) * nsets;
            if(!(d & 1)) red_relaxed_add(p.done_a + t, ITEM_WARPS);   // stage-B tiles were announced when their scratch tile had been read
          }
          jr += k;
          st_release_cta(released, jr);
        }
      }
    }
    else if(tid >= FP_GROUPS * 256 + 32)
    {
      // whole warp: lanes 0-15 sample the counters of the next 16 stage-B items, lanes 16-31 those of the next 16 stage-A
      // items (relaxed loads, all in flight together: one L2 round trip — ~2 000 cycles under load — per 32 items at best);
      // the leading run of complete ones is published after an acquire fence
      const int l = tid & 15;
      const bool for_b = (tid & 16) == 0;
      int na = 0, nb = 0;
      while(na < my_tiles || nb < my_tiles)
      {
        const int n = (for_b ? nb : na) + l;
        bool ok = false;
        if(n < my_tiles)
        {
          if(for_b) ok = ld_acquire(p.done_a + set + n * nsets) >= full_cnt;                              // all 16 column tiles written
          else ok = n < p.ring || ld_acquire(p.done_b + set + (n - p.ring) * nsets) >= full_cnt;          // scratch slot free again
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        const int rb = __ffs(~(m & 0xffffu)) - 1, ra = __ffs(~(m >> 16)) - 1;          // leading ready items of each list
        if(ra + rb > 0)
        {
          fence_proxy_async_all();   // the other CTAs' generic-proxy stores before the async-proxy reads that follow the hand-over
          __syncwarp();
          na += ra;
          nb += rb;
          if((tid & 31) == 0)
          {
            st_release_cta(ready + 0, na);
            st_release_cta(ready + 1, nb);
          }
        }
        else __nanosleep(100);
      }
    }
    else if(tid == FP_GROUPS * 256)
    {
      // a stage-B tile goes first whenever one is ready (keeps the scratch window short: it stays in L2), otherwise the next
      // ready stage-A tile; a not-yet-ready stage-B tile never blocks the stage-A prefetch behind it
      int ia = 0, ib = 0, j = 0;
      const uint64_t pol_first = l2_policy_evict_first();
      long long w_empty = 0, w_flag = 0;
      const long long t_start = clock64();
      while(ia < my_tiles || ib < my_tiles)
      {
        const long long c0 = clock64();
        bool take_b;
        for(;;)
        {
          if(ib < my_tiles && ib < ld_volatile_s(ready + 1)) { take_b = true; break; }
          if(ia < my_tiles && ia < ld_volatile_s(ready + 0)) { take_b = false; break; }
          __nanosleep(32);
        }
        const long long c1 = clock64();
        const int slot = j % FP_SLOTS, use = j / FP_SLOTS;
        j++;
        if(use > 0) mbar_wait(empty + slot, (unsigned) (use - 1) & 1u);
        w_empty += clock64() - c1;
        w_flag += c1 - c0;
        float2 *dst = slots + slot * 4096;
        const int n = take_b ? ib : ia, t = set + n * nsets;
        desc[slot] = (n << 1) | (take_b ? 1 : 0);
        while(j - ld_volatile_s(released) > FP_HIST) __nanosleep(64);   // never in practice: the releaser is at most a few tiles behind
        hist[(j - 1) % FP_HIST] = (n << 1) | (take_b ? 1 : 0);
        if(p.trace && blockIdx.x == 0 && j <= 2048)
        {
          int *tr = p.trace + (j - 1) * 6;
          tr[0] = (int) (c0 - t_start); tr[1] = (int) (c1 - t_start); tr[2] = (int) (clock64() - t_start);
          tr[3] = take_b ? 1 : 0; tr[4] = ia | (ib << 16); tr[5] = ld_volatile_s(ready + 0) | (ld_volatile_s(ready + 1) << 16);
        }
        mbar_expect_tx(full + slot, FP_SLOT_BYTES);
        if(!take_b)
        {
          if(p.hints) tma_load_3d_hint(dst, &xmap, 32 * g, 0, t, full + slot, pol_first);
          else tma_load_3d(dst, &xmap, 32 * g, 0, t, full + slot);
          ia++;
        }
        else
        {
          bulk_g2s(dst, p.scratch + ((long long) set * p.ring + n % p.ring) * 65536 + (long long) g * 4096, FP_SLOT_BYTES, full + slot);
          ib++;
        }
      }
      if(p.prof)
      {
        p.prof[blockIdx.x * 8 + 3] = (unsigned long long) w_empty;
        p.prof[blockIdx.x * 8 + 4] = (unsigned long long) w_flag;
        p.prof[blockIdx.x * 8 + 5] = (unsigned long long) (clock64() - t_start);
      }
    }
    return;
  }

  // ------------------------------------------------------------------ consumers
  const int gi = tid >> 8, gt = tid & 255, hi = gt >> 4, lo = gt & 15;
  const float inv256 = 1.0f / 256.0f;
  const uint64_t pol_last = l2_policy_evict_last(), pol_first = l2_policy_evict_first();
  int j;
  long long w_full = 0, n_items = 0;
  const long long t_begin = clock64();
  for(j = gi; j < 2 * my_tiles; j += FP_GROUPS)
  {
    const int slot = j % FP_SLOTS, use = j / FP_SLOTS;
    float2 *sl = slots + slot * 4096;
    float2 v[16];
    const long long c0 = clock64();
    mbar_wait(full + slot, (unsigned) use & 1u);
    w_full += clock64() - c0;
    n_items++;
    const int d = desc[slot], n = d >> 1, t = set + n * nsets;
    // a stage-B tile is in shared memory: its scratch tile may be overwritten (by stage A of transform n + ring of this set)
    // from now on — nothing was written here, so no fence: the counter only has to move after the read has completed
    if((d & 1) && gt == 0) red_relaxed_add(p.done_b + t, ITEM_WARPS);
    if(!(d & 1))
    {
      // ---- stage A: columns n2 in [16g, 16g+16), transform over n1; slot = [n1][16]
      const int n2 = 16 * g + lo;
      const float2 tb = tw4_load<INV>(p.tw4 + hi * 256 + n2), ts = tw4_load<INV>(p.tw4 + 8192 + n2);
#pragma unroll
      for(int q = 0; q < 16; q++) v[q] = sl[(16 * q + hi) * 16 + lo];
      fft16<INV>(v);
      mul_table<INV, true>(v, tw, hi);   // W256^(hi*k1)
      // in place: thread (hi, lo) owns rows {16 q + hi}; value k1 goes where q = k1 was
#pragma unroll
      for(int k1 = 0; k1 < 16; k1++) sl[(16 * k1 + hi) * 16 + lo] = v[k1];
      tc::named_bar(1 + gi, 256);
#pragma unroll
      for(int b = 0; b < 16; b++) v[b] = sl[(16 * hi + b) * 16 + lo];
      tc::mbar_arrive(empty + slot);
      fft16<INV>(v);
      // v[p2] = Y[k1 = hi + 16 p2][n2]; four-step twiddle W_N^(n2*k1)
      mul_geometric(v, tb, ts);
      float2 *sc = p.scratch + ((long long) set * p.ring + n % p.ring) * 65536 + hi * 256 + n2;
#pragma unroll
      for(int p2 = 0; p2 < 16; p2++)
      {
        if(p.hints) stg_hint(sc + p2 * 4096, v[p2], pol_last);
        else sc[p2 * 4096] = v[p2];
      }
      tile_done(done_cnt + gi);
    }
    else
    {
      // ---- stage B: rows k1 in [16g, 16g+16), transform over n2, natural-order output; slot = [16 rows][256]
#pragma unroll
      for(int q = 0; q < 16; q++) v[q] = sl[hi * 256 + 16 * q + lo];
      fft16<INV>(v);
      mul_table<INV, true>(v, tw, lo);   // W256^(b*k1)
      tc::named_bar(1 + gi, 256);        // every thread has consumed the linear layout
#pragma unroll
      for(int k1 = 0; k1 < 16; k1++) sl[k1 * 256 + lo * 16 + ((hi + lo) & 15)] = v[k1];
      tc::named_bar(1 + gi, 256);
#pragma unroll
      for(int b = 0; b < 16; b++) v[b] = sl[hi * 256 + b * 16 + ((lo + b) & 15)];
      tc::mbar_arrive(empty + slot);
      fft16<INV>(v);
      // thread (hi = k', lo = r): v[k2] = X[(16g + r) + 256*(k' + 16*k2)]
      float2 *y = p.y + (long long) t * p.y_stride + hi * 256 + 16 * g + lo;
#pragma unroll
      for(int k2 = 0; k2 < 16; k2++)
      {
        if(p.hints) stg_hint(y + k2 * 4096, make_float2(v[k2].x * inv256, v[k2].y * inv256), pol_first);
        else stg_stream(y + k2 * 4096, make_float2(v[k2].x * inv256, v[k2].y * inv256));
      }
      tile_done(done_cnt + gi);
    }
  }
  if(p.prof && gt == 0)
  {
    if(gi == 0)
    {
      p.prof[blockIdx.x * 8 + 0] = (unsigned long long) (clock64() - t_begin);
      p.prof[blockIdx.x * 8 + 2] = (unsigned long long) n_items;
    }
    atomicAdd(p.prof + blockIdx.x * 8 + 1, (unsigned long long) w_full);
  }
}

// ---------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn()
{
  static EncodeTiledFn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return (EncodeTiledFn) f;
  }();
  return fn;
}

bool fft64k_pipe_usable(const float2 *x, long long xs, const float2 *y, long long ys)
{
  // tensor-map constraints: 16-byte aligned base and row strides
  return encode_fn() && ((uintptr_t) x & 15) == 0 && ((uintptr_t) y & 15) == 0 && (xs & 1) == 0 && (ys & 1) == 0 && xs >= 65536;
}

int fft64k_pipe_run(tsdgpu_fft_s *p, const float2 *x, long long xs, float2 *y, long long ys, int batch, bool forward)
{
  Runtime &r = rt();
  if(!p->pipe_ready)
  {
    TSD_CUDA(cudaFuncSetAttribute(fft64k_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FP_SMEM));
    TSD_CUDA(cudaFuncSetAttribute(fft64k_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FP_SMEM));
    p->pipe_ready = true;
  }
  // input viewed as float32 [batch][256 rows][512]: a stage-A tile is the box {32 floats, 256 rows, 1 transform}
  CUtensorMap map;
  const cuuint64_t dims[3] = {512, 256, (cuuint64_t) batch};
  const cuuint64_t strides[2] = {2048, (cuuint64_t) xs * 8};
  const cuuint32_t box[3] = {32, 256, 1}, estr[3] = {1, 1, 1};
  const CUresult rc = encode_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float2 *>(x), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if(rc != CUDA_SUCCESS) return fail("tsdgpu_fft_exec: cuTensorMapEncodeTiled failed (" + std::to_string((int) rc) + ")");
  TSD_CUDA(cudaMemsetAsync(p->flags, 0, ((size_t) 2 * batch + 1) * sizeof(unsigned), r.stream));
  FftPipeParams q;
  q.y = y;
  q.y_stride = ys;
  q.scratch = p->scratch;
  q.done_a = p->flags;
  q.done_b = p->flags + batch;
  q.batch = batch;
  q.ring = p->pipe_ring;
  static const int hints = getenv("TSDGPU_FFT_HINTS") ? atoi(getenv("TSDGPU_FFT_HINTS")) : 1;
  q.hints = hints;
  q.tw = r.tw256;
  q.tw4 = r.tw4;
  q.prof = nullptr;
  q.trace = nullptr;
  static const bool prof = getenv("TSDGPU_FFT_PROF") != nullptr;
  if(prof)
  {
    TSD_CUDA(cudaMalloc(&q.prof, (size_t) r.num_sms * 8 * sizeof(unsigned long long)));
    TSD_CUDA(cudaMemsetAsync(q.prof, 0, (size_t) r.num_sms * 8 * sizeof(unsigned long long), r.stream));
    TSD_CUDA(cudaMalloc(&q.trace, 2048 * 6 * sizeof(int)));
    TSD_CUDA(cudaMemsetAsync(q.trace, 0, 2048 * 6 * sizeof(int), r.stream));
  }
  const int nsets = std::max(1, std::min(r.num_sms / 16, batch)), grid = nsets * 16;
  {
    KernelTimer timer;
    if(forward) fft64k_pipe_kernel<false><<<grid, FP_THREADS, FP_SMEM, r.stream>>>(map, q);
    else fft64k_pipe_kernel<true><<<grid, FP_THREADS, FP_SMEM, r.stream>>>(map, q);
    TSD_LAUNCH_CHECK();
  }
  if(prof)
  {
    std::vector<unsigned long long> h((size_t) r.num_sms * 8);
    TSD_CUDA(cudaStreamSynchronize(r.stream));
    TSD_CUDA(cudaMemcpy(h.data(), q.prof, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(q.prof);
    {
      std::vector<int> tr(2048 * 6);
      TSD_CUDA(cudaMemcpy(tr.data(), q.trace, tr.size() * sizeof(int), cudaMemcpyDeviceToHost));
      cudaFree(q.trace);
      if(FILE *f = fopen("gpurun_out/fft_trace.txt", "w"))
      {
        for(int i = 0; i < 2048 && (i == 0 || tr[i * 6 + 2]); i++)
          fprintf(f, "%d %d %d %d %d %d %d %d\n", tr[i * 6], tr[i * 6 + 1], tr[i * 6 + 2], tr[i * 6 + 3], tr[i * 6 + 4] & 0xffff, tr[i * 6 + 4] >> 16,
                  tr[i * 6 + 5] & 0xffff, tr[i * 6 + 5] >> 16);
        fclose(f);
      }
    }
    double tot = 0, wf = 0, items = 0, we = 0, wfl = 0, pt = 0;
    for(int c = 0; c < grid; c++)
    {
      tot += (double) h[c * 8 + 0]; wf += (double) h[c * 8 + 1]; items += (double) h[c * 8 + 2];
      we += (double) h[c * 8 + 3]; wfl += (double) h[c * 8 + 4]; pt += (double) h[c * 8 + 5];
    }
    fprintf(stderr, "[fft64k_pipe] per CTA: life %.0f cycles, group-0 items %.1f, consumers waiting for data %.0f cycles per group "
                    "(%.1f %% of life); producer: life %.0f, waiting for a free slot %.0f, for flags %.0f\n",
            tot / grid, items / grid, wf / grid / FP_GROUPS, 100.0 * wf / FP_GROUPS / tot, pt / grid, we / grid, wfl / grid);
  }
  return 0;
}

} // namespace tsdgpu
