// Shared host/device helpers of libtsdgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

namespace tsdgpu {

// ---------------------------------------------------------------- host runtime state
// staging of the host-memory entry points (host_pipe.cuh): two device slots each way + their events
struct HostStage
{
  void *in[2] = {nullptr, nullptr}, *out[2] = {nullptr, nullptr};
  size_t in_bytes = 0, out_bytes = 0;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  // pageable caller memory (what a libtsd application's Veccf is): pinned bounce buffers per slot, filled / drained by the
  // copy-thread pool (runtime.cu) instead of the driver's single-threaded staging
  void *pin_in[2] = {nullptr, nullptr}, *pin_out[2] = {nullptr, nullptr};
  size_t pin_in_bytes[2] = {0, 0}, pin_out_bytes[2] = {0, 0};
  cudaEvent_t ev_pin_in[2] = {nullptr, nullptr}, ev_pin_out[2] = {nullptr, nullptr};
  struct Pending { bool active = false; void *dst = nullptr; size_t dpitch = 0, width = 0, height = 0; } pend[2];
};
// One Runtime per CUDA device (runtime.cu keeps a table).  A host thread works on one device at a time
// (tsdgpu_init / tsdgpu_set_device; every object remembers the device it was created on and its entry points switch
// to it); calls that target the same device are serialised by `mu`, calls on different devices run concurrently.
struct Runtime
{
  std::recursive_mutex mu;
  HostStage hs;
  bool ols_ready = false;            // ols16k.cu: constant table + shared-memory opt-in done on this device
  // shared-memory opt-ins (cudaFuncSetAttribute is per device) of the tensor-core kernels and the banded resampler kernel
  bool fir_tc1_ready = false, fir_tc2_ready = false, resamp_tc_ready = false;
  bool ola_sandwich_ready = false;
  size_t resamp2_smem_set = 0;
  int device = -1;
  int num_sms = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;     // stream every launch goes to
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  // auxiliary compute streams: the staged FFT / FFT-filter paths spread their chunks over them so that
  // the stage kernels of different chunks overlap (fork from / join into `stream` with events)
  static constexpr int MAX_AUX = 8;
  cudaStream_t aux[MAX_AUX] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_AUX] = {};
  float4 *tw256 = nullptr;           // [512] local twiddles of the 256-point transforms: forward, then conjugated
  // four-step twiddles of the 65536-point transform, W = exp(-2 pi i / 65536) (host-built in double):
  // [0, 4096) W^(n2*hi) at [hi*256 + n2]; [4096, 8192) the same values at [k1*16 + lo]; [8192, 8448) W^(16*n)
  float2 *tw4 = nullptr;
  long long launches = 0;
  // optional device-side timing of the dominant kernels (tsdgpu_timing_*)
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;
};
// RAII bracket: records an event pair around a main-kernel launch when timing is enabled
struct KernelTimer
{
  cudaEvent_t a = nullptr, b = nullptr;
  KernelTimer();
  ~KernelTimer();
};
Runtime &rt();               // runtime of the calling thread's current device
int ensure_init();           // makes sure the calling thread has a usable device (default: the first one initialised, else 0)
int enter_device(int device);   // switches the calling thread to `device` (< 0: keep / default) and initialises it if needed
// first statement of every C-ABI entry that touches a device: select it, then hold its lock for the whole call
#define TSD_ENTER(dev)                                                                  \
  if(::tsdgpu::enter_device(dev)) return 1;                                             \
  std::lock_guard<std::recursive_mutex> tsd_guard__(::tsdgpu::rt().mu)
int aux_init();          // creates the auxiliary streams, their events and the twiddle table (idempotent)
int aux_fork(int n);     // aux[0..n) wait for everything enqueued so far on rt().stream
int aux_join(int n);     // rt().stream waits for everything enqueued on aux[0..n)
void set_error(const std::string &s);
int fail(const std::string &s);

#define TSD_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if(e__ != cudaSuccess)                                                              \
      return ::tsdgpu::fail(std::string(#expr) + ": " + cudaGetErrorString(e__));       \
  } while(0)

#define TSD_LAUNCH_CHECK()                                                              \
  do {                                                                                  \
    ::tsdgpu::rt().launches++;                                                          \
    cudaError_t e__ = cudaGetLastError();                                               \
    if(e__ != cudaSuccess)                                                              \
      return ::tsdgpu::fail(std::string("kernel launch: ") + cudaGetErrorString(e__));  \
  } while(0)

// ---------------------------------------------------------------- device helpers
// ---- packed FP32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2): one issue slot per complex add.
// A float2 lives in an aligned 64-bit register pair; scalar broadcast (make_float2(s, s)) and half
// swap (make_float2(v.y, v.x)) become operand modifiers (.F32 / .F32x2.LO_HI) in SASS.
__device__ __forceinline__ unsigned long long f2_bits(float2 v) { return *reinterpret_cast<unsigned long long *>(&v); }
__device__ __forceinline__ float2 bits_f2(unsigned long long u) { return *reinterpret_cast<float2 *>(&u); }
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(d);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b)
{
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(d);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
  return bits_f2(d);
}
__device__ __forceinline__ float2 bcast2(float s) { return make_float2(s, s); }
// i * w
__device__ __forceinline__ float2 rot90(float2 w) { return make_float2(-w.y, w.x); }
// a * w given w and rw = i*w : (a.x*w.x - a.y*w.y, a.x*w.y + a.y*w.x) in two packed instructions
__device__ __forceinline__ float2 cmul_rot(float2 a, float2 w, float2 rw) { return fma2(bcast2(a.y), rw, mul2(bcast2(a.x), w)); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return cmul_rot(a, b, rot90(b)); }
// a * conj(b)
__device__ __forceinline__ float2 cmulc(float2 a, float2 b)
{
  return cmul_rot(a, make_float2(b.x, -b.y), make_float2(b.y, b.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return add2(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return sub2(a, b); }

// exp(-+ 2*pi*i * num/den), den a power of two <= 2^24, 0 <= num: exact angle in revolutions.
template<bool INV> __device__ __forceinline__ float2 twiddle(unsigned num, float two_over_den)
{
  float s, c;
  sincospif((float) num * two_over_den, &s, &c);
  return make_float2(c, INV ? s : -s);
}

// ---- mbarrier + 1-D bulk async copy (TMA unit, SASS UBLKCP) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TSD_MBAR_SLEEP adds a suspend-time hint (the hardware parks the warp instead of re-issuing try_wait every few cycles).
// Measured: +3 % for ols16k, whose waits are long (its own mbar_wait_sleep), -2..4 % for the tensor-core kernels, whose
// waits are short and frequent (the wake-up costs more than the spinning): off by default here.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
#ifdef TSD_MBAR_SLEEP
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "WAIT_%=:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
    "@p bra DONE_%=;\n\t"
    "bra WAIT_%=;\n\t"
    "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
#else
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "WAIT_%=:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
    "@p bra DONE_%=;\n\t"
    "bra WAIT_%=;\n\t"
    "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
}
// Wait of a role that is idle for a long time by design (an epilogue warp waiting a whole tile time for its accumulator,
// a loader waiting for a free staging slot): try_wait with a suspend-time hint, so that the warp is parked instead of
// re-issuing SYNCS / BRA every few cycles next to the warps that have work.  TSD_LONGWAIT=0 builds spin like mbar_wait.
#ifndef TSD_LONGWAIT
#define TSD_LONGWAIT 1
#endif
__device__ __forceinline__ void mbar_wait_long(uint64_t *bar, unsigned parity)
{
#if TSD_LONGWAIT
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "WAIT_%=:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
    "@p bra DONE_%=;\n\t"
    "bra WAIT_%=;\n\t"
    "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
#else
  mbar_wait(bar, parity);
#endif
}
// global -> shared, completion counted on the mbarrier (bytes multiple of 16, both 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                 smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, unsigned bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
// hint: bring [src, src + bytes) into L2 (both 16-byte aligned); no completion to wait for
__device__ __forceinline__ void bulk_prefetch_l2(const void *src_gmem, unsigned bytes)
{
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- Ampere-style async copy (LDGSTS), 8 bytes per thread
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- inter-CTA flags (release/acquire at gpu scope) -------------------------------------------
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned *p, unsigned v)
{
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// streaming global accesses (touched once: do not keep in L1)
__device__ __forceinline__ float2 ldg_stream(const float2 *p)
{
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(float2 *p, float2 v)
{
  asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

} // namespace tsdgpu
