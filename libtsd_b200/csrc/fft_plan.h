// Internal view of an FFT plan (shared by fft.cu and ola.cu).
#pragma once
#include <cuda_runtime.h>

struct tsdgpu_fft_s
{
  int device = 0;              // CUDA device the plan lives on
  int n = 0, batch = 0;
  // N = 65536 pipeline
  float2 *scratch = nullptr;   // ring of L2-resident intermediates
  unsigned *flags = nullptr;   // done_a[batch], done_b[batch], ticket
  int ring = 0, lag = 0, ctas = 0;
  int staged = 1, chunk = 32, nstreams = 4;   // staged form: transforms per stage kernel, auxiliary streams
  // TMA-fed persistent form (fft64k_pipe.cu): default (112 vs 103 G round trips/s and less DRAM write-back than the staged
  // kernels); TSDGPU_FFT_MODE=staged|persistent selects the older schedules, unaligned buffers fall back to staged
  int pipe = 1, pipe_ring = 8;   // scratch slots per set of 16 CTAs
  bool pipe_ready = false;
  bool smem_optin = false;     // shared-memory kernel: > 48 KiB of dynamic shared memory enabled
  // generic radix-2 path
  float2 *work[2] = {nullptr, nullptr};
  size_t work_cap[2] = {0, 0};   // elements allocated
  int batch_created = 0;          // batch the plan was created with
  // n not a power of two (TFRPlanDefaut::configure, fourier.cc:372-405): even n -> two transforms of n/2 + one
  // radix-2 combine (fourier.cc:438-462); odd n -> chirp-z through a power-of-two plan of n2 = p2(2n-1) (fourier.cc:237-255)
  tsdgpu_fft_s *sub = nullptr;
  int n2 = 0;
  float2 *d_rot = nullptr;     // even n: tfr_rotation(n), [n]
  float2 *d_chirp = nullptr;   // odd n: chirp.tail(n), [n]
  float2 *d_Xc = nullptr;      // odd n: unitary transform of conj(chirp) zero-padded to n2, [n2]
  float2 *nwork = nullptr;     // [2*batch][n/2] or [batch][n2]
};

namespace tsdgpu {
int fft_plan_create(int n, int batch, tsdgpu_fft_s **out);
void fft_plan_destroy(tsdgpu_fft_s *p);
// device pointers, enqueued on the library stream
int fft_exec_device(tsdgpu_fft_s *p, const float2 *x, long long xs, float2 *y, long long ys, bool forward);
// fft64k_pipe.cu
bool fft64k_pipe_usable(const float2 *x, long long xs, const float2 *y, long long ys);
int fft64k_pipe_run(tsdgpu_fft_s *p, const float2 *x, long long xs, float2 *y, long long ys, int batch, bool forward);
}
