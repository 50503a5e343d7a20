// Internal view of an FFT plan (shared by fft.cu and ola.cu).
#pragma once
#include <cuda_runtime.h>

struct tsdgpu_fft_s
{
  int n = 0, batch = 0;
  // N = 65536 pipeline
  float2 *scratch = nullptr;   // ring of L2-resident intermediates
  unsigned *flags = nullptr;   // done_a[batch], done_b[batch], ticket
  int ring = 0, lag = 0, ctas = 0;
  int staged = 1, chunk = 32, nstreams = 4;   // staged form: transforms per stage kernel, auxiliary streams
  // generic radix-2 path
  float2 *work[2] = {nullptr, nullptr};
};

namespace tsdgpu {
int fft_plan_create(int n, int batch, tsdgpu_fft_s **out);
void fft_plan_destroy(tsdgpu_fft_s *p);
// device pointers, enqueued on the library stream
int fft_exec_device(tsdgpu_fft_s *p, const float2 *x, long long xs, float2 *y, long long ys, bool forward);
}
