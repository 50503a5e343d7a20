// Tensor-core (tcgen05, 3xTF32) direct FIR for cf32 data and K <= 127 real taps (fir_tc.cu).
#pragma once
#include <cuda_runtime.h>

namespace tsdgpu {

struct FirTcParams
{
  const float2 *x;         // [nchan][x_stride]
  float2 *y;               // [nchan][y_stride]
  const float2 *hist;      // [nchan][halo] samples before x[0] (oldest first)
  const float *taps_rev;   // taps in accumulation order: taps_rev[m] = h[K-1-m]
  long long x_stride, y_stride;
  int n, K, halo, nchan;
  int ntiles, span;        // filled by fir_tc_launch
  int real = 0;            // 1: real-valued data (x, y, hist point to float rows; strides / halo count floats)
};

// 16-byte aligned rows (the producers use 128-bit loads) and K <= 127
bool fir_tc_eligible(int kind_cf32_f32, int K, const void *x, long long x_stride, const void *y, long long y_stride, const void *hist, int halo);
bool fir_tc_real_eligible(int K, const void *x, long long x_stride, const void *y, long long y_stride, const void *hist, int halo);
int fir_tc_launch(const FirTcParams &p);

}
