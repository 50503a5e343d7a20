// Tensor-core (tcgen05, 3xTF32) LUT resampler (resamp_tc.cu).
#pragma once
#include <cuda_runtime.h>

namespace tsdgpu {

struct ResampTcParams
{
  const float2 *x;        // [nchan][x_stride] this call's input
  float2 *y;              // [nchan][y_stride]
  const float2 *hist;     // [nchan][hist_len] inputs preceding x[0]
  const float *lut;       // [(nphases+1)][K]
  const int2 *sched;      // per output of this chunk: {in_idx (call-relative), lut_idx}
  long long x_stride, y_stride;
  long long out0;         // first output (call-relative) of this chunk
  long long n_out;        // outputs of this chunk
  int n;                  // input samples of the call (positions >= n read as zero)
  int K, hist_len, nchan;
  int lut_elems;          // K * (nphases + 1)
  int max_tile_chunks;    // from resamp_tc_eligible: most 32-input chunks feeding one tile of 128 outputs
  int ntiles, span, vec_store, band, groups;   // filled by resamp_tc_launch
};

bool resamp_tc_eligible(const int2 *sched_host, long long n_out, int K, int nphases, const void *x, long long x_stride, int *max_tile_chunks);
int resamp_tc_launch(const ResampTcParams &p);

}
