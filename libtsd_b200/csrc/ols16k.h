// Internal interface of the single-SM overlap-save kernel (ols16k.cu), used by ola.cu.
#pragma once
#include <cuda_runtime.h>
#include <complex>

namespace tsdgpu {

struct Ols16k
{
  int O = 0, L = 0, K = 0;     // overlap, outputs per internal block (16384 - O), taps
  float *d_init = nullptr;     // [512 threads][128 floats]: per-thread gains and twiddles (-> tensor memory)
};

// Builds the device constants from the reference-layout gains H[N] = DFT([0^(N-K), h]) (fourier.cc:962-965).
// *out stays null (status 0) when this path does not serve the case (K-1 > 8192, or H is not the transform of K taps).
int ols16k_create(const float *H, int N, int K, Ols16k **out);
// Same from the K taps themselves (complex, double): y_fir[u] = sum_m taps[m] stream[u - m].
int ols16k_create_taps(const std::complex<double> *taps, int K, Ols16k **out);
void ols16k_destroy(Ols16k *o);
int ols16k_smem_bytes(int O);
// y[c][i] = sum_m h[m] stream[t0 + i - delay - m] for i in [0, out_count); x[0] of this call is stream sample
// t0 + residual, earlier samples come from carry[c][carry_len + pos] (pos < 0), samples before the stream are zero.
int ols16k_run(Ols16k *o, const float2 *x, long long xs, int n, const float2 *carry, int carry_len, float2 *y, long long ys,
               long long out_count, int delay, int residual, int nchan);

} // namespace tsdgpu
