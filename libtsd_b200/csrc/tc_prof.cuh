// Per-role cycle accounting for the warp-specialised tensor-core kernels (compiled in with -DTSD_TC_PROF only).
// The including file defines PROF_ARRAY = a __device__ long long [1024][24][4] array ([cta][warp][wait A, wait B, work, total]).
#ifdef TSD_TC_PROF
// timing experiment: per-CTA cycle totals [cta][role 0..15][what 0..3]; role = warp, what: 0 wait A, 1 wait B, 2 work, 3 total
#define PROF_DECL long long pf_[4] = {0, 0, 0, 0}; const long long pf_t0 = clock64();
#define PROF_BEGIN(v) const long long v = clock64();
#define PROF_ADD(k, v) pf_[k] += clock64() - (v);
#define PROF_END                                                                                             \
  if(lane == 0 && blockIdx.y == 0 && blockIdx.x < 1024)                                                      \
  {                                                                                                          \
    pf_[3] = clock64() - pf_t0;                                                                              \
    for(int k = 0; k < 4; k++) PROF_ARRAY[blockIdx.x][warp][k] = pf_[k];                                       \
  }
#else
#define PROF_DECL
#define PROF_BEGIN(v)
#define PROF_ADD(k, v)
#define PROF_END
#endif

