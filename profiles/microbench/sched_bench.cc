// Host-side phase recurrence of the resampler (ra.cc:64-73): float32 loop as in resamp.cu vs an integer-exact run formulation
// (prototype, not merged).  g++ -O3 sched_bench.cc && ./a.out
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <chrono>
#include <vector>
#include <algorithm>
struct int2 { int x, y; };
static float sched_f(float phase, float increment, int nphases, int i0, int i1, int2 *out, size_t *count)
{
  size_t j = 0; const float nph = (float) nphases;
  for(int i = i0; i < i1; i++) { if(phase < 1) { out[j].x = i; out[j].y = (int)(phase * nph); j++; phase += increment; } phase--; }
  *count = j; return phase;
}
static bool sched_runs(float *phase_io, float increment, int nphases, int i0, int i1, int2 *out, size_t *count)
{
  const float ph = *phase_io;
  if(!(increment > 1.0f && increment < 2.0f) || !(ph >= 0.0f && ph < 2.0f)) return false;
  const float t = ph * 8388608.0f;
  if(t != std::floor(t)) return false;
  const uint32_t ONE = 1u << 23;
  uint32_t P = (uint32_t) t;
  const uint32_t I = (uint32_t) (increment * 8388608.0f), d = I - ONE;
  const uint32_t mc = (ONE + d - 1) / d;            // ceil(ONE / d) >= 2
  const uint32_t m_lo = mc >= 2 ? mc - 2 : 0;
  const float sc = 0x1p-23f, nph = (float) nphases;
  size_t j = 0;
  int i = i0;
  if(i < i1 && P >= ONE) { P -= ONE; i++; }
  while(i < i1)
  {
    uint32_t m = std::max(1u, m_lo);
    while(P + m * d < ONE) m++;
    const uint32_t k_emit = std::min<uint32_t>(m, (uint32_t) (i1 - i));
    int2 *o = out + j;
    for(uint32_t k = 0; k < k_emit; k++)
    {
      const uint32_t Pk = P + k * d;
      o[k].x = i + (int) k;
      o[k].y = (int) ((float) Pk * sc * nph);
    }
    j += k_emit;
    if(k_emit < m) { P += k_emit * d; i += (int) k_emit; break; }
    uint32_t S = P + m * d + ONE;
    S = (S + ((S >> 1) & 1u)) & ~1u;
    P = S - ONE;
    i += (int) m;
    if(i < i1) { P -= ONE; i++; }
  }
  *phase_io = (float) P * sc;
  *count = j;
  return true;
}
int main()
{
  const int n = 1 << 23; std::vector<int2> a(n + 64), b(n + 64);
  const double ratios[] = {147.0/160, 0.999999, 0.50001, 0.75, 2.0/3, 0.6180339, 0.97, 0.51, 44100.0/48000, 0.9999, 0.5000001, 0.53};
  for(double r : ratios)
  {
    const float inc = 1.0f / (float) r;
    float pa = 0.f, pb = 0.f; bool same = true; double ta = 0, tb = 0; size_t tot = 0;
    int pos = 0; const int chunks[] = {1, 10, 1000, 65536, 777, 300001, 1 << 22, 12345, 3};
    for(int c : chunks)
    {
      size_t ca = 0, cb = 0;
      auto t0 = std::chrono::steady_clock::now();
      pa = sched_f(pa, inc, 256, pos, pos + c, a.data(), &ca);
      auto t1 = std::chrono::steady_clock::now();
      bool ok = sched_runs(&pb, inc, 256, pos, pos + c, b.data(), &cb);
      auto t2 = std::chrono::steady_clock::now();
      ta += std::chrono::duration<double, std::milli>(t1 - t0).count(); tb += std::chrono::duration<double, std::milli>(t2 - t1).count();
      same = same && ok && ca == cb && pa == pb;
      for(size_t k = 0; k < ca && same; k++) same = a[k].x == b[k].x && a[k].y == b[k].y;
      pos += c; tot += ca;
    }
    printf("ratio %.7f: same %d, float %.2f ms, runs %.2f ms (%zu outputs)\n", r, (int) same, ta, tb, tot);
  }
}
