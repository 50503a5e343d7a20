/* TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.
 *
 * Plain-C restatement of the libtsd CPU algorithms on the filtering hot path.
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/core).  Build: gcc -O2 -march=x86-64 -ffp-contract=off (no FMA
 * contraction: same arithmetic as the reference's release flags
 * -O3 -march=x86-64, std-makefile-defs:171).
 *
 * PARITY PINNED: tests/test_oracle_vs_ref.py checks this file against the
 * reference's own sources compiled in place (oracle/_ref/libtsdref.so) — bit-exact
 * for the streaming paths (FIR, radix-2 FFT, OLA, resampler step, re-blocking, p2),
 * <=1e-6 for the design helpers (taps, LUT) — and tests/golden/ holds vectors produced
 * by that reference build.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef float _Complex cf32;
typedef double _Complex cf64;

#define TSD_PI 3.14159265358979323846
static const float TSD_PI_F = 3.14159265358979323846f;

/* ------------------------------------------------------------------ p2 / cost model */

/* tsd.cc:287-291  prochaine_puissance_de_2 (float log, kept as is) */
int tsdo_p2(int i)
{
  int lg2 = (int) ceilf(logf((float) i) / logf(2.0f));
  return (int) (1l << lg2);
}

/* fourier.cc:705-713  ola_complexité */
void tsdo_ola_complexite(int M, int Ne, float *C, int *Nf, int *Nz)
{
  *Nf = tsdo_p2(Ne + M - 1);
  *Nz = *Nf - Ne;
  *C = (1.0f / Ne) * 2 * 5 * *Nf * logf(1.0f * *Nf) / logf(2.0f);
}

/* fourier.cc:715-735  ola_complexité_optimise */
void tsdo_ola_complexite_optimise(int M, float *C_, int *Nf_, int *Nz_, int *Ne_)
{
  int kmin = (int) ceil(log((double) M) / log(2.0));
  for(int k = kmin; (k < kmin + 20) && (k < 31); k++)
  {
    int Nf, Nz, Ne = (1 << k) - (M - 1);
    float C;
    tsdo_ola_complexite(M, Ne, &C, &Nf, &Nz);
    if((k == kmin) || (C < *C_))
    {
      *Nf_ = Nf;
      *Nz_ = Nf - Ne;
      *Ne_ = Ne;
      *C_ = C;
    }
  }
}

/* ------------------------------------------------------------------ design helpers */

/* divers.cc:6-12 */
static float tsdo_sinc(float T, float f)
{
  float a = TSD_PI_F * T * f;
  if(fabsf(a) < 1e-7f)
    return T;
  return sinf(a) / (TSD_PI_F * f);
}

/* tsd.hpp:916-931 */
static void tsdo_linspace(float a, float b, int n, float *x)
{
  if(n > 0)
    x[0] = a;
  if(n > 1)
  {
    double step = ((double) b - a) / (n - 1);
    for(int i = 1; i < n; i++)
      x[i] = (float) (a + step * i);
  }
}

/* fenetres.cc:17-59,125-128,222  symmetric or periodic Hann / rectangular window.
 * win: "hn" or "re".  Returns 0 on success. */
int tsdo_fenetre(const char *win, int n, int sym, float *w)
{
  if(!strcmp(win, "re") || !strcmp(win, ""))
  {
    for(int i = 0; i < n; i++) w[i] = 1.0f;
    return 0;
  }
  if(strcmp(win, "hn"))
    return 1;
  float tmin = -(n / 2), tmax;
  if((n & 1) == 0) tmax = sym ? n / 2 : (n - 1) / 2;
  else tmax = sym ? n / 2 : n / 2 - ((float) n - 1) / n;
  tsdo_linspace(tmin / n, tmax / n, n, w);
  float two_pi = (float) (2 * TSD_PI);
  for(int i = 0; i < n; i++)
    w[i] = 0.5f + (1 - 0.5f) * cosf(two_pi * w[i]);
  return 0;
}

/* rif-fen.cc:30-41,57-61,85-107  design_rif_fen(n, "lp", fc, win) */
int tsdo_design_rif_fen_lp(int n, float fc, const char *win, float *h)
{
  float *w = (float *) malloc(sizeof(float) * (n > 0 ? n : 1));
  if(tsdo_fenetre(win, n, 1, w)) { free(w); return 1; }
  int c = (n & 1) ? n / 2 : (n - 1) / 2;
  double acc = 0;
  for(int i = 0; i < n; i++)
  {
    h[i] = tsdo_sinc(2 * fc, (float) (i - c)) * w[i];
    acc += h[i];
  }
  float s = (float) acc;
  for(int i = 0; i < n; i++)
    h[i] = h[i] / s;
  free(w);
  return 0;
}

/* itrp.cc:24-54  windowed-sinc interpolator LUT: lut[p*K + i], p = 0..nphases.
 * win: "hn" or anything else (no window). */
void tsdo_itrp_sinc_lut(int K, int nphases, float fcut, const char *win, float *lut)
{
  float *lin = (float *) malloc(sizeof(float) * (K > 0 ? K : 1));
  tsdo_linspace((float) (-K / 2), (float) ((K - 1) / 2), K, lin);
  float scale = (float) (2 * TSD_PI / K);
  int hann = !strcmp(win, "hn");
  for(int p = 0; p <= nphases; p++)
  {
    float tau = (float) ((1.0 * p) / nphases);
    float *col = lut + (size_t) p * K;
    for(int i = 0; i < K; i++)
    {
      float h = tsdo_sinc(2 * fcut, (float) (i - K / 2) - tau);
      if(hann)
      {
        float t = (lin[i] - tau) * scale;
        float r2 = 0.5f + (2 * 0.25f) * cosf(t);
        h = h * r2;
      }
      col[i] = h;
    }
  }
  free(lin);
}

/* ------------------------------------------------------------------ direct FIR */

/* filtre-rt.cc:53-109  FiltreRIF<T,Tc>.  kind 0: float data/float taps, 1: cfloat data/float taps,
 * 2: cfloat data/cfloat taps (the three instantiations of filtre-rt.cc:816-818). */
typedef struct
{
  int kind, K, index;
  float *coefs; /* K floats (kind 0,1) or K cfloat (kind 2) */
  float *fen;   /* K floats (kind 0) or K cfloat */
} tsdo_fir;

tsdo_fir *tsdo_fir_new(int kind, const float *taps, int K)
{
  if(K <= 0 || kind < 0 || kind > 2) return NULL;
  tsdo_fir *f = (tsdo_fir *) calloc(1, sizeof(*f));
  f->kind = kind;
  f->K = K;
  size_t ct = (kind == 2) ? 2 : 1, dt = (kind == 0) ? 1 : 2;
  f->coefs = (float *) malloc(sizeof(float) * ct * K);
  memcpy(f->coefs, taps, sizeof(float) * ct * K);
  f->fen = (float *) calloc(dt * K, sizeof(float));
  return f;
}
void tsdo_fir_free(tsdo_fir *f)
{
  if(!f) return;
  free(f->coefs);
  free(f->fen);
  free(f);
}
int tsdo_fir_index(const tsdo_fir *f) { return f->index; }

/* one sample in, one out; accumulation starts at the OLDEST sample with h[K-1] (filtre-rt.cc:82-107) */
void tsdo_fir_step(tsdo_fir *f, const float *x, int n, float *y)
{
  const int K = f->K;
  if(f->kind == 0)
  {
    for(int j = 0; j < n; j++)
    {
      f->fen[f->index] = x[j];
      f->index = (f->index + 1) % K;
      float s = 0;
      int c = K - 1;
      for(int i = f->index; i < K; i++) s += f->fen[i] * f->coefs[c--];
      for(int i = 0; i < f->index; i++) s += f->fen[i] * f->coefs[c--];
      y[j] = s;
    }
  }
  else if(f->kind == 1)
  {
    for(int j = 0; j < n; j++)
    {
      f->fen[2 * f->index] = x[2 * j];
      f->fen[2 * f->index + 1] = x[2 * j + 1];
      f->index = (f->index + 1) % K;
      float sr = 0, si = 0;
      int c = K - 1;
      for(int i = f->index; i < K; i++, c--)
      {
        sr += f->fen[2 * i] * f->coefs[c];
        si += f->fen[2 * i + 1] * f->coefs[c];
      }
      for(int i = 0; i < f->index; i++, c--)
      {
        sr += f->fen[2 * i] * f->coefs[c];
        si += f->fen[2 * i + 1] * f->coefs[c];
      }
      y[2 * j] = sr;
      y[2 * j + 1] = si;
    }
  }
  else
  {
    cf32 *fen = (cf32 *) f->fen;
    const cf32 *co = (const cf32 *) f->coefs;
    const cf32 *xi = (const cf32 *) x;
    cf32 *yo = (cf32 *) y;
    for(int j = 0; j < n; j++)
    {
      fen[f->index] = xi[j];
      f->index = (f->index + 1) % K;
      cf32 s = 0;
      int c = K - 1;
      for(int i = f->index; i < K; i++) s += fen[i] * co[c--];
      for(int i = 0; i < f->index; i++) s += fen[i] * co[c--];
      yo[j] = s;
    }
  }
}

/* ------------------------------------------------------------------ FFT plan */

/* fourier.cc:32-46  twiddles by double-precision recurrence, stored as cfloat */
static void tsdo_rotations(int n, cf32 *w)
{
  double th = (-1 * 2 * TSD_PI) / n;
  cf64 r = 1, w0 = cos(th) + sin(th) * I;
  for(int i = 0; i < n; i++)
  {
    w[i] = (float) creal(r) + (float) cimag(r) * I;
    r *= w0;
  }
}

typedef struct
{
  int n;
  cf32 *rot, *a, *b;
} tsdo_fft;

/* TFRPlanDefaut::configure for n = 2^k (fourier.cc:372-405).  The normalise flag is
 * ignored by the reference (always unitary, fourier.cc:119-120,362). */
tsdo_fft *tsdo_fft_new(int n)
{
  if(n <= 0 || (n & (n - 1))) return NULL;
  tsdo_fft *p = (tsdo_fft *) calloc(1, sizeof(*p));
  p->n = n;
  p->rot = (cf32 *) malloc(sizeof(cf32) * n);
  p->a = (cf32 *) malloc(sizeof(cf32) * n);
  p->b = (cf32 *) malloc(sizeof(cf32) * n);
  tsdo_rotations(n, p->rot);
  return p;
}
void tsdo_fft_free(tsdo_fft *p)
{
  if(!p) return;
  free(p->rot);
  free(p->a);
  free(p->b);
  free(p);
}

/* fourier.cc:61-121  tfr_radix2: autosort radix-2, log2(N) ping-pong passes, then X /= sqrt((float) N)
 * (complex / complex, tableau.hpp:1228-1232 -> tableau.cc:1323-1332). */
void tsdo_fft_step(tsdo_fft *p, const float *xin, int forward, float *yout)
{
  const int N = p->n;
  const cf32 *x = (const cf32 *) xin;
  cf32 *y = (cf32 *) yout;
  if(N == 1)
  {
    y[0] = x[0];
    return;
  }
  const cf32 *src = x;
  cf32 *dst = p->a;
  for(int n = 1; n < N; n *= 2)
  {
    int pas = N / (2 * n);
    const cf32 *E = src;
    cf32 *lo = dst, *hi = dst + N / 2;
    for(int k = 0; k < n; k++)
    {
      cf32 rot = p->rot[k * pas];
      float tr = crealf(rot), ti = forward ? cimagf(rot) : -cimagf(rot);
      for(int m = 0; m < pas; m++)
      {
        cf32 g = E[pas], e = E[0];
        float gx = crealf(g), gy = cimagf(g);
        float pr = tr * gx - ti * gy, pi = tr * gy + ti * gx;
        *lo++ = (crealf(e) + pr) + (cimagf(e) + pi) * I;
        *hi++ = (crealf(e) - pr) + (cimagf(e) - pi) * I;
        E++;
      }
      E += pas;
    }
    src = dst;
    dst = (dst == p->a) ? p->b : p->a;
  }
  cf32 s = sqrtf((float) N);
  for(int i = 0; i < N; i++)
    y[i] = src[i] / s;
}

/* ------------------------------------------------------------------ re-blocking */

/* tsd.cc:307-371  TamponNv2: number of full N-blocks fired by a chunk of n samples and the
 * new residual; the sample routing itself is in tsdo_ola_step. */
int tsdo_tampon_blocks(int N, int *windex, int n)
{
  int total = *windex + n;
  *windex = total % N;
  return total / N;
}

/* ------------------------------------------------------------------ OLA (filtre_fft) */

/* fourier.cc:737-932  OLA<cfloat>, plain mode (avec_fenetrage = non) and Hann-window 50 % overlap mode
 * (avec_fenetrage = oui), callback X *= H (fourier.cc:956-959) or identity when H == NULL. */
typedef struct
{
  int Ne, N, Nz, windex, avec_fenetrage;
  int64_t cnt_ech;
  tsdo_fft *plan;
  cf32 *padded, *X, *x2, *svg, *svg_own, *tampon, *H, *last;
  float *fen;
} tsdo_ola;

tsdo_ola *tsdo_ola_new(int dim_blocs_temporel, int nb_zeros_min, const float *H);

/* fourier.cc:794-798: fenêtre("hn", Ne, non) */
tsdo_ola *tsdo_ola_new2(int dim_blocs_temporel, int nb_zeros_min, const float *H, int avec_fenetrage)
{
  tsdo_ola *o = tsdo_ola_new(dim_blocs_temporel, nb_zeros_min, H);
  if(!o || !avec_fenetrage) return o;
  o->avec_fenetrage = 1;
  o->last = (cf32 *) calloc(o->Ne, sizeof(cf32));                /* :784 */
  o->fen = (float *) malloc(sizeof(float) * o->Ne);
  tsdo_fenetre("hn", o->Ne, 0, o->fen);
  return o;
}

tsdo_ola *tsdo_ola_new(int dim_blocs_temporel, int nb_zeros_min, const float *H)
{
  tsdo_ola *o = (tsdo_ola *) calloc(1, sizeof(*o));
  o->Ne = dim_blocs_temporel;
  if(o->Ne <= 0) o->Ne = 512;                                   /* fourier.cc:769-770 */
  o->N = tsdo_p2(o->Ne + nb_zeros_min);                         /* :775 */
  o->Nz = o->N - o->Ne;
  o->cnt_ech = -(o->Ne / 2);                                    /* :779 */
  o->plan = tsdo_fft_new(o->N);
  if(!o->plan) { free(o); return NULL; }
  o->padded = (cf32 *) calloc(o->N, sizeof(cf32));
  o->X = (cf32 *) calloc(o->N, sizeof(cf32));
  o->x2 = (cf32 *) calloc(o->N, sizeof(cf32));
  o->svg = o->svg_own = (cf32 *) calloc(o->Ne, sizeof(cf32));
  o->tampon = (cf32 *) calloc(o->Ne, sizeof(cf32));
  if(H)
  {
    o->H = (cf32 *) malloc(sizeof(cf32) * o->N);
    memcpy(o->H, H, sizeof(cf32) * o->N);
  }
  return o;
}
void tsdo_ola_free(tsdo_ola *o)
{
  if(!o) return;
  tsdo_fft_free(o->plan);
  free(o->padded); free(o->X); free(o->x2); free(o->svg_own); free(o->tampon); free(o->H); free(o->last); free(o->fen);
  free(o);
}
void tsdo_ola_dims(const tsdo_ola *o, int *Ne, int *N, int *Nz, int *residual)
{
  *Ne = o->Ne; *N = o->N; *Nz = o->Nz; *residual = o->windex;
}

/* fourier.cc:837-882 step_interne (plain branch).  Returns 1 where the reference runs off its
 * buffers (N_zeros > Ne: svg.tail(N_zeros) starts before svg, tableau.cc:520 lets it through). */
static int tsdo_ola_block(tsdo_ola *o, const cf32 *x, cf32 *y)
{
  const int Ne = o->Ne, N = o->N, Nz = o->Nz;
  if(Nz > Ne) return 1;
  memcpy(o->padded + Nz, x, sizeof(cf32) * Ne);
  tsdo_fft_step(o->plan, (const float *) o->padded, 1, (float *) o->X);
  if(o->H)
    for(int i = 0; i < N; i++) o->X[i] *= o->H[i];
  tsdo_fft_step(o->plan, (const float *) o->X, 0, (float *) o->x2);
  for(int i = 0; i < Nz; i++) o->svg[Ne - Nz + i] += o->x2[i];
  memcpy(y, o->svg, sizeof(cf32) * Ne);
  memcpy(o->svg, o->x2 + (N - Ne), sizeof(cf32) * Ne);
  o->cnt_ech += Ne;
  return 0;
}

static void tsdo_ola_spectral(tsdo_ola *o)
{
  tsdo_fft_step(o->plan, (const float *) o->padded, 1, (float *) o->X);
  if(o->H)
    for(int i = 0; i < o->N; i++) o->X[i] *= o->H[i];
  tsdo_fft_step(o->plan, (const float *) o->X, 0, (float *) o->x2);
}

/* fourier.cc:884-930 step_interne, windowed branch: two transforms per block (previous half + new half, then the
 * new block alone), both multiplied by the window, halves recombined through `last`.  *ny = Ne, or 0 for the
 * blocks seen while cnt_ech < 0 (the first one).
 * Aliasing of the reference, kept as is: `svg = x2.segment(N_zeros, Ne)` (:923) assigns a temporary VIEW to an
 * owning vector, which moves the view in (tableau.hpp:545-566): from the end of the first block on, svg IS
 * x2[N_zeros..N).  x2 is rewritten in place by every inverse transform (resize to the same size keeps the buffer,
 * tableau.cc:702-705) and the later `svg = ...` assignments copy that memory onto itself (est_reference() and same
 * dimensions, tableau.hpp:579-590).  So after the first block the "saved" frame is the CURRENT frame: each frame is
 * folded onto itself (tail(N_zeros) += head(N_zeros)) instead of onto its predecessor. */
static int tsdo_ola_block_fen(tsdo_ola *o, const cf32 *x, cf32 *y, int *ny)
{
  const int Ne = o->Ne, N = o->N, Nz = o->Nz, h = Ne / 2;
  if(Nz > Ne) return 1;
  cf32 *tail = o->padded + (N - Ne);
  /* 1) previous + new */
  memcpy(o->padded + (N - h), x, sizeof(cf32) * h);                          /* :887 */
  for(int i = 0; i < Ne; i++) tail[i] *= (cf32) o->fen[i];                    /* :888 (window as complex) */
  tsdo_ola_spectral(o);
  for(int i = 0; i < Nz; i++) o->svg[Ne - Nz + i] += o->x2[i];               /* :896 */
  for(int i = 0; i < h; i++) o->last[Ne - h + i] += o->svg[i] / 2.0f;        /* :899 */
  if(o->cnt_ech >= 0)
  {
    memcpy(y, o->last, sizeof(cf32) * Ne);
    *ny = Ne;
  }
  else
    *ny = 0;
  for(int i = 0; i < h; i++) o->last[i] = o->svg[Ne - h + i] / 2.0f;         /* :905 */
  for(int i = 0; i < h; i++) o->last[Ne - h + i] = 0;                        /* :906 */
  if(o->svg != o->x2 + Nz) memcpy(o->svg, o->x2 + (N - Ne), sizeof(cf32) * Ne);   /* :909 (self-copy once aliased) */
  o->cnt_ech += Ne / 2;
  /* 2) new block alone */
  for(int i = 0; i < Ne; i++) tail[i] = x[i] * (cf32) o->fen[i];              /* :914 */
  tsdo_ola_spectral(o);
  for(int i = 0; i < Nz; i++) o->svg[Ne - Nz + i] += o->x2[i];               /* :921 */
  for(int i = 0; i < Ne; i++) o->last[i] += o->svg[i] / 2.0f;                /* :922 */
  o->svg = o->x2 + Nz;                                                       /* :923 svg becomes a view of x2 */
  o->cnt_ech += Ne / 2;
  memcpy(o->padded + Nz, x + (Ne - h), sizeof(cf32) * h);                    /* :928 */
  return 0;
}

/* Windowed mode through the re-blocking of OLA::step (fourier.cc:813-833).  y must hold Ne * blocks samples;
 * *n_out receives what was emitted (the first block of the stream emits nothing). */
int tsdo_ola_step_fen(tsdo_ola *o, const float *xin, int n, float *yout, int *n_out)
{
  const cf32 *x = (const cf32 *) xin;
  cf32 *y = (cf32 *) yout;
  const int Ne = o->Ne;
  int produced = 0, i = 0, ny = 0;
  while(i < n)
  {
    const cf32 *blk = NULL;
    if(o->windex == 0 && n - i >= Ne)
    {
      blk = x + i;
      i += Ne;
    }
    else
    {
      int take = Ne - o->windex;
      if(take > n - i) take = n - i;
      memcpy(o->tampon + o->windex, x + i, sizeof(cf32) * take);
      o->windex += take;
      i += take;
      if(o->windex == Ne)
      {
        blk = o->tampon;
        o->windex = 0;
      }
    }
    if(blk)
    {
      if(tsdo_ola_block_fen(o, blk, y + produced, &ny)) return 1;
      produced += ny;
    }
  }
  *n_out = produced;
  return 0;
}

/* fourier.cc:813-833 OLA::step + tsd.cc:332-370 TamponNv2::step.  y must hold
 * Ne * ((residual + n) / Ne) samples; *n_out receives that count. */
int tsdo_ola_step(tsdo_ola *o, const float *xin, int n, float *yout, int *n_out)
{
  if(o->avec_fenetrage) return tsdo_ola_step_fen(o, xin, n, yout, n_out);
  const cf32 *x = (const cf32 *) xin;
  cf32 *y = (cf32 *) yout;
  const int Ne = o->Ne;
  int produced = 0, i = 0;
  while(i < n)
  {
    if(o->windex == 0 && n - i >= Ne)
    {
      /* aligned: same samples as copying through the buffer */
      if(tsdo_ola_block(o, x + i, y + produced)) return 1;
      produced += Ne;
      i += Ne;
      continue;
    }
    int take = Ne - o->windex;
    if(take > n - i) take = n - i;
    memcpy(o->tampon + o->windex, x + i, sizeof(cf32) * take);
    o->windex += take;
    i += take;
    if(o->windex == Ne)
    {
      if(tsdo_ola_block(o, o->tampon, y + produced)) return 1;
      produced += Ne;
      o->windex = 0;
    }
  }
  *n_out = produced;
  return 0;
}

/* fourier.cc:962-965  H of the FiltreFFTRIF convention: h2 = zeros(N), h2.tail(K) = h,
 * H = fft(h2) * sqrt(N).  The reference takes the real-input route (RTFRPlan,
 * fourier.cc:280-355); here the complex plan is used, which differs by rounding only
 * (<= a few 1e-7 of max|H|).  H is set-up data handed identically to both sides of a parity test. */
int tsdo_ola_make_H(const float *h, int K, int N, float *H)
{
  tsdo_fft *p = tsdo_fft_new(N);
  if(!p || K > N) return 1;
  cf32 *h2 = (cf32 *) calloc(N, sizeof(cf32));
  for(int i = 0; i < K; i++) h2[N - K + i] = h[i];
  tsdo_fft_step(p, (const float *) h2, 1, H);
  cf32 s = (float) sqrt((double) N);
  cf32 *Hc = (cf32 *) H;
  for(int i = 0; i < N; i++) Hc[i] *= s;
  free(h2);
  tsdo_fft_free(p);
  return 0;
}

/* ------------------------------------------------------------------ resampler */

/* ra.cc:13-79 AdaptationRythmeSimple + filtrage.hpp:1873-1881 InterpolateurRIF::step +
 * itrp.cc:16-22 InterpolateurSinc::coefs.  LUT given as data: lut[p*K + i]. */
typedef struct
{
  float phase, ratio, increment;
  int K, nphases;
  float *lut;
  cf32 *fen;
} tsdo_itrp;

tsdo_itrp *tsdo_itrp_new(float ratio, const float *lut, int K, int nphases)
{
  tsdo_itrp *r = (tsdo_itrp *) calloc(1, sizeof(*r));
  r->ratio = ratio;
  r->increment = 1 / ratio;
  r->phase = 0;
  r->K = K;
  r->nphases = nphases;
  r->lut = (float *) malloc(sizeof(float) * K * (nphases + 1));
  memcpy(r->lut, lut, sizeof(float) * K * (nphases + 1));
  r->fen = (cf32 *) calloc(K, sizeof(cf32));
  return r;
}
void tsdo_itrp_free(tsdo_itrp *r)
{
  if(!r) return;
  free(r->lut);
  free(r->fen);
  free(r);
}
float tsdo_itrp_phase(const tsdo_itrp *r) { return r->phase; }
/* upper bound used by the reference for its temporary (ra.cc:43) */
int tsdo_itrp_capacity(const tsdo_itrp *r, int n) { return (int) (ceilf(r->ratio * n) + 10); }

int tsdo_itrp_step(tsdo_itrp *r, const float *xin, int n, float *yout, int cap, int *n_out)
{
  const cf32 *x = (const cf32 *) xin;
  float *y = yout;
  const int K = r->K;
  int j = 0;
  for(int i = 0; i < n; i++)
  {
    memmove(r->fen, r->fen + 1, sizeof(cf32) * (K - 1));
    r->fen[K - 1] = x[i];
    while(r->phase < 1)
    {
      int idx = (int) (r->phase * r->nphases);
      if(idx < 0 || idx > r->nphases) return 2;
      const float *h = r->lut + (size_t) idx * K;
      float sr = 0, si = 0;
      for(int k = 0; k < K; k++)
      {
        sr += crealf(r->fen[k]) * h[k];
        si += cimagf(r->fen[k]) * h[k];
      }
      if(j >= cap) return 1;
      y[2 * j] = sr;
      y[2 * j + 1] = si;
      j++;
      r->phase += r->increment;
    }
    r->phase--;
  }
  *n_out = j;
  return 0;
}

/* Schedule of the same recurrence without data (ra.cc:58-73): for input sample i of this call,
 * emits (i, lut index) pairs.  Used to check the GPU host scheduler. */
int tsdo_itrp_schedule(float *phase_io, float ratio, int nphases, int n, int32_t *in_idx, int32_t *lut_idx,
                       int cap, int *n_out)
{
  float phase = *phase_io, inc = 1 / ratio;
  int j = 0;
  for(int i = 0; i < n; i++)
  {
    while(phase < 1)
    {
      if(j >= cap) return 1;
      in_idx[j] = i;
      lut_idx[j] = (int) (phase * nphases);
      j++;
      phase += inc;
    }
    phase--;
  }
  *phase_io = phase;
  *n_out = j;
  return 0;
}

/* ra.cc:104-156 stage planner of filtre_reechan (AdaptationRythmeArbitraire::configure_impl) */
void tsdo_reechan_plan(float ratio_, int *ndec, int *nups, float *post, float *fcut, int *use_itrp)
{
  float ratio = ratio_;
  if((ratio <= 0) || isinf(ratio) || (ratio >= 1e9))
    ratio = 1;
  float f = ratio;
  int d = 0, u = 0;
  while(f < 0.5) { d++; f *= 2; }
  while(f >= 2) { u++; f /= 2; }
  *ndec = d;
  *nups = u;
  *post = f;
  *fcut = fminf(0.4f, f / 2);
  *use_itrp = (ratio != 1) && !(fabsf(f - 1) < 1e-6f);
}

/* ------------------------------------------------------------------ polyphase rate-change stages */

/* polyphase.cc:54-341.  kind 0: FiltreRIFUps<T,float>(c, R) (:246-341), 1: FiltreRIFDemiBande<T,float>(c) (:54-149),
 * 2: FiltreRIFDecim<T,float>(c, R) (:156-239); cplx selects T = cfloat (else float).  The loops below are the
 * reference's, pointer walk included (ring `fenêtre`, `index`, counter `odd` / `cnt`). */
typedef struct
{
  int kind, cplx, K, R, index, cnt;
  float *coefs; /* K floats (ups: scaled by R and zero-padded to a multiple of R, :259-269) */
  float *fen;   /* ring: K (decimators) or K/R (ups) samples of T */
} tsdo_poly;

tsdo_poly *tsdo_poly_new(int kind, const float *taps, int K, int R, int cplx)
{
  if(K <= 0 || kind < 0 || kind > 2) return NULL;
  if(kind == 1) R = 2; /* :60 */
  if(R < 1) return NULL;
  tsdo_poly *f = (tsdo_poly *) calloc(1, sizeof(*f));
  f->kind = kind;
  f->cplx = cplx;
  f->R = R;
  int Kp = K;
  if(kind == 0 && (K % R) != 0) Kp = K + (R - (K % R)); /* :264-268 */
  f->K = Kp;
  f->coefs = (float *) calloc(Kp, sizeof(float));
  for(int i = 0; i < K; i++) f->coefs[i] = (kind == 0) ? taps[i] * R : taps[i]; /* coefs = c * R (:258) */
  const int W = (kind == 0) ? Kp / R : Kp;
  f->fen = (float *) calloc((size_t) W * (cplx ? 2 : 1), sizeof(float));
  return f;
}
void tsdo_poly_free(tsdo_poly *f)
{
  if(!f) return;
  free(f->coefs);
  free(f->fen);
  free(f);
}
int tsdo_poly_index(const tsdo_poly *f) { return f->index; }
int tsdo_poly_cnt(const tsdo_poly *f) { return f->cnt; }
/* y.resize(n*R) (:290) / y.resize((n + cnt) / R) (:78,181) */
int tsdo_poly_out_count(const tsdo_poly *f, int n) { return f->kind == 0 ? n * f->R : (n + f->cnt) / f->R; }

#define TSDO_POLY_BODY(T)                                                                                  \
  const T *iptr = (const T *) x;                                                                           \
  T *optr = (T *) y;                                                                                       \
  T *fen = (T *) f->fen;                                                                                   \
  const int K = f->K, R = f->R;                                                                            \
  if(f->kind == 0)                                                                                         \
  {                                                                                                        \
    const int W = K / R;                                                                                   \
    for(int j = 0; j < n; j++)                                                                             \
    {                                                                                                      \
      fen[f->index] = *iptr++;                                                                             \
      f->index = (f->index + 1) % W;                                                                       \
      for(int i = 0; i < R; i++)                                                                           \
      {                                                                                                    \
        T sum = 0;                                                                                         \
        const float *cptr = f->coefs + (R - 1) - i;                                                        \
        const T *wptr = fen + f->index;                                                                    \
        const int K1 = W - f->index, K2 = W - K1;                                                          \
        for(int k = 0; k < K1; k++) { sum += *wptr++ * *cptr; cptr += R; }                                 \
        wptr = fen;                                                                                        \
        for(int k = 0; k < K2; k++) { sum += *wptr++ * *cptr; cptr += R; }                                 \
        *optr++ = sum;                                                                                     \
      }                                                                                                    \
    }                                                                                                      \
  }                                                                                                        \
  else                                                                                                     \
  {                                                                                                        \
    for(int j = 0; j < n; j++)                                                                             \
    {                                                                                                      \
      const float *cptr = f->coefs;                                                                        \
      T somme = 0;                                                                                         \
      fen[f->index] = *iptr++;                                                                             \
      f->index = (f->index + 1) % K;                                                                       \
      if(f->cnt < R - 1) { f->cnt++; continue; }                                                           \
      f->cnt = 0;                                                                                          \
      const T *wptr = fen + f->index;                                                                      \
      const int K1 = K - f->index, K2 = K - K1;                                                            \
      if(f->kind == 2)                                                                                     \
      {                                                                                                    \
        for(int i = 0; i < K1; i++) somme += *wptr++ * *cptr++;                                            \
        wptr = fen;                                                                                        \
        for(int i = 0; i < K2; i++) somme += *wptr++ * *cptr++;                                            \
      }                                                                                                    \
      else                                                                                                 \
      {                                                                                                    \
        int i;                                                                                             \
        for(i = 0; i < K1; i += 2) { somme += *wptr * *cptr; wptr += 2; cptr += 2; }                       \
        i = i - K1;                                                                                        \
        wptr = fen + i;                                                                                    \
        for(; i < K2; i += 2) { somme += *wptr * *cptr; wptr += 2; cptr += 2; }                            \
        somme += 0.5f * fen[(f->index + K / 2) % K];                                                       \
      }                                                                                                    \
      *optr++ = somme;                                                                                     \
    }                                                                                                      \
  }

void tsdo_poly_step(tsdo_poly *f, const float *x, int n, float *y)
{
  if(f->cplx) { TSDO_POLY_BODY(cf32) }
  else { TSDO_POLY_BODY(float) }
}
