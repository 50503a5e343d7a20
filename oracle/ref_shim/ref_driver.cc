// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// C-ABI wrappers around the reference's own factories so that Python tests can
// drive the UNMODIFIED libtsd CPU implementation (compiled in place from
// /root/reference by oracle/Makefile into oracle/_ref/libtsdref.so).
// Every entry returns 0 on success, non-zero on failure (message through
// tsdref_last_error()), mirroring the reference's "exceptions only" error model
// (commun.hpp:152-163, tsd.cc:114-118).
#include "tsd/tsd.hpp"
#include "tsd/filtrage.hpp"
#include "tsd/fourier.hpp"
#include "tsd/filtrage/spline.hpp"

#include <cstring>
#include <stdexcept>
#include <string>

using namespace tsd;
using namespace tsd::filtrage;
using namespace tsd::fourier;

static thread_local std::string g_err;

// Same policy as the reference's default logger (tsd.cc:45-119) minus the printing.
static void quiet_logger(const char *, entier, entier niveau, cstring str)
{
  if(niveau >= 4) throw std::runtime_error(str);
}
static void __attribute__((constructor(101))) install_logger() { get_logger() = quiet_logger; }

template<typename F> static int guarded(F &&f)
{
  get_logger() = quiet_logger;
  try { f(); return 0; }
  catch(const std::exception &e) { g_err = e.what(); }
  catch(const std::string &s) { g_err = s; }
  catch(...) { g_err = "unknown exception"; }
  return 1;
}

struct RefFilter
{
  bool cplx = true;
  sptr<FiltreGen<float>> fr;
  sptr<FiltreGen<cfloat>> fc;
  sptr<Interpolateur<cfloat>> keep_c;
  sptr<Interpolateur<float>> keep_r;
  Veccf H;   // spectral gain captured by the OLA callback
};

template<typename T> static sptr<InterpolateurRIF<T>> make_itrp(int kind, int ncoefs, int nphases, float fcut, int degree)
{
  if(kind == 0) return itrp_sinc<T>({ncoefs, nphases, fcut, "hn"});
  if(kind == 1) return itrp_cspline<T>();
  if(kind == 2) return itrp_lineaire<T>();
  return itrp_lagrange<T>(degree);
}

extern "C" {

const char *tsdref_last_error() { return g_err.c_str(); }

// tsd.cc:287-291
int tsdref_p2(int i) { return prochaine_puissance_de_2(i); }

// fourier.cc:715-735
int tsdref_ola_complexite_optimise(int M, float *C, int *Nf, int *Nz, int *Ne)
{
  return guarded([&] { ola_complexité_optimise(M, *C, *Nf, *Nz, *Ne); });
}

// rif-fen.cc:30-107
int tsdref_design_rif_fen(int n, const char *type, float fc, const char *fen, float *h)
{
  return guarded([&] {
    Vecf v = design_rif_fen(n, type, fc, fen);
    memcpy(h, v.data(), sizeof(float) * n);
  });
}

// itrp.cc:10-55 : LUT column p (p = 0..nphases) as used by coefs(tau)
int tsdref_itrp_sinc_lut(int ncoefs, int nphases, float fcut, const char *fen, float *lut)
{
  return guarded([&] {
    auto it = itrp_sinc<cfloat>({ncoefs, nphases, fcut, fen});
    for(int p = 0; p <= nphases; p++)
    {
      // tau chosen in the interior of LUT cell p so that (int)(tau*nphases) == p
      float tau = (p == nphases) ? 1.0f : (p + 0.5f) / nphases;
      Vecf c = it->coefs(tau);
      memcpy(lut + (size_t) p * ncoefs, c.data(), sizeof(float) * ncoefs);
    }
  });
}

// kind: 0 = filtre_rif<float,float>, 1 = filtre_rif<float,cfloat>, 2 = filtre_rif<cfloat,cfloat>
// (filtre-rt.cc:171-175,816-818)
void *tsdref_fir_new(int kind, const float *taps, int K)
{
  RefFilter *f = new RefFilter;
  int rc = guarded([&] {
    if(kind == 0) { f->cplx = false; f->fr = filtre_rif<float, float>(Vecf::map(taps, K).clone()); }
    else if(kind == 1) f->fc = filtre_rif<float, cfloat>(Vecf::map(taps, K).clone());
    else f->fc = filtre_rif<cfloat, cfloat>(Veccf::map((const cfloat *) taps, K).clone());
  });
  if(rc) { delete f; return nullptr; }
  return f;
}

// filtre_rif_fft<T>(h) (fourier.cc:946-990); kind 0 = float, 1 = cfloat
void *tsdref_rif_fft_new(int kind, const float *taps, int K)
{
  RefFilter *f = new RefFilter;
  int rc = guarded([&] {
    if(kind == 0) { f->cplx = false; f->fr = filtre_rif_fft<float>(Vecf::map(taps, K).clone()); }
    else f->fc = filtre_rif_fft<cfloat>(Vecf::map(taps, K).clone());
  });
  if(rc) { delete f; return nullptr; }
  return f;
}

// filtre_fft(config) with traitement_freq = "X *= H" (fourier.cc:935-940,956-959).
// H == NULL gives the identity callback.
// avec_fenetrage != 0 selects the Hann-window, 50 % overlap mode (fourier.cc:884-930).
void *tsdref_ola_new2(int Ne, int nb_zeros_min, const float *H, int N_expected, int *N_out, int avec_fenetrage)
{
  RefFilter *f = new RefFilter;
  int rc = guarded([&] {
    FiltreFFTConfig cfg;
    cfg.dim_blocs_temporel = Ne;
    cfg.nb_zeros_min = nb_zeros_min;
    cfg.avec_fenetrage = avec_fenetrage != 0;
    if(H)
    {
      f->H = Veccf::map((const cfloat *) H, N_expected).clone();
      RefFilter *self = f;
      cfg.traitement_freq = [self](Veccf &X) { X *= self->H; };
    }
    else
      cfg.traitement_freq = [](Veccf &) {};
    auto [flt, N] = filtre_fft(cfg);
    f->fc = flt;
    if(N_out) *N_out = N;
    if(H && N != N_expected) échec("tsdref_ola_new: N = {} but H has {} bins", N, N_expected);
  });
  if(rc) { delete f; return nullptr; }
  return f;
}

void *tsdref_ola_new(int Ne, int nb_zeros_min, const float *H, int N_expected, int *N_out)
{
  return tsdref_ola_new2(Ne, nb_zeros_min, H, N_expected, N_out, 0);
}

// fenêtre(nom, n, sym) as the OLA object calls it (fourier.cc:796; built by the shim, tab_shim.cc)
int tsdref_fenetre(const char *nom, int n, int sym, float *w)
{
  return guarded([&] {
    Vecf v = tsd::filtrage::fenêtre(std::string(nom), n, sym != 0);
    memcpy(w, v.data(), sizeof(float) * n);
  });
}

// periodogramme_tfd(x, N) (fourier.cc:1451-1481): out[rows][cols] row-major (rows = frames, cols = N2 / 2)
int tsdref_periodogramme_tfd(const float *x, int n, int N, float *out, int cap, int *rows, int *cols)
{
  return guarded([&] {
    Tabf M = tsd::tf::periodogramme_tfd(Veccf::map((const cfloat *) x, n).clone(), N);
    *rows = M.rows();
    *cols = M.cols();
    if((long long) M.rows() * M.cols() > cap) échec("tsdref_periodogramme_tfd: capacity");
    for(int i = 0; i < M.rows(); i++)
      for(int j = 0; j < M.cols(); j++) out[(size_t) i * M.cols() + j] = M(i, j);
  });
}

// rt_spectrum(SpectrumConfig) (fourier.hpp:909-952, fourier.cc:1162-1343): averaged power spectrum in dB
void *tsdref_spectrum_new(int BS, int nmeans, int nsubs, int sweep_active, int sweep_step, int masque_bf, int masque_hf, int fenetre,
                          int *Nf, int *Ns)
{
  void *res = nullptr;
  guarded([&] {
    tsd::fourier::SpectrumConfig c;
    c.BS = BS;
    c.nmeans = nmeans;
    c.nsubs = nsubs;
    c.sweep.active = sweep_active != 0;
    c.sweep.step = sweep_step;
    c.sweep.masque_bf = masque_bf;
    c.sweep.masque_hf = masque_hf;
    c.fenetre = (tsd::filtrage::Fenetre) fenetre;
    *Nf = c.Nf();
    *Ns = c.Ns();
    res = new sptr<Filtre<cfloat, float, tsd::fourier::SpectrumConfig>>(tsd::fourier::rt_spectrum(c));
  });
  return res;
}
// one block of BS samples; *n_out = length of the spectrum this call returned (0 while the average is incomplete)
int tsdref_spectrum_step(void *h, const float *x, int n, float *y, int cap, int *n_out)
{
  return guarded([&] {
    auto &f = *(sptr<Filtre<cfloat, float, tsd::fourier::SpectrumConfig>> *) h;
    Vecf r;
    f->step(Veccf::map((const cfloat *) x, n).clone(), r);
    *n_out = r.rows();
    if(r.rows() > cap) échec("tsdref_spectrum_step: capacity");
    if(r.rows() > 0) memcpy(y, r.data(), sizeof(float) * r.rows());
  });
}
void tsdref_spectrum_free(void *h) { delete (sptr<Filtre<cfloat, float, tsd::fourier::SpectrumConfig>> *) h; }

// rééchan_freq<T>(x, lom) (fourier.cc:1391-1419); kind 0 = float, 1 = cfloat.  y needs round(n * lom) elements.
int tsdref_reechan_freq(int kind, const void *x, int n, float lom, void *y, int cap, int *n_out)
{
  return guarded([&] {
    if(kind == 0)
    {
      Vecf r = rééchan_freq<float>(Vecf::map((const float *) x, n).clone(), lom);
      *n_out = r.rows();
      if(r.rows() > cap) échec("tsdref_reechan_freq: capacity");
      memcpy(y, r.data(), sizeof(float) * r.rows());
    }
    else
    {
      Veccf r = rééchan_freq<cfloat>(Veccf::map((const cfloat *) x, n).clone(), lom);
      *n_out = r.rows();
      if(r.rows() > cap) échec("tsdref_reechan_freq: capacity");
      memcpy(y, r.data(), sizeof(cfloat) * r.rows());
    }
  });
}

// H of the FiltreFFTRIF convention (fourier.cc:962-965): h2.tail(K) = h ; H = fft(h2) * sqrt(N)
int tsdref_ola_make_H(const float *h, int K, int N, float *H)
{
  return guarded([&] {
    Vecf h2 = Vecf::zeros(N);
    h2.tail(K) = Vecf::map(h, K);
    Veccf Hv = fft(h2);
    Hv *= sqrt(N);
    memcpy(H, Hv.data(), sizeof(cfloat) * N);
  });
}

// filtre_itrp<cfloat>(ratio, itrp_sinc<cfloat>({ncoefs,nphases,fcut,"hn"})) (ra.cc:185-188)
void *tsdref_itrp_new(float ratio, int ncoefs, int nphases, float fcut)
{
  RefFilter *f = new RefFilter;
  int rc = guarded([&] {
    f->keep_c = itrp_sinc<cfloat>({ncoefs, nphases, fcut, "hn"});
    f->fc = filtre_itrp<cfloat>(ratio, f->keep_c);
  });
  if(rc) { delete f; return nullptr; }
  return f;
}

// filtre_itrp<T>(ratio, itrp) with any of the reference's interpolators (itrp.cc:130-157) and T = float (cplx = 0) or
// cfloat: kind 0 = itrp_sinc({ncoefs, nphases, fcut, "hn"}), 1 = itrp_cspline, 2 = itrp_lineaire, 3 = itrp_lagrange(degree)
void *tsdref_itrp_new2(float ratio, int kind, int cplx, int ncoefs, int nphases, float fcut, int degree)
{
  RefFilter *f = new RefFilter;
  f->cplx = cplx != 0;
  int rc = guarded([&] {
    if(cplx)
    {
      f->keep_c = make_itrp<cfloat>(kind, ncoefs, nphases, fcut, degree);
      f->fc = filtre_itrp<cfloat>(ratio, f->keep_c);
    }
    else
    {
      f->keep_r = make_itrp<float>(kind, ncoefs, nphases, fcut, degree);
      f->fr = filtre_itrp<float>(ratio, f->keep_r);
    }
  });
  if(rc) { delete f; return nullptr; }
  return f;
}
// filtre_reechan<float>(ratio) (ra.cc:190-191), the instantiation tests/test-ra.cc:161-164 drives
void *tsdref_reechan_new_f32(float ratio)
{
  RefFilter *f = new RefFilter;
  f->cplx = false;
  int rc = guarded([&] { f->fr = filtre_reechan<float>(ratio); });
  if(rc) { delete f; return nullptr; }
  return f;
}
// cspline_calc_lut(n, c) (itrp.cc:314-320): [4][n+1] column-major -> lut[p*4 + i]
int tsdref_cspline_lut(int n, float c, float *lut)
{
  return guarded([&] {
    Tabf L = cspline_calc_lut(n, c);
    for(int p2 = 0; p2 <= n; p2++)
      for(int i = 0; i < 4; i++) lut[p2 * 4 + i] = L(i, p2);
  });
}

// détecteur_création(config) (detection.cc:68-516): normalised-correlation detector, OLA (mode 0) or FIR (mode 1)
// correlator.  step() returns the score signal and the detections the reference hands to gere_detection, as rows of
// {position, position_prec, score, gain, theta, SNR_dB, sigma_noise}.
struct RefDetect
{
  sptr<Detecteur> d;
  std::vector<Detection> dets;
};
void *tsdref_detect_new(const float *motif, int M, int Ne, float seuil, int mode)
{
  RefDetect *r = new RefDetect;
  int rc = guarded([&] {
    DetecteurConfig c;
    c.Ne = Ne;
    c.motif = Veccf::map((const cfloat *) motif, M).clone();
    c.seuil = seuil;
    c.mode = mode ? DetecteurConfig::MODE_RIF : DetecteurConfig::MODE_OLA;
    c.gere_detection = [r](const Detection &det) { r->dets.push_back(det); };
    r->d = détecteur_création(c);
  });
  if(rc) { delete r; return nullptr; }
  return r;
}
int tsdref_detect_step(void *h, const float *x, int n, float *score, float *dets, int cap, int *ndet)
{
  RefDetect *r = (RefDetect *) h;
  return guarded([&] {
    r->dets.clear();
    const Veccf xv = Veccf::map((const cfloat *) x, n);
    Vecf y;
    r->d->step(xv, y);
    if(y.rows() != n) échec("tsdref_detect_step: {} scores for {} samples", y.rows(), n);
    memcpy(score, y.data(), sizeof(float) * n);
    *ndet = (int) r->dets.size();
    if(*ndet > cap) échec("tsdref_detect_step: {} detections > capacity {}", *ndet, cap);
    for(int i = 0; i < *ndet; i++)
    {
      const Detection &d = r->dets[i];
      float *o = dets + 7 * i;
      o[0] = (float) d.position; o[1] = d.position_prec; o[2] = d.score; o[3] = d.gain; o[4] = d.θ; o[5] = d.SNR_dB; o[6] = d.σ_noise;
    }
  });
}
void tsdref_detect_free(void *h) { delete(RefDetect *) h; }

// filtre_reechan<cfloat>(ratio) (ra.cc:180-183) = what resample()/rééchan() builds (tsd.hpp:700-705)
void *tsdref_reechan_new(float ratio)
{
  RefFilter *f = new RefFilter;
  int rc = guarded([&] { f->fc = filtre_reechan<cfloat>(ratio); });
  if(rc) { delete f; return nullptr; }
  return f;
}

// polyphase.cc: kind 0 = filtre_rif_demi_bande<float,cfloat>(h), 1 = filtre_rif_ups<float,cfloat>(h,R),
// 2 = filtre_rif_decim<float,cfloat>(h,R)
void *tsdref_polyphase_new(int kind, const float *taps, int K, int R)
{
  RefFilter *f = new RefFilter;
  int rc = guarded([&] {
    Vecf h = Vecf::map(taps, K).clone();
    if(kind == 0) f->fc = filtre_rif_demi_bande<float, cfloat>(h);
    else if(kind == 1) f->fc = filtre_rif_ups<float, cfloat>(h, R);
    else f->fc = filtre_rif_decim<float, cfloat>(h, R);
  });
  if(rc) { delete f; return nullptr; }
  return f;
}

int tsdref_filter_is_complex(void *h) { return ((RefFilter *) h)->cplx ? 1 : 0; }

// FiltreGen<T>::step(x, y) (tsd.hpp:626-657).  x, y in units of T (float or cfloat).
// On return *n_out = y.rows(); fails if it exceeds cap.
int tsdref_filter_step(void *h, const void *x, int n, void *y, int cap, int *n_out)
{
  RefFilter *f = (RefFilter *) h;
  return guarded([&] {
    if(f->cplx)
    {
      const Veccf xv = Veccf::map((const cfloat *) x, n);
      Veccf yv;
      f->fc->step(xv, yv);
      *n_out = yv.rows();
      if(yv.rows() > cap) échec("tsdref_filter_step: output {} > capacity {}", yv.rows(), cap);
      if(yv.rows()) memcpy(y, yv.data(), sizeof(cfloat) * yv.rows());
    }
    else
    {
      const Vecf xv = Vecf::map((const float *) x, n);
      Vecf yv;
      f->fr->step(xv, yv);
      *n_out = yv.rows();
      if(yv.rows() > cap) échec("tsdref_filter_step: output {} > capacity {}", yv.rows(), cap);
      if(yv.rows()) memcpy(y, yv.data(), sizeof(float) * yv.rows());
    }
  });
}

void tsdref_filter_free(void *h) { delete(RefFilter *) h; }

// FFTPlan (fourier.hpp:19-32) obtained through tfrplan_création (fourier.cc:475-481)
void *tsdref_fftplan_new(int n, int forward)
{
  sptr<FFTPlan> *p = new sptr<FFTPlan>;
  int rc = guarded([&] { *p = tfrplan_création(n, forward != 0, true); });
  if(rc) { delete p; return nullptr; }
  return p;
}
int tsdref_fftplan_step(void *plan, const float *x, int n, int forward, float *y)
{
  sptr<FFTPlan> &p = *(sptr<FFTPlan> *) plan;
  return guarded([&] {
    const Veccf xv = Veccf::map((const cfloat *) x, n);
    Veccf yv;
    p->step(xv, yv, forward != 0);
    memcpy(y, yv.data(), sizeof(cfloat) * yv.rows());
  });
}
void tsdref_fftplan_free(void *plan) { delete(sptr<FFTPlan> *) plan; }

// rfft (fourier.hpp:116-122, fourier.cc:280-355): n real -> n complex
int tsdref_rfft(const float *x, int n, float *y)
{
  return guarded([&] {
    Veccf yv = rfft(Vecf::map(x, n).clone());
    memcpy(y, yv.data(), sizeof(cfloat) * yv.rows());
  });
}

// filtrer(h, x) one-shot (filtrage.hpp:1684-1711) = README example path, float data
int tsdref_filtrer_f32(const float *taps, int K, const float *x, int n, float *y)
{
  return guarded([&] {
    Vecf h = Vecf::map(taps, K).clone();
    Vecf xv = Vecf::map(x, n).clone();
    Vecf yv = filtrer(h, xv);
    if(yv.rows() != n) échec("filtrer: len(y) = {} != {}", yv.rows(), n);
    memcpy(y, yv.data(), sizeof(float) * n);
  });
}

// TamponNv2 (tsd.cc:307-380): feeds chunk sizes, records the size of every emitted block.
int tsdref_tampon_trace(int N, const int *chunks, int nchunks, int *emitted, int cap, int *n_emitted)
{
  return guarded([&] {
    int cnt = 0;
    float marker = 0;
    std::vector<float> firsts;
    auto t = tampon_création<float>(N, [&](const Vecteur<float> &b) {
      if(cnt < cap) emitted[cnt] = b.rows();
      cnt++;
    });
    for(int i = 0; i < nchunks; i++)
    {
      Vecf v = Vecf::zeros(chunks[i]);
      for(int k = 0; k < chunks[i]; k++) v(k) = marker++;
      t->step(v);
    }
    *n_emitted = cnt;
  });
}

}
