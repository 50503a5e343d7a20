// tsd_gpu_adapters.hpp — the binding a libtsd maintainer adds to route the filtering hot path to the
// B200 library.  It is compiled AGAINST THE REFERENCE'S OWN HEADERS (core/include/tsd/*.hpp) and only
// forwards to the C ABI of include/tsdgpu.h; no arithmetic lives here.
//
//   #include "tsd/tsd.hpp" / "tsd/filtrage.hpp" / "tsd/fourier.hpp"   (reference)
//   #include "tsdgpu.h"                                               (this repo)
//   link: -ltsdgpu
//
// Drop-in points (reference file:line):
//   filtre_rif<Tc,T>(h)            core/src/filtrage/filtre-rt.cc:171-175   -> tsd::gpu::filtre_rif_gpu<Tc,T>(h)
//   filtre_fft(config)             core/src/fourier/fourier.cc:935-940      -> tsd::gpu::filtre_fft_gpu(config, H, K)
//   filtre_itrp<T>(r, itrp)        core/src/reechan/ra.cc:185-195           -> tsd::gpu::filtre_itrp_gpu<T>(r, itrp)
//   filtre_rif_fft<T>(h)           core/src/fourier/fourier.cc:946-990      -> tsd::gpu::filtre_rif_fft_gpu<T>(h)
//   fftplan_defaut (global hook)   core/src/fourier/fourier.cc:469-472      -> tsd::gpu::installe_fftplan_gpu()
//   filtre_rif_ups / _demi_bande / _decim   core/src/reechan/polyphase.cc:344-360 -> tsd::gpu::filtre_rif_*_gpu<T>(c[, R])
//   filtre_reechan<T>(ratio)       core/src/reechan/ra.cc:180-183,190-191   -> tsd::gpu::filtre_reechan_gpu<T>(ratio)
// integration/tsd_gpu_dropin.cc DEFINES the reference's own factory symbols on top of these classes, so that linking it
// in front of the library routes filtre_rif / filtrer / filter_fir / filtre_reechan / resample / filtre_itrp /
// filtre_rif_fft / filtre_fft to the GPU with no source change in the callers.
#pragma once
#include "tsd/tsd.hpp"
#include "tsd/filtrage.hpp"
#include "tsd/fourier.hpp"
#include "tsdgpu.h"

#include <cstdlib>
#include <exception>
#include <string>
#include <tuple>
#include <vector>

namespace tsd::gpu {

// same error model as the reference: log at level 4 (the default logger throws), then throw (commun.hpp:152-157)
inline void verifie(int rc, const char *quoi)
{
  if(rc != 0) échec("{} : {}", quoi, tsdgpu_last_error());
}

template<typename T, typename Tc> struct FiltreRIFGpu: FiltreGen<T>
{
  tsdgpu_fir_t h = nullptr;
  FiltreRIFGpu(const Vecteur<Tc> &c)
  {
    constexpr int kind = std::is_same<T, float>::value ? TSDGPU_FIR_F32_F32
                         : (std::is_same<Tc, float>::value ? TSDGPU_FIR_CF32_F32 : TSDGPU_FIR_CF32_CF32);
    verifie(tsdgpu_fir_create(kind, (const float *) c.data(), c.rows(), 1, &h), "filtre_rif (gpu)");
  }
  ~FiltreRIFGpu() { tsdgpu_fir_destroy(h); }
  void step(const Vecteur<T> &x, Vecteur<T> &y) override
  {
    soit n = x.rows();
    si(x.data() != y.data())      // same in-place rule as FiltreRIF::step (filtre-rt.cc:76-80)
      y.resize(n);
    verifie(tsdgpu_fir_step(h, x.data(), n, n, y.data(), n, TSDGPU_HOST), "filtre_rif::step (gpu)");
  }
};

template<typename Tc, typename T> sptr<FiltreGen<T>> filtre_rif_gpu(const Vecteur<Tc> &c)
{
  retourne std::make_shared<FiltreRIFGpu<T, Tc>>(c);
}

// FFTPlan (fourier.hpp:19-32).  Like TFRPlanDefaut: always unitary, re-plans when the size changes.
struct FFTPlanGpu: FiltreGen<cfloat>, tsd::fourier::FFTPlan
{
  tsdgpu_fft_t h = nullptr;
  entier n = -1;
  bouléen avant = oui;
  ~FFTPlanGpu() { tsdgpu_fft_destroy(h); }
  void configure(entier n_, bouléen avant_, bouléen) override
  {
    avant = avant_;
    si(n_ == n || n_ < 0) retourne;
    tsdgpu_fft_destroy(h);
    h = nullptr;
    n = n_;
    verifie(tsdgpu_fft_plan(n, 1, &h), "tfrplan (gpu)");
  }
  void step(const Veccf &x, Veccf &y) override { step(x, y, avant); }
  void step(const Veccf &x, Veccf &y, bouléen av) override
  {
    assertion(x.rows() > 0);
    si(x.rows() != n) configure(x.rows(), avant, oui);
    y.resize(n);
    verifie(tsdgpu_fft_exec(h, x.data(), n, y.data(), n, av ? 1 : 0, TSDGPU_HOST), "tfrplan::step (gpu)");
  }
};
// every fft()/ifft()/rfft()/Spectrum in the library then uses the GPU plan (any n: 2^k, even split, chirp-z)
inline void installe_fftplan_gpu()
{
  tsd::fourier::fftplan_defaut = []() -> sptr<tsd::fourier::FFTPlan> { retourne std::make_shared<FFTPlanGpu>(); };
}

// filtre_fft(config).  Two modes:
//  * gains as data: the callback of FiltreFFTRIF, "X *= H" (fourier.cc:956-959), is replaced by the vector H itself; K > 0
//    declares H = fft([0^(N-K), h]) * sqrt(N) (fourier.cc:962-965) and selects the single-SM overlap-save kernel;
//  * generic: no H given -> config.traitement_freq, an arbitrary host std::function (fourier.hpp:319), is called once per
//    transformed block exactly like the reference does (fourier.cc:863,895,915); the spectra travel device -> host ->
//    callback -> device.  An exception thrown by the callback is re-thrown from step().
struct OLAGpu: Filtre<cfloat, cfloat, tsd::fourier::FiltreFFTConfig>
{
  tsdgpu_ola_t h = nullptr;
  Veccf H;
  entier K = 0, N = 0;
  bouléen generique = non;
  std::exception_ptr erreur_rappel;
  OLAGpu(const Veccf &H_, entier K_): H(H_), K(K_) {}
  OLAGpu(): generique(oui) {}
  ~OLAGpu() { tsdgpu_ola_destroy(h); }
  static void rappel(void *user, int, float *X, int n)
  {
    soit moi = (OLAGpu *) user;
    si(moi->erreur_rappel) retourne;
    try
    {
      Veccf Xv = Veccf::map((cfloat *) X, n);     // non-owning view: the callback works in place
      moi->lis_config().traitement_freq(Xv);
    }
    catch(...) { moi->erreur_rappel = std::current_exception(); }
  }
  void configure_impl(const tsd::fourier::FiltreFFTConfig &c) override
  {
    tsdgpu_ola_destroy(h);
    h = nullptr;
    const float *Hp = H.rows() ? (const float *) H.data() : nullptr;
    Vecf fen;
    si(c.avec_fenetrage)
    {
      // Hann-window 50 % overlap mode (fourier.cc:794-798,884-930): same window as the reference object
      soit Ne = c.dim_blocs_temporel > 0 ? c.dim_blocs_temporel : 512;
      fen = tsd::filtrage::fenêtre("hn", Ne, non);
    }
    si(generique && c.traitement_freq)
      verifie(tsdgpu_ola_create_cb(c.dim_blocs_temporel, c.nb_zeros_min, &OLAGpu::rappel, this,
                                   c.avec_fenetrage ? fen.data() : nullptr, 1, &h), "filtre_fft (gpu)");
    sinon si(c.avec_fenetrage)
      verifie(tsdgpu_ola_create_fen(c.dim_blocs_temporel, c.nb_zeros_min, Hp, fen.data(), 1, &h), "filtre_fft (gpu)");
    sinon
      verifie(tsdgpu_ola_create(c.dim_blocs_temporel, c.nb_zeros_min, Hp, K, 1, &h), "filtre_fft (gpu)");
    tsdgpu_ola_dims(h, nullptr, &N, nullptr, nullptr);
  }
  void step(const Veccf &x, Veccf &y) override
  {
    y.resize((entier) tsdgpu_ola_out_count(h, x.rows()));
    long long n_out = 0;
    soit rc = tsdgpu_ola_step(h, x.data(), x.rows(), x.rows(), y.data(), std::max(1, y.rows()), &n_out, TSDGPU_HOST);
    si(erreur_rappel)
    {
      soit e = erreur_rappel;
      erreur_rappel = nullptr;
      std::rethrow_exception(e);
    }
    verifie(rc, "filtre_fft::step (gpu)");
  }
};
inline std::tuple<sptr<Filtre<cfloat, cfloat, tsd::fourier::FiltreFFTConfig>>, entier>
filtre_fft_gpu(const tsd::fourier::FiltreFFTConfig &config, const Veccf &H, entier K = 0)
{
  soit res = std::make_shared<OLAGpu>(H, K);
  res->configure(config);
  retourne {res, res->N};
}
// exact signature of the reference factory (fourier.hpp:370): the callback stays a callback
inline std::tuple<sptr<Filtre<cfloat, cfloat, tsd::fourier::FiltreFFTConfig>>, entier>
filtre_fft_gpu(const tsd::fourier::FiltreFFTConfig &config)
{
  soit res = std::make_shared<OLAGpu>();
  res->configure(config);
  retourne {res, res->N};
}

// filtre_rif_fft<T>(h) (fourier.cc:946-990): Ne = 512, nb_zeros_min = K, H = rfft(h2) * sqrt(N) with h2.tail(K) = h, and the
// reference's output convention y = real(ola.step(x.as_complex())) (fourier.cc:976) for BOTH instantiations.
template<typename T> struct FiltreFFTRIFGpu: FiltreGen<T>
{
  sptr<Filtre<cfloat, cfloat, tsd::fourier::FiltreFFTConfig>> ola;
  FiltreFFTRIFGpu(const Vecf &h)
  {
    tsd::fourier::FiltreFFTConfig c;
    c.nb_zeros_min = h.rows();
    c.dim_blocs_temporel = 512;                                   // fourier.cc:954-960 builds the object with the default
    soit N = prochaine_puissance_de_2(512 + h.rows());
    Vecf h2 = Vecf::zeros(N);
    h2.tail(h.rows()) = h;
    Veccf H = tsd::fourier::rfft(h2);                             // fourier.cc:962-965 (plan from fftplan_defaut)
    H *= std::sqrt((float) N);
    ola = std::get<0>(filtre_fft_gpu(c, H, h.rows()));
  }
  void step(const Vecteur<T> &x, Vecteur<T> &y) override
  {
    Veccf yc;
    si constexpr(std::is_same<T, cfloat>::value) ola->step(x, yc);
    sinon ola->step(x.as_complex(), yc);
    y = real(yc);                                                 // converts back to T (fourier.cc:976)
  }
};
template<typename T> sptr<FiltreGen<T>> filtre_rif_fft_gpu(const Vecf &h) { retourne std::make_shared<FiltreFFTRIFGpu<T>>(h); }

// filtre_itrp<T>(ratio, itrp), T = float or cfloat (ra.cc:185-195).  The interpolator classes are private to itrp.cc; what
// is public is coefs(tau), K and the description `nom` they all set in their constructor:
//   "sinc - ncoefs=.., nphases=P, .."  LUT of P + 1 columns, column (int)(tau * P)      (itrp.cc:16-22,43-44)
//   "cspline"                          LUT of 257 columns, column (int)(tau * 256)      (itrp.cc:62-67,74)
//   "linéaire", "Lagrange degré d"     coefficients evaluated at the EXACT tau          (itrp.cc:84-87,113-127)
// LUT interpolators are handed over as their table (read through coefs() at a delay inside every cell), the exact ones as
// (kind, degree): the device evaluates the same float32 expressions at the float32 phase of every output.  Anything else
// is refused rather than sampled wrongly.
template<typename T> struct AdaptationRythmeGpu: FiltreGen<T>
{
  tsdgpu_resamp_t h = nullptr;
  AdaptationRythmeGpu(float ratio, sptr<tsd::filtrage::InterpolateurRIF<T>> itrp, entier nphases_force = 0)
  {
    constexpr int cplx = std::is_same<T, cfloat>::value ? 1 : 0;
    const std::string &nom = itrp->nom;
    soit K = itrp->K;
    si(nphases_force <= 0 && nom.rfind("lin", 0) == 0)
    {
      verifie(tsdgpu_resamp_create_exact(ratio, TSDGPU_ITRP_LINEAIRE, 1, cplx, 1, &h), "filtre_itrp (gpu)");
      retourne;
    }
    si(nphases_force <= 0 && nom.rfind("Lagrange", 0) == 0)
    {
      verifie(tsdgpu_resamp_create_exact(ratio, TSDGPU_ITRP_LAGRANGE, K - 1, cplx, 1, &h), "filtre_itrp (gpu)");
      retourne;
    }
    entier nphases = nphases_force;
    si(nphases <= 0)
    {
      si(nom == "cspline") nphases = 256;
      sinon si(nom.rfind("sinc", 0) == 0)
      {
        soit pos = nom.find("nphases=");
        si(pos != std::string::npos) nphases = std::atoi(nom.c_str() + pos + 8);
      }
    }
    si(nphases <= 0)
      échec("filtre_itrp (gpu) : interpolateur \"{}\" inconnu du chemin GPU (ni LUT sinc / cspline, ni linéaire / Lagrange).", nom);
    Vecf lut(K * (nphases + 1));
    pour(auto p = 0; p <= nphases; p++)
    {
      // tau inside LUT cell p so that (int)(tau*nphases) == p (itrp.cc:16-22)
      soit c = itrp->coefs(p == nphases ? 1.0f : (p + 0.5f) / nphases);
      pour(auto i = 0; i < K; i++) lut(p * K + i) = c(i);
    }
    verifie(tsdgpu_resamp_create_ex(ratio, lut.data(), K, nphases, cplx, 1, &h), "filtre_itrp (gpu)");
  }
  ~AdaptationRythmeGpu() { tsdgpu_resamp_destroy(h); }
  void step(const Vecteur<T> &x, Vecteur<T> &y) override
  {
    soit n = x.rows();
    y.resize((entier) tsdgpu_resamp_out_count(h, n));
    si(n == 0) retourne;                                  // ra.cc:45-49
    long long n_out = 0;
    verifie(tsdgpu_resamp_step(h, x.data(), n, n, y.rows() ? y.data() : nullptr, std::max(1, y.rows()), y.rows(), &n_out, TSDGPU_HOST),
            "filtre_itrp::step (gpu)");
  }
};
template<typename T = cfloat>
sptr<FiltreGen<T>> filtre_itrp_gpu(float ratio, sptr<tsd::filtrage::InterpolateurRIF<T>> itrp, entier nphases = 0)
{
  retourne std::make_shared<AdaptationRythmeGpu<T>>(ratio, itrp, nphases);
}

// polyphase.cc stages: filtre_rif_ups<float,T>(c, R), filtre_rif_demi_bande<float,T>(c), filtre_rif_decim<float,T>(c, R)
template<typename T> struct FiltrePolyphaseGpu: FiltreGen<T>
{
  tsdgpu_poly_t h = nullptr;
  FiltrePolyphaseGpu(int kind, const Vecf &c, entier R)
  {
    verifie(tsdgpu_poly_create(kind, c.data(), c.rows(), R, std::is_same<T, cfloat>::value ? 1 : 0, 1, &h), "filtre polyphase (gpu)");
  }
  ~FiltrePolyphaseGpu() { tsdgpu_poly_destroy(h); }
  void step(const Vecteur<T> &x, Vecteur<T> &y) override
  {
    soit n = x.rows();
    y.resize((entier) tsdgpu_poly_out_count(h, n));
    long long n_out = 0;
    verifie(tsdgpu_poly_step(h, x.data(), n, n, y.rows() ? y.data() : nullptr, std::max(1, y.rows()), &n_out, TSDGPU_HOST),
            "filtre polyphase::step (gpu)");
  }
};
template<typename T> sptr<FiltreGen<T>> filtre_rif_ups_gpu(const Vecf &c, entier R)
{
  retourne std::make_shared<FiltrePolyphaseGpu<T>>(TSDGPU_POLY_UPS, c, R);
}
template<typename T> sptr<FiltreGen<T>> filtre_rif_demi_bande_gpu(const Vecf &c)
{
  retourne std::make_shared<FiltrePolyphaseGpu<T>>(TSDGPU_POLY_DEMI_BANDE, c, 2);
}
template<typename T> sptr<FiltreGen<T>> filtre_rif_decim_gpu(const Vecf &c, entier R)
{
  retourne std::make_shared<FiltrePolyphaseGpu<T>>(TSDGPU_POLY_DECIM, c, R);
}

// filtre_reechan<T>(ratio), T = float or cfloat (ra.cc:180-183,190-191): the reference's own planner (ra.cc:104-156) with
// every stage on the GPU
template<typename T> struct AdaptationRythmeArbitraireGpu: Filtre<T, T, float>
{
  std::vector<sptr<FiltreGen<T>>> etages;
  float ratio = 1;
  AdaptationRythmeArbitraireGpu(float r) { Configurable<float>::configure(r); }
  void configure_impl(const float &ratio_) override
  {
    ratio = ratio_;
    si((ratio <= 0) || std::isinf(ratio) || (ratio >= 1e9))
    {
      msg_erreur("filtre_reechan (gpu) : facteur de décimation invalide : {}.", ratio);
      ratio = 1;
    }
    etages.clear();
    float f = ratio;
    soit coefs = tsd::filtrage::design_rif_fen(15, "lp", 0.25, "hn");
    tantque(f < 0.5) { etages.push_back(filtre_rif_demi_bande_gpu<T>(coefs)); f *= 2; }
    std::vector<sptr<FiltreGen<T>>> ups;
    tantque(f >= 2) { ups.push_back(filtre_rif_ups_gpu<T>(coefs, 2)); f /= 2; }
    etages.insert(etages.end(), ups.begin(), ups.end());
    si(!(std::abs(f - 1) < 1e-6f))
      etages.push_back(filtre_itrp_gpu<T>(f, tsd::filtrage::itrp_sinc<T>({15, 256, std::min(0.4f, f / 2), "hn"})));
  }
  void step(const Vecteur<T> &x, Vecteur<T> &y) override
  {
    y = x;
    si(ratio == 1) retourne;
    pour(auto &e: etages) y = e->step(y);
  }
};
template<typename T = cfloat> sptr<Filtre<T, T, float>> filtre_reechan_gpu(float ratio)
{
  retourne std::make_shared<AdaptationRythmeArbitraireGpu<T>>(ratio);
}

// periodogramme_tfd(x, N) (fourier.hpp:967, fourier.cc:1451-1481): frames x N2/2 matrix of 10 log10(|X|^2 + 1e-20)
inline Tabf periodogramme_tfd_gpu(const Veccf &x, entier N)
{
  Vecf fen = tsd::filtrage::fenêtre("hn", N, non);                 // fourier.cc:796
  soit N2 = prochaine_puissance_de_2(N);
  std::vector<float> buf((size_t) 2 * (x.rows() / N) * (N2 / 2) + 1);
  int nf = 0, nb = 0;
  verifie(tsdgpu_periodogramme_tfd(x.data(), x.rows(), x.rows(), 1, N, fen.data(), buf.data(), (long long) buf.size(), &nf, &nb,
                                   TSDGPU_HOST),
          "periodogramme_tfd (gpu)");
  Tabf M(nf, nb);
  pour(auto i = 0; i < nf; i++)
    pour(auto j = 0; j < nb; j++)
      M(i, j) = buf[(size_t) i * nb + j];
  retourne M;
}

// rt_spectrum(config) (fourier.hpp:952; Spectrum, fourier.cc:1162-1343): same Filtre<cfloat,float,SpectrumConfig> interface,
// the windowing / transforms / fft-shifted accumulation / dB conversion on the device (tsdgpu_spectrum_*).  config.plan (an
// optional user FFT plan) has no role here.
struct SpectrumGpu : Filtre<cfloat, float, tsd::fourier::SpectrumConfig>
{
  tsdgpu_spectrum_t h = nullptr;
  int Ns = 0;
  ~SpectrumGpu() { si(h) tsdgpu_spectrum_destroy(h); }
  void configure_impl(const tsd::fourier::SpectrumConfig &c)
  {
    si(h) { tsdgpu_spectrum_destroy(h); h = nullptr; }
    soit Nf = c.Nf();
    Vecf f = tsd::filtrage::fenêtre(c.fenetre, Nf, non);
    f = sqrt(Nf / abs2(f).somme()) * f;                                  // fourier.cc:1203
    verifie(tsdgpu_spectrum_create(c.BS, c.nmeans, c.nsubs, c.sweep.active ? 1 : 0, c.sweep.step, c.sweep.masque_bf, c.sweep.masque_hf,
                                   f.data(), 1, &h),
            "rt_spectrum (gpu)");
    Ns = c.Ns();
  }
  void step(const Vecteur<cfloat> &x, Vecf &y)
  {
    y.resize(Ns);
    int n_out = 0;
    verifie(tsdgpu_spectrum_step(h, x.data(), x.rows(), x.rows(), y.data(), Ns, &n_out, TSDGPU_HOST), "Spectrum::step (gpu)");
    si(n_out != Ns) y.resize(n_out);
  }
};
inline sptr<Filtre<cfloat, float, tsd::fourier::SpectrumConfig>> rt_spectrum_gpu(const tsd::fourier::SpectrumConfig &config)
{
  soit res = std::make_shared<SpectrumGpu>();
  res->configure(config);
  retourne res;
}

} // namespace tsd::gpu
