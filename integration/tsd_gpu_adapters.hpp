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
//   filtre_itrp<cfloat>(r, itrp)   core/src/reechan/ra.cc:185-188           -> tsd::gpu::filtre_itrp_gpu(r, itrp)
//   fftplan_defaut (global hook)   core/src/fourier/fourier.cc:469-472      -> tsd::gpu::installe_fftplan_gpu()
#pragma once
#include "tsd/tsd.hpp"
#include "tsd/filtrage.hpp"
#include "tsd/fourier.hpp"
#include "tsdgpu.h"

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
// every fft()/ifft()/rfft()/Spectrum in the library then uses the GPU plan (power-of-two sizes)
inline void installe_fftplan_gpu()
{
  tsd::fourier::fftplan_defaut = []() -> sptr<tsd::fourier::FFTPlan> { retourne std::make_shared<FFTPlanGpu>(); };
}

// filtre_fft(config) with the callback of FiltreFFTRIF, "X *= H", passed as data (fourier.cc:956-959).
// K > 0 declares H = fft([0^(N-K), h]) * sqrt(N) (fourier.cc:962-965): overlap-save form.
struct OLAGpu: Filtre<cfloat, cfloat, tsd::fourier::FiltreFFTConfig>
{
  tsdgpu_ola_t h = nullptr;
  Veccf H;
  entier K = 0, N = 0;
  OLAGpu(const Veccf &H_, entier K_): H(H_), K(K_) {}
  ~OLAGpu() { tsdgpu_ola_destroy(h); }
  void configure_impl(const tsd::fourier::FiltreFFTConfig &c) override
  {
    si(c.avec_fenetrage) échec("filtre_fft (gpu) : mode fenêtré non disponible");
    tsdgpu_ola_destroy(h);
    verifie(tsdgpu_ola_create(c.dim_blocs_temporel, c.nb_zeros_min, H.rows() ? (const float *) H.data() : nullptr, K, 1, &h),
            "filtre_fft (gpu)");
    tsdgpu_ola_dims(h, nullptr, &N, nullptr, nullptr);
  }
  void step(const Veccf &x, Veccf &y) override
  {
    y.resize((entier) tsdgpu_ola_out_count(h, x.rows()));
    long long n_out = 0;
    verifie(tsdgpu_ola_step(h, x.data(), x.rows(), x.rows(), y.data(), std::max(1, y.rows()), &n_out, TSDGPU_HOST),
            "filtre_fft::step (gpu)");
  }
};
inline std::tuple<sptr<Filtre<cfloat, cfloat, tsd::fourier::FiltreFFTConfig>>, entier>
filtre_fft_gpu(const tsd::fourier::FiltreFFTConfig &config, const Veccf &H, entier K = 0)
{
  soit res = std::make_shared<OLAGpu>(H, K);
  res->configure(config);
  retourne {res, res->N};
}

// filtre_itrp<cfloat>(ratio, itrp): the interpolator's LUT is read through its public coefs(tau)
struct AdaptationRythmeGpu: FiltreGen<cfloat>
{
  tsdgpu_resamp_t h = nullptr;
  AdaptationRythmeGpu(float ratio, sptr<tsd::filtrage::InterpolateurRIF<cfloat>> itrp, entier nphases)
  {
    soit K = itrp->K;
    Vecf lut(K * (nphases + 1));
    pour(auto p = 0; p <= nphases; p++)
    {
      // tau inside LUT cell p so that (int)(tau*nphases) == p (itrp.cc:16-22)
      soit c = itrp->coefs(p == nphases ? 1.0f : (p + 0.5f) / nphases);
      pour(auto i = 0; i < K; i++) lut(p * K + i) = c(i);
    }
    verifie(tsdgpu_resamp_create(ratio, lut.data(), K, nphases, 1, &h), "filtre_itrp (gpu)");
  }
  ~AdaptationRythmeGpu() { tsdgpu_resamp_destroy(h); }
  void step(const Veccf &x, Veccf &y) override
  {
    soit n = x.rows();
    y.resize((entier) tsdgpu_resamp_out_count(h, n));
    si(n == 0) retourne;                                  // ra.cc:45-49
    long long n_out = 0;
    verifie(tsdgpu_resamp_step(h, x.data(), n, n, y.data(), std::max(1, y.rows()), y.rows(), &n_out, TSDGPU_HOST),
            "filtre_itrp::step (gpu)");
  }
};
inline sptr<FiltreGen<cfloat>> filtre_itrp_gpu(float ratio, sptr<tsd::filtrage::InterpolateurRIF<cfloat>> itrp, entier nphases = 256)
{
  retourne std::make_shared<AdaptationRythmeGpu>(ratio, itrp, nphases);
}

} // namespace tsd::gpu
