// tsd_gpu_dropin.cc — the reference's OWN factory symbols, defined on top of the GPU adapters.
//
// The reference defines its streaming-block factories as function templates with explicit instantiations
// (filtre_rif<Tc,T>: filtre-rt.cc:171-175,816-818; filtre_reechan<T>, filtre_itrp<T>: ra.cc:180-195; filtre_rif_fft<T>:
// fourier.cc:980-990).  Instantiations have vague linkage (weak symbols); the explicit SPECIALISATIONS below are ordinary
// strong definitions of the very same symbols, so an application that links this translation unit together with libtsd
// gets them instead of the CPU ones — with no change to any caller:
//   tsd::filtrage::filtrer / filtfilt / convol            (filtrage.hpp:1684-1780 -> filtre_rif<float,T>)
//   dsp::filter_fir / dsp::filter / dsp::resample / dsp::filter_resample / dsp::filter_itrp / dsp::filter_fir_fft
//                                                         (dsp/filter.hpp:1333-1381,1662-1666,1897-1913, dsp/dsp.hpp:499-503)
//   tsd::rééchan                                          (tsd.hpp:700-705 -> filtre_reechan<T>)
// fft() / ifft() / rfft() / Spectrum go through the fftplan_defaut hook, installed by the static initialiser below.
// filtre_fft(config) (fourier.cc:935-940) and rt_spectrum(config) (:1339) are plain functions, hence strong symbols in libtsd itself: routing them needs the
// one-line `__attribute__((weak))` on the reference definition (or `objcopy --weaken-symbol`, which is what the check
// build in oracle/Makefile does on its private copy of fourier.o); -DTSD_GPU_DROPIN_FILTRE_FFT then defines it here.
#include "tsd_gpu_adapters.hpp"

namespace tsd::filtrage {

template<> sptr<FiltreGen<float>> filtre_rif<float, float>(const Vecf &c) { retourne tsd::gpu::filtre_rif_gpu<float, float>(c); }
template<> sptr<FiltreGen<cfloat>> filtre_rif<float, cfloat>(const Vecf &c) { retourne tsd::gpu::filtre_rif_gpu<float, cfloat>(c); }
template<> sptr<FiltreGen<cfloat>> filtre_rif<cfloat, cfloat>(const Veccf &c) { retourne tsd::gpu::filtre_rif_gpu<cfloat, cfloat>(c); }

template<> sptr<Filtre<float, float, float>> filtre_reechan<float>(float ratio) { retourne tsd::gpu::filtre_reechan_gpu<float>(ratio); }
template<> sptr<Filtre<cfloat, cfloat, float>> filtre_reechan<cfloat>(float ratio) { retourne tsd::gpu::filtre_reechan_gpu<cfloat>(ratio); }

// the reference takes the abstract Interpolateur<T>; every interpolator it ships is an InterpolateurRIF<T> (itrp.cc:130-157)
template<typename T> static sptr<FiltreGen<T>> itrp_gpu(float ratio, sptr<Interpolateur<T>> itrp)
{
  soit rif = std::dynamic_pointer_cast<InterpolateurRIF<T>>(itrp);
  si(!rif) échec("filtre_itrp (gpu) : interpolateur \"{}\" sans coefs() : non pris en charge.", itrp ? itrp->nom : std::string("?"));
  retourne tsd::gpu::filtre_itrp_gpu<T>(ratio, rif);
}
template<> sptr<FiltreGen<float>> filtre_itrp<float>(float ratio, sptr<Interpolateur<float>> itrp) { retourne itrp_gpu<float>(ratio, itrp); }
template<> sptr<FiltreGen<cfloat>> filtre_itrp<cfloat>(float ratio, sptr<Interpolateur<cfloat>> itrp) { retourne itrp_gpu<cfloat>(ratio, itrp); }

template<> sptr<FiltreGen<float>> filtre_rif_fft<float>(const Vecf &h) { retourne tsd::gpu::filtre_rif_fft_gpu<float>(h); }
template<> sptr<FiltreGen<cfloat>> filtre_rif_fft<cfloat>(const Vecf &h) { retourne tsd::gpu::filtre_rif_fft_gpu<cfloat>(h); }

} // namespace tsd::filtrage

#ifdef TSD_GPU_DROPIN_FILTRE_FFT
namespace tsd::fourier {
std::tuple<sptr<Filtre<cfloat, cfloat, FiltreFFTConfig>>, entier> filtre_fft(const FiltreFFTConfig &config)
{
  retourne tsd::gpu::filtre_fft_gpu(config);
}
// rt_spectrum(config) (fourier.cc:1339-1344): a plain function as well, same treatment
sptr<Filtre<cfloat, float, SpectrumConfig>> rt_spectrum(const SpectrumConfig &config) { retourne tsd::gpu::rt_spectrum_gpu(config); }
}
#endif

namespace {
struct InstalleurPlan
{
  InstalleurPlan() { tsd::gpu::installe_fftplan_gpu(); }
} installeur_plan;
}
