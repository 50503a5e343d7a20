// Drop-in proof, second form: this program contains NO GPU-specific call.  It uses only the reference's public API
// (tsd::filtrage::filtrer, filtre_rif, rééchan, filtre_rif_fft, filtre_fft, fft) and is linked twice by oracle/Makefile:
//   dropin_check_cpu   reference objects only                      -> prints a checksum per call, on the CPU
//   dropin_check_gpu   + integration/tsd_gpu_dropin.cc + libtsdgpu -> same source, the factories now build GPU objects
// tests/test_gpu_parity.py::test_cpp_dropin_tu runs both and compares the printed samples (<= 1e-5 of RMS) and, through
// tsdgpu_launch_count(), that the second binary really launched kernels.
#include "tsd/tsd.hpp"
#include "tsd/filtrage.hpp"
#include "tsd/fourier.hpp"
#include <cstdio>
#include <random>

using namespace tsd;
using namespace tsd::filtrage;
using namespace tsd::fourier;

extern "C" long long tsdgpu_launch_count(int) __attribute__((weak));

static Veccf bruit(entier n, unsigned seed)
{
  std::mt19937 g(seed);
  std::normal_distribution<float> d(0, 1);
  Veccf x(n);
  pour(auto i = 0; i < n; i++) x(i) = cfloat(d(g), d(g));
  retourne x;
}
static void imprime(const char *nom, const Veccf &y)
{
  printf("%s %d", nom, y.rows());
  pour(auto i = 0; i < y.rows(); i += std::max(1, y.rows() / 16)) printf(" %.6e %.6e", y(i).real(), y(i).imag());
  double e = 0;
  pour(auto i = 0; i < y.rows(); i++) e += std::norm(y(i));
  printf(" rms %.6e\n", std::sqrt(e / std::max(1, y.rows())));
}

int main()
{
  get_logger() = [](const char *, entier, entier niveau, cstring s) { if(niveau >= 4) throw std::runtime_error(s); };
  try
  {
    soit x = bruit(30000, 21);
    soit h = design_rif_fen(127, "lp", 0.1);
    imprime("filtrer", filtrer(h, x));                                           // filtrage.hpp:1684-1711 -> filtre_rif<float,cfloat>
    soit f = filtre_rif<float, cfloat>(h);
    Veccf y1 = f->step(x.head(10000)), y2 = f->step(x.segment(10000, 20000));
    imprime("filtre_rif_blocs", vconcat(y1, y2));
    imprime("reechan_0.3", rééchan(x, 0.3f));                                    // tsd.hpp:700-705 -> filtre_reechan<cfloat>
    imprime("reechan_147_160", rééchan(x, 147.0f / 160.0f));
    imprime("filtre_itrp_lagrange", filtre_itrp<cfloat>(1.37f, itrp_lagrange<cfloat>(3))->step(x));
    imprime("filtre_rif_fft", filtre_rif_fft<cfloat>(h)->step(x));
    imprime("fft", fft(x.head(16384)));
    FiltreFFTConfig c;
    c.dim_blocs_temporel = 2048;
    c.nb_zeros_min = 2048;
    c.traitement_freq = [](Veccf &X) { pour(auto i = X.rows() / 4; i < X.rows() / 2; i++) X(i) = 0; };   // test-filtres.cc:428-432 style
    soit [ola, N] = filtre_fft(c);
    imprime("filtre_fft_rappel", ola->step(x));
    SpectrumConfig sc;                                                            // fourier.hpp:909-952 -> rt_spectrum
    sc.BS = 4096;
    sc.nmeans = 2;
    sc.nsubs = 2;
    soit sp = rt_spectrum(sc);
    Vecf s1, s2;
    sp->step(x.head(4096), s1);
    sp->step(x.segment(4096, 4096), s2);
    Veccf sdb(s2.rows());
    pour(auto i = 0; i < s2.rows(); i++) sdb(i) = cfloat(s2(i), (float) s1.rows());
    imprime("rt_spectrum_dB", sdb);
    printf("noyaux_gpu %lld\n", tsdgpu_launch_count ? tsdgpu_launch_count(0) : 0LL);
  }
  catch(const std::exception &e) { printf("exception: %s\n", e.what()); retourne 2; }
  catch(const std::string &s) { printf("exception: %s\n", s.c_str()); retourne 2; }
  retourne 0;
}
