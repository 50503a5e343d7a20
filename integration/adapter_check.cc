// Drop-in proof: drives the reference's OWN interfaces (FiltreGen<T>::step, FFTPlan, fft()) once through
// the reference CPU classes and once through the adapters of tsd_gpu_adapters.hpp, on the same inputs.
// Built here by oracle/Makefile target `adapter` against /root/reference headers + the Tab shim; the
// binary (oracle/_ref/adapter_check) travels to the GPU box.  Exit code 0 = all within 1e-5 of RMS.
#include "tsd_gpu_adapters.hpp"
#include <cstdio>
#include <random>
#include <thread>

using namespace tsd;
using namespace tsd::filtrage;
using namespace tsd::fourier;

static Veccf bruit(entier n, unsigned seed)
{
  std::mt19937 g(seed);
  std::normal_distribution<float> d(0, 1);
  Veccf x(n);
  pour(auto i = 0; i < n; i++) x(i) = cfloat(d(g), d(g));
  retourne x;
}
static Vecf bruit_reel(entier n, unsigned seed)
{
  std::mt19937 g(seed);
  std::normal_distribution<float> d(0, 1);
  Vecf x(n);
  pour(auto i = 0; i < n; i++) x(i) = d(g);
  retourne x;
}
static double ecart_reel(const Vecf &a, const Vecf &b)
{
  si(a.rows() != b.rows()) retourne 1e9;
  double m = 0;
  pour(auto i = 0; i < a.rows(); i++) m = std::max(m, (double) std::abs(a(i) - b(i)));
  retourne m;
}
static double ecart(const Veccf &a, const Veccf &b, const Veccf &x)
{
  si(a.rows() != b.rows()) retourne 1e9;
  double m = 0, e = 0;
  pour(auto i = 0; i < a.rows(); i++) m = std::max(m, (double) std::abs(a(i) - b(i)));
  pour(auto i = 0; i < x.rows(); i++) e += std::norm(x(i));
  retourne m / std::sqrt(e / std::max(1, x.rows()));
}

int main()
{
  get_logger() = [](const char *, entier, entier niveau, cstring s) { if(niveau >= 4) throw std::runtime_error(s); };
  int bad = 0;
  try
  {
    // filtre_rif<float,cfloat>, streaming in blocks (test-filtres.cc:9-31 pattern)
    soit h = design_rif_fen(127, "lp", 0.1);
    soit x = bruit(20000, 1);
    soit fc = filtre_rif<float, cfloat>(h);
    soit fg = tsd::gpu::filtre_rif_gpu<float, cfloat>(h);
    pour(auto i = 0; i < 20000; i += 5000)
    {
      Veccf xb = x.segment(i, 5000).clone();
      soit e = ecart(fg->step(xb), fc->step(xb), xb);
      printf("filtre_rif   bloc %5d : ecart %.2e\n", i, e);
      bad += e > 1e-5;
    }
    // fft()/ifft() through the global plan hook
    soit xf = bruit(65536, 2);
    soit Xc = fft(xf);
    tsd::gpu::installe_fftplan_gpu();
    soit Xg = fft(xf);
    soit e1 = ecart(Xg, Xc, xf), e2 = ecart(ifft(Xg), xf, xf);
    printf("fft 65536 via fftplan_defaut : ecart %.2e, aller-retour %.2e\n", e1, e2);
    bad += (e1 > 1e-5) + (e2 > 1e-5);
    // filtre_fft, K = 4095, Ne = 61441
    soit h4 = design_rif_fen(4095, "lp", 0.1);
    FiltreFFTConfig cfg;
    cfg.dim_blocs_temporel = 61441;
    cfg.nb_zeros_min = 4095;
    Veccf H;
    {
      soit h2 = Vecf::zeros(65536);
      h2.tail(4095) = h4;
      H = fft(h2);          // GPU plan is installed; rfft path of the reference on top of it
      H *= sqrt(65536.0f);
    }
    cfg.traitement_freq = [&](Veccf &X) { X *= H; };
    soit [oc, Nc] = filtre_fft(cfg);
    soit [og, Ng] = tsd::gpu::filtre_fft_gpu(cfg, H, 4095);
    soit xo = bruit(200000, 3);
    soit yc = oc->step(xo), yg = og->step(xo);
    soit e3 = ecart(yg, yc, xo);
    printf("filtre_fft   N %d/%d, %d/%d echantillons : ecart %.2e\n", Nc, Ng, yc.rows(), yg.rows(), e3);
    bad += (e3 > 1e-5) + (Nc != Ng);
    {
      // windowed (Hann, 50 % overlap) mode with the same callback: first block silent, then Ne samples per block
      FiltreFFTConfig cf2;
      cf2.dim_blocs_temporel = 4096;
      cf2.nb_zeros_min = 4096;
      cf2.avec_fenetrage = oui;
      Veccf H2 = bruit(8192, 7);
      cf2.traitement_freq = [&](Veccf &X) { X *= H2; };
      soit [wc, Nwc] = filtre_fft(cf2);
      soit [wg, Nwg] = tsd::gpu::filtre_fft_gpu(cf2, H2);
      soit xw = bruit(50000, 8);
      soit ywc = wc->step(xw), ywg = wg->step(xw);
      soit e3b = ywc.rows() == ywg.rows() ? ecart(ywg, ywc, ywc) : 1.0f;
      printf("filtre_fft   fenetre N %d/%d, %d/%d echantillons : ecart %.2e\n", Nwc, Nwg, ywc.rows(), ywg.rows(), e3b);
      bad += (e3b > 1e-5) + (Nwc != Nwg);
    }
    {
      // periodogramme_tfd: same matrix (dB) from the reference and from the GPU frames
      soit xs = bruit(20000, 9);
      Tabf Mc = tsd::tf::periodogramme_tfd(xs, 512), Mg = tsd::gpu::periodogramme_tfd_gpu(xs, 512);
      double e = (Mc.rows() == Mg.rows() && Mc.cols() == Mg.cols()) ? 0.0 : 1e9;
      pour(auto i = 0; i < Mc.rows() && e < 1e9; i++)
        pour(auto j = 0; j < Mc.cols(); j++)
          e = std::max(e, (double) std::abs(Mc(i, j) - Mg(i, j)));
      printf("periodogramme_tfd %d x %d : ecart max %.2e dB\n", Mg.rows(), Mg.cols(), e);
      bad += e > 1e-3;
    }
    {
      // rt_spectrum: sub-blocks + sweep + masks, two averages; empty results in between like the reference
      SpectrumConfig sc;
      sc.BS = 4096;
      sc.nmeans = 2;
      sc.nsubs = 4;
      sc.sweep.active = oui;
      sc.sweep.step = 512;
      sc.sweep.masque_bf = 8;
      sc.sweep.masque_hf = 16;
      soit sc_c = rt_spectrum(sc), sc_g = tsd::gpu::rt_spectrum_gpu(sc);
      double e = 0;
      int vides = 0;
      pour(auto b = 0; b < 4; b++)
      {
        soit xs = bruit(4096, 20 + b);
        Vecf yc, yg;
        sc_c->step(xs, yc);
        sc_g->step(xs, yg);
        si(yc.rows() != yg.rows()) e = 1e9;
        sinon si(yc.rows() == 0) vides++;
        sinon e = std::max(e, ecart_reel(yc, yg));
      }
      printf("rt_spectrum  BS 4096, 4 sous-blocs, balayage : %d resultats vides, ecart max %.2e dB\n", vides, e);
      bad += (e > 1e-3) + (vides != 2);
    }
    // filtre_itrp 147/160, sinc 64 x 257
    soit it = itrp_sinc<cfloat>({64, 256, 0.4f, "hn"});
    soit rc = filtre_itrp<cfloat>(147.0f / 160.0f, it);
    soit rg = tsd::gpu::filtre_itrp_gpu(147.0f / 160.0f, it, 256);
    soit xr = bruit(50000, 4);
    soit e4 = ecart(rg->step(xr), rc->step(xr), xr);
    printf("filtre_itrp  147/160 : ecart %.2e\n", e4);
    bad += e4 > 1e-5;
    // polyphase stages and full resample() chains (ratios outside [0.5, 2) use half-band / x2 stages)
    soit h15 = design_rif_fen(15, "lp", 0.25);
    soit xp = bruit(30001, 5);
    {
      soit e5 = ecart(tsd::gpu::filtre_rif_ups_gpu<cfloat>(h15, 2)->step(xp), filtre_rif_ups<float, cfloat>(h15, 2)->step(xp), xp);
      soit e6 = ecart(tsd::gpu::filtre_rif_demi_bande_gpu<cfloat>(h15)->step(xp), filtre_rif_demi_bande<float, cfloat>(h15)->step(xp), xp);
      soit e7 = ecart(tsd::gpu::filtre_rif_decim_gpu<cfloat>(h15, 3)->step(xp), filtre_rif_decim<float, cfloat>(h15, 3)->step(xp), xp);
      printf("polyphase    ups %.2e, demi-bande %.2e, decim %.2e\n", e5, e6, e7);
      bad += (e5 > 1e-5) + (e6 > 1e-5) + (e7 > 1e-5);
    }
    pour(float ratio: {0.1f, 7.3f, 147.0f / 160.0f})
    {
      soit e8 = ecart(tsd::gpu::filtre_reechan_gpu(ratio)->step(xp), filtre_reechan<cfloat>(ratio)->step(xp), xp);
      printf("filtre_reechan %.4f : ecart %.2e\n", ratio, e8);
      bad += e8 > 1e-5;
    }
    {
      // every interpolator of itrp.cc:130-157, T = float (the instantiation tests/test-ra.cc drives) and cfloat
      soit xf = bruit_reel(20000, 11);
      soit xc = bruit(20000, 12);
      pour(float ratio: {1.2f, 0.73f})
      {
        double er = 0, ec = 0;
        er = std::max(er, ecart_reel(tsd::gpu::filtre_itrp_gpu<float>(ratio, itrp_cspline<float>())->step(xf), filtre_itrp<float>(ratio, itrp_cspline<float>())->step(xf)));
        er = std::max(er, ecart_reel(tsd::gpu::filtre_itrp_gpu<float>(ratio, itrp_lineaire<float>())->step(xf), filtre_itrp<float>(ratio, itrp_lineaire<float>())->step(xf)));
        er = std::max(er, ecart_reel(tsd::gpu::filtre_itrp_gpu<float>(ratio, itrp_lagrange<float>(3))->step(xf), filtre_itrp<float>(ratio, itrp_lagrange<float>(3))->step(xf)));
        er = std::max(er, ecart_reel(tsd::gpu::filtre_itrp_gpu<float>(ratio, itrp_sinc<float>({31, 100, 0.4f, "hn"}))->step(xf), filtre_itrp<float>(ratio, itrp_sinc<float>({31, 100, 0.4f, "hn"}))->step(xf)));
        er = std::max(er, ecart_reel(tsd::gpu::filtre_reechan_gpu<float>(ratio)->step(xf), filtre_reechan<float>(ratio)->step(xf)));
        ec = std::max(ec, ecart(tsd::gpu::filtre_itrp_gpu<cfloat>(ratio, itrp_lineaire<cfloat>())->step(xc), filtre_itrp<cfloat>(ratio, itrp_lineaire<cfloat>())->step(xc), xc));
        ec = std::max(ec, ecart(tsd::gpu::filtre_itrp_gpu<cfloat>(ratio, itrp_lagrange<cfloat>(5))->step(xc), filtre_itrp<cfloat>(ratio, itrp_lagrange<cfloat>(5))->step(xc), xc));
        ec = std::max(ec, ecart(tsd::gpu::filtre_itrp_gpu<cfloat>(ratio, itrp_cspline<cfloat>())->step(xc), filtre_itrp<cfloat>(ratio, itrp_cspline<cfloat>())->step(xc), xc));
        printf("interpolateurs %.2f : float (cspline, lineaire, lagrange, sinc/100 phases, reechan) %.2e, cfloat %.2e\n", ratio, er, ec);
        bad += (er > 1e-5) + (ec > 1e-5);
      }
    }
    {
      // filtre_rif_fft<T>(h) with the reference's real() quirk (fourier.cc:976), both instantiations
      soit h127 = design_rif_fen(127, "lp", 0.2);
      soit xq = bruit(9000, 13);
      soit eq = ecart(tsd::gpu::filtre_rif_fft_gpu<cfloat>(h127)->step(xq), filtre_rif_fft<cfloat>(h127)->step(xq), xq);
      soit xqr = bruit_reel(9000, 14);
      soit eqr = ecart_reel(tsd::gpu::filtre_rif_fft_gpu<float>(h127)->step(xqr), filtre_rif_fft<float>(h127)->step(xqr));
      printf("filtre_rif_fft<cfloat> %.2e, <float> %.2e\n", eq, eqr);
      bad += (eq > 1e-5) + (eqr > 1e-5);
    }
    {
      // filtre_fft(config) with an ARBITRARY host callback (not a multiplication by a fixed vector): the callback is
      // called once per block on the host, like the reference does
      FiltreFFTConfig cg;
      cg.dim_blocs_temporel = 1000;
      cg.nb_zeros_min = 24;
      int appels_c = 0, appels_g = 0;
      cg.traitement_freq = [&](Veccf &X) { appels_c++; pour(auto i = 0; i < X.rows(); i++) X(i) = (i % 7 == 0) ? cfloat(0, 0) : X(i) * cfloat(0.5f, (float) (appels_c % 3)); };
      soit [gc, Ngc] = filtre_fft(cg);
      soit cg2 = cg;
      cg2.traitement_freq = [&](Veccf &X) { appels_g++; pour(auto i = 0; i < X.rows(); i++) X(i) = (i % 7 == 0) ? cfloat(0, 0) : X(i) * cfloat(0.5f, (float) (appels_g % 3)); };
      soit [gg, Ngg] = tsd::gpu::filtre_fft_gpu(cg2);
      soit xg = bruit(7777, 15);
      soit eg = ecart(gg->step(xg), gc->step(xg), xg);
      printf("filtre_fft rappel generique : %d/%d appels, ecart %.2e\n", appels_c, appels_g, eg);
      bad += (eg > 1e-5) + (appels_c != appels_g) + (Ngc != Ngg);
    }
    {
      // independent objects stepped from different threads (reference semantics): same results as sequentially
      soit xa = bruit(40000, 16), xb = bruit(40000, 17);
      soit ya_ref = filtre_rif<float, cfloat>(h)->step(xa);
      soit yb_ref = filtre_reechan<cfloat>(1.3f)->step(xb);
      Veccf ya, yb;
      std::string err;
      std::thread ta([&] { try { soit f = tsd::gpu::filtre_rif_gpu<float, cfloat>(h); pour(auto r = 0; r < 8; r++) { f = tsd::gpu::filtre_rif_gpu<float, cfloat>(h); ya = f->step(xa); } } catch(...) { err = "thread a"; } });
      std::thread tb([&] { try { pour(auto r = 0; r < 8; r++) yb = tsd::gpu::filtre_reechan_gpu(1.3f)->step(xb); } catch(...) { err = "thread b"; } });
      ta.join();
      tb.join();
      soit e_fils = std::max(ecart(ya, ya_ref, xa), ecart(yb, yb_ref, xb));
      printf("deux fils d'execution : ecart %.2e %s\n", e_fils, err.c_str());
      bad += (e_fils > 1e-5) + !err.empty();
    }
  }
  catch(const std::exception &e) { printf("exception: %s\n", e.what()); retourne 2; }
  catch(const std::string &s) { printf("exception: %s\n", s.c_str()); retourne 2; }
  printf(bad ? "ADAPTER CHECK FAILED\n" : "ADAPTER CHECK OK\n");
  retourne bad ? 1 : 0;
}
