// TEST INFRASTRUCTURE ONLY (oracle).  The reference's detector (core/src/fourier/detection.cc) draws debug figures when
// DetecteurConfig::debug_actif is set; the plotting library (src/vue, freetype / png / gtkmm) is out of scope and not
// built.  These empty definitions satisfy the linker; they are never called (debug_actif stays `non` in the oracle).
#include "tsd/tsd-all.hpp"

namespace tsd::vue {

sptr<const Rendable> Figure::rendable() const { return nullptr; }
Figure::Figure(cstring) {}
Figure::Courbe Figure::plot_int(const Vecf &, cstring, cstring) { return Courbe(); }
Figure::Courbe Figure::plot(float, float, cstring, cstring) { return Courbe(); }

sptr<const Rendable> Figures::rendable() const { return nullptr; }
Figures::Figures(entier, entier) {}
Figure Figures::subplot(entier) { return Figure(); }
void Figures::afficher(cstring, const Dim &) const {}

} // namespace tsd::vue
