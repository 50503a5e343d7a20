// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Eigen-free stand-in for the reference's core/src/tableau.cc so that the
// reference's OWN hot-path sources (filtre-rt.cc, fourier.cc, ra.cc, itrp.cc,
// polyphase.cc, tsd.cc, rif-fen.cc, divers.cc, moniteur-cpu.cc) can be compiled
// unmodified, in place, from /root/reference and linked into oracle/_ref/.
// Eigen is not installed in this image; the reference reaches Eigen only through
// the element-wise members of tsd::Tab (tableau.cc:856-872,1243-1533,1648-1717),
// which are restated here as plain loops.  Only the members that the hot-path
// objects leave undefined are provided.
//
// Behaviour kept from the reference container (tableau.cc):
//   * storage is malloc'd, column-major, dims are int (143-185, 693-722)
//   * segment()/col() return non-owning aliases flagged est_reference() (500-530)
//   * operator= copies values when dims match, otherwise re-seats on a deep clone (1114-1133)
//   * copie() converts between real/complex when formats differ (1055-1112)
//   * element-wise operators compute in the dtype of *this (1243-1533)
#include "tsd/tsd.hpp"
#include "tsd/filtrage.hpp"
#include "tsd/moniteur-cpu.hpp"
#include "tsd/vue.hpp"
#include "tsd/filtrage/frat.hpp"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>

namespace tsd {

struct Tab::Impl
{
  Scalaire kind = ℝ;
  entier nbits = 32;
  bool alias = false;        // view on somebody else's storage
  void *vals = nullptr;
  NTenseurDim dims;
  sptr<Impl> parent;         // keeps the aliased storage alive

  ~Impl() { if(vals && !alias) free(vals); }
  entier esize() const { return nbits / 8; }
  entier count() const { return dims.total(); }
  void shape(entier n, entier m)
  {
    dims = (m > 1) ? NTenseurDim::dim2(n, m) : NTenseurDim::dim1(n);
    vals = (count() > 0) ? malloc((size_t) count() * esize()) : nullptr;
  }
};

// ---- NTenseurDim (tableau.cc:60-141) -------------------------------------
NTenseurDim NTenseurDim::dim1(entier n) { NTenseurDim d; d.dims = {n}; return d; }
NTenseurDim NTenseurDim::dim2(entier n, entier m) { NTenseurDim d; d.dims = {n, m}; return d; }
entier NTenseurDim::total() const
{
  if(dims.empty()) return 0;
  entier t = 1;
  for(auto v: dims) t *= v;
  return t;
}
bouléen NTenseurDim::operator ==(const NTenseurDim &o) const
{
  entier a = ndims(), b = o.ndims();
  if(a == 0 && b >= 1 && o[0] != 0) return false;
  if(b == 0 && a >= 1 && dims[0] != 0) return false;
  for(entier i = 0; i < std::max(a, b); i++)
  {
    entier u = (i < a) ? dims[i] : 1, v = (i < b) ? o[i] : 1;
    if(u != v) return false;
  }
  return true;
}
std::ostream &operator<<(std::ostream &s, const NTenseurDim &t)
{
  for(entier i = 0; i < t.ndims(); i++) s << t(i) << (i + 1 < t.ndims() ? "x" : "");
  return s;
}
std::ostream &operator<<(std::ostream &s, const Scalaire &k)
{
  static const char *names[] = {"R", "C", "N", "Z", "B"};
  return s << (((int) k < 5) ? names[(int) k] : "?");
}
std::ostream &operator<<(std::ostream &s, const Tab &t)
{
  return s << "Tab[" << t.rows() << "x" << t.cols() << "]";
}

// ---- dtype dispatch -------------------------------------------------------
template<typename F> static void with_type(const Tab &t, F &&f)
{
  auto k = t.tscalaire();
  auto nb = t.nbits();
  if(k == ℝ && nb == 32) f((float *) nullptr);
  else if(k == ℝ && nb == 64) f((double *) nullptr);
  else if(k == ℂ && nb == 64) f((cfloat *) nullptr);
  else if(k == ℂ && nb == 128) f((cdouble *) nullptr);
  else if(k == ℤ && nb == 32) f((int32_t *) nullptr);
  else if(k == B && nb == 8) f((char *) nullptr);
  else échec("tab_shim: unsupported scalar kind {} / {} bits", (int) k, nb);
}
template<typename T> struct is_cplx: std::false_type {};
template<typename T> struct is_cplx<std::complex<T>>: std::true_type {};

// ---- construction / shape -------------------------------------------------
Tab::Tab() { impl = std::make_shared<Impl>(); }
Tab::Tab(Scalaire s, entier reso, entier n, entier m)
{
  impl = std::make_shared<Impl>();
  impl->kind = s;
  impl->nbits = reso;
  impl->shape(n, m);
}
void *Tab::rawptr() { return impl->vals; }
const void *Tab::rawptr() const { return impl->vals; }
entier Tab::rows() const { return impl->dims.ndims() == 0 ? 0 : impl->dims(0); }
entier Tab::cols() const
{
  if(impl->dims.ndims() < 2) return rows() == 0 ? 0 : 1;
  return impl->dims(1);
}
entier Tab::nelems() const { return impl->count(); }
Scalaire Tab::tscalaire() const { return impl->kind; }
entier Tab::nbits() const { return impl->nbits; }
entier Tab::dim_scalaire() const { return impl->esize(); }
bouléen Tab::est_reference() const { return impl && impl->alias; }
bool Tab::est_de_même_dimensions(const Tab &t) const { return impl->dims == t.impl->dims; }

void Tab::resize(entier n, entier m)
{
  if(rows() == n && cols() == m) return;
  if(impl->vals && !impl->alias) free(impl->vals);
  impl->shape(n, m);
}
void Tab::setZero(entier n, entier m)
{
  if(n >= 0 || impl->dims.ndims() == 0) resize(n, m);
  if(nelems() > 0) memset(rawptr(), 0, (size_t) nelems() * impl->esize());
}
Tab Tab::zeros(Scalaire s, entier reso, entier n, entier m)
{
  Tab t(s, reso, n, m);
  t.setZero(n, m);
  return t;
}
Tab Tab::ones(Scalaire s, entier reso, entier n, entier m)
{
  Tab t(s, reso, n, m);
  with_type(t, [&]<typename T>(T *) {
    T *p = (T *) t.rawptr();
    for(entier i = 0; i < t.nelems(); i++) p[i] = (T) 1;
  });
  return t;
}
static Tab like(const Tab &x)
{
  Tab y;
  y.impl->kind = x.impl->kind;
  y.impl->nbits = x.impl->nbits;
  y.impl->dims = x.impl->dims;
  y.impl->vals = x.nelems() > 0 ? malloc((size_t) x.nelems() * x.impl->esize()) : nullptr;
  return y;
}
Tab Tab::clone() const
{
  Tab y = like(*this);
  if(nelems() > 0) memcpy(y.rawptr(), rawptr(), (size_t) nelems() * impl->esize());
  return y;
}
static Tab view(const Tab &src, entier first, entier n)
{
  Tab r;
  r.impl->alias = true;
  r.impl->kind = src.impl->kind;
  r.impl->nbits = src.impl->nbits;
  r.impl->dims = NTenseurDim::dim1(n);
  r.impl->vals = (char *) src.impl->vals + (size_t) first * src.impl->esize();
  r.impl->parent = src.impl;
  return r;
}
Tab Tab::map(Scalaire s, entier reso, entier n, entier m, void *data)
{
  // non-owning view on caller memory (tableau.cc:724-738)
  Tab r;
  r.impl->alias = true;
  r.impl->kind = s;
  r.impl->nbits = reso;
  r.impl->dims = (m > 1) ? NTenseurDim::dim2(n, m) : NTenseurDim::dim1(n);
  r.impl->vals = data;
  return r;
}
Tab Tab::segment(entier i0, entier n) const
{
  // same failure condition as the reference (tableau.cc:520)
  assertion_msg(i0 + n <= rows(), "Tab::segment({},{}): dépassement ({} éléments).", i0, n, rows());
  return view(*this, i0, n);
}
Tab Tab::col(entier num) const
{
  assertion_msg(num <= cols(), "Tab::col({}): dépassement ({} colonnes).", num, cols());
  return view(*this, num * rows(), rows());
}

// ---- copy / assignment ----------------------------------------------------
template<typename D, typename S> static D conv(const S &s)
{
  if constexpr(is_cplx<S>::value && !is_cplx<D>::value) return (D) std::real(s);
  else if constexpr(is_cplx<S>::value && is_cplx<D>::value) return D((typename D::value_type) s.real(), (typename D::value_type) s.imag());
  else if constexpr(is_cplx<D>::value) return D((typename D::value_type) s);
  else return (D) s;
}
void Tab::copie(const Tab &src)
{
  if(!est_de_même_dimensions(src))
    échec("Tab::copie : source et destination non compatibles.");
  entier n = nelems();
  if(impl->kind == src.impl->kind && impl->nbits == src.impl->nbits)
  {
    if(n > 0) memmove(impl->vals, src.impl->vals, (size_t) n * impl->esize());
    return;
  }
  with_type(*this, [&]<typename D>(D *) {
    with_type(src, [&]<typename S>(S *) {
      D *d = (D *) impl->vals;
      const S *s = (const S *) src.impl->vals;
      for(entier i = 0; i < n; i++) d[i] = conv<D, S>(s[i]);
    });
  });
}
Tab &Tab::operator =(const Tab &src)
{
  if(est_de_même_dimensions(src)) copie(src);
  else impl = src.clone().impl;
  return *this;
}

// ---- element-wise arithmetic (dtype of *this) -----------------------------
template<typename Op> static Tab binary(const Tab &a, const Tab &b, const char *name, bool check_dims, Op op)
{
  if(a.tscalaire() != b.tscalaire())
    échec("Tab::Opérateur {} : types incompatibles", name);
  if(check_dims && !a.est_de_même_dimensions(b))
    échec("Tab::Opérateur {} : dimensions incompatibles", name);
  Tab y = like(a);
  with_type(a, [&]<typename T>(T *) {
    const T *p = (const T *) a.rawptr(), *q = (const T *) b.rawptr();
    T *r = (T *) y.rawptr();
    for(entier i = 0; i < a.nelems(); i++) r[i] = op(p[i], q[i]);
  });
  return y;
}
Tab Tab::operator +(const Tab &t) const { return binary(*this, t, "+", true, [](auto a, auto b) { return (decltype(a)) (a + b); }); }
Tab Tab::operator_minus(const Tab &t) const { return binary(*this, t, "-", false, [](auto a, auto b) { return (decltype(a)) (a - b); }); }
Tab Tab::operator *(const Tab &t) const { return binary(*this, t, "*", true, [](auto a, auto b) { return (decltype(a)) (a * b); }); }
Tab Tab::operator /(const Tab &t) const { return binary(*this, t, "/", true, [](auto a, auto b) { return (decltype(a)) (a / b); }); }

template<typename Op> static void inplace(Tab &a, const Tab &b, Op op)
{
  with_type(a, [&]<typename T>(T *) {
    T *p = (T *) a.rawptr();
    const T *q = (const T *) b.rawptr();
    for(entier i = 0; i < a.nelems(); i++) op(p[i], q[i]);
  });
}
Tab &Tab::operator +=(const Tab &x) { inplace(*this, x, [](auto &a, auto b) { a = a + b; }); return *this; }
Tab &Tab::operator *=(const Tab &x) { inplace(*this, x, [](auto &a, auto b) { a = a * b; }); return *this; }
Tab &Tab::operator /=(const Tab &x) { inplace(*this, x, [](auto &a, auto b) { a = a / b; }); return *this; }

template<typename Op> static Tab scalar_op(const Tab &a, float x, Op op)
{
  Tab y = like(a);
  with_type(a, [&]<typename T>(T *) {
    const T *p = (const T *) a.rawptr();
    T *r = (T *) y.rawptr();
    T s = (T) x;
    for(entier i = 0; i < a.nelems(); i++) r[i] = op(p[i], s);
  });
  return y;
}
Tab Tab::operator *(const float &x) const { return scalar_op(*this, x, [](auto a, auto s) { return (decltype(a)) (a * s); }); }
Tab Tab::operator /(const float &x) const { return scalar_op(*this, x, [](auto a, auto s) { return (decltype(a)) (a / s); }); }
Tab Tab::operator +(const float &x) const { return scalar_op(*this, x, [](auto a, auto s) { return (decltype(a)) (a + s); }); }
Tab Tab::operator -(const float &x) const { return scalar_op(*this, x, [](auto a, auto s) { return (decltype(a)) (a - s); }); }
Tab &Tab::operator *=(const float &x)
{
  with_type(*this, [&]<typename T>(T *) { T *p = (T *) rawptr(); T s = (T) x; for(entier i = 0; i < nelems(); i++) p[i] = p[i] * s; });
  return *this;
}
Tab &Tab::operator /=(const float &x)
{
  // complex vectors divide by (T) x, i.e. a complex scalar (tableau.cc:1306-1315)
  with_type(*this, [&]<typename T>(T *) { T *p = (T *) rawptr(); T s = (T) x; for(entier i = 0; i < nelems(); i++) p[i] = p[i] / s; });
  return *this;
}
Tab &Tab::operator *=(const cfloat &x)
{
  with_type(*this, [&]<typename T>(T *) {
    if constexpr(is_cplx<T>::value) { T *p = (T *) rawptr(); T s = (T) x; for(entier i = 0; i < nelems(); i++) p[i] = p[i] * s; }
    else échec("Tab *= cfloat : tableau réel");
  });
  return *this;
}
Tab &Tab::operator /=(const cfloat &x)
{
  with_type(*this, [&]<typename T>(T *) {
    if constexpr(is_cplx<T>::value) { T *p = (T *) rawptr(); T s = (T) x; for(entier i = 0; i < nelems(); i++) p[i] = p[i] / s; }
    else échec("Tab /= cfloat : tableau réel");
  });
  return *this;
}
Tab Tab::operator *(const cfloat &x) const { Tab y = clone(); y *= x; return y; }
Tab Tab::operator /(const cfloat &x) const { Tab y = clone(); y /= x; return y; }
Tab Tab::operator-() const
{
  Tab y = like(*this);
  with_type(*this, [&]<typename T>(T *) {
    const T *p = (const T *) rawptr(); T *r = (T *) y.rawptr();
    for(entier i = 0; i < nelems(); i++) r[i] = -p[i];
  });
  return y;
}
Tab Tab::conjugate() const
{
  Tab y = like(*this);
  with_type(*this, [&]<typename T>(T *) {
    const T *p = (const T *) rawptr(); T *r = (T *) y.rawptr();
    for(entier i = 0; i < nelems(); i++)
    {
      if constexpr(is_cplx<T>::value) r[i] = std::conj(p[i]);
      else r[i] = p[i];
    }
  });
  return y;
}
Tab Tab::reverse() const
{
  Tab y = like(*this);
  with_type(*this, [&]<typename T>(T *) {
    const T *p = (const T *) rawptr(); T *r = (T *) y.rawptr();
    entier n = nelems();
    for(entier i = 0; i < n; i++) r[i] = p[n - 1 - i];
  });
  return y;
}
Tab Tab::transpose_int() const
{
  entier r = rows(), c = cols();
  Tab y(impl->kind, impl->nbits, c, r);
  if(c <= 1) y.impl->dims = NTenseurDim::dim2(1, r);
  with_type(*this, [&]<typename T>(T *) {
    const T *p = (const T *) rawptr(); T *q = (T *) y.rawptr();
    for(entier j = 0; j < c; j++)
      for(entier i = 0; i < r; i++)
        q[j + (size_t) i * c] = p[i + (size_t) j * r];
  });
  return y;
}
bouléen Tab::hasNaN() const
{
  bool res = false;
  with_type(*this, [&]<typename T>(T *) {
    const T *p = (const T *) rawptr();
    for(entier i = 0; i < nelems(); i++)
    {
      if constexpr(is_cplx<T>::value) res |= std::isnan(p[i].real()) || std::isnan(p[i].imag());
      else if constexpr(std::is_floating_point<T>::value) res |= std::isnan(p[i]);
    }
  });
  return res;
}
bouléen Tab::est_nul() const
{
  bool res = true;
  with_type(*this, [&]<typename T>(T *) {
    const T *p = (const T *) rawptr();
    for(entier i = 0; i < nelems(); i++) res &= (p[i] == (T) 0);
  });
  return res;
}
double Tab::maxCoeff(entier *oi, entier *oj) const
{
  double best = 0;
  entier bi = 0;
  with_type(*this, [&]<typename T>(T *) {
    if constexpr(!is_cplx<T>::value)
    {
      const T *p = (const T *) rawptr();
      for(entier i = 0; i < nelems(); i++) if(i == 0 || (double) p[i] > best) { best = (double) p[i]; bi = i; }
    }
  });
  if(oi) *oi = rows() ? bi % rows() : 0;
  if(oj) *oj = rows() ? bi / rows() : 0;
  return best;
}
double Tab::minCoeff(entier *oi, entier *oj) const
{
  double best = 0;
  entier bi = 0;
  with_type(*this, [&]<typename T>(T *) {
    if constexpr(!is_cplx<T>::value)
    {
      const T *p = (const T *) rawptr();
      for(entier i = 0; i < nelems(); i++) if(i == 0 || (double) p[i] < best) { best = (double) p[i]; bi = i; }
    }
  });
  if(oi) *oi = rows() ? bi % rows() : 0;
  if(oj) *oj = rows() ? bi / rows() : 0;
  return best;
}

// ---- element-wise math (tableau.cc:1648-1717) -----------------------------
template<typename Op> static Tab unary(const Tab &x, Op op)
{
  Tab y = like(x);
  with_type(x, [&]<typename T>(T *) {
    const T *p = (const T *) x.rawptr(); T *r = (T *) y.rawptr();
    for(entier i = 0; i < x.nelems(); i++) r[i] = op(p[i]);
  });
  return y;
}
template<typename T> static T f_cos(T v)
{
  if constexpr(std::is_integral<T>::value) return (T) std::cos((double) v); else return std::cos(v);
}
template<typename T> static T f_exp(T v)
{
  if constexpr(std::is_integral<T>::value) return (T) std::exp((double) v); else return std::exp(v);
}
template<typename T> static T f_log10(T v)
{
  if constexpr(std::is_integral<T>::value) return (T) std::log10((double) v); else return std::log10(v);
}
Tab cos_i(const Tab &x) { return unary(x, [](auto v) { return f_cos(v); }); }
Tab exp_i(const Tab &x) { return unary(x, [](auto v) { return f_exp(v); }); }
Tab log10_i(const Tab &x) { return unary(x, [](auto v) { return f_log10(v); }); }
Tab square_i(const Tab &x) { return unary(x, [](auto v) { return (decltype(v)) (v * v); }); }
template<typename T> static T f_sqrt(T v)
{
  if constexpr(std::is_integral<T>::value) return (T) std::sqrt((double) v); else return std::sqrt(v);
}
Tab sqrt_i(const Tab &x) { return unary(x, [](auto v) { return f_sqrt(v); }); }   // tableau.cc:1704
// tableau.cc:1172-1183: element > (T) x, result = boolean table (B, 8 bits)
Tab Tab::operator>(double x) const
{
  Tab y(B, 8, rows(), cols());
  with_type(*this, [&]<typename T>(T *) {
    if constexpr(is_cplx<T>::value) échec("tab_shim: operator>(double) on a complex table");
    else
    {
      const T *p = (const T *) rawptr();
      char *r = (char *) y.rawptr();
      for(entier i = 0; i < nelems(); i++) r[i] = p[i] > (T) x;
    }
  });
  return y;
}
template<typename Op> static Tab c2r(const Tab &x, Op op)
{
  Tab y;
  with_type(x, [&]<typename T>(T *) {
    const T *p = (const T *) x.rawptr();
    if constexpr(is_cplx<T>::value)
    {
      using R = typename T::value_type;
      y = Tab(ℝ, x.nbits() / 2, x.rows(), x.cols());
      R *r = (R *) y.rawptr();
      for(entier i = 0; i < x.nelems(); i++) r[i] = op(p[i]);
    }
    else
    {
      y = like(x);
      T *r = (T *) y.rawptr();
      for(entier i = 0; i < x.nelems(); i++) r[i] = (T) op(p[i]);
    }
  });
  return y;
}
Tab abs_i(const Tab &x) { return c2r(x, [](auto v) { return std::abs(v); }); }
Tab abs2_i(const Tab &x) { return c2r(x, [](auto v) { return std::norm(std::complex<double>(v)) ; }); }

// ---- stubs for symbols outside the hot path -------------------------------
template<> Vecteur<cfloat> Poly<float>::roots() const { échec("tab_shim: Poly::roots needs Eigen (not on the hot path)"); return {}; }
template<> Vecteur<cfloat> Poly<cfloat>::roots() const { échec("tab_shim: Poly::roots needs Eigen (not on the hot path)"); return {}; }
std::ostream &operator<<(std::ostream &os, const FRat<cfloat> &) { return os << "FRat<cfloat>"; }
std::ostream &operator<<(std::ostream &os, const FRat<float> &) { return os << "FRat<float>"; }

namespace vue {
Stdo stdo;
void Stdo::printf(cstring) {}
void Stdo::flush() {}
}

namespace filtrage {

// Normalised abscissa of the window, [-1/2, 1/2) (fenetres.cc:17-59).
static Vecf window_axis(entier n, bool sym)
{
  float tmin = -(n / 2), tmax;
  if((n & 1) == 0) tmax = sym ? n / 2 : (n - 1) / 2;
  else tmax = sym ? n / 2 : n / 2 - ((float) n - 1) / n;
  return linspace(tmin / n, tmax / n, n);
}
// "re" (none), "hn" (Hann), "hm" (Hamming) windows: a + (1-a) cos(2 pi t) (fenetres.cc:125-128,222-232).
Vecf fenêtre(Fenetre type, entier n, bouléen sym)
{
  Vecf x = Vecf::ones(n);
  if(type == Fenetre::AUCUNE) return x;
  float a;
  if(type == Fenetre::HANN) a = 0.5f;
  else if(type == Fenetre::HAMMING) a = 0.54f;
  else { échec("tab_shim: window type not available without Eigen"); return x; }
  Vecf t = window_axis(n, sym);
  // float arithmetic like the reference's Tab expression a + (1-a) * cos(2*π*t)
  float w = (float) (2 * π);
  for(entier i = 0; i < n; i++) x(i) = a + (1 - a) * std::cos(w * t(i));
  return x;
}
Vecf fenêtre(cstring nom, entier n, bouléen sym)
{
  // name table of fenetres.cc:178-200 restricted to the windows this shim can build
  if(nom == "" || nom == "aucune" || nom == "none" || nom == "re") return fenêtre(Fenetre::AUCUNE, n, sym);
  if(nom == "hn" || nom == "hann") return fenêtre(Fenetre::HANN, n, sym);
  if(nom == "hm" || nom == "hamming") return fenêtre(Fenetre::HAMMING, n, sym);
  échec("tab_shim: window '{}' not available without Eigen", nom);
  return {};
}
Vecf fenêtre_chebychev(entier, float, bouléen) { échec("tab_shim: chebychev window needs Eigen"); return {}; }
Vecf fenêtre_kaiser(float, float, bouléen) { échec("tab_shim: kaiser window not built"); return {}; }
float lexp_coef(Fréquence fc) { return 1.0f - std::exp(-fc.value * 2 * π_f); }
void verifie_frequence_normalisee(float f, cstring msg)
{
  if(f < 0 || f > 0.5) échec("{}: fréquence normalisée attendue, f = {}", msg, f);
}

} // namespace filtrage
} // namespace tsd
