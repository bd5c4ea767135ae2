// TEST INFRASTRUCTURE ONLY (see oracle/README.md) - never linked into the product library.
//
// Forward-mode automatic differentiation with nestable dual numbers.  The reference
// never writes a force by hand: gradient = jax.grad(potential)
// (/root/reference/streamsculptor/main.py:37-40), tidal tensor = jax.jacfwd/jacrev of
// that (main.py:59-65, fields.py:193), third derivatives = one more jacfwd
// (perturbative.py:281-296 through main.py:76-104).  The oracle restates exactly that:
// potentials are written ONCE as scalar formulas Phi(x,t) and every derivative comes
// from Dual<> nesting, so that the hand-derived closed forms in the CUDA kernels are
// checked against an independent differentiation of the reference's own formulas.
#ifndef ORC_AD_H
#define ORC_AD_H
#include <cmath>

namespace orc {

template <class T, int N>
struct Dual {
    T v;
    T d[N];
    Dual() : v(), d() {}
    Dual(double c) : v(c), d() {}                       // constant
    static Dual var(const T& val, int i) { Dual r; r.v = val; r.d[i] = T(1.0); return r; }
};

inline double val(double x) { return x; }
template <class T, int N> inline double val(const Dual<T, N>& x) { return val(x.v); }

// --- arithmetic: Dual (+-*/) Dual, Dual (+-*/) double, double (+-*/) Dual ---
template <class T, int N> inline Dual<T, N> operator-(const Dual<T, N>& a) {
    Dual<T, N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <class T, int N> inline Dual<T, N> operator+(const Dual<T, N>& a, const Dual<T, N>& b) {
    Dual<T, N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <class T, int N> inline Dual<T, N> operator-(const Dual<T, N>& a, const Dual<T, N>& b) {
    Dual<T, N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <class T, int N> inline Dual<T, N> operator*(const Dual<T, N>& a, const Dual<T, N>& b) {
    Dual<T, N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <class T, int N> inline Dual<T, N> operator/(const Dual<T, N>& a, const Dual<T, N>& b) {
    Dual<T, N> r; T inv = 1.0 / b.v; r.v = a.v * inv;
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r; }

template <class T, int N> inline Dual<T, N> operator+(const Dual<T, N>& a, double b) { Dual<T, N> r = a; r.v = a.v + b; return r; }
template <class T, int N> inline Dual<T, N> operator+(double b, const Dual<T, N>& a) { return a + b; }
template <class T, int N> inline Dual<T, N> operator-(const Dual<T, N>& a, double b) { Dual<T, N> r = a; r.v = a.v - b; return r; }
template <class T, int N> inline Dual<T, N> operator-(double b, const Dual<T, N>& a) { return (-a) + b; }
template <class T, int N> inline Dual<T, N> operator*(const Dual<T, N>& a, double b) {
    Dual<T, N> r; r.v = a.v * b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b; return r; }
template <class T, int N> inline Dual<T, N> operator*(double b, const Dual<T, N>& a) { return a * b; }
template <class T, int N> inline Dual<T, N> operator/(const Dual<T, N>& a, double b) {
    Dual<T, N> r; r.v = a.v / b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / b; return r; }
template <class T, int N> inline Dual<T, N> operator/(double b, const Dual<T, N>& a) {
    Dual<T, N> r; T inv = 1.0 / a.v; r.v = b * inv; T f = -(r.v * inv);
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * f;
    return r; }
template <class T, int N> inline Dual<T, N>& operator+=(Dual<T, N>& a, const Dual<T, N>& b) { a = a + b; return a; }

// --- elementary functions (ADL finds the right overload at every nesting level) ---
using std::sqrt; using std::log; using std::pow; using std::fabs;
template <class T, int N> inline Dual<T, N> sqrt(const Dual<T, N>& a) {
    Dual<T, N> r; r.v = sqrt(a.v); T f = 0.5 / r.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * f; return r; }
template <class T, int N> inline Dual<T, N> log(const Dual<T, N>& a) {
    Dual<T, N> r; r.v = log(a.v); T f = 1.0 / a.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * f; return r; }
inline double powc(double a, double p) { return std::pow(a, p); }
template <class T, int N> inline Dual<T, N> powc(const Dual<T, N>& a, double p) {   // a ** p, p constant
    Dual<T, N> r; r.v = powc(a.v, p); T f = p * powc(a.v, p - 1.0); for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * f; return r; }
template <class T> inline T sq(const T& a) { return a * a; }

}  // namespace orc
#endif
