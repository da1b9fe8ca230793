/* fftw3.h SHIM -- test infrastructure only (oracle/), never on the product path.
 *
 * The reference's FFT is the external FFTW3 library (configure.ac:141,
 * call sites apf/apf/convolver.h:146,176-181,421-423,612 and
 * apf/apf/fftwtools.h:44-72). FFTW3 is not installed in this image and there is
 * no network, so this header provides exactly the subset the reference binds:
 * fftw{f,,l}_{malloc,free,plan_r2r_1d,execute,execute_r2r,destroy_plan},
 * fftw_r2r_kind, FFTW_R2HC / FFTW_HC2R / FFTW_PATIENT.
 *
 * Semantics follow FFTW's published halfcomplex convention:
 *   R2HC: out = [r0, r1, ..., r_{n/2}, i_{(n+1)/2-1}, ..., i_1]
 *   HC2R: unnormalised inverse (result is n times the input of R2HC).
 *
 * The float path (the one the convolver uses) is a float32 radix-4 Stockham
 * complex FFT of size n/2 plus a real split pass, for power-of-two n; it is
 * what every CPU-baseline number in this repo is measured with ("shim FFT:
 * hand-written float32 radix-4 Stockham").  Non-power-of-two sizes and the
 * double / long double flavours fall back to an O(n^2) DFT (correctness only).
 */
#ifndef SSR_B200_ORACLE_FFTW3_SHIM_H
#define SSR_B200_ORACLE_FFTW3_SHIM_H

#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <vector>

enum fftw_r2r_kind_do_not_use_me { FFTW_R2HC = 0, FFTW_HC2R = 1 };
typedef enum fftw_r2r_kind_do_not_use_me fftw_r2r_kind;

#define FFTW_MEASURE (0U)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)

namespace ssr_b200_fftshim
{

template<typename T>
struct Plan
{
  int n;
  fftw_r2r_kind kind;
  T* in;
  T* out;
  bool pow2;
  int m;  // n / 2
  // per-stage twiddles for the complex FFT of size m (split re / im)
  struct Stage { int radix; int Ns; std::vector<T> wr, wi; };
  std::vector<Stage> stages;
  std::vector<T> sr, si;  // split twiddles e^{-2 pi i k / n}, k = 0..m/2
  // scratch (a plan is executed by one thread at a time, as with FFTW's
  // new-array execute on distinct buffers the arrays differ but the scratch
  // does not; therefore scratch is allocated per call on the stack/heap below)
};

template<typename T>
static Plan<T>* make_plan(int n, T* in, T* out, fftw_r2r_kind kind)
{
  Plan<T>* p = new Plan<T>();
  p->n = n; p->kind = kind; p->in = in; p->out = out;
  p->pow2 = n >= 4 && (n & (n - 1)) == 0;
  p->m = n / 2;
  if (!p->pow2) return p;
  const int m = p->m;
  const double sign = (kind == FFTW_R2HC) ? -1.0 : 1.0;
  int Ns = 1;
  int rem = m;
  while (rem > 1)
  {
    int radix = (rem % 4 == 0) ? 4 : 2;
    typename Plan<T>::Stage st;
    st.radix = radix; st.Ns = Ns;
    st.wr.resize(size_t(radix - 1) * Ns);
    st.wi.resize(size_t(radix - 1) * Ns);
    for (int r = 1; r < radix; ++r)
    {
      for (int k = 0; k < Ns; ++k)
      {
        double a = sign * 2.0 * M_PI * double(r) * double(k) / double(Ns * radix);
        st.wr[size_t(r - 1) * Ns + k] = T(std::cos(a));
        st.wi[size_t(r - 1) * Ns + k] = T(std::sin(a));
      }
    }
    p->stages.push_back(st);
    Ns *= radix;
    rem /= radix;
  }
  p->sr.resize(m / 2 + 1);
  p->si.resize(m / 2 + 1);
  for (int k = 0; k <= m / 2; ++k)
  {
    double a = sign * 2.0 * M_PI * double(k) / double(n);
    p->sr[k] = T(std::cos(a));
    p->si[k] = T(std::sin(a));
  }
  return p;
}

// Complex FFT of size m on split arrays (ar, ai) using (br, bi) as ping-pong.
// Returns true if the result ended up in (br, bi).
template<typename T>
static bool cfft(const Plan<T>& p, T* ar, T* ai, T* br, T* bi, bool inverse)
{
  const int m = p.m;
  T* xr = ar; T* xi = ai; T* yr = br; T* yi = bi;
  const T sgn = inverse ? T(-1) : T(1);  // multiplies the "-i" rotation
  for (const auto& st: p.stages)
  {
    const int Ns = st.Ns;
    if (st.radix == 4)
    {
      const int q = m / 4;
      const int nb = q / Ns;
      const T* w1r = &st.wr[0];        const T* w1i = &st.wi[0];
      const T* w2r = &st.wr[Ns];       const T* w2i = &st.wi[Ns];
      const T* w3r = &st.wr[2 * Ns];   const T* w3i = &st.wi[2 * Ns];
      for (int b = 0; b < nb; ++b)
      {
        const T* i0r = xr + b * Ns;  const T* i0i = xi + b * Ns;
        const T* i1r = i0r + q;      const T* i1i = i0i + q;
        const T* i2r = i1r + q;      const T* i2i = i1i + q;
        const T* i3r = i2r + q;      const T* i3i = i2i + q;
        T* o0r = yr + b * Ns * 4;    T* o0i = yi + b * Ns * 4;
        T* o1r = o0r + Ns;           T* o1i = o0i + Ns;
        T* o2r = o1r + Ns;           T* o2i = o1i + Ns;
        T* o3r = o2r + Ns;           T* o3i = o2i + Ns;
        for (int k = 0; k < Ns; ++k)
        {
          T a0r = i0r[k], a0i = i0i[k];
          T a1r = i1r[k] * w1r[k] - i1i[k] * w1i[k];
          T a1i = i1r[k] * w1i[k] + i1i[k] * w1r[k];
          T a2r = i2r[k] * w2r[k] - i2i[k] * w2i[k];
          T a2i = i2r[k] * w2i[k] + i2i[k] * w2r[k];
          T a3r = i3r[k] * w3r[k] - i3i[k] * w3i[k];
          T a3i = i3r[k] * w3i[k] + i3i[k] * w3r[k];
          T s02r = a0r + a2r, s02i = a0i + a2i;
          T d02r = a0r - a2r, d02i = a0i - a2i;
          T s13r = a1r + a3r, s13i = a1i + a3i;
          T d13r = a1r - a3r, d13i = a1i - a3i;
          // forward: multiply d13 by -i  -> ( d13i, -d13r); inverse: by +i
          T rr = sgn * d13i, ri = -sgn * d13r;
          o0r[k] = s02r + s13r;  o0i[k] = s02i + s13i;
          o1r[k] = d02r + rr;    o1i[k] = d02i + ri;
          o2r[k] = s02r - s13r;  o2i[k] = s02i - s13i;
          o3r[k] = d02r - rr;    o3i[k] = d02i - ri;
        }
      }
    }
    else
    {
      const int q = m / 2;
      const int nb = q / Ns;
      const T* w1r = &st.wr[0]; const T* w1i = &st.wi[0];
      for (int b = 0; b < nb; ++b)
      {
        const T* i0r = xr + b * Ns;  const T* i0i = xi + b * Ns;
        const T* i1r = i0r + q;      const T* i1i = i0i + q;
        T* o0r = yr + b * Ns * 2;    T* o0i = yi + b * Ns * 2;
        T* o1r = o0r + Ns;           T* o1i = o0i + Ns;
        for (int k = 0; k < Ns; ++k)
        {
          T a1r = i1r[k] * w1r[k] - i1i[k] * w1i[k];
          T a1i = i1r[k] * w1i[k] + i1i[k] * w1r[k];
          o0r[k] = i0r[k] + a1r;  o0i[k] = i0i[k] + a1i;
          o1r[k] = i0r[k] - a1r;  o1i[k] = i0i[k] - a1i;
        }
      }
    }
    T* t;
    t = xr; xr = yr; yr = t;
    t = xi; xi = yi; yi = t;
  }
  return xr == br;
}

template<typename T>
static void slow_dft(int n, fftw_r2r_kind kind, const T* in, T* out)
{
  std::vector<long double> tmp(n);
  if (kind == FFTW_R2HC)
  {
    for (int k = 0; k <= n / 2; ++k)
    {
      long double re = 0, im = 0;
      for (int j = 0; j < n; ++j)
      {
        long double a = -2.0L * M_PIl * (long double)((long long)k * j % n) / n;
        re += in[j] * std::cos(a); im += in[j] * std::sin(a);
      }
      tmp[k] = re;
      if (k > 0 && k < (n + 1) / 2) tmp[n - k] = im;
    }
  }
  else
  {
    for (int j = 0; j < n; ++j)
    {
      long double acc = in[0];
      for (int k = 1; k < (n + 1) / 2; ++k)
      {
        long double a = 2.0L * M_PIl * (long double)((long long)k * j % n) / n;
        acc += 2.0L * (in[k] * std::cos(a) - in[n - k] * std::sin(a));
      }
      if (n % 2 == 0) acc += in[n / 2] * ((j % 2) ? -1.0L : 1.0L);
      tmp[j] = acc;
    }
  }
  for (int j = 0; j < n; ++j) out[j] = T(tmp[j]);
}

template<typename T>
static void execute(const Plan<T>* p, T* in, T* out)
{
  const int n = p->n;
  if (!p->pow2) { slow_dft(n, p->kind, in, out); return; }
  const int m = p->m;
  // scratch: 4 arrays of m
  T stackbuf[4 * 2048];
  std::vector<T> heapbuf;
  T* buf = stackbuf;
  if (m > 2048) { heapbuf.resize(size_t(4) * m); buf = heapbuf.data(); }
  T* ar = buf; T* ai = buf + m; T* br = buf + 2 * m; T* bi = buf + 3 * m;

  if (p->kind == FFTW_R2HC)
  {
    for (int j = 0; j < m; ++j) { ar[j] = in[2 * j]; ai[j] = in[2 * j + 1]; }
    bool inb = cfft(*p, ar, ai, br, bi, false);
    const T* zr = inb ? br : ar; const T* zi = inb ? bi : ai;
    // split: X[k] = E[k] + w^k O[k], E = (Z[k] + conj Z[m-k]) / 2,
    //        O = (Z[k] - conj Z[m-k]) / (2i),  w = e^{-2 pi i / n}
    out[0] = zr[0] + zi[0];
    out[m] = zr[0] - zi[0];
    for (int k = 1; k <= m / 2; ++k)
    {
      const int k2 = m - k;
      T er = T(0.5) * (zr[k] + zr[k2]), ei = T(0.5) * (zi[k] - zi[k2]);
      T orr = T(0.5) * (zi[k] + zi[k2]), oi = T(-0.5) * (zr[k] - zr[k2]);
      T wr = p->sr[k], wi = p->si[k];
      T tr = orr * wr - oi * wi, ti = orr * wi + oi * wr;
      T xr = er + tr, xi = ei + ti;      // X[k]
      T yr = er - tr, yi = -(ei - ti);   // X[m-k] = conj(E[k] - w^k O[k])
      out[k] = xr;
      if (k != k2) { out[k2] = yr; out[n - k2] = yi; }
      out[n - k] = xi;
    }
  }
  else
  {
    // Z'[k] = (X[k] + conj X[m-k]) + i w^{-k} (X[k] - conj X[m-k])
    const T x0 = in[0], xm = in[m];
    ar[0] = x0 + xm; ai[0] = x0 - xm;
    for (int k = 1; k <= m / 2; ++k)
    {
      const int k2 = m - k;
      T xr = in[k], xi = in[n - k];
      T yr = in[k2], yi = (k2 == k) ? xi : in[n - k2];
      T er = xr + yr, ei = xi - yi;
      T dr = xr - yr, di = xi + yi;
      T wr = p->sr[k], wi = p->si[k];  // e^{+2 pi i k / n} (inverse plan)
      // i * w * d
      T tr = -(dr * wi + di * wr), ti = dr * wr - di * wi;
      ar[k] = er + tr; ai[k] = ei + ti;
      if (k != k2) { ar[k2] = er - tr; ai[k2] = -(ei - ti); }
    }
    bool inb = cfft(*p, ar, ai, br, bi, true);
    const T* zr = inb ? br : ar; const T* zi = inb ? bi : ai;
    for (int j = 0; j < m; ++j) { out[2 * j] = zr[j]; out[2 * j + 1] = zi[j]; }
  }
}

static inline void* aligned_malloc(size_t n)
{
  void* p = nullptr;
  if (n == 0) n = 64;
  if (posix_memalign(&p, 64, n) != 0) return nullptr;
  return p;
}

}  // namespace ssr_b200_fftshim

#define SSR_B200_FFTW_SHIM_FLAVOUR(T, PFX) \
  typedef ssr_b200_fftshim::Plan<T>* PFX##plan; \
  static inline void* PFX##malloc(size_t n) { \
    return ssr_b200_fftshim::aligned_malloc(n); } \
  static inline void PFX##free(void* p) { std::free(p); } \
  static inline PFX##plan PFX##plan_r2r_1d(int n, T* in, T* out \
      , fftw_r2r_kind kind, unsigned) { \
    return ssr_b200_fftshim::make_plan<T>(n, in, out, kind); } \
  static inline void PFX##execute(const PFX##plan p) { \
    ssr_b200_fftshim::execute<T>(p, p->in, p->out); } \
  static inline void PFX##execute_r2r(const PFX##plan p, T* in, T* out) { \
    ssr_b200_fftshim::execute<T>(p, in, out); } \
  static inline void PFX##destroy_plan(PFX##plan p) { delete p; }

SSR_B200_FFTW_SHIM_FLAVOUR(float, fftwf_)
SSR_B200_FFTW_SHIM_FLAVOUR(double, fftw_)
SSR_B200_FFTW_SHIM_FLAVOUR(long double, fftwl_)

#undef SSR_B200_FFTW_SHIM_FLAVOUR

#endif
