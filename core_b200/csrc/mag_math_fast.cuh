// mag_math_fast.cuh -- MAG_FP_FAST device arithmetic: algebraically equivalent restatements of
// the reference formulas (see SURVEY.md section 8a "identities"), written for the fp64 pipe of
// sm_100a: FMA-contracted, one reciprocal and one square root per Gauss point, no vector
// normalisations.  Values agree with the strict path to ~1e-15 relative for the anisotropy
// ratios MeshAdapt meets (tests assert 1e-12); flags never depend on it because every value that
// lands within 1e-12 of a threshold is re-evaluated by the strict kernels.
//
// Every output depends on the transform Q only through M = Q Q^T and det Q > 0:
//   edge length at a Gauss point = sqrt(d^T M d),  d = (x1-x0)/2
//   AnisoSizeField: M = Rt diag(1/h^2) Rt^T, Rt = Gram-Schmidt of the interpolated frame, so with
//     c0, c1 the interpolated frame columns, n0 = |c0|^2, c1' = n0 c1 - (c0.c1) c0, n1 = |c1'|^2,
//     w = c0 x c1' (|w|^2 = n0 n1):
//     d^T M d = (d.c0)^2/(n0 h0^2) + (d.c1')^2/(n1 h1^2) + (d.w)^2/(n0 n1 h2^2)
#pragma once
#include "mag_math.cuh"

namespace magfa {

// shape values at the two Gauss points, as apfShape.cc:123-124 evaluates them
__device__ constexpr double kXI = 0.577350269189626;
__device__ constexpr double kNP0 = (1.0 - kXI) / 2.0, kNP1 = (1.0 + kXI) / 2.0;

// (contractions written out: the whole-part and the sub-range kernels must return the same bits, whatever nvcc would fuse)
__device__ __forceinline__ double edge_identity(const double ra[4], const double rb[4])
{
  double dx = rb[0] - ra[0], dy = rb[1] - ra[1], dz = rb[2] - ra[2];
  return sqrt(fma(dz, dz, fma(dy, dy, __dmul_rn(dx, dx))));
}

// iso: Q = I/h  ->  len = |d| (1/h+ + 1/h-) = |x1-x0| (h+ + h-) / (2 h+ h-)
__device__ __forceinline__ double edge_iso(const double ra[4], const double rb[4])
{
  double dx = rb[0] - ra[0], dy = rb[1] - ra[1], dz = rb[2] - ra[2];
  double l = sqrt(fma(dz, dz, fma(dy, dy, __dmul_rn(dx, dx))));
  double hp = fma(ra[3], kNP0, __dmul_rn(rb[3], kNP1));
  double hm = fma(ra[3], kNP1, __dmul_rn(rb[3], kNP0));
  return __dmul_rn(l, hp + hm) / __dmul_rn(__dmul_rn(2.0, hp), hm);
}

// Unnormalised Gram-Schmidt at one Gauss point, t = interpolation weight of vertex b (c = a + t (b - a)):
//   X = c0 x c1, N0 = |c0|^2, E = |X|^2, G = c0.c1, A0 = d.c0, A1 = d.c1, D = d.X
//   Gram-Schmidt frame: r0 = c0/sqrt(N0), r1 = (N0 c1 - G c0)/sqrt(N0^2 E ... ), r2 = X/sqrt(E), so
//   d^T M d = A0^2/(N0 h0^2) + (N0 A1 - G A0)^2/(N0 E h1^2) + D^2/(E h2^2)
// brought to the common denominator N0 E (h0 h1 h2)^2.  Returns numerator and denominator (both >= 0);
// the metric length is sqrt(num/den) = num * rsqrt(num * den).
// Every contraction is written out (fma / __dmul_rn / __dadd_rn): the same edge must get the same bits from every kernel
// that evaluates it (whole-part rows, lean rows, sub-range tiles), whatever nvcc would choose to fuse in each context.
__device__ __forceinline__ double dot3(double ax, double ay, double az, double bx, double by, double bz)
{
  return fma(az, bz, fma(ay, by, __dmul_rn(ax, bx)));
}
// a b - c d
__device__ __forceinline__ double diffprod(double a, double b, double c, double d) { return fma(a, b, -__dmul_rn(c, d)); }

// one Gauss point: c = interpolated {h0, h1, h2, c0x, c0y, c0z, c1x, c1y, c1z}, (dx, dy, dz) = x1 - x0
__device__ __forceinline__ void aniso_point_nd(const double* __restrict__ c, double dx, double dy, double dz, double& num, double& den)
{
  const double h0 = c[0], h1 = c[1], h2 = c[2];
  const double c0x = c[3], c0y = c[4], c0z = c[5], c1x = c[6], c1y = c[7], c1z = c[8];
  const double Xx = diffprod(c0y, c1z, c0z, c1y), Xy = diffprod(c0z, c1x, c0x, c1z), Xz = diffprod(c0x, c1y, c0y, c1x);
  const double N0 = dot3(c0x, c0y, c0z, c0x, c0y, c0z);
  const double E = dot3(Xx, Xy, Xz, Xx, Xy, Xz);
  const double G = dot3(c0x, c0y, c0z, c1x, c1y, c1z);
  const double A0 = dot3(dx, dy, dz, c0x, c0y, c0z);
  const double A1 = dot3(dx, dy, dz, c1x, c1y, c1z);
  const double D = dot3(dx, dy, dz, Xx, Xy, Xz);
  const double a1 = diffprod(N0, A1, G, A0);
  const double q0 = __dmul_rn(h1, h2), q1 = __dmul_rn(h0, h2), q2 = __dmul_rn(h0, h1), pp = __dmul_rn(q0, h0);
  const double t0 = __dmul_rn(A0, q0), t1 = __dmul_rn(a1, q1), t2 = __dmul_rn(D, q2);
  num = fma(__dmul_rn(t2, t2), N0, fma(t1, t1, __dmul_rn(__dmul_rn(t0, t0), E)));
  den = __dmul_rn(__dmul_rn(N0, E), __dmul_rn(pp, pp));
}

// 1 / sqrt(x) for a positive normal x: MUFU.RSQ64H (2^-22) + one cubic step, as CUDA's rsqrt() does it, without the
// special-case branch (the caller tests the range once for both Gauss points)
__device__ __forceinline__ double rsqrt_core(double x)
{
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, __dmul_rn(y, y), 1.0);          // 1 - x y^2
  const double p = fma(e, 0.375, 0.5);
  return fma(__dmul_rn(y, e), p, y);                       // y (1 + e/2 + 3 e^2/8)
}

// 0.5 (sqrt(np/dp) + sqrt(nm/dm)) with the square roots as n rsqrt(n d); exact zero for a zero-length edge
__device__ __forceinline__ double half_sum_sqrt_ratios(double np, double dp, double nm, double dm)
{
  const double xp = __dmul_rn(np, dp), xm = __dmul_rn(nm, dm);
  double sp, sm;
  // both products positive, normal and far from the ends of the exponent range (2^-900 .. 2^+900): an integer test of the
  // high words (negative values and NaN fail it through the sign bit / the top exponent)
  const unsigned hp = (unsigned)__double2hiint(xp) - 0x07b00000u, hm = (unsigned)__double2hiint(xm) - 0x07b00000u;
  if (hp < 0x70800000u && hm < 0x70800000u) {
    sp = __dmul_rn(np, rsqrt_core(xp));
    sm = __dmul_rn(nm, rsqrt_core(xm));
  } else {                                                 // zero-length edges, sizes beyond 1e+-23: rare
    sp = xp > 0.0 ? __dmul_rn(np, rsqrt(xp)) : 0.0;
    sm = xm > 0.0 ? __dmul_rn(nm, rsqrt(xm)) : 0.0;
  }
  return __dmul_rn(0.5, __dadd_rn(sp, sm));
}

// The records are consumed first -- both Gauss points' interpolated values c = a + t (b - a) -- so a and b are dead
// before the long part starts (the kernels keep other records in flight meanwhile: registers)
__device__ __forceinline__ double edge_aniso(const double* __restrict__ a, const double* __restrict__ b)
{
  const double dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2]; // 2d; the 1/2 is applied at the end
  double cp[9], cm[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const double dl = b[3 + i] - a[3 + i];
    cp[i] = fma(kNP1, dl, a[3 + i]);   // xi = +XI: weights (kNP0, kNP1)
    cm[i] = fma(kNP0, dl, a[3 + i]);   // xi = -XI: weights (kNP1, kNP0)
  }
  double np, dp, nm, dm;
  aniso_point_nd(cp, dx, dy, dz, np, dp);
  aniso_point_nd(cm, dx, dy, dz, nm, dm);
  return half_sum_sqrt_ratios(np, dp, nm, dm);
}

// log-Euclidean field.  The length at a Gauss point is sqrt(sum_k exp(lambda_k) (v_k . j)^2), j = (x1 - x0) / 2, with
// (lambda_k, v_k) the eigenpairs of the interpolated logM -- AS THE REFERENCE'S SOLVER LEAVES THEM: mth::eigenQR
// (mth/mthQR.cc:186-265) stops as soon as an off-diagonal drops below 1e-10, so its eigenvectors carry an error of that
// order and the reference's lengths differ from the exact matrix exponential by up to ~3e-11 relative (measured with a
// Jacobi solver on random frames).  Staying within 1e-12 of the reference therefore means running the same iteration --
// Householder reduction of column 0, then Wilkinson-shifted QR steps with the same shift, the same 1e-10 deflation
// tests and the same "column norm < 1e-10: no reflection" rule -- but nothing obliges us to run it the same way:
//   * every reflector acts on two rows only (the matrix is tridiagonal after the reduction), applied in the
//     beta-form H x = x - beta v (v . x), beta = 2 / |v|^2 = 1 / (n (n + |a|)): no normalisation divides;
//   * the eigenvector matrix is never formed: each reflection is applied to w = Q^T j instead;
//   * the last reflector of every QR step (a 1 x 1 sign flip) cancels in every quantity used.
// First form (round 1 until r1j): ~160 fp64 operations per iteration instead of ~350 + two dense 3 x 3 products.
// Second form (this one): the iterate stays a SYMMETRIC 3 x 3 matrix (six numbers d0 d1 d2 / e0 = (1,0), e1 = (2,1),
// f = (2,0); the reference's own iterate is symmetric up to rounding), and one QR step T' = R Q + mu = H1 H0 (T - mu) H0 H1
// + mu is applied as two two-sided 2 x 2 reflections: each touches one diagonal block (14 flops), one off-block pair and one
// pair of w.  The reflection that maps (a, b) to (-sign(a) n, 0) is H = [[-C, -S], [-S, C]] with (C, S) = sign(a) (a, b) / n,
// so one rsqrt replaces the square root and the divide of the beta-form.  The reflector of column 1 is taken from the
// LEFT product H0 (T - mu), as the reference forms R before it multiplies from the right.  ~105 fp64 operations per
// iteration against ~160 for the full-matrix form, and a third fewer live registers.  Agreement with the compiled
// reference on the fixtures: <= 1.3e-14 relative (prototype checked on the CPU against its lengths before porting).
struct Hh2 { double C, S, rn; bool on; };
__device__ __forceinline__ Hh2 make_hh2(double a, double b)
{
  Hh2 h;
  const double n2 = fma(a, a, b * b);
  h.on = !(n2 < 1e-20);                     // get_reflector: cnorm < 1e-10 -> no reflection (mthQR.cc:27)
  const double r = h.on ? rsqrt(n2) : 0.0;
  const double sg = a < 0 ? -r : r;
  h.C = sg * a;
  h.S = sg * b;
  h.rn = (a < 0 ? n2 : -n2) * r;            // -sign(a) n
  return h;
}
__device__ __forceinline__ void hh2_pair(const Hh2& h, double& x, double& y)
{
  const double t0 = -fma(h.C, x, h.S * y), t1 = fma(h.C, y, -h.S * x);
  x = t0;
  y = t1;
}
// H [[p, q], [q, r]] H
__device__ __forceinline__ void hh2_block(const Hh2& h, double& p, double& q, double& r)
{
  const double x0 = -fma(h.C, p, h.S * q), x1 = -fma(h.C, q, h.S * r);
  const double y0 = fma(h.C, q, -h.S * p), y1 = fma(h.C, r, -h.S * q);
  p = -fma(h.C, x0, h.S * x1);
  q = fma(h.C, x1, -h.S * x0);
  r = fma(h.C, y1, -h.S * y0);
}
// returns w^T V exp(Lambda) V^T w through the reference's iteration; *fail is set if the reference would assert
#ifndef MAG_QR_INLINE
#define MAG_QR_INLINE __forceinline__
#endif
#ifndef MAG_QR_FAST_SHIFT
#define MAG_QR_FAST_SHIFT 1   /* shift through rsqrt + reciprocal instead of IEEE sqrt + divide: 2.73 -> 2.67 ms, same flags */
#endif
__device__ MAG_QR_INLINE double quad_expm_qr3(double m00, double m01, double m02, double m11, double m12, double m22,
                                             double w0, double w1, double w2, int* fail)
{
  double d0 = m00, d1 = m11, d2 = m22, e0 = m01, e1 = m12, f = m02;
  {  // reduction to Hessenberg (= tridiagonal) form: reflector from rows 1..2 of column 0 (mthQR.cc:193-200)
    const Hh2 h = make_hh2(e0, f);
    if (h.on) {
      hh2_block(h, d1, e1, d2);
      e0 = h.rn;
      f = 0.0;
      hh2_pair(h, w1, w2);
    }
  }
  int red = 3;
  for (int it = 0; it < 100; ++it) {
    if (red == 3 && fabs(e1) < 1e-10) red = 2;
    if (red == 2 && fabs(e0) < 1e-10) red = 1;
    if (red == 1) return exp(d0) * (w0 * w0) + exp(d1) * (w1 * w1) + exp(d2) * (w2 * w2);
    const double amm1 = red == 3 ? d1 : d0, am = red == 3 ? d2 : d1, bmm1 = red == 3 ? e1 : e0;
    const double sig = 0.5 * (amm1 - am);
    const double b2 = bmm1 * bmm1;
#if MAG_QR_FAST_SHIFT
    const double s2 = fma(sig, sig, b2);
    const double denom = fabs(sig) + s2 * rsqrt(s2);          // s2 == 0: NaN, caught by the test below like a zero denominator
    if (!(fabs(denom) > 1e-10)) { *fail = 1; return 0.0; }
    const double mu = fma(sig < 0 ? b2 : -b2, __drcp_rn(denom), am);
#else
    const double denom = fabs(sig) + sqrt(fma(sig, sig, b2));
    if (!(fabs(denom) > 1e-10)) { *fail = 1; return 0.0; }
    const double mu = am - (sig < 0 ? -b2 : b2) / denom;
#endif
    double a0 = d0 - mu, a1 = d1 - mu, a2 = d2 - mu;
    const Hh2 h0 = make_hh2(a0, e0);
    const double t11 = h0.on ? fma(h0.C, a1, -h0.S * e0) : a1;   // (H0 (T - mu))(1,1)
    const Hh2 h1 = make_hh2(t11, e1);                             // column 1 of the left product, rows 1..2
    if (h0.on) {
      hh2_block(h0, a0, e0, a1);
      hh2_pair(h0, f, e1);
      hh2_pair(h0, w0, w1);
    }
    if (h1.on) {
      hh2_block(h1, a1, e1, a2);
      hh2_pair(h1, e0, f);
      hh2_pair(h1, w1, w2);
    }
    d0 = a0 + mu; d1 = a1 + mu; d2 = a2 + mu;
  }
  *fail = 1;   // not converged in 100 iterations: apf::eigen asserts (apfMatrix.cc:76)
  return 0.0;
}
__device__ __forceinline__ double edge_logm(const double* __restrict__ a, const double* __restrict__ b, int* eig_fail)
{
  const double jx = 0.5 * (b[0] - a[0]), jy = 0.5 * (b[1] - a[1]), jz = 0.5 * (b[2] - a[2]);
  double len = 0;
#pragma unroll 1
  for (int p = 0; p < 2; ++p) {
    const double wa = p ? kNP1 : kNP0, wb = p ? kNP0 : kNP1;
    double L[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) L[i] = a[3 + i] * wa + b[3 + i] * wb;
    // the field is symmetric up to the rounding of its construction (maSize.cc:343-346,491-499): use the symmetric part
    len += sqrt(quad_expm_qr3(L[0], 0.5 * (L[1] + L[3]), 0.5 * (L[2] + L[6]), L[4], 0.5 * (L[5] + L[7]), L[8], jx, jy, jz, eig_fail));
  }
  return len;
}

// mean ratio cubed with fixed Q, evaluated in metric space: y_i = (x_i - x_0) Q, the six edges are
// y1, y2, y3, y2-y1, y3-y1, y3-y2 (the reference's 01,12,20,03,13,23 up to sign), V = det[y1;y2;y3] / 6
// (= det(J) det(Q) / 6).  Differences are taken in physical space first so no accuracy is lost far from the origin.
// quality from the metric-space edge vectors y_i = (x_i - x_0) Q, i = 1..3
__device__ __forceinline__ double tet_quality_y(const double y[3][3])
{
  // sum of the six squared edge lengths of the points 0, y1, y2, y3:  4 sum |y_i|^2 - |sum y_i|^2  (every term of the
  // difference is bounded by 4 s, so no accuracy is lost: s >= sum |y_i|^2)
  const double n0 = dot3(y[0][0], y[0][1], y[0][2], y[0][0], y[0][1], y[0][2]);
  const double n1 = dot3(y[1][0], y[1][1], y[1][2], y[1][0], y[1][1], y[1][2]);
  const double n2 = dot3(y[2][0], y[2][1], y[2][2], y[2][0], y[2][1], y[2][2]);
  const double Sx = __dadd_rn(__dadd_rn(y[0][0], y[1][0]), y[2][0]), Sy = __dadd_rn(__dadd_rn(y[0][1], y[1][1]), y[2][1]),
               Sz = __dadd_rn(__dadd_rn(y[0][2], y[1][2]), y[2][2]);
  const double s = fma(4.0, __dadd_rn(__dadd_rn(n0, n1), n2), -dot3(Sx, Sy, Sz, Sx, Sy, Sz));
  const double m0 = diffprod(y[1][1], y[2][2], y[1][2], y[2][1]), m1 = diffprod(y[1][0], y[2][2], y[1][2], y[2][0]),
               m2 = diffprod(y[1][0], y[2][1], y[1][1], y[2][0]);
  const double det = fma(y[0][2], m2, fma(-y[0][1], m1, __dmul_rn(y[0][0], m0)));
  const double V = __dmul_rn(det, 1.0 / 6.0);
  const double q = __ddiv_rn(__dmul_rn(15552.0, __dmul_rn(V, V)), __dmul_rn(__dmul_rn(s, s), s));
  return V < 0 ? -q : q;
}
__device__ __forceinline__ double tet_quality(const V3 x[4], const M3& Q, double /*detQ*/)
{
  double y[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double ex = x[i + 1].x - x[0].x, ey = x[i + 1].y - x[0].y, ez = x[i + 1].z - x[0].z;
#pragma unroll
    for (int k = 0; k < 3; ++k) y[i][k] = fma(ez, Q.m[2][k], fma(ey, Q.m[1][k], __dmul_rn(ex, Q.m[0][k])));
  }
  return tet_quality_y(y);
}

// triangle mean ratio in metric space: y_i = (x_i - x_0) Q, edges y1, y2 - y1, y2; A = |y1 x y2| / 2
__device__ __forceinline__ double tri_quality(const V3 x[3], const M3& Q)
{
  double y[2][3];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double ex = x[i + 1].x - x[0].x, ey = x[i + 1].y - x[0].y, ez = x[i + 1].z - x[0].z;
#pragma unroll
    for (int k = 0; k < 3; ++k) y[i][k] = ex * Q.m[0][k] + ey * Q.m[1][k] + ez * Q.m[2][k];
  }
  const double u = y[1][0] - y[0][0], v = y[1][1] - y[0][1], w = y[1][2] - y[0][2];
  const double s = (y[0][0] * y[0][0] + y[0][1] * y[0][1] + y[0][2] * y[0][2]) + (u * u + v * v + w * w) +
                   (y[1][0] * y[1][0] + y[1][1] * y[1][1] + y[1][2] * y[1][2]);
  const double cx = y[0][1] * y[1][2] - y[0][2] * y[1][1], cy = y[0][2] * y[1][0] - y[0][0] * y[1][2],
               cz = y[0][0] * y[1][1] - y[0][1] * y[1][0];
  const double A2 = 0.25 * (cx * cx + cy * cy + cz * cz);
  return 48.0 * A2 / (s * s);
}

} // namespace magfa
