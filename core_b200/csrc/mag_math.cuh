// mag_math.cuh -- device arithmetic of the MeshAdapt sweep in the reference's evaluation order,
// templated on an arithmetic policy:
//   StrictOps: every operation is a correctly rounded IEEE fp64 add / mul / div / sqrt issued through
//              the __d*_rn intrinsics, which nvcc never contracts into FMAs -> bit-identical to the
//              reference's x86-64 (no-FMA) build;
//   FusedOps:  the same expressions with plain operators, so nvcc contracts a*b+c into DFMA.
// Evaluation order follows (paths relative to the SCOREC/core tree):
//   apf/apfVector.h:58-130 (dot, length, normalize, cross)
//   apf/apfMatrix.h:94-106 (mat*mat), apf/apfMatrix.cc:85-120 (cofactor determinant)
//   apf/apfElement.cc:106-114 + apf/apfShape.cc:116-139,203-230 (linear interpolation)
//   apf/apfVectorElement.cc:44-91 (Jacobian, |row0|), apf/apfIntegrate.cc:27-53,316-327
//   ma/maSize.cc:94-142,158-216,395-413,506-522; ma/maQuality.cc:35-167
//   mth/mthQR.cc:7-124,186-265 (Householder Hessenberg + Wilkinson-shift QR)
// Additions of an exact zero that the reference performs (0 + x, x + y*0) are
// dropped: they only ever change the sign of a zero result.
#pragma once
#include <cuda_runtime.h>

struct V3 { double x, y, z; };
struct M3 { double m[3][3]; };

struct StrictOps {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double sqrt_(double a) { return __dsqrt_rn(a); }
};
struct FusedOps {
  static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
  static __device__ __forceinline__ double add(double a, double b) { return a + b; }
  static __device__ __forceinline__ double sub(double a, double b) { return a - b; }
  static __device__ __forceinline__ double div(double a, double b) { return a / b; }
  static __device__ __forceinline__ double sqrt_(double a) { return sqrt(a); }
};

template <class OPS>
struct MagMath {
  static __device__ __forceinline__ double mul(double a, double b) { return OPS::mul(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return OPS::add(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return OPS::sub(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return OPS::div(a, b); }
  static __device__ __forceinline__ double sqrt_(double a) { return OPS::sqrt_(a); }



// r = 0; r += a0*b0; r += a1*b1; r += a2*b2
static __device__ __forceinline__ double dot(const V3& a, const V3& b)
{
  return add(add(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z));
}
static __device__ __forceinline__ double length(const V3& a) { return sqrt_(dot(a, a)); }
static __device__ __forceinline__ V3 normalize(const V3& a)
{
  double l = length(a);
  return V3{div(a.x, l), div(a.y, l), div(a.z, l)};
}
static __device__ __forceinline__ V3 cross(const V3& a, const V3& b)
{
  return V3{sub(mul(a.y, b.z), mul(a.z, b.y)),
            sub(mul(a.z, b.x), mul(a.x, b.z)),
            sub(mul(a.x, b.y), mul(a.y, b.x))};
}

// 2x2: d = 0 + B00*(+B11) + B10*(-B01)
static __device__ __forceinline__ double det2(double b00, double b01, double b10, double b11)
{
  return add(mul(b00, b11), mul(b10, -b01));
}
// cofactor expansion down column 0 (apfMatrix.cc:104-120)
static __device__ __forceinline__ double det3(const M3& A)
{
  double m0 = det2(A.m[1][1], A.m[1][2], A.m[2][1], A.m[2][2]);
  double m1 = det2(A.m[0][1], A.m[0][2], A.m[2][1], A.m[2][2]);
  double m2 = det2(A.m[0][1], A.m[0][2], A.m[1][1], A.m[1][2]);
  return add(add(mul(A.m[0][0], m0), mul(A.m[1][0], -m1)), mul(A.m[2][0], m2));
}

// rows of RT are the frame vectors; maSize.cc:108-116 / :135-138
static __device__ __forceinline__ void gram_schmidt(V3& r0, V3& r1, V3& r2)
{
  r0 = normalize(r0);
  double d = dot(r0, r1);
  r1 = V3{sub(r1.x, mul(r0.x, d)), sub(r1.y, mul(r0.y, d)), sub(r1.z, mul(r0.z, d))};
  r1 = normalize(r1);
  r2 = cross(r0, r1);
}

// Q = R * diag(s), R = transpose([r0;r1;r2]): Q[i][k] = r_k[i] * s_k
static __device__ __forceinline__ void frame_times_diag(const V3& r0, const V3& r1, const V3& r2,
                                                 double s0, double s1, double s2, M3& Q)
{
  Q.m[0][0] = mul(r0.x, s0); Q.m[1][0] = mul(r0.y, s0); Q.m[2][0] = mul(r0.z, s0);
  Q.m[0][1] = mul(r1.x, s1); Q.m[1][1] = mul(r1.y, s1); Q.m[2][1] = mul(r1.z, s1);
  Q.m[0][2] = mul(r2.x, s2); Q.m[1][2] = mul(r2.y, s2); Q.m[2][2] = mul(r2.z, s2);
}

// AnisoSizeField::getTransform after interpolation (maSize.cc:406-412).
// c0, c1 = columns 0 and 1 of the interpolated frame matrix (column 2 is
// discarded by orthogonalizeR).
static __device__ __forceinline__ void transform_aniso(V3 c0, V3 c1, double h0, double h1, double h2, M3& Q)
{
  V3 c2;
  gram_schmidt(c0, c1, c2);
  frame_times_diag(c0, c1, c2, div(1.0, h0), div(1.0, h1), div(1.0, h2), Q);
}

// ---------------------------------------------------------------- mth::eigenQR
template <int K, int O>
static __device__ __forceinline__ bool get_reflector(const M3& a, double v[3])
{
  double cnorm = 0;
#pragma unroll
  for (int i = K + O; i < 3; ++i) cnorm = add(cnorm, mul(a.m[i][K], a.m[i][K]));
  cnorm = sqrt_(cnorm);
  if (cnorm < 1e-10) return false;
#pragma unroll
  for (int i = 0; i < K + O; ++i) v[i] = 0;
#pragma unroll
  for (int i = K + O; i < 3; ++i) v[i] = a.m[i][K];
  v[K + O] = add(v[K + O], mul((a.m[K + O][K] < 0) ? -1.0 : 1.0, cnorm));
  double rnorm = 0;
#pragma unroll
  for (int i = K + O; i < 3; ++i) rnorm = add(rnorm, mul(v[i], v[i]));
  rnorm = sqrt_(rnorm);
#pragma unroll
  for (int i = K + O; i < 3; ++i) v[i] = div(v[i], rnorm);
  return true;
}
template <int KO>
static __device__ __forceinline__ void reflect_columns(const double v[3], M3& a)
{
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    double dt = 0;
#pragma unroll
    for (int i = KO; i < 3; ++i) dt = add(dt, mul(a.m[i][j], v[i]));
#pragma unroll
    for (int i = KO; i < 3; ++i) a.m[i][j] = sub(a.m[i][j], mul(mul(2.0, dt), v[i]));
  }
}
template <int KO>
static __device__ __forceinline__ void reflect_rows(const double v[3], M3& q)
{
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double dt = 0;
#pragma unroll
    for (int j = KO; j < 3; ++j) dt = add(dt, mul(q.m[i][j], v[j]));
#pragma unroll
    for (int j = KO; j < 3; ++j) q.m[i][j] = sub(q.m[i][j], mul(mul(2.0, dt), v[j]));
  }
}
static __device__ __forceinline__ void identity(M3& q)
{
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) q.m[i][j] = (i == j) ? 1.0 : 0.0;
}
template <int K>
static __device__ __forceinline__ void qr_step(M3& r, M3& q)
{
  double v[3];
  if (get_reflector<K, 0>(r, v)) {
    reflect_columns<K>(v, r);
    reflect_rows<K>(v, q);
  }
}
// c = 0; c += a(i,l)*b(l,j)  (mth_def.h:235-248)
static __device__ __forceinline__ void mth_multiply(const M3& a, const M3& b, M3& c)
{
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c.m[i][j] = add(add(add(0.0, mul(a.m[i][0], b.m[0][j])), mul(a.m[i][1], b.m[1][j])), mul(a.m[i][2], b.m[2][j]));
}
// returns 1 converged, 0 not converged in 100 iterations, -1 zero Wilkinson denominator
static __device__ __noinline__ int eigen_qr(const M3& a, M3& l, M3& q)
{
  double v[3];
  identity(q);
  l = a;
  if (get_reflector<0, 1>(l, v)) {
    reflect_columns<1>(v, l);
    reflect_rows<1>(v, l);
    reflect_rows<1>(v, q);
  }
  int red_m = 3;
  for (int it = 0; it < 100; ++it) {
    if (red_m == 3 && fabs(l.m[1][2]) < 1e-10 && fabs(l.m[2][1]) < 1e-10) red_m = 2;
    if (red_m == 2 && fabs(l.m[0][1]) < 1e-10 && fabs(l.m[1][0]) < 1e-10) red_m = 1;
    if (red_m == 1) return 1;
    double amm1, am, bmm1;
    if (red_m == 3) { amm1 = l.m[1][1]; am = l.m[2][2]; bmm1 = l.m[1][2]; }
    else { amm1 = l.m[0][0]; am = l.m[1][1]; bmm1 = l.m[0][1]; }
    double sig = div(sub(amm1, am), 2.0);
    double denom = add(fabs(sig), sqrt_(add(mul(sig, sig), mul(bmm1, bmm1))));
    if (!(fabs(denom) > 1e-10)) return -1;
    double mu = sub(am, div(mul((sig < 0) ? -1.0 : 1.0, mul(bmm1, bmm1)), denom));
#pragma unroll
    for (int i = 0; i < 3; ++i) l.m[i][i] = sub(l.m[i][i], mu);
    M3 qk, rk = l;
    identity(qk);
    qr_step<0>(rk, qk);
    qr_step<1>(rk, qk);
    qr_step<2>(rk, qk);
    mth_multiply(rk, qk, l);
#pragma unroll
    for (int i = 0; i < 3; ++i) l.m[i][i] = sub(l.m[i][i], -mu);
    M3 t;
    mth_multiply(q, qk, t);
    q = t;
  }
  return 0;
}

// LogAnisoSizeField::getTransform after interpolation (maSize.cc:511-521);
// apf::eigen puts eigenvector j (column j of q) in row j (apfMatrix.cc:79-81).
static __device__ __forceinline__ int transform_logm(const M3& logM, M3& Q)
{
  M3 L, E;
  int rc = eigen_qr(logM, L, E);
  V3 r0{E.m[0][0], E.m[1][0], E.m[2][0]};
  V3 r1{E.m[0][1], E.m[1][1], E.m[2][1]};
  V3 r2;
  gram_schmidt(r0, r1, r2);
  frame_times_diag(r0, r1, r2, sqrt_(exp(L.m[0][0])), sqrt_(exp(L.m[1][1])), sqrt_(exp(L.m[2][2])), Q);
  return rc;
}

// c = 0; c += a*Na; c += b*Nb  (two-node interpolation)
static __device__ __forceinline__ double lerp2(double a, double na, double b, double nb)
{
  return add(mul(a, na), mul(b, nb));
}

// row 0 of the edge Jacobian: x0*(-0.5) + x1*0.5 (apfVectorElement.cc:44-52)
static __device__ __forceinline__ V3 edge_j0(const V3& x0, const V3& x1)
{
  return V3{add(mul(x0.x, -0.5), mul(x1.x, 0.5)),
            add(mul(x0.y, -0.5), mul(x1.y, 0.5)),
            add(mul(x0.z, -0.5), mul(x1.z, 0.5))};
}
// |row0(J*Q)|
static __device__ __forceinline__ double row0_length(const V3& j, const M3& Q)
{
  V3 r;
  r.x = add(add(mul(j.x, Q.m[0][0]), mul(j.y, Q.m[1][0])), mul(j.z, Q.m[2][0]));
  r.y = add(add(mul(j.x, Q.m[0][1]), mul(j.y, Q.m[1][1])), mul(j.z, Q.m[2][1]));
  r.z = add(add(mul(j.x, Q.m[0][2]), mul(j.y, Q.m[1][2])), mul(j.z, Q.m[2][2]));
  return length(r);
}

// measureTetQuality with a fixed Q (maQuality.cc:155-166)
static __device__ __forceinline__ double tet_quality(const V3 x[4], const M3& Q)
{
  // edges in tet_edge_verts order {01,12,20,03,13,23}; l = 0 + 2*|row0|
  const int ea[6] = {0, 1, 2, 0, 1, 2}, eb[6] = {1, 2, 0, 3, 3, 3};
  double s = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double l = mul(2.0, row0_length(edge_j0(x[ea[i]], x[eb[i]]), Q));
    s = add(s, mul(l, l));
  }
  // J rows = x1-x0, x2-x0, x3-x0  (-x0 + xn)
  M3 J, JQ;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    J.m[i][0] = add(-x[0].x, x[i + 1].x);
    J.m[i][1] = add(-x[0].y, x[i + 1].y);
    J.m[i][2] = add(-x[0].z, x[i + 1].z);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      JQ.m[i][j] = add(add(mul(J.m[i][0], Q.m[0][j]), mul(J.m[i][1], Q.m[1][j])), mul(J.m[i][2], Q.m[2][j]));
  double V = mul(1.0 / 6.0, det3(JQ));
  double s3 = mul(mul(s, s), s);
  double c = (V < 0) ? -15552.0 : 15552.0;
  return div(mul(c, mul(V, V)), s3);
}


// measureTriQuality with a fixed Q (maQuality.cc:126-135): edges in tri_edge_verts order {01,12,20}, l = 2*|row0|;
// A = |row0(JQ) x row1(JQ)| / 2 with J rows x1-x0, x2-x0 (triangle N1 rule, w = 1/2)
static __device__ __forceinline__ double tri_quality(const V3 x[3], const M3& Q)
{
  const int ea[3] = {0, 1, 2}, eb[3] = {1, 2, 0};
  double s = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double l = mul(2.0, row0_length(edge_j0(x[ea[i]], x[eb[i]]), Q));
    s = add(s, mul(l, l));
  }
  V3 jq[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double j0 = add(-x[0].x, x[i + 1].x), j1 = add(-x[0].y, x[i + 1].y), j2 = add(-x[0].z, x[i + 1].z);
    jq[i].x = add(add(mul(j0, Q.m[0][0]), mul(j1, Q.m[1][0])), mul(j2, Q.m[2][0]));
    jq[i].y = add(add(mul(j0, Q.m[0][1]), mul(j1, Q.m[1][1])), mul(j2, Q.m[2][1]));
    jq[i].z = add(add(mul(j0, Q.m[0][2]), mul(j1, Q.m[1][2])), mul(j2, Q.m[2][2]));
  }
  double A = mul(1.0 / 2.0, length(cross(jq[0], jq[1])));
  return div(mul(48.0, mul(A, A)), mul(s, s));
}
};

typedef MagMath<StrictOps> magst;
typedef MagMath<FusedOps> magfu;
