/* ma_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded restatement of the arithmetic SCOREC/core's
 * MeshAdapt executes on the marking / quality hot path, operation by operation
 * and in the reference's evaluation order, so that (compiled for baseline
 * x86-64 with -ffp-contract=off, as the reference is) it reproduces the
 * reference's fp64 results BIT FOR BIT.  Parity status: PINNED -- tests/
 * compare every function here against the compiled, unmodified reference
 * (oracle/_ref/libref_oracle.so, built from /root/reference by
 * oracle/ref/Makefile) and against golden vectors that reference produced
 * (tests/golden/, made by tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * use this file.  The product (core_b200/) never links or calls it.
 *
 * Citations are relative to /root/reference.
 */
#include "ma_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- apf::Vector
 * apf/apfVector.h:58-130 */
static double dot3(const double* a, const double* b)
{
  double r = 0; /* :103-109 */
  r += a[0] * b[0];
  r += a[1] * b[1];
  r += a[2] * b[2];
  return r;
}
static double len3(const double* a) { return sqrt(dot3(a, a)); } /* :111 */
static void normalize3(double* a) /* :113 = (*this)/getLength(), :90-96 */
{
  double l = len3(a);
  a[0] = a[0] / l; a[1] = a[1] / l; a[2] = a[2] / l;
}
static void cross3(const double* a, const double* b, double* r) /* :121-128 */
{
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}

/* ---------------------------------------------------------------- apf::Matrix
 * apf/apfMatrix.h:94-106 (mat*mat), apf/apfMatrix.cc:85-120 (determinant) */
static void matmul3(const double a[3][3], const double b[3][3], double r[3][3])
{
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = a[i][0] * b[0][j];
      s += a[i][1] * b[1][j];
      s += a[i][2] * b[2][j];
      r[i][j] = s;
    }
}
static double det2(double b00, double b01, double b10, double b11)
{
  /* getDeterminant<2,2>: d=0; d+=B00*cof(0,0); d+=B10*cof(1,0);
     cof(0,0)=+1*B11, cof(1,0)=-1*B01 */
  double d = 0;
  d += b00 * (1.0 * b11);
  d += b10 * (-1.0 * b01);
  return d;
}
double mao_det3(const double A[3][3])
{
  /* expansion down column 0, minors drop row i and column 0 */
  double d = 0;
  d += A[0][0] * (1.0 * det2(A[1][1], A[1][2], A[2][1], A[2][2]));
  d += A[1][0] * (-1.0 * det2(A[0][1], A[0][2], A[2][1], A[2][2]));
  d += A[2][0] * (1.0 * det2(A[0][1], A[0][2], A[1][1], A[1][2]));
  return d;
}

/* ------------------------------------------------------------ mth::eigenQR
 * mth/mthQR.cc:7-124,186-265; matrices row-major a[i][j] = a(i,j) */
static double sgn(double x) { return (x < 0) ? -1 : 1; }
static double sq(double x) { return x * x; }

static int get_reflector(const double a[3][3], double v[3], unsigned k, unsigned o)
{
  const unsigned m = 3;
  double cnorm = 0;
  for (unsigned i = k + o; i < m; ++i) cnorm += sq(a[i][k]);
  cnorm = sqrt(cnorm);
  if (cnorm < 1e-10) return 0;
  for (unsigned i = 0; i < k + o; ++i) v[i] = 0;
  for (unsigned i = k + o; i < m; ++i) v[i] = a[i][k];
  v[k + o] += sgn(a[k + o][k]) * cnorm;
  double rnorm = 0;
  for (unsigned i = k + o; i < m; ++i) rnorm += sq(v[i]);
  rnorm = sqrt(rnorm);
  for (unsigned i = k + o; i < m; ++i) v[i] /= rnorm;
  return 1;
}
static void reflect_columns(const double v[3], double a[3][3], unsigned k, unsigned o)
{
  for (unsigned j = 0; j < 3; ++j) {
    double dot = 0;
    for (unsigned i = k + o; i < 3; ++i) dot += a[i][j] * v[i];
    for (unsigned i = k + o; i < 3; ++i) a[i][j] -= 2 * dot * v[i];
  }
}
static void reflect_rows(const double v[3], double q[3][3], unsigned k, unsigned o)
{
  for (unsigned i = 0; i < 3; ++i) {
    double dot = 0;
    for (unsigned j = k + o; j < 3; ++j) dot += q[i][j] * v[j];
    for (unsigned j = k + o; j < 3; ++j) q[i][j] -= 2 * dot * v[j];
  }
}
static void identity3(double q[3][3])
{
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) q[i][j] = (double)(i == j);
}
static void decompose_qr(const double a[3][3], double q[3][3], double r[3][3])
{
  double v[3];
  identity3(q);
  memcpy(r, a, 9 * sizeof(double));
  for (unsigned k = 0; k < 3; ++k)
    if (get_reflector(r, v, k, 0)) {
      reflect_columns(v, r, k, 0);
      reflect_rows(v, q, k, 0);
    }
}
static void mth_multiply(const double a[3][3], const double b[3][3], double c[3][3])
{
  /* mth/mth_def.h:235-248: c=0; c += a(i,l)*b(l,j) */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int l = 0; l < 3; ++l) s += a[i][l] * b[l][j];
      c[i][j] = s;
    }
}
/* returns 1 when converged; l = final iterate (eigenvalues on the diagonal),
   q = accumulated orthogonal matrix (eigenvectors in COLUMNS).  *iters gets the
   number of shifted-QR iterations taken; returns -1 if the reference would have
   hit PCU_ALWAYS_ASSERT(fabs(denom) > 1e-10) (mthQR.cc:228). */
int mao_eigen_qr(const double a[3][3], double l[3][3], double q[3][3], int* iters)
{
  /* reduceToHessenberg (m=3 => only k=0, offset 1) */
  double v[3];
  identity3(q);
  memcpy(l, a, 9 * sizeof(double));
  if (get_reflector(l, v, 0, 1)) {
    reflect_columns(v, l, 0, 1);
    reflect_rows(v, l, 0, 1);
    reflect_rows(v, q, 0, 1);
  }
  unsigned red_m = 3;
  double r_k[3][3], q_k[3][3], tmp[3][3];
  if (iters) *iters = 0;
  for (unsigned it = 0; it < 100; ++it) {
    /* reduce() */
    int more = 0;
    while (red_m > 1) {
      if (fabs(l[red_m - 2][red_m - 1]) < 1e-10 && fabs(l[red_m - 1][red_m - 2]) < 1e-10)
        --red_m;
      else { more = 1; break; }
    }
    if (!more) return 1;
    /* get_wilkinson_shift() */
    double amm1 = l[red_m - 2][red_m - 2];
    double am = l[red_m - 1][red_m - 1];
    double bmm1 = l[red_m - 2][red_m - 1];
    double sig = (amm1 - am) / 2;
    double denom = (fabs(sig) + sqrt(sq(sig) + sq(bmm1)));
    if (!(fabs(denom) > 1e-10)) return -1;
    double mu = am - ((sgn(sig) * sq(bmm1)) / denom);
    for (int i = 0; i < 3; ++i) l[i][i] -= mu;
    decompose_qr(l, q_k, r_k);
    mth_multiply(r_k, q_k, l);
    for (int i = 0; i < 3; ++i) l[i][i] -= -mu;
    mth_multiply(q, q_k, tmp);
    memcpy(q, tmp, sizeof(tmp));
    if (iters) *iters = (int)it + 1;
  }
  return 0;
}
/* apf::eigen (apf/apfMatrix.cc:68-83): eigenvector j in ROW j of vecs */
int mao_eigen(const double A[3][3], double vecs[3][3], double vals[3])
{
  double L[3][3], Q[3][3];
  int rc = mao_eigen_qr(A, L, Q, 0);
  for (int i = 0; i < 3; ++i) vals[i] = L[i][i];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) vecs[j][i] = Q[i][j];
  return rc;
}

/* ----------------------------------------------------- frame clean-up + Q=R*S
 * ma/maSize.cc:94-121 (orthogonalizeR) and :123-142 share the row operations */
static void gram_schmidt_rows(double RT[3][3])
{
  normalize3(RT[0]);
  double d = dot3(RT[0], RT[1]);
  for (int i = 0; i < 3; ++i) { double t = RT[0][i] * d; RT[1][i] = RT[1][i] - t; }
  normalize3(RT[1]);
  double c[3];
  cross3(RT[0], RT[1], c);
  RT[2][0] = c[0]; RT[2][1] = c[1]; RT[2][2] = c[2];
}
static void transpose3(const double a[3][3], double r[3][3])
{
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r[j][i] = a[i][j];
}
/* AnisoSizeField::getTransform body after interpolation, maSize.cc:406-412 */
void mao_transform_aniso(const double h[3], const double Rin[3][3], double Q[3][3])
{
  double RT[3][3], R[3][3];
  transpose3(Rin, RT);
  gram_schmidt_rows(RT);
  transpose3(RT, R);
  double S[3][3] = {{1 / h[0], 0, 0}, {0, 1 / h[1], 0}, {0, 0, 1 / h[2]}};
  matmul3(R, S, Q);
}
/* LogAnisoSizeField::getTransform body after interpolation, maSize.cc:511-521 */
int mao_transform_logm(const double logM[3][3], double Q[3][3])
{
  double vals[3], RT[3][3], R[3][3];
  int rc = mao_eigen(logM, RT, vals);
  gram_schmidt_rows(RT);
  transpose3(RT, R);
  double S[3][3] = {{sqrt(exp(vals[0])), 0, 0}, {0, sqrt(exp(vals[1])), 0}, {0, 0, sqrt(exp(vals[2]))}};
  matmul3(R, S, Q);
  return rc;
}

/* logM vertex field construction.
   variant 0: LogAnisoSizeField::init from sizes/frames fields, maSize.cc:491-499
   variant 1: LogMEval from a user function, maSize.cc:343-346 */
void mao_logm_from_frame(int variant, const double h[3], const double R[3][3], double out[3][3])
{
  double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, RT[3][3], T[3][3];
  for (int i = 0; i < 3; ++i)
    S[i][i] = variant ? -2 * log(h[i]) : log(1 / h[i] / h[i]);
  transpose3(R, RT);
  matmul3(R, S, T);
  matmul3(T, RT, out);
}

/* ----------------------------------------------------- field interpolation
 * apf/apfElement.cc:106-114 with Linear shapes apf/apfShape.cc:116-139,203-230 */
static void interp(const double* const* node, const double* N, int nen, int nc, double* c)
{
  for (int ci = 0; ci < nc; ++ci) c[ci] = 0;
  for (int ni = 0; ni < nen; ++ni)
    for (int ci = 0; ci < nc; ++ci)
      c[ci] += node[ni][ci] * N[ni];
}

typedef struct {
  int kind;
  const double* a; /* ISO: s[nv]; ANISO: h[nv][3] */
  const double* b; /* ANISO: R[nv][9]; LOGM: logM[nv][9] */
} metric_t;

/* SizeField::getTransform at a point of an entity with nen nodes / shape
   values N.  ISO goes through the generic anisotropic path with R=I,
   h=(s,s,s) exactly as IsoWrapper does (maSize.cc:242-257). */
static int transform_at(const metric_t* mt, const int32_t* verts, const double* N, int nen, double Q[3][3])
{
  const double* node[4];
  if (mt->kind == MAO_IDENTITY) { identity3(Q); return 1; }
  if (mt->kind == MAO_LOGM) {
    double M[3][3];
    for (int i = 0; i < nen; ++i) node[i] = mt->b + 9 * (size_t)verts[i];
    interp(node, N, nen, 9, &M[0][0]);
    return mao_transform_logm(M, Q);
  }
  double h[3], R[3][3];
  if (mt->kind == MAO_ISO) {
    static const double I9[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double hv[4][3];
    for (int i = 0; i < nen; ++i) {
      double s = mt->a[verts[i]];
      hv[i][0] = s; hv[i][1] = s; hv[i][2] = s;
      node[i] = hv[i];
    }
    interp(node, N, nen, 3, h);
    for (int i = 0; i < nen; ++i) node[i] = I9;
    interp(node, N, nen, 9, &R[0][0]);
  } else {
    for (int i = 0; i < nen; ++i) node[i] = mt->a + 3 * (size_t)verts[i];
    interp(node, N, nen, 3, h);
    for (int i = 0; i < nen; ++i) node[i] = mt->b + 9 * (size_t)verts[i];
    interp(node, N, nen, 9, &R[0][0]);
  }
  mao_transform_aniso(h, (const double(*)[3])R, Q);
  return 1;
}

/* ----------------------------------------------------- edge metric length
 * MetricSizeField::measure + SizeFieldIntegrator, maSize.cc:158-216;
 * order = max(1,1)+1 = 2 -> EdgeIntegration::N2 (apfIntegrate.cc:41-53);
 * IdentitySizeField::measure = apf::measure with the mesh order (1) -> N1
 * (maSize.cc:54-60, apfIntegrate.cc:27-40,689-694). */
static void edge_jacobian_row0(const double* x0, const double* x1, double* j0)
{
  /* J = tp(grad0,x0) + tp(grad1,x1), grads (-0.5,0,0),(0.5,0,0):
     apfVectorElement.cc:44-52, apfShape.cc:126-131 */
  for (int c = 0; c < 3; ++c) j0[c] = x0[c] * -0.5 + x1[c] * 0.5;
}
static double row0_length(const double* j0, const double Q[3][3])
{
  /* (J*Q) row 0, then apf::getJacobianDeterminant(.,1) = |row0|,
     apfVectorElement.cc:86-91 */
  double r[3];
  for (int j = 0; j < 3; ++j) {
    double s = j0[0] * Q[0][j];
    s += j0[1] * Q[1][j];
    s += j0[2] * Q[2][j];
    r[j] = s;
  }
  return len3(r);
}
double mao_edge_length(int kind, const double* xyz, const double* ma, const double* mb,
                       const int32_t* ev, int* status)
{
  metric_t mt = {kind, ma, mb};
  const double* x0 = xyz + 3 * (size_t)ev[0];
  const double* x1 = xyz + 3 * (size_t)ev[1];
  double j0[3];
  edge_jacobian_row0(x0, x1, j0);
  if (kind == MAO_IDENTITY) {
    /* apf::measure: N1 rule, w=2, dV = |row0(J)| */
    double m = 0;
    m += 2 * len3(j0);
    return m;
  }
  static const double xi[2] = {0.577350269189626, -0.577350269189626};
  double measurement = 0;
  for (int p = 0; p < 2; ++p) {
    double N[2] = {(1.0 - xi[p]) / 2.0, (1.0 + xi[p]) / 2.0};
    double Q[3][3];
    int rc = transform_at(&mt, ev, N, 2, Q);
    if (rc != 1 && status) *status = rc;
    double dV2 = row0_length(j0, (const double(*)[3])Q);
    measurement += 1.0 * dV2;
  }
  return measurement;
}

/* per-vertex transform = getTransform(vertex MeshElement, xi=0): Vertex shape
   value 1.0 (apfShape.cc:99-104) */
int mao_vertex_transform(int kind, const double* ma, const double* mb, int32_t v, double Q[3][3])
{
  metric_t mt = {kind, ma, mb};
  double N[1] = {1.0};
  return transform_at(&mt, &v, N, 1, Q);
}

/* ----------------------------------------------------- tet mean-ratio quality
 * measureTetQuality, maQuality.cc:139-167; getMetricWithMaxJacobean :83-108;
 * qMeasure / FixedMetricIntegrator :35-81 */
static const int tet_edge_verts[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}}; /* apfMesh.cc:53-60 */

double mao_tet_quality(int kind, const double* xyz, const double* ma, const double* mb,
                       const int32_t* tv, int use_max, int* status)
{
  metric_t mt = {kind, ma, mb};
  double Q[3][3];
  if (use_max) {
    double maxJ = -1.0;
    for (int i = 0; i < 4; ++i) {
      double cq[3][3];
      double N[1] = {1.0};
      int rc = transform_at(&mt, tv + i, N, 1, cq);
      if (rc != 1 && status) *status = rc;
      double cj = mao_det3((const double(*)[3])cq);
      if (cj > maxJ) { maxJ = cj; memcpy(Q, cq, sizeof(Q)); }
    }
    if (maxJ == -1.0) memset(Q, 0, sizeof(Q)); /* reference leaves Q unset; not reachable for det>0 */
  } else {
    double N[4] = {1 - 0.25 - 0.25 - 0.25, 0.25, 0.25, 0.25};
    int rc = transform_at(&mt, tv, N, 4, Q);
    if (rc != 1 && status) *status = rc;
  }
  const double* x[4];
  for (int i = 0; i < 4; ++i) x[i] = xyz + 3 * (size_t)tv[i];
  double l[6];
  for (int i = 0; i < 6; ++i) {
    double j0[3];
    edge_jacobian_row0(x[tet_edge_verts[i][0]], x[tet_edge_verts[i][1]], j0);
    double m = 0;
    m += 2 * row0_length(j0, (const double(*)[3])Q); /* N1: xi=0, w=2 */
    l[i] = m;
  }
  /* tet Jacobian: grads (-1,-1,-1),(1,0,0),(0,1,0),(0,0,1) (apfShape.cc:218-224) */
  static const double g[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double J[3][3];
  for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) J[i][c] = x[0][c] * g[0][i];
  for (int n = 1; n < 4; ++n)
    for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) J[i][c] = J[i][c] + x[n][c] * g[n][i];
  double JQ[3][3];
  matmul3((const double(*)[3])J, (const double(*)[3])Q, JQ);
  double V = 0;
  V += (1.0 / 6.0) * mao_det3((const double(*)[3])JQ);
  double s = 0;
  for (int i = 0; i < 6; ++i) s += l[i] * l[i];
  if (V < 0) return -15552 * (V * V) / (s * s * s);
  return 15552 * (V * V) / (s * s * s);
}

/* ----------------------------------------------------- triangle mean-ratio quality (2-D meshes)
 * measureTriQuality, maQuality.cc:110-136: 48 A^2 / (sum l^2)^2 with a fixed Q;
 * getMetricWithMaxJacobean uses getJacobianDeterminant(Q, 2) = |row0 x row1|
 * (apfVectorElement.cc:75-84) on a 2-D mesh; triangle N1 rule: xi=(1/3,1/3), w=1/2
 * (apfIntegrate.cc:134-145); triangle shape functions (apfShape.cc:141-170):
 * N = (1-xi0-xi1, xi0, xi1), grads (-1,-1),(1,0),(0,1). */
static const int tri_edge_verts[3][2] = {{0, 1}, {1, 2}, {2, 0}}; /* apfMesh.cc:28-33 */
static double gen_det2(const double A[3][3])
{
  double c[3];
  cross3(A[0], A[1], c);
  return len3(c);
}
static double tri_quality_dim(int kind, const double* xyz, const double* ma, const double* mb,
                              const int32_t* tv, int use_max, int mesh_dim, int* status)
{
  /* mesh_dim = m->getDimension() (maQuality.cc:86,100): on a 3-D mesh the "largest Jacobian" vertex of a FACE is still
     chosen by the 3x3 determinant of its transform */
  metric_t mt = {kind, ma, mb};
  double Q[3][3];
  if (use_max) {
    double maxJ = -1.0;
    for (int i = 0; i < 3; ++i) {
      double cq[3][3];
      double N[1] = {1.0};
      int rc = transform_at(&mt, tv + i, N, 1, cq);
      if (rc != 1 && status) *status = rc;
      double cj = mesh_dim == 3 ? mao_det3((const double(*)[3])cq) : gen_det2((const double(*)[3])cq);
      if (cj > maxJ) { maxJ = cj; memcpy(Q, cq, sizeof(Q)); }
    }
    if (maxJ == -1.0) memset(Q, 0, sizeof(Q));
  } else {
    double N[3] = {1 - 1. / 3. - 1. / 3., 1. / 3., 1. / 3.};
    int rc = transform_at(&mt, tv, N, 3, Q);
    if (rc != 1 && status) *status = rc;
  }
  const double* x[3];
  for (int i = 0; i < 3; ++i) x[i] = xyz + 3 * (size_t)tv[i];
  double l[3];
  for (int i = 0; i < 3; ++i) {
    double j0[3];
    edge_jacobian_row0(x[tri_edge_verts[i][0]], x[tri_edge_verts[i][1]], j0);
    double m = 0;
    m += 2 * row0_length(j0, (const double(*)[3])Q);
    l[i] = m;
  }
  static const double g[3][3] = {{-1, -1, 0}, {1, 0, 0}, {0, 1, 0}};
  double J[3][3];
  for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) J[i][c] = x[0][c] * g[0][i];
  for (int n = 1; n < 3; ++n)
    for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) J[i][c] = J[i][c] + x[n][c] * g[n][i];
  double JQ[3][3];
  matmul3((const double(*)[3])J, (const double(*)[3])Q, JQ);
  double A = 0;
  A += (1.0 / 2.0) * gen_det2((const double(*)[3])JQ);
  double s = 0;
  for (int i = 0; i < 3; ++i) s += l[i] * l[i];
  return 48 * (A * A) / (s * s);
}
double mao_tri_quality(int kind, const double* xyz, const double* ma, const double* mb,
                       const int32_t* tv, int use_max, int* status)
{
  return tri_quality_dim(kind, xyz, ma, mb, tv, use_max, 2, status);
}
int mao_tri_qualities(int kind, const double* xyz, const double* ma, const double* mb,
                      int64_t nt, const int32_t* tri_v, int use_max, double* out)
{
  int status = 1;
  for (int64_t t = 0; t < nt; ++t)
    out[t] = mao_tri_quality(kind, xyz, ma, mb, tri_v + 3 * t, use_max, &status);
  return status;
}

/* ----------------------------------------------------- sliver classification
 * getSliverCode / matchSliver, maShape.cc:35-120 (the sweep fixElementShapes' LargeAngleTetFixer runs over the
 * BAD_QUALITY tets).  J and Q at the centroid; the quality of the tet's FIRST face decides which projection is used.
 * face0_v = the vertices of getDownward(tet, 2)[0] in that face's OWN order (measureTriQuality walks the face entity).
 * apf::project apfVector.h:134-137, apf::invert apfMatrix.h:165-173 (rows = cross products of the columns, divided by the
 * cofactor determinant), matrix * vector = row dot products. */
static void project3(const double* a, const double* b, double* r)
{
  double s = dot3(a, b) / dot3(b, b);
  for (int i = 0; i < 3; ++i) r[i] = b[i] * s;
}
static void invert3(const double m[3][3], double r[3][3])
{
  double x[3][3];
  transpose3(m, x);
  cross3(x[1], x[2], r[0]);
  cross3(x[2], x[0], r[1]);
  cross3(x[0], x[1], r[2]);
  double d = mao_det3(m);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r[i][j] = r[i][j] / d;
}
int mao_sliver_code(int kind, const double* xyz, const double* ma, const double* mb, const int32_t* tv,
                    const int32_t* face0_v, double good_quality, int* status)
{
  metric_t mt = {kind, ma, mb};
  const double* x[4];
  for (int i = 0; i < 4; ++i) x[i] = xyz + 3 * (size_t)tv[i];
  static const double g[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double J0[3][3], Q[3][3], J[3][3];
  for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) J0[i][c] = x[0][c] * g[0][i];
  for (int n = 1; n < 4; ++n)
    for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) J0[i][c] = J0[i][c] + x[n][c] * g[n][i];
  double N[4] = {1 - .25 - .25 - .25, .25, .25, .25};
  int rc = transform_at(&mt, tv, N, 4, Q);
  if (rc != 1 && status) *status = rc;
  matmul3((const double(*)[3])J0, (const double(*)[3])Q, J);
  int code = 0;
  double f0 = tri_quality_dim(kind, xyz, ma, mb, face0_v, 1, 3, status);
  double JT[3][3], inv[3][3], projected[3], basis[3];
  if (f0 * f0 * f0 > good_quality * good_quality) {
    double v03[3] = {J[2][0], J[2][1], J[2][2]};
    cross3(J[0], J[1], J[2]);
    double pr[3];
    project3(v03, J[2], pr);
    for (int i = 0; i < 3; ++i) projected[i] = v03[i] - pr[i];
    transpose3((const double(*)[3])J, JT);
    invert3((const double(*)[3])JT, inv);
    for (int i = 0; i < 3; ++i) basis[i] = dot3(inv[i], projected);
    double area[3] = {1 - basis[0] - basis[1], basis[0], basis[1]};
    for (int i = 0; i < 3; ++i) if (area[i] > 0) code |= (1 << i);
    for (int i = 0; i < 3; ++i) if (area[i] > -0.10 && area[i] < 0.10) code |= ((1 << i) << 3);
  } else {
    code |= (1 << 6);
    double v02[3] = {J[1][0], J[1][1], J[1][2]};
    project3(v02, J[0], projected);
    cross3(J[0], J[1], J[2]);
    transpose3((const double(*)[3])J, JT);
    invert3((const double(*)[3])JT, inv);
    for (int i = 0; i < 3; ++i) basis[i] = dot3(inv[i], projected);
    double area[3] = {1 - basis[0] - basis[1], basis[0], basis[1]};
    for (int i = 0; i < 2; ++i) if (area[i] > 0) code |= ((1 << i) << 7);
    for (int i = 0; i < 3; ++i) if (area[i] > -0.20 && area[i] < 0.20) code |= ((1 << i) << 9);
  }
  return code;
}
/* matchSliver's two tables (maShape.cc:93-119): {rotation, code_index}, {-1,-1} = no match */
static const signed char sliver_table2d[4][4][2] =
  {{{-1,-1}, {-1,-1}, {-1,-1}, {-1,-1}},
   {{ 7, 2}, {-1,-1}, { 3, 3}, {-1,-1}},
   {{ 1, 2}, { 2, 3}, {-1,-1}, {-1,-1}},
   {{ 3, 2}, { 2, 3}, { 3, 3}, {-1,-1}}};
static const signed char sliver_table[8][8][2] =
  {{{-1,-1}, {-1,-1}, {-1,-1}, {-1,-1}, {-1,-1}, {-1,-1}, {-1,-1}, {-1,-1}},
   {{ 4, 1}, {-1,-1}, {10, 2}, { 6, 3}, { 4, 2}, { 5, 3}, { 0, 3}, {-1,-1}},
   {{ 1, 1}, { 8, 2}, {-1,-1}, { 6, 3}, { 9, 2}, { 5, 3}, { 0, 3}, {-1,-1}},
   {{ 2, 0}, { 8, 2}, {10, 2}, {-1,-1}, { 0, 2}, { 5, 3}, { 0, 3}, {-1,-1}},
   {{ 2, 1}, {11, 2}, { 2, 2}, { 6, 3}, {-1,-1}, { 5, 3}, { 0, 3}, {-1,-1}},
   {{ 0, 0}, {11, 2}, { 6, 2}, { 6, 3}, { 4, 2}, {-1,-1}, { 0, 3}, {-1,-1}},
   {{ 1, 0}, { 5, 2}, { 2, 2}, { 6, 3}, { 9, 2}, { 5, 3}, {-1,-1}, {-1,-1}},
   {{ 0, 1}, { 5, 2}, { 6, 2}, { 6, 3}, { 0, 2}, { 5, 3}, { 0, 3}, {-1,-1}}};
void mao_match_sliver(int code, int* rotation, int* code_index)
{
  const signed char* m = ((code >> 6) & 1) ? sliver_table2d[(code >> 7) & 3][(code >> 9) & 3]
                                           : sliver_table[code & 7][(code >> 3) & 7];
  *rotation = m[0];
  *code_index = m[1];
}
int mao_sliver_codes(int kind, const double* xyz, const double* ma, const double* mb, int64_t nt, const int32_t* tet_v,
                     const int32_t* face0_v, double good_quality, int32_t* codes, int32_t* match /*[nt][2]*/)
{
  int status = 1;
  for (int64_t t = 0; t < nt; ++t) {
    int c = mao_sliver_code(kind, xyz, ma, mb, tet_v + 4 * t, face0_v + 3 * t, good_quality, &status);
    codes[t] = c;
    if (match) { int r, k; mao_match_sliver(c, &r, &k); match[2 * t] = r; match[2 * t + 1] = k; }
  }
  return status;
}

/* ----------------------------------------------------- prism / pyramid validity
 * isPrismOk / isPyramidOk, maQuality.cc:490-560; apf::Plane apfGeometry.cc:28-45;
 * prism_rotation / pyramid_rotation maTables.cc */
static const int prism_rotation[6][6] = {
  {0, 1, 2, 3, 4, 5}, {1, 2, 0, 4, 5, 3}, {2, 0, 1, 5, 3, 4},
  {3, 5, 4, 0, 2, 1}, {4, 3, 5, 1, 0, 2}, {5, 4, 3, 2, 1, 0}}; /* maTables.cc:204-210 */
static const int pyramid_rotation[4][5] = {
  {0, 1, 2, 3, 4}, {1, 2, 3, 0, 4}, {2, 3, 0, 1, 4}, {3, 0, 1, 2, 4}};

typedef struct { double n[3]; double r; } plane_t;
static plane_t plane_from_points(const double* a, const double* b, const double* c)
{
  /* fromPoints: normal = cross(a-c, b-c).normalize(); radius = c*normal;
     then the Plane(n, r) constructor normalizes n AGAIN and scales r by the
     length of the (already unit) n -- apfGeometry.cc:28-40 */
  double u[3], w[3], n[3];
  plane_t p;
  for (int i = 0; i < 3; ++i) { u[i] = a[i] - c[i]; w[i] = b[i] - c[i]; }
  cross3(u, w, n);
  normalize3(n);
  double radius = dot3(c, n);
  double l = len3(n);
  p.n[0] = n[0] / l; p.n[1] = n[1] / l; p.n[2] = n[2] / l;
  p.r = radius * l;
  return p;
}
static double plane_distance(const plane_t* p, const double* x) { return dot3(p->n, x) - p->r; }

static int unrotate_prism_diagonal_code(int code, int rot)
{
  static const int shift_table[6] = {0, 1, 2, 2, 0, 1};
  int out = 0;
  for (int i = 0; i < 3; ++i)
    if (code & (1 << i)) out |= (1 << ((i + shift_table[rot]) % 3));
  return out;
}
int mao_prism_ok(const double* xyz, const int32_t* pv, int* good_codes)
{
  const double* p[6];
  for (int i = 0; i < 6; ++i) p[i] = xyz + 3 * (size_t)pv[i];
  int all_good = 1, codes = 0xFF;
  for (int i = 0; i < 6; ++i) {
    const int* n2o = prism_rotation[i];
    plane_t pl = plane_from_points(p[n2o[0]], p[n2o[1]], p[n2o[5]]);
    if (plane_distance(&pl, p[n2o[3]]) <= 0) { all_good = 0; codes &= ~(1 << unrotate_prism_diagonal_code(5, i)); }
    if (plane_distance(&pl, p[n2o[4]]) <= 0) { all_good = 0; codes &= ~(1 << unrotate_prism_diagonal_code(4, i)); }
    if (plane_distance(&pl, p[n2o[2]]) >= 0) {
      all_good = 0;
      codes &= ~(1 << unrotate_prism_diagonal_code(5, i));
      codes &= ~(1 << unrotate_prism_diagonal_code(4, i));
    }
  }
  if (good_codes) *good_codes = codes;
  return all_good;
}
int mao_pyramid_ok(const double* xyz, const int32_t* pv, int* good_rotation)
{
  const double* p[5];
  for (int i = 0; i < 5; ++i) p[i] = xyz + 3 * (size_t)pv[i];
  int all_good = 1, rot = -1;
  for (int i = 0; i < 2; ++i) {
    const int* n2o = pyramid_rotation[i];
    plane_t pl = plane_from_points(p[n2o[0]], p[n2o[2]], p[n2o[4]]);
    if (plane_distance(&pl, p[n2o[1]]) <= 0) { all_good = 0; continue; }
    if (plane_distance(&pl, p[n2o[3]]) >= 0) { all_good = 0; continue; }
    rot = i;
  }
  if (good_rotation) *good_rotation = rot;
  return all_good;
}

/* ----------------------------------------------------- predictive element weight
 * ma::getElementWeight, maBalance.cc:74-81: getSizeWeight (:21-39) -> SizeField::getWeight =
 * measure(e) / parentMeasure[type] (maSize.cc:225-229, parentMeasure[TET] = 1.0/6.0 :147-156), then
 * clampForIterations (:41-52).  measure(tet): SizeFieldIntegrator with order 2 -> TetrahedronIntegration::N2,
 * 4 points, weights 0.25/6.0 (apfIntegrate.cc:328-342); at each point dV2 = det(J*Q(xi)) with Q from
 * getTransform at the point (tet shape values apfShape.cc:203-210).  IdentitySizeField::getWeight = 1.0
 * (maSize.cc:89-92). */
double mao_tet_weight(int kind, const double* xyz, const double* ma, const double* mb,
                      const int32_t* tv, int* status)
{
  if (kind == MAO_IDENTITY) return 1.0;
  metric_t mt = {kind, ma, mb};
  static const double P[4][3] = {
    {0.138196601125011, 0.138196601125011, 0.138196601125011},
    {0.585410196624969, 0.138196601125011, 0.138196601125011},
    {0.138196601125011, 0.585410196624969, 0.138196601125011},
    {0.138196601125011, 0.138196601125011, 0.585410196624969}};
  static const double g[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  const double* x[4];
  for (int i = 0; i < 4; ++i) x[i] = xyz + 3 * (size_t)tv[i];
  double measurement = 0;
  for (int p = 0; p < 4; ++p) {
    const double* xi = P[p];
    double N[4] = {1 - xi[0] - xi[1] - xi[2], xi[0], xi[1], xi[2]};
    double Q[3][3];
    int rc = transform_at(&mt, tv, N, 4, Q);
    if (rc != 1 && status) *status = rc;
    /* apf::getJacobian at the point (constant for a linear tet, recomputed per point by the reference) */
    double J[3][3];
    for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) J[i][c] = x[0][c] * g[0][i];
    for (int n = 1; n < 4; ++n)
      for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) J[i][c] = J[i][c] + x[n][c] * g[n][i];
    double JQ[3][3];
    matmul3((const double(*)[3])J, (const double(*)[3])Q, JQ);
    double dV2 = mao_det3((const double(*)[3])JQ);
    measurement += (0.25 / 6.0) * dV2;
  }
  return measurement / (1.0 / 6.0);
}
/* measure(triangle) on a 2-D mesh: TriangleIntegration::N2 (apfIntegrate.cc:146-159), 3 points, weights 1/3/2;
 * dV2 = |row0(J Q) x row1(J Q)| (getJacobianDeterminant(., 2), apfVectorElement.cc:75-84); triangle shape values and
 * gradients apfShape.cc:141-160; parentMeasure[TRIANGLE] = 1.0/2.0 (maSize.cc:150). */
double mao_tri_weight(int kind, const double* xyz, const double* ma, const double* mb,
                      const int32_t* tv, int* status)
{
  if (kind == MAO_IDENTITY) return 1.0;
  metric_t mt = {kind, ma, mb};
  static const double P[3][2] = {{0.666666666666667, 0.166666666666667}, {0.166666666666667, 0.666666666666667},
                                 {0.166666666666667, 0.166666666666667}};
  static const double g[3][3] = {{-1, -1, 0}, {1, 0, 0}, {0, 1, 0}};
  const double* x[3];
  for (int i = 0; i < 3; ++i) x[i] = xyz + 3 * (size_t)tv[i];
  double measurement = 0;
  for (int p = 0; p < 3; ++p) {
    double N[3] = {1 - P[p][0] - P[p][1], P[p][0], P[p][1]};
    double Q[3][3];
    int rc = transform_at(&mt, tv, N, 3, Q);
    if (rc != 1 && status) *status = rc;
    double J[3][3];
    for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) J[i][c] = x[0][c] * g[0][i];
    for (int n = 1; n < 3; ++n)
      for (int i = 0; i < 3; ++i) for (int c = 0; c < 3; ++c) J[i][c] = J[i][c] + x[n][c] * g[n][i];
    double JQ[3][3];
    matmul3((const double(*)[3])J, (const double(*)[3])Q, JQ);
    measurement += (1. / 3. / 2.0) * gen_det2((const double(*)[3])JQ);
  }
  return measurement / (1.0 / 2.0);
}
int mao_tri_weights(int kind, const double* xyz, const double* ma, const double* mb,
                    int64_t nt, const int32_t* tri_v, double w_max, double w_min, double* out)
{
  int status = 1;
  for (int64_t t = 0; t < nt; ++t) {
    double w = mao_tri_weight(kind, xyz, ma, mb, tri_v + 3 * t, &status);
    out[t] = (w > w_max) ? w_max : ((w < w_min) ? w_min : w);
  }
  return status;
}
/* clamp of maBalance.cc:14-19 */
double mao_clamp(double x, double max, double min)
{
  if (x > max) return max;
  if (x < min) return min;
  return x;
}
/* weights of nt tets; w_max = pow(2, dim*refinesLeft), w_min = pow(4, -coarsensLeft) (maBalance.cc:41-52),
   pass +-HUGE_VAL for the raw SizeField::getWeight */
int mao_tet_weights(int kind, const double* xyz, const double* ma, const double* mb,
                    int64_t nt, const int32_t* tet_v, double w_max, double w_min, double* out)
{
  int status = 1;
  for (int64_t t = 0; t < nt; ++t)
    out[t] = mao_clamp(mao_tet_weight(kind, xyz, ma, mb, tet_v + 4 * t, &status), w_max, w_min);
  return status;
}

/* ----------------------------------------------------- size-field transfer to split vertices
 * ma::makeSplitVert, maRefine.cc:129-151: the new vertex sits at xi = 0 of the edge:
 * point = mapLocalToGlobal (coordinate interpolation, N = (0.5, 0.5), apfShape.cc:123-124), then
 * SizeField::interpolate(edge element, xi, vert):
 *   AnisoSizeField (maSize.cc:414-429): h = interpolated sizes, R = orthogonalizeR(interpolated frames), all 9 entries
 *   LogAnisoSizeField (:523-534): logM = interpolated logM
 *   Iso (IsoSizeField is an AnisoSizeField over h=(s,s,s), R=I): the interpolated scalar
 * out_a / out_b follow the (ma, mb) layout of the kind. */
void mao_split_vertex(int kind, const double* xyz, const double* ma, const double* mb, const int32_t* ev,
                      double* out_xyz, double* out_a, double* out_b)
{
  const double N[2] = {(1.0 - 0.0) / 2.0, (1.0 + 0.0) / 2.0};
  const double* node[2];
  node[0] = xyz + 3 * (size_t)ev[0]; node[1] = xyz + 3 * (size_t)ev[1];
  interp(node, N, 2, 3, out_xyz);
  if (kind == MAO_ISO) {
    node[0] = ma + ev[0]; node[1] = ma + ev[1];
    interp(node, N, 2, 1, out_a);
  } else if (kind == MAO_ANISO) {
    node[0] = ma + 3 * (size_t)ev[0]; node[1] = ma + 3 * (size_t)ev[1];
    interp(node, N, 2, 3, out_a);
    double R[3][3], RT[3][3];
    node[0] = mb + 9 * (size_t)ev[0]; node[1] = mb + 9 * (size_t)ev[1];
    interp(node, N, 2, 9, &R[0][0]);
    transpose3((const double(*)[3])R, RT);
    gram_schmidt_rows(RT);
    transpose3((const double(*)[3])RT, R);
    memcpy(out_b, R, sizeof(R));
  } else if (kind == MAO_LOGM) {
    node[0] = mb + 9 * (size_t)ev[0]; node[1] = mb + 9 * (size_t)ev[1];
    interp(node, N, 2, 9, out_b);
  }
}
void mao_split_vertices(int kind, const double* xyz, const double* ma, const double* mb, int64_t n,
                        const int32_t* edge_v, double* out_xyz, double* out_a, double* out_b)
{
  const int sa = kind == MAO_ISO ? 1 : 3;
  for (int64_t i = 0; i < n; ++i)
    mao_split_vertex(kind, xyz, ma, mb, edge_v + 2 * i, out_xyz + 3 * i, out_a ? out_a + sa * i : 0, out_b ? out_b + 9 * i : 0);
}

/* ----------------------------------------------------- sweeps over arrays */
int mao_edge_lengths(int kind, const double* xyz, const double* ma, const double* mb,
                     int64_t ne, const int32_t* edge_v, double* out)
{
  int status = 1;
  for (int64_t e = 0; e < ne; ++e)
    out[e] = mao_edge_length(kind, xyz, ma, mb, edge_v + 2 * e, &status);
  return status;
}
int mao_tet_qualities(int kind, const double* xyz, const double* ma, const double* mb,
                      int64_t nt, const int32_t* tet_v, int use_max, double* out)
{
  int status = 1;
  for (int64_t t = 0; t < nt; ++t)
    out[t] = mao_tet_quality(kind, xyz, ma, mb, tet_v + 4 * t, use_max, &status);
  return status;
}
void mao_vertex_transforms(int kind, const double* ma, const double* mb, int64_t nv, double* Q9)
{
  for (int64_t v = 0; v < nv; ++v)
    mao_vertex_transform(kind, ma, mb, (int32_t)v, (double(*)[3])(Q9 + 9 * v));
}

/* ma::markEntities (maAdapt.cc:293-324) over precomputed predicate values.
   cmp 0: value > thr (ShouldSplit, maSize.cc:217-220)
   cmp 1: value < thr (ShouldCollapse :221-224, IsBadQuality maShape.cc:122-130)
   returns owned-true count, or -1-(index) of the first entity that already
   carries trueFlag (the reference asserts, maAdapt.cc:308). */
int64_t mao_mark_entities(int64_t n, const double* value, int cmp, double thr,
                          int32_t* flags, const uint8_t* owned,
                          int32_t true_flag, int32_t set_false_flag, int32_t all_false_flags)
{
  if (!all_false_flags) all_false_flags = set_false_flag;
  int64_t count = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (flags[i] & true_flag) return -1 - i;
    if (all_false_flags & flags[i]) continue;
    int pred = cmp ? (value[i] < thr) : (value[i] > thr);
    if (pred) {
      flags[i] |= true_flag;
      if (!owned || owned[i]) ++count;
    } else
      flags[i] |= set_false_flag;
  }
  return count;
}

/* ma::getMinQuality (maShape.cc:152-169): min over simplex elements, init 1 */
double mao_min_quality(int64_t n, const double* q)
{
  double minqual = 1;
  for (int64_t i = 0; i < n; ++i) if (q[i] < minqual) minqual = q[i];
  return minqual;
}
/* ma::getMaximumEdgeLength (maSize.cc:673-691): max over owned edges, init 0 */
double mao_max_length(int64_t n, const double* len, const uint8_t* owned)
{
  double mx = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    if (owned && !owned[i]) continue;
    if (len[i] > mx) mx = len[i];
  }
  return mx;
}
