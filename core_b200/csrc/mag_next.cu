// mag_next.cu -- the sweeps either side of the marking path, over the same device-resident part:
//   mag_element_weights   predictive load-balance weight of every element
//                         ma::getElementWeights / getElementWeight (ma/maBalance.cc:21-52,74-97),
//                         SizeField::getWeight = measure(element) / parentMeasure (ma/maSize.cc:147-156,225-229)
//   mag_split_vertices    position and size-field values of the vertex that will split every SPLIT-marked edge
//                         ma::makeSplitVert (ma/maRefine.cc:129-151), SizeField::interpolate (ma/maSize.cc:414-429,523-534)
// MAG_FP_STRICT follows the reference's operation order (StrictOps of mag_math.cuh: bit-identical); MAG_FP_FAST uses
// algebraically equivalent, cheaper forms (within 1e-12 relative).
#include "mag_internal.h"
#include "mag_layout.cuh"
#include "mag_math.cuh"
#include <cmath>

using namespace maglay;

int magk_vertex_pass(mag_ctx* c);
int magk_cavity_quality(mag_ctx* c, int fp_mode, int64_t ncav, const int64_t* d_off, const int32_t* d_tv, int use_max, double* d_worst, double* d_qual);
int magk_build_v2t(mag_ctx* c);
int magk_collapse_quality(mag_ctx* c, int fp_mode, int64_t ncand, const int32_t* d_edge, const uint8_t* d_end, int use_max,
                          double* d_new, double* d_old, int32_t* d_keep);

namespace {

constexpr int kWThreads = 128;

// ------------------------------------------------------------------ element weights
// measure(tet): SizeFieldIntegrator, order 2 -> TetrahedronIntegration::N2 (apf/apfIntegrate.cc:328-342): 4 points,
// weights 0.25/6; at each point dV2 = det(J * Q(xi)), Q = getTransform at the point from the four vertex values
// (tet shape values apf/apfShape.cc:203-210; interpolation c = 0; c += node_n * N_n in node order, apfElement.cc:106-114).
template <int KIND, class OPS>
__device__ __forceinline__ double tet_weight(const double* __restrict__ vedge, const int4& tv, int* eig_fail)
{
  typedef MagMath<OPS> MM;
  if (KIND == MAG_KIND_IDENTITY) return 1.0;   // IdentitySizeField::getWeight (maSize.cc:89-92)
  constexpr int N = (KIND == MAG_KIND_ISO) ? 4 : 12;
  const int32_t vid[4] = {tv.x & kVidMask, tv.y, tv.z, tv.w};
  double rec[4][N];
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    if (KIND == MAG_KIND_ISO) load_rec4(vedge, vid[n], rec[n]);
    else {
      Rec12 r = load_rec12(vedge, vid[n]);
#pragma unroll
      for (int i = 0; i < 12; ++i) rec[n][i] = r.v[i];
    }
  }
  // J rows = -x0 + xn (apfVectorElement.cc:44-52 with grads (-1,-1,-1),(1,0,0),(0,1,0),(0,0,1); the x*0 terms add +-0)
  M3 J;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int c = 0; c < 3; ++c) J.m[i][c] = MM::add(-rec[0][c], rec[i + 1][c]);
  constexpr double A = 0.138196601125011, B = 0.585410196624969;
  double measurement = 0.0;
#pragma unroll 1
  for (int p = 0; p < 4; ++p) {
    const double xi0 = p == 1 ? B : A, xi1 = p == 2 ? B : A, xi2 = p == 3 ? B : A;
    const double Ns[4] = {1 - xi0 - xi1 - xi2, xi0, xi1, xi2};
    double c[N - 3];
#pragma unroll
    for (int i = 0; i < N - 3; ++i) {
      double v = MM::mul(rec[0][3 + i], Ns[0]);
#pragma unroll
      for (int n = 1; n < 4; ++n) v = MM::add(v, MM::mul(rec[n][3 + i], Ns[n]));
      c[i] = v;
    }
    M3 Q;
    if (KIND == MAG_KIND_ISO) {
      MM::identity(Q);
      const double ih = MM::div(1.0, c[0]);
      Q.m[0][0] = ih; Q.m[1][1] = ih; Q.m[2][2] = ih;
    } else if (KIND == MAG_KIND_ANISO) {
      MM::transform_aniso(V3{c[3], c[4], c[5]}, V3{c[6], c[7], c[8]}, c[0], c[1], c[2], Q);
    } else {
      M3 L;
#pragma unroll
      for (int i = 0; i < 9; ++i) L.m[i / 3][i % 3] = c[i];
      if (MM::transform_logm(L, Q) != 1) *eig_fail = 1;
    }
    M3 JQ;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        JQ.m[i][j] = MM::add(MM::add(MM::mul(J.m[i][0], Q.m[0][j]), MM::mul(J.m[i][1], Q.m[1][j])), MM::mul(J.m[i][2], Q.m[2][j]));
    const double wdv = MM::mul(0.25 / 6.0, MM::det3(JQ));
    measurement = p == 0 ? wdv : MM::add(measurement, wdv);
  }
  return MM::div(measurement, 1.0 / 6.0);   // parentMeasure[TET]
}

// MAG_FP_FAST: det(J Q) = det(J) det(Q), and det(Q) needs no frame at all: Gram-Schmidt yields a rotation (maSize.cc:116,
// 138), so det Q = 1 / (h0 h1 h2) for AnisoSizeField, exp(trace(logM) / 2) for LogAnisoSizeField (the eigenvalues of the
// interpolated logM sum to its trace, which interpolates linearly) and 1 / h^3 for an isotropic field:
//     weight = 0.25 det(J) sum_p det Q(xi_p)
// Only the chunks holding the positions and the sizes / the diagonal of logM are gathered.
template <int KIND>
__device__ __forceinline__ double tet_weight_fast(const double* __restrict__ vedge, const int4& tv)
{
  if (KIND == MAG_KIND_IDENTITY) return 1.0;
  constexpr int K = (KIND == MAG_KIND_ISO) ? 2 : 6;
  const int32_t vid[4] = {tv.x & kVidMask, tv.y, tv.z, tv.w};
  double x[4][3], s[4][3];
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const double2* p = chunk_ptr<K>(vedge, 0, vid[n]);
    const double2 c0 = __ldg(p), c1 = __ldg(p + kVB);
    x[n][0] = c0.x; x[n][1] = c0.y; x[n][2] = c1.x;
    s[n][0] = c1.y;                                   // iso: s; aniso: h0; logm: M00
    if (KIND == MAG_KIND_ANISO) { const double2 c2 = __ldg(p + 2 * kVB); s[n][1] = c2.x; s[n][2] = c2.y; }
    if (KIND == MAG_KIND_LOGM) { s[n][0] += __ldg(p + 3 * kVB).y + __ldg(p + 5 * kVB).y; }   // trace: M00 + M11 + M22
  }
  double e[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int c = 0; c < 3; ++c) e[i][c] = x[i + 1][c] - x[0][c];
  const double detJ = e[0][0] * (e[1][1] * e[2][2] - e[1][2] * e[2][1]) - e[0][1] * (e[1][0] * e[2][2] - e[1][2] * e[2][0]) +
                      e[0][2] * (e[1][0] * e[2][1] - e[1][1] * e[2][0]);
  constexpr double A = 0.138196601125011, B = 0.585410196624969;
  double sum = 0.0;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const double xi0 = p == 1 ? B : A, xi1 = p == 2 ? B : A, xi2 = p == 3 ? B : A;
    const double Ns[4] = {1 - xi0 - xi1 - xi2, xi0, xi1, xi2};
    constexpr int NC = KIND == MAG_KIND_ANISO ? 3 : 1;
    double c[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) c[i] = s[0][i] * Ns[0] + s[1][i] * Ns[1] + s[2][i] * Ns[2] + s[3][i] * Ns[3];
    if (KIND == MAG_KIND_ISO) sum += 1.0 / (c[0] * c[0] * c[0]);
    else if (KIND == MAG_KIND_ANISO) sum += 1.0 / (c[0] * c[1] * c[2]);
    else sum += exp(0.5 * c[0]);
  }
  return 0.25 * detJ * sum;
}

template <int KIND, bool FAST>
__global__ void __launch_bounds__(FAST ? 256 : kWThreads)
k_tet_weights(int32_t nt, int32_t elem_off, const int4* __restrict__ tet_v, const double* __restrict__ vedge,
              double w_max, double w_min, double* __restrict__ weight, MagDevStats* st)
{
  int eig = 0;
  constexpr int T = FAST ? 256 : kWThreads;
  for (int32_t t = blockIdx.x * T + threadIdx.x; t < nt; t += gridDim.x * T) {
    double w = FAST ? tet_weight_fast<KIND>(vedge, __ldg(tet_v + t)) : tet_weight<KIND, StrictOps>(vedge, __ldg(tet_v + t), &eig);
    // clamp of maBalance.cc:14-19
    if (w > w_max) w = w_max;
    else if (w < w_min) w = w_min;
    weight[elem_off + t] = w;
  }
  if (eig) atomicAdd(&st->n_eigen_aux, 1ull);
}

// measure(triangle) on a 2-D part: TriangleIntegration::N2 (apf/apfIntegrate.cc:146-159), three points of weight 1/3/2;
// dV2 = |row0(J Q) x row1(J Q)| (getJacobianDeterminant(., 2), apf/apfVectorElement.cc:75-84); triangle shape values
// apf/apfShape.cc:141-160; parentMeasure[TRIANGLE] = 1/2 (ma/maSize.cc:150).  MAG_FP_FAST runs the same steps with
// contracted arithmetic (a 2-D part is never large enough for the gathers to matter).
template <int KIND, class OPS>
__device__ __forceinline__ double tri_weight(const double* __restrict__ vedge, const int32_t* __restrict__ tv, int* eig_fail)
{
  typedef MagMath<OPS> MM;
  if (KIND == MAG_KIND_IDENTITY) return 1.0;
  constexpr int N = (KIND == MAG_KIND_ISO) ? 4 : 12;
  double rec[3][N];
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    const int32_t v = __ldg(tv + n) & kVidMask;
    if (KIND == MAG_KIND_ISO) load_rec4(vedge, v, rec[n]);
    else {
      Rec12 r = load_rec12(vedge, v);
#pragma unroll
      for (int i = 0; i < 12; ++i) rec[n][i] = r.v[i];
    }
  }
  double J[2][3];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int c = 0; c < 3; ++c) J[i][c] = MM::add(-rec[0][c], rec[i + 1][c]);
  constexpr double A = 0.666666666666667, B = 0.166666666666667;
  double measurement = 0.0;
#pragma unroll 1
  for (int p = 0; p < 3; ++p) {
    const double xi0 = p == 0 ? A : B, xi1 = p == 1 ? A : B;
    const double Ns[3] = {1 - xi0 - xi1, xi0, xi1};
    double c[N - 3];
#pragma unroll
    for (int i = 0; i < N - 3; ++i) {
      double v = MM::mul(rec[0][3 + i], Ns[0]);
#pragma unroll
      for (int n = 1; n < 3; ++n) v = MM::add(v, MM::mul(rec[n][3 + i], Ns[n]));
      c[i] = v;
    }
    M3 Q;
    if (KIND == MAG_KIND_ISO) {
      MM::identity(Q);
      const double ih = MM::div(1.0, c[0]);
      Q.m[0][0] = ih; Q.m[1][1] = ih; Q.m[2][2] = ih;
    } else if (KIND == MAG_KIND_ANISO) {
      MM::transform_aniso(V3{c[3], c[4], c[5]}, V3{c[6], c[7], c[8]}, c[0], c[1], c[2], Q);
    } else {
      M3 L;
#pragma unroll
      for (int i = 0; i < 9; ++i) L.m[i / 3][i % 3] = c[i];
      if (MM::transform_logm(L, Q) != 1) *eig_fail = 1;
    }
    V3 jq[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      jq[i].x = MM::add(MM::add(MM::mul(J[i][0], Q.m[0][0]), MM::mul(J[i][1], Q.m[1][0])), MM::mul(J[i][2], Q.m[2][0]));
      jq[i].y = MM::add(MM::add(MM::mul(J[i][0], Q.m[0][1]), MM::mul(J[i][1], Q.m[1][1])), MM::mul(J[i][2], Q.m[2][1]));
      jq[i].z = MM::add(MM::add(MM::mul(J[i][0], Q.m[0][2]), MM::mul(J[i][1], Q.m[1][2])), MM::mul(J[i][2], Q.m[2][2]));
    }
    const double wdv = MM::mul(1. / 3. / 2.0, MM::length(MM::cross(jq[0], jq[1])));
    measurement = p == 0 ? wdv : MM::add(measurement, wdv);
  }
  return MM::div(measurement, 1.0 / 2.0);   // parentMeasure[TRIANGLE]
}

template <int KIND, bool FAST>
__global__ void __launch_bounds__(kWThreads)
k_tri_weights(int32_t nt, const int32_t* __restrict__ tri_v, const double* __restrict__ vedge, double w_max, double w_min,
              double* __restrict__ weight, MagDevStats* st)
{
  int eig = 0;
  for (int32_t t = blockIdx.x * kWThreads + threadIdx.x; t < nt; t += gridDim.x * kWThreads) {
    double w = FAST ? tri_weight<KIND, FusedOps>(vedge, tri_v + 3 * (int64_t)t, &eig)
                    : tri_weight<KIND, StrictOps>(vedge, tri_v + 3 * (int64_t)t, &eig);
    if (w > w_max) w = w_max;
    else if (w < w_min) w = w_min;
    weight[t] = w;
  }
  if (eig) atomicAdd(&st->n_eigen_aux, 1ull);
}

// a prism as ma::getElementWeight weighs it (maBalance.cc:21-81): measure(base triangle) / (1/2) -- the triangle walked in the
// FACE's own vertex order --, clampForIterations, clampForLayerPermissions, accountForTets
template <int KIND, bool FAST>
__global__ void __launch_bounds__(kWThreads)
k_prism_weights(int32_t np, const int32_t* __restrict__ base_v, const double* __restrict__ vedge, double w_max, double w_min,
                int refine_layer, int coarsen_layer, int to_tets, double* __restrict__ weight, MagDevStats* st)
{
  int eig = 0;
  for (int32_t t = blockIdx.x * kWThreads + threadIdx.x; t < np; t += gridDim.x * kWThreads) {
    double w = FAST ? tri_weight<KIND, FusedOps>(vedge, base_v + 3 * (int64_t)t, &eig)
                    : tri_weight<KIND, StrictOps>(vedge, base_v + 3 * (int64_t)t, &eig);
    if (w > w_max) w = w_max;
    else if (w < w_min) w = w_min;
    if (!refine_layer) w = w > 1.0 ? w : 1.0;      // std::max(1.0, weight)
    if (!coarsen_layer) w = w < 1.0 ? w : 1.0;     // std::min(1.0, weight)
    if (to_tets) w = __dmul_rn(w, 3.0);
    weight[t] = w;
  }
  if (eig) atomicAdd(&st->n_eigen_aux, 1ull);
}

// ------------------------------------------------------------------ sliver classification
// ma::getSliverCode / matchSliver (ma/maShape.cc:35-120), the classification LargeAngleTetFixer runs over the BAD_QUALITY
// tets: J (apfVectorElement.cc:44-52) and Q = getTransform at the centroid, J = J Q; the quality of the tet's FIRST face
// (measureTriQuality with the transform of its largest-determinant vertex -- m->getDimension() is 3, maQuality.cc:86-100 --
// walking the FACE's own vertex order) against goodQuality^2 decides between projecting vertex 3 onto the face and
// projecting edge 0-2 onto edge 0-1.  apf::project apfVector.h:134-137; apf::invert apfMatrix.h:165-173.  Always in the
// reference's operation order (StrictOps): the outputs are bit codes decided by comparisons.
__constant__ signed char c_sliver_table2d[4][4][2] =
  {{{-1,-1}, {-1,-1}, {-1,-1}, {-1,-1}},
   {{ 7, 2}, {-1,-1}, { 3, 3}, {-1,-1}},
   {{ 1, 2}, { 2, 3}, {-1,-1}, {-1,-1}},
   {{ 3, 2}, { 2, 3}, { 3, 3}, {-1,-1}}};
__constant__ signed char c_sliver_table[8][8][2] =
  {{{-1,-1}, {-1,-1}, {-1,-1}, {-1,-1}, {-1,-1}, {-1,-1}, {-1,-1}, {-1,-1}},
   {{ 4, 1}, {-1,-1}, {10, 2}, { 6, 3}, { 4, 2}, { 5, 3}, { 0, 3}, {-1,-1}},
   {{ 1, 1}, { 8, 2}, {-1,-1}, { 6, 3}, { 9, 2}, { 5, 3}, { 0, 3}, {-1,-1}},
   {{ 2, 0}, { 8, 2}, {10, 2}, {-1,-1}, { 0, 2}, { 5, 3}, { 0, 3}, {-1,-1}},
   {{ 2, 1}, {11, 2}, { 2, 2}, { 6, 3}, {-1,-1}, { 5, 3}, { 0, 3}, {-1,-1}},
   {{ 0, 0}, {11, 2}, { 6, 2}, { 6, 3}, { 4, 2}, {-1,-1}, { 0, 3}, {-1,-1}},
   {{ 1, 0}, { 5, 2}, { 2, 2}, { 6, 3}, { 9, 2}, { 5, 3}, {-1,-1}, {-1,-1}},
   {{ 0, 1}, { 5, 2}, { 6, 2}, { 6, 3}, { 0, 2}, { 5, 3}, { 0, 3}, {-1,-1}}};

__device__ __forceinline__ V3 project_onto(const V3& a, const V3& b)
{
  const double s = magst::div(magst::dot(a, b), magst::dot(b, b));
  return V3{magst::mul(b.x, s), magst::mul(b.y, s), magst::mul(b.z, s)};
}
// basisPoint = invert(transpose(J)) * p with J rows j0, j1, j2: rows of the inverse = cross products of J's rows over det(J^T)
__device__ __forceinline__ void area_point(const V3& j0, const V3& j1, const V3& j2, const V3& p, double area[3])
{
  M3 JT;
  JT.m[0][0] = j0.x; JT.m[1][0] = j0.y; JT.m[2][0] = j0.z;
  JT.m[0][1] = j1.x; JT.m[1][1] = j1.y; JT.m[2][1] = j1.z;
  JT.m[0][2] = j2.x; JT.m[1][2] = j2.y; JT.m[2][2] = j2.z;
  const double d = magst::det3(JT);
  const V3 r0 = magst::cross(j1, j2), r1 = magst::cross(j2, j0);
  const double b0 = magst::dot(V3{magst::div(r0.x, d), magst::div(r0.y, d), magst::div(r0.z, d)}, p);
  const double b1 = magst::dot(V3{magst::div(r1.x, d), magst::div(r1.y, d), magst::div(r1.z, d)}, p);
  area[0] = magst::sub(magst::sub(1.0, b0), b1);
  area[1] = b0;
  area[2] = b1;
}

template <int KIND>
__global__ void __launch_bounds__(kWThreads)
k_sliver_codes(int32_t nt, int32_t elem_off, const int4* __restrict__ tet_v, const int32_t* __restrict__ face0_v,
               const double* __restrict__ vedge, const double* __restrict__ vpos, const double* __restrict__ vq,
               const int32_t* __restrict__ flags, int only_bad, double good_quality, int32_t* __restrict__ codes,
               int32_t* __restrict__ match, MagDevStats* st)
{
  int eig = 0;
  for (int32_t t = blockIdx.x * kWThreads + threadIdx.x; t < nt; t += gridDim.x * kWThreads) {
    int code = 0, rot = -1, idx = -1;
    if (!only_bad || (flags[elem_off + t] & MAG_BAD_QUALITY)) {
      const int4 tv = __ldg(tet_v + t);
      const int32_t vid[4] = {tv.x & kVidMask, tv.y, tv.z, tv.w};
      // centroid transform: N = (1 - .25 - .25 - .25, .25, .25, .25)
      constexpr int N = (KIND == MAG_KIND_ISO || KIND == MAG_KIND_IDENTITY) ? 4 : 12;
      double rec[4][N];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        if (N == 4) load_rec4(vedge, vid[n], rec[n]);
        else {
          Rec12 r = load_rec12(vedge, vid[n]);
#pragma unroll
          for (int i = 0; i < 12; ++i) rec[n][i] = r.v[i];
        }
      }
      M3 Q;
      if (KIND == MAG_KIND_IDENTITY) magst::identity(Q);
      else {
        const double Ns[4] = {1 - .25 - .25 - .25, .25, .25, .25};
        double cf[N - 3];
#pragma unroll
        for (int i = 0; i < N - 3; ++i) {
          double v = magst::mul(rec[0][3 + i], Ns[0]);
#pragma unroll
          for (int n = 1; n < 4; ++n) v = magst::add(v, magst::mul(rec[n][3 + i], Ns[n]));
          cf[i] = v;
        }
        if (KIND == MAG_KIND_ISO) {
          magst::identity(Q);
          const double ih = magst::div(1.0, cf[0]);
          Q.m[0][0] = ih; Q.m[1][1] = ih; Q.m[2][2] = ih;
        } else if (KIND == MAG_KIND_ANISO) {
          magst::transform_aniso(V3{cf[3], cf[4], cf[5]}, V3{cf[6], cf[7], cf[8]}, cf[0], cf[1], cf[2], Q);
        } else {
          M3 L;
#pragma unroll
          for (int i = 0; i < 9; ++i) L.m[i / 3][i % 3] = cf[i];
          if (magst::transform_logm(L, Q) != 1) eig = 1;
        }
      }
      V3 j[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double a = magst::add(-rec[0][0], rec[i + 1][0]), b = magst::add(-rec[0][1], rec[i + 1][1]),
                     cz = magst::add(-rec[0][2], rec[i + 1][2]);
        j[i].x = magst::add(magst::add(magst::mul(a, Q.m[0][0]), magst::mul(b, Q.m[1][0])), magst::mul(cz, Q.m[2][0]));
        j[i].y = magst::add(magst::add(magst::mul(a, Q.m[0][1]), magst::mul(b, Q.m[1][1])), magst::mul(cz, Q.m[2][1]));
        j[i].z = magst::add(magst::add(magst::mul(a, Q.m[0][2]), magst::mul(b, Q.m[1][2])), magst::mul(cz, Q.m[2][2]));
      }
      // quality of the first face with the transform of its largest-determinant vertex (strict >, first maximum wins)
      int32_t fv[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) fv[i] = face0_v ? __ldg(face0_v + 3 * (int64_t)t + i) : vid[i];
      V3 fx[3];
      int32_t vb = fv[0];
      double maxJ = -1.0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double p[4];
        load_rec4(vpos, fv[i], p);
        fx[i] = V3{p[0], p[1], p[2]};
        if (p[3] > maxJ) { maxJ = p[3]; vb = fv[i]; }
      }
      M3 Qf;
      {
        const double2 q0 = __ldg(chunk_ptr<5>(vq, 0, vb)), q1 = __ldg(chunk_ptr<5>(vq, 1, vb)), q2 = __ldg(chunk_ptr<5>(vq, 2, vb)),
                      q3 = __ldg(chunk_ptr<5>(vq, 3, vb)), q4 = __ldg(chunk_ptr<5>(vq, 4, vb));
        Qf.m[0][0] = q0.x; Qf.m[0][1] = q0.y; Qf.m[0][2] = q1.x;
        Qf.m[1][0] = q1.y; Qf.m[1][1] = q2.x; Qf.m[1][2] = q2.y;
        Qf.m[2][0] = q3.x; Qf.m[2][1] = q3.y; Qf.m[2][2] = q4.x;
      }
      const double f0 = magst::tri_quality(fx, Qf);
      double area[3];
      if (magst::mul(magst::mul(f0, f0), f0) > magst::mul(good_quality, good_quality)) {
        const V3 v03 = j[2];
        const V3 nrm = magst::cross(j[0], j[1]);
        const V3 pr = project_onto(v03, nrm);
        const V3 projected = V3{magst::sub(v03.x, pr.x), magst::sub(v03.y, pr.y), magst::sub(v03.z, pr.z)};
        area_point(j[0], j[1], nrm, projected, area);
#pragma unroll
        for (int i = 0; i < 3; ++i) if (area[i] > 0) code |= (1 << i);
#pragma unroll
        for (int i = 0; i < 3; ++i) if (area[i] > -0.10 && area[i] < 0.10) code |= ((1 << i) << 3);
        rot = c_sliver_table[code & 7][(code >> 3) & 7][0];
        idx = c_sliver_table[code & 7][(code >> 3) & 7][1];
      } else {
        code |= (1 << 6);
        const V3 projected = project_onto(j[1], j[0]);
        const V3 nrm = magst::cross(j[0], j[1]);
        area_point(j[0], j[1], nrm, projected, area);
#pragma unroll
        for (int i = 0; i < 2; ++i) if (area[i] > 0) code |= ((1 << i) << 7);
#pragma unroll
        for (int i = 0; i < 3; ++i) if (area[i] > -0.20 && area[i] < 0.20) code |= ((1 << i) << 9);
        rot = c_sliver_table2d[(code >> 7) & 3][(code >> 9) & 3][0];
        idx = c_sliver_table2d[(code >> 7) & 3][(code >> 9) & 3][1];
      }
    }
    codes[elem_off + t] = code;
    match[2 * (int64_t)(elem_off + t)] = rot;
    match[2 * (int64_t)(elem_off + t) + 1] = idx;
  }
  if (eig) atomicAdd(&st->n_eigen_aux, 1ull);
}

// ma::clearFlagFromDimension (ma/maAdapt.cc:139-147) / ma::unMarkBadQuality (ma/maShape.cc:138-150) on the resident words
__global__ void __launch_bounds__(256)
k_clear_bits(int64_t n, int32_t mask, int32_t* __restrict__ flags)
{
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int32_t f = flags[i];
    if (f & mask) flags[i] = f & ~mask;
  }
}

// ------------------------------------------------------------------ short-edge classification
// ShortEdgeFixer::shouldApply (ma/maShape.cc:188-219), the sweep fixElementShapes runs right after markBadQuality: for
// every element carrying BAD_QUALITY, the measured lengths of its six edges in getDownward(tet, 1) order; if
// max / min < maximumEdgeRatio the element is not a short-edge case and BAD_QUALITY is cleared, otherwise the FIRST
// shortest edge is the one to remove.  The lengths are the resident results of the last MAG_OP_LENGTHS sweep.
__global__ void __launch_bounds__(256)
k_short_edges(int32_t nt, int32_t elem_off, const int32_t* __restrict__ tet_e, const double* __restrict__ len, double max_ratio,
              int32_t* __restrict__ flags, int32_t* __restrict__ short_edge, unsigned long long* __restrict__ counts)
{
  unsigned cleared = 0, kept = 0;
  const int32_t nloop = (nt + 255) / 256 * 256;
  for (int32_t t = blockIdx.x * 256 + threadIdx.x; t < nloop; t += gridDim.x * 256) {
    if (t < nt) {
      const int32_t f = flags[elem_off + t];
      int32_t pick = -1;
      if (f & MAG_BAD_QUALITY) {
        int32_t e[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) e[i] = __ldg(tet_e + 6 * (int64_t)t + i);
        double lmax = __ldg(len + e[0]), lmin = lmax;
        pick = e[0];
#pragma unroll
        for (int i = 1; i < 6; ++i) {
          const double l = __ldg(len + e[i]);
          if (l > lmax) lmax = l;
          if (l < lmin) { lmin = l; pick = e[i]; }
        }
        if (__ddiv_rn(lmax, lmin) < max_ratio) {
          flags[elem_off + t] = f & ~MAG_BAD_QUALITY;
          pick = -1;
          ++cleared;
        } else ++kept;
      }
      short_edge[elem_off + t] = pick;
    }
  }
  cleared = __reduce_add_sync(0xffffffffu, cleared);
  kept = __reduce_add_sync(0xffffffffu, kept);
  if ((threadIdx.x & 31) == 0) {
    if (cleared) atomicAdd(&counts[0], (unsigned long long)cleared);
    if (kept) atomicAdd(&counts[1], (unsigned long long)kept);
  }
}

template <int KIND>
int launch_weights(mag_ctx* c, double w_max, double w_min, bool fast, double* d_w)
{
  const int T = fast ? 256 : kWThreads;
  const int64_t blocks = (c->nt + T - 1) / T;
  const unsigned g = (unsigned)(blocks < (int64_t)c->n_sms * 32 ? blocks : (int64_t)c->n_sms * 32);
  const int4* tv = reinterpret_cast<const int4*>(c->d_tet_v);
  const int32_t off = (int32_t)(c->np + c->npy);
  if (fast) k_tet_weights<KIND, true><<<g, 256, 0, c->stream>>>((int32_t)c->nt, off, tv, c->d_vedge, w_max, w_min, d_w, c->d_stats);
  else k_tet_weights<KIND, false><<<g, kWThreads, 0, c->stream>>>((int32_t)c->nt, off, tv, c->d_vedge, w_max, w_min, d_w, c->d_stats);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}

template <int KIND>
int launch_tri_weights(mag_ctx* c, double w_max, double w_min, bool fast, double* d_w)
{
  const int64_t blocks = (c->ntri + kWThreads - 1) / kWThreads;
  const unsigned g = (unsigned)(blocks < (int64_t)c->n_sms * 32 ? blocks : (int64_t)c->n_sms * 32);
  if (fast) k_tri_weights<KIND, true><<<g, kWThreads, 0, c->stream>>>((int32_t)c->ntri, c->d_tri_v, c->d_vedge, w_max, w_min, d_w, c->d_stats);
  else k_tri_weights<KIND, false><<<g, kWThreads, 0, c->stream>>>((int32_t)c->ntri, c->d_tri_v, c->d_vedge, w_max, w_min, d_w, c->d_stats);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}

// ------------------------------------------------------------------ split vertices
// ordered compaction of the SPLIT-marked edges: tiles of kSTile edges -> tile counts -> exclusive scan -> scatter
constexpr int kSThreads = 256, kSPasses = 4, kSTile = kSThreads * kSPasses;

__global__ void __launch_bounds__(kSThreads)
k_split_count(int32_t ne, const int32_t* __restrict__ flags, int32_t* __restrict__ tile_count)
{
  __shared__ int sh[kSThreads / 32];
  int n = 0;
  const int32_t e0 = blockIdx.x * kSTile;
#pragma unroll
  for (int p = 0; p < kSPasses; ++p) {
    const int32_t e = e0 + p * kSThreads + threadIdx.x;
    if (e < ne && (flags[e] & MAG_SPLIT)) ++n;
  }
  n = __reduce_add_sync(0xffffffffu, n);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < kSThreads / 32; ++i) t += sh[i];
    tile_count[blockIdx.x] = t;
  }
}
// single block: exclusive scan of the tile counts in place, total to *total
__global__ void __launch_bounds__(1024)
k_split_scan(int32_t ntiles, int32_t* __restrict__ tile_count, long long* __restrict__ total)
{
  __shared__ long long sh[1024];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int32_t base = 0; base < ntiles; base += 1024) {
    const int32_t i = base + threadIdx.x;
    const long long v = i < ntiles ? tile_count[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      long long t = threadIdx.x >= (unsigned)o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < ntiles) tile_count[i] = (int32_t)(carry + sh[threadIdx.x] - v);
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

// makeSplitVert at xi = 0: edge shape values ((1-xi)/2, (1+xi)/2) = (0.5, 0.5) (apfShape.cc:123-124)
template <int KIND, class OPS>
__device__ __forceinline__ void split_vertex(const double* __restrict__ vedge, const double* __restrict__ raw_b, int2 ev, long long k,
                                             double* __restrict__ oxyz, double* __restrict__ oa, double* __restrict__ ob)
{
  typedef MagMath<OPS> MM;
  constexpr double N0 = (1.0 - 0.0) / 2.0, N1 = (1.0 + 0.0) / 2.0;
  ev.x &= kVidMask;
  if (KIND == MAG_KIND_IDENTITY || KIND == MAG_KIND_ISO) {
    double a[4], b[4];
    load_rec4(vedge, ev.x, a);
    load_rec4(vedge, ev.y, b);
#pragma unroll
    for (int i = 0; i < 3; ++i) oxyz[3 * k + i] = MM::lerp2(a[i], N0, b[i], N1);
    if (KIND == MAG_KIND_ISO) oa[k] = MM::lerp2(a[3], N0, b[3], N1);
    return;
  }
  const Rec12 a = load_rec12(vedge, ev.x), b = load_rec12(vedge, ev.y);
#pragma unroll
  for (int i = 0; i < 3; ++i) oxyz[3 * k + i] = MM::lerp2(a.v[i], N0, b.v[i], N1);
  if (KIND == MAG_KIND_LOGM) {
#pragma unroll
    for (int i = 0; i < 9; ++i) ob[9 * k + i] = MM::lerp2(a.v[3 + i], N0, b.v[3 + i], N1);
    return;
  }
  // AnisoSizeField::interpolate: h interpolated, R = orthogonalizeR(interpolated frame) (maSize.cc:414-429, 94-121);
  // column 2 of the interpolated frame is discarded by orthogonalizeR, so columns 0 and 1 of the records suffice
  (void)raw_b;
#pragma unroll
  for (int i = 0; i < 3; ++i) oa[3 * k + i] = MM::lerp2(a.v[3 + i], N0, b.v[3 + i], N1);
  V3 r0{MM::lerp2(a.v[6], N0, b.v[6], N1), MM::lerp2(a.v[7], N0, b.v[7], N1), MM::lerp2(a.v[8], N0, b.v[8], N1)};
  V3 r1{MM::lerp2(a.v[9], N0, b.v[9], N1), MM::lerp2(a.v[10], N0, b.v[10], N1), MM::lerp2(a.v[11], N0, b.v[11], N1)};
  V3 r2;
  MM::gram_schmidt(r0, r1, r2);
  double* R = ob + 9 * k;   // row-major, frame vectors in the columns
  R[0] = r0.x; R[3] = r0.y; R[6] = r0.z;
  R[1] = r1.x; R[4] = r1.y; R[7] = r1.z;
  R[2] = r2.x; R[5] = r2.y; R[8] = r2.z;
}

template <int KIND, class OPS>
__global__ void __launch_bounds__(kSThreads)
k_split_fill(int32_t ne, const int32_t* __restrict__ flags, const int2* __restrict__ edge_v, const double* __restrict__ vedge,
             const int32_t* __restrict__ tile_off, int32_t* __restrict__ oidx, double* __restrict__ oxyz,
             double* __restrict__ oa, double* __restrict__ ob)
{
  __shared__ int sh[kSThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  long long base = tile_off[blockIdx.x];
  const int32_t e0 = blockIdx.x * kSTile;
  for (int p = 0; p < kSPasses; ++p) {
    const int32_t e = e0 + p * kSThreads + threadIdx.x;
    const bool s = e < ne && (flags[e] & MAG_SPLIT);
    const unsigned m = __ballot_sync(0xffffffffu, s);
    __syncthreads();                       // sh is free again
    if (lane == 0) sh[w] = __popc(m);
    __syncthreads();
    int before = 0, pass_total = 0;
#pragma unroll
    for (int i = 0; i < kSThreads / 32; ++i) { before += i < w ? sh[i] : 0; pass_total += sh[i]; }
    if (s) {
      const long long k = base + before + __popc(m & ((1u << lane) - 1u));
      oidx[k] = e;
      split_vertex<KIND, OPS>(vedge, nullptr, __ldg(edge_v + e), k, oxyz, oa, ob);
    }
    base += pass_total;
  }
}

template <int KIND>
int launch_split_fill(mag_ctx* c, bool fast, unsigned tiles, const int32_t* d_off, int32_t* d_idx, double* d_xyz, double* d_a, double* d_b)
{
  const int2* ev = reinterpret_cast<const int2*>(c->d_edge_v);
  if (fast) k_split_fill<KIND, FusedOps><<<tiles, kSThreads, 0, c->stream>>>((int32_t)c->ne, c->d_edge_flags, ev, c->d_vedge, d_off, d_idx, d_xyz, d_a, d_b);
  else k_split_fill<KIND, StrictOps><<<tiles, kSThreads, 0, c->stream>>>((int32_t)c->ne, c->d_edge_flags, ev, c->d_vedge, d_off, d_idx, d_xyz, d_a, d_b);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}

struct DevBuf {   // frees on scope exit
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
};

} // namespace

extern "C" {

int mag_element_weights(mag_ctx* c, double w_max, double w_min, int fp_mode, double* out)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  if (c->kind == MAG_KIND_NONE) return mag_fail(c, MAG_ERR_ARG, "mag_element_weights: no size field set");
  if (fp_mode != MAG_FP_STRICT && fp_mode != MAG_FP_FAST) return mag_fail(c, MAG_ERR_ARG, "mag_element_weights: bad fp_mode %d", fp_mode);
  const int64_t nel = c->np + c->npy + c->nt + c->ntri;
  if (nel == 0) return MAG_OK;
  if (!c->d_weight) MAG_CUDA(c, cudaMalloc((void**)&c->d_weight, (size_t)nel * sizeof(double)));
  // layer elements: the reference weighs a prism by its base triangle (maBalance.cc:31-37), which needs the face's own
  // vertex order; they get 0 here and stay with the reference
  if (c->np + c->npy) MAG_CUDA(c, cudaMemsetAsync(c->d_weight, 0, (size_t)(c->np + c->npy) * sizeof(double), c->stream));
  MAG_CUDA(c, cudaMemsetAsync(&c->d_stats->n_eigen_aux, 0, sizeof(unsigned long long), c->stream));
  if (c->nt) {
    const bool fast = fp_mode == MAG_FP_FAST;
    int rc;
    switch (c->kind) {
      case MAG_KIND_IDENTITY: rc = launch_weights<MAG_KIND_IDENTITY>(c, w_max, w_min, fast, c->d_weight); break;
      case MAG_KIND_ISO: rc = launch_weights<MAG_KIND_ISO>(c, w_max, w_min, fast, c->d_weight); break;
      case MAG_KIND_ANISO: rc = launch_weights<MAG_KIND_ANISO>(c, w_max, w_min, fast, c->d_weight); break;
      default: rc = launch_weights<MAG_KIND_LOGM>(c, w_max, w_min, fast, c->d_weight); break;
    }
    if (rc) return rc;
  }
  if (c->ntri) {
    const bool fast = fp_mode == MAG_FP_FAST;
    int rc;
    switch (c->kind) {
      case MAG_KIND_IDENTITY: rc = launch_tri_weights<MAG_KIND_IDENTITY>(c, w_max, w_min, fast, c->d_weight); break;
      case MAG_KIND_ISO: rc = launch_tri_weights<MAG_KIND_ISO>(c, w_max, w_min, fast, c->d_weight); break;
      case MAG_KIND_ANISO: rc = launch_tri_weights<MAG_KIND_ANISO>(c, w_max, w_min, fast, c->d_weight); break;
      default: rc = launch_tri_weights<MAG_KIND_LOGM>(c, w_max, w_min, fast, c->d_weight); break;
    }
    if (rc) return rc;
  }
  if (out) MAG_CUDA(c, cudaMemcpyAsync(out, c->d_weight, (size_t)nel * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaMemcpyAsync(&c->h_stats->n_eigen_aux, &c->d_stats->n_eigen_aux, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->h_stats->n_eigen_aux)
    return mag_fail(c, MAG_ERR_EIGEN, "eigenQR failed in %llu blocks of the weight sweep (apf::eigen asserts convergence, apfMatrix.cc:76)", c->h_stats->n_eigen_aux);
  return MAG_OK;
}

int mag_prism_weights(mag_ctx* c, const int32_t* base_v, double w_max, double w_min, int should_refine_layer, int should_coarsen_layer,
                      int should_turn_layer_to_tets, int fp_mode, double* out)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  if (c->kind == MAG_KIND_NONE) return mag_fail(c, MAG_ERR_ARG, "mag_prism_weights: no size field set");
  if (fp_mode != MAG_FP_STRICT && fp_mode != MAG_FP_FAST) return mag_fail(c, MAG_ERR_ARG, "mag_prism_weights: bad fp_mode %d", fp_mode);
  if (c->np == 0) return MAG_OK;
  if (!base_v) return mag_fail(c, MAG_ERR_ARG, "mag_prism_weights: no base triangles");
  for (int64_t i = 0; i < 3 * c->np; ++i)
    if (base_v[i] < 0 || base_v[i] >= c->nv) return mag_fail(c, MAG_ERR_ARG, "mag_prism_weights: vertex id %d out of range", base_v[i]);
  const int64_t nel = c->np + c->npy + c->nt + c->ntri;
  if (!c->d_weight) {
    MAG_CUDA(c, cudaMalloc((void**)&c->d_weight, (size_t)nel * sizeof(double)));
    MAG_CUDA(c, cudaMemsetAsync(c->d_weight, 0, (size_t)nel * sizeof(double), c->stream));
  }
  DevBuf d_base;
  MAG_CUDA(c, cudaMalloc(&d_base.p, (size_t)c->np * 3 * sizeof(int32_t)));
  MAG_CUDA(c, cudaMemcpyAsync(d_base.p, base_v, (size_t)c->np * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  MAG_CUDA(c, cudaMemsetAsync(&c->d_stats->n_eigen_aux, 0, sizeof(unsigned long long), c->stream));
  const int64_t blocks = (c->np + kWThreads - 1) / kWThreads;
  const unsigned g = (unsigned)(blocks < (int64_t)c->n_sms * 32 ? blocks : (int64_t)c->n_sms * 32);
  const bool fast = fp_mode == MAG_FP_FAST;
#define MAG_PW(K)                                                                                                                    \
  do {                                                                                                                               \
    if (fast) k_prism_weights<K, true><<<g, kWThreads, 0, c->stream>>>((int32_t)c->np, (const int32_t*)d_base.p, c->d_vedge, w_max, w_min, should_refine_layer, \
                                                                       should_coarsen_layer, should_turn_layer_to_tets, c->d_weight, c->d_stats); \
    else k_prism_weights<K, false><<<g, kWThreads, 0, c->stream>>>((int32_t)c->np, (const int32_t*)d_base.p, c->d_vedge, w_max, w_min, should_refine_layer,     \
                                                                   should_coarsen_layer, should_turn_layer_to_tets, c->d_weight, c->d_stats);  \
  } while (0)
  switch (c->kind) {
    case MAG_KIND_IDENTITY: MAG_PW(MAG_KIND_IDENTITY); break;
    case MAG_KIND_ISO: MAG_PW(MAG_KIND_ISO); break;
    case MAG_KIND_ANISO: MAG_PW(MAG_KIND_ANISO); break;
    default: MAG_PW(MAG_KIND_LOGM); break;
  }
#undef MAG_PW
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  if (out) MAG_CUDA(c, cudaMemcpyAsync(out, c->d_weight, (size_t)c->np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaMemcpyAsync(&c->h_stats->n_eigen_aux, &c->d_stats->n_eigen_aux, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->h_stats->n_eigen_aux)
    return mag_fail(c, MAG_ERR_EIGEN, "eigenQR failed in %llu blocks of the prism weight sweep (apf::eigen asserts convergence, apfMatrix.cc:76)", c->h_stats->n_eigen_aux);
  return MAG_OK;
}

int mag_cavity_quality(mag_ctx* c, int64_t ncav, const int64_t* offsets, const int32_t* tet_v, int use_max_metric, int fp_mode,
                       double* worst, double* qualities)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  if (c->kind == MAG_KIND_NONE) return mag_fail(c, MAG_ERR_ARG, "mag_cavity_quality: no size field set");
  if (c->dim != 3) return mag_fail(c, MAG_ERR_ARG, "mag_cavity_quality: 3-D parts only");
  if (fp_mode != MAG_FP_STRICT && fp_mode != MAG_FP_FAST) return mag_fail(c, MAG_ERR_ARG, "mag_cavity_quality: bad fp_mode %d", fp_mode);
  if (ncav < 0 || (ncav && (!offsets || !worst))) return mag_fail(c, MAG_ERR_ARG, "mag_cavity_quality: bad arguments");
  if (ncav == 0) return MAG_OK;
  const int64_t ntet = offsets[ncav];
  if (offsets[0] != 0 || ntet < ncav || (ntet && !tet_v)) return mag_fail(c, MAG_ERR_ARG, "mag_cavity_quality: bad offsets");
  for (int64_t k = 0; k < ncav; ++k)   // getWorstQuality asserts n > 0 (maQuality.cc:186)
    if (offsets[k + 1] <= offsets[k]) return mag_fail(c, MAG_ERR_ARG, "mag_cavity_quality: cavity %lld is empty", (long long)k);
  for (int64_t i = 0; i < 4 * ntet; ++i)
    if (tet_v[i] < 0 || tet_v[i] >= c->nv) return mag_fail(c, MAG_ERR_ARG, "mag_cavity_quality: vertex id %d out of range", tet_v[i]);
  int rc;
  if (!c->vertex_pass_valid) {
    if ((rc = magk_vertex_pass(c))) return rc;
    c->vertex_pass_valid = true;
  }
  DevBuf off, tv, dw, dq;
  MAG_CUDA(c, cudaMalloc(&off.p, (size_t)(ncav + 1) * 8));
  MAG_CUDA(c, cudaMalloc(&tv.p, (size_t)ntet * 16));
  MAG_CUDA(c, cudaMalloc(&dw.p, (size_t)ncav * 8));
  if (qualities) MAG_CUDA(c, cudaMalloc(&dq.p, (size_t)ntet * 8));
  MAG_CUDA(c, cudaMemcpyAsync(off.p, offsets, (size_t)(ncav + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  MAG_CUDA(c, cudaMemcpyAsync(tv.p, tet_v, (size_t)ntet * 16, cudaMemcpyHostToDevice, c->stream));
  MAG_CUDA(c, cudaMemsetAsync(&c->d_stats->n_eigen_aux, 0, sizeof(unsigned long long), c->stream));
  if ((rc = magk_cavity_quality(c, fp_mode, ncav, (const int64_t*)off.p, (const int32_t*)tv.p, use_max_metric, (double*)dw.p, (double*)dq.p))) return rc;
  MAG_CUDA(c, cudaMemcpyAsync(worst, dw.p, (size_t)ncav * 8, cudaMemcpyDeviceToHost, c->stream));
  if (qualities) MAG_CUDA(c, cudaMemcpyAsync(qualities, dq.p, (size_t)ntet * 8, cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaMemcpyAsync(&c->h_stats->n_eigen_aux, &c->d_stats->n_eigen_aux, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->h_stats->n_eigen_aux)
    return mag_fail(c, MAG_ERR_EIGEN, "eigenQR failed in %llu blocks of the cavity sweep (apf::eigen asserts convergence, apfMatrix.cc:76)", c->h_stats->n_eigen_aux);
  return MAG_OK;
}

int mag_sliver_codes(mag_ctx* c, const int32_t* face0_v, double good_quality, int only_bad, int32_t* codes, int32_t* match)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  if (c->kind == MAG_KIND_NONE) return mag_fail(c, MAG_ERR_ARG, "mag_sliver_codes: no size field set");
  if (c->dim != 3) return mag_fail(c, MAG_ERR_ARG, "mag_sliver_codes: 3-D parts only");
  if (!codes || !match) return mag_fail(c, MAG_ERR_ARG, "mag_sliver_codes: null output");
  if (face0_v)
    for (int64_t i = 0; i < 3 * c->nt; ++i)
      if (face0_v[i] < 0 || face0_v[i] >= c->nv) return mag_fail(c, MAG_ERR_ARG, "mag_sliver_codes: vertex id %d out of range", face0_v[i]);
  const int64_t nel = c->np + c->npy + c->nt;
  if (nel == 0) return MAG_OK;
  int rc;
  if (only_bad && (rc = magi_materialize_flags(c))) return rc;
  if (!c->vertex_pass_valid) {
    if ((rc = magk_vertex_pass(c))) return rc;
    c->vertex_pass_valid = true;
  }
  DevBuf f0, dc, dm;
  if (face0_v && c->nt) {
    MAG_CUDA(c, cudaMalloc(&f0.p, (size_t)c->nt * 12));
    MAG_CUDA(c, cudaMemcpyAsync(f0.p, face0_v, (size_t)c->nt * 12, cudaMemcpyHostToDevice, c->stream));
  }
  MAG_CUDA(c, cudaMalloc(&dc.p, (size_t)nel * 4));
  MAG_CUDA(c, cudaMalloc(&dm.p, (size_t)nel * 8));
  // layer elements: code 0, no match
  MAG_CUDA(c, cudaMemsetAsync(dc.p, 0, (size_t)nel * 4, c->stream));
  MAG_CUDA(c, cudaMemsetAsync(dm.p, 0xff, (size_t)nel * 8, c->stream));
  MAG_CUDA(c, cudaMemsetAsync(&c->d_stats->n_eigen_aux, 0, sizeof(unsigned long long), c->stream));
  if (c->nt) {
    const int64_t blocks = (c->nt + kWThreads - 1) / kWThreads;
    const unsigned g = (unsigned)(blocks < (int64_t)c->n_sms * 32 ? blocks : (int64_t)c->n_sms * 32);
    const int4* tv = reinterpret_cast<const int4*>(c->d_tet_v);
    const int32_t off = (int32_t)(c->np + c->npy);
#define MAG_SLIVER(K) k_sliver_codes<K><<<g, kWThreads, 0, c->stream>>>((int32_t)c->nt, off, tv, (const int32_t*)f0.p, c->d_vedge, \
      c->d_vpos, c->d_vq, c->d_elem_flags, only_bad, good_quality, (int32_t*)dc.p, (int32_t*)dm.p, c->d_stats)
    switch (c->kind) {
      case MAG_KIND_IDENTITY: MAG_SLIVER(MAG_KIND_IDENTITY); break;
      case MAG_KIND_ISO: MAG_SLIVER(MAG_KIND_ISO); break;
      case MAG_KIND_ANISO: MAG_SLIVER(MAG_KIND_ANISO); break;
      default: MAG_SLIVER(MAG_KIND_LOGM); break;
    }
#undef MAG_SLIVER
    MAG_CUDA(c, cudaGetLastError());
    c->n_launches++;
  }
  MAG_CUDA(c, cudaMemcpyAsync(codes, dc.p, (size_t)nel * 4, cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaMemcpyAsync(match, dm.p, (size_t)nel * 8, cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaMemcpyAsync(&c->h_stats->n_eigen_aux, &c->d_stats->n_eigen_aux, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->h_stats->n_eigen_aux)
    return mag_fail(c, MAG_ERR_EIGEN, "eigenQR failed in %llu blocks of the sliver sweep (apf::eigen asserts convergence, apfMatrix.cc:76)", c->h_stats->n_eigen_aux);
  return MAG_OK;
}

int mag_clear_flag(mag_ctx* c, int dimension, int32_t flag)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  if (dimension != 1 && dimension != c->dim) return mag_fail(c, MAG_ERR_ARG, "mag_clear_flag: dimension %d holds no resident flag words (edges = 1, elements = %d)", dimension, c->dim);
  const bool edges = dimension == 1;
  if (edges ? c->edge_flags_zero : c->elem_flags_zero) return MAG_OK;   // all words are zero: nothing to clear
  const int64_t n = edges ? c->ne : c->np + c->npy + c->nt + c->ntri;
  if (n == 0) return MAG_OK;
  const int64_t blocks = (n + 255) / 256;
  const unsigned g = (unsigned)(blocks < (int64_t)c->n_sms * 16 ? blocks : (int64_t)c->n_sms * 16);
  k_clear_bits<<<g, 256, 0, c->stream>>>(n, flag, edges ? c->d_edge_flags : c->d_elem_flags);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}

int mag_collapse_quality(mag_ctx* c, int64_t ncand, const int32_t* edges, const uint8_t* which_end, int use_max_metric, int fp_mode,
                         double* new_worst, double* old_worst, int32_t* n_keep)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  if (c->kind == MAG_KIND_NONE) return mag_fail(c, MAG_ERR_ARG, "mag_collapse_quality: no size field set");
  if (c->dim != 3) return mag_fail(c, MAG_ERR_ARG, "mag_collapse_quality: 3-D parts only");
  if (fp_mode != MAG_FP_STRICT && fp_mode != MAG_FP_FAST) return mag_fail(c, MAG_ERR_ARG, "mag_collapse_quality: bad fp_mode %d", fp_mode);
  if (ncand < 0 || (ncand && (!edges || !which_end || !new_worst || !old_worst))) return mag_fail(c, MAG_ERR_ARG, "mag_collapse_quality: bad arguments");
  if (ncand == 0) return MAG_OK;
  for (int64_t k = 0; k < ncand; ++k)
    if (edges[k] < 0 || edges[k] >= c->ne) return mag_fail(c, MAG_ERR_ARG, "mag_collapse_quality: edge %d out of range", edges[k]);
  int rc;
  if (!c->schedule_valid) return mag_fail(c, MAG_ERR_ARG, "mag_collapse_quality: the part came through mag_sweep_host; call mag_set_mesh");
  if (!c->vertex_pass_valid) {
    if ((rc = magk_vertex_pass(c))) return rc;
    c->vertex_pass_valid = true;
  }
  if ((rc = magk_build_v2t(c))) return rc;
  DevBuf de, dn, dnew, dold, dkeep;
  MAG_CUDA(c, cudaMalloc(&de.p, (size_t)ncand * 4));
  MAG_CUDA(c, cudaMalloc(&dn.p, (size_t)ncand));
  MAG_CUDA(c, cudaMalloc(&dnew.p, (size_t)ncand * 8));
  MAG_CUDA(c, cudaMalloc(&dold.p, (size_t)ncand * 8));
  MAG_CUDA(c, cudaMalloc(&dkeep.p, (size_t)ncand * 4));
  MAG_CUDA(c, cudaMemcpyAsync(de.p, edges, (size_t)ncand * 4, cudaMemcpyHostToDevice, c->stream));
  MAG_CUDA(c, cudaMemcpyAsync(dn.p, which_end, (size_t)ncand, cudaMemcpyHostToDevice, c->stream));
  MAG_CUDA(c, cudaMemsetAsync(&c->d_stats->n_eigen_aux, 0, sizeof(unsigned long long), c->stream));
  if ((rc = magk_collapse_quality(c, fp_mode, ncand, (const int32_t*)de.p, (const uint8_t*)dn.p, use_max_metric, (double*)dnew.p,
                                  (double*)dold.p, (int32_t*)dkeep.p)))
    return rc;
  MAG_CUDA(c, cudaMemcpyAsync(new_worst, dnew.p, (size_t)ncand * 8, cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaMemcpyAsync(old_worst, dold.p, (size_t)ncand * 8, cudaMemcpyDeviceToHost, c->stream));
  if (n_keep) MAG_CUDA(c, cudaMemcpyAsync(n_keep, dkeep.p, (size_t)ncand * 4, cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaMemcpyAsync(&c->h_stats->n_eigen_aux, &c->d_stats->n_eigen_aux, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->h_stats->n_eigen_aux)
    return mag_fail(c, MAG_ERR_EIGEN, "eigenQR failed in %llu blocks of the collapse sweep (apf::eigen asserts convergence, apfMatrix.cc:76)", c->h_stats->n_eigen_aux);
  return MAG_OK;
}

int mag_short_edge_test(mag_ctx* c, const int32_t* tet_edges, double max_edge_ratio, int32_t* short_edge,
                        int64_t* n_cleared, int64_t* n_short)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  if (c->dim != 3) return mag_fail(c, MAG_ERR_ARG, "mag_short_edge_test: 3-D parts only");
  if (c->nt && !tet_edges) return mag_fail(c, MAG_ERR_ARG, "mag_short_edge_test: null tet_edges");
  if (!(c->last_ops & MAG_OP_LENGTHS)) return mag_fail(c, MAG_ERR_ARG, "mag_short_edge_test: the last sweep did not measure the edges (MAG_OP_LENGTHS)");
  for (int64_t i = 0; i < 6 * c->nt; ++i)
    if (tet_edges[i] < 0 || tet_edges[i] >= c->ne) return mag_fail(c, MAG_ERR_ARG, "mag_short_edge_test: edge index %d out of range", tet_edges[i]);
  int rc;
  if ((rc = magi_materialize_flags(c))) return rc;
  const int64_t nel = c->np + c->npy + c->nt;
  if (n_cleared) *n_cleared = 0;
  if (n_short) *n_short = 0;
  if (c->nt == 0) return MAG_OK;
  DevBuf te, se, cnt;
  MAG_CUDA(c, cudaMalloc(&te.p, (size_t)c->nt * 24));
  MAG_CUDA(c, cudaMalloc(&se.p, (size_t)nel * 4));
  MAG_CUDA(c, cudaMalloc(&cnt.p, 16));
  MAG_CUDA(c, cudaMemcpyAsync(te.p, tet_edges, (size_t)c->nt * 24, cudaMemcpyHostToDevice, c->stream));
  MAG_CUDA(c, cudaMemsetAsync(se.p, 0xff, (size_t)nel * 4, c->stream));      // -1 on layer elements
  MAG_CUDA(c, cudaMemsetAsync(cnt.p, 0, 16, c->stream));
  const int64_t blocks = (c->nt + 255) / 256;
  const unsigned g = (unsigned)(blocks < (int64_t)c->n_sms * 16 ? blocks : (int64_t)c->n_sms * 16);
  k_short_edges<<<g, 256, 0, c->stream>>>((int32_t)c->nt, (int32_t)(c->np + c->npy), (const int32_t*)te.p, c->d_len, max_edge_ratio,
                                          c->d_elem_flags, (int32_t*)se.p, (unsigned long long*)cnt.p);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  unsigned long long h[2] = {0, 0};
  if (short_edge) MAG_CUDA(c, cudaMemcpyAsync(short_edge, se.p, (size_t)nel * 4, cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaMemcpyAsync(h, cnt.p, 16, cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (n_cleared) *n_cleared = (int64_t)h[0];
  if (n_short) *n_short = (int64_t)h[1];
  return MAG_OK;
}

int mag_split_vertices(mag_ctx* c, int fp_mode, int64_t cap, int64_t* n, int32_t* edge_idx, double* xyz, double* field_a, double* field_b)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  if (!n) return mag_fail(c, MAG_ERR_ARG, "mag_split_vertices: null count pointer");
  if (c->kind == MAG_KIND_NONE) return mag_fail(c, MAG_ERR_ARG, "mag_split_vertices: no size field set");
  if (fp_mode != MAG_FP_STRICT && fp_mode != MAG_FP_FAST) return mag_fail(c, MAG_ERR_ARG, "mag_split_vertices: bad fp_mode %d", fp_mode);
  *n = 0;
  if (c->ne == 0) return MAG_OK;
  const unsigned tiles = (unsigned)((c->ne + kSTile - 1) / kSTile);
  { int rc = magi_materialize_flags(c); if (rc) return rc; }
  DevBuf off, tot, didx, dxyz, da, db;
  MAG_CUDA(c, cudaMalloc(&off.p, (size_t)tiles * 4));
  MAG_CUDA(c, cudaMalloc(&tot.p, 8));
  k_split_count<<<tiles, kSThreads, 0, c->stream>>>((int32_t)c->ne, c->d_edge_flags, (int32_t*)off.p);
  k_split_scan<<<1, 1024, 0, c->stream>>>((int32_t)tiles, (int32_t*)off.p, (long long*)tot.p);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches += 2;
  long long total = 0;
  MAG_CUDA(c, cudaMemcpyAsync(&total, tot.p, 8, cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  *n = total;
  if (total == 0 || (!edge_idx && !xyz && !field_a && !field_b)) return MAG_OK;   // count only
  if (cap < total) return mag_fail(c, MAG_ERR_ARG, "mag_split_vertices: %lld edges carry SPLIT, buffers hold %lld", total, (long long)cap);
  const size_t na = c->kind == MAG_KIND_ISO ? 1 : (c->kind == MAG_KIND_ANISO ? 3 : 0);
  const size_t nb = (c->kind == MAG_KIND_ANISO || c->kind == MAG_KIND_LOGM) ? 9 : 0;
  MAG_CUDA(c, cudaMalloc(&didx.p, (size_t)total * 4));
  MAG_CUDA(c, cudaMalloc(&dxyz.p, (size_t)total * 24));
  if (na) MAG_CUDA(c, cudaMalloc(&da.p, (size_t)total * na * 8));
  if (nb) MAG_CUDA(c, cudaMalloc(&db.p, (size_t)total * nb * 8));
  const bool fast = fp_mode == MAG_FP_FAST;
  int rc;
  switch (c->kind) {
    case MAG_KIND_IDENTITY: rc = launch_split_fill<MAG_KIND_IDENTITY>(c, fast, tiles, (int32_t*)off.p, (int32_t*)didx.p, (double*)dxyz.p, (double*)da.p, (double*)db.p); break;
    case MAG_KIND_ISO: rc = launch_split_fill<MAG_KIND_ISO>(c, fast, tiles, (int32_t*)off.p, (int32_t*)didx.p, (double*)dxyz.p, (double*)da.p, (double*)db.p); break;
    case MAG_KIND_ANISO: rc = launch_split_fill<MAG_KIND_ANISO>(c, fast, tiles, (int32_t*)off.p, (int32_t*)didx.p, (double*)dxyz.p, (double*)da.p, (double*)db.p); break;
    default: rc = launch_split_fill<MAG_KIND_LOGM>(c, fast, tiles, (int32_t*)off.p, (int32_t*)didx.p, (double*)dxyz.p, (double*)da.p, (double*)db.p); break;
  }
  if (rc) return rc;
  if (edge_idx) MAG_CUDA(c, cudaMemcpyAsync(edge_idx, didx.p, (size_t)total * 4, cudaMemcpyDeviceToHost, c->stream));
  if (xyz) MAG_CUDA(c, cudaMemcpyAsync(xyz, dxyz.p, (size_t)total * 24, cudaMemcpyDeviceToHost, c->stream));
  if (field_a && na) MAG_CUDA(c, cudaMemcpyAsync(field_a, da.p, (size_t)total * na * 8, cudaMemcpyDeviceToHost, c->stream));
  if (field_b && nb) MAG_CUDA(c, cudaMemcpyAsync(field_b, db.p, (size_t)total * nb * 8, cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  return MAG_OK;
}

} // extern "C"
