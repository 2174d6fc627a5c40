// mag_kernels.cu -- sm_100a kernels of the MeshAdapt marking / quality sweep.
//
// One thread per entity, one launch per entity dimension:
//   k_pack_*        raw uploads -> 32-byte-sector-aligned per-vertex gather records
//   k_vertex_pass   per-vertex transform Q_v and det Q_v   (hoists getMetricWithMaxJacobean's
//                   4 getTransform calls per tet, ma/maQuality.cc:83-108, to one per vertex)
//   k_edges         metric length (2-point Gauss), SPLIT / COLLAPSE flags, owned counts, max / sum
//   k_tets          mean-ratio quality, BAD_QUALITY flags, owned count, min
//   k_layer         prism / pyramid validity
//   k_fix_*         strict re-evaluation of the entities the fast kernels found within 1e-12 of a threshold
// Warp-shuffle + shared-memory block reductions feed one atomic per block.
#include "mag_internal.h"
#include "mag_math.cuh"
#include "mag_math_fast.cuh"
#include <cstring>

namespace {

constexpr int kThreads = 256;

// order-preserving map double -> uint64 (for atomicMin on possibly negative qualities)
__host__ __device__ inline unsigned long long dkey(double d)
{
  unsigned long long b;
#ifdef __CUDA_ARCH__
  b = (unsigned long long)__double_as_longlong(d);
#else
  memcpy(&b, &d, 8);
#endif
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_max(unsigned long long v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { unsigned long long w = __shfl_down_sync(0xffffffffu, v, o); v = w > v ? w : v; }
  return v;
}
__device__ __forceinline__ unsigned long long warp_min(unsigned long long v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { unsigned long long w = __shfl_down_sync(0xffffffffu, v, o); v = w < v ? w : v; }
  return v;
}

// packs up to 4 small counters into one u64 (16 bits each, a block has <= 1024 threads)
__device__ __forceinline__ void block_count4(unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                             unsigned long long* g0, unsigned long long* g1,
                                             unsigned long long* g2, unsigned long long* g3)
{
  __shared__ unsigned long long sh[kThreads / 32];
  unsigned long long p = (unsigned long long)c0 | ((unsigned long long)c1 << 16) |
                         ((unsigned long long)c2 << 32) | ((unsigned long long)c3 << 48);
  p = warp_sum(p);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = p;
  __syncthreads();
  if (w == 0) {
    p = lane < (kThreads / 32) ? sh[lane] : 0ull;
    p = warp_sum(p);
    if (lane == 0) {
      unsigned a = (unsigned)(p & 0xffff), b = (unsigned)((p >> 16) & 0xffff);
      unsigned cc = (unsigned)((p >> 32) & 0xffff), d = (unsigned)((p >> 48) & 0xffff);
      if (a) atomicAdd(g0, (unsigned long long)a);
      if (b) atomicAdd(g1, (unsigned long long)b);
      if (cc) atomicAdd(g2, (unsigned long long)cc);
      if (d) atomicAdd(g3, (unsigned long long)d);
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void block_max_u64(unsigned long long v, unsigned long long* g)
{
  __shared__ unsigned long long sh[kThreads / 32];
  v = warp_max(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < (kThreads / 32) ? sh[lane] : 0ull;
    v = warp_max(v);
    if (lane == 0 && v) atomicMax(g, v);
  }
  __syncthreads();
}
__device__ __forceinline__ void block_min_u64(unsigned long long v, unsigned long long* g)
{
  __shared__ unsigned long long sh[kThreads / 32];
  v = warp_min(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < (kThreads / 32) ? sh[lane] : ~0ull;
    v = warp_min(v);
    if (lane == 0 && v != ~0ull) atomicMin(g, v);
  }
  __syncthreads();
}
// deterministic per-block partial sum -> d_block_sums[blockIdx.x]
__device__ __forceinline__ void block_sum_f64(double v, double* out)
{
  __shared__ double sh[kThreads / 32];
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < (kThreads / 32) ? sh[lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0) *out = v;
  }
  __syncthreads();
}

// ------------------------------------------------------------------ record loads
struct Rec12 { double v[12]; };
__device__ __forceinline__ Rec12 load_rec12(const double* __restrict__ base, int32_t vid)
{
  const double2* p = reinterpret_cast<const double2*>(base + 12 * (size_t)vid);
  Rec12 r;
#pragma unroll
  for (int i = 0; i < 6; ++i) { double2 t = __ldg(p + i); r.v[2 * i] = t.x; r.v[2 * i + 1] = t.y; }
  return r;
}
__device__ __forceinline__ void load_rec4(const double* __restrict__ base, int32_t vid, double out[4])
{
  const double2* p = reinterpret_cast<const double2*>(base + 4 * (size_t)vid);
  double2 a = __ldg(p), b = __ldg(p + 1);
  out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
}

// ------------------------------------------------------------------ pack kernels
__global__ void k_pack4(int64_t nv, const double* __restrict__ xyz, const double* __restrict__ s, double* __restrict__ rec)
{
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= nv) return;
  double2* o = reinterpret_cast<double2*>(rec + 4 * v);
  o[0] = make_double2(xyz[3 * v], xyz[3 * v + 1]);
  o[1] = make_double2(xyz[3 * v + 2], s ? s[v] : 0.0);
}
// aniso: {x,y,z,h0,h1,h2,R00,R10,R20,R01,R11,R21} (frame columns 0 and 1; column 2 is
// overwritten by orthogonalizeR before use, maSize.cc:94-121)
__global__ void k_pack12_aniso(int64_t nv, const double* __restrict__ xyz, const double* __restrict__ h,
                               const double* __restrict__ R, double* __restrict__ rec)
{
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= nv) return;
  double* o = rec + 12 * v;
  const double* r = R + 9 * v;
  o[0] = xyz[3 * v]; o[1] = xyz[3 * v + 1]; o[2] = xyz[3 * v + 2];
  o[3] = h[3 * v]; o[4] = h[3 * v + 1]; o[5] = h[3 * v + 2];
  o[6] = r[0]; o[7] = r[3]; o[8] = r[6];
  o[9] = r[1]; o[10] = r[4]; o[11] = r[7];
}
__global__ void k_pack12_logm(int64_t nv, const double* __restrict__ xyz, const double* __restrict__ M, double* __restrict__ rec)
{
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= nv) return;
  double* o = rec + 12 * v;
  o[0] = xyz[3 * v]; o[1] = xyz[3 * v + 1]; o[2] = xyz[3 * v + 2];
#pragma unroll
  for (int i = 0; i < 9; ++i) o[3 + i] = M[9 * v + i];
}

// ------------------------------------------------------------------ per-vertex pass
// Q_v = SizeField::getTransform(vertex, xi=0) and det Q_v (apf::getJacobianDeterminant(Q,3)).
// The vertex shape value is exactly 1.0 so "interpolation" returns the node value.
template <int KIND>
__global__ void k_vertex_pass(int64_t nv, const double* __restrict__ vedge, double* __restrict__ vpos,
                              double* __restrict__ vq, MagDevStats* st)
{
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= nv) return;
  M3 Q;
  double x, y, z;
  if (KIND == MAG_KIND_IDENTITY || KIND == MAG_KIND_ISO) {
    double r[4];
    load_rec4(vedge, (int32_t)v, r);
    x = r[0]; y = r[1]; z = r[2];
    magst::identity(Q);
    if (KIND == MAG_KIND_ISO) {
      // generic path with R = I, h = (s,s,s): Gram-Schmidt of I is I exactly, Q = diag(1/s)
      double ih = magst::div(1.0, r[3]);
      Q.m[0][0] = ih; Q.m[1][1] = ih; Q.m[2][2] = ih;
    }
  } else {
    Rec12 r = load_rec12(vedge, (int32_t)v);
    x = r.v[0]; y = r.v[1]; z = r.v[2];
    if (KIND == MAG_KIND_ANISO) {
      magst::transform_aniso(V3{r.v[6], r.v[7], r.v[8]}, V3{r.v[9], r.v[10], r.v[11]}, r.v[3], r.v[4], r.v[5], Q);
    } else {
      M3 A;
#pragma unroll
      for (int i = 0; i < 9; ++i) A.m[i / 3][i % 3] = r.v[3 + i];
      int rc = magst::transform_logm(A, Q);
      if (rc != 1) atomicAdd(&st->n_eigen_fail, 1ull);
    }
  }
  double det = magst::det3(Q);
  double2* o = reinterpret_cast<double2*>(vpos + 4 * v);
  o[0] = make_double2(x, y);
  o[1] = make_double2(z, det);
  double2* q = reinterpret_cast<double2*>(vq + 10 * v);
  q[0] = make_double2(Q.m[0][0], Q.m[0][1]);
  q[1] = make_double2(Q.m[0][2], Q.m[1][0]);
  q[2] = make_double2(Q.m[1][1], Q.m[1][2]);
  q[3] = make_double2(Q.m[2][0], Q.m[2][1]);
  q[4] = make_double2(Q.m[2][2], det);
}

// ------------------------------------------------------------------ edge metric length (strict)
// MetricSizeField::measure: order 2 -> EdgeIntegration::N2, points +-0.577350269189626, weights 1
template <int KIND>
__device__ __forceinline__ double edge_length_strict(const double* __restrict__ vedge, int32_t a, int32_t b, int* eig_fail)
{
  constexpr double XI = 0.577350269189626;
  // shape values exactly as apfShape.cc:123-124 computes them
  constexpr double NP0 = (1.0 - XI) / 2.0, NP1 = (1.0 + XI) / 2.0;       // point 0: xi = +XI
  constexpr double NM0 = (1.0 - (-XI)) / 2.0, NM1 = (1.0 + (-XI)) / 2.0; // point 1: xi = -XI
  if (KIND == MAG_KIND_IDENTITY || KIND == MAG_KIND_ISO) {
    double ra[4], rb[4];
    load_rec4(vedge, a, ra);
    load_rec4(vedge, b, rb);
    V3 j = magst::edge_j0(V3{ra[0], ra[1], ra[2]}, V3{rb[0], rb[1], rb[2]});
    if (KIND == MAG_KIND_IDENTITY) return magst::mul(2.0, magst::length(j)); // apf::measure, N1 rule
    double len[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      double h = magst::lerp2(ra[3], p ? NM0 : NP0, rb[3], p ? NM1 : NP1);
      double ih = magst::div(1.0, h);
      V3 r{magst::mul(j.x, ih), magst::mul(j.y, ih), magst::mul(j.z, ih)};
      len[p] = magst::length(r);
    }
    return magst::add(len[0], len[1]);
  } else {
    Rec12 ra = load_rec12(vedge, a), rb = load_rec12(vedge, b);
    V3 j = magst::edge_j0(V3{ra.v[0], ra.v[1], ra.v[2]}, V3{rb.v[0], rb.v[1], rb.v[2]});
    double len[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const double n0 = p ? NM0 : NP0, n1 = p ? NM1 : NP1;
      double c[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) c[i] = magst::lerp2(ra.v[3 + i], n0, rb.v[3 + i], n1);
      M3 Q;
      if (KIND == MAG_KIND_ANISO) {
        magst::transform_aniso(V3{c[3], c[4], c[5]}, V3{c[6], c[7], c[8]}, c[0], c[1], c[2], Q);
      } else {
        M3 A;
#pragma unroll
        for (int i = 0; i < 9; ++i) A.m[i / 3][i % 3] = c[i];
        if (magst::transform_logm(A, Q) != 1) *eig_fail = 1;
      }
      len[p] = magst::row0_length(j, Q);
    }
    return magst::add(len[0], len[1]);
  }
}

template <int KIND>
__device__ __forceinline__ double edge_length_fast(const double* __restrict__ vedge, int32_t a, int32_t b, int* eig_fail)
{
  if (KIND == MAG_KIND_IDENTITY || KIND == MAG_KIND_ISO) {
    double ra[4], rb[4];
    load_rec4(vedge, a, ra);
    load_rec4(vedge, b, rb);
    if (KIND == MAG_KIND_IDENTITY) return magfa::edge_identity(ra, rb);
    return magfa::edge_iso(ra, rb);
  } else {
    Rec12 ra = load_rec12(vedge, a), rb = load_rec12(vedge, b);
    if (KIND == MAG_KIND_ANISO) return magfa::edge_aniso(ra.v, rb.v);
    return magfa::edge_logm(ra.v, rb.v, eig_fail);
  }
}

struct SweepParams {
  uint32_t ops;
  double max_len, min_len, good_q;
  int use_max;
};

__device__ __forceinline__ bool near_thr(double v, double thr)
{
  return fabs(thr) <= 1.79e308 && fabs(v - thr) <= MAG_NEAR_REL * fabs(thr);
}

// shared tail of k_edges and k_fix_edges: flag update + per-thread counters
__device__ __forceinline__ void mark_edge(double len, int32_t& f, bool need_split, bool need_coll, bool owned,
                                          const SweepParams& P, unsigned& c_split, unsigned& c_coll)
{
  if (need_split) {
    if (len > P.max_len) { f |= MAG_SPLIT; if (owned) ++c_split; }
    else f |= MAG_NEED_NOT_SPLIT;
  }
  if (need_coll) {
    if (len < P.min_len) { f |= MAG_COLLAPSE; if (owned) ++c_coll; }
    else f |= MAG_NEED_NOT_COLLAPSE;
  }
}

template <int KIND, bool FAST>
__global__ void __launch_bounds__(kThreads)
k_edges(int64_t ne, const int2* __restrict__ edge_v, const double* __restrict__ vedge,
        const uint8_t* __restrict__ owned_arr, int32_t* __restrict__ flags, double* __restrict__ lengths,
        SweepParams P, MagDevStats* st, double* __restrict__ block_sums, int64_t* __restrict__ near_list)
{
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  unsigned c_split = 0, c_coll = 0, c_eval = 0, c_err = 0;
  unsigned long long maxbits = 0;
  double sum = 0;
  if (e < ne) {
    int32_t f = flags[e];
    const int32_t f_in = f;
    const bool do_split = P.ops & MAG_OP_MARK_SPLIT, do_coll = P.ops & MAG_OP_MARK_COLLAPSE;
    // markEntities asserts the true flag is clear on every entity it visits (maAdapt.cc:308)
    if ((do_split && (f & MAG_SPLIT)) || (do_coll && (f & MAG_COLLAPSE))) ++c_err;
    const bool need_split = do_split && !(f & (MAG_DONT_SPLIT | MAG_NEED_NOT_SPLIT));
    const bool need_coll = do_coll && !(f & (MAG_DONT_COLLAPSE | MAG_NEED_NOT_COLLAPSE));
    const bool want_len = P.ops & MAG_OP_LENGTHS;
    if (want_len || need_split || need_coll) {
      int2 ev = __ldg(edge_v + e);
      int eig = 0;
      double len = FAST ? edge_length_fast<KIND>(vedge, ev.x, ev.y, &eig)
                        : edge_length_strict<KIND>(vedge, ev.x, ev.y, &eig);
      if (eig) atomicAdd(&st->n_eigen_fail, 1ull);
      const bool owned = owned_arr ? (owned_arr[e] != 0) : true;
      if (want_len) {
        lengths[e] = len;
        if (owned) { maxbits = (unsigned long long)__double_as_longlong(len > 0 ? len : 0.0); sum = len; }
      }
      bool nr = (need_split && near_thr(len, P.max_len)) || (need_coll && near_thr(len, P.min_len));
      if (nr) {
        unsigned long long k = atomicAdd(&st->n_near_edge, 1ull);
        if (k < MAG_NEAR_CAP) near_list[k] = e;
      }
      if (FAST && nr) {
        flags[e] = f | MAG_PENDING_BIT; // k_fix_edges re-evaluates it in strict arithmetic
      } else if (need_split || need_coll) {
        ++c_eval;
        mark_edge(len, f, need_split, need_coll, owned, P, c_split, c_coll);
        if (f != f_in) flags[e] = f;
      }
    }
  }
  block_count4(c_split, c_coll, c_eval, c_err, &st->n_split, &st->n_collapse, &st->n_edges_eval, &st->n_flag_err);
  if (P.ops & MAG_OP_LENGTHS) {
    block_max_u64(maxbits, &st->max_len_bits);
    block_sum_f64(sum, block_sums + blockIdx.x);
  }
}

// strict re-evaluation of the edges a FAST sweep found within 1e-12 of a threshold
// (walks the recorded list, or -- if more than MAG_NEAR_CAP entities were near -- every entity
// carrying the pending bit)
template <int KIND>
__global__ void __launch_bounds__(kThreads)
k_fix_edges(int64_t ne, const int2* __restrict__ edge_v, const double* __restrict__ vedge, const uint8_t* __restrict__ owned_arr,
            int32_t* __restrict__ flags, SweepParams P, MagDevStats* st, const int64_t* __restrict__ near_list)
{
  const unsigned long long nn = st->n_near_edge;
  const bool listed = nn <= MAG_NEAR_CAP;
  const unsigned long long n = listed ? nn : (unsigned long long)ne;
  unsigned c_split = 0, c_coll = 0, c_eval = 0;
  for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < n;
       k += (unsigned long long)gridDim.x * blockDim.x) {
    int64_t e = listed ? near_list[k] : (int64_t)k;
    int32_t f = flags[e];
    if (!(f & MAG_PENDING_BIT)) continue;
    f &= ~MAG_PENDING_BIT;
    const bool need_split = (P.ops & MAG_OP_MARK_SPLIT) && !(f & (MAG_DONT_SPLIT | MAG_NEED_NOT_SPLIT));
    const bool need_coll = (P.ops & MAG_OP_MARK_COLLAPSE) && !(f & (MAG_DONT_COLLAPSE | MAG_NEED_NOT_COLLAPSE));
    int2 ev = __ldg(edge_v + e);
    int eig = 0;
    double len = edge_length_strict<KIND>(vedge, ev.x, ev.y, &eig);
    const bool owned = owned_arr ? (owned_arr[e] != 0) : true;
    ++c_eval;
    mark_edge(len, f, need_split, need_coll, owned, P, c_split, c_coll);
    flags[e] = f;
  }
  block_count4(c_split, c_coll, c_eval, 0, &st->n_split, &st->n_collapse, &st->n_edges_eval, &st->n_flag_err);
}

// ------------------------------------------------------------------ tets
// centroid transform for useMax == false (maQuality.cc:148-153): N = (1-.25-.25-.25, .25, .25, .25)
template <int KIND>
__device__ __forceinline__ void centroid_transform(const double* __restrict__ vedge, const int4& tv, M3& Q, int* eig)
{
  constexpr double N0 = 1 - 0.25 - 0.25 - 0.25;
  const int32_t vid[4] = {tv.x, tv.y, tv.z, tv.w};
  if (KIND == MAG_KIND_IDENTITY) { magst::identity(Q); return; }
  if (KIND == MAG_KIND_ISO) {
    double h = 0;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      double r[4];
      load_rec4(vedge, vid[n], r);
      h = magst::add(h, magst::mul(r[3], n ? 0.25 : N0));
    }
    magst::identity(Q);
    double ih = magst::div(1.0, h);
    Q.m[0][0] = ih; Q.m[1][1] = ih; Q.m[2][2] = ih;
    return;
  }
  double c[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) c[i] = 0;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    Rec12 r = load_rec12(vedge, vid[n]);
#pragma unroll
    for (int i = 0; i < 9; ++i) c[i] = magst::add(c[i], magst::mul(r.v[3 + i], n ? 0.25 : N0));
  }
  if (KIND == MAG_KIND_ANISO) {
    magst::transform_aniso(V3{c[3], c[4], c[5]}, V3{c[6], c[7], c[8]}, c[0], c[1], c[2], Q);
  } else {
    M3 A;
#pragma unroll
    for (int i = 0; i < 9; ++i) A.m[i / 3][i % 3] = c[i];
    if (magst::transform_logm(A, Q) != 1) *eig = 1;
  }
}

template <int KIND, bool FAST>
__device__ __forceinline__ double tet_quality_eval(const int4& tv, const double* __restrict__ vpos,
                                                   const double* __restrict__ vq, const double* __restrict__ vedge,
                                                   int use_max, int* eig)
{
  double p[4][4];
  load_rec4(vpos, tv.x, p[0]);
  load_rec4(vpos, tv.y, p[1]);
  load_rec4(vpos, tv.z, p[2]);
  load_rec4(vpos, tv.w, p[3]);
  M3 Q;
  double detQ;
  if (use_max) {
    // getMetricWithMaxJacobean: strict >, first maximum wins (maQuality.cc:97-104)
    int best = 0;
    double maxJ = -1.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (p[i][3] > maxJ) { maxJ = p[i][3]; best = i; }
    int32_t vb = best == 0 ? tv.x : best == 1 ? tv.y : best == 2 ? tv.z : tv.w;
    const double2* q = reinterpret_cast<const double2*>(vq + 10 * (size_t)vb);
    double2 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3), q4 = __ldg(q + 4);
    Q.m[0][0] = q0.x; Q.m[0][1] = q0.y; Q.m[0][2] = q1.x;
    Q.m[1][0] = q1.y; Q.m[1][1] = q2.x; Q.m[1][2] = q2.y;
    Q.m[2][0] = q3.x; Q.m[2][1] = q3.y; Q.m[2][2] = q4.x;
    detQ = q4.y;
  } else {
    centroid_transform<KIND>(vedge, tv, Q, eig);
    detQ = FAST ? magst::det3(Q) : 0.0;
  }
  V3 x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = V3{p[i][0], p[i][1], p[i][2]};
  return FAST ? magfa::tet_quality(x, Q, detQ) : magst::tet_quality(x, Q);
}

__device__ __forceinline__ void mark_tet(double q, int32_t& f, bool owned, const SweepParams& P, unsigned& c_bad)
{
  if (q < P.good_q) { f |= MAG_BAD_QUALITY; if (owned) ++c_bad; }
  else f |= MAG_OK_QUALITY;
}

template <int KIND, bool FAST>
__global__ void __launch_bounds__(kThreads)
k_tets(int64_t nt, int64_t elem_off, const int4* __restrict__ tet_v, const double* __restrict__ vpos,
       const double* __restrict__ vq, const double* __restrict__ vedge, const uint8_t* __restrict__ owned_arr,
       int32_t* __restrict__ flags, double* __restrict__ qual, SweepParams P, MagDevStats* st,
       int64_t* __restrict__ near_list)
{
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  unsigned c_bad = 0, c_eval = 0, c_err = 0;
  unsigned long long minkey = ~0ull;
  if (t < nt) {
    const int64_t el = elem_off + t;
    int32_t f = flags[el];
    const int32_t f_in = f;
    const bool do_bad = P.ops & MAG_OP_MARK_BAD;
    if (do_bad && (f & MAG_BAD_QUALITY)) ++c_err;
    const bool need_bad = do_bad && !(f & MAG_OK_QUALITY);
    const bool want_q = P.ops & MAG_OP_QUALITIES;
    if (want_q || need_bad) {
      int4 tv = __ldg(tet_v + t);
      int eig = 0;
      double q = tet_quality_eval<KIND, FAST>(tv, vpos, vq, vedge, P.use_max, &eig);
      if (eig) atomicAdd(&st->n_eigen_fail, 1ull);
      if (want_q) { qual[el] = q; minkey = dkey(q); }
      const bool owned = owned_arr ? (owned_arr[el] != 0) : true;
      bool nr = need_bad && near_thr(q, P.good_q);
      if (nr) {
        unsigned long long k = atomicAdd(&st->n_near_elem, 1ull);
        if (k < MAG_NEAR_CAP) near_list[k] = el;
      }
      if (FAST && nr) {
        flags[el] = f | MAG_PENDING_BIT;
      } else if (need_bad) {
        ++c_eval;
        mark_tet(q, f, owned, P, c_bad);
        if (f != f_in) flags[el] = f;
      }
    }
  }
  block_count4(c_bad, c_eval, c_err, 0, &st->n_bad, &st->n_elems_eval, &st->n_flag_err, &st->n_flag_err);
  if (P.ops & MAG_OP_QUALITIES) block_min_u64(minkey, &st->min_q_key);
}

template <int KIND>
__global__ void __launch_bounds__(kThreads)
k_fix_tets(int64_t nt, int64_t elem_off, const int4* __restrict__ tet_v, const double* __restrict__ vpos,
           const double* __restrict__ vq, const double* __restrict__ vedge, const uint8_t* __restrict__ owned_arr,
           int32_t* __restrict__ flags, SweepParams P, MagDevStats* st, const int64_t* __restrict__ near_list)
{
  const unsigned long long nn = st->n_near_elem;
  const bool listed = nn <= MAG_NEAR_CAP;
  const unsigned long long n = listed ? nn : (unsigned long long)nt;
  unsigned c_bad = 0, c_eval = 0;
  for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < n;
       k += (unsigned long long)gridDim.x * blockDim.x) {
    int64_t el = listed ? near_list[k] : elem_off + (int64_t)k;
    int32_t f = flags[el];
    if (!(f & MAG_PENDING_BIT)) continue;
    f &= ~MAG_PENDING_BIT;
    int4 tv = __ldg(tet_v + (el - elem_off));
    int eig = 0;
    double q = tet_quality_eval<KIND, false>(tv, vpos, vq, vedge, P.use_max, &eig);
    const bool owned = owned_arr ? (owned_arr[el] != 0) : true;
    ++c_eval;
    mark_tet(q, f, owned, P, c_bad);
    flags[el] = f;
  }
  block_count4(c_bad, c_eval, 0, 0, &st->n_bad, &st->n_elems_eval, &st->n_flag_err, &st->n_flag_err);
}

// ------------------------------------------------------------------ prisms / pyramids
// a non-simplex element that reaches markBadQuality without OK_QUALITY would make the
// reference call a null table entry (maQuality.cc:169-182): report instead of crash.
__global__ void k_nonsimplex_guard(int64_t n, const int32_t* __restrict__ flags, SweepParams P, MagDevStats* st)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t f = flags[i];
  if (f & MAG_BAD_QUALITY) atomicAdd(&st->n_flag_err, 1ull);
  if (!(f & MAG_OK_QUALITY)) atomicAdd(&st->n_nonsimplex, 1ull);
}

struct Plane { V3 n; double r; };
// apf::Plane::fromPoints + Plane ctor (apfGeometry.cc:28-40): the constructor normalizes the
// already-unit normal again and scales the radius by its length
__device__ __forceinline__ Plane plane_from_points(const V3& a, const V3& b, const V3& c)
{
  V3 u{magst::sub(a.x, c.x), magst::sub(a.y, c.y), magst::sub(a.z, c.z)};
  V3 w{magst::sub(b.x, c.x), magst::sub(b.y, c.y), magst::sub(b.z, c.z)};
  V3 n = magst::normalize(magst::cross(u, w));
  double radius = magst::dot(c, n);
  double l = magst::length(n);
  Plane p;
  p.n = V3{magst::div(n.x, l), magst::div(n.y, l), magst::div(n.z, l)};
  p.r = magst::mul(radius, l);
  return p;
}
__device__ __forceinline__ double plane_distance(const Plane& p, const V3& x)
{
  return magst::sub(magst::dot(p.n, x), p.r);
}
__constant__ int c_prism_rotation[6][6] = {   // ma/maTables.cc:204-210
  {0, 1, 2, 3, 4, 5}, {1, 2, 0, 4, 5, 3}, {2, 0, 1, 5, 3, 4},
  {3, 5, 4, 0, 2, 1}, {4, 3, 5, 1, 0, 2}, {5, 4, 3, 2, 1, 0}};
__constant__ int c_pyramid_rotation[2][5] = {{0, 1, 2, 3, 4}, {1, 2, 3, 0, 4}}; // ma/maTables.cc:212-216 (first two)
__constant__ int c_shift_table[6] = {0, 1, 2, 2, 0, 1};                         // ma/maQuality.cc:482
__device__ __forceinline__ int unrotate_code(int code, int rot)
{
  int out = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (code & (1 << i)) out |= (1 << ((i + c_shift_table[rot]) % 3));
  return out;
}
__device__ __forceinline__ V3 load_pos(const double* __restrict__ vpos, int32_t v)
{
  double r[4];
  load_rec4(vpos, v, r);
  return V3{r[0], r[1], r[2]};
}
// isPrismOk (maQuality.cc:490-530) / isPyramidOk (:532-560)
__global__ void k_layer(int64_t np, int64_t npy, const int32_t* __restrict__ prism_v, const int32_t* __restrict__ pyr_v,
                        const double* __restrict__ vpos, int32_t* __restrict__ ok, int32_t* __restrict__ codes, MagDevStats* st)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= np + npy) return;
  int good = 1, code;
  if (i < np) {
    V3 p[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) p[k] = load_pos(vpos, prism_v[6 * i + k]);
    code = 0xFF;
    for (int r = 0; r < 6; ++r) {
      const int* n2o = c_prism_rotation[r];
      Plane pl = plane_from_points(p[n2o[0]], p[n2o[1]], p[n2o[5]]);
      if (plane_distance(pl, p[n2o[3]]) <= 0) { good = 0; code &= ~(1 << unrotate_code(5, r)); }
      if (plane_distance(pl, p[n2o[4]]) <= 0) { good = 0; code &= ~(1 << unrotate_code(4, r)); }
      if (plane_distance(pl, p[n2o[2]]) >= 0) {
        good = 0;
        code &= ~(1 << unrotate_code(5, r));
        code &= ~(1 << unrotate_code(4, r));
      }
    }
  } else {
    int64_t j = i - np;
    V3 p[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) p[k] = load_pos(vpos, pyr_v[5 * j + k]);
    code = -1;
    for (int r = 0; r < 2; ++r) {
      const int* n2o = c_pyramid_rotation[r];
      Plane pl = plane_from_points(p[n2o[0]], p[n2o[2]], p[n2o[4]]);
      if (plane_distance(pl, p[n2o[1]]) <= 0) { good = 0; continue; }
      if (plane_distance(pl, p[n2o[3]]) >= 0) { good = 0; continue; }
      code = r;
    }
  }
  ok[i] = good;
  codes[i] = code;
  if (!good) atomicAdd(&st->n_layer_unsafe, 1ull);
}

// ------------------------------------------------------------------ misc
__global__ void k_init_stats(MagDevStats* st)
{
  MagDevStats z;
  memset(&z, 0, sizeof(z));
  z.max_len_bits = 0;                 // bits of +0.0 : getMaximumEdgeLength starts at 0.0
  z.min_q_key = dkey(1.0);            // getMinQuality starts at 1
  *st = z;
}
// fixed-order tree sum of the per-block partials (deterministic for a given grid)
__global__ void k_finish_sum(int64_t n, const double* __restrict__ part, MagDevStats* st)
{
  __shared__ double sh[kThreads];
  double s = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) st->sum_len = sh[0];
}

inline unsigned grid_for(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

} // namespace

// ====================================================================== host-side launchers
int magk_pack(mag_ctx* c)
{
  if (c->nv == 0) return MAG_OK;
  unsigned g = grid_for(c->nv);
  switch (c->kind) {
    case MAG_KIND_IDENTITY: k_pack4<<<g, kThreads, 0, c->stream>>>(c->nv, c->d_xyz, nullptr, c->d_vedge); break;
    case MAG_KIND_ISO: k_pack4<<<g, kThreads, 0, c->stream>>>(c->nv, c->d_xyz, c->d_ma, c->d_vedge); break;
    case MAG_KIND_ANISO: k_pack12_aniso<<<g, kThreads, 0, c->stream>>>(c->nv, c->d_xyz, c->d_ma, c->d_mb, c->d_vedge); break;
    case MAG_KIND_LOGM: k_pack12_logm<<<g, kThreads, 0, c->stream>>>(c->nv, c->d_xyz, c->d_mb, c->d_vedge); break;
    default: return mag_fail(c, MAG_ERR_ARG, "no size field set");
  }
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}

int magk_init_stats(mag_ctx* c)
{
  k_init_stats<<<1, 1, 0, c->stream>>>(c->d_stats);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}

int magk_vertex_pass(mag_ctx* c)
{
  if (c->nv == 0) return MAG_OK;
  unsigned g = grid_for(c->nv);
  switch (c->kind) {
    case MAG_KIND_IDENTITY: k_vertex_pass<MAG_KIND_IDENTITY><<<g, kThreads, 0, c->stream>>>(c->nv, c->d_vedge, c->d_vpos, c->d_vq, c->d_stats); break;
    case MAG_KIND_ISO: k_vertex_pass<MAG_KIND_ISO><<<g, kThreads, 0, c->stream>>>(c->nv, c->d_vedge, c->d_vpos, c->d_vq, c->d_stats); break;
    case MAG_KIND_ANISO: k_vertex_pass<MAG_KIND_ANISO><<<g, kThreads, 0, c->stream>>>(c->nv, c->d_vedge, c->d_vpos, c->d_vq, c->d_stats); break;
    case MAG_KIND_LOGM: k_vertex_pass<MAG_KIND_LOGM><<<g, kThreads, 0, c->stream>>>(c->nv, c->d_vedge, c->d_vpos, c->d_vq, c->d_stats); break;
    default: return mag_fail(c, MAG_ERR_ARG, "no size field set");
  }
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}

template <int KIND>
static int launch_edges(mag_ctx* c, const SweepParams& P, bool fast)
{
  unsigned g = grid_for(c->ne);
  const int2* ev = reinterpret_cast<const int2*>(c->d_edge_v);
  if (fast) {
    k_edges<KIND, true><<<g, kThreads, 0, c->stream>>>(c->ne, ev, c->d_vedge, c->d_edge_owned, c->d_edge_flags, c->d_len, P, c->d_stats, c->d_block_sums, c->d_near_edge);
    if (P.ops & (MAG_OP_MARK_SPLIT | MAG_OP_MARK_COLLAPSE))
    { k_fix_edges<KIND><<<148, kThreads, 0, c->stream>>>(c->ne, ev, c->d_vedge, c->d_edge_owned, c->d_edge_flags, P, c->d_stats, c->d_near_edge); c->n_launches++; }
  } else {
    k_edges<KIND, false><<<g, kThreads, 0, c->stream>>>(c->ne, ev, c->d_vedge, c->d_edge_owned, c->d_edge_flags, c->d_len, P, c->d_stats, c->d_block_sums, c->d_near_edge);
  }
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  if (P.ops & MAG_OP_LENGTHS) {
    k_finish_sum<<<1, kThreads, 0, c->stream>>>((int64_t)g, c->d_block_sums, c->d_stats);
    MAG_CUDA(c, cudaGetLastError());
    c->n_launches++;
  }
  return MAG_OK;
}

template <int KIND>
static int launch_tets(mag_ctx* c, const SweepParams& P, bool fast)
{
  unsigned g = grid_for(c->nt);
  const int4* tv = reinterpret_cast<const int4*>(c->d_tet_v);
  const int64_t off = c->np + c->npy;
  if (fast) {
    k_tets<KIND, true><<<g, kThreads, 0, c->stream>>>(c->nt, off, tv, c->d_vpos, c->d_vq, c->d_vedge, c->d_elem_owned, c->d_elem_flags, c->d_qual, P, c->d_stats, c->d_near_elem);
    if (P.ops & MAG_OP_MARK_BAD)
    { k_fix_tets<KIND><<<148, kThreads, 0, c->stream>>>(c->nt, off, tv, c->d_vpos, c->d_vq, c->d_vedge, c->d_elem_owned, c->d_elem_flags, P, c->d_stats, c->d_near_elem); c->n_launches++; }
  } else {
    k_tets<KIND, false><<<g, kThreads, 0, c->stream>>>(c->nt, off, tv, c->d_vpos, c->d_vq, c->d_vedge, c->d_elem_owned, c->d_elem_flags, c->d_qual, P, c->d_stats, c->d_near_elem);
  }
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}

int magk_sweep(mag_ctx* c, uint32_t ops, double max_len, double min_len, double good_q, int use_max, int fp_mode)
{
  SweepParams P{ops, max_len, min_len, good_q, use_max};
  if (c->kind == MAG_KIND_IDENTITY) {
    // IdentitySizeField::shouldSplit / shouldCollapse are constant false (maSize.cc:64-72): no length exceeds +inf
    P.max_len = INFINITY;
    P.min_len = -INFINITY;
  }
  const bool fast = fp_mode == MAG_FP_FAST;
  int rc;
  cudaEvent_t* tev = (c->t_used < c->t_slots) ? &c->tev[(size_t)4 * c->t_used] : nullptr;
  if (tev) MAG_CUDA(c, cudaEventRecord(tev[0], c->stream));
  if (ops & (MAG_OP_QUALITIES | MAG_OP_MARK_BAD | MAG_OP_LAYER_CHECK)) {
    // per-vertex transforms are part of every quality sweep (never cached across sweeps)
    if ((rc = magk_vertex_pass(c))) return rc;
  }
  if (tev) MAG_CUDA(c, cudaEventRecord(tev[1], c->stream));
  if (c->ne && (ops & (MAG_OP_LENGTHS | MAG_OP_MARK_SPLIT | MAG_OP_MARK_COLLAPSE))) {
    switch (c->kind) {
      case MAG_KIND_IDENTITY: rc = launch_edges<MAG_KIND_IDENTITY>(c, P, fast); break;
      case MAG_KIND_ISO: rc = launch_edges<MAG_KIND_ISO>(c, P, fast); break;
      case MAG_KIND_ANISO: rc = launch_edges<MAG_KIND_ANISO>(c, P, fast); break;
      default: rc = launch_edges<MAG_KIND_LOGM>(c, P, fast); break;
    }
    if (rc) return rc;
  }
  if (tev) MAG_CUDA(c, cudaEventRecord(tev[2], c->stream));
  if (ops & (MAG_OP_QUALITIES | MAG_OP_MARK_BAD)) {
    if (c->nt) {
      switch (c->kind) {
        case MAG_KIND_IDENTITY: rc = launch_tets<MAG_KIND_IDENTITY>(c, P, fast); break;
        case MAG_KIND_ISO: rc = launch_tets<MAG_KIND_ISO>(c, P, fast); break;
        case MAG_KIND_ANISO: rc = launch_tets<MAG_KIND_ANISO>(c, P, fast); break;
        default: rc = launch_tets<MAG_KIND_LOGM>(c, P, fast); break;
      }
      if (rc) return rc;
    }
    if ((ops & MAG_OP_MARK_BAD) && (c->np + c->npy)) {
      k_nonsimplex_guard<<<grid_for(c->np + c->npy), kThreads, 0, c->stream>>>(c->np + c->npy, c->d_elem_flags, P, c->d_stats);
      MAG_CUDA(c, cudaGetLastError());
      c->n_launches++;
    }
  }
  if ((ops & MAG_OP_LAYER_CHECK) && (c->np + c->npy)) {
    k_layer<<<grid_for(c->np + c->npy), kThreads, 0, c->stream>>>(c->np, c->npy, c->d_prism_v, c->d_pyr_v, c->d_vpos, c->d_layer_ok, c->d_layer_codes, c->d_stats);
    MAG_CUDA(c, cudaGetLastError());
    c->n_launches++;
  }
  if (tev) { MAG_CUDA(c, cudaEventRecord(tev[3], c->stream)); c->t_used++; }
  return MAG_OK;
}

double magk_key_to_double(unsigned long long k)
{
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double d;
  memcpy(&d, &b, 8);
  return d;
}
