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
#include "mag_layout.cuh"
#include "mag_math.cuh"
#include "mag_math_fast.cuh"
#include <cstring>
#include <cmath>

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

// Streaming accesses of the persistent kernels (connectivity and flag words in, lengths / qualities / flag words out) are
// touched once per sweep; the vertex records are re-read ~14 times.  MAG_STREAM_HINTS=1 marks the former evict-first
// (ld.global.cs / st.global.cs) so they do not displace vertex records in L1 / L2.
#ifndef MAG_STREAM_HINTS
#define MAG_STREAM_HINTS 1   /* measured: tets 0.807 -> 0.794 ms, edges unchanged */
#endif
template <class T> __device__ __forceinline__ T ld_stream(const T* p) { return MAG_STREAM_HINTS ? __ldcs(p) : __ldg(p); }
template <class T> __device__ __forceinline__ T ld_stream_rw(const T* p) { return MAG_STREAM_HINTS ? __ldcs(p) : *p; }
template <class T> __device__ __forceinline__ void st_stream(T* p, T v) { if (MAG_STREAM_HINTS) __stcs(p, v); else *p = v; }

// ------------------------------------------------------------------ reductions
// The edge / tet kernels are persistent (grid-stride over tiles): every thread carries its counters in registers
// for the whole launch and the block reduces ONCE at the end -- warp REDUX, then one atomic per warp and counter.
__device__ __forceinline__ void warp_count_to(unsigned c, unsigned long long* g)
{
  unsigned v = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(g, (unsigned long long)v);
}
// 64-bit max / min through two 32-bit REDUX steps (high word first)
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v)
{
  unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
  unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return ((unsigned long long)mh << 32) | ml;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v)
{
  unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
  unsigned mh = __reduce_min_sync(0xffffffffu, hi);
  unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
  return ((unsigned long long)mh << 32) | ml;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

using namespace maglay;

// ------------------------------------------------------------------ pack kernels
__global__ void k_pack4(int64_t nv, const double* __restrict__ xyz, const double* __restrict__ s, double* __restrict__ rec)
{
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= nv) return;
  *chunk_ptr_w<2>(rec, 0, v) = make_double2(xyz[3 * v], xyz[3 * v + 1]);
  *chunk_ptr_w<2>(rec, 1, v) = make_double2(xyz[3 * v + 2], s ? s[v] : 0.0);
}
__global__ void k_pack12_aniso(int64_t nv, const double* __restrict__ xyz, const double* __restrict__ h,
                               const double* __restrict__ R, double* __restrict__ rec)
{
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const double* r = R + 9 * v;
  *chunk_ptr_w<6>(rec, 0, v) = make_double2(xyz[3 * v], xyz[3 * v + 1]);
  *chunk_ptr_w<6>(rec, 1, v) = make_double2(xyz[3 * v + 2], h[3 * v]);
  *chunk_ptr_w<6>(rec, 2, v) = make_double2(h[3 * v + 1], h[3 * v + 2]);
  *chunk_ptr_w<6>(rec, 3, v) = make_double2(r[0], r[3]);
  *chunk_ptr_w<6>(rec, 4, v) = make_double2(r[6], r[1]);
  *chunk_ptr_w<6>(rec, 5, v) = make_double2(r[4], r[7]);
}
__global__ void k_pack12_logm(int64_t nv, const double* __restrict__ xyz, const double* __restrict__ M, double* __restrict__ rec)
{
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const double* m = M + 9 * v;
  *chunk_ptr_w<6>(rec, 0, v) = make_double2(xyz[3 * v], xyz[3 * v + 1]);
  *chunk_ptr_w<6>(rec, 1, v) = make_double2(xyz[3 * v + 2], m[0]);
  *chunk_ptr_w<6>(rec, 2, v) = make_double2(m[1], m[2]);
  *chunk_ptr_w<6>(rec, 3, v) = make_double2(m[3], m[4]);
  *chunk_ptr_w<6>(rec, 4, v) = make_double2(m[5], m[6]);
  *chunk_ptr_w<6>(rec, 5, v) = make_double2(m[7], m[8]);
}

// sets the sign bit of the first vertex id of every entity this part does not own (stride = vertices per entity)
__global__ void k_fold_owned(int64_t n, int stride, const uint8_t* __restrict__ owned, int32_t* __restrict__ conn)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && !owned[i]) conn[i * stride] |= (int32_t)0x80000000;
}

// ------------------------------------------------------------------ per-vertex pass
// Q_v = SizeField::getTransform(vertex, xi=0) and det Q_v (apf::getJacobianDeterminant(Q,3)).
// The vertex shape value is exactly 1.0 so "interpolation" returns the node value.
template <int KIND>
__device__ __forceinline__ void vertex_pass_one(int64_t v, int dim, const double* __restrict__ vedge, double* __restrict__ vpos,
                                                double* __restrict__ vq, int* eig_fail)
{
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
      if (rc != 1) *eig_fail = 1;
    }
  }
  // apf::getJacobianDeterminant(Q, mesh dimension) (apfVectorElement.cc:68-91): det in 3-D, |row0 x row1| in 2-D
  double det = dim == 3 ? magst::det3(Q)
                        : magst::length(magst::cross(V3{Q.m[0][0], Q.m[0][1], Q.m[0][2]}, V3{Q.m[1][0], Q.m[1][1], Q.m[1][2]}));
  *chunk_ptr_w<2>(vpos, 0, v) = make_double2(x, y);
  *chunk_ptr_w<2>(vpos, 1, v) = make_double2(z, det);
  *chunk_ptr_w<5>(vq, 0, v) = make_double2(Q.m[0][0], Q.m[0][1]);
  *chunk_ptr_w<5>(vq, 1, v) = make_double2(Q.m[0][2], Q.m[1][0]);
  *chunk_ptr_w<5>(vq, 2, v) = make_double2(Q.m[1][1], Q.m[1][2]);
  *chunk_ptr_w<5>(vq, 3, v) = make_double2(Q.m[2][0], Q.m[2][1]);
  *chunk_ptr_w<5>(vq, 4, v) = make_double2(Q.m[2][2], det);
}
template <int KIND>
__global__ void k_vertex_pass(int64_t nv, int dim, const double* __restrict__ vedge, double* __restrict__ vpos,
                              double* __restrict__ vq, unsigned long long* eig_fail)
{
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= nv) return;
  int eig = 0;
  vertex_pass_one<KIND>(v, dim, vedge, vpos, vq, &eig);
  if (eig) atomicAdd(eig_fail, 1ull);
}

// shape values of the edge's two Gauss points exactly as apfShape.cc:123-124 computes them (xi = +-0.577350269189626)
constexpr double kXI = 0.577350269189626;
constexpr double kNP0s = (1.0 - kXI) / 2.0, kNP1s = (1.0 + kXI) / 2.0;       // point 0: xi = +XI
constexpr double kNM0s = (1.0 - (-kXI)) / 2.0, kNM1s = (1.0 + (-kXI)) / 2.0; // point 1: xi = -XI
static_assert(kNM0s == kNP1s && kNM1s == kNP0s, "the two Gauss points swap their shape values");

// Q_u(v): the transform BOTH Gauss points of an edge see when its two ends carry vertex v's size-field values bit for bit (an edge
// along a direction the field does not vary in, any edge of a uniform region): the interpolated values are a N0 + a N1, the same
// sum in either order (edge_length_strict below), so the transform depends on the vertex alone and every such edge around it shares
// it.  Strict arithmetic, the reference's operation order: what edge_length_strict would compute in place, computed once per vertex
// with the per-vertex pass.  Chunk 4 carries the eigen-solver's failure flag of the log-Euclidean field in its second half.
template <int KIND>
__global__ void k_vertex_uniform(int64_t nv, const double* __restrict__ vedge, double* __restrict__ vqu)
{
  static_assert(KIND == MAG_KIND_ANISO || KIND == MAG_KIND_LOGM, "only the frame-carrying fields have a transform worth caching");
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const Rec12 r = load_rec12(vedge, (int32_t)v);
  double c[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) c[i] = magst::lerp2(r.v[3 + i], kNP0s, r.v[3 + i], kNP1s);
  M3 Q;
  double fail = 0.0;
  if (KIND == MAG_KIND_ANISO) {
    magst::transform_aniso(V3{c[3], c[4], c[5]}, V3{c[6], c[7], c[8]}, c[0], c[1], c[2], Q);
  } else {
    M3 A;
#pragma unroll
    for (int i = 0; i < 9; ++i) A.m[i / 3][i % 3] = c[i];
    if (magst::transform_logm(A, Q) != 1) fail = 1.0;
  }
  *chunk_ptr_w<5>(vqu, 0, v) = make_double2(Q.m[0][0], Q.m[0][1]);
  *chunk_ptr_w<5>(vqu, 1, v) = make_double2(Q.m[0][2], Q.m[1][0]);
  *chunk_ptr_w<5>(vqu, 2, v) = make_double2(Q.m[1][1], Q.m[1][2]);
  *chunk_ptr_w<5>(vqu, 3, v) = make_double2(Q.m[2][0], Q.m[2][1]);
  *chunk_ptr_w<5>(vqu, 4, v) = make_double2(Q.m[2][2], fail);
}

// ------------------------------------------------------------------ edge metric length
// the two vertex records of an edge, in registers: 4 doubles each (identity / iso) or 12 (aniso / logm)
template <int KIND>
struct EdgeRecs {
  static constexpr int N = (KIND == MAG_KIND_ANISO || KIND == MAG_KIND_LOGM) ? 12 : 4;
  double a[N], b[N];
};
template <int KIND>
__device__ __forceinline__ void load_edge_recs(const double* __restrict__ vedge, int2 ev, EdgeRecs<KIND>& R)
{
  constexpr int N = EdgeRecs<KIND>::N;
  const double2* pa = chunk_ptr<N / 2>(vedge, 0, ev.x);
  const double2* pb = chunk_ptr<N / 2>(vedge, 0, ev.y);
#pragma unroll
  for (int i = 0; i < N / 2; ++i) { double2 t = __ldg(pa + i * kVB); R.a[2 * i] = t.x; R.a[2 * i + 1] = t.y; }
#pragma unroll
  for (int i = 0; i < N / 2; ++i) { double2 t = __ldg(pb + i * kVB); R.b[2 * i] = t.x; R.b[2 * i + 1] = t.y; }
}

// MetricSizeField::measure: order 2 -> EdgeIntegration::N2, points +-0.577350269189626, weights 1
template <int KIND>
__device__ __forceinline__ double edge_length_strict(const EdgeRecs<KIND>& R, int* eig_fail, const double* __restrict__ vqu = nullptr, int32_t va = 0)
{
  constexpr double NP0 = kNP0s, NP1 = kNP1s, NM0 = kNM0s, NM1 = kNM1s;
  const double* ra = R.a;
  const double* rb = R.b;
  V3 j = magst::edge_j0(V3{ra[0], ra[1], ra[2]}, V3{rb[0], rb[1], rb[2]});
  if (KIND == MAG_KIND_IDENTITY) return magst::mul(2.0, magst::length(j)); // apf::measure, N1 rule
  // The second Gauss point interpolates with the two shape values swapped ((1 - (-XI)) / 2 is the same operation as (1 + XI) / 2,
  // bit for bit).  When both ends carry the SAME size-field values -- an edge along a direction the field does not vary in, any
  // edge of a uniform region -- a * N0 + a * N1 is therefore the same sum with its terms exchanged, floating-point addition
  // commutes, both points see identical interpolated values, identical transforms and identical lengths, and one evaluation
  // serves both: exact, not an approximation.  (The benchmark lattice's z edges are such edges, and they are the ones that sit
  // on the collapse threshold and come here to be re-evaluated.)  That one transform is a function of vertex `va`'s record alone:
  // with the per-vertex pass in place it is read from vqu (k_vertex_uniform: the same operations on the same values) and the
  // edge costs its Jacobian row and one 3 x 3 product.
  bool same = true;
#pragma unroll
  for (int i = 3; i < EdgeRecs<KIND>::N; ++i) same = same && (__double_as_longlong(ra[i]) == __double_as_longlong(rb[i]));
  if ((KIND == MAG_KIND_ANISO || KIND == MAG_KIND_LOGM) && vqu != nullptr && same) {
    const double2 q0 = __ldg(chunk_ptr<5>(vqu, 0, va)), q1 = __ldg(chunk_ptr<5>(vqu, 1, va)), q2 = __ldg(chunk_ptr<5>(vqu, 2, va)),
                  q3 = __ldg(chunk_ptr<5>(vqu, 3, va)), q4 = __ldg(chunk_ptr<5>(vqu, 4, va));
    M3 Q;
    Q.m[0][0] = q0.x; Q.m[0][1] = q0.y; Q.m[0][2] = q1.x;
    Q.m[1][0] = q1.y; Q.m[1][1] = q2.x; Q.m[1][2] = q2.y;
    Q.m[2][0] = q3.x; Q.m[2][1] = q3.y; Q.m[2][2] = q4.x;
    if (q4.y != 0.0) *eig_fail = 1;
    const double len = magst::row0_length(j, Q);
    return magst::add(len, len);
  }
  auto point = [&](const double n0, const double n1) -> double {
    if (KIND == MAG_KIND_ISO) {
      double h = magst::lerp2(ra[3], n0, rb[3], n1);
      double ih = magst::div(1.0, h);
      V3 r{magst::mul(j.x, ih), magst::mul(j.y, ih), magst::mul(j.z, ih)};
      return magst::length(r);
    }
    double c[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) c[i] = magst::lerp2(ra[(3 + i) % EdgeRecs<KIND>::N], n0, rb[(3 + i) % EdgeRecs<KIND>::N], n1);
    M3 Q;
    if (KIND == MAG_KIND_ANISO) {
      magst::transform_aniso(V3{c[3], c[4], c[5]}, V3{c[6], c[7], c[8]}, c[0], c[1], c[2], Q);
    } else {
      M3 A;
#pragma unroll
      for (int i = 0; i < 9; ++i) A.m[i / 3][i % 3] = c[i];
      if (magst::transform_logm(A, Q) != 1) *eig_fail = 1;
    }
    return magst::row0_length(j, Q);
  };
  const double len0 = point(NP0, NP1);
  const double len1 = same ? len0 : point(NM0, NM1);
  return magst::add(len0, len1);
}

template <int KIND>
__device__ __forceinline__ double edge_length_fast(const EdgeRecs<KIND>& R, int* eig_fail)
{
  if (KIND == MAG_KIND_IDENTITY) return magfa::edge_identity(R.a, R.b);
  if (KIND == MAG_KIND_ISO) return magfa::edge_iso(R.a, R.b);
  if (KIND == MAG_KIND_ANISO) return magfa::edge_aniso(R.a, R.b);
  return magfa::edge_logm(R.a, R.b, eig_fail);
}

struct SweepParams {
  uint32_t ops;
  double max_len, min_len, good_q;
  int use_max;
  int reeval = 1;   // fast sweeps: re-evaluate near-threshold edges in strict arithmetic (0: MAG_FP_FAST_LISTED)
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

// Near-threshold work queue (one per warp, in shared memory).  An entity whose value lands within 1e-12
// relative of a threshold is pushed here instead of being decided on the spot; whenever a warp has 32 of them it
// drains the queue with all lanes active: the entity is appended to the global near-threshold list and, in
// MAG_FP_FAST, re-evaluated in strict arithmetic so the flag is the reference's.  Deferring keeps the strict code
// out of the divergent path of the main loop (structured meshes put whole families of edges exactly ON a threshold).
constexpr int kWarps = 16;   // queue rows: the largest block any kernel here uses is 512 threads
constexpr int kQCap = 64;
struct NearQueue { int32_t e[kWarps][kQCap]; int32_t f[kWarps][kQCap]; };

// push: warp-collective.  qn is warp-uniform.  returns true when >= 32 entries wait
__device__ __forceinline__ bool queue_push(NearQueue& q, int& qn, bool nr, int32_t e, int32_t f)
{
  const unsigned m = __ballot_sync(0xffffffffu, nr);
  if (!m) return false;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (nr) {
    const int pos = qn + __popc(m & ((1u << lane) - 1u));
    q.e[w][pos] = e;
    q.f[w][pos] = f;
  }
  qn += __popc(m);
  __syncwarp();
  return qn >= 32;
}
// reserve n slots of the global near list (warp-aggregated: one atomic per drain)
__device__ __forceinline__ unsigned long long near_reserve(unsigned long long* counter, int n)
{
  unsigned long long base = 0;
  if ((threadIdx.x & 31) == 0) base = atomicAdd(counter, (unsigned long long)n);
  return __shfl_sync(0xffffffffu, base, 0);
}

// Work distribution of the persistent kernels: chunks of consecutive entities (kEdgeChunk / kTetChunk, a whole number
// of tiles for every block size used) handed out through an atomic counter.  Consecutive tiles on one SM keep the vertex
// records they share (the next grid row of a box mesh, the neighbouring elements of any reasonably numbered mesh) in
// that SM's L1; dynamic hand-out evens out chunks made expensive by near-threshold re-evaluations or cheap by skip flags.
//
// Occupancy (measured on B200, n = 203 lattice, scripts/run_variants.sh): both kernels are latency-bound, so resident
// warps matter more than registers per thread.  The fast edge kernel fits 80 registers without spilling: 3 x 256 threads
// per SM (24 warps) runs 1.49 -> 1.33 ms against 2 x 256 at 120 registers; 2 x 320 (96 registers) 1.40; 4 x 256 (64
// registers, spills) 1.64.  The fast tet kernel fits 96 registers: 2 x 320 threads 0.90 -> 0.79 ms; 3 x 256 (80
// registers, spills) 0.90; without the {z,det} prefetch 0.86 at 3 x 256.  Strict kernels keep 2 x 256.
#ifndef MAG_EDGE_THREADS
#define MAG_EDGE_THREADS 256
#endif
#ifndef MAG_EDGE_BLOCKS
#define MAG_EDGE_BLOCKS 3
#endif
#ifndef MAG_TET_THREADS
#define MAG_TET_THREADS 320
#endif
#ifndef MAG_TET_BLOCKS
#define MAG_TET_BLOCKS 2
#endif
#ifndef MAG_TET_PIPE
#define MAG_TET_PIPE 1   /* 1: {z,det} chunks of the next tile are prefetched; 0: everything is loaded at the tile */
#endif
#ifndef MAG_STRICT_BLOCKS
#define MAG_STRICT_BLOCKS 2
#endif
constexpr int kStrictThreads = 256, kStrictBlocks = MAG_STRICT_BLOCKS;
#ifndef MAG_EDGE_TILE_SCHED
#define MAG_EDGE_TILE_SCHED 1   /* 1: the edge schedule orders tiles (256 edges); 0: chunks of MAG_EDGE_CHUNK edges */
#endif
#ifndef MAG_FUSE_VERTEX
#define MAG_FUSE_VERTEX 0   /* 1: in fast sweeps the per-vertex pass rides in the edge kernel's launch (k_edges<..,VERT>).
                               Measured on B200 (n = 203): 2.428 -> 2.386 ms lattice, 1.95 -> 1.91 ms jittered with a vertex
                               chunk every 8th ticket -- the vertex work costs the same warp-time inside the persistent
                               kernel as alone (both are latency-bound per warp), so it stays a separate launch by default */
#endif
#ifndef MAG_VERT_EVERY
#define MAG_VERT_EVERY 8
#endif
#ifndef MAG_EDGE_CHUNK
#define MAG_EDGE_CHUNK 8192   /* 32 tiles of 256 */
#endif
#ifndef MAG_TET_CHUNK
#define MAG_TET_CHUNK 7680    /* 24 tiles of 320 = 30 tiles of 256 */
#endif
constexpr int kEdgeChunk = MAG_EDGE_CHUNK;
constexpr int kTetChunk = MAG_TET_CHUNK;
#ifndef MAG_EDGE_BLOCKS_LOGM
#define MAG_EDGE_BLOCKS_LOGM 2   /* the log-Euclidean kernel (QR iteration) wants 128 registers: 5.4 ms against 6.3 ms at 3 x 256 */
#endif
template <int KIND, bool FAST> struct EdgeCfg {
  static constexpr int T = FAST ? MAG_EDGE_THREADS : kStrictThreads;
  static constexpr int B = FAST ? (KIND == MAG_KIND_LOGM ? MAG_EDGE_BLOCKS_LOGM : MAG_EDGE_BLOCKS) : kStrictBlocks;
};
template <bool FAST> struct TetCfg { static constexpr int T = FAST ? MAG_TET_THREADS : kStrictThreads, B = FAST ? MAG_TET_BLOCKS : kStrictBlocks; };
__device__ __forceinline__ long long next_chunk(unsigned long long* counter, long long* slot)
{
  __syncthreads();                       // everyone is done with the previous value of *slot
  if (threadIdx.x == 0) *slot = (long long)atomicAdd(counter, 1ull);
  __syncthreads();
  return *slot;
}
// What the edge kernel needs of the sweep parameters, decoded once on the host so that the per-edge flag logic is a
// handful of integer operations:
//   off_bits  OR-ed into the flag word before the skip tests: the skip bit of every mark that was NOT requested
//   err_mask  true flags of the requested marks; markEntities asserts they are clear on every entity (maAdapt.cc:308)
//   tol_*     MAG_NEAR_REL * |threshold|, or -1 for an infinite threshold (nothing is ever near it)
struct EdgeParams {
  int32_t off_bits, err_mask;
  int want_len;
  int zero_in;   // the incoming flag words are all zero and the array was not materialised: do not read it
  int reeval;    // fast sweeps: near-threshold edges are re-evaluated in strict arithmetic (MAG_FP_FAST) or only listed (MAG_FP_FAST_LISTED)
  uint32_t ops;
  double max_len, min_len, tol_max, tol_min;
  const double* vqu;   // k_vertex_uniform's per-vertex transforms when they are in place for this size field, else null
};
constexpr int32_t kSkipSplit = MAG_DONT_SPLIT | MAG_NEED_NOT_SPLIT, kSkipColl = MAG_DONT_COLLAPSE | MAG_NEED_NOT_COLLAPSE;

// drain n (<= 32) queued edges: entry i is handled by lane i.  Returns bit 0: evaluated, bit 1: counted SPLIT, bit 2: counted COLLAPSE
template <int KIND, bool FAST>
__device__ __noinline__ unsigned drain_edges(NearQueue& q, int first, int n, const int2* __restrict__ edge_v,
                                             const double* __restrict__ vedge,
                                             int32_t* __restrict__ flags, double* __restrict__ lengths, uint32_t ops,
                                             double max_len, double min_len,
                                             MagDevStats* st, int32_t* __restrict__ near_list, int32_t id_base, bool reeval,
                                             const double* __restrict__ vqu)
{
  SweepParams P{ops, max_len, min_len, 0.0, 0};
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned long long base = near_reserve(&st->n_near_edge, n);
  unsigned out = 0;
  if (lane < n) {
    const int32_t e = q.e[w][first + lane];
    near_list[base + lane] = e + id_base;
    if (FAST && reeval) {                     // MAG_FP_FAST_LISTED: listed only, the flag was written from the fast value
      int32_t f = q.f[w][first + lane];
      const bool need_split = (P.ops & MAG_OP_MARK_SPLIT) && !(f & kSkipSplit);
      const bool need_coll = (P.ops & MAG_OP_MARK_COLLAPSE) && !(f & kSkipColl);
      EdgeRecs<KIND> R;
      int2 ev = __ldg(edge_v + e);
      const bool owned = ev.x >= 0;
      ev.x &= kVidMask;
      load_edge_recs<KIND>(vedge, ev, R);
      int eig = 0;
      const double len = edge_length_strict<KIND>(R, &eig, vqu, ev.x);
      unsigned cs = 0, cc = 0;
      mark_edge(len, f, need_split, need_coll, owned, P, cs, cc);
      out = 1u | (cs << 1) | (cc << 2);
      flags[e] = f;
      if (P.ops & MAG_OP_LENGTHS) lengths[e] = len;
    }
  }
  __syncwarp();
  return out;
}

// Persistent edge kernel.  Per thread and tile: the flag word and the end vertices are loaded one tile ahead; the two
// vertex records are gathered only when the flag word says the edge has to be evaluated at all (skipped edges cost 12
// bytes).  Everything is indexed with int32 (mag_set_mesh rejects meshes with 2^31 or more entities of one dimension).
// VERT: the per-vertex pass rides in the same launch.  Every kVertEvery-th ticket is a chunk of vertices instead of a
// chunk of edges, so the HBM-bound streaming of the vertex pass (96 B in, 112 B out per vertex) overlaps the
// latency-bound edge work of the other resident CTAs instead of preceding it, and both sweep the vertex array in step
// (the vertex pass finds the gather records in L2).  The tet kernel of the same sweep is launched after this one.
struct VertArgs { int32_t nv; int dim; double* vpos; double* vq; };
constexpr int kVertChunk = 8192, kVertEvery = MAG_VERT_EVERY;
template <int KIND, bool FAST, bool VERT>
__global__ void __launch_bounds__(EdgeCfg<KIND, FAST>::T, EdgeCfg<KIND, FAST>::B)
k_edges(int32_t ne, const int2* __restrict__ edge_v, const double* __restrict__ vedge,
        int32_t* __restrict__ flags, double* __restrict__ lengths,
        EdgeParams P, MagDevStats* st, int32_t* __restrict__ near_list, const int32_t* __restrict__ chunk_order, int32_t id_base,
        VertArgs V)
{
  __shared__ NearQueue q;
  __shared__ long long chunk_slot;
  unsigned c_split = 0, c_coll = 0, c_eval = 0, c_err = 0;
  double maxlen = 0.0;                  // getMaximumEdgeLength starts at 0 and ignores NaN (maSize.cc:673-691)
  int qn = 0, eig_any = 0;
  constexpr int kEdgeThreads = EdgeCfg<KIND, FAST>::T, kChunkEdges = kEdgeChunk;
  const int nchunks = (ne + kChunkEdges - 1) / kChunkEdges;
  const int nvchunks = VERT ? (V.nv + kVertChunk - 1) / kVertChunk : 0;
  for (;;) {
    long long ticket = next_chunk(&st->edge_chunk, &chunk_slot);
    if (VERT) {
      // tickets 0, kVertEvery, 2 kVertEvery, ... are the vertex chunks; the others are edge chunks in schedule order
      const long long nint = (long long)nvchunks * kVertEvery;      // tickets of the interleaved stretch
      const long long ntickets = nint > (long long)nchunks + nvchunks ? nint : (long long)nchunks + nvchunks;
      if (ticket >= ntickets) break;
      if (ticket < nint && ticket % kVertEvery == 0) {
        const int v0 = (int)(ticket / kVertEvery) * kVertChunk;
        const int v_end = (V.nv - v0 < kVertChunk) ? V.nv : v0 + kVertChunk;
        for (int v = v0 + (int)threadIdx.x; v < v_end; v += kEdgeThreads) vertex_pass_one<KIND>(v, V.dim, vedge, V.vpos, V.vq, &eig_any);
        continue;
      }
      ticket = ticket < nint ? ticket - ticket / kVertEvery - 1 : ticket - nvchunks;
      if (ticket >= nchunks) continue;                                // fewer edge chunks than slots between vertex chunks
    } else if (ticket >= nchunks) break;
#if MAG_EDGE_TILE_SCHED
    // tile-granular schedule: the ticket is a group of kChunkEdges / T consecutive SCHEDULE positions; position p holds
    // the index of a tile of T consecutive edges.  The schedule is sorted by the smallest vertex id a tile touches, so
    // the tiles a CTA works through back to back belong to different edge families over the same window of vertices
    // (box mesh: x, y, z, three face diagonals, body diagonal) and find each other's records in L1.
    static_assert(kEdgeThreads == kStrictThreads, "the tile schedule is built for tiles of kStrictThreads edges");
    constexpr int kGroup = kChunkEdges / kEdgeThreads;
    const int ntiles = (ne + kEdgeThreads - 1) / kEdgeThreads;
    const int k0 = (int)ticket * kGroup;
    const int tiles = (ntiles - k0 < kGroup) ? ntiles - k0 : kGroup;
    const int e_end = ne;
    auto tile_base = [&](int p) { return (chunk_order ? __ldg(chunk_order + k0 + p) : k0 + p) * kEdgeThreads + (int)threadIdx.x; };
#else
    const int e0 = (chunk_order ? chunk_order[ticket] : (int)ticket) * kChunkEdges;   // no schedule: a sub-range sweep (mag_sweep_host)
    const int e_end = (ne - e0 < kChunkEdges) ? ne : e0 + kChunkEdges;   // first edge past this chunk
    const int tiles = (e_end - e0 + kEdgeThreads - 1) / kEdgeThreads;
    auto tile_base = [&](int p) { return e0 + p * kEdgeThreads + (int)threadIdx.x; };
#endif
    int e = tile_base(0);
    int e_nx = tiles > 1 ? tile_base(1) : e_end;
    int32_t f = 0;
    int2 ev = make_int2(0, 0);
    if (e < e_end) { f = P.zero_in ? 0 : ld_stream_rw(flags + e); ev = ld_stream(edge_v + e); }
    for (int tile = 0; tile < tiles; ++tile) {
      const int e_nx2 = tile + 2 < tiles ? tile_base(tile + 2) : e_end;
      int32_t f_nx = 0;
      int2 ev_nx = make_int2(0, 0);
      if (e_nx < e_end) { f_nx = P.zero_in ? 0 : ld_stream_rw(flags + e_nx); ev_nx = ld_stream(edge_v + e_nx); }
      bool nr = false;
      if (e < e_end) {
        const int32_t fe = f | P.off_bits;
        const bool need_split = !(fe & kSkipSplit), need_coll = !(fe & kSkipColl);
        if (f & P.err_mask) ++c_err;
        if (P.want_len || need_split || need_coll) {
          EdgeRecs<KIND> R;
          const bool owned = ev.x >= 0;                 // sign bit of the first vertex id = "not owned" (k_fold_owned)
          load_edge_recs<KIND>(vedge, make_int2(ev.x & kVidMask, ev.y), R);
          const double len = FAST ? edge_length_fast<KIND>(R, &eig_any) : edge_length_strict<KIND>(R, &eig_any);   // (no Q_u here: a third of a tile's lanes taking the short way costs the warp both ways -- lattice 2.95 -> 3.19 ms)
          if (P.want_len) {
            st_stream(lengths + e, len);
            if (owned && len > maxlen) maxlen = len;
          }
          if (need_split || need_coll) {
            nr = (need_split && fabs(len - P.max_len) <= P.tol_max) || (need_coll && fabs(len - P.min_len) <= P.tol_min);
            // A near-threshold edge of a MAG_FP_FAST sweep is re-evaluated in the reference's operation order (drain_edges) so
            // that its flag is the strict one; MAG_FP_FAST_LISTED decides it by the value at hand and only lists it
            const bool kReevaluate = FAST && P.reeval;
            if (!(kReevaluate && nr)) {
              ++c_eval;
              int32_t g = f;
              if (need_split) {
                const bool t = len > P.max_len;
                g |= t ? MAG_SPLIT : MAG_NEED_NOT_SPLIT;
                c_split += (t && owned) ? 1u : 0u;
              }
              if (need_coll) {
                const bool t = len < P.min_len;
                g |= t ? MAG_COLLAPSE : MAG_NEED_NOT_COLLAPSE;
                c_coll += (t && owned) ? 1u : 0u;
              }
              st_stream(flags + e, g);
            }
          }
        }
      }
      if (queue_push(q, qn, nr, (int32_t)e, f)) {
        qn -= 32;
        const unsigned r = drain_edges<KIND, FAST>(q, qn, 32, edge_v, vedge, flags, lengths, P.ops, P.max_len, P.min_len, st, near_list, id_base, P.reeval != 0, P.vqu);
        c_eval += r & 1u; c_split += (r >> 1) & 1u; c_coll += (r >> 2) & 1u;
      }
      f = f_nx;
      ev = ev_nx;
      e = e_nx;
      e_nx = e_nx2;
    }
  }
  if (qn) {
    const unsigned r = drain_edges<KIND, FAST>(q, 0, qn, edge_v, vedge, flags, lengths, P.ops, P.max_len, P.min_len, st, near_list, id_base, P.reeval != 0, P.vqu);
    c_eval += r & 1u; c_split += (r >> 1) & 1u; c_coll += (r >> 2) & 1u;
  }
  if (eig_any) atomicAdd(&st->n_eigen_fail, 1ull);
  warp_count_to(c_split, &st->n_split);
  warp_count_to(c_coll, &st->n_collapse);
  warp_count_to(c_eval, &st->n_edges_eval);
  warp_count_to(c_err, &st->n_flag_err);
  if (P.want_len) {
    // a non-negative double orders like its bit pattern
    const unsigned long long m = warp_max_u64((unsigned long long)__double_as_longlong(maxlen));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&st->max_len_bits, m);
  }
}

// ma::getAverageEdgeLength's sum of the stored lengths over all edges of the part (MAG_OP_LENGTH_SUM): a fixed-shape
// two-level tree (per-thread strided partial -> warp -> block -> k_finish_sum), deterministic for a given ne
__global__ void __launch_bounds__(kThreads)
k_sum_lengths(int64_t ne, const double* __restrict__ lengths, const uint8_t* __restrict__ owned_arr, double* __restrict__ block_sums)
{
  __shared__ double sh[kWarps];
  double s = 0;
  for (int64_t e = blockIdx.x * (int64_t)kThreads + threadIdx.x; e < ne; e += (int64_t)gridDim.x * kThreads)
    if (!owned_arr || owned_arr[e]) s += lengths[e];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < kThreads / 32; ++i) t += sh[i];   // the block's warps (sh is sized for the largest block)
    block_sums[blockIdx.x] = t;
  }
}

// ------------------------------------------------------------------ tets
// centroid transform for useMax == false (maQuality.cc:148-153): N = (1-.25-.25-.25, .25, .25, .25)
template <int KIND>
__device__ __forceinline__ void centroid_transform(const double* __restrict__ vedge, int64_t nv, const int4& tv, M3& Q, int* eig)
{
  constexpr double N0 = 1 - 0.25 - 0.25 - 0.25;
  const int32_t vid[4] = {tv.x, tv.y, tv.z, tv.w};
  if constexpr (KIND == MAG_KIND_IDENTITY) { magst::identity(Q); return; }
  else if constexpr (KIND == MAG_KIND_ISO) {
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
  } else {
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
}

// getMetricWithMaxJacobean: strict >, first maximum wins (maQuality.cc:97-104)
__device__ __forceinline__ int32_t best_vertex(const int4& tv, double d0, double d1, double d2, double d3)
{
  int32_t vb = tv.x;
  double maxJ = -1.0;
  if (d0 > maxJ) { maxJ = d0; vb = tv.x; }
  if (d1 > maxJ) { maxJ = d1; vb = tv.y; }
  if (d2 > maxJ) { maxJ = d2; vb = tv.z; }
  if (d3 > maxJ) { maxJ = d3; vb = tv.w; }
  return vb;
}
__device__ __forceinline__ void load_q(const double* __restrict__ vq, int32_t vb, M3& Q, double& detQ)
{
  double2 q0 = __ldg(chunk_ptr<5>(vq, 0, vb)), q1 = __ldg(chunk_ptr<5>(vq, 1, vb)), q2 = __ldg(chunk_ptr<5>(vq, 2, vb)),
          q3 = __ldg(chunk_ptr<5>(vq, 3, vb)), q4 = __ldg(chunk_ptr<5>(vq, 4, vb));
  Q.m[0][0] = q0.x; Q.m[0][1] = q0.y; Q.m[0][2] = q1.x;
  Q.m[1][0] = q1.y; Q.m[1][1] = q2.x; Q.m[1][2] = q2.y;
  Q.m[2][0] = q3.x; Q.m[2][1] = q3.y; Q.m[2][2] = q4.x;
  detQ = q4.y;
}

// HAVE_DETS: the four det Q_v were loaded one tile ahead (dets[]), so the transform of the best vertex can be
// requested together with the coordinates instead of after them
template <int KIND, bool FAST, bool HAVE_DETS>
__device__ __forceinline__ double tet_quality_eval(const int4& tv, int64_t nv, const double* __restrict__ vpos,
                                                   const double* __restrict__ vq, const double* __restrict__ vedge,
                                                   int use_max, int* eig, const double* dets)
{
  M3 Q;
  double detQ = 0.0;
  if (HAVE_DETS && use_max) load_q(vq, best_vertex(tv, dets[0], dets[1], dets[2], dets[3]), Q, detQ);
  double p[4][4];
  load_rec4(vpos, tv.x, p[0]);
  load_rec4(vpos, tv.y, p[1]);
  load_rec4(vpos, tv.z, p[2]);
  load_rec4(vpos, tv.w, p[3]);
  if (use_max) {
    if (!HAVE_DETS) load_q(vq, best_vertex(tv, p[0][3], p[1][3], p[2][3], p[3][3]), Q, detQ);
  } else {
    centroid_transform<KIND>(vedge, nv, tv, Q, eig);
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

// What the tet kernel needs of the sweep parameters, decoded once on the host (see EdgeParams)
struct TetParams {
  uint32_t ops;
  int do_bad, want_q, use_max;
  int zero_in;   // see EdgeParams
  double good_q, tol_q;   // tol_q = MAG_NEAR_REL * |good_q|, or -1 when good_q is not finite
};

// returns bit 0: evaluated, bit 1: counted BAD_QUALITY
template <int KIND, bool FAST>
__device__ __noinline__ unsigned drain_tets(NearQueue& q, int first, int n, int32_t elem_off, int64_t nv, const int4* __restrict__ tet_v,
                                            const double* __restrict__ vpos, const double* __restrict__ vq,
                                            const double* __restrict__ vedge,
                                            int32_t* __restrict__ flags, double* __restrict__ qual, uint32_t ops,
                                            double good_q, int use_max,
                                            MagDevStats* st, int32_t* __restrict__ near_list)
{
  SweepParams P{ops, 0.0, 0.0, good_q, use_max};
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned long long base = near_reserve(&st->n_near_elem, n);
  unsigned out = 0;
  if (lane < n) {
    const int32_t t = q.e[w][first + lane];
    const int32_t el = elem_off + t;
    near_list[base + lane] = el;
    if (FAST) {
      int32_t f = q.f[w][first + lane];
      int4 tv = __ldg(tet_v + t);
      const bool owned = tv.x >= 0;
      tv.x &= kVidMask;
      int eig = 0;
      const double qv = tet_quality_eval<KIND, false, false>(tv, nv, vpos, vq, vedge, P.use_max, &eig, nullptr);
      unsigned cb = 0;
      mark_tet(qv, f, owned, P, cb);
      out = 1u | (cb << 1);
      flags[el] = f;
      if (P.ops & MAG_OP_QUALITIES) qual[el] = qv;
    }
  }
  __syncwarp();
  return out;
}

// Persistent tet kernel, int32 indexing throughout (mag_set_mesh rejects larger meshes).  Software pipeline per thread
// and tile k: flag word + vertices two tiles ahead; the four {z, det Q_v} chunks one tile ahead, so that the choice of
// the max-Jacobian vertex (getMetricWithMaxJacobean, maQuality.cc:83-108) does not sit between two dependent gathers and
// no chunk is fetched twice: at the tile itself only the four {x,y} chunks and the winner's transform are loaded.
template <int KIND, bool FAST, bool USE_MAX>
__global__ void __launch_bounds__(TetCfg<FAST>::T, TetCfg<FAST>::B)
k_tets(int32_t nt, int32_t elem_off, int64_t nv, const int4* __restrict__ tet_v, const double* __restrict__ vpos,
       const double* __restrict__ vq, const double* __restrict__ vedge,
       int32_t* __restrict__ flags, double* __restrict__ qual, TetParams P, MagDevStats* st,
       int32_t* __restrict__ near_list, const int32_t* __restrict__ chunk_order)
{
  __shared__ NearQueue q;
  __shared__ long long chunk_slot;
  unsigned c_bad = 0, c_eval = 0, c_err = 0;
  unsigned long long minkey = ~0ull;
  int qn = 0, eig_any = 0;
  constexpr int kTetThreads = TetCfg<FAST>::T, kChunkTets = kTetChunk;
  const int nchunks = (nt + kChunkTets - 1) / kChunkTets;
  flags += elem_off;
  qual += elem_off;
  auto load_zd = [&](const int4& tv, double2* zd) {
    zd[0] = __ldg(chunk_ptr<2>(vpos, 1, tv.x & kVidMask)); zd[1] = __ldg(chunk_ptr<2>(vpos, 1, tv.y));
    zd[2] = __ldg(chunk_ptr<2>(vpos, 1, tv.z)); zd[3] = __ldg(chunk_ptr<2>(vpos, 1, tv.w));
  };
  for (;;) {
    const long long ticket = next_chunk(&st->elem_chunk, &chunk_slot);
    if (ticket >= nchunks) break;
    const int t0 = (chunk_order ? chunk_order[ticket] : (int)ticket) * kChunkTets;
    const int t_end = (nt - t0 < kChunkTets) ? nt : t0 + kChunkTets;
    const int tiles = (t_end - t0 + kTetThreads - 1) / kTetThreads;
    int t = t0 + (int)threadIdx.x;
    int32_t f_cur = 0, f_nx = 0;
    int4 tv_cur = make_int4(0, 0, 0, 0), tv_nx = tv_cur;
    double2 zd[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) zd[i] = make_double2(0.0, 0.0);
    if (t < t_end) {
      f_cur = P.zero_in ? 0 : ld_stream_rw(flags + t);
      tv_cur = ld_stream(tet_v + t);
      if (MAG_TET_PIPE && (P.want_q || (P.do_bad && !(f_cur & MAG_OK_QUALITY)))) load_zd(tv_cur, zd);
    }
    if (t + kTetThreads < t_end) { f_nx = P.zero_in ? 0 : ld_stream_rw(flags + t + kTetThreads); tv_nx = ld_stream(tet_v + t + kTetThreads); }
    for (int tile = 0; tile < tiles; ++tile, t += kTetThreads) {
      const int32_t f_in = f_cur;
      int4 tv = tv_cur;
      const bool owned = tv.x >= 0;
      tv.x &= kVidMask;
      if (!MAG_TET_PIPE && (P.want_q || (P.do_bad && !(f_in & MAG_OK_QUALITY)))) load_zd(tv, zd);
      const double2 z0 = zd[0], z1 = zd[1], z2 = zd[2], z3 = zd[3];
      // next tile: its {z,det} chunks; tile after: flag word + vertices
      f_cur = f_nx;
      tv_cur = tv_nx;
      if (MAG_TET_PIPE && t + kTetThreads < t_end && (P.want_q || (P.do_bad && !(f_cur & MAG_OK_QUALITY)))) load_zd(tv_cur, zd);
      if (t + 2 * kTetThreads < t_end) { f_nx = P.zero_in ? 0 : ld_stream_rw(flags + t + 2 * kTetThreads); tv_nx = ld_stream(tet_v + t + 2 * kTetThreads); }
      bool nr = false;
      if (t < t_end) {
        int32_t f = f_in;
        if (P.do_bad && (f & MAG_BAD_QUALITY)) ++c_err;
        const bool need_bad = P.do_bad && !(f & MAG_OK_QUALITY);
        if (P.want_q || need_bad) {
          M3 Q;
          double detQ = 0.0;
          if (USE_MAX) load_q(vq, best_vertex(tv, z0.y, z1.y, z2.y, z3.y), Q, detQ);
          const double2 a0 = __ldg(chunk_ptr<2>(vpos, 0, tv.x)), a1 = __ldg(chunk_ptr<2>(vpos, 0, tv.y)),
                        a2 = __ldg(chunk_ptr<2>(vpos, 0, tv.z)), a3 = __ldg(chunk_ptr<2>(vpos, 0, tv.w));
          if (!USE_MAX) {   // centroid metric (maQuality.cc:148-153): a separate instantiation keeps its registers out of the default path
            centroid_transform<KIND>(vedge, nv, tv, Q, &eig_any);
            detQ = FAST ? magst::det3(Q) : 0.0;
          }
          const V3 x[4] = {V3{a0.x, a0.y, z0.x}, V3{a1.x, a1.y, z1.x}, V3{a2.x, a2.y, z2.x}, V3{a3.x, a3.y, z3.x}};
          const double qv = FAST ? magfa::tet_quality(x, Q, detQ) : magst::tet_quality(x, Q);
          if (P.want_q) {
            st_stream(qual + t, qv);
            const unsigned long long k = dkey(qv);
            minkey = k < minkey ? k : minkey;
          }
          nr = need_bad && fabs(qv - P.good_q) <= P.tol_q;
          if (need_bad && !(FAST && nr)) {
            ++c_eval;
            const bool bad = qv < P.good_q;
            c_bad += (bad && owned) ? 1u : 0u;
            st_stream(flags + t, (int32_t)(f | (bad ? MAG_BAD_QUALITY : MAG_OK_QUALITY)));
          }
        }
      }
      if (queue_push(q, qn, nr, (int32_t)t, f_in)) {
        qn -= 32;
        const unsigned r = drain_tets<KIND, FAST>(q, qn, 32, elem_off, nv, tet_v, vpos, vq, vedge, flags - elem_off, qual - elem_off, P.ops, P.good_q, P.use_max, st, near_list);
        c_eval += r & 1u; c_bad += (r >> 1) & 1u;
      }
    }
  }
  if (qn) {
    const unsigned r = drain_tets<KIND, FAST>(q, 0, qn, elem_off, nv, tet_v, vpos, vq, vedge, flags - elem_off, qual - elem_off, P.ops, P.good_q, P.use_max, st, near_list);
    c_eval += r & 1u; c_bad += (r >> 1) & 1u;
  }
  if (eig_any) atomicAdd(&st->n_eigen_fail, 1ull);
  warp_count_to(c_bad, &st->n_bad);
  warp_count_to(c_eval, &st->n_elems_eval);
  warp_count_to(c_err, &st->n_flag_err);
  if (P.want_q) {
    const unsigned long long m = warp_min_u64(minkey);
    if ((threadIdx.x & 31) == 0 && m != ~0ull) atomicMin(&st->min_q_key, m);
  }
}

// ------------------------------------------------------------------ cavity batches
// Batch form of ma::getWorstQuality (ma/maQuality.cc:184-210) for many candidate cavities at once: cavity k is the list
// of tets [offsets[k], offsets[k+1]) given by their four vertex ids -- existing vertices of the resident part, but the tets
// need not exist in the mesh (collapse / swap / snap evaluate would-be elements, maCollapse.cc:37-113,
// maEdgeSwap.cc:598-740).  One warp per cavity, lanes stride over its tets, warp-wide minimum: no atomics, deterministic.
template <int KIND, bool FAST>
__global__ void __launch_bounds__(256)
k_cavity_quality(int64_t ncav, const int64_t* __restrict__ offsets, const int4* __restrict__ tet_v, int64_t nv,
                 const double* __restrict__ vpos, const double* __restrict__ vq, const double* __restrict__ vedge,
                 int use_max, double* __restrict__ worst, double* __restrict__ qual, MagDevStats* st)
{
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int eig = 0;
  for (int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < ncav; k += nwarps) {
    const int64_t lo = offsets[k], hi = offsets[k + 1];
    unsigned long long key = ~0ull;
    for (int64_t t = lo + lane; t < hi; t += 32) {
      const int4 tv = __ldg(tet_v + t);
      const double q = tet_quality_eval<KIND, FAST, false>(tv, nv, vpos, vq, vedge, use_max, &eig, nullptr);
      if (qual) qual[t] = q;
      const unsigned long long kq = dkey(q);
      key = kq < key ? kq : key;
    }
    key = warp_min_u64(key);
    if (lane == 0) {
      const unsigned long long b = (key >> 63) ? (key & 0x7fffffffffffffffull) : ~key;   // inverse of dkey()
      worst[k] = __longlong_as_double((long long)b);
    }
  }
  if (eig) atomicAdd(&st->n_eigen_aux, 1ull);
}

// ------------------------------------------------------------------ edge-collapse candidates (the consumer of the batch)
// What ma::Collapse decides a collapse on (ma/maCollapse.cc:37-113,353-383,425-433), for many candidates at once, on the resident
// part alone: candidate k collapses vertex vc of edge e onto the other end vk.  Elements around vc that contain the edge
// disappear (elementsToCollapse); every other element around vc is rebuilt with vc replaced by vk, vertex order kept
// (rebuildElements -> ma::rebuildElement).  new_worst[k] = getWorstQuality(newElements), old_worst[k] = getOldQuality() =
// getWorstQuality of ALL elements around vc, n_keep[k] = number of rebuilt elements.  The caller accepts when
// !(new_worst < min(goodQuality, max(old_worst, validQuality))) (tryBothDirections).  Needs the vertex -> tet incidence, built
// on the device from the resident connectivity on first use (count, scan, fill); one warp per candidate, lanes stride over the
// tets around vc, warp-wide minima.
__global__ void __launch_bounds__(kThreads)
k_v2t_count(int64_t nt, const int4* __restrict__ tet_v, int32_t* __restrict__ cnt)
{
  const int64_t t = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (t >= nt) return;
  const int4 tv = tet_v[t];
  atomicAdd(cnt + (tv.x & kVidMask), 1); atomicAdd(cnt + tv.y, 1); atomicAdd(cnt + tv.z, 1); atomicAdd(cnt + tv.w, 1);
}
__global__ void __launch_bounds__(kThreads)
k_v2t_fill(int64_t nt, const int4* __restrict__ tet_v, int32_t* __restrict__ cursor, int32_t* __restrict__ list)
{
  const int64_t t = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (t >= nt) return;
  const int4 tv = tet_v[t];
  list[atomicAdd(cursor + (tv.x & kVidMask), 1)] = (int32_t)t;
  list[atomicAdd(cursor + tv.y, 1)] = (int32_t)t;
  list[atomicAdd(cursor + tv.z, 1)] = (int32_t)t;
  list[atomicAdd(cursor + tv.w, 1)] = (int32_t)t;
}
template <int KIND, bool FAST>
__global__ void __launch_bounds__(256)
k_collapse_quality(int64_t ncand, const int32_t* __restrict__ cand_edge, const uint8_t* __restrict__ cand_end, const int2* __restrict__ edge_v,
                   const int32_t* __restrict__ v2t_off, const int32_t* __restrict__ v2t, const int4* __restrict__ tet_v, int64_t nv,
                   const double* __restrict__ vpos, const double* __restrict__ vq, const double* __restrict__ vedge, int use_max,
                   double* __restrict__ new_worst, double* __restrict__ old_worst, int32_t* __restrict__ n_keep, MagDevStats* st)
{
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  int eig = 0;
  for (int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < ncand; k += nwarps) {
    int2 ev = __ldg(edge_v + cand_edge[k]);
    ev.x &= kVidMask;
    const int32_t vc = cand_end[k] ? ev.y : ev.x, vk = cand_end[k] ? ev.x : ev.y;
    const int32_t lo = v2t_off[vc], hi = v2t_off[vc + 1];
    unsigned long long key_new = ~0ull, key_old = ~0ull;
    int keep = 0;
    for (int32_t i = lo + lane; i < hi; i += 32) {
      int4 tv = __ldg(tet_v + v2t[i]);
      tv.x &= kVidMask;
      const unsigned long long ko = dkey(tet_quality_eval<KIND, FAST, false>(tv, nv, vpos, vq, vedge, use_max, &eig, nullptr));
      key_old = ko < key_old ? ko : key_old;
      if (tv.x == vk || tv.y == vk || tv.z == vk || tv.w == vk) continue;       // contains the edge: collapses away
      if (tv.x == vc) tv.x = vk; else if (tv.y == vc) tv.y = vk; else if (tv.z == vc) tv.z = vk; else tv.w = vk;
      const unsigned long long kn = dkey(tet_quality_eval<KIND, FAST, false>(tv, nv, vpos, vq, vedge, use_max, &eig, nullptr));
      key_new = kn < key_new ? kn : key_new;
      ++keep;
    }
    key_new = warp_min_u64(key_new);
    key_old = warp_min_u64(key_old);
    keep = __reduce_add_sync(0xffffffffu, keep);
    if (lane == 0) {
      const unsigned long long bn = (key_new >> 63) ? (key_new & 0x7fffffffffffffffull) : ~key_new;   // inverse of dkey()
      const unsigned long long bo = (key_old >> 63) ? (key_old & 0x7fffffffffffffffull) : ~key_old;
      new_worst[k] = keep ? __longlong_as_double((long long)bn) : __longlong_as_double(0x7ff0000000000000ll);   // nothing rebuilt: +inf
      old_worst[k] = hi > lo ? __longlong_as_double((long long)bo) : __longlong_as_double(0x7ff0000000000000ll);
      n_keep[k] = keep;
    }
  }
  if (eig) atomicAdd(&st->n_eigen_aux, 1ull);
}

// ------------------------------------------------------------------ triangles (2-D meshes)
// measureTriQuality (maQuality.cc:110-136): 48 A^2 / (sum l^2)^2 with the transform of the max-"Jacobian" vertex
// (|row0 x row1| of Q_v on a 2-D mesh) or of the centroid (xi = 1/3, 1/3).  2-D meshes are small next to the 3-D
// benchmark parts, so this kernel is a plain one-thread-per-triangle grid-stride loop; near-threshold triangles are
// re-evaluated in place.
template <int KIND>
__device__ __forceinline__ void centroid_transform_tri(const double* __restrict__ vedge, int64_t nv, const int32_t vid[3], M3& Q, int* eig)
{
  constexpr double N1 = 1. / 3., N0 = 1 - 1. / 3. - 1. / 3.;
  if constexpr (KIND == MAG_KIND_IDENTITY) { magst::identity(Q); return; }
  else if constexpr (KIND == MAG_KIND_ISO) {
    double h = 0;
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      double r[4];
      load_rec4(vedge, vid[n], r);
      h = magst::add(h, magst::mul(r[3], n ? N1 : N0));
    }
    magst::identity(Q);
    double ih = magst::div(1.0, h);
    Q.m[0][0] = ih; Q.m[1][1] = ih; Q.m[2][2] = ih;
    return;
  } else {
    double c[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) c[i] = 0;
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      Rec12 r = load_rec12(vedge, vid[n]);
#pragma unroll
      for (int i = 0; i < 9; ++i) c[i] = magst::add(c[i], magst::mul(r.v[3 + i], n ? N1 : N0));
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
}

template <int KIND, bool FAST>
__device__ __forceinline__ double tri_quality_eval(const int32_t vid[3], int64_t nv, const double* __restrict__ vpos,
                                                   const double* __restrict__ vq, const double* __restrict__ vedge,
                                                   int use_max, int* eig)
{
  double p[3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i) load_rec4(vpos, vid[i], p[i]);
  M3 Q;
  double detQ;
  if (use_max) {
    int32_t vb = vid[0];
    double maxJ = -1.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (p[i][3] > maxJ) { maxJ = p[i][3]; vb = vid[i]; }
    load_q(vq, vb, Q, detQ);
  } else {
    centroid_transform_tri<KIND>(vedge, nv, vid, Q, eig);
  }
  V3 x[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) x[i] = V3{p[i][0], p[i][1], p[i][2]};
  return FAST ? magfa::tri_quality(x, Q) : magst::tri_quality(x, Q);
}

template <int KIND, bool FAST>
__global__ void __launch_bounds__(kThreads)
k_tris(int64_t nt, int64_t nv, const int32_t* __restrict__ tri_v, const double* __restrict__ vpos,
       const double* __restrict__ vq, const double* __restrict__ vedge, const uint8_t* __restrict__ owned_arr,
       int32_t* __restrict__ flags, double* __restrict__ qual, SweepParams P, MagDevStats* st,
       int32_t* __restrict__ near_list)
{
  unsigned c_bad = 0, c_eval = 0, c_err = 0;
  unsigned long long minkey = ~0ull;
  int eig_any = 0;
  const bool do_bad = P.ops & MAG_OP_MARK_BAD, want_q = P.ops & MAG_OP_QUALITIES;
  const int64_t nloop = (nt + kThreads - 1) / kThreads * kThreads;   // whole warps stay together for the reductions
  for (int64_t t = blockIdx.x * (int64_t)kThreads + threadIdx.x; t < nloop; t += (int64_t)gridDim.x * kThreads) {
    if (t >= nt) continue;
    int32_t f = flags[t];
    const int32_t f_in = f;
    if (do_bad && (f & MAG_BAD_QUALITY)) ++c_err;
    const bool need_bad = do_bad && !(f & MAG_OK_QUALITY);
    if (!(want_q || need_bad)) continue;
    const int32_t vid[3] = {tri_v[3 * t], tri_v[3 * t + 1], tri_v[3 * t + 2]};
    double qv = tri_quality_eval<KIND, FAST>(vid, nv, vpos, vq, vedge, P.use_max, &eig_any);
    if (need_bad && near_thr(qv, P.good_q)) {
      const unsigned long long k = atomicAdd(&st->n_near_elem, 1ull);
      near_list[k] = (int32_t)t;
      if (FAST) qv = tri_quality_eval<KIND, false>(vid, nv, vpos, vq, vedge, P.use_max, &eig_any);
    }
    if (want_q) {
      qual[t] = qv;
      const unsigned long long k = dkey(qv);
      minkey = k < minkey ? k : minkey;
    }
    if (need_bad) {
      const bool owned = owned_arr ? (owned_arr[t] != 0) : true;
      ++c_eval;
      mark_tet(qv, f, owned, P, c_bad);
      if (f != f_in) flags[t] = f;
    }
  }
  if (eig_any) atomicAdd(&st->n_eigen_fail, 1ull);
  warp_count_to(c_bad, &st->n_bad);
  warp_count_to(c_eval, &st->n_elems_eval);
  warp_count_to(c_err, &st->n_flag_err);
  if (want_q) {
    const unsigned long long m = warp_min_u64(minkey);
    if ((threadIdx.x & 31) == 0 && m != ~0ull) atomicMin(&st->min_q_key, m);
  }
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
__global__ void k_layer(int64_t np, int64_t npy, int64_t nv, const int32_t* __restrict__ prism_v, const int32_t* __restrict__ pyr_v,
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
// vstat: eigen-solver failures of the cached per-vertex pass; they belong to every sweep that uses those transforms
__global__ void k_init_stats(MagDevStats* st, const unsigned long long* vstat)
{
  MagDevStats z;
  memset(&z, 0, sizeof(z));
  z.max_len_bits = 0;                 // bits of +0.0 : getMaximumEdgeLength starts at 0.0
  z.min_q_key = dkey(1.0);            // getMinQuality starts at 1
  z.n_eigen_fail = vstat ? *vstat : 0ull;
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

// ------------------------------------------------------------------ chunk schedule (export time)
// key of a chunk = smallest vertex id any of its entities touches.  Handing the chunks out in key order makes the
// axis-edge / face-diagonal / body-diagonal families of a box mesh (or whatever families the caller's numbering has) sweep
// the vertex array together, so a vertex record is fetched from HBM once per sweep instead of once per family.
template <int NV>
__global__ void __launch_bounds__(kThreads)
k_chunk_keys(int64_t n, int64_t chunk_len, const int32_t* __restrict__ conn, int32_t* __restrict__ keys)
{
  __shared__ int sh[kThreads / 32];
  const int64_t lo = blockIdx.x * chunk_len, hi = (lo + chunk_len < n) ? lo + chunk_len : n;
  int m = 0x7fffffff;
  for (int64_t i = lo * NV + threadIdx.x; i < hi * NV; i += kThreads) { int v = conn[i] & kVidMask; m = v < m ? v : m; }
  m = __reduce_min_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < kThreads / 32; ++i) m = sh[i] < m ? sh[i] : m;
    keys[blockIdx.x] = m;
  }
}

// export-time validation: every vertex id of a connectivity array must lie in [0, nv).  One streaming pass per array
// (edges + tets of the 50 M-tet part: 1.3 GB, 0.2 ms), so a bad id becomes MAG_ERR_ARG instead of an out-of-bounds gather.
__global__ void __launch_bounds__(kThreads)
k_check_conn(int64_t n, int32_t* __restrict__ conn, int32_t nv, unsigned long long* __restrict__ bad)
{
  unsigned b = 0;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    if ((unsigned)conn[i] >= (unsigned)nv) { conn[i] = 0; ++b; }   // made harmless for kernels already queued behind this one
  b = __reduce_add_sync(0xffffffffu, b);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(bad, (unsigned long long)b);
}

// ------------------------------------------------------------------ mark bytes (mag_set_mark_bytes / mag_get_mark_bytes)
// the eight bits of an "ma_flags" word the sweep reads or writes, packed into one byte for the host link:
//   bits 0-3 = SPLIT, DONT_SPLIT, COLLAPSE, DONT_COLLAPSE (word bits 0-3); 4 = NEED_NOT_SPLIT (17); 5 = NEED_NOT_COLLAPSE (18);
//   6 = BAD_QUALITY (5); 7 = OK_QUALITY (6)
__device__ __forceinline__ int32_t mark_expand(unsigned b)
{
  return (int32_t)((b & 0xFu) | ((b & 0x10u) << 13) | ((b & 0x20u) << 13) | ((b & 0x40u) >> 1) | ((b & 0x80u) >> 1));
}
__device__ __forceinline__ unsigned mark_compress(int32_t f)
{
  const unsigned u = (unsigned)f;
  return (u & 0xFu) | ((u >> 13) & 0x10u) | ((u >> 13) & 0x20u) | ((u << 1) & 0x40u) | ((u << 1) & 0x80u);
}
// four entities per thread: one 32-bit load of bytes <-> one 128-bit store of words
__global__ void __launch_bounds__(kThreads)
k_marks_expand(int64_t n, const uint8_t* __restrict__ bytes, int32_t* __restrict__ words)
{
  const int64_t i = (blockIdx.x * (int64_t)kThreads + threadIdx.x) * 4;
  if (i + 3 < n) {
    const unsigned b = *reinterpret_cast<const unsigned*>(bytes + i);
    *reinterpret_cast<int4*>(words + i) = make_int4(mark_expand(b & 0xFF), mark_expand((b >> 8) & 0xFF), mark_expand((b >> 16) & 0xFF), mark_expand(b >> 24));
  } else {
    for (int64_t j = i; j < n; ++j) words[j] = mark_expand(bytes[j]);
  }
}
__global__ void __launch_bounds__(kThreads)
k_marks_compress(int64_t n, const int32_t* __restrict__ words, uint8_t* __restrict__ bytes)
{
  const int64_t i = (blockIdx.x * (int64_t)kThreads + threadIdx.x) * 4;
  if (i + 3 < n) {
    const int4 w = *reinterpret_cast<const int4*>(words + i);
    *reinterpret_cast<unsigned*>(bytes + i) = mark_compress(w.x) | (mark_compress(w.y) << 8) | (mark_compress(w.z) << 16) | (mark_compress(w.w) << 24);
  } else {
    for (int64_t j = i; j < n; ++j) bytes[j] = (uint8_t)mark_compress(words[j]);
  }
}

inline unsigned grid_for(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

#include "mag_rows.cuh"
#include "mag_lean.cuh"

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

// counts the out-of-range vertex ids of conn[0..n) into d_stats->n_bad_conn (read back by magk_conn_result, or with the
// statistics of mag_sweep_host) and replaces them by 0; runs on the compute stream BEFORE ownership is folded into the
// sign bits
int magk_check_conn(mag_ctx* c, int32_t* d_conn, int64_t n)
{
  if (n <= 0) return MAG_OK;
  int64_t g = (n + kThreads - 1) / kThreads;
  if (g > (int64_t)c->n_sms * 16) g = (int64_t)c->n_sms * 16;
  k_check_conn<<<(unsigned)g, kThreads, 0, c->stream>>>(n, d_conn, (int32_t)c->nv, &c->d_stats->n_bad_conn);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}
int magk_conn_begin(mag_ctx* c)
{
  MAG_CUDA(c, cudaMemsetAsync(&c->d_stats->n_bad_conn, 0, sizeof(unsigned long long), c->stream));
  return MAG_OK;
}
// synchronizes the compute stream
int magk_conn_result(mag_ctx* c, unsigned long long* bad)
{
  MAG_CUDA(c, cudaMemcpyAsync(&c->h_stats->n_bad_conn, &c->d_stats->n_bad_conn, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  *bad = c->h_stats->n_bad_conn;
  return MAG_OK;
}

// after the connectivity and owned arrays are uploaded (mag_set_mesh): fold ownership into the tet / edge vertex ids
int magk_fold_owned(mag_ctx* c)
{
  if (c->d_edge_owned && c->ne) {
    k_fold_owned<<<grid_for(c->ne), kThreads, 0, c->stream>>>(c->ne, 2, c->d_edge_owned, c->d_edge_v);
    c->n_launches++;
  }
  if (c->d_elem_owned && c->nt) {
    k_fold_owned<<<grid_for(c->nt), kThreads, 0, c->stream>>>(c->nt, 4, c->d_elem_owned + (c->np + c->npy), c->d_tet_v);
    c->n_launches++;
  }
  MAG_CUDA(c, cudaGetLastError());
  return MAG_OK;
}

// words[first .. first+n) <-> bytes[0 .. n) on the given stream (first must be a multiple of 4: cudaMalloc alignment + int4 access)
int magk_marks_expand(mag_ctx* c, cudaStream_t s, const uint8_t* d_bytes, int32_t* d_words, int64_t n)
{
  if (n <= 0) return MAG_OK;
  k_marks_expand<<<grid_for((n + 3) / 4), kThreads, 0, s>>>(n, d_bytes, d_words);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}
int magk_marks_compress(mag_ctx* c, cudaStream_t s, const int32_t* d_words, uint8_t* d_bytes, int64_t n)
{
  if (n <= 0) return MAG_OK;
  k_marks_compress<<<grid_for((n + 3) / 4), kThreads, 0, s>>>(n, d_words, d_bytes);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}

int magk_init_stats(mag_ctx* c)
{
  k_init_stats<<<1, 1, 0, c->stream>>>(c->d_stats, c->vertex_pass_valid ? c->d_vstat : nullptr);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}

// Q_v and det Q_v of every vertex (strict arithmetic in both modes).  They depend on the coordinates and the size field
// only, so they are computed when either changes (repack in mag_api.cu, mag_sweep_host) and reused by every sweep,
// cavity batch and sliver classification until the next change; eigen-solver failures are kept in c->d_vstat.
int magk_tet_winners(mag_ctx* c);
int magk_vertex_pass(mag_ctx* c)
{
  c->winners_valid = false;     // det Q_v is about to change: the winner bits of the tet slots (k_tet_winners) go with it
  if (c->nv == 0) return MAG_OK;
  MAG_CUDA(c, cudaMemsetAsync(c->d_vstat, 0, sizeof(unsigned long long), c->stream));
  unsigned g = grid_for(c->nv);
  switch (c->kind) {
    case MAG_KIND_IDENTITY: k_vertex_pass<MAG_KIND_IDENTITY><<<g, kThreads, 0, c->stream>>>(c->nv, c->dim, c->d_vedge, c->d_vpos, c->d_vq, c->d_vstat); break;
    case MAG_KIND_ISO: k_vertex_pass<MAG_KIND_ISO><<<g, kThreads, 0, c->stream>>>(c->nv, c->dim, c->d_vedge, c->d_vpos, c->d_vq, c->d_vstat); break;
    case MAG_KIND_ANISO: k_vertex_pass<MAG_KIND_ANISO><<<g, kThreads, 0, c->stream>>>(c->nv, c->dim, c->d_vedge, c->d_vpos, c->d_vq, c->d_vstat); break;
    case MAG_KIND_LOGM: k_vertex_pass<MAG_KIND_LOGM><<<g, kThreads, 0, c->stream>>>(c->nv, c->dim, c->d_vedge, c->d_vpos, c->d_vq, c->d_vstat); break;
    default: return mag_fail(c, MAG_ERR_ARG, "no size field set");
  }
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  c->vqu_kind = MAG_KIND_NONE;
  if (c->ne && (c->kind == MAG_KIND_ANISO || c->kind == MAG_KIND_LOGM)) {
    if (c->kind == MAG_KIND_ANISO) k_vertex_uniform<MAG_KIND_ANISO><<<g, kThreads, 0, c->stream>>>(c->nv, c->d_vedge, c->d_vqu);
    else k_vertex_uniform<MAG_KIND_LOGM><<<g, kThreads, 0, c->stream>>>(c->nv, c->d_vedge, c->d_vqu);
    MAG_CUDA(c, cudaGetLastError());
    c->n_launches++;
    c->vqu_kind = c->kind;
  }
  c->vertex_pass_valid = true;
  // the winner bits of the tet slots follow det Q_v (a no-op until the row layout exists; launch_tet_rows_w checks again)
  return magk_tet_winners(c);
}

template <int KIND>
static int launch_tris(mag_ctx* c, const SweepParams& P, bool fast)
{
  unsigned g = grid_for(c->ntri);
  const unsigned cap = (unsigned)c->n_sms * 8;
  if (g > cap) g = cap;
  if (fast) k_tris<KIND, true><<<g, kThreads, 0, c->stream>>>(c->ntri, c->nv, c->d_tri_v, c->d_vpos, c->d_vq, c->d_vedge, c->d_elem_owned, c->d_elem_flags, c->d_qual, P, c->d_stats, c->d_near_elem);
  else k_tris<KIND, false><<<g, kThreads, 0, c->stream>>>(c->ntri, c->nv, c->d_tri_v, c->d_vpos, c->d_vq, c->d_vedge, c->d_elem_owned, c->d_elem_flags, c->d_qual, P, c->d_stats, c->d_near_elem);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}

// persistent grids: resident blocks per SM x number of SMs (queried once per context)
static int blocks_per_sm(mag_ctx* c, const void* kernel, int threads)
{
  auto it = c->occupancy.find(kernel);
  if (it != c->occupancy.end()) return it->second;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  c->occupancy[kernel] = per_sm;
  return per_sm;
}
static unsigned persistent_grid(mag_ctx* c, const void* kernel, int64_t n, int threads, int chunk)
{
  const int per_sm = blocks_per_sm(c, kernel, threads);
  int64_t g = (int64_t)per_sm * c->n_sms;
  const int64_t chunks = (n + chunk - 1) / chunk;
  if (g > chunks) g = chunks;
  return (unsigned)(g < 1 ? 1 : g);
}

static EdgeParams edge_params(const mag_ctx* c, const SweepParams& P, bool zero_in)
{
  EdgeParams E;
  E.vqu = (c->vertex_pass_valid && c->vqu_kind == c->kind && c->vqu_kind != MAG_KIND_NONE) ? c->d_vqu : nullptr;
  E.zero_in = zero_in ? 1 : 0;
  E.reeval = P.reeval;
  const bool do_split = P.ops & MAG_OP_MARK_SPLIT, do_coll = P.ops & MAG_OP_MARK_COLLAPSE;
  E.off_bits = (do_split ? 0 : MAG_DONT_SPLIT) | (do_coll ? 0 : MAG_DONT_COLLAPSE);
  E.err_mask = (do_split ? MAG_SPLIT : 0) | (do_coll ? MAG_COLLAPSE : 0);
  E.want_len = (P.ops & MAG_OP_LENGTHS) ? 1 : 0;
  E.ops = P.ops;
  E.max_len = P.max_len;
  E.min_len = P.min_len;
  E.tol_max = std::isfinite(P.max_len) ? MAG_NEAR_REL * std::fabs(P.max_len) : -1.0;
  E.tol_min = std::isfinite(P.min_len) ? MAG_NEAR_REL * std::fabs(P.min_len) : -1.0;
  return E;
}

// [first, first + n) = the whole part with the chunk schedule, or a sub-range in natural order (mag_sweep_host)
struct Range { int64_t first, n; bool whole; };

template <int KIND, bool FAST, bool VERT>
static int launch_edges_t(mag_ctx* c, const SweepParams& P, const Range& r)
{
  constexpr int kEdgeThreads = EdgeCfg<KIND, FAST>::T;
  // the grid is sized for the edge chunks plus, with VERT, the vertex chunks that ride along
  const int64_t work = r.n + (VERT ? (c->nv + kVertChunk - 1) / kVertChunk * (int64_t)kEdgeChunk : 0);
  const unsigned g = persistent_grid(c, (const void*)k_edges<KIND, FAST, VERT>, work, kEdgeThreads, kEdgeChunk);
  const VertArgs V{(int32_t)c->nv, c->dim, c->d_vpos, c->d_vq};
  k_edges<KIND, FAST, VERT><<<g, kEdgeThreads, 0, c->stream>>>((int32_t)r.n, reinterpret_cast<const int2*>(c->d_edge_v) + r.first, c->d_vedge,
                                                               c->d_edge_flags + r.first, c->d_len + r.first, edge_params(c, P, r.whole && c->edge_flags_zero), c->d_stats,
                                                               c->d_near_edge, r.whole ? c->d_edge_order : nullptr, (int32_t)r.first, V);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}
template <int KIND>
static int launch_edges(mag_ctx* c, const SweepParams& P, bool fast, const Range& r, bool with_vertex_pass)
{
#if MAG_FUSE_VERTEX
  if (fast && with_vertex_pass) return launch_edges_t<KIND, true, true>(c, P, r);
#endif
  (void)with_vertex_pass;
  return fast ? launch_edges_t<KIND, true, false>(c, P, r) : launch_edges_t<KIND, false, false>(c, P, r);
}

static TetParams tet_params(const SweepParams& P, bool zero_in)
{
  TetParams T;
  T.zero_in = zero_in ? 1 : 0;
  T.ops = P.ops;
  T.do_bad = (P.ops & MAG_OP_MARK_BAD) ? 1 : 0;
  T.want_q = (P.ops & MAG_OP_QUALITIES) ? 1 : 0;
  T.use_max = P.use_max;
  T.good_q = P.good_q;
  T.tol_q = std::isfinite(P.good_q) ? MAG_NEAR_REL * std::fabs(P.good_q) : -1.0;
  return T;
}

template <int KIND, bool FAST, bool USE_MAX>
static int launch_tets_t(mag_ctx* c, const SweepParams& P, const Range& r)
{
  constexpr int kTetThreads = TetCfg<FAST>::T;
  const unsigned g = persistent_grid(c, (const void*)k_tets<KIND, FAST, USE_MAX>, r.n, kTetThreads, kTetChunk);
  k_tets<KIND, FAST, USE_MAX><<<g, kTetThreads, 0, c->stream>>>((int32_t)r.n, (int32_t)(c->np + c->npy + r.first), c->nv,
                                                                 reinterpret_cast<const int4*>(c->d_tet_v) + r.first,
                                                                 c->d_vpos, c->d_vq, c->d_vedge, c->d_elem_flags, c->d_qual, tet_params(P, r.whole && c->elem_flags_zero),
                                                                 c->d_stats, c->d_near_elem, r.whole ? c->d_tet_order : nullptr);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}
template <int KIND>
static int launch_tets(mag_ctx* c, const SweepParams& P, bool fast, const Range& r)
{
  if (P.use_max) return fast ? launch_tets_t<KIND, true, true>(c, P, r) : launch_tets_t<KIND, false, true>(c, P, r);
  return fast ? launch_tets_t<KIND, true, false>(c, P, r) : launch_tets_t<KIND, false, false>(c, P, r);
}

static int launch_edges_kind(mag_ctx* c, const SweepParams& P, bool fast, const Range& r, bool with_vertex_pass = false)
{
  switch (c->kind) {
    case MAG_KIND_IDENTITY: return launch_edges<MAG_KIND_IDENTITY>(c, P, fast, r, with_vertex_pass);
    case MAG_KIND_ISO: return launch_edges<MAG_KIND_ISO>(c, P, fast, r, with_vertex_pass);
    case MAG_KIND_ANISO: return launch_edges<MAG_KIND_ANISO>(c, P, fast, r, with_vertex_pass);
    default: return launch_edges<MAG_KIND_LOGM>(c, P, fast, r, false);   // the eigen-solver twice in one kernel does not fit
  }
}
static int launch_tets_kind(mag_ctx* c, const SweepParams& P, bool fast, const Range& r)
{
  switch (c->kind) {
    case MAG_KIND_IDENTITY: return launch_tets<MAG_KIND_IDENTITY>(c, P, fast, r);
    case MAG_KIND_ISO: return launch_tets<MAG_KIND_ISO>(c, P, fast, r);
    case MAG_KIND_ANISO: return launch_tets<MAG_KIND_ANISO>(c, P, fast, r);
    default: return launch_tets<MAG_KIND_LOGM>(c, P, fast, r);
  }
}
static SweepParams sweep_params(const mag_ctx* c, uint32_t ops, double max_len, double min_len, double good_q, int use_max)
{
  SweepParams P{ops, max_len, min_len, good_q, use_max};
  if (c->kind == MAG_KIND_IDENTITY) {
    // IdentitySizeField::shouldSplit / shouldCollapse are constant false (maSize.cc:64-72): no length exceeds +inf
    // ma::UniformRefiner (maSize.h:75-85) answers shouldSplit with a constant true: every length exceeds -inf
    P.max_len = c->uniform_refiner ? -INFINITY : INFINITY;
    P.min_len = -INFINITY;
  }
  return P;
}

// sub-range sweeps of the streaming entry point (mag_sweep_host): the slice's connectivity and flag words are already on
// the device; entities are taken in natural order; the work-distribution ticket is reset for every launch
int magk_edges_range(mag_ctx* c, uint32_t ops, double max_len, double min_len, int fp_mode, int64_t first, int64_t n)
{
  if (n <= 0) return MAG_OK;
  if (c->d_edge_owned) {
    k_fold_owned<<<grid_for(n), kThreads, 0, c->stream>>>(n, 2, c->d_edge_owned + first, c->d_edge_v + 2 * first);
    c->n_launches++;
  }
  if (!(ops & (MAG_OP_LENGTHS | MAG_OP_MARK_SPLIT | MAG_OP_MARK_COLLAPSE))) return MAG_OK;   // export only
  MAG_CUDA(c, cudaMemsetAsync(&c->d_stats->edge_chunk, 0, sizeof(unsigned long long), c->stream));
  return launch_edges_kind(c, sweep_params(c, ops, max_len, min_len, 0.0, 1), fp_mode == MAG_FP_FAST, Range{first, n, false});
}
int magk_tets_range(mag_ctx* c, uint32_t ops, double good_q, int use_max, int fp_mode, int64_t first, int64_t n)
{
  if (n <= 0) return MAG_OK;
  if (c->d_elem_owned) {
    k_fold_owned<<<grid_for(n), kThreads, 0, c->stream>>>(n, 4, c->d_elem_owned + (c->np + c->npy) + first, c->d_tet_v + 4 * first);
    c->n_launches++;
  }
  if (!(ops & (MAG_OP_QUALITIES | MAG_OP_MARK_BAD))) return MAG_OK;
  MAG_CUDA(c, cudaMemsetAsync(&c->d_stats->elem_chunk, 0, sizeof(unsigned long long), c->stream));
  return launch_tets_kind(c, sweep_params(c, ops, 0.0, 0.0, good_q, use_max), fp_mode == MAG_FP_FAST, Range{first, n, false});
}
int magk_length_sum(mag_ctx* c)
{
  // ma::getAverageEdgeLength (maSize.cc:654-671) adds up EVERY edge a part holds, owned or not, before the PCU Add
  k_sum_lengths<<<MAG_SUM_BLOCKS, kThreads, 0, c->stream>>>(c->ne, c->d_len, nullptr, c->d_block_sums);
  k_finish_sum<<<1, kThreads, 0, c->stream>>>((int64_t)MAG_SUM_BLOCKS, c->d_block_sums, c->d_stats);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches += 2;
  return MAG_OK;
}

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
namespace {
__global__ void k_iota(int64_t n, int32_t* __restrict__ out)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)i;
}
// export-time temporaries: freed on every exit path
// Temporaries of an export, from the device's stream-ordered pool (mag_create keeps freed blocks in the pool instead of
// returning them to the driver): with plain cudaMalloc / cudaFree the ~25 temporaries of one export cost 35 - 700 ms on a
// 663 k-tet part (r2o trace), every free being a device-wide synchronisation and an unmap; from the pool the same export is a
// few milliseconds, every time.
// Blocks of kPoolMax bytes or more keep cudaMalloc / cudaFree: for them the driver call is a small part of the work that
// follows, and a pool grown by the gigabytes of a 50 M-tet export would hold them for the life of the process.
constexpr size_t kPoolMax = (size_t)64 << 20;
cudaError_t temp_alloc(void** p, size_t bytes, cudaStream_t stream, bool* pooled)
{
  *pooled = bytes < kPoolMax;
  return *pooled ? cudaMallocAsync(p, bytes, stream) : cudaMalloc(p, bytes);
}
void temp_free(void* p, cudaStream_t stream, bool pooled)
{
  if (!p) return;
  if (pooled) cudaFreeAsync(p, stream); else cudaFree(p);
}
struct Scratch {
  cudaStream_t stream;
  std::vector<std::pair<void*, bool> > ptrs;
  explicit Scratch(cudaStream_t s) : stream(s) {}
  ~Scratch() { for (auto& p : ptrs) temp_free(p.first, stream, p.second); }
  template <class T> cudaError_t get(T*& p, size_t count)
  {
    p = nullptr;
    bool pooled = true;
    cudaError_t e = temp_alloc((void**)&p, (count ? count : 1) * sizeof(T), stream, &pooled);
    if (e == cudaSuccess) ptrs.push_back(std::make_pair((void*)p, pooled));
    return e;
  }
};
int bits_for(int64_t n) { int b = 1; while (b < 31 && ((int64_t)1 << b) <= n) ++b; return b; }
int sort_pairs(mag_ctx* c, Scratch& S, const int32_t* k_in, int32_t* k_out, const int32_t* v_in, int32_t* v_out, int64_t n, int end_bit)
{
  size_t tmp_bytes = 0;
  void* d_tmp = nullptr;
  MAG_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, v_out, (int)n, 0, end_bit, c->stream));
  MAG_CUDA(c, S.get(reinterpret_cast<char*&>(d_tmp), tmp_bytes));
  MAG_CUDA(c, cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, k_in, k_out, v_in, v_out, (int)n, 0, end_bit, c->stream));
  return MAG_OK;
}
int exclusive_scan(mag_ctx* c, Scratch& S, const int32_t* in, int32_t* out, int64_t n)
{
  size_t tmp_bytes = 0;
  void* d_tmp = nullptr;
  MAG_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (int)n, c->stream));
  MAG_CUDA(c, S.get(reinterpret_cast<char*&>(d_tmp), tmp_bytes));
  MAG_CUDA(c, cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, in, out, (int)n, c->stream));
  return MAG_OK;
}
void free_rows(MagRows& r, cudaStream_t stream)
{
  // (allocated from the stream-ordered pool on the context's compute stream, the only stream that ever reads them)
  temp_free(r.d_anchor, stream, r.pooled[0]);
  temp_free(r.d_slice_off, stream, r.pooled[1]);
  temp_free(r.d_slots, stream, r.pooled[2]);
  r.d_anchor = r.d_slice_off = r.d_slots = nullptr;
  r.n_rows = r.n_slices = r.n_slots = 0;
  r.valid = false;
}
} // namespace
// builds c->d_edge_order / c->d_tet_order (legacy tile kernels): chunk (tile) indices sorted by key on the device (LSD radix
// sort is stable, so equal keys keep the caller's order); nothing travels to the host
static int build_order(mag_ctx* c, int64_t n, int64_t chunk_len, const int32_t* d_conn, int nv_per, int32_t*& d_order, int64_t& n_chunks)
{
  n_chunks = (n + chunk_len - 1) / chunk_len;
  if (d_order) { MAG_CUDA(c, cudaFreeAsync(d_order, c->stream)); d_order = nullptr; }
  if (n_chunks == 0) return MAG_OK;
  Scratch S(c->stream);
  int32_t *d_keys = nullptr, *d_keys_out = nullptr, *d_idx = nullptr;
  MAG_CUDA(c, S.get(d_keys, (size_t)n_chunks));
  MAG_CUDA(c, S.get(d_keys_out, (size_t)n_chunks));
  MAG_CUDA(c, S.get(d_idx, (size_t)n_chunks));
  MAG_CUDA(c, cudaMallocAsync((void**)&d_order, (size_t)n_chunks * 4, c->stream));
  if (nv_per == 2) k_chunk_keys<2><<<(unsigned)n_chunks, kThreads, 0, c->stream>>>(n, chunk_len, d_conn, d_keys);
  else k_chunk_keys<4><<<(unsigned)n_chunks, kThreads, 0, c->stream>>>(n, chunk_len, d_conn, d_keys);
  k_iota<<<grid_for(n_chunks), kThreads, 0, c->stream>>>(n_chunks, d_idx);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches += 2;
  int rc = sort_pairs(c, S, d_keys, d_keys_out, d_idx, d_order, n_chunks, 31);
  if (rc) return rc;
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  return MAG_OK;
}

// the anchor-row layout of one entity dimension (mag_rows.cuh), built on the device from the resident connectivity
// (ownership already folded into the sign bit of the first vertex id)
template <int NV>
static int build_rows(mag_ctx* c, int64_t n, const int32_t* d_conn, MagRows& rows)
{
  free_rows(rows, c->stream);
  if (NV == 4) c->winners_valid = false;
  if (n == 0 || c->nv == 0) { rows.valid = true; return MAG_OK; }
  const int64_t nv = c->nv;
  Scratch S(c->stream);
  int32_t *key = nullptr, *val = nullptr, *key2 = nullptr, *sorted_e = nullptr, *deg = nullptr, *start = nullptr, *nrows = nullptr, *rowstart = nullptr;
  MAG_CUDA(c, S.get(key, (size_t)n));
  MAG_CUDA(c, S.get(val, (size_t)n));
  MAG_CUDA(c, S.get(key2, (size_t)n));
  MAG_CUDA(c, S.get(sorted_e, (size_t)n));
  MAG_CUDA(c, S.get(deg, (size_t)nv + 1));
  MAG_CUDA(c, S.get(start, (size_t)nv + 1));
  MAG_CUDA(c, S.get(nrows, (size_t)nv + 1));
  MAG_CUDA(c, S.get(rowstart, (size_t)nv + 1));
  MAG_CUDA(c, cudaMemsetAsync(deg, 0, ((size_t)nv + 1) * 4, c->stream));
  k_row_keys<NV><<<grid_for(n), kThreads, 0, c->stream>>>(n, d_conn, key, val, deg);
  MAG_CUDA(c, cudaGetLastError());
  int rc;
  if ((rc = sort_pairs(c, S, key, key2, val, sorted_e, n, bits_for(nv)))) return rc;
  if ((rc = exclusive_scan(c, S, deg, start, nv + 1))) return rc;
  k_row_counts<<<grid_for(nv + 1), kThreads, 0, c->stream>>>(nv, deg, nrows);
  MAG_CUDA(c, cudaGetLastError());
  if ((rc = exclusive_scan(c, S, nrows, rowstart, nv + 1))) return rc;
  int32_t n_rows32 = 0;
  MAG_CUDA(c, cudaMemcpyAsync(&n_rows32, rowstart + nv, 4, cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  const int64_t R = n_rows32, nslices = (R + 31) / 32, Rpad = nslices * 32;
  int32_t *row_anchor = nullptr, *row_len = nullptr, *row_first = nullptr, *rkey = nullptr, *ridx = nullptr, *rkey2 = nullptr, *order = nullptr;
  MAG_CUDA(c, S.get(row_anchor, (size_t)R));
  MAG_CUDA(c, S.get(row_len, (size_t)R));
  MAG_CUDA(c, S.get(row_first, (size_t)R));
  MAG_CUDA(c, S.get(rkey, (size_t)R));
  MAG_CUDA(c, S.get(ridx, (size_t)R));
  MAG_CUDA(c, S.get(rkey2, (size_t)R));
  MAG_CUDA(c, S.get(order, (size_t)R));
  k_row_records<<<grid_for(nv), kThreads, 0, c->stream>>>(nv, deg, start, rowstart, row_anchor, row_len, row_first, rkey, ridx);
  MAG_CUDA(c, cudaGetLastError());
  if ((rc = sort_pairs(c, S, rkey, rkey2, ridx, order, R, bits_for(nv >> kRowWindowLog2) + 6))) return rc;
  MAG_CUDA(c, temp_alloc((void**)&rows.d_slice_off, ((size_t)nslices + 1) * 4, c->stream, &rows.pooled[1]));
  MAG_CUDA(c, temp_alloc((void**)&rows.d_anchor, (size_t)Rpad * 4, c->stream, &rows.pooled[0]));
  int32_t* width32 = nullptr;
  MAG_CUDA(c, S.get(width32, (size_t)nslices + 1));
  k_slice_width<<<grid_for((nslices + 1) * 32), kThreads, 0, c->stream>>>(R, nslices, order, row_len, width32);
  MAG_CUDA(c, cudaGetLastError());
  if ((rc = exclusive_scan(c, S, width32, rows.d_slice_off, nslices + 1))) return rc;
  // the scan runs in int32: check the total in 64 bits on the host side of the widths (each <= 32 * kRowMax)
  int32_t n_slots32 = 0;
  MAG_CUDA(c, cudaMemcpyAsync(&n_slots32, rows.d_slice_off + nslices, 4, cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (n_slots32 < 0 || (int64_t)n_slots32 < n) {
    free_rows(rows, c->stream);
    return mag_fail(c, MAG_ERR_ARG, "row layout: slot count overflows int32 (%lld entities)", (long long)n);
  }
  const int64_t n_slots = n_slots32;
  MAG_CUDA(c, temp_alloc((void**)&rows.d_slots, (size_t)n_slots * NV * 4, c->stream, &rows.pooled[2]));
  MAG_CUDA(c, cudaMemsetAsync(rows.d_slots, 0xFF, (size_t)n_slots * NV * 4, c->stream));
  k_slots_fill<NV><<<grid_for(Rpad), kThreads, 0, c->stream>>>(R, Rpad, order, row_anchor, row_len, row_first, sorted_e, d_conn,
                                                               rows.d_slice_off, rows.d_anchor, rows.d_slots);
  MAG_CUDA(c, cudaGetLastError());
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n_launches += 5;
  rows.n_rows = R;
  rows.n_slices = nslices;
  rows.n_slots = n_slots;
  rows.valid = true;
  return MAG_OK;
}

void magk_free_rows(mag_ctx* c) { free_rows(c->erows, c->stream); free_rows(c->trows, c->stream); }

int magk_build_schedule(mag_ctx* c)
{
  int rc;
  // the tile schedule of k_edges / k_tets (a few bytes per 256 entities) is always built: those kernels serve every sweep
  // the lean row kernels do not (incoming flag words, single marks, strict arithmetic, the log-Euclidean field)
  int64_t nch;
  if ((rc = build_order(c, c->ne, MAG_EDGE_TILE_SCHED ? (int64_t)kStrictThreads : (int64_t)kEdgeChunk, c->d_edge_v, 2, c->d_edge_order, nch))) return rc;
  if ((rc = build_order(c, c->nt, (int64_t)kTetChunk, c->d_tet_v, 4, c->d_tet_order, nch))) return rc;
  if (c->legacy_sweep) return MAG_OK;
  if ((rc = build_rows<2>(c, c->ne, c->d_edge_v, c->erows))) return rc;
  if ((rc = build_rows<4>(c, c->nt, c->d_tet_v, c->trows))) return rc;
  return magk_tet_winners(c);      // (a no-op until the per-vertex pass has run)
}

// ---- whole-part sweeps over the anchor rows
// Which kernel family serves a whole-part sweep (measured on B200, n = 203, r2c / r2d):
//   lean row kernels (mag_lean.cuh)  MAG_FP_FAST over all-zero incoming flag words, every output of the dimension requested,
//                                    max-Jacobian metric: 1.26 / 0.78 ms (edges / tets) against 1.35 / 0.80 for the tiles;
//   tile kernels (k_edges / k_tets)  everything else.  The log-Euclidean edge kernel stays with the tiles as well (its QR
//                                    iteration wants every register: 4.7 ms against 5.4 ms in a row kernel);
//   (general row kernels -- the lean ones plus incoming words and single marks -- were built and measured: 1.43 / 0.82 ms,
//    slower than the tiles, and removed.)
static bool lean_edges_ok(const mag_ctx* c, const SweepParams& P, bool fast)
{
  constexpr uint32_t kEdgeFull = MAG_OP_LENGTHS | MAG_OP_MARK_SPLIT | MAG_OP_MARK_COLLAPSE;
  return fast && c->edge_flags_zero && c->lean_sweep && (P.ops & kEdgeFull) == kEdgeFull && c->kind != MAG_KIND_LOGM;
}
static bool lean_tets_ok(const mag_ctx* c, const SweepParams& P, bool fast)
{
  constexpr uint32_t kElemFull = MAG_OP_QUALITIES | MAG_OP_MARK_BAD;
  return fast && (c->elem_flags_zero || c->tet_words_zero) && P.use_max && c->lean_sweep && (P.ops & kElemFull) == kElemFull;
}
static bool use_edge_rows(const mag_ctx* c, const SweepParams& P, bool fast)
{
  return !c->legacy_sweep && lean_edges_ok(c, P, fast);
}
static bool use_tet_rows(const mag_ctx* c, const SweepParams& P, bool fast)
{
  return !c->legacy_sweep && lean_tets_ok(c, P, fast);
}

// mag_sweep_reconciled runs the part-boundary exchange of the edge marks on a side stream under the element kernel.  A
// persistent element kernel that fills every SM (2 CTAs x 128 registers x 256 threads = the whole register file) leaves no SM
// on which the pack / NCCL / merge kernels of the exchange could start, so the "overlapped" exchange ran AFTER the element
// kernel (r2z, 8 GPUs: element phase 0.82 ms for a 0.67 ms kernel).  A few CTAs fewer leave half-empty SMs for it.
static int64_t overlap_reserve(const mag_ctx* c, int64_t g)
{
  if (!c->comm_pending) return 0;
  const int64_t r = c->n_sms / 16 > 4 ? c->n_sms / 16 : 4;
  return g > 4 * r ? r : 0;
}

// the lean kernels (mag_lean.cuh): MAG_FP_FAST sweeps over all-zero incoming flag words
template <int KIND>
static int launch_edge_rows_z(mag_ctx* c, const SweepParams& P)
{
  constexpr int T = EdgeLeanCfg<KIND>::T;
  const int per_sm = blocks_per_sm(c, (const void*)k_edge_rows_z<KIND>, T);
  int64_t g = (int64_t)per_sm * c->n_sms;
  const int64_t groups = (c->erows.n_slices + kEZGroup - 1) / kEZGroup;
  const int64_t need = (groups + T / 32 - 1) / (T / 32);
  if (g > need) g = need;
  k_edge_rows_z<KIND><<<(unsigned)(g < 1 ? 1 : g), T, 0, c->stream>>>(
      (int32_t)c->erows.n_slices, c->erows.d_anchor, c->erows.d_slice_off, reinterpret_cast<const int2*>(c->erows.d_slots), c->d_vedge,
      c->d_edge_flags, c->d_len, edge_params(c, P, true), c->d_stats, c->d_near_edge);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}
template <int KIND>
static int launch_tet_rows_z(mag_ctx* c, const SweepParams& P)
{
  constexpr int T = MAG_TZ_THREADS;
  const int per_sm = blocks_per_sm(c, (const void*)k_tet_rows_z<KIND>, T);
  int64_t g = (int64_t)per_sm * c->n_sms;
  const int64_t groups = (c->trows.n_slices + kTZGroup - 1) / kTZGroup;
  const int64_t need = (groups + T / 32 - 1) / (T / 32);
  if (g > need) g = need;
  g -= overlap_reserve(c, g);
  k_tet_rows_z<KIND><<<(unsigned)(g < 1 ? 1 : g), T, 0, c->stream>>>(
      (int32_t)c->trows.n_slices, c->trows.d_anchor, c->trows.d_slice_off, reinterpret_cast<const int4*>(c->trows.d_slots),
      (int32_t)(c->np + c->npy), c->nv, c->d_vpos, c->d_vq, c->d_vedge, c->d_elem_flags, c->d_qual, tet_params(P, true),
      c->d_stats, c->d_near_elem);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}

// the winner-in-slot tet kernel: the max-Jacobian vertex of every tet, written into its slot word after the per-vertex pass
static bool tet_winner_ok(const mag_ctx* c) { return c->tet_winner && c->dim == 3 && c->nt < ((int64_t)1 << kWinShift); }
int magk_tet_winners(mag_ctx* c)
{
  if (c->winners_valid || !tet_winner_ok(c) || c->legacy_sweep || !c->trows.valid || !c->vertex_pass_valid) return MAG_OK;
  if (c->trows.n_slices) {
    k_tet_winners<<<grid_for(c->trows.n_slices * 32), kThreads, 0, c->stream>>>((int32_t)c->trows.n_slices, c->trows.d_anchor, c->trows.d_slice_off,
                                                                               reinterpret_cast<int4*>(c->trows.d_slots), c->d_vpos);
    MAG_CUDA(c, cudaGetLastError());
    c->n_launches++;
  }
  c->winners_valid = true;
  return MAG_OK;
}
template <int KIND>
static int launch_tet_rows_w(mag_ctx* c, const SweepParams& P)
{
  int rc = magk_tet_winners(c);
  if (rc) return rc;
  if (!c->winners_valid) return mag_fail(c, MAG_ERR_ARG, "internal: tet winners not in place (rows %d, vertex pass %d)", (int)c->trows.valid, (int)c->vertex_pass_valid);
  constexpr int T = MAG_TZ_THREADS;
  const int per_sm = blocks_per_sm(c, (const void*)k_tet_rows_w<KIND>, T);
  int64_t g = (int64_t)per_sm * c->n_sms;
  const int64_t groups = (c->trows.n_slices + kTZGroup - 1) / kTZGroup;
  const int64_t need = (groups + T / 32 - 1) / (T / 32);
  if (g > need) g = need;
  g -= overlap_reserve(c, g);
  k_tet_rows_w<KIND><<<(unsigned)(g < 1 ? 1 : g), T, 0, c->stream>>>(
      (int32_t)c->trows.n_slices, c->trows.d_anchor, c->trows.d_slice_off, reinterpret_cast<const int4*>(c->trows.d_slots),
      (int32_t)(c->np + c->npy), c->nv, c->d_vpos, c->d_vq, c->d_vedge, c->d_elem_flags, c->d_qual, tet_params(P, true),
      c->d_stats, c->d_near_elem);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}

static int launch_edge_rows(mag_ctx* c, const SweepParams& P, bool)
{
  switch (c->kind) {
    case MAG_KIND_IDENTITY: return launch_edge_rows_z<MAG_KIND_IDENTITY>(c, P);
    case MAG_KIND_ISO: return launch_edge_rows_z<MAG_KIND_ISO>(c, P);
    default: return launch_edge_rows_z<MAG_KIND_ANISO>(c, P);     // lean_edges_ok excludes the log-Euclidean field
  }
}
static int launch_tet_rows(mag_ctx* c, const SweepParams& P, bool)
{
  if (tet_winner_ok(c)) {
    switch (c->kind) {
      case MAG_KIND_IDENTITY: return launch_tet_rows_w<MAG_KIND_IDENTITY>(c, P);
      case MAG_KIND_ISO: return launch_tet_rows_w<MAG_KIND_ISO>(c, P);
      case MAG_KIND_ANISO: return launch_tet_rows_w<MAG_KIND_ANISO>(c, P);
      default: return launch_tet_rows_w<MAG_KIND_LOGM>(c, P);
    }
  }
  switch (c->kind) {
    case MAG_KIND_IDENTITY: return launch_tet_rows_z<MAG_KIND_IDENTITY>(c, P);
    case MAG_KIND_ISO: return launch_tet_rows_z<MAG_KIND_ISO>(c, P);
    case MAG_KIND_ANISO: return launch_tet_rows_z<MAG_KIND_ANISO>(c, P);
    default: return launch_tet_rows_z<MAG_KIND_LOGM>(c, P);
  }
}

int magk_sweep(mag_ctx* c, uint32_t ops, double max_len, double min_len, double good_q, int use_max, int fp_mode)
{
  SweepParams P = sweep_params(c, ops, max_len, min_len, good_q, use_max);
  const bool fast = fp_mode != MAG_FP_STRICT;
  P.reeval = fp_mode == MAG_FP_FAST_LISTED ? 0 : 1;
  int rc;
  cudaEvent_t* tev = (c->t_used < c->t_slots) ? &c->tev[(size_t)4 * c->t_used] : nullptr;
  if (tev) MAG_CUDA(c, cudaEventRecord(tev[0], c->stream));
  const bool need_vertex = ops & (MAG_OP_QUALITIES | MAG_OP_MARK_BAD | MAG_OP_LAYER_CHECK);
  // the per-vertex transforms are normally in place (computed when the coordinates or the size field were set)
  if (need_vertex && !c->vertex_pass_valid && (rc = magk_vertex_pass(c))) return rc;
  if (tev) MAG_CUDA(c, cudaEventRecord(tev[1], c->stream));
  if (c->ne && (ops & (MAG_OP_LENGTHS | MAG_OP_MARK_SPLIT | MAG_OP_MARK_COLLAPSE))) {
    if ((rc = use_edge_rows(c, P, fast) ? launch_edge_rows(c, P, fast) : launch_edges_kind(c, P, fast, Range{0, c->ne, true}, false))) return rc;
    // a requested mark writes the flag word of EVERY edge when the incoming words are zero (nothing is skipped)
    if (ops & (MAG_OP_MARK_SPLIT | MAG_OP_MARK_COLLAPSE)) c->edge_flags_zero = false;
    if ((ops & MAG_OP_LENGTH_SUM) && (rc = magk_length_sum(c))) return rc;
  }
  if (tev) MAG_CUDA(c, cudaEventRecord(tev[2], c->stream));
  // mag_sweep_reconciled: the edge marks are final; their part-boundary exchange runs under the element kernels
  if (c->overlap_mask && (rc = magc_overlap_begin(c, c->overlap_mask))) return rc;
  if (ops & (MAG_OP_QUALITIES | MAG_OP_MARK_BAD)) {
    // only the tet kernel understands "all zero, not materialised" (it then writes every word it marks)
    if ((c->ntri || c->np + c->npy) && (rc = magi_materialize_flags(c))) return rc;
    if (c->nt) {
      if ((rc = use_tet_rows(c, P, fast) ? launch_tet_rows(c, P, fast) : launch_tets_kind(c, P, fast, Range{0, c->nt, true}))) return rc;
      if (ops & MAG_OP_MARK_BAD) c->elem_flags_zero = c->tet_words_zero = false;
    }
    if (c->ntri) {
      switch (c->kind) {
        case MAG_KIND_IDENTITY: rc = launch_tris<MAG_KIND_IDENTITY>(c, P, fast); break;
        case MAG_KIND_ISO: rc = launch_tris<MAG_KIND_ISO>(c, P, fast); break;
        case MAG_KIND_ANISO: rc = launch_tris<MAG_KIND_ANISO>(c, P, fast); break;
        default: rc = launch_tris<MAG_KIND_LOGM>(c, P, fast); break;
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
    k_layer<<<grid_for(c->np + c->npy), kThreads, 0, c->stream>>>(c->np, c->npy, c->nv, c->d_prism_v, c->d_pyr_v, c->d_vpos, c->d_layer_ok, c->d_layer_codes, c->d_stats);
    MAG_CUDA(c, cudaGetLastError());
    c->n_launches++;
  }
  if (tev) { MAG_CUDA(c, cudaEventRecord(tev[3], c->stream)); c->t_used++; }
  return MAG_OK;
}

template <int KIND>
static int launch_cavities(mag_ctx* c, bool fast, int64_t ncav, const int64_t* d_off, const int32_t* d_tv, int use_max, double* d_worst, double* d_qual)
{
  const int64_t blocks = (ncav + 7) / 8;
  const unsigned g = (unsigned)(blocks < (int64_t)c->n_sms * 16 ? (blocks < 1 ? 1 : blocks) : (int64_t)c->n_sms * 16);
  const int4* tv = reinterpret_cast<const int4*>(d_tv);
  if (fast) k_cavity_quality<KIND, true><<<g, 256, 0, c->stream>>>(ncav, d_off, tv, c->nv, c->d_vpos, c->d_vq, c->d_vedge, use_max, d_worst, d_qual, c->d_stats);
  else k_cavity_quality<KIND, false><<<g, 256, 0, c->stream>>>(ncav, d_off, tv, c->nv, c->d_vpos, c->d_vq, c->d_vedge, use_max, d_worst, d_qual, c->d_stats);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}
int magk_cavity_quality(mag_ctx* c, int fp_mode, int64_t ncav, const int64_t* d_off, const int32_t* d_tv, int use_max, double* d_worst, double* d_qual)
{
  const bool fast = fp_mode == MAG_FP_FAST;
  switch (c->kind) {
    case MAG_KIND_IDENTITY: return launch_cavities<MAG_KIND_IDENTITY>(c, fast, ncav, d_off, d_tv, use_max, d_worst, d_qual);
    case MAG_KIND_ISO: return launch_cavities<MAG_KIND_ISO>(c, fast, ncav, d_off, d_tv, use_max, d_worst, d_qual);
    case MAG_KIND_ANISO: return launch_cavities<MAG_KIND_ANISO>(c, fast, ncav, d_off, d_tv, use_max, d_worst, d_qual);
    default: return launch_cavities<MAG_KIND_LOGM>(c, fast, ncav, d_off, d_tv, use_max, d_worst, d_qual);
  }
}

// vertex -> tet incidence of the resident part (CSR: c->d_v2t_off [nv + 1], c->d_v2t [4 nt]); kept until the mesh changes
int magk_build_v2t(mag_ctx* c)
{
  if (c->v2t_valid) return MAG_OK;
  if (c->d_v2t_off) { MAG_CUDA(c, cudaFree(c->d_v2t_off)); c->d_v2t_off = nullptr; }      // (cudaFree takes blocks of either kind)
  if (c->d_v2t) { MAG_CUDA(c, cudaFree(c->d_v2t)); c->d_v2t = nullptr; }
  bool pooled_unused;
  MAG_CUDA(c, temp_alloc((void**)&c->d_v2t_off, ((size_t)c->nv + 1) * 4, c->stream, &pooled_unused));
  MAG_CUDA(c, temp_alloc((void**)&c->d_v2t, ((size_t)c->nt * 4 + 1) * 4, c->stream, &pooled_unused));
  Scratch S(c->stream);
  int32_t *cnt = nullptr, *cursor = nullptr;
  MAG_CUDA(c, S.get(cnt, (size_t)c->nv + 1));
  MAG_CUDA(c, S.get(cursor, (size_t)c->nv + 1));
  MAG_CUDA(c, cudaMemsetAsync(cnt, 0, ((size_t)c->nv + 1) * 4, c->stream));
  const int4* tv = reinterpret_cast<const int4*>(c->d_tet_v);
  if (c->nt) k_v2t_count<<<grid_for(c->nt), kThreads, 0, c->stream>>>(c->nt, tv, cnt);
  MAG_CUDA(c, cudaGetLastError());
  int rc;
  if ((rc = exclusive_scan(c, S, cnt, c->d_v2t_off, c->nv + 1))) return rc;
  MAG_CUDA(c, cudaMemcpyAsync(cursor, c->d_v2t_off, ((size_t)c->nv + 1) * 4, cudaMemcpyDeviceToDevice, c->stream));
  if (c->nt) k_v2t_fill<<<grid_for(c->nt), kThreads, 0, c->stream>>>(c->nt, tv, cursor, c->d_v2t);
  MAG_CUDA(c, cudaGetLastError());
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n_launches += 3;
  c->v2t_valid = true;
  return MAG_OK;
}
template <int KIND>
static int launch_collapse(mag_ctx* c, bool fast, int64_t ncand, const int32_t* d_edge, const uint8_t* d_end, int use_max,
                           double* d_new, double* d_old, int32_t* d_keep)
{
  const int64_t blocks = (ncand + 7) / 8;
  const unsigned g = (unsigned)(blocks < (int64_t)c->n_sms * 16 ? (blocks < 1 ? 1 : blocks) : (int64_t)c->n_sms * 16);
  const int2* ev = reinterpret_cast<const int2*>(c->d_edge_v);
  const int4* tv = reinterpret_cast<const int4*>(c->d_tet_v);
  if (fast) k_collapse_quality<KIND, true><<<g, 256, 0, c->stream>>>(ncand, d_edge, d_end, ev, c->d_v2t_off, c->d_v2t, tv, c->nv, c->d_vpos, c->d_vq, c->d_vedge, use_max, d_new, d_old, d_keep, c->d_stats);
  else k_collapse_quality<KIND, false><<<g, 256, 0, c->stream>>>(ncand, d_edge, d_end, ev, c->d_v2t_off, c->d_v2t, tv, c->nv, c->d_vpos, c->d_vq, c->d_vedge, use_max, d_new, d_old, d_keep, c->d_stats);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches++;
  return MAG_OK;
}
int magk_collapse_quality(mag_ctx* c, int fp_mode, int64_t ncand, const int32_t* d_edge, const uint8_t* d_end, int use_max,
                          double* d_new, double* d_old, int32_t* d_keep)
{
  const bool fast = fp_mode == MAG_FP_FAST;
  switch (c->kind) {
    case MAG_KIND_IDENTITY: return launch_collapse<MAG_KIND_IDENTITY>(c, fast, ncand, d_edge, d_end, use_max, d_new, d_old, d_keep);
    case MAG_KIND_ISO: return launch_collapse<MAG_KIND_ISO>(c, fast, ncand, d_edge, d_end, use_max, d_new, d_old, d_keep);
    case MAG_KIND_ANISO: return launch_collapse<MAG_KIND_ANISO>(c, fast, ncand, d_edge, d_end, use_max, d_new, d_old, d_keep);
    default: return launch_collapse<MAG_KIND_LOGM>(c, fast, ncand, d_edge, d_end, use_max, d_new, d_old, d_keep);
  }
}

double magk_key_to_double(unsigned long long k)
{
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double d;
  memcpy(&d, &b, 8);
  return d;
}
