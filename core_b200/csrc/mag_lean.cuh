// mag_lean.cuh -- the whole-part sweep kernels for the common case, written for instruction count.
//
// Included by mag_kernels.cu after mag_rows.cuh.  Same anchor-row layout, same results bit for bit as k_edge_rows /
// k_tet_rows (tests/test_gpu_parity.py::test_lean_kernels_equal_general), used when the sweep is
//     MAG_FP_FAST, incoming flag words all zero (the state after mag_set_flags(NULL, NULL) / ma::Adapt's constructor on a
//     mesh without layer elements), lengths + both edge marks (edges) / qualities + BAD_QUALITY with the max-Jacobian metric,
//     the reference's default (tets)
// -- a full marking sweep.  Everything else runs the general kernels.
//
// Why a second pair of kernels (ncu source page of the general ones, r2c, n = 203): 292 warp-instructions per edge and 301
// per tet, of which 141 / 100 on the fp64 pipe; the rest is flag logic for words that are known to be zero, 64-bit address
// arithmetic, register moves of the software pipeline (24 - 35 IMAD.MOV per entity: the prefetch buffers are rotated by
// copying) and loop-invariant loads ptxas re-issues every iteration.  With issue slots 50 - 56 % busy and the fp64 pipe 33 -
// 52 %, the path is bound by instruction issue under latency, not by HBM, so instructions are what has to go:
//   * the k loop is unrolled by two over PING-PONG prefetch buffers (no register copies);
//   * skip tests, incoming-word loads and error-flag tests are compiled out (the words are zero);
//   * the other end's record (edges) / the other three vertices' {x,y,z,det} (tets) of the NEXT slot row are requested before
//     this row is evaluated, the slot words two rows ahead, the next slice's header a slice ahead and its first slot words
//     two rows before the slice ends: a warp waits for memory once per slice (anchor record), not once per row;
//   * the transform of the max-Jacobian vertex stays in registers while consecutive tets of a row pick the same vertex
//     (the six Kuhn tets of a cell share two vertices; any tet fan around an edge does);
//   * tickets hand out groups of 2 - 4 slices: fewer atomics on one address, and the next slice is known without waiting.
#pragma once

#ifndef MAG_EZ_THREADS
#define MAG_EZ_THREADS 256
#endif
#ifndef MAG_EZ_BLOCKS
#define MAG_EZ_BLOCKS 2
#endif
#ifndef MAG_TZ_THREADS
#define MAG_TZ_THREADS 256
#endif
#ifndef MAG_TZ_BLOCKS
#define MAG_TZ_BLOCKS 2
#endif
#ifndef MAG_EZ_GROUP
#define MAG_EZ_GROUP 2   /* slices per ticket, edges: 2 -> 0.845 / 1.233 ms, 4 -> 0.875 / 1.259 ms (jittered / lattice, r2d) */
#endif
#ifndef MAG_TZ_GROUP
#define MAG_TZ_GROUP 4   /* tets: 4 -> 0.853 / 0.783 ms, 2 -> 0.880 / 0.811 ms */
#endif
constexpr int kEZGroup = MAG_EZ_GROUP, kTZGroup = MAG_TZ_GROUP;

// Groups of G consecutive slices per ticket, two tickets in flight: the id of the slice after the current one is
// known at the top of every slice (so its header can travel during the slice) without ever waiting for the atomic.
template <int kZGroup>
struct GroupWalk {
  int s, s_end, nx_end, raw;
  __device__ __forceinline__ void begin(unsigned long long* counter, int nslices)
  {
    const int r0 = SliceWalk::issue(counter);
    const int g = __shfl_sync(0xffffffffu, r0, 0);
    s = g * kZGroup;
    s_end = s + kZGroup < nslices ? s + kZGroup : nslices;
    nx_end = 0;
    raw = SliceWalk::issue(counter);
  }
  // top of a slice: the slice that follows (>= nslices: none)
  __device__ __forceinline__ int next_slice(unsigned long long* counter, int nslices)
  {
    if (s + 1 < s_end) return s + 1;
    const int g = __shfl_sync(0xffffffffu, raw, 0);   // requested a whole group ago
    raw = SliceWalk::issue(counter);
    const int f = g * kZGroup;
    nx_end = f + kZGroup < nslices ? f + kZGroup : nslices;
    return f < nslices ? f : nslices;
  }
  __device__ __forceinline__ void advance(int s_nx)
  {
    if (s + 1 >= s_end) s_end = nx_end;
    s = s_nx;
  }
};

template <int KIND>
__device__ __forceinline__ double edge_length_fast_p(const double* __restrict__ a, const double* __restrict__ b, int* eig_fail)
{
  if (KIND == MAG_KIND_IDENTITY) return magfa::edge_identity(a, b);
  if (KIND == MAG_KIND_ISO) return magfa::edge_iso(a, b);
  if (KIND == MAG_KIND_ANISO) return magfa::edge_aniso(a, b);
  return magfa::edge_logm(a, b, eig_fail);
}

template <int KIND> struct EdgeLeanCfg {
  static constexpr int T = MAG_EZ_THREADS, B = KIND == MAG_KIND_LOGM ? MAG_EROW_BLOCKS_LOGM : MAG_EZ_BLOCKS;
};

#ifndef MAG_VQU_PREFETCH
#define MAG_VQU_PREFETCH 1
#endif
#ifndef MAG_EZ_ASMEM
#define MAG_EZ_ASMEM 1   /* 1: the anchor record lives in shared memory (one 16-byte chunk plane per warp and chunk, read back with
                            conflict-free LDS.128) instead of 24 registers: the other end's record of the next row stays in flight */
#endif
// lengths + SPLIT + COLLAPSE of every edge, flag words written from zero
template <int KIND>
__global__ void __launch_bounds__(EdgeLeanCfg<KIND>::T, EdgeLeanCfg<KIND>::B)
k_edge_rows_z(int32_t nslices, const int32_t* __restrict__ anchor, const int32_t* __restrict__ slice_off, const int2* __restrict__ slots,
              const double* __restrict__ vedge, int32_t* __restrict__ flags, double* __restrict__ lengths, EdgeParams P,
              MagDevStats* st, int32_t* __restrict__ near_list)
{
  constexpr int N = EdgeRecs<KIND>::N, C = N / 2;
#if MAG_EZ_ASMEM
  __shared__ double2 sh_a[EdgeLeanCfg<KIND>::T / 32][C][32];
  double2 (*as)[32] = sh_a[threadIdx.x >> 5];
#endif
  const int lane = threadIdx.x & 31;
  const double max_len = P.max_len, min_len = P.min_len, tol_max = P.tol_max, tol_min = P.tol_min;
  const bool reeval = P.reeval != 0;
  unsigned c_split = 0, c_coll = 0, c_eval = 0;
  double maxlen = 0.0;                  // getMaximumEdgeLength starts at 0 and ignores NaN (maSize.cc:673-691)
  int eig_any = 0;
  const int2 kNone = make_int2(0, -1);
  // one slot: length of edge sl.y between the anchor (record a) and the other end (record b); flag word from zero
  auto item = [&](const int2 sl, const int k, const double* __restrict__ a_reg, const double* __restrict__ b, unsigned& nearmask) {
    const int e = sl.y;
    if (e < 0) return;
#if MAG_EZ_ASMEM
    double a[N];
#pragma unroll
    for (int i = 0; i < C; ++i) { const double2 t = as[i][lane]; a[2 * i] = t.x; a[2 * i + 1] = t.y; }
    (void)a_reg;
#else
    const double* a = a_reg;
#endif
    const double len = edge_length_fast_p<KIND>(a, b, &eig_any);
    const bool owned = sl.x >= 0;               // sign bit of the other vertex id = "not owned"
    st_stream(lengths + e, len);
    if (owned && len > maxlen) maxlen = len;
    const bool nr = fabs(len - max_len) <= tol_max || fabs(len - min_len) <= tol_min;
    nearmask |= (nr ? 1u : 0u) << k;
    if (!nr || !reeval) {                       // near ones are decided in strict arithmetic after the row (near_edges), or -- MAG_FP_FAST_LISTED -- here and only listed
      ++c_eval;
      const bool sp = len > max_len, co = len < min_len;
      c_split += (sp && owned) ? 1u : 0u;
      c_coll += (co && owned) ? 1u : 0u;
      st_stream(flags + e, (int32_t)((sp ? MAG_SPLIT : MAG_NEED_NOT_SPLIT) | (co ? MAG_COLLAPSE : MAG_NEED_NOT_COLLAPSE)));
    }
  };
  GroupWalk<kEZGroup> w;
  w.begin(&st->edge_chunk, nslices);
  int off = 0, off1 = 0, va = -1;
  int2 s0 = kNone, s1 = kNone;
  if (w.s < nslices) {
    off = __ldg(slice_off + w.s); off1 = __ldg(slice_off + w.s + 1); va = __ldg(anchor + (w.s << 5) + lane);
    s0 = ld_stream(slots + off + lane);
    if (off1 - off > 32) s1 = ld_stream(slots + off + lane + 32);
  }
  while (w.s < nslices) {
    const int s_nx = w.next_slice(&st->edge_chunk, nslices);
    int off_nx = 0, off1_nx = 0, va_nx = -1;
    if (s_nx < nslices) { off_nx = __ldg(slice_off + s_nx); off1_nx = __ldg(slice_off + s_nx + 1); va_nx = __ldg(anchor + (s_nx << 5) + lane); }
    const int K = (off1 - off) >> 5;
    const int2* sp = slots + off + lane;
    double b0[N], b1[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { b0[i] = 0.0; b1[i] = 0.0; }
#if MAG_EZ_ASMEM
    {
      const double2 z = make_double2(0.0, 0.0);
      const double2* pa = chunk_ptr<C>(vedge, 0, va < 0 ? 0 : va);
      __syncwarp();                              // every lane is done with the previous slice's anchor
#pragma unroll
      for (int i = 0; i < C; ++i) as[i][lane] = va >= 0 ? __ldg(pa + i * kVB) : z;
      __syncwarp();
    }
    const double* a = nullptr;
#else
    double a_r[N];
#pragma unroll
    for (int i = 0; i < N; ++i) a_r[i] = 0.0;
    if (va >= 0) load_half_rec<KIND>(vedge, va, a_r);
    const double* a = a_r;
#endif
    if (s0.y >= 0) load_half_rec<KIND>(vedge, s0.x & kVidMask, b0);
    int2 n0 = kNone, n1 = kNone;                  // first two slot words of the next slice
    bool have_next = false;
    unsigned nearmask = 0;
    for (int k = 0; k < K; k += 2) {
      if (k == 2 && s_nx < nslices) {             // the next slice's header has arrived by now
        n0 = ld_stream(slots + off_nx + lane);
        if (off1_nx - off_nx > 32) n1 = ld_stream(slots + off_nx + lane + 32);
        have_next = true;
      }
      int2 s2 = kNone, s3 = kNone;
      if (k + 2 < K) s2 = ld_stream(sp + (k + 2) * 32);
#if MAG_SLOT_PREFETCH_E
      {  // two rows (this iteration's worth) of slot words, MAG_SLOT_PREFETCH_E rows ahead: in this slice, or the first rows of the next
        const int kp = k + MAG_SLOT_PREFETCH_E;
        if (kp + 1 < K) { if (lane < 4) prefetch_l2(reinterpret_cast<const char*>(sp - lane + kp * 32) + lane * 128); }
        else {
          const int kq = kp - K < 0 ? 0 : kp - K;
          if (s_nx < nslices && off_nx + (kq + 2) * 32 <= off1_nx && lane < 4) prefetch_l2(reinterpret_cast<const char*>(slots + off_nx + kq * 32) + lane * 128);
        }
      }
#endif
      if (s1.y >= 0) load_half_rec<KIND>(vedge, s1.x & kVidMask, b1);
      item(s0, k, a, b0, nearmask);
      if (k + 3 < K) s3 = ld_stream(sp + (k + 3) * 32);
      if (s2.y >= 0) load_half_rec<KIND>(vedge, s2.x & kVidMask, b0);
      item(s1, k + 1, a, b1, nearmask);
      s0 = s2;
      s1 = s3;
    }
#if MAG_VQU_PREFETCH
    // a lane with a near-threshold edge: its anchor's Q_u (k_vertex_uniform) is asked for now, so that it travels while near_edges
    // fetches the slot word and the two records (inside the k loop the request costs the loop registers: 24 bytes of spills)
    if (nearmask && P.vqu && reeval) {
#pragma unroll
      for (int i = 0; i < 5; ++i) prefetch_l2(chunk_ptr<5>(P.vqu, i, va < 0 ? 0 : va));
    }
#endif
    if (!have_next && s_nx < nslices) {
      n0 = ld_stream(slots + off_nx + lane);
      if (off1_nx - off_nx > 32) n1 = ld_stream(slots + off_nx + lane + 32);
    }
    // near-threshold entities of this slice, one slot row at a time (see k_edge_rows)
    for (unsigned any = __reduce_or_sync(0xffffffffu, nearmask); any; any &= any - 1) {
      const int k = __ffs(any) - 1;
      const bool nr = (nearmask >> k) & 1u;
      const int2 q = nr ? __ldg(sp + k * 32) : kNone;
      const unsigned r = near_edges<KIND, true>(nr, q.y, va, q.x, 0, vedge, flags, lengths, P.ops, max_len, min_len, st, near_list, reeval, P.vqu);
      c_eval += r & 1u; c_split += (r >> 1) & 1u; c_coll += (r >> 2) & 1u; eig_any |= (int)(r >> 3);
    }
    w.advance(s_nx);
    off = off_nx; off1 = off1_nx; va = va_nx;
    s0 = n0; s1 = n1;
  }
  if (eig_any) atomicAdd(&st->n_eigen_fail, 1ull);
  warp_count_to(c_split, &st->n_split);
  warp_count_to(c_coll, &st->n_collapse);
  warp_count_to(c_eval, &st->n_edges_eval);
  const unsigned long long m = warp_max_u64((unsigned long long)__double_as_longlong(maxlen));
  if (lane == 0 && m) atomicMax(&st->max_len_bits, m);
}

// ------------------------------------------------------------------ tets
// (A second form was built and measured, r2e: both dependent gathers -- {z, det} two rows ahead, the winner's transform and {x, y}
//  one row ahead -- pipelined so that nothing is waited for inside a slice, at 168 registers = 3 x 128 threads per SM: 0.87 ms
//  against 0.78 ms for this one at 2 x 256.  Resident warps beat prefetch depth here: see DESIGN.md section 4, fp64 latency.)
// qualities + BAD_QUALITY of every tet.  slot = {o1 | not-owned << 31, o2, o3, tet index}; the anchor is the tet's first vertex (see k_tet_rows).  Per lane: the
// anchor's {x,y,z,det Q_v} for the whole row; the other three vertices' {x,y,z,det Q_v} one row ahead (ping-pong); the
// transform of the max-Jacobian vertex (getMetricWithMaxJacobean, maQuality.cc:83-108) re-read only when the vertex changes.
template <int KIND>
__global__ void __launch_bounds__(MAG_TZ_THREADS, MAG_TZ_BLOCKS)
k_tet_rows_z(int32_t nslices, const int32_t* __restrict__ anchor, const int32_t* __restrict__ slice_off, const int4* __restrict__ slots,
             int32_t elem_off, int64_t nv, const double* __restrict__ vpos, const double* __restrict__ vq, const double* __restrict__ vedge,
             int32_t* __restrict__ flags, double* __restrict__ qual, TetParams P, MagDevStats* st, int32_t* __restrict__ near_list)
{
  const int lane = threadIdx.x & 31;
  const double good_q = P.good_q, tol_q = P.tol_q;
  unsigned c_bad = 0, c_eval = 0;
  unsigned long long minkey = ~0ull;
  int eig_any = 0;
  flags += elem_off;
  qual += elem_off;
  const int4 kNone = make_int4(0, 0, 0, -1);
  M3 Q;
  double detQ = 0.0;
  int32_t q_of = -1;                              // vertex whose transform Q holds
#pragma unroll
  for (int i = 0; i < 9; ++i) Q.m[i / 3][i % 3] = 0.0;
  auto load_pos = [&](const int4& sl, double2* p) {   // p[0..5] = xy, zd of o1, o2, o3
    const int32_t v1 = sl.x & kVidMask;
    p[0] = __ldg(chunk_ptr<2>(vpos, 0, v1)); p[1] = __ldg(chunk_ptr<2>(vpos, 1, v1));
    p[2] = __ldg(chunk_ptr<2>(vpos, 0, sl.y)); p[3] = __ldg(chunk_ptr<2>(vpos, 1, sl.y));
    p[4] = __ldg(chunk_ptr<2>(vpos, 0, sl.z)); p[5] = __ldg(chunk_ptr<2>(vpos, 1, sl.z));
  };
  auto item = [&](const int4 sl, const int k, const int32_t va, const double2 a_xy, const double2 a_zd, const double2* __restrict__ p,
                  unsigned& nearmask) {
    const int t = sl.w;
    if (t < 0) return;
    const int4 tv = make_int4(va, sl.x & kVidMask, sl.y, sl.z);
    const int32_t vb = best_vertex(tv, a_zd.y, p[1].y, p[3].y, p[5].y);
    if (vb != q_of) { load_q(vq, vb, Q, detQ); q_of = vb; }
    const V3 x[4] = {V3{a_xy.x, a_xy.y, a_zd.x}, V3{p[0].x, p[0].y, p[1].x}, V3{p[2].x, p[2].y, p[3].x}, V3{p[4].x, p[4].y, p[5].x}};
    const double qv = magfa::tet_quality(x, Q, detQ);
    st_stream(qual + t, qv);
    const unsigned long long kq = dkey(qv);
    minkey = kq < minkey ? kq : minkey;
    const bool nr = fabs(qv - good_q) <= tol_q;
    nearmask |= (nr ? 1u : 0u) << k;
    if (!nr) {
      ++c_eval;
      const bool bad = qv < good_q;
      c_bad += (bad && sl.x >= 0) ? 1u : 0u;
      st_stream(flags + t, (int32_t)(bad ? MAG_BAD_QUALITY : MAG_OK_QUALITY));
    }
  };
  GroupWalk<kTZGroup> w;
  w.begin(&st->elem_chunk, nslices);
  int off = 0, off1 = 0, va = -1;
  int4 s0 = kNone, s1 = kNone;
  if (w.s < nslices) {
    off = __ldg(slice_off + w.s); off1 = __ldg(slice_off + w.s + 1); va = __ldg(anchor + (w.s << 5) + lane);
    s0 = ld_stream(slots + off + lane);
    if (off1 - off > 32) s1 = ld_stream(slots + off + lane + 32);
  }
  while (w.s < nslices) {
    const int s_nx = w.next_slice(&st->elem_chunk, nslices);
    int off_nx = 0, off1_nx = 0, va_nx = -1;
    if (s_nx < nslices) { off_nx = __ldg(slice_off + s_nx); off1_nx = __ldg(slice_off + s_nx + 1); va_nx = __ldg(anchor + (s_nx << 5) + lane); }
    const int K = (off1 - off) >> 5;
    const int4* sp = slots + off + lane;
    double2 a_xy = make_double2(0.0, 0.0), a_zd = a_xy, p0[6], p1[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { p0[i] = make_double2(0.0, 0.0); p1[i] = p0[i]; }
    if (va >= 0) { a_xy = __ldg(chunk_ptr<2>(vpos, 0, va)); a_zd = __ldg(chunk_ptr<2>(vpos, 1, va)); }
    if (s0.w >= 0) load_pos(s0, p0);
    int4 n0 = kNone, n1 = kNone;
    bool have_next = false;
    unsigned nearmask = 0;
    for (int k = 0; k < K; k += 2) {
      if (k == 2 && s_nx < nslices) {
        n0 = ld_stream(slots + off_nx + lane);
        if (off1_nx - off_nx > 32) n1 = ld_stream(slots + off_nx + lane + 32);
        have_next = true;
      }
      int4 s2 = kNone, s3 = kNone;
      if (k + 2 < K) s2 = ld_stream(sp + (k + 2) * 32);
#if MAG_SLOT_PREFETCH_T
      {  // two rows of slot words (2 x 512 bytes), MAG_SLOT_PREFETCH_T rows ahead
        const int kp = k + MAG_SLOT_PREFETCH_T;
        if (kp + 1 < K) { if (lane < 8) prefetch_l2(reinterpret_cast<const char*>(sp - lane + kp * 32) + lane * 128); }
        else {
          const int kq = kp - K < 0 ? 0 : kp - K;
          if (s_nx < nslices && off_nx + (kq + 2) * 32 <= off1_nx && lane < 8) prefetch_l2(reinterpret_cast<const char*>(slots + off_nx + kq * 32) + lane * 128);
        }
      }
#endif
      if (s1.w >= 0) load_pos(s1, p1);
      item(s0, k, va, a_xy, a_zd, p0, nearmask);
      if (k + 3 < K) s3 = ld_stream(sp + (k + 3) * 32);
      if (s2.w >= 0) load_pos(s2, p0);
      item(s1, k + 1, va, a_xy, a_zd, p1, nearmask);
      s0 = s2;
      s1 = s3;
    }
    if (!have_next && s_nx < nslices) {
      n0 = ld_stream(slots + off_nx + lane);
      if (off1_nx - off_nx > 32) n1 = ld_stream(slots + off_nx + lane + 32);
    }
    for (unsigned any = __reduce_or_sync(0xffffffffu, nearmask); any; any &= any - 1) {
      const int k = __ffs(any) - 1;
      const bool nr = (nearmask >> k) & 1u;
      const int4 q = nr ? __ldg(sp + k * 32) : kNone;
      const unsigned r = near_tets<KIND, true>(nr, q.w, elem_off, make_int4(va, q.x, q.y, q.z), 0, nv, vpos, vq, vedge, flags - elem_off,
                                               qual - elem_off, P.ops, good_q, P.use_max, st, near_list);
      c_eval += r & 1u; c_bad += (r >> 1) & 1u; eig_any |= (int)(r >> 2);
    }
    w.advance(s_nx);
    off = off_nx; off1 = off1_nx; va = va_nx;
    s0 = n0; s1 = n1;
  }
  if (eig_any) atomicAdd(&st->n_eigen_fail, 1ull);
  warp_count_to(c_bad, &st->n_bad);
  warp_count_to(c_eval, &st->n_elems_eval);
  const unsigned long long m = warp_min_u64(minkey);
  if (lane == 0 && m != ~0ull) atomicMin(&st->min_q_key, m);
}

// ------------------------------------------------------------------ tets, winner known from the slot word
// r2f / r2t profiles of k_tet_rows_z: 55 % of the warp time is "long scoreboard", and the largest single place is the first use
// of the winner's transform -- getMetricWithMaxJacobean (maQuality.cc:83-108) picks the vertex AFTER the four det Q_v have
// arrived, so the gather of its 80-byte transform starts when the row is already being evaluated.  Which vertex wins depends on
// the size field and the connectivity only, like det Q_v itself: k_tet_winners writes it (2 bits) into the slot word of every
// tet whenever the per-vertex pass has run (mag_set_metric_* / mag_set_coords), with the reference's own rule (strict >, first
// wins).  Here the slot word, read two rows ahead, names the winner, and its transform travels one row ahead like the
// coordinates: by cp.async into a per-warp shared-memory stage (no registers are held in flight; a lane whose winner is the
// previous row's keeps the registers it has), read back with conflict-free LDS.128 when the row is evaluated.
constexpr int kWinShift = 29;                      // slot.w = tet index | winner << 29 (k_tet_winners); parts of 2^29 tets or more use k_tet_rows_z
constexpr int32_t kTidMask = (1 << kWinShift) - 1;
__global__ void __launch_bounds__(kThreads)
k_tet_winners(int32_t nslices, const int32_t* __restrict__ anchor, const int32_t* __restrict__ slice_off, int4* __restrict__ slots,
              const double* __restrict__ vpos)
{
  const int64_t s = (blockIdx.x * (int64_t)kThreads + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= nslices) return;
  const int off = slice_off[s], K = (slice_off[s + 1] - off) >> 5;
  const int32_t va = anchor[(s << 5) + lane];
  const double d0 = va >= 0 ? __ldg(chunk_ptr<2>(vpos, 1, va)).y : 0.0;
  for (int k = 0; k < K; ++k) {
    int4* p = slots + off + 32 * k + lane;
    const int4 sl = *p;
    if (sl.w < 0) continue;
    const double d1 = __ldg(chunk_ptr<2>(vpos, 1, sl.x & kVidMask)).y, d2 = __ldg(chunk_ptr<2>(vpos, 1, sl.y)).y,
                 d3 = __ldg(chunk_ptr<2>(vpos, 1, sl.z)).y;
    const int32_t w = best_vertex(make_int4(0, 1, 2, 3), d0, d1, d2, d3);
    p->w = (sl.w & kTidMask) | (w << kWinShift);
  }
}

template <int KIND>
__global__ void __launch_bounds__(MAG_TZ_THREADS, MAG_TZ_BLOCKS)
k_tet_rows_w(int32_t nslices, const int32_t* __restrict__ anchor, const int32_t* __restrict__ slice_off, const int4* __restrict__ slots,
             int32_t elem_off, int64_t nv, const double* __restrict__ vpos, const double* __restrict__ vq, const double* __restrict__ vedge,
             int32_t* __restrict__ flags, double* __restrict__ qual, TetParams P, MagDevStats* st, int32_t* __restrict__ near_list)
{
  __shared__ double2 sh_q[MAG_TZ_THREADS / 32][2][5][32];       // the winner's transform of the row in flight / being evaluated
  const int lane = threadIdx.x & 31;
  double2 (*stage)[5][32] = sh_q[threadIdx.x >> 5];
  const double good_q = P.good_q, tol_q = P.tol_q;
  unsigned c_bad = 0, c_eval = 0;
  unsigned long long minkey = ~0ull;
  int eig_any = 0;
  flags += elem_off;
  qual += elem_off;
  const int4 kNone = make_int4(0, 0, 0, -1);
  M3 Q;
  int32_t q_of = -1;                              // vertex whose transform Q holds
  int32_t q_sent = -1;                            // winner of the row whose transform was requested last
#pragma unroll
  for (int i = 0; i < 9; ++i) Q.m[i / 3][i % 3] = 0.0;
  auto winner = [&](const int4& sl, const int32_t va) -> int32_t {
    const int w = (sl.w >> kWinShift) & 3;
    return w == 0 ? va : (w == 1 ? (sl.x & kVidMask) : (w == 2 ? sl.y : sl.z));
  };
  auto load_pos = [&](const int4& sl, double2* p) {   // p[0..5] = xy, z. of o1, o2, o3
    const int32_t v1 = sl.x & kVidMask;
    p[0] = __ldg(chunk_ptr<2>(vpos, 0, v1)); p[1] = __ldg(chunk_ptr<2>(vpos, 1, v1));
    p[2] = __ldg(chunk_ptr<2>(vpos, 0, sl.y)); p[3] = __ldg(chunk_ptr<2>(vpos, 1, sl.y));
    p[4] = __ldg(chunk_ptr<2>(vpos, 0, sl.z)); p[5] = __ldg(chunk_ptr<2>(vpos, 1, sl.z));
  };
  // one row's requests: coordinates into registers, the winner's transform into stage[buf] unless the previous row asked for the same
  auto request = [&](const int4& sl, const int32_t va, double2* p, const int buf) {
    if (sl.w >= 0) {
      load_pos(sl, p);
      const int32_t vb = winner(sl, va);
      if (vb != q_sent) {
        const double2* src = chunk_ptr<5>(vq, 0, vb);
#pragma unroll
        for (int i = 0; i < 5; ++i)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&stage[buf][i][lane])), "l"(src + i * kVB) : "memory");
        q_sent = vb;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto item = [&](const int4 sl, const int k, const int32_t va, const double2 a_xy, const double2 a_zd, const double2* __restrict__ p,
                  const int buf, unsigned& nearmask) {
    const int t = sl.w & kTidMask;
    if (sl.w < 0) return;
    const int32_t vb = winner(sl, va);
    if (vb != q_of) {
      const double2 q0 = stage[buf][0][lane], q1 = stage[buf][1][lane], q2 = stage[buf][2][lane], q3 = stage[buf][3][lane], q4 = stage[buf][4][lane];
      Q.m[0][0] = q0.x; Q.m[0][1] = q0.y; Q.m[0][2] = q1.x;
      Q.m[1][0] = q1.y; Q.m[1][1] = q2.x; Q.m[1][2] = q2.y;
      Q.m[2][0] = q3.x; Q.m[2][1] = q3.y; Q.m[2][2] = q4.x;
      q_of = vb;
    }
    const V3 x[4] = {V3{a_xy.x, a_xy.y, a_zd.x}, V3{p[0].x, p[0].y, p[1].x}, V3{p[2].x, p[2].y, p[3].x}, V3{p[4].x, p[4].y, p[5].x}};
    const double qv = magfa::tet_quality(x, Q, 0.0);
    st_stream(qual + t, qv);
    const unsigned long long kq = dkey(qv);
    minkey = kq < minkey ? kq : minkey;
    const bool nr = fabs(qv - good_q) <= tol_q;
    nearmask |= (nr ? 1u : 0u) << k;
    if (!nr) {
      ++c_eval;
      const bool bad = qv < good_q;
      c_bad += (bad && sl.x >= 0) ? 1u : 0u;
      st_stream(flags + t, (int32_t)(bad ? MAG_BAD_QUALITY : MAG_OK_QUALITY));
    }
  };
  GroupWalk<kTZGroup> w;
  w.begin(&st->elem_chunk, nslices);
  int off = 0, off1 = 0, va = -1;
  int4 s0 = kNone, s1 = kNone;
  if (w.s < nslices) {
    off = __ldg(slice_off + w.s); off1 = __ldg(slice_off + w.s + 1); va = __ldg(anchor + (w.s << 5) + lane);
    s0 = ld_stream(slots + off + lane);
    if (off1 - off > 32) s1 = ld_stream(slots + off + lane + 32);
  }
  while (w.s < nslices) {
    const int s_nx = w.next_slice(&st->elem_chunk, nslices);
    int off_nx = 0, off1_nx = 0, va_nx = -1;
    if (s_nx < nslices) { off_nx = __ldg(slice_off + s_nx); off1_nx = __ldg(slice_off + s_nx + 1); va_nx = __ldg(anchor + (s_nx << 5) + lane); }
    const int K = (off1 - off) >> 5;
    const int4* sp = slots + off + lane;
    double2 a_xy = make_double2(0.0, 0.0), a_zd = a_xy, p0[6], p1[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { p0[i] = make_double2(0.0, 0.0); p1[i] = p0[i]; }
    if (va >= 0) { a_xy = __ldg(chunk_ptr<2>(vpos, 0, va)); a_zd = __ldg(chunk_ptr<2>(vpos, 1, va)); }
    request(s0, va, p0, 0);
    int4 n0 = kNone, n1 = kNone;
    bool have_next = false;
    unsigned nearmask = 0;
    for (int k = 0; k < K; k += 2) {
      if (k == 2 && s_nx < nslices) {
        n0 = ld_stream(slots + off_nx + lane);
        if (off1_nx - off_nx > 32) n1 = ld_stream(slots + off_nx + lane + 32);
        have_next = true;
      }
      int4 s2 = kNone, s3 = kNone;
      if (k + 2 < K) s2 = ld_stream(sp + (k + 2) * 32);
      request(s1, va, p1, 1);                                    // row k + 1 (an empty request past the end of the slice)
      asm volatile("cp.async.wait_group 1;" ::: "memory");      // row k's transform has landed
      item(s0, k, va, a_xy, a_zd, p0, 0, nearmask);
      if (k + 3 < K) s3 = ld_stream(sp + (k + 3) * 32);
      request(s2, va, p0, 0);                                    // row k + 2
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      item(s1, k + 1, va, a_xy, a_zd, p1, 1, nearmask);
      s0 = s2;
      s1 = s3;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");        // (only empty requests are left)
    if (!have_next && s_nx < nslices) {
      n0 = ld_stream(slots + off_nx + lane);
      if (off1_nx - off_nx > 32) n1 = ld_stream(slots + off_nx + lane + 32);
    }
    for (unsigned any = __reduce_or_sync(0xffffffffu, nearmask); any; any &= any - 1) {
      const int k = __ffs(any) - 1;
      const bool nr = (nearmask >> k) & 1u;
      const int4 q = nr ? __ldg(sp + k * 32) : kNone;
      const unsigned r = near_tets<KIND, true>(nr, q.w & kTidMask, elem_off, make_int4(va, q.x, q.y, q.z), 0, nv, vpos, vq, vedge, flags - elem_off,
                                               qual - elem_off, P.ops, good_q, P.use_max, st, near_list);
      c_eval += r & 1u; c_bad += (r >> 1) & 1u; eig_any |= (int)(r >> 2);
    }
    w.advance(s_nx);
    off = off_nx; off1 = off1_nx; va = va_nx;
    s0 = n0; s1 = n1;
  }
  if (eig_any) atomicAdd(&st->n_eigen_fail, 1ull);
  warp_count_to(c_bad, &st->n_bad);
  warp_count_to(c_eval, &st->n_elems_eval);
  const unsigned long long m = warp_min_u64(minkey);
  if (lane == 0 && m != ~0ull) atomicMin(&st->min_q_key, m);
}
