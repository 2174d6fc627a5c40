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

// ------------------------------------------------------------------ the stream kernels
// ncu source page of the first lean kernels (r2d / r2e, n = 203): a third of the edge kernel's and half of the tet kernel's
// warp time was "long scoreboard" -- not at the gathers themselves, which were requested a row ahead, but at the TEST OF A SLOT
// WORD that had been loaded two rows earlier, and at the start of every slice (header -> slot words -> anchor -> first record:
// four dependent round trips for seven rows of work).  A warp has six scoreboard entries; a load that shares one with a younger
// load is only "complete" when the younger one is, so a deep register pipeline of loads of different ages stalls on its
// youngest member.  The stream kernels therefore keep exactly ONE kind of long-latency load in flight in the steady state --
// the next row's gather -- and obey one rule: whatever a long-latency load returned is consumed at a point where nothing
// younger has been issued.
//   * slot words and the anchor's data of a whole slice travel into per-warp shared memory with cp.async (its own completion
//     counter, no register, no scoreboard entry), for the NEXT slice while this one is evaluated: double buffer, each lane
//     reads back only what it copied itself (no barrier);
//   * a row is evaluated in phases: the gathered record is consumed first (aniso_pre: 21 numbers), the first Gauss point is
//     evaluated, THEN the next row's gather is issued -- also across a slice boundary -- into registers that are free by then,
//     and the second Gauss point and the square roots run while it travels;
//   * slice headers are requested a slice ahead and tickets a group ahead, and both are consumed at the first row of a slice
//     right after phase one, where no gather is in flight.
#ifndef MAG_EZ_GROUP
#define MAG_EZ_GROUP 2
#endif
#ifndef MAG_TZ_GROUP
#define MAG_TZ_GROUP 4
#endif
constexpr int kEZGroup = MAG_EZ_GROUP, kTZGroup = MAG_TZ_GROUP;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// successive slice ids for one warp: groups of G consecutive slices per ticket, the next ticket always requested a group ahead
template <int G>
struct TicketStream {
  int s, s_end, raw;
  __device__ __forceinline__ int first(unsigned long long* counter, int n)
  {
    const int r0 = SliceWalk::issue(counter);
    const int g = __shfl_sync(0xffffffffu, r0, 0);
    raw = SliceWalk::issue(counter);
    s = g * G;
    s_end = s + G < n ? s + G : n;
    return s < n ? s : n;
  }
  __device__ __forceinline__ int next(unsigned long long* counter, int n)
  {
    if (s + 1 < s_end) return ++s;
    const int g = __shfl_sync(0xffffffffu, raw, 0);   // requested a whole group ago
    raw = SliceWalk::issue(counter);
    s = g * G;
    s_end = s + G < n ? s + G : n;
    return s < n ? s : n;
  }
};

// the phases of one edge, per size-field kind (mag_math_fast.cuh: aniso_pre / aniso_point / half_sum_sqrt_ratios); the other
// kinds have four-number records and nothing worth splitting: phase one copies, the last phase evaluates
template <int KIND> struct EdgePre { double a[4], b[4]; };
template <> struct EdgePre<MAG_KIND_ANISO> { magfa::AnisoPre p; double np, dp; };
template <int KIND>
__device__ __forceinline__ void edge_pre(const double* __restrict__ a, const double* __restrict__ b, EdgePre<KIND>& e)
{
#pragma unroll
  for (int i = 0; i < 4; ++i) { e.a[i] = a[i]; e.b[i] = b[i]; }
}
template <>
__device__ __forceinline__ void edge_pre<MAG_KIND_ANISO>(const double* __restrict__ a, const double* __restrict__ b, EdgePre<MAG_KIND_ANISO>& e)
{
  magfa::aniso_pre(a, b, e.p);
}
// the part of the evaluation that runs BEFORE the next row's gather is issued (first Gauss point)
template <int KIND> __device__ __forceinline__ void edge_mid(EdgePre<KIND>&) {}
template <>
__device__ __forceinline__ void edge_mid<MAG_KIND_ANISO>(EdgePre<MAG_KIND_ANISO>& e)
{
  magfa::aniso_point(e.p, 0, e.np, e.dp);
  magfa::aniso_fence(e.p, e.np, e.dp);
}
template <int KIND>
__device__ __forceinline__ double edge_post(const EdgePre<KIND>& e)
{
  return KIND == MAG_KIND_ISO ? magfa::edge_iso(e.a, e.b) : magfa::edge_identity(e.a, e.b);
}
template <>
__device__ __forceinline__ double edge_post<MAG_KIND_ANISO>(const EdgePre<MAG_KIND_ANISO>& e)
{
  double nm, dm;
  magfa::aniso_point(e.p, 1, nm, dm);
  return magfa::half_sum_sqrt_ratios(e.np, e.dp, nm, dm);
}

template <int KIND> struct EdgeLeanCfg {
  static constexpr int T = MAG_EZ_THREADS, B = MAG_EZ_BLOCKS;
  static constexpr int C = EdgeRecs<KIND>::N / 2;
  struct WarpBuf { double2 a[C][32]; int2 sl[kRowMax][32]; };   // one slice: the anchors' records and every slot word, lane-major
  static constexpr size_t kSmem = sizeof(WarpBuf) * 2 * (T / 32);
};

// lengths + SPLIT + COLLAPSE of every edge, flag words written from zero (MAG_FP_FAST; identity / iso / aniso fields)
template <int KIND>
__global__ void __launch_bounds__(EdgeLeanCfg<KIND>::T, EdgeLeanCfg<KIND>::B)
k_edge_rows_z(int32_t nslices, const int32_t* __restrict__ anchor, const int32_t* __restrict__ slice_off, const int2* __restrict__ slots,
              const double* __restrict__ vedge, int32_t* __restrict__ flags, double* __restrict__ lengths, EdgeParams P,
              MagDevStats* st, int32_t* __restrict__ near_list)
{
  constexpr int N = EdgeRecs<KIND>::N, C = N / 2;
  typedef typename EdgeLeanCfg<KIND>::WarpBuf WarpBuf;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpBuf* wb = reinterpret_cast<WarpBuf*>(smem_raw) + 2 * (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  unsigned c_split = 0, c_coll = 0, c_eval = 0;
  double maxlen = 0.0;                  // getMaximumEdgeLength starts at 0 and ignores NaN (maSize.cc:673-691)
  int eig_any = 0;
  const int2 kNone = make_int2(0, -1);
  unsigned long long* counter = &st->edge_chunk;
  auto fetch_header = [&](int s, int& off, int& K, int& va) {
    off = __ldg(slice_off + s);
    K = (__ldg(slice_off + s + 1) - off) >> 5;
    va = __ldg(anchor + (s << 5) + lane);
  };
  // the anchors' records and all slot words of one slice -> shared memory buffer bi (asynchronous)
  auto prefetch_slice = [&](int bi, int off, int K, int va) {
    if (va >= 0) {
      const double2* pa = chunk_ptr<C>(vedge, 0, va);
#pragma unroll
      for (int i = 0; i < C; ++i) cp_async16(&wb[bi].a[i][lane], pa + i * kVB);
    }
    const int2* sp = slots + off + lane;
    for (int k = 0; k < K; ++k) cp_async8(&wb[bi].sl[k][lane], sp + 32 * k);
    cp_async_commit();
  };
  TicketStream<kEZGroup> ts;
  int s_cur = ts.first(counter, nslices);
  if (s_cur < nslices) {
    int off_c, K, va_c;
    fetch_header(s_cur, off_c, K, va_c);
    prefetch_slice(0, off_c, K, va_c);
    int s_n1 = ts.next(counter, nslices), s_n2 = nslices;
    int hb_off = 0, hb_K = 0, hb_va = -1;          // header of the slice after the current one: requested, consumed at the safe point
    if (s_n1 < nslices) fetch_header(s_n1, hb_off, hb_K, hb_va);
    cp_async_wait_all();
    int K_n1 = 0, va_n1 = -1;
    int cb = 0, k = 0;
    int2 cur = wb[0].sl[0][lane];
    // Every row is evaluated, also the empty slots of short rows (their gather reads vertex 0, their result is dropped):
    // no value is then live across a divergent branch, and the gather's registers are dead between phase one and its issue.
    double b[N];
    load_half_rec<KIND>(vedge, cur.y >= 0 ? (cur.x & kVidMask) : 0, b);
    unsigned nearmask = 0;
    for (;;) {
      // phase one: the gathered record and the anchor's are consumed
      EdgePre<KIND> pre;
      {
        double a[N];
#pragma unroll
        for (int i = 0; i < C; ++i) { const double2 t = wb[cb].a[i][lane]; a[2 * i] = t.x; a[2 * i + 1] = t.y; }
        edge_pre<KIND>(a, b, pre);
      }
      // safe point, once per slice: no gather is in flight.  The next slice's header has arrived (requested a slice ago): its
      // data starts travelling into the other buffer; the header after that and the next ticket are requested.
      if (k == 0) {
        K_n1 = hb_K;
        va_n1 = hb_va;
        if (s_n1 < nslices) prefetch_slice(cb ^ 1, hb_off, hb_K, hb_va);
        s_n2 = ts.next(counter, nslices);
        if (s_n2 < nslices) fetch_header(s_n2, hb_off, hb_K, hb_va);
      }
      edge_mid<KIND>(pre);                       // first Gauss point
      // the next row of the stream: its slot word from shared memory, its gather into the registers phase one freed
      const bool last = k + 1 >= K;
      int2 nx = kNone;
      if (!last) nx = wb[cb].sl[k + 1][lane];
      else if (s_n1 < nslices) { cp_async_wait_all(); nx = wb[cb ^ 1].sl[0][lane]; }
      load_half_rec<KIND>(vedge, nx.y >= 0 ? (nx.x & kVidMask) : 0, b);
      // second Gauss point, square roots, and the outputs of this row
      const double len = edge_post<KIND>(pre);
      if (cur.y >= 0) {
        const int e = cur.y;
        const bool owned = cur.x >= 0;               // sign bit of the other vertex id = "not owned"
        st_stream(lengths + e, len);
        if (owned && len > maxlen) maxlen = len;
        const bool nr = fabs(len - P.max_len) <= P.tol_max || fabs(len - P.min_len) <= P.tol_min;
        nearmask |= (nr ? 1u : 0u) << k;
        if (!nr) {                                  // near ones are decided in strict arithmetic after the slice (near_edges)
          ++c_eval;
          const bool sp = len > P.max_len, co = len < P.min_len;
          c_split += (sp && owned) ? 1u : 0u;
          c_coll += (co && owned) ? 1u : 0u;
          st_stream(flags + e, (int32_t)((sp ? MAG_SPLIT : MAG_NEED_NOT_SPLIT) | (co ? MAG_COLLAPSE : MAG_NEED_NOT_COLLAPSE)));
        }
      }
      if (!last) ++k;
      else {
        // near-threshold entities of the finished slice, one slot row at a time (warp-collective; its buffer is still intact)
        for (unsigned any = __reduce_or_sync(0xffffffffu, nearmask); any; any &= any - 1) {
          const int kk = __ffs(any) - 1;
          const bool nr = (nearmask >> kk) & 1u;
          const int2 q = nr ? wb[cb].sl[kk][lane] : kNone;
          const unsigned r = near_edges<KIND, true>(nr, q.y, va_c, q.x, 0, vedge, flags, lengths, P.ops, P.max_len, P.min_len, st, near_list);
          c_eval += r & 1u; c_split += (r >> 1) & 1u; c_coll += (r >> 2) & 1u; eig_any |= (int)(r >> 3);
        }
        if (s_n1 >= nslices) break;
        cb ^= 1; K = K_n1; k = 0; va_c = va_n1; s_cur = s_n1; s_n1 = s_n2; nearmask = 0;
      }
      cur = nx;
    }
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
