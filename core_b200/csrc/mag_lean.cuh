// mag_lean.cuh -- the whole-part sweep kernels of the full marking sweep: stream kernels over the anchor rows.
//
// Included by mag_kernels.cu after mag_rows.cuh.  Same anchor-row layout, same results bit for bit as the tile kernels
// k_edges / k_tets (tests/test_gpu_parity.py::test_lean_kernels_equal_general), used when the sweep is
//     MAG_FP_FAST, incoming flag words all zero (the state after mag_set_flags(NULL, NULL) / ma::Adapt's constructor on a
//     mesh without layer elements), lengths + both edge marks (edges; identity / iso / aniso fields) / qualities + BAD_QUALITY
//     with the max-Jacobian metric, the reference's default (tets)
// -- a full marking sweep.  Everything else runs the tile kernels.
//
// What the design answers (ncu source pages of the kernels that came before, profiles/r2_kernel_history.md):
//   1. The tile kernels spend 292 warp-instructions per edge and 301 per tet, 141 / 100 of them fp64: flag logic for words
//      that are known to be zero, 64-bit address arithmetic, register copies of software pipelines.  Here skip tests, incoming
//      words and error tests are compiled out, and the anchor's data is not gathered again for every entity of its row.
//   2. A first pair of lean kernels prefetched the next row's gather into registers (ping-pong) and lost a third (edges) to a
//      half (tets) of their time to "long scoreboard" anyway -- not at the gathers, but at the TEST OF A SLOT WORD loaded two
//      rows earlier and at the start of every slice (header -> slot words -> anchor -> first record: dependent round trips
//      for seven rows of work).  A warp has six scoreboard entries; a load that shares one with a younger load is complete
//      only when the younger one is.  And every register ptxas spills right after a load forces a wait for that load.
//   So the stream kernels keep ONE kind of long-latency load in flight in the steady state, the next row's gather, issued
//   where its registers are free and consumed where nothing younger has been issued; everything else a slice needs -- its
//   header, the anchors' data, every slot word -- travels into per-warp shared memory with cp.async (own completion counter,
//   no register, no scoreboard entry) one to two slices ahead, and each lane reads back only what it copied itself.
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
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_slot(int2* smem, const int2* gmem) { cp_async8(smem, gmem); }
__device__ __forceinline__ void cp_async_slot(int4* smem, const int4* gmem) { cp_async16(smem, gmem); }

// Successive slice ids for one warp: groups of G consecutive slices per ticket, the next ticket always requested a group
// ahead.  The atomic is written in PTX: left to itself, ptxas turns a one-lane atomicAdd into its warp-aggregated form, whose
// broadcast shuffle waits for the atomic right where it is issued.
template <int G>
struct TicketStream {
  int s, s_end;
  unsigned raw;
  static __device__ __forceinline__ unsigned issue(unsigned long long* counter)
  {
    unsigned long long t = 0;
    if ((threadIdx.x & 31) == 0) asm volatile("atom.global.add.u64 %0, [%1], 1;" : "=l"(t) : "l"(counter) : "memory");
    return (unsigned)t;
  }
  __device__ __forceinline__ int take(int n)
  {
    const unsigned g = __shfl_sync(0xffffffffu, raw, 0);
    const unsigned long long f = (unsigned long long)g * G;
    s = f < (unsigned long long)n ? (int)f : n;
    s_end = s + G < n ? s + G : n;
    return s;
  }
  __device__ __forceinline__ int first(unsigned long long* counter, int n)
  {
    raw = issue(counter);
    const int r = take(n);
    raw = issue(counter);
    return r;
  }
  __device__ __forceinline__ int next(unsigned long long* counter, int n)
  {
    if (s + 1 < s_end) return ++s;
    const int r = take(n);                            // requested a whole group ago
    raw = issue(counter);
    return r;
  }
};

// Per-warp staging of slices in shared memory.  Data (anchor records, slot words) is double-buffered: slice i of the warp's
// stream lives in buffer i & 1.  Headers (slice_off[s], slice_off[s+1], the lanes' anchor vertex ids) are triple-buffered, slot
// i % 3: the header of slice i+2 is announced while slices i and i+1 still need theirs.  Everything arrives through cp.async and
// is read back only by the lane that copied it -- except the two slice_off words (copied by lanes 0 and 1, read by all), which
// is why a __syncwarp follows every cp.async.wait_group before a header is read.
template <int AW, class SlotT>
struct WarpStage {
  double2 a[2][AW][32];
  SlotT sl[2][kRowMax][32];
  int va[3][32];
  int off[3][2];
  // header of slice s -> header slot h (asynchronous)
  __device__ __forceinline__ void announce(int h, int s, const int32_t* __restrict__ slice_off, const int32_t* __restrict__ anchor, int lane)
  {
    if (lane < 2) cp_async4(&off[h][lane], slice_off + s + lane);
    cp_async4(&va[h][lane], anchor + (s << 5) + lane);
    cp_async_commit();
  }
  __device__ __forceinline__ int rows(int h) const { return (off[h][1] - off[h][0]) >> 5; }
  // the data of the slice whose header sits in slot h -> data buffer b (asynchronous); copy_anchor copies the anchor's chunks
  template <class AFn>
  __device__ __forceinline__ void stage(int b, int h, const SlotT* __restrict__ slots, int lane, AFn copy_anchor)
  {
    const int v = va[h][lane];
    if (v >= 0) copy_anchor(&a[b][0][lane], v);
    const int K = rows(h);
    const SlotT* sp = slots + off[h][0] + lane;
    for (int k = 0; k < K; ++k) cp_async_slot(&sl[b][k][lane], sp + 32 * k);
    cp_async_commit();
  }
};
__device__ __forceinline__ int next3(int h) { return h == 2 ? 0 : h + 1; }

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
  typedef WarpStage<C, int2> Stage;
  static constexpr size_t kSmem = sizeof(Stage) * (T / 32);
};

// lengths + SPLIT + COLLAPSE of every edge, flag words written from zero (MAG_FP_FAST; identity / iso / aniso fields).
// A row is evaluated in phases: the gathered record is consumed first (aniso_pre: 21 numbers), the first Gauss point is
// evaluated, THEN the next row's gather is issued -- also across a slice boundary -- into registers that are free by then, and
// the second Gauss point and the square roots run while it travels.  Every row is evaluated, also the empty slots of short
// rows (their gather reads vertex 0, their result is dropped): no value is live across a divergent branch.
template <int KIND>
__global__ void __launch_bounds__(EdgeLeanCfg<KIND>::T, EdgeLeanCfg<KIND>::B)
k_edge_rows_z(int32_t nslices, const int32_t* __restrict__ anchor, const int32_t* __restrict__ slice_off, const int2* __restrict__ slots,
              const double* __restrict__ vedge, int32_t* __restrict__ flags, double* __restrict__ lengths, EdgeParams P,
              MagDevStats* st, int32_t* __restrict__ near_list)
{
  constexpr int N = EdgeRecs<KIND>::N, C = N / 2;
  typedef typename EdgeLeanCfg<KIND>::Stage Stage;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Stage& ws = reinterpret_cast<Stage*>(smem_raw)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  unsigned c_split = 0, c_coll = 0, c_eval = 0;
  double maxlen = 0.0;                  // getMaximumEdgeLength starts at 0 and ignores NaN (maSize.cc:673-691)
  int eig_any = 0;
  const int2 kNone = make_int2(0, -1);
  unsigned long long* counter = &st->edge_chunk;
  auto copy_anchor = [&](double2* dst, int v) {
    const double2* pa = chunk_ptr<C>(vedge, 0, v);
#pragma unroll
    for (int i = 0; i < C; ++i) cp_async16(dst + i * 32, pa + i * kVB);
  };
  TicketStream<kEZGroup> ts;
  const int s_first = ts.first(counter, nslices);
  if (s_first < nslices) {
    // prologue: the first slice announced and staged, the second announced
    const int s_second = ts.next(counter, nslices);
    bool more = s_second < nslices, more2 = false;    // a slice follows the current one; a slice follows that one
    ws.announce(0, s_first, slice_off, anchor, lane);
    if (more) ws.announce(1, s_second, slice_off, anchor, lane);
    cp_async_wait_all();
    __syncwarp();
    ws.stage(0, 0, slots, lane, copy_anchor);
    cp_async_wait_all();
    int hc = 0, cb = 0, k = 0;                        // header slot and data buffer of the current slice, current row
    int K = ws.rows(0);
    int2 cur = ws.sl[0][0][lane];
    double b[N];
    load_half_rec<KIND>(vedge, cur.y >= 0 ? (cur.x & kVidMask) : 0, b);
    unsigned nearmask = 0;
    for (;;) {
      // phase one: the gathered record and the anchor's are consumed
      EdgePre<KIND> pre;
      {
        double a[N];
#pragma unroll
        for (int i = 0; i < C; ++i) { const double2 t = ws.a[cb][i][lane]; a[2 * i] = t.x; a[2 * i + 1] = t.y; }
        edge_pre<KIND>(a, b, pre);
      }
      // safe point, once per slice: no gather is in flight.  The next slice's header arrived a slice ago: its data starts
      // travelling into the other buffer; the next ticket is drawn and the header of the slice after that announced.
      if (k == 0 && more) {
        ws.stage(cb ^ 1, next3(hc), slots, lane, copy_anchor);
        const int s2 = ts.next(counter, nslices);
        more2 = s2 < nslices;
        if (more2) ws.announce(next3(next3(hc)), s2, slice_off, anchor, lane);
      }
      edge_mid<KIND>(pre);                            // first Gauss point
      // the next row of the stream: its slot word from shared memory, its gather into the registers phase one freed
      const bool last = k + 1 >= K;
      int2 nx = kNone;
      if (!last) nx = ws.sl[cb][k + 1][lane];
      else if (more) { cp_async_wait_all(); __syncwarp(); nx = ws.sl[cb ^ 1][0][lane]; }
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
          const int2 q = nr ? ws.sl[cb][kk][lane] : kNone;
          const unsigned r = near_edges<KIND, true>(nr, q.y, ws.va[hc][lane], q.x, 0, vedge, flags, lengths, P.ops, P.max_len, P.min_len, st, near_list);
          c_eval += r & 1u; c_split += (r >> 1) & 1u; c_coll += (r >> 2) & 1u; eig_any |= (int)(r >> 3);
        }
        if (!more) break;
        more = more2;
        more2 = false;
        cb ^= 1; hc = next3(hc); K = ws.rows(hc); k = 0; nearmask = 0;
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
// The quality of a tet needs its four vertices and then -- a DEPENDENT gather -- the transform of the vertex with the largest
// det Q_v (getMetricWithMaxJacobean, maQuality.cc:83-108).  The stream is therefore three rows deep:
//     E0  the row being evaluated: its {x, y} and the winner's transform were requested one row ago, its z two rows ago
//     E1  the next row: its {z, det Q_v} arrived during the last row; the winner is chosen now and its transform + {x, y} requested
//     E2  the row after: its {z, det Q_v} are requested now
// Per row:  y = (x_i - x_0) Q from what arrived (coordinates and transform are dead from here on)  ->  winner of E1  ->  the two
// groups of gathers are issued into the registers just freed  ->  quality from y (sum of squared edges, volume) and the outputs.
// Every gather has one whole row to arrive and is consumed where nothing younger is in flight.  Rows come from a read cursor
// that runs two rows ahead of the evaluation through the staged slices (at most one slice boundary ahead: a row it cannot
// reach yet is a bubble -- an empty row -- and is fetched again next time).
template <int KIND> struct TetLeanCfg {
  static constexpr int T = MAG_TZ_THREADS, B = MAG_TZ_BLOCKS;
  typedef WarpStage<2, int4> Stage;   // anchors' {x,y} {z,det Q_v}; slot = {o1 | not-owned << 31, o2, o3, tet index}
  static constexpr size_t kSmem = sizeof(Stage) * (T / 32);
};
// qualities + BAD_QUALITY of every tet, flag words written from zero (MAG_FP_FAST, max-Jacobian metric)
template <int KIND>
__global__ void __launch_bounds__(TetLeanCfg<KIND>::T, TetLeanCfg<KIND>::B)
k_tet_rows_z(int32_t nslices, const int32_t* __restrict__ anchor, const int32_t* __restrict__ slice_off, const int4* __restrict__ slots,
             int32_t elem_off, int64_t nv, const double* __restrict__ vpos, const double* __restrict__ vq, const double* __restrict__ vedge,
             int32_t* __restrict__ flags, double* __restrict__ qual, TetParams P, MagDevStats* st, int32_t* __restrict__ near_list)
{
  typedef typename TetLeanCfg<KIND>::Stage Stage;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Stage& ws = reinterpret_cast<Stage*>(smem_raw)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  unsigned c_bad = 0, c_eval = 0;
  unsigned long long minkey = ~0ull;
  int eig_any = 0;
  flags += elem_off;
  qual += elem_off;
  const int4 kNone = make_int4(0, 0, 0, -1);
  unsigned long long* counter = &st->elem_chunk;
  auto copy_anchor = [&](double2* dst, int v) {
    cp_async16(dst, chunk_ptr<2>(vpos, 0, v));
    cp_async16(dst + 32, chunk_ptr<2>(vpos, 1, v));
  };
  // an entry of the stream: the lane's slot word and, warp-uniform, where the row lives
  //   meta bit 0 data buffer, bits 1-3 row, bit 4 first row of its slice, bit 5 last row, bit 6 a row (clear: bubble / end),
  //   bits 8-9 header slot of its slice
  struct Ent { int4 sl; int meta; };
  TicketStream<kTZGroup> ts;
  const int s_first = ts.first(counter, nslices);
  if (s_first < nslices) {
    // prologue: two slices announced and staged, the third announced
    int s_next = ts.next(counter, nslices);         // the slice after the one the evaluation is in (>= nslices: none)
    int s_next2 = nslices;                          // the slice after that
    ws.announce(0, s_first, slice_off, anchor, lane);
    if (s_next < nslices) ws.announce(1, s_next, slice_off, anchor, lane);
    cp_async_wait_all();
    __syncwarp();
    ws.stage(0, 0, slots, lane, copy_anchor);
    if (s_next < nslices) {
      ws.stage(1, 1, slots, lane, copy_anchor);
      s_next2 = ts.next(counter, nslices);
      if (s_next2 < nslices) ws.announce(2, s_next2, slice_off, anchor, lane);
    }
    cp_async_wait_all();
    __syncwarp();
    int hc = 0;                                     // header slot of the slice the evaluation is in
    // read cursor: data buffer, header slot, next row, rows of its slice, whether it is one slice ahead of the evaluation
    int rb = 0, rh = 0, rk = 0, rK = ws.rows(0), ahead = 0;
    bool r_done = false;
    auto fetch = [&]() {
      Ent e;
      e.sl = kNone;
      e.meta = 0;
      if (r_done) return e;
      if (rk >= rK) {                               // this slice is exhausted
        if (ahead) return e;                        // the slice after the next is not staged yet: bubble
        if (s_next >= nslices) { r_done = true; return e; }
        cp_async_wait_all();
        __syncwarp();
        rb ^= 1; rh = next3(rh); rK = ws.rows(rh); rk = 0; ahead = 1;
      }
      e.sl = ws.sl[rb][rk][lane];
      e.meta = 64 | rb | (rk << 1) | (rk == 0 ? 16 : 0) | (rk + 1 == rK ? 32 : 0) | (rh << 8);
      ++rk;
      return e;
    };
    auto load_zd = [&](const Ent& e, double2* zd) {
      const bool ok = e.sl.w >= 0;
      zd[0] = __ldg(chunk_ptr<2>(vpos, 1, ok ? (e.sl.x & kVidMask) : 0));
      zd[1] = __ldg(chunk_ptr<2>(vpos, 1, ok ? e.sl.y : 0));
      zd[2] = __ldg(chunk_ptr<2>(vpos, 1, ok ? e.sl.z : 0));
    };
    // winner of the entry's tet from the determinants, then its transform and the {x, y} of the three other vertices
    auto load_xyq = [&](const Ent& e, const double2* zd, double2* xy, double2* q) {
      const bool ok = e.sl.w >= 0;
      const int32_t v1 = e.sl.x & kVidMask;
      int32_t vb = 0;
      if (ok) vb = best_vertex(make_int4(ws.va[(e.meta >> 8) & 3][lane], v1, e.sl.y, e.sl.z), ws.a[e.meta & 1][1][lane].y, zd[0].y, zd[1].y, zd[2].y);
      const double2* pq = chunk_ptr<5>(vq, 0, vb);
      xy[0] = __ldg(chunk_ptr<2>(vpos, 0, ok ? v1 : 0));
      xy[1] = __ldg(chunk_ptr<2>(vpos, 0, ok ? e.sl.y : 0));
      xy[2] = __ldg(chunk_ptr<2>(vpos, 0, ok ? e.sl.z : 0));
#pragma unroll
      for (int i = 0; i < 5; ++i) q[i] = __ldg(pq + i * kVB);
    };
    // fill the stream
    Ent E0 = fetch(), E1 = fetch(), E2 = fetch();
    double2 zd[3], xy[3], q[5];
    double z0[3];
    {
      double2 zd0[3];
      load_zd(E0, zd0);
      load_zd(E1, zd);
      load_xyq(E0, zd0, xy, q);
      z0[0] = zd0[0].x; z0[1] = zd0[1].x; z0[2] = zd0[2].x;
    }
    bool first_slice = true;
    unsigned nearmask = 0;
    while (((E0.meta | E1.meta | E2.meta) & 64) || !r_done) {
      // the row's edge vectors in metric space, from what arrived during the last row
      double y[3][3];
      {
        const double2 a_xy = ws.a[E0.meta & 1][0][lane], a_zd = ws.a[E0.meta & 1][1][lane];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double ex = xy[i].x - a_xy.x, ey = xy[i].y - a_xy.y, ez = z0[i] - a_zd.x;
          y[i][0] = fma(ez, q[3].x, fma(ey, q[1].y, __dmul_rn(ex, q[0].x)));   // Q row-major in q: {00,01} {02,10} {11,12} {20,21} {22,det}
          y[i][1] = fma(ez, q[3].y, fma(ey, q[2].x, __dmul_rn(ex, q[0].y)));
          y[i][2] = fma(ez, q[4].x, fma(ey, q[2].y, __dmul_rn(ex, q[1].x)));
        }
      }
      // safe point, when the evaluation enters a slice (not the first one: the prologue did its work): no gather is in flight.
      // The slice before is finished, so the slice after this one is staged into its buffer (its header was announced a slice
      // ago), the next ticket is drawn and the header after that announced.
      if ((E0.meta & 16) && !first_slice) {
        hc = next3(hc);
        s_next = s_next2;
        s_next2 = nslices;
        if (s_next < nslices) {
          cp_async_wait_all();
          __syncwarp();
          ws.stage((E0.meta & 1) ^ 1, next3(hc), slots, lane, copy_anchor);
          s_next2 = ts.next(counter, nslices);
          if (s_next2 < nslices) ws.announce(next3(next3(hc)), s_next2, slice_off, anchor, lane);
        }
        ahead = 0;
      }
      if (E0.meta & 16) first_slice = false;
      // next row: winner, transform, {x, y}; the row after: {z, det}
      load_xyq(E1, zd, xy, q);
      z0[0] = zd[0].x; z0[1] = zd[1].x; z0[2] = zd[2].x;
      load_zd(E2, zd);
      // quality and outputs of this row
      const double qv = magfa::tet_quality_y(y);
      const int t = E0.sl.w;
      const int k = (E0.meta >> 1) & 7;
      if (t >= 0) {
        st_stream(qual + t, qv);
        const unsigned long long kq = dkey(qv);
        minkey = kq < minkey ? kq : minkey;
        const bool nr = fabs(qv - P.good_q) <= P.tol_q;
        nearmask |= (nr ? 1u : 0u) << k;
        if (!nr) {
          ++c_eval;
          const bool bad = qv < P.good_q;
          c_bad += (bad && E0.sl.x >= 0) ? 1u : 0u;
          st_stream(flags + t, (int32_t)(bad ? MAG_BAD_QUALITY : MAG_OK_QUALITY));
        }
      }
      if (E0.meta & 32) {                           // last row of its slice: near-threshold tets, one slot row at a time
        const int bi = E0.meta & 1;
        for (unsigned any = __reduce_or_sync(0xffffffffu, nearmask); any; any &= any - 1) {
          const int kk = __ffs(any) - 1;
          const bool nr = (nearmask >> kk) & 1u;
          const int4 qs = nr ? ws.sl[bi][kk][lane] : kNone;
          const unsigned r = near_tets<KIND, true>(nr, qs.w, elem_off, make_int4(ws.va[hc][lane], qs.x, qs.y, qs.z), 0, nv, vpos, vq, vedge,
                                                   flags - elem_off, qual - elem_off, P.ops, P.good_q, P.use_max, st, near_list);
          c_eval += r & 1u; c_bad += (r >> 1) & 1u; eig_any |= (int)(r >> 2);
        }
        nearmask = 0;
      }
      E0 = E1;
      E1 = E2;
      E2 = fetch();
    }
  }
  if (eig_any) atomicAdd(&st->n_eigen_fail, 1ull);
  warp_count_to(c_bad, &st->n_bad);
  warp_count_to(c_eval, &st->n_elems_eval);
  const unsigned long long m = warp_min_u64(minkey);
  if (lane == 0 && m != ~0ull) atomicMin(&st->min_q_key, m);
}
