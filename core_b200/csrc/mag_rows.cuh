// mag_rows.cuh -- the whole-part sweep kernels of round 2: anchor rows.
//
// Included by mag_kernels.cu after the arithmetic and the parameter structs.  Replaces, for whole-part sweeps, the tile
// kernels k_edges / k_tets (which stay for the sub-range sweeps of mag_sweep_host), for this reason (ncu, round 1): those
// kernels gather BOTH 96-byte vertex records of every edge with scattered 16-byte loads -- 12 requests x 7.3 L1 wavefronts
// per warp, the L1 data pipe at 62 %, 37 % of the warps resident -- so the L1 -> register path, not HBM and not the
// fp64 pipe, set their speed.
//
// Layout (built once at export, build_rows below).  Every edge / tet is filed under its FIRST vertex, the anchor.  An
// anchor's entities form a row (rows longer than kRowMax are cut); rows are ordered by vertex id, and inside windows of
// kRowWindow vertices by descending length (SELL-C-sigma of sparse matrix-vector products: C = 32, sigma = kRowWindow), cut
// into slices of 32 rows -- one per warp, one row per lane -- and each slice is stored slot-major:
//     slot (s, k, lane)  at  slice_off[s] + 32 k + lane     k < width of slice s = its longest row
// holding the OTHER vertex ids, the ownership bit and the entity's index in the caller's order (-1 = empty).
// A lane keeps its anchor's record in registers for the whole row and gathers only the other end of each edge (6
// requests instead of 12; the other three vertices of a tet), and since neighbouring lanes hold neighbouring anchors
// whose k-th entities belong to the same family on any structured numbering (box meshes: the +x edge of every vertex,
// then +y, ...), the lanes of one request read consecutive records: 4-5 wavefronts instead of 7-8.  Lengths, qualities
// and flag words are written to the caller's entity order through the slot's entity index, so nothing else in the library
// (getters, reconciliation lists, the sweeps either side of the path) sees the layout.
//
// Work is handed out one slice per warp through an atomic ticket, so all warps of the device advance through the vertex
// array together (a record is fetched from HBM about once per sweep) and slow slices (strict re-evaluations) even out.
// Entities whose value lands within 1e-12 of a threshold are re-evaluated in the reference's operation order under a
// warp-uniform branch: on the structured benchmark whole families sit ON a threshold (the z edges of config 3 measure
// exactly 0.5), i.e. whole warps take the branch together; on an unstructured mesh it is almost never taken.
#pragma once

constexpr int kRowMax = 32;          // longest row; an anchor with more entities gets several rows
constexpr int kRowWindowLog2 = 11;   // rows are sorted by length inside windows of 2048 vertices

#ifndef MAG_EROW_THREADS
#define MAG_EROW_THREADS 256
#endif
#ifndef MAG_EROW_BLOCKS
#define MAG_EROW_BLOCKS 3
#endif
#ifndef MAG_EROW_BLOCKS_LOGM
#define MAG_EROW_BLOCKS_LOGM 2
#endif
#ifndef MAG_TROW_THREADS
#define MAG_TROW_THREADS 256
#endif
#ifndef MAG_TROW_BLOCKS
#define MAG_TROW_BLOCKS 2
#endif
template <int KIND, bool FAST> struct EdgeRowCfg {
  static constexpr int T = FAST ? MAG_EROW_THREADS : kStrictThreads;
  static constexpr int B = FAST ? (KIND == MAG_KIND_LOGM ? MAG_EROW_BLOCKS_LOGM : MAG_EROW_BLOCKS) : kStrictBlocks;
};
template <bool FAST> struct TetRowCfg {
  static constexpr int T = FAST ? MAG_TROW_THREADS : kStrictThreads, B = FAST ? MAG_TROW_BLOCKS : kStrictBlocks;
};

#ifndef MAG_ROW_STATIC
#define MAG_ROW_STATIC 0   /* 1: slice s goes to warp s mod (warps of the grid); 0: atomic ticket per slice.  Measured (B200, n = 203,
                              r2b): static 1.94 / 1.77 ms (lattice / jittered) with the SMs idle 45 % of the launch -- warps that fall behind lose
                              the L2 window the others share and fall further behind; ticket 1.42 / 1.02 ms */
#endif
#ifndef MAG_EROW_PF
#define MAG_EROW_PF 0      /* 1: the other end's record of the NEXT slot row is requested before this one is evaluated (24 registers) */
#endif
// Slice hand-out.  Static: all warps of the grid stride through the slices together (so they sweep the vertex array
// together) and a warp knows its next slice at once -- its header is requested a slice ahead.  Ticket: lane 0 draws the
// warp's next slice from an atomic counter; the value is broadcast (and thereby waited for) only when it is needed.
struct SliceWalk {
  int s, raw, step;
  __device__ __forceinline__ void begin(unsigned long long* counter)
  {
    if (MAG_ROW_STATIC) {
      const int wpb = blockDim.x >> 5;
      s = blockIdx.x * wpb + (threadIdx.x >> 5);
      step = gridDim.x * wpb;
    } else {
      raw = issue(counter);
      s = __shfl_sync(0xffffffffu, raw, 0);
    }
  }
  // call at the top of a slice: returns the slice after this one when it is already known (static), else -1
  __device__ __forceinline__ int peek(unsigned long long* counter)
  {
    if (MAG_ROW_STATIC) return s + step;
    raw = issue(counter);
    return -1;
  }
  __device__ __forceinline__ void next()
  {
    if (MAG_ROW_STATIC) s += step; else s = __shfl_sync(0xffffffffu, raw, 0);
  }
  static __device__ __forceinline__ int issue(unsigned long long* counter)
  {
    unsigned long long t = 0;
    if ((threadIdx.x & 31) == 0) t = atomicAdd(counter, 1ull);
    return (int)t;
  }
};

template <int KIND>
__device__ __forceinline__ void load_half_rec(const double* __restrict__ vedge, int32_t v, double* r)
{
  constexpr int N = EdgeRecs<KIND>::N;
  const double2* p = chunk_ptr<N / 2>(vedge, 0, v);
#pragma unroll
  for (int i = 0; i < N / 2; ++i) { double2 t = __ldg(p + i * kVB); r[2 * i] = t.x; r[2 * i + 1] = t.y; }
}

// the lanes with nr set: listed, and in MAG_FP_FAST re-evaluated in strict arithmetic (the flag word and the length
// become the reference's).  Warp-collective.  Returns bit 0: evaluated, bit 1: counted SPLIT, bit 2: counted COLLAPSE,
// bit 3: the eigen-solver failed
template <int KIND, bool FAST>
__device__ __noinline__ unsigned near_edges(bool nr, int32_t e, int32_t va, int32_t vb_word, int32_t f, const double* __restrict__ vedge,
                                            int32_t* __restrict__ flags, double* __restrict__ lengths, uint32_t ops,
                                            double max_len, double min_len, MagDevStats* st, int32_t* __restrict__ near_list)
{
  const unsigned m = __ballot_sync(0xffffffffu, nr);
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(&st->n_near_edge, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  unsigned out = 0;
  if (nr) {
    near_list[base + __popc(m & ((1u << lane) - 1u))] = e;
    if (FAST) {
      const SweepParams P{ops, max_len, min_len, 0.0, 0};
      const bool need_split = (ops & MAG_OP_MARK_SPLIT) && !(f & kSkipSplit);
      const bool need_coll = (ops & MAG_OP_MARK_COLLAPSE) && !(f & kSkipColl);
      EdgeRecs<KIND> R;
      load_edge_recs<KIND>(vedge, make_int2(va, vb_word & kVidMask), R);
      int eig = 0;
      const double len = edge_length_strict<KIND>(R, &eig);
      unsigned cs = 0, cc = 0;
      mark_edge(len, f, need_split, need_coll, vb_word >= 0, P, cs, cc);
      out = 1u | (cs << 1) | (cc << 2) | (eig ? 8u : 0u);
      flags[e] = f;
      if (ops & MAG_OP_LENGTHS) lengths[e] = len;
    }
  }
  __syncwarp();
  return out;
}

template <int KIND, bool FAST>
__global__ void __launch_bounds__(EdgeRowCfg<KIND, FAST>::T, EdgeRowCfg<KIND, FAST>::B)
k_edge_rows(int32_t nslices, const int32_t* __restrict__ anchor, const int32_t* __restrict__ slice_off, const int2* __restrict__ slots,
            const double* __restrict__ vedge, int32_t* __restrict__ flags, double* __restrict__ lengths, EdgeParams P,
            MagDevStats* st, int32_t* __restrict__ near_list, PfArgs pf)
{
  constexpr int N = EdgeRecs<KIND>::N;
  const int lane = threadIdx.x & 31;
  unsigned c_split = 0, c_coll = 0, c_eval = 0, c_err = 0;
  double maxlen = 0.0;                  // getMaximumEdgeLength starts at 0 and ignores NaN (maSize.cc:673-691)
  int eig_any = 0;
  SliceWalk w;
  w.begin(&st->edge_chunk);
  int off = 0, off1 = 0, va = -1;
  if (w.s < nslices) { off = __ldg(slice_off + w.s); off1 = __ldg(slice_off + w.s + 1); va = __ldg(anchor + (w.s << 5) + lane); }
  while (w.s < nslices) {
    const int s_nx = w.peek(&st->edge_chunk);
    const int K = (off1 - off) >> 5;
    int pf_b0 = 0, pf_b1 = 0;                      // lane 0: the vertex blocks slice s + dist adds (requested at the end of this slice)
    if (pf.blk && lane == 0 && w.s + pf.dist < pf.n) { pf_b0 = __ldg(pf.blk + w.s + pf.dist - 1); pf_b1 = __ldg(pf.blk + w.s + pf.dist); }
    const int2* sp = slots + off + lane;
    int2 sl = ld_stream(sp);                      // every slice is at least one slot wide
    int2 sl1 = make_int2(0, -1);
    if (K > 1) sl1 = ld_stream(sp + 32);
    EdgeRecs<KIND> R;
#pragma unroll
    for (int i = 0; i < N; ++i) R.a[i] = 0.0;
    if (va >= 0) load_half_rec<KIND>(vedge, va, R.a);
    // header of the next slice (static hand-out: known now; ticket hand-out: read when the ticket has arrived, below)
    int off_nx = 0, off1_nx = 0, va_nx = -1;
    if (MAG_ROW_STATIC && s_nx < nslices) { off_nx = __ldg(slice_off + s_nx); off1_nx = __ldg(slice_off + s_nx + 1); va_nx = __ldg(anchor + (s_nx << 5) + lane); }
    int32_t f = (!P.zero_in && sl.y >= 0) ? ld_stream_rw(flags + sl.y) : 0;
#if MAG_EROW_PF
    double bn[N];
#pragma unroll
    for (int i = 0; i < N; ++i) bn[i] = 0.0;
    if (sl.y >= 0) load_half_rec<KIND>(vedge, sl.x & kVidMask, bn);
#endif
    unsigned nearmask = 0;                        // bit k: this lane's k-th entity landed within 1e-12 of a threshold
    for (int k = 0; k < K; ++k) {
      int2 sl2 = make_int2(0, -1);
      if (k + 2 < K) sl2 = ld_stream(sp + (k + 2) * 32);
      const int32_t f1 = (!P.zero_in && sl1.y >= 0) ? ld_stream_rw(flags + sl1.y) : 0;
      const int e = sl.y;
#if MAG_EROW_PF
#pragma unroll
      for (int i = 0; i < N; ++i) R.b[i] = bn[i];
      if (sl1.y >= 0) load_half_rec<KIND>(vedge, sl1.x & kVidMask, bn);   // the next slot row's records travel during this evaluation
#endif
      if (e >= 0) {
        const int32_t fe = f | P.off_bits;
        const bool need_split = !(fe & kSkipSplit), need_coll = !(fe & kSkipColl);
        if (f & P.err_mask) ++c_err;
        if (P.want_len || need_split || need_coll) {
          const bool owned = sl.x >= 0;           // sign bit of the other vertex id = "not owned"
#if !MAG_EROW_PF
          load_half_rec<KIND>(vedge, sl.x & kVidMask, R.b);
#endif
          const double len = FAST ? edge_length_fast<KIND>(R, &eig_any) : edge_length_strict<KIND>(R, &eig_any);
          if (P.want_len) {
            st_stream(lengths + e, len);
            if (owned && len > maxlen) maxlen = len;
          }
          if (need_split || need_coll) {
            const bool nr = (need_split && fabs(len - P.max_len) <= P.tol_max) || (need_coll && fabs(len - P.min_len) <= P.tol_min);
            nearmask |= (nr ? 1u : 0u) << k;
            if (!(FAST && nr)) {
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
      sl = sl1;
      sl1 = sl2;
      f = f1;
    }
    // near-threshold entities of this slice, one slot row at a time (the anchor record is dead by now: the call costs the
    // main loop no registers).  In MAG_FP_FAST their flag words have not been written yet, so the incoming word is re-read.
    for (unsigned any = __reduce_or_sync(0xffffffffu, nearmask); any; any &= any - 1) {
      const int k = __ffs(any) - 1;
      const bool nr = (nearmask >> k) & 1u;
      const int2 w = nr ? __ldg(sp + k * 32) : make_int2(0, -1);
      const int32_t fw = (nr && !P.zero_in && FAST) ? flags[w.y] : 0;
      const unsigned r = near_edges<KIND, FAST>(nr, w.y, va, w.x, fw, vedge, flags, lengths, P.ops, P.max_len, P.min_len, st, near_list);
      c_eval += r & 1u; c_split += (r >> 1) & 1u; c_coll += (r >> 2) & 1u; eig_any |= (int)(r >> 3);
    }
    if (pf_b1 > pf_b0) l2_prefetch_blocks<N / 2>(vedge, pf_b0, pf_b1);
    w.next();
    if (MAG_ROW_STATIC) { off = off_nx; off1 = off1_nx; va = va_nx; }
    else if (w.s < nslices) { off = __ldg(slice_off + w.s); off1 = __ldg(slice_off + w.s + 1); va = __ldg(anchor + (w.s << 5) + lane); }
  }
  if (eig_any) atomicAdd(&st->n_eigen_fail, 1ull);
  warp_count_to(c_split, &st->n_split);
  warp_count_to(c_coll, &st->n_collapse);
  warp_count_to(c_eval, &st->n_edges_eval);
  warp_count_to(c_err, &st->n_flag_err);
  if (P.want_len) {
    const unsigned long long m = warp_max_u64((unsigned long long)__double_as_longlong(maxlen));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&st->max_len_bits, m);
  }
}

// ------------------------------------------------------------------ tets
// returns bit 0: evaluated, bit 1: counted BAD_QUALITY, bit 2: the eigen-solver failed
template <int KIND, bool FAST>
__device__ __noinline__ unsigned near_tets(bool nr, int32_t t, int32_t elem_off, int4 tv_word, int32_t f, int64_t nv,
                                           const double* __restrict__ vpos, const double* __restrict__ vq, const double* __restrict__ vedge,
                                           int32_t* __restrict__ flags, double* __restrict__ qual, uint32_t ops, double good_q, int use_max,
                                           MagDevStats* st, int32_t* __restrict__ near_list)
{
  const unsigned m = __ballot_sync(0xffffffffu, nr);
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(&st->n_near_elem, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  unsigned out = 0;
  if (nr) {
    const int32_t el = elem_off + t;
    near_list[base + __popc(m & ((1u << lane) - 1u))] = el;
    if (FAST) {
      const SweepParams P{ops, 0.0, 0.0, good_q, use_max};
      const bool owned = tv_word.y >= 0;
      int4 tv = tv_word;
      tv.y &= kVidMask;
      int eig = 0;
      const double qv = tet_quality_eval<KIND, false, false>(tv, nv, vpos, vq, vedge, use_max, &eig, nullptr);
      unsigned cb = 0;
      mark_tet(qv, f, owned, P, cb);
      out = 1u | (cb << 1) | (eig ? 4u : 0u);
      flags[el] = f;
      if (ops & MAG_OP_QUALITIES) qual[el] = qv;
    }
  }
  __syncwarp();
  return out;
}

// slot = {o1 | not-owned << 31, o2, o3, tet index}; the anchor is the tet's first vertex, so (anchor, o1, o2, o3) is the
// caller's vertex order.  Per lane: the anchor's {x,y} {z,det Q_v} stay in registers for the whole row; slots are read two
// rows ahead, the three {z, det Q_v} chunks one row ahead (the choice of the max-Jacobian vertex, maQuality.cc:83-108,
// then does not sit between two dependent gathers).
template <int KIND, bool FAST, bool USE_MAX>
__global__ void __launch_bounds__(TetRowCfg<FAST>::T, TetRowCfg<FAST>::B)
k_tet_rows(int32_t nslices, const int32_t* __restrict__ anchor, const int32_t* __restrict__ slice_off, const int4* __restrict__ slots,
           int32_t elem_off, int64_t nv, const double* __restrict__ vpos, const double* __restrict__ vq, const double* __restrict__ vedge,
           int32_t* __restrict__ flags, double* __restrict__ qual, TetParams P, MagDevStats* st, int32_t* __restrict__ near_list, PfArgs pf)
{
  const int lane = threadIdx.x & 31;
  unsigned c_bad = 0, c_eval = 0, c_err = 0;
  unsigned long long minkey = ~0ull;
  int eig_any = 0;
  flags += elem_off;
  qual += elem_off;
  auto wanted = [&](int32_t t, int32_t fw) { return t >= 0 && (P.want_q || (P.do_bad && !(fw & MAG_OK_QUALITY))); };
  auto load_zd = [&](const int4& w, double2* zd) {
    zd[0] = __ldg(chunk_ptr<2>(vpos, 1, w.x & kVidMask)); zd[1] = __ldg(chunk_ptr<2>(vpos, 1, w.y)); zd[2] = __ldg(chunk_ptr<2>(vpos, 1, w.z));
  };
  SliceWalk w;
  w.begin(&st->elem_chunk);
  int off = 0, off1 = 0, va = -1;
  if (w.s < nslices) { off = __ldg(slice_off + w.s); off1 = __ldg(slice_off + w.s + 1); va = __ldg(anchor + (w.s << 5) + lane); }
  while (w.s < nslices) {
    const int s_nx = w.peek(&st->elem_chunk);
    const int K = (off1 - off) >> 5;
    int pf_b0 = 0, pf_b1 = 0;
    if (pf.blk && lane == 0 && w.s + pf.dist < pf.n) { pf_b0 = __ldg(pf.blk + w.s + pf.dist - 1); pf_b1 = __ldg(pf.blk + w.s + pf.dist); }
    int off_nx = 0, off1_nx = 0, va_nx = -1;
    if (MAG_ROW_STATIC && s_nx < nslices) { off_nx = __ldg(slice_off + s_nx); off1_nx = __ldg(slice_off + s_nx + 1); va_nx = __ldg(anchor + (s_nx << 5) + lane); }
    const int4* sp = slots + off + lane;
    int4 sl = ld_stream(sp);
    int4 sl1 = make_int4(0, 0, 0, -1);
    if (K > 1) sl1 = ld_stream(sp + 32);
    double2 a_xy = make_double2(0.0, 0.0), a_zd = a_xy;
    if (va >= 0) { a_xy = __ldg(chunk_ptr<2>(vpos, 0, va)); a_zd = __ldg(chunk_ptr<2>(vpos, 1, va)); }
    int32_t f = (!P.zero_in && sl.w >= 0) ? ld_stream_rw(flags + sl.w) : 0;
    double2 zd[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) zd[i] = make_double2(0.0, 0.0);
    if (wanted(sl.w, f)) load_zd(sl, zd);
    unsigned nearmask = 0;
    for (int k = 0; k < K; ++k) {
      int4 sl2 = make_int4(0, 0, 0, -1);
      if (k + 2 < K) sl2 = ld_stream(sp + (k + 2) * 32);
      const int32_t f1 = (!P.zero_in && sl1.w >= 0) ? ld_stream_rw(flags + sl1.w) : 0;
      const double2 z1 = zd[0], z2 = zd[1], z3 = zd[2];
      if (wanted(sl1.w, f1)) load_zd(sl1, zd);      // next row's {z, det}
      const int t = sl.w;
      const int4 tv = make_int4(va, sl.x & kVidMask, sl.y, sl.z);
      if (t >= 0) {
        if (P.do_bad && (f & MAG_BAD_QUALITY)) ++c_err;
        const bool need_bad = P.do_bad && !(f & MAG_OK_QUALITY);
        if (P.want_q || need_bad) {
          const bool owned = sl.x >= 0;
          M3 Q;
          double detQ = 0.0;
          if (USE_MAX) load_q(vq, best_vertex(tv, a_zd.y, z1.y, z2.y, z3.y), Q, detQ);
          const double2 b1 = __ldg(chunk_ptr<2>(vpos, 0, tv.y)), b2 = __ldg(chunk_ptr<2>(vpos, 0, tv.z)), b3 = __ldg(chunk_ptr<2>(vpos, 0, tv.w));
          if (!USE_MAX) {   // centroid metric (maQuality.cc:148-153)
            centroid_transform<KIND>(vedge, nv, tv, Q, &eig_any);
            detQ = FAST ? magst::det3(Q) : 0.0;
          }
          const V3 x[4] = {V3{a_xy.x, a_xy.y, a_zd.x}, V3{b1.x, b1.y, z1.x}, V3{b2.x, b2.y, z2.x}, V3{b3.x, b3.y, z3.x}};
          const double qv = FAST ? magfa::tet_quality(x, Q, detQ) : magst::tet_quality(x, Q);
          if (P.want_q) {
            st_stream(qual + t, qv);
            const unsigned long long kq = dkey(qv);
            minkey = kq < minkey ? kq : minkey;
          }
          const bool nr = need_bad && fabs(qv - P.good_q) <= P.tol_q;
          nearmask |= (nr ? 1u : 0u) << k;
          if (need_bad && !(FAST && nr)) {
            ++c_eval;
            const bool bad = qv < P.good_q;
            c_bad += (bad && owned) ? 1u : 0u;
            st_stream(flags + t, (int32_t)(f | (bad ? MAG_BAD_QUALITY : MAG_OK_QUALITY)));
          }
        }
      }
      sl = sl1;
      sl1 = sl2;
      f = f1;
    }
    for (unsigned any = __reduce_or_sync(0xffffffffu, nearmask); any; any &= any - 1) {
      const int k = __ffs(any) - 1;
      const bool nr = (nearmask >> k) & 1u;
      const int4 w = nr ? __ldg(sp + k * 32) : make_int4(0, 0, 0, -1);
      const int32_t fw = (nr && !P.zero_in && FAST) ? flags[w.w] : 0;
      const unsigned r = near_tets<KIND, FAST>(nr, w.w, elem_off, make_int4(va, w.x, w.y, w.z), fw, nv, vpos, vq, vedge, flags - elem_off,
                                               qual - elem_off, P.ops, P.good_q, P.use_max, st, near_list);
      c_eval += r & 1u; c_bad += (r >> 1) & 1u; eig_any |= (int)(r >> 2);
    }
    if (pf_b1 > pf_b0) {
      l2_prefetch_blocks<2>(vpos, pf_b0, pf_b1);
      if (USE_MAX) l2_prefetch_blocks<5>(vq, pf_b0, pf_b1);
      else l2_prefetch_blocks<EdgeRecs<KIND>::N / 2>(vedge, pf_b0, pf_b1);
    }
    w.next();
    if (MAG_ROW_STATIC) { off = off_nx; off1 = off1_nx; va = va_nx; }
    else if (w.s < nslices) { off = __ldg(slice_off + w.s); off1 = __ldg(slice_off + w.s + 1); va = __ldg(anchor + (w.s << 5) + lane); }
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

// ------------------------------------------------------------------ export-time construction of the rows
// 1. (anchor, entity) pairs sorted by anchor (stable: an anchor's entities keep the caller's order); histogram -> degrees
template <int NV>
__global__ void __launch_bounds__(kThreads)
k_row_keys(int64_t n, const int32_t* __restrict__ conn, int32_t* __restrict__ key, int32_t* __restrict__ val, int32_t* __restrict__ deg)
{
  const int64_t e = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (e >= n) return;
  const int32_t a = conn[e * NV] & kVidMask;
  key[e] = a;
  val[e] = (int32_t)e;
  atomicAdd(deg + a, 1);
}
__global__ void __launch_bounds__(kThreads)
k_row_counts(int64_t nv, const int32_t* __restrict__ deg, int32_t* __restrict__ nrows)
{
  const int64_t v = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (v <= nv) nrows[v] = v < nv ? (deg[v] + kRowMax - 1) / kRowMax : 0;
}
// 2. one record per row + its sort key: window of the anchor (major), descending length (minor)
__global__ void __launch_bounds__(kThreads)
k_row_records(int64_t nv, const int32_t* __restrict__ deg, const int32_t* __restrict__ start, const int32_t* __restrict__ rowstart,
              int32_t* __restrict__ row_anchor, int32_t* __restrict__ row_len, int32_t* __restrict__ row_first,
              int32_t* __restrict__ rkey, int32_t* __restrict__ ridx)
{
  const int64_t v = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (v >= nv) return;
  const int d = deg[v];
  int r = rowstart[v];
  for (int j = 0; j < d; j += kRowMax, ++r) {
    const int len = d - j < kRowMax ? d - j : kRowMax;
    row_anchor[r] = (int32_t)v;
    row_len[r] = len;
    row_first[r] = start[v] + j;
    rkey[r] = (int32_t)((v >> kRowWindowLog2) << 6) | (kRowMax - len);
    ridx[r] = r;
  }
}
// 3. width of every slice = its longest row (one warp per slice); out[s] = 32 * width, out[nslices] = 0 (scanned in place)
__global__ void __launch_bounds__(kThreads)
k_slice_width(int64_t nrows, int64_t nslices, const int32_t* __restrict__ order, const int32_t* __restrict__ row_len, int32_t* __restrict__ out)
{
  const int64_t s = (blockIdx.x * (int64_t)kThreads + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s > nslices) return;
  int len = 0;
  const int64_t p = s * 32 + lane;
  if (s < nslices && p < nrows) len = row_len[order[p]];
  len = __reduce_max_sync(0xffffffffu, len);
  if (lane == 0) out[s] = 32 * len;
}
// 4. the slots.  NV = 2: int2 {other | not-owned, e};  NV = 4: int4 {o1 | not-owned, o2, o3, t}
template <int NV>
__global__ void __launch_bounds__(kThreads)
k_slots_fill(int64_t nrows, int64_t nrows_pad, const int32_t* __restrict__ order, const int32_t* __restrict__ row_anchor,
             const int32_t* __restrict__ row_len, const int32_t* __restrict__ row_first, const int32_t* __restrict__ sorted_e,
             const int32_t* __restrict__ conn, const int32_t* __restrict__ slice_off, int32_t* __restrict__ anchor_out,
             int32_t* __restrict__ slots, int32_t* __restrict__ slice_vmax)
{
  const int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (p >= nrows_pad) return;                     // nrows_pad is a multiple of 32: whole warps leave together
  int vmax = 0;                                   // largest vertex id this row touches (-> L2 prefetch table)
  if (p >= nrows) anchor_out[p] = -1;
  else {
    const int r = order[p];
    vmax = anchor_out[p] = row_anchor[r];
    const int len = row_len[r], first = row_first[r];
    const int64_t base = (int64_t)slice_off[p >> 5] + (p & 31);
    for (int k = 0; k < len; ++k) {
      const int32_t e = sorted_e[first + k];
      const int32_t* cv = conn + (int64_t)e * NV;
      const int32_t notowned = cv[0] & (int32_t)0x80000000;
      const int64_t d = base + 32 * k;
      vmax = cv[1] > vmax ? cv[1] : vmax;
      if (NV == 2) {
        reinterpret_cast<int2*>(slots)[d] = make_int2(cv[1] | notowned, e);
      } else {
        reinterpret_cast<int4*>(slots)[d] = make_int4(cv[1] | notowned, cv[2], cv[3], e);
        vmax = cv[2] > vmax ? cv[2] : vmax;
        vmax = cv[3] > vmax ? cv[3] : vmax;
      }
    }
  }
  vmax = __reduce_max_sync(0xffffffffu, vmax);
  if ((p & 31) == 0) slice_vmax[p >> 5] = vmax;
}
// running maximum of vertex ids -> number of kVB-vertex blocks covering [0, vmax]
__global__ void __launch_bounds__(kThreads)
k_vmax_to_blocks(int64_t n, int32_t* __restrict__ a)
{
  const int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (i < n) a[i] = a[i] / kVB + 1;
}
