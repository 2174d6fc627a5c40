// mag_rows.cuh -- the anchor-row layout of a part's edges and tets, and what every kernel over it shares.
//
// Included by mag_kernels.cu after the arithmetic and the parameter structs; the kernels themselves are in mag_lean.cuh.
//
// Layout (built once at export, build_rows in mag_kernels.cu).  Every edge / tet is filed under its FIRST vertex, the anchor.
// An anchor's entities form a row (rows longer than kRowMax are cut); rows are ordered by vertex id, and inside windows of
// 2^kRowWindowLog2 vertices by descending length (SELL-C-sigma of sparse matrix-vector products: C = 32), cut into slices of 32
// rows -- one per warp, one row per lane -- and each slice is stored slot-major:
//     slot (s, k, lane)  at  slice_off[s] + 32 k + lane     k < width of slice s = its longest row
// holding the OTHER vertex ids, the ownership bit and the entity's index in the caller's order (-1 = empty).
// A lane keeps its anchor's data for the whole row and gathers only the other end of each edge (6 requests instead of 12;
// the other three vertices of a tet), and since neighbouring lanes hold neighbouring anchors whose k-th entities belong to the
// same family on any structured numbering (box meshes: the +x edge of every vertex, then +y, ...), the lanes of one request
// read consecutive records: coalesced gathers.  Lengths, qualities and flag words are written to the caller's entity order
// through the slot's entity index, so nothing else in the library (getters, reconciliation lists, the sweeps either side of
// the path) sees the layout.
//
// Entities whose value lands within 1e-12 of a threshold are re-evaluated in the reference's operation order under a
// warp-uniform branch (near_edges / near_tets below): on the structured benchmark whole families sit ON a threshold (the z
// edges of config 3 measure exactly 0.5), i.e. whole warps take the branch together; on an unstructured mesh it is almost
// never taken.
//
// History (measured on B200, n = 203, profiles/r2_kernel_history.md): general row kernels -- any flag words, any subset of the
// marks, both arithmetic modes -- were written first and lost to the round-1 tile kernels (1.43 / 0.82 ms against 1.35 / 0.80);
// handing slices out statically (slice s to warp s mod W) left the SMs idle 45 % of the launch (1.94 ms).  They are gone; the
// lean kernels of mag_lean.cuh serve the full marking sweep, the tile kernels everything else.
#pragma once

constexpr int kRowMax = 32;          // longest row; an anchor with more entities gets several rows
constexpr int kRowWindowLog2 = 11;   // rows are sorted by length inside windows of 2048 vertices

#ifndef MAG_EROW_BLOCKS_LOGM
#define MAG_EROW_BLOCKS_LOGM 2
#endif
// lane 0 draws a ticket from an atomic counter; the value is broadcast (and thereby waited for) only where it is needed
// (PTX atomic: nvcc turns atomicAdd under `lane == 0` into its warp-aggregated form -- vote, ATOMG, and a SHFL that broadcasts the
//  result AT ONCE, i.e. the warp waits ~1 us for every ticket it meant to draw ahead of time: 6 % of the edge kernel's warp time
//  in the r2f profile.  Written as PTX the result stays in lane 0's register until GroupWalk::next_slice shuffles it.)
#ifndef MAG_TICKET_PTX
#define MAG_TICKET_PTX 1
#endif
struct SliceWalk {
  static __device__ __forceinline__ int issue(unsigned long long* counter)
  {
    unsigned long long t = 0;
#if MAG_TICKET_PTX
    if ((threadIdx.x & 31) == 0) asm volatile("atom.global.add.u64 %0, [%1], 1;" : "=l"(t) : "l"(counter) : "memory");
#else
    if ((threadIdx.x & 31) == 0) t = atomicAdd(counter, 1ull);
#endif
    return (int)t;
  }
};
// the slot words are the one stream of these kernels that comes straight from DRAM, a row (256 / 512 bytes) at a time; they
// are loaded one row before the gather they address is issued, and a DRAM access under load takes longer than a row (r2f
// profile: 15 % of the warp time of both kernels is spent waiting for a slot word).  Rows further ahead are pulled into L2.
#ifndef MAG_SLOT_PREFETCH_E
#define MAG_SLOT_PREFETCH_E 3   /* edges: rows ahead; 0 = off.  r2t / r2u (jittered / lattice): off 0.846 / 1.235 ms, 2: 0.846 / 1.220, 3: 0.786 / 1.173, 4: 0.844 / 1.237, 6: 0.830 / 1.223, 10: 0.843 / 1.238 */
#endif
#ifndef MAG_SLOT_PREFETCH_T
#define MAG_SLOT_PREFETCH_T 0   /* tets: 0.855 / 0.784 ms off, 3: 0.891 / 0.806, 6: 0.866 / 0.781 -- the slot wait only gives way to the wait for the winner's transform */
#endif
// (an L2 prefetch of the other-end RECORDS two rows before their gather, every eighth lane asking for its lines, was measured
//  too, r2u: 0.826 - 0.831 ms against 0.786 without it)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

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
                                            double max_len, double min_len, MagDevStats* st, int32_t* __restrict__ near_list,
                                            bool reeval = true, const double* __restrict__ vqu = nullptr)
{
  const unsigned m = __ballot_sync(0xffffffffu, nr);
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(&st->n_near_edge, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  unsigned out = 0;
  if (nr) {
    near_list[base + __popc(m & ((1u << lane) - 1u))] = e;
    if (FAST && reeval) {
      const SweepParams P{ops, max_len, min_len, 0.0, 0};
      const bool need_split = (ops & MAG_OP_MARK_SPLIT) && !(f & kSkipSplit);
      const bool need_coll = (ops & MAG_OP_MARK_COLLAPSE) && !(f & kSkipColl);
      EdgeRecs<KIND> R;
      load_edge_recs<KIND>(vedge, make_int2(va, vb_word & kVidMask), R);
      int eig = 0;
      const double len = edge_length_strict<KIND>(R, &eig, vqu, va);
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
             int32_t* __restrict__ slots)
{
  const int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (p >= nrows_pad) return;
  if (p >= nrows) { anchor_out[p] = -1; return; }
  const int r = order[p];
  anchor_out[p] = row_anchor[r];
  const int len = row_len[r], first = row_first[r];
  const int64_t base = (int64_t)slice_off[p >> 5] + (p & 31);
  for (int k = 0; k < len; ++k) {
    const int32_t e = sorted_e[first + k];
    const int32_t* cv = conn + (int64_t)e * NV;
    const int32_t notowned = cv[0] & (int32_t)0x80000000;
    const int64_t d = base + 32 * k;
    if (NV == 2) {
      reinterpret_cast<int2*>(slots)[d] = make_int2(cv[1] | notowned, e);
    } else {
      reinterpret_cast<int4*>(slots)[d] = make_int4(cv[1] | notowned, cv[2], cv[3], e);
    }
  }
}
