// mag_layer.cu -- ma::resetLayer on the resident part: markLayerElements + freezeLayer (ma/maLayer.cc:11-71, 94-103).
//
// Every non-simplex element (and every element the caller tags, Input::userDefinedLayerTagName, maLayer.cc:24-39) puts
// LAYER on its closure; LAYER is synchronised across part boundaries (maLayer.cc:45-46); then LAYER edges get
// DONT_COLLAPSE | DONT_SPLIT | DONT_SWAP and LAYER elements OK_QUALITY, which is what makes all three marks of the sweep
// skip the boundary layer.  Of the closure the library holds the edges and the elements (vertex / face flag words are not
// part of the sweep path; the adapter leaves them to the reference).
//
// An element knows its vertices, not its edges, so the closure needs (vertex pair) -> edge index: an open-addressing hash
// table over the part's edges, built on the device the first time it is needed and kept while the connectivity is resident
// (12 bytes per slot, 2 slots per edge).
#include "mag_internal.h"

namespace {

constexpr unsigned long long kEmpty = ~0ull;
constexpr int32_t kVidMask = 0x7fffffff;
constexpr int kT = 256;

__device__ __forceinline__ unsigned long long pair_key(int32_t a, int32_t b)
{
  const unsigned lo = (unsigned)(a < b ? a : b), hi = (unsigned)(a < b ? b : a);
  return ((unsigned long long)lo << 32) | hi;
}
__device__ __forceinline__ unsigned slot_of(unsigned long long key, int bits)
{
  return (unsigned)((key * 0x9E3779B97F4A7C15ull) >> (64 - bits));
}

__global__ void __launch_bounds__(kT)
k_pair_insert(int64_t ne, const int32_t* __restrict__ edge_v, unsigned long long* __restrict__ keys, int32_t* __restrict__ vals, int bits)
{
  const int64_t e = blockIdx.x * (int64_t)kT + threadIdx.x;
  if (e >= ne) return;
  const unsigned long long key = pair_key(edge_v[2 * e] & kVidMask, edge_v[2 * e + 1]);
  const unsigned mask = (1u << bits) - 1u;
  for (unsigned s = slot_of(key, bits);; s = (s + 1) & mask) {
    const unsigned long long prev = atomicCAS(keys + s, kEmpty, key);
    if (prev == kEmpty || prev == key) { vals[s] = (int32_t)e; return; }
  }
}
__device__ __forceinline__ int32_t pair_find(unsigned long long key, const unsigned long long* __restrict__ keys,
                                             const int32_t* __restrict__ vals, int bits)
{
  const unsigned mask = (1u << bits) - 1u;
  for (unsigned s = slot_of(key, bits);; s = (s + 1) & mask) {
    const unsigned long long k = keys[s];
    if (k == key) return vals[s];
    if (k == kEmpty) return -1;
  }
}

// apf's canonical element -> edge vertex tables (apf/apfMesh.cc:53-78)
__constant__ int c_tet_edges[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
__constant__ int c_prism_edges[9][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 4}, {2, 5}, {3, 4}, {4, 5}, {5, 3}};
__constant__ int c_pyr_edges[8][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {0, 4}, {1, 4}, {2, 4}, {3, 4}};

// one thread per element: LAYER on the element and on the edges of its closure (markLayerElements, maLayer.cc:11-40)
__global__ void __launch_bounds__(kT)
k_layer_mark(int64_t np, int64_t npy, int64_t nt, const int32_t* __restrict__ prism_v, const int32_t* __restrict__ pyr_v,
             const int32_t* __restrict__ tet_v, const int32_t* __restrict__ user_tag, const unsigned long long* __restrict__ keys,
             const int32_t* __restrict__ vals, int bits, int32_t* __restrict__ edge_flags, int32_t* __restrict__ elem_flags,
             unsigned long long* __restrict__ counters /* [0] layer elements, [1] closure edges the part does not hold */)
{
  const int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x;
  if (i >= np + npy + nt) return;
  const bool nonsimplex = i < np + npy;
  if (!nonsimplex && !(user_tag && user_tag[i])) return;
  int v[6], n_edges;
  const int (*tab)[2];
  if (i < np) {
    for (int k = 0; k < 6; ++k) v[k] = prism_v[6 * i + k] & kVidMask;
    n_edges = 9; tab = c_prism_edges;
  } else if (i < np + npy) {
    for (int k = 0; k < 5; ++k) v[k] = pyr_v[5 * (i - np) + k] & kVidMask;
    n_edges = 8; tab = c_pyr_edges;
  } else {
    for (int k = 0; k < 4; ++k) v[k] = tet_v[4 * (i - np - npy) + k] & kVidMask;
    n_edges = 6; tab = c_tet_edges;
  }
  elem_flags[i] |= MAG_LAYER;
  unsigned missing = 0;
  for (int k = 0; k < n_edges; ++k) {
    const int32_t e = pair_find(pair_key(v[tab[k][0]], v[tab[k][1]]), keys, vals, bits);
    if (e < 0) { ++missing; continue; }
    if (!(edge_flags[e] & MAG_LAYER)) atomicOr(edge_flags + e, MAG_LAYER);
  }
  atomicAdd(counters, 1ull);
  if (missing) atomicAdd(counters + 1, (unsigned long long)missing);
}
// freezeLayer (maLayer.cc:51-71) on the words the library holds
__global__ void __launch_bounds__(kT)
k_layer_freeze(int64_t n, int32_t add, int32_t* __restrict__ flags)
{
  const int64_t i = blockIdx.x * (int64_t)kT + threadIdx.x;
  if (i >= n) return;
  const int32_t f = flags[i];
  if ((f & MAG_LAYER) && (f & add) != add) flags[i] = f | add;
}

unsigned grid_for(int64_t n) { return (unsigned)((n + kT - 1) / kT); }

} // namespace

void magl_free_pairs(mag_ctx* c)
{
  cudaFree(c->d_pair_keys); cudaFree(c->d_pair_vals);
  c->d_pair_keys = nullptr; c->d_pair_vals = nullptr;
  c->pair_bits = 0;
}

static int build_pairs(mag_ctx* c)
{
  if (c->pair_bits) return MAG_OK;
  int bits = 4;
  while (((int64_t)1 << bits) < 2 * c->ne) ++bits;
  const size_t cap = (size_t)1 << bits;
  MAG_CUDA(c, cudaMalloc((void**)&c->d_pair_keys, cap * 8));
  MAG_CUDA(c, cudaMalloc((void**)&c->d_pair_vals, cap * 4));
  MAG_CUDA(c, cudaMemsetAsync(c->d_pair_keys, 0xFF, cap * 8, c->stream));
  if (c->ne) {
    k_pair_insert<<<grid_for(c->ne), kT, 0, c->stream>>>(c->ne, c->d_edge_v, c->d_pair_keys, c->d_pair_vals, bits);
    MAG_CUDA(c, cudaGetLastError());
    c->n_launches++;
  }
  c->pair_bits = bits;
  return MAG_OK;
}

extern "C" int mag_reset_layer(mag_ctx* c, const int32_t* user_layer_tag, int64_t* n_layer_elements)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  if (c->dim != 3) return mag_fail(c, MAG_ERR_ARG, "mag_reset_layer: 3-D parts only (2-D layer elements are quads, which the library does not hold)");
  const int64_t nel = c->np + c->npy + c->nt;
  int rc;
  if ((rc = magi_materialize_flags(c))) return rc;
  int32_t* d_tag = nullptr;
  if (user_layer_tag) c->tet_words_zero = false;    // a user tag may put LAYER on tets; without one only prisms / pyramids get words
  if (user_layer_tag && nel) {
    MAG_CUDA(c, cudaMalloc((void**)&d_tag, (size_t)nel * 4));
    cudaError_t e = cudaMemcpyAsync(d_tag, user_layer_tag, (size_t)nel * 4, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) { cudaFree(d_tag); return mag_fail(c, MAG_ERR_CUDA, "mag_reset_layer: %s", cudaGetErrorString(e)); }
  }
  unsigned long long h_count[2] = {0, 0};
  if ((c->np + c->npy > 0 || d_tag) && nel) {
    if ((rc = build_pairs(c))) { cudaFree(d_tag); return rc; }
    cudaMemsetAsync(c->d_layer_count, 0, 16, c->stream);
    k_layer_mark<<<grid_for(nel), kT, 0, c->stream>>>(c->np, c->npy, c->nt, c->d_prism_v, c->d_pyr_v, c->d_tet_v, d_tag, c->d_pair_keys,
                                                     c->d_pair_vals, c->pair_bits, c->d_edge_flags, c->d_elem_flags, c->d_layer_count);
    c->n_launches++;
    cudaMemcpyAsync(h_count, c->d_layer_count, 16, cudaMemcpyDeviceToHost, c->stream);
  }
  cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(d_tag);
  if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) return mag_fail(c, MAG_ERR_CUDA, "mag_reset_layer: %s", cudaGetErrorString(e));
  if (n_layer_elements) *n_layer_elements = (int64_t)h_count[0];
  if (h_count[1])
    return mag_fail(c, MAG_ERR_ARG, "mag_reset_layer: %llu edges of layer elements are not among the part's edges", h_count[1]);
  // LAYER travels to the other copies of shared edges (syncFlag, maLayer.cc:45-46) before the freeze looks at it; a->hasLayer
  // is a global property, so parts without layer elements of their own still take part
  if (!c->links.empty() && (rc = mag_sync_edge_flags(c, MAG_LAYER))) return rc;
  if (c->ne) k_layer_freeze<<<grid_for(c->ne), kT, 0, c->stream>>>(c->ne, MAG_DONT_COLLAPSE | MAG_DONT_SPLIT | MAG_DONT_SWAP, c->d_edge_flags);
  if (nel) k_layer_freeze<<<grid_for(nel), kT, 0, c->stream>>>(nel, MAG_OK_QUALITY, c->d_elem_flags);
  MAG_CUDA(c, cudaGetLastError());
  c->n_launches += 2;
  return MAG_OK;
}
