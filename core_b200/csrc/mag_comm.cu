// mag_comm.cu -- part-boundary flag exchange and global statistics over NCCL (NVLink 5 / NVSwitch).
// Replaces, on the sweep path only, PCU's phased p2p exchange in ma::checkFlagConsistency /
// ma::syncFlag (ma/maAdapt.cc:226-256, 498-520) and the PCU Add / Min / Max scalar reductions
// (ma/maAdapt.cc:323, ma/maShape.cc:168, ma/maSize.cc:689).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a process that already loaded a NCCL (e.g.
// through torch.distributed) shares that one; a single-GPU user never needs NCCL installed.
#include "mag_internal.h"
#include <dlfcn.h>
#include <cstring>
#include <nccl.h>

namespace {

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};
NcclApi g_nccl;

bool load_nccl()
{
  if (g_nccl.h) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { g_nccl.err = dlerror(); return false; }
#define BIND(name)                                                                 \
  *(void**)(&g_nccl.name) = dlsym(h, "nccl" #name);                                \
  if (!g_nccl.name) { g_nccl.err = "missing symbol nccl" #name; dlclose(h); return false; }
  BIND(GetUniqueId) BIND(CommInitRank) BIND(CommDestroy) BIND(Send) BIND(Recv) BIND(AllGather)
  BIND(GroupStart) BIND(GroupEnd) BIND(GetErrorString)
#undef BIND
  g_nccl.h = h;
  return true;
}

#define MAG_NCCL(c, call)                                                                       \
  do {                                                                                          \
    ncclResult_t r_ = (call);                                                                   \
    if (r_ != ncclSuccess)                                                                      \
      return mag_fail((c), MAG_ERR_NCCL, "%s:%d: %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
  } while (0)

__global__ void k_gather_flags(int64_t n, const int32_t* __restrict__ idx, const int32_t* __restrict__ flags,
                               int32_t mask, int32_t* __restrict__ out)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = flags[idx[i]] & mask;
}
// mode 0: compare, count mismatches, take the peer's bits where the peer owns the entity (owner wins)
// mode 1: OR the peer's bits in (ma::syncFlag)
__global__ void k_merge_flags(int64_t n, const int32_t* __restrict__ idx, const int32_t* __restrict__ recv,
                              const uint8_t* __restrict__ peer_owns, int32_t mask, int mode,
                              int32_t* __restrict__ flags, MagDevStats* st)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t e = idx[i];
  int32_t mine = flags[e], theirs = recv[i] & mask;
  if (mode == 1) {
    if (theirs & ~mine) atomicOr(&flags[e], theirs); // an edge can sit in several peers' lists
    return;
  }
  if ((mine & mask) != theirs) {
    atomicAdd(&st->n_flag_mismatch, 1ull);
    if (mode == 0 && peer_owns && peer_owns[i]) flags[e] = (mine & ~mask) | theirs;   // mode 2: count only
  }
}

int exchange(mag_ctx* c, int32_t mask, int mode, cudaStream_t stream)
{
  if (c->links.empty()) return MAG_OK;
  if (!c->nccl_comm) return mag_fail(c, MAG_ERR_ARG, "flag exchange: call mag_comm_init first");
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  // the exchange reads and writes EDGE words only: the element words may stay "all zero, not materialised", which is what lets the
  // element sweep that follows (or runs beside it, mag_sweep_reconciled) use the lean tet kernel.  Until r2z this materialised
  // both, and every multi-part sweep ran the tile tet kernel: 0.82 instead of 0.67 ms.
  { int rc = magi_materialize_edge_flags(c); if (rc) return rc; }
  for (auto& L : c->links)
    if (L.n) k_gather_flags<<<(unsigned)((L.n + 255) / 256), 256, 0, stream>>>(L.n, L.d_idx, c->d_edge_flags, mask, L.d_send);
  MAG_CUDA(c, cudaGetLastError());
  MAG_NCCL(c, g_nccl.GroupStart());
  for (auto& L : c->links) {
    if (!L.n) continue;
    MAG_NCCL(c, g_nccl.Send(L.d_send, (size_t)L.n, ncclInt32, L.peer, comm, stream));
    MAG_NCCL(c, g_nccl.Recv(L.d_recv, (size_t)L.n, ncclInt32, L.peer, comm, stream));
  }
  MAG_NCCL(c, g_nccl.GroupEnd());
  for (auto& L : c->links)
    if (L.n) k_merge_flags<<<(unsigned)((L.n + 255) / 256), 256, 0, stream>>>(L.n, L.d_idx, L.d_recv, L.d_peer_owns, mask, mode, c->d_edge_flags, c->d_stats);
  MAG_CUDA(c, cudaGetLastError());
  return MAG_OK;
}

void free_links(mag_ctx* c)
{
  for (auto& L : c->links) { cudaFree(L.d_idx); cudaFree(L.d_send); cudaFree(L.d_recv); cudaFree(L.d_peer_owns); }
  c->links.clear();
}

} // namespace

// The part-boundary exchange of the edge marks, started on a side stream as soon as the edge kernel has finished, so that it
// runs under the element kernel (an edge's marks are final once the edge sweep ends).  magc_overlap_end makes the compute
// stream wait for it.  Measured at N = 2 (bench.py multi_part_overhead): 0.13 ms when it runs after the sweep.
int magc_overlap_begin(mag_ctx* c, int32_t mask)
{
  if (c->links.empty() || !c->nccl_comm) return MAG_OK;
  if (!c->s_comm) {   // highest priority: its few blocks go first whenever an SM has room
    int lo = 0, hi = 0;
    MAG_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    MAG_CUDA(c, cudaStreamCreateWithPriority(&c->s_comm, cudaStreamNonBlocking, hi));
  }
  if (!c->ev_comm[0]) {
    MAG_CUDA(c, cudaEventCreateWithFlags(&c->ev_comm[0], cudaEventDisableTiming));
    MAG_CUDA(c, cudaEventCreateWithFlags(&c->ev_comm[1], cudaEventDisableTiming));
  }
  { int rc = magi_materialize_edge_flags(c); if (rc) return rc; }   // on the compute stream, before the hand-over (edge words only)
  MAG_CUDA(c, cudaEventRecord(c->ev_comm[0], c->stream));
  MAG_CUDA(c, cudaStreamWaitEvent(c->s_comm, c->ev_comm[0], 0));
  int rc = exchange(c, mask, 0, c->s_comm);
  if (rc) return rc;
  MAG_CUDA(c, cudaEventRecord(c->ev_comm[1], c->s_comm));
  c->comm_pending = true;
  return MAG_OK;
}
int magc_overlap_end(mag_ctx* c)
{
  if (!c->comm_pending) return MAG_OK;
  c->comm_pending = false;
  MAG_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_comm[1], 0));
  return MAG_OK;
}

void magc_destroy(mag_ctx* c)
{
  if (c->s_comm) { cudaStreamSynchronize(c->s_comm); cudaStreamDestroy(c->s_comm); c->s_comm = nullptr; }
  for (int i = 0; i < 2; ++i) if (c->ev_comm[i]) { cudaEventDestroy(c->ev_comm[i]); c->ev_comm[i] = nullptr; }
  free_links(c);
  cudaFree(c->d_gather); c->d_gather = nullptr;
  cudaFreeHost(c->h_gather); c->h_gather = nullptr;
  if (c->nccl_comm && g_nccl.h) g_nccl.CommDestroy((ncclComm_t)c->nccl_comm);
  c->nccl_comm = nullptr;
}

extern "C" {

int mag_comm_unique_id(void* out_id)
{
  static_assert(sizeof(ncclUniqueId) <= MAG_UNIQUE_ID_BYTES, "unique id size");
  if (!out_id) return MAG_ERR_ARG;
  if (!load_nccl()) return mag_fail(nullptr, MAG_ERR_NCCL, "cannot load NCCL: %s", g_nccl.err.c_str());
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return mag_fail(nullptr, MAG_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
  memset(out_id, 0, MAG_UNIQUE_ID_BYTES);
  memcpy(out_id, &id, sizeof(id));
  return MAG_OK;
}

int mag_comm_init(mag_ctx* c, int nranks, int rank, const void* unique_id)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  if (nranks < 1 || rank < 0 || rank >= nranks || !unique_id) return mag_fail(c, MAG_ERR_ARG, "mag_comm_init: bad arguments");
  if (!load_nccl()) return mag_fail(c, MAG_ERR_NCCL, "cannot load NCCL: %s", g_nccl.err.c_str());
  if (c->nccl_comm) { g_nccl.CommDestroy((ncclComm_t)c->nccl_comm); c->nccl_comm = nullptr; }
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclComm_t comm;
  MAG_NCCL(c, g_nccl.CommInitRank(&comm, nranks, id, rank));
  c->nccl_comm = comm;
  c->nranks = nranks;
  c->rank = rank;
  if (c->d_gather) { cudaFree(c->d_gather); c->d_gather = nullptr; }
  if (c->h_gather) { cudaFreeHost(c->h_gather); c->h_gather = nullptr; }
  MAG_CUDA(c, cudaMalloc((void**)&c->d_gather, sizeof(MagDevStats) * (size_t)nranks));
  MAG_CUDA(c, cudaMallocHost((void**)&c->h_gather, sizeof(MagDevStats) * (size_t)nranks));
  return MAG_OK;
}

int mag_set_edge_links(mag_ctx* c, int npeers, const int32_t* peer, const int64_t* n, const int32_t* const* idx,
                       const uint8_t* const* peer_owns)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  if (npeers < 0 || (npeers && (!peer || !n || !idx))) return mag_fail(c, MAG_ERR_ARG, "mag_set_edge_links: bad arguments");
  for (int k = 0; k < npeers; ++k) {   // the lists are boundary-sized: checked on the host
    if (n[k] < 0 || (n[k] && !idx[k])) return mag_fail(c, MAG_ERR_ARG, "mag_set_edge_links: bad list for peer %d", peer[k]);
    for (int64_t i = 0; i < n[k]; ++i)
      if (idx[k][i] < 0 || idx[k][i] >= c->ne)
        return mag_fail(c, MAG_ERR_ARG, "mag_set_edge_links: edge index %d (peer %d, entry %lld) outside [0, %lld)", idx[k][i], peer[k],
                        (long long)i, (long long)c->ne);
  }
  free_links(c);
  for (int k = 0; k < npeers; ++k) {
    MagLinks L;
    L.peer = peer[k];
    L.n = n[k];
    L.d_idx = L.d_send = L.d_recv = nullptr;
    L.d_peer_owns = nullptr;
    c->links.push_back(L);               // registered first: free_links releases whatever was allocated if a step below fails
    MagLinks& R = c->links.back();
    cudaError_t e = cudaSuccess;
    if (R.n) {
      if (e == cudaSuccess) e = cudaMalloc((void**)&R.d_idx, (size_t)R.n * 4);
      if (e == cudaSuccess) e = cudaMalloc((void**)&R.d_send, (size_t)R.n * 4);
      if (e == cudaSuccess) e = cudaMalloc((void**)&R.d_recv, (size_t)R.n * 4);
      if (e == cudaSuccess) e = cudaMemcpyAsync(R.d_idx, idx[k], (size_t)R.n * 4, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess && peer_owns && peer_owns[k]) {
        e = cudaMalloc((void**)&R.d_peer_owns, (size_t)R.n);
        if (e == cudaSuccess) e = cudaMemcpyAsync(R.d_peer_owns, peer_owns[k], (size_t)R.n, cudaMemcpyHostToDevice, c->stream);
      }
    }
    if (e != cudaSuccess) {
      cudaStreamSynchronize(c->stream);
      free_links(c);
      return mag_fail(c, MAG_ERR_CUDA, "mag_set_edge_links: %s", cudaGetErrorString(e));
    }
  }
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  return MAG_OK;
}

int mag_reconcile_edge_flags(mag_ctx* c, int32_t flag_mask)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  return exchange(c, flag_mask, 0, c->stream);
}
/* ma::checkFlagConsistency (maAdapt.cc:226-256): the reference asserts that every copy of a shared edge carries the same
   bits; here the copies are compared (nothing is repaired) and a disagreement is MAG_ERR_INCONSISTENT */
int mag_check_edge_flag_consistency(mag_ctx* c, int32_t flag_mask, int64_t* n_mismatch)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  unsigned long long before = 0, after = 0;
  MAG_CUDA(c, cudaMemcpyAsync(&before, &c->d_stats->n_flag_mismatch, sizeof(before), cudaMemcpyDeviceToHost, c->stream));
  int rc = exchange(c, flag_mask, 2, c->stream);
  if (rc) return rc;
  MAG_CUDA(c, cudaMemcpyAsync(&after, &c->d_stats->n_flag_mismatch, sizeof(after), cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (n_mismatch) *n_mismatch = (int64_t)(after - before);
  if (after != before)
    return mag_fail(c, MAG_ERR_INCONSISTENT, "%llu copies of shared edges disagree with their peer's copy under mask 0x%x", after - before, flag_mask);
  return MAG_OK;
}
int mag_sync_edge_flags(mag_ctx* c, int32_t flag_mask)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  return exchange(c, flag_mask, 1, c->stream);
}

int mag_allreduce_stats(mag_ctx* c, mag_stats* global)
{
  if (!c || !global) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  if (c->nranks == 1 || !c->nccl_comm) return mag_get_stats(c, global);
  // one all-gather of the device accumulators (sum / min / max are different operators, so a single allreduce cannot
  // carry them), one copy to pinned memory, one host synchronisation; reduced locally, identically on every rank
  const size_t sz = sizeof(MagDevStats);
  MAG_NCCL(c, g_nccl.AllGather(c->d_stats, c->d_gather, sz, ncclChar, (ncclComm_t)c->nccl_comm, c->stream));
  MAG_CUDA(c, cudaMemcpyAsync(c->h_gather, c->d_gather, sz * (size_t)c->nranks, cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  MagDevStats g = c->h_gather[0];
  for (int r = 1; r < c->nranks; ++r) {
    const MagDevStats& s = c->h_gather[r];
    g.n_split += s.n_split; g.n_collapse += s.n_collapse; g.n_bad += s.n_bad;
    g.n_edges_eval += s.n_edges_eval; g.n_elems_eval += s.n_elems_eval;
    g.n_near_edge += s.n_near_edge; g.n_near_elem += s.n_near_elem; g.n_layer_unsafe += s.n_layer_unsafe;
    g.n_flag_err += s.n_flag_err; g.n_eigen_fail += s.n_eigen_fail; g.n_nonsimplex += s.n_nonsimplex;
    g.n_flag_mismatch += s.n_flag_mismatch;
    g.n_bad_conn += s.n_bad_conn;
    if (s.min_q_key < g.min_q_key) g.min_q_key = s.min_q_key;
    if (s.max_len_bits > g.max_len_bits) g.max_len_bits = s.max_len_bits;
    g.sum_len += s.sum_len;
  }
  return mag_stats_from_dev(c, g, global);
}

} // extern "C"
