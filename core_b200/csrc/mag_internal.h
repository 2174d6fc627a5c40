// mag_internal.h -- context layout shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>
#include "../../include/mag.h"

enum MagKind { MAG_KIND_NONE = -1, MAG_KIND_IDENTITY = 0, MAG_KIND_ISO = 1, MAG_KIND_ANISO = 2, MAG_KIND_LOGM = 3 };

// device-side accumulators of one sweep (zeroed by mag_sweep)
struct MagDevStats {
  unsigned long long n_split, n_collapse, n_bad;
  unsigned long long n_edges_eval, n_elems_eval;
  unsigned long long n_near_edge, n_near_elem;
  unsigned long long n_layer_unsafe;
  unsigned long long n_flag_err, n_eigen_fail, n_nonsimplex;
  unsigned long long n_flag_mismatch;
  unsigned long long n_bad_conn;   // vertex ids outside [0, nv) met by the export-time check (and replaced by 0)
  unsigned long long edge_chunk, elem_chunk; // work-distribution counters of the persistent kernels
  unsigned long long max_len_bits; // bits of a non-negative double: integer order == fp order
  unsigned long long min_q_key;    // order-preserving key of a double (see dkey())
  double sum_len;
  unsigned long long n_eigen_aux;  // eigen-solver failures of the auxiliary sweeps (weights, cavities, sliver codes): their own
                                   // counter, so a deferred MAG_ERR_EIGEN of the last mag_sweep survives them
};

#define MAG_SUM_BLOCKS 1184 /* 8 x 148: fixed shape of the length-sum tree */
#define MAG_NEAR_REL 1e-12
// per-vertex device arrays are stored in blocks of MAG_VBLOCK vertices (see mag_kernels.cu) and padded to whole blocks
#define MAG_VBLOCK 32
static inline int64_t vpad(int64_t nv) { return (nv + MAG_VBLOCK - 1) / MAG_VBLOCK * MAG_VBLOCK; }
// entity ids are int32 (MDS_ID_TYPE=int, mds/CMakeLists.txt:7); the kernels index one chunk past the end in int32
#define MAG_MAX_ENTITIES (0x7fffffffLL - (1LL << 20))
// transient bit (never visible to the caller): entity awaits strict re-evaluation
#define MAG_PENDING_BIT (1 << 30)

// Anchor-row layout of one entity dimension (mag_rows.cuh): every entity is filed under its FIRST vertex (the anchor); the
// rows are cut into slices of 32 (one per warp), stored slot-major inside a slice (sliced ELLPACK).  Built at export.
struct MagRows {
  int64_t n_rows;        // rows holding at least one entity
  int64_t n_slices;      // ceil(n_rows / 32)
  int64_t n_slots;       // 32 * sum of the slice widths
  int32_t* d_anchor;     // [32 n_slices] anchor vertex of row position p, -1 = padding
  int32_t* d_slice_off;  // [n_slices + 1] first slot of every slice
  int32_t* d_slots;      // edges: int2 {other vertex | not-owned << 31, edge id}; tets: int4 {o1 | not-owned << 31, o2, o3, tet id}; id -1 = empty
  bool valid;
  bool pooled[3];        // d_anchor / d_slice_off / d_slots came from the stream-ordered pool (small blocks) or from cudaMalloc
};

struct MagLinks {
  int peer;
  int64_t n;
  int32_t* d_idx;       // local edge indices shared with this peer
  int32_t* d_send;      // packed flag words to send
  int32_t* d_recv;      // packed flag words received
  uint8_t* d_peer_owns; // optional: 1 where the peer's copy is the owner
};

struct mag_ctx {
  int device;
  cudaStream_t own_stream, stream;
  std::string err;

  int64_t nv, ne, nt, np, npy;
  int64_t ntri; // 2-D meshes: the elements are triangles (nt = np = npy = 0)
  int dim;      // mesh dimension (3, or 2 after mag_set_mesh_2d)
  int kind;
  bool uniform_refiner; // identity kind only: ma::UniformRefiner, shouldSplit constant true (maSize.h:75-85)
  bool vertex_pass_valid;
  // "ma_flags" words are logically all zero (ma::getFlags returns 0 when the tag is absent, maAdapt.cc:80-88) but the
  // device arrays have not been zeroed: the whole-part edge / tet kernels then skip reading them; every other consumer
  // calls magi_materialize_flags first
  bool edge_flags_zero, elem_flags_zero;
  // the words of the TETS are all zero (logically, or as materialised zeros), whatever the layer elements in front of them
  // carry: after mag_reset_layer on a mixed part the element words are materialised (LAYER | OK_QUALITY on prisms / pyramids)
  // but no tet has been touched, and the element sweep may still use the lean tet kernel, which writes every tet word from zero
  bool tet_words_zero;
  bool schedule_valid;   // d_edge_order / d_tet_order match the resident connectivity

  // raw uploads (kept so coordinates or metric can be replaced independently)
  double* d_xyz;   // [nv][3]
  double* d_ma;    // iso: s[nv]; aniso: h[nv][3]
  double* d_mb;    // aniso: R[nv][9]; logm: logM[nv][9]
  // packed gather records
  double* d_vedge; // per vertex: iso/identity 4 doubles {x,y,z,s}; aniso/logm 12 doubles
  double* d_vpos;  // per vertex 4 doubles {x,y,z,det Q_v}
  double* d_vq;    // per vertex 10 doubles {Q_v row-major, det Q_v}
  double* d_vqu;   // per vertex 10 doubles {Q_u row-major, eigen-solver failure}: the transform both Gauss points of an edge see whose
                   // two ends carry THIS vertex's size-field values (k_vertex_uniform); valid with the vertex pass, for vqu_kind
  int vqu_kind;    // size-field kind d_vqu was computed for (MAG_KIND_NONE: none)
  int32_t* d_edge_v;  // [ne][2]
  int32_t* d_tet_v;   // [nt][4]
  int32_t* d_prism_v; // [np][6]
  int32_t* d_pyr_v;   // [npy][5]
  int32_t* d_tri_v;   // [ntri][3]
  uint8_t* d_edge_owned; // may be null
  uint8_t* d_elem_owned; // may be null
  int32_t* d_edge_flags; // [ne]
  int32_t* d_elem_flags; // [np+npy+nt]
  double* d_len;         // [ne]
  double* d_qual;        // [np+npy+nt]
  double* d_weight;      // [np+npy+nt] element weights of the last mag_element_weights (allocated on first use)
  int32_t* d_layer_ok;   // [np+npy]
  int32_t* d_layer_codes;
  MagDevStats* d_stats;
  MagDevStats* h_stats;  // pinned
  double* d_block_sums;  // [MAG_SUM_BLOCKS] partial sums of the edge lengths (MAG_OP_LENGTH_SUM)
  int32_t* d_v2t_off;    // [nv+1] vertex -> tet incidence (CSR) of mag_collapse_quality, built on first use per mesh
  int32_t* d_v2t;        // [4 nt]
  bool v2t_valid;
  int32_t* d_near_edge;  // [ne]  near-threshold edge indices of the last sweep (an entity is listed at most once)
  int32_t* d_near_elem;  // [np+npy+nt]
  int n_sms;
  int32_t* d_edge_order; // chunk schedule of the legacy edge kernel (tile indices sorted by smallest vertex id)
  int32_t* d_tet_order;
  bool lean_sweep;       // MAG_LEAN_SWEEP=0: never use the lean kernels of mag_lean.cuh (A/B measurements, tests)
  MagRows erows, trows;  // anchor-row layout of the edges / tets (whole-part sweeps)
  bool tet_winner;       // MAG_TET_WINNER=0: the tet rows keep the dependent gather of round-2f (A/B measurements); default on
  bool winners_valid;    // the slot words of trows carry the max-Jacobian vertex of their tet (bits 29-30) for the current det Q_v
  bool legacy_sweep;     // MAG_LEGACY_SWEEP=1: whole-part sweeps run the round-1 tile kernels (A/B measurements)
  // (vertex pair) -> edge index hash table of mag_reset_layer (mag_layer.cu); pair_bits = 0: not built
  unsigned long long* d_pair_keys;
  int32_t* d_pair_vals;
  int pair_bits;
  unsigned long long* d_layer_count; // [2]
  uint8_t* d_edge_bytes;  // [ne] mark bytes in transit (mag_set/get_mark_bytes, mag_resweep_host); allocated on first use
  uint8_t* d_elem_bytes;  // [np+npy+nt+ntri]
  unsigned long long* d_vstat; // [1] eigen-solver failures of the cached per-vertex pass (folded into every sweep's statistics)
  size_t cap_vedge, cap_ma, cap_mb;

  std::map<const void*, int> occupancy; // resident blocks per SM of the persistent kernels (queried once per context)

  // last sweep parameters (for the near-threshold fix-up and getters)
  uint32_t last_ops;
  int last_fp_mode;

  // instrumentation
  std::vector<cudaEvent_t> tev; // 4 events per armed sweep slot
  int t_slots, t_used;
  int64_t n_launches;

  // mag_sweep_host pipeline: upload / download streams and an event pool
  cudaStream_t s_up, s_down;
  std::vector<cudaEvent_t> pipe_ev;

  // multi-GPU
  cudaStream_t s_comm;       // side stream of the overlapped flag exchange (mag_sweep_reconciled)
  cudaEvent_t ev_comm[2];
  bool comm_pending;
  int32_t overlap_mask;      // != 0 while mag_sweep_reconciled runs: magk_sweep starts the exchange right after the edge kernel
  void* nccl_comm;
  int nranks, rank;
  std::vector<MagLinks> links;
  MagDevStats* d_gather; // [nranks] all-gathered accumulators (mag_allreduce_stats)
  MagDevStats* h_gather; // pinned
};

int mag_fail(mag_ctx* c, int code, const char* fmt, ...);
extern "C" {
int magi_reshape(mag_ctx* c, int dim, int64_t nv, int64_t ne, int64_t nt, int64_t np, int64_t npy, int64_t ntri,
                 bool has_edge_owned, bool has_elem_owned);
int magi_materialize_flags(mag_ctx* c);
int magi_materialize_edge_flags(mag_ctx* c);
int magi_reserve_metric(mag_ctx* c, int kind, size_t na, size_t nb);
}
void magl_free_pairs(mag_ctx* c);
int magc_overlap_begin(mag_ctx* c, int32_t mask);
int magc_overlap_end(mag_ctx* c);
int mag_stats_from_dev(mag_ctx* c, const MagDevStats& s, mag_stats* out);
#define MAG_CUDA(c, call)                                                                      \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return mag_fail((c), MAG_ERR_CUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)
