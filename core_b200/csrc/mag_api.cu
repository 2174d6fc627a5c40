// mag_api.cu -- the C ABI declared in include/mag.h: context, device memory, uploads, getters.
// No CPU fallback anywhere: every entry point needs a live CUDA device.
#include "mag_internal.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cstdlib>

int magk_pack(mag_ctx* c);
int magk_init_stats(mag_ctx* c);
int magk_vertex_pass(mag_ctx* c);
int magk_build_schedule(mag_ctx* c);
int magk_fold_owned(mag_ctx* c);
int magk_check_conn(mag_ctx* c, int32_t* d_conn, int64_t n);
int magk_conn_begin(mag_ctx* c);
int magk_conn_result(mag_ctx* c, unsigned long long* bad);
int magk_sweep(mag_ctx* c, uint32_t ops, double max_len, double min_len, double good_q, int use_max, int fp_mode);
double magk_key_to_double(unsigned long long k);
void magc_destroy(mag_ctx* c);
void magk_free_rows(mag_ctx* c);

static thread_local std::string g_create_err;

int mag_fail(mag_ctx* c, int code, const char* fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_err = buf;
  return code;
}

namespace {

template <class T>
int dev_free(mag_ctx* c, T*& p)
{
  if (p) { MAG_CUDA(c, cudaFree(p)); p = nullptr; }
  return MAG_OK;
}
template <class T>
int dev_alloc(mag_ctx* c, T*& p, size_t count)
{
  int rc = dev_free(c, p);
  if (rc) return rc;
  if (count) MAG_CUDA(c, cudaMalloc((void**)&p, count * sizeof(T)));
  return MAG_OK;
}
// grow-only buffer
template <class T>
int dev_reserve(mag_ctx* c, T*& p, size_t& cap, size_t count)
{
  if (count <= cap && p) return MAG_OK;
  int rc = dev_alloc(c, p, count);
  if (rc) return rc;
  cap = count;
  return MAG_OK;
}
template <class T>
int upload(mag_ctx* c, T* dst, const T* src, size_t count)
{
  if (count) MAG_CUDA(c, cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  return MAG_OK;
}
template <class T>
int download(mag_ctx* c, T* dst, const T* src, size_t count)
{
  if (count) MAG_CUDA(c, cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
  return MAG_OK;
}

int use_device(mag_ctx* c)
{
  MAG_CUDA(c, cudaSetDevice(c->device));
  return MAG_OK;
}

size_t rec_doubles(int kind) { return (kind == MAG_KIND_ANISO || kind == MAG_KIND_LOGM) ? 12 : 4; }

// (re)build the gather records after coordinates or the size field changed
int repack(mag_ctx* c)
{
  if (c->kind == MAG_KIND_NONE || c->nv == 0 || !c->d_xyz) return MAG_OK;
  int rc = dev_reserve(c, c->d_vedge, c->cap_vedge, (size_t)vpad(c->nv) * rec_doubles(c->kind));
  if (rc) return rc;
  c->vertex_pass_valid = false;
  if ((rc = magk_pack(c))) return rc;
  // Q_v and det Q_v depend on the coordinates and the size field only: computed here, reused by every sweep until either
  // changes again (ma/maQuality.cc:83-108 evaluates them per tet; round 1 recomputed them per sweep)
  return magk_vertex_pass(c);
}

#define CHECK_CTX(c) do { if (!(c)) return MAG_ERR_ARG; int rc_ = use_device(c); if (rc_) return rc_; } while (0)

} // namespace

// device accumulators -> the public struct; the reference's assertions on this path become return codes
int mag_stats_from_dev(mag_ctx* c, const MagDevStats& s, mag_stats* out)
{
  out->n_split = (int64_t)s.n_split;
  out->n_collapse = (int64_t)s.n_collapse;
  out->n_bad = (int64_t)s.n_bad;
  out->n_edges_evaluated = (int64_t)s.n_edges_eval;
  out->n_elems_evaluated = (int64_t)s.n_elems_eval;
  out->n_near_threshold = (int64_t)(s.n_near_edge + s.n_near_elem);
  out->n_layer_unsafe = (int64_t)s.n_layer_unsafe;
  out->n_flag_mismatch = (int64_t)s.n_flag_mismatch;
  out->min_quality = magk_key_to_double(s.min_q_key);
  memcpy(&out->max_length, &s.max_len_bits, 8);
  out->sum_length = s.sum_len;
  if (s.n_bad_conn)
    return mag_fail(c, MAG_ERR_ARG, "%llu vertex ids of the connectivity lie outside [0, nv); the results of this call are void", s.n_bad_conn);
  if (s.n_flag_err)
    return mag_fail(c, MAG_ERR_FLAG_STATE, "%llu entities already carried the flag being marked (ma::markEntities asserts, maAdapt.cc:308)", s.n_flag_err);
  if (s.n_eigen_fail)
    return mag_fail(c, MAG_ERR_EIGEN, "eigenQR failed on %llu evaluations (apf::eigen asserts convergence, apfMatrix.cc:76)", s.n_eigen_fail);
  if (s.n_nonsimplex)
    return mag_fail(c, MAG_ERR_NONSIMPLEX, "%llu prisms/pyramids reached markBadQuality without OK_QUALITY (maQuality.cc:169-182 has no entry for them)", s.n_nonsimplex);
  return MAG_OK;
}

extern "C" {

const char* mag_last_error(const mag_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int mag_create(mag_ctx** out, int device)
{
  if (!out) return mag_fail(nullptr, MAG_ERR_ARG, "mag_create: null out pointer");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return mag_fail(nullptr, MAG_ERR_CUDA, "mag_create: no CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= ndev) return mag_fail(nullptr, MAG_ERR_ARG, "mag_create: device %d out of range [0,%d)", device, ndev);
  mag_ctx* c = new mag_ctx();
  c->device = device;
  c->own_stream = c->stream = nullptr;
  c->nv = c->ne = c->nt = c->np = c->npy = c->ntri = 0;
  c->dim = 3;
  c->kind = MAG_KIND_NONE;
  c->uniform_refiner = false;
  c->vertex_pass_valid = false; c->schedule_valid = false; c->edge_flags_zero = c->elem_flags_zero = false; c->tet_words_zero = false; c->s_up = c->s_down = nullptr;
  c->d_xyz = c->d_ma = c->d_mb = c->d_vedge = c->d_vpos = c->d_vq = c->d_vqu = nullptr; c->vqu_kind = MAG_KIND_NONE;
  c->d_edge_v = c->d_tet_v = c->d_prism_v = c->d_pyr_v = c->d_tri_v = nullptr;
  c->d_edge_owned = c->d_elem_owned = nullptr;
  c->d_edge_flags = c->d_elem_flags = nullptr;
  c->d_len = c->d_qual = nullptr; c->d_weight = nullptr;
  c->d_layer_ok = c->d_layer_codes = nullptr;
  c->d_stats = nullptr; c->h_stats = nullptr;
  c->d_block_sums = nullptr; c->n_sms = 148;
  c->d_edge_order = c->d_tet_order = nullptr;
  c->erows = MagRows{0, 0, 0, nullptr, nullptr, nullptr, false, {true, true, true}};
  c->trows = c->erows;
  { const char* e = getenv("MAG_LEGACY_SWEEP"); c->legacy_sweep = e && e[0] == '1'; }
  { const char* e = getenv("MAG_LEAN_SWEEP"); c->lean_sweep = !(e && e[0] == '0'); }
  { const char* e = getenv("MAG_TET_WINNER"); c->tet_winner = !(e && e[0] == '0'); }
  c->winners_valid = false;
  c->d_vstat = nullptr;
  c->s_comm = nullptr; c->ev_comm[0] = c->ev_comm[1] = nullptr; c->comm_pending = false; c->overlap_mask = 0;
  c->d_edge_bytes = c->d_elem_bytes = nullptr;
  c->d_pair_keys = nullptr; c->d_pair_vals = nullptr; c->pair_bits = 0; c->d_layer_count = nullptr;
  c->d_near_edge = c->d_near_elem = nullptr;
  c->d_v2t_off = c->d_v2t = nullptr; c->v2t_valid = false;
  c->cap_vedge = c->cap_ma = c->cap_mb = 0;
  c->last_ops = 0; c->last_fp_mode = 0;
  c->nccl_comm = nullptr; c->nranks = 1; c->rank = 0; c->d_gather = nullptr; c->h_gather = nullptr;
  c->t_slots = c->t_used = 0; c->n_launches = 0;
  auto fail = [&](cudaError_t err, const char* what) {
    int rc = mag_fail(nullptr, MAG_ERR_CUDA, "mag_create: %s: %s", what, cudaGetErrorString(err));
    delete c;
    return rc;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(e, "cudaSetDevice");
  {
    // the temporaries and the row layout of every export come from the device's stream-ordered pool (mag_kernels.cu: Scratch);
    // freed blocks stay in the pool for the next export instead of going back to the driver (MAG_POOL_RELEASE=1: default policy)
    cudaMemPool_t pool;
    if (!getenv("MAG_POOL_RELEASE") && cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
  c->stream = c->own_stream;
  if ((e = cudaMalloc((void**)&c->d_stats, sizeof(MagDevStats))) != cudaSuccess) return fail(e, "cudaMalloc stats");
  if ((e = cudaMallocHost((void**)&c->h_stats, sizeof(MagDevStats))) != cudaSuccess) return fail(e, "cudaMallocHost stats");
  if ((e = cudaMalloc((void**)&c->d_block_sums, sizeof(double) * MAG_SUM_BLOCKS)) != cudaSuccess) return fail(e, "cudaMalloc block sums");
  if ((e = cudaMalloc((void**)&c->d_vstat, sizeof(unsigned long long))) != cudaSuccess) return fail(e, "cudaMalloc vertex-pass counter");
  if ((e = cudaMemset(c->d_vstat, 0, sizeof(unsigned long long))) != cudaSuccess) return fail(e, "cudaMemset vertex-pass counter");
  if ((e = cudaMalloc((void**)&c->d_layer_count, 2 * sizeof(unsigned long long))) != cudaSuccess) return fail(e, "cudaMalloc layer counters");
  if ((e = cudaDeviceGetAttribute(&c->n_sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return fail(e, "SM count");
  int rc = magk_init_stats(c);
  if (rc) { g_create_err = c->err; delete c; return rc; }
  *out = c;
  return MAG_OK;
}

void mag_destroy(mag_ctx* c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  magc_destroy(c);
  cudaFree(c->d_xyz); cudaFree(c->d_ma); cudaFree(c->d_mb); cudaFree(c->d_vedge); cudaFree(c->d_vpos); cudaFree(c->d_vq); cudaFree(c->d_vqu);
  cudaFree(c->d_edge_v); cudaFree(c->d_tet_v); cudaFree(c->d_prism_v); cudaFree(c->d_pyr_v); cudaFree(c->d_tri_v);
  cudaFree(c->d_edge_owned); cudaFree(c->d_elem_owned); cudaFree(c->d_edge_flags); cudaFree(c->d_elem_flags);
  cudaFree(c->d_len); cudaFree(c->d_qual); cudaFree(c->d_weight); cudaFree(c->d_layer_ok); cudaFree(c->d_layer_codes);
  cudaFree(c->d_stats); cudaFreeHost(c->h_stats); cudaFree(c->d_block_sums);
  cudaFree(c->d_near_edge); cudaFree(c->d_near_elem); cudaFree(c->d_edge_order); cudaFree(c->d_tet_order);
  cudaFree(c->d_edge_bytes); cudaFree(c->d_elem_bytes);
  cudaFree(c->d_v2t_off); cudaFree(c->d_v2t);
  magk_free_rows(c); cudaFree(c->d_vstat); magl_free_pairs(c); cudaFree(c->d_layer_count);
  for (cudaEvent_t e : c->tev) cudaEventDestroy(e);
  for (cudaEvent_t e : c->pipe_ev) cudaEventDestroy(e);
  if (c->s_up) cudaStreamDestroy(c->s_up);
  if (c->s_down) cudaStreamDestroy(c->s_down);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

int mag_set_stream(mag_ctx* c, void* cuda_stream)
{
  CHECK_CTX(c);
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
  return MAG_OK;
}

int mag_synchronize(mag_ctx* c)
{
  CHECK_CTX(c);
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  return MAG_OK;
}

// validates the shape of a part and (re)allocates every per-entity device array for it; no data is moved
int magi_reshape(mag_ctx* c, int dim, int64_t nv, int64_t ne, int64_t nt, int64_t np, int64_t npy, int64_t ntri,
                 bool has_edge_owned, bool has_elem_owned)
{
  if (nv < 0 || ne < 0 || nt < 0 || np < 0 || npy < 0 || ntri < 0) return mag_fail(c, MAG_ERR_ARG, "mag_set_mesh: negative count");
  if (nv > MAG_MAX_ENTITIES || ne > MAG_MAX_ENTITIES || np + npy + nt + ntri > MAG_MAX_ENTITIES)
    return mag_fail(c, MAG_ERR_ARG, "mag_set_mesh: entity ids are int32 (MDS_ID_TYPE=int, mds/CMakeLists.txt:7)");
  int rc;
  const int64_t nel = np + npy + nt + ntri;
  if (dim != c->dim) c->vertex_pass_valid = false;
  c->dim = dim;
  c->schedule_valid = false;
  magl_free_pairs(c);   // new connectivity follows
  const bool same_shape = nv == c->nv && ne == c->ne && nt == c->nt && np == c->np && npy == c->npy && ntri == c->ntri &&
                          has_edge_owned == (c->d_edge_owned != nullptr) && has_elem_owned == (c->d_elem_owned != nullptr);
  if (same_shape) return MAG_OK;
  // the context holds NO part while its arrays are replaced: a failed allocation below (a 50 M-tet part needs ~10 GB)
  // leaves it empty, so that later calls report MAG_ERR_ARG instead of running over half-allocated arrays
  const bool keep_field = nv == c->nv;
  c->nv = c->ne = c->nt = c->np = c->npy = c->ntri = 0;
  c->vertex_pass_valid = false;
  c->vqu_kind = MAG_KIND_NONE;
  magk_free_rows(c);
  if ((rc = dev_free(c, c->d_weight)) || (rc = dev_free(c, c->d_edge_bytes)) || (rc = dev_free(c, c->d_elem_bytes))) return rc;
  if (!keep_field) { // size field arrays are per vertex: drop them
    c->kind = MAG_KIND_NONE;
    if ((rc = dev_free(c, c->d_ma)) || (rc = dev_free(c, c->d_mb)) || (rc = dev_free(c, c->d_vedge))) return rc;
    c->cap_ma = c->cap_mb = c->cap_vedge = 0;
  }
  auto fail_empty = [&](int code) { c->kind = MAG_KIND_NONE; return code; };
  if ((rc = dev_alloc(c, c->d_xyz, (size_t)nv * 3)) || (rc = dev_alloc(c, c->d_vpos, (size_t)vpad(nv) * 4)) ||
      (rc = dev_alloc(c, c->d_vq, (size_t)vpad(nv) * 10)) || (rc = dev_alloc(c, c->d_vqu, (size_t)vpad(nv) * 10)) ||
      (rc = dev_alloc(c, c->d_edge_v, (size_t)ne * 2)) ||
      (rc = dev_alloc(c, c->d_tet_v, (size_t)nt * 4)) || (rc = dev_alloc(c, c->d_prism_v, (size_t)np * 6)) ||
      (rc = dev_alloc(c, c->d_pyr_v, (size_t)npy * 5)) || (rc = dev_alloc(c, c->d_tri_v, (size_t)ntri * 3)) ||
      (rc = dev_alloc(c, c->d_edge_owned, has_edge_owned ? (size_t)ne : 0)) ||
      (rc = dev_alloc(c, c->d_elem_owned, has_elem_owned ? (size_t)nel : 0)) ||
      (rc = dev_alloc(c, c->d_edge_flags, (size_t)ne)) || (rc = dev_alloc(c, c->d_elem_flags, (size_t)nel)) ||
      (rc = dev_alloc(c, c->d_len, (size_t)ne)) || (rc = dev_alloc(c, c->d_qual, (size_t)nel)) ||
      (rc = dev_alloc(c, c->d_layer_ok, (size_t)(np + npy))) || (rc = dev_alloc(c, c->d_layer_codes, (size_t)(np + npy))) ||
      (rc = dev_alloc(c, c->d_near_edge, (size_t)ne)) || (rc = dev_alloc(c, c->d_near_elem, (size_t)nel)))
    return fail_empty(rc);
  c->nv = nv; c->ne = ne; c->nt = nt; c->np = np; c->npy = npy; c->ntri = ntri;
  if (nel) MAG_CUDA(c, cudaMemsetAsync(c->d_qual, 0, (size_t)nel * 8, c->stream));
  return MAG_OK;
}

// the edge words only (the part-boundary exchange works on them while the element kernel still wants to see "all zero, not
// materialised" for the element words: otherwise a multi-part sweep falls back from the lean tet kernel to the tiles)
int magi_materialize_edge_flags(mag_ctx* c)
{
  if (c->edge_flags_zero && c->ne) MAG_CUDA(c, cudaMemsetAsync(c->d_edge_flags, 0, (size_t)c->ne * 4, c->stream));
  c->edge_flags_zero = false;
  return MAG_OK;
}

int magi_materialize_flags(mag_ctx* c)
{
  if (c->edge_flags_zero && c->ne) MAG_CUDA(c, cudaMemsetAsync(c->d_edge_flags, 0, (size_t)c->ne * 4, c->stream));
  const int64_t nel = c->np + c->npy + c->nt + c->ntri;
  if (c->elem_flags_zero && nel) MAG_CUDA(c, cudaMemsetAsync(c->d_elem_flags, 0, (size_t)nel * 4, c->stream));
  c->edge_flags_zero = c->elem_flags_zero = false;
  return MAG_OK;
}

// grows the raw size-field arrays and the packed gather records for `kind`
int magi_reserve_metric(mag_ctx* c, int kind, size_t na, size_t nb)
{
  int rc;
  if ((rc = dev_reserve(c, c->d_ma, c->cap_ma, na)) || (rc = dev_reserve(c, c->d_mb, c->cap_mb, nb))) return rc;
  return dev_reserve(c, c->d_vedge, c->cap_vedge, (size_t)vpad(c->nv) * rec_doubles(kind));
}

static int set_mesh_impl(mag_ctx* c, int dim, int64_t nv, const double* xyz, int64_t ne, const int32_t* edge_v,
                         int64_t nt, const int32_t* tet_v, int64_t np, const int32_t* prism_v,
                         int64_t npy, const int32_t* pyr_v, int64_t ntri, const int32_t* tri_v,
                         const uint8_t* edge_owned, const uint8_t* elem_owned)
{
  CHECK_CTX(c);
  if ((nv > 0 && !xyz) || (ne > 0 && !edge_v) || (nt > 0 && !tet_v) || (np > 0 && !prism_v) || (npy > 0 && !pyr_v) || (ntri > 0 && !tri_v))
    return mag_fail(c, MAG_ERR_ARG, "mag_set_mesh: null array with non-zero count");
  if (nv == 0 && (ne > 0 || nt > 0 || np > 0 || npy > 0 || ntri > 0)) return mag_fail(c, MAG_ERR_ARG, "mag_set_mesh: entities without vertices");
  int rc;
  // MAG_TRACE=1: wall clock of the stages of an export on stderr (each stage synchronised: diagnosis only)
  static const bool trace = getenv("MAG_TRACE") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_stage = trace ? now() : 0.0;
  auto lap = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(c->stream);
    const double t = now();
    fprintf(stderr, "[mag_set_mesh] %-14s %8.3f ms\n", what, 1e3 * (t - t_stage));
    t_stage = t;
  };
  if ((rc = magi_reshape(c, dim, nv, ne, nt, np, npy, ntri, edge_owned != nullptr, elem_owned != nullptr))) return rc;
  lap("reshape");
  const int64_t nel = np + npy + nt + ntri;
  // a new mesh starts with no flags (ma::getFlags returns 0 when the tag is absent, maAdapt.cc:80-88)
  c->edge_flags_zero = c->elem_flags_zero = true;
  c->tet_words_zero = true;
  c->v2t_valid = false;
  if ((rc = upload(c, c->d_xyz, xyz, (size_t)nv * 3)) || (rc = upload(c, c->d_edge_v, edge_v, (size_t)ne * 2)) ||
      (rc = upload(c, c->d_tet_v, tet_v, (size_t)nt * 4)) || (rc = upload(c, c->d_prism_v, prism_v, (size_t)np * 6)) ||
      (rc = upload(c, c->d_pyr_v, pyr_v, (size_t)npy * 5)) || (rc = upload(c, c->d_tri_v, tri_v, (size_t)ntri * 3)))
    return rc;
  if (edge_owned && (rc = upload(c, c->d_edge_owned, edge_owned, (size_t)ne))) return rc;
  if (elem_owned && (rc = upload(c, c->d_elem_owned, elem_owned, (size_t)nel))) return rc;
  lap("upload");
  // every vertex id must address a vertex: checked on the device before anything gathers through it
  unsigned long long bad = 0;
  if ((rc = magk_conn_begin(c)) || (rc = magk_check_conn(c, c->d_edge_v, ne * 2)) || (rc = magk_check_conn(c, c->d_tet_v, nt * 4)) ||
      (rc = magk_check_conn(c, c->d_prism_v, np * 6)) || (rc = magk_check_conn(c, c->d_pyr_v, npy * 5)) ||
      (rc = magk_check_conn(c, c->d_tri_v, ntri * 3)) || (rc = magk_conn_result(c, &bad)))
    return rc;
  if (bad) {
    c->nv = c->ne = c->nt = c->np = c->npy = c->ntri = 0;   // the part is unusable: leave the context empty
    c->kind = MAG_KIND_NONE;
    return mag_fail(c, MAG_ERR_ARG, "mag_set_mesh: %llu vertex ids outside [0, %lld)", bad, (long long)nv);
  }
  lap("check_conn");
  if ((rc = magk_fold_owned(c)) || (rc = magk_build_schedule(c))) return rc;
  lap("schedule+rows");
  c->schedule_valid = true;
  rc = repack(c);
  lap("repack");
  return rc;
}

int mag_set_mesh(mag_ctx* c, int64_t nv, const double* xyz, int64_t ne, const int32_t* edge_v,
                 int64_t nt, const int32_t* tet_v, int64_t np, const int32_t* prism_v,
                 int64_t npy, const int32_t* pyr_v, const uint8_t* edge_owned, const uint8_t* elem_owned)
{
  return set_mesh_impl(c, 3, nv, xyz, ne, edge_v, nt, tet_v, np, prism_v, npy, pyr_v, 0, nullptr, edge_owned, elem_owned);
}

int mag_set_mesh_2d(mag_ctx* c, int64_t nv, const double* xyz, int64_t ne, const int32_t* edge_v,
                    int64_t ntri, const int32_t* tri_v, const uint8_t* edge_owned, const uint8_t* elem_owned)
{
  return set_mesh_impl(c, 2, nv, xyz, ne, edge_v, 0, nullptr, 0, nullptr, 0, nullptr, ntri, tri_v, edge_owned, elem_owned);
}

int mag_set_coords(mag_ctx* c, const double* xyz)
{
  CHECK_CTX(c);
  if (!c->d_xyz || !xyz) return mag_fail(c, MAG_ERR_ARG, "mag_set_coords: no mesh set / null xyz");
  int rc = upload(c, c->d_xyz, xyz, (size_t)c->nv * 3);
  if (rc) return rc;
  return repack(c);
}

static int set_metric(mag_ctx* c, int kind, const double* a, size_t na, const double* b, size_t nb)
{
  CHECK_CTX(c);
  if (!c->d_xyz && c->nv) return mag_fail(c, MAG_ERR_ARG, "set metric: call mag_set_mesh first");
  if ((na && !a) || (nb && !b)) return mag_fail(c, MAG_ERR_ARG, "set metric: null array");
  int rc;
  if ((rc = dev_reserve(c, c->d_ma, c->cap_ma, na)) || (rc = dev_reserve(c, c->d_mb, c->cap_mb, nb))) return rc;
  if ((rc = upload(c, c->d_ma, a, na)) || (rc = upload(c, c->d_mb, b, nb))) return rc;
  c->kind = kind;
  c->uniform_refiner = false;
  return repack(c);
}
int mag_set_metric_identity(mag_ctx* c) { return set_metric(c, MAG_KIND_IDENTITY, nullptr, 0, nullptr, 0); }
int mag_set_metric_uniform_refiner(mag_ctx* c)
{
  int rc = set_metric(c, MAG_KIND_IDENTITY, nullptr, 0, nullptr, 0);
  if (rc == MAG_OK) c->uniform_refiner = true;
  return rc;
}
int mag_set_metric_iso(mag_ctx* c, const double* size) { return c ? set_metric(c, MAG_KIND_ISO, size, (size_t)c->nv, nullptr, 0) : MAG_ERR_ARG; }
int mag_set_metric_aniso(mag_ctx* c, const double* h, const double* R)
{
  return c ? set_metric(c, MAG_KIND_ANISO, h, (size_t)c->nv * 3, R, (size_t)c->nv * 9) : MAG_ERR_ARG;
}
int mag_set_metric_logm(mag_ctx* c, const double* logM) { return c ? set_metric(c, MAG_KIND_LOGM, nullptr, 0, logM, (size_t)c->nv * 9) : MAG_ERR_ARG; }

int mag_set_flags(mag_ctx* c, const int32_t* edge_flags, const int32_t* elem_flags)
{
  CHECK_CTX(c);
  const int64_t nel = c->np + c->npy + c->nt + c->ntri;
  // NULL = all zero: nothing is written now; the kernels of the next sweep skip the reads, anything else that looks
  // at the words materialises them first (magi_materialize_flags)
  if (edge_flags) { int rc = upload(c, c->d_edge_flags, edge_flags, (size_t)c->ne); if (rc) return rc; }
  c->edge_flags_zero = edge_flags == nullptr;
  if (elem_flags) { int rc = upload(c, c->d_elem_flags, elem_flags, (size_t)nel); if (rc) return rc; }
  c->elem_flags_zero = elem_flags == nullptr;
  c->tet_words_zero = elem_flags == nullptr;
  (void)nel;
  return MAG_OK;
}

int mag_sweep(mag_ctx* c, uint32_t ops, double max_len, double min_len, double good_quality, int use_max_metric, int fp_mode)
{
  CHECK_CTX(c);
  if (c->kind == MAG_KIND_NONE) return mag_fail(c, MAG_ERR_ARG, "mag_sweep: no size field set");
  if (fp_mode != MAG_FP_STRICT && fp_mode != MAG_FP_FAST && fp_mode != MAG_FP_FAST_LISTED) return mag_fail(c, MAG_ERR_ARG, "mag_sweep: bad fp_mode %d", fp_mode);
  if (ops & ~(uint32_t)(MAG_OP_ALL | MAG_OP_LENGTH_SUM)) return mag_fail(c, MAG_ERR_ARG, "mag_sweep: unknown op bits 0x%x", ops);
  if ((ops & MAG_OP_LENGTH_SUM) && !(ops & MAG_OP_LENGTHS)) return mag_fail(c, MAG_ERR_ARG, "mag_sweep: MAG_OP_LENGTH_SUM needs MAG_OP_LENGTHS");
  int rc;
  if (!c->schedule_valid) {   // the mesh came through mag_sweep_host, which defers the chunk schedule
    if ((rc = magk_build_schedule(c))) return rc;
    c->schedule_valid = true;
  }
  if ((rc = magk_init_stats(c))) return rc;
  c->last_ops = ops;
  c->last_fp_mode = fp_mode;
  return magk_sweep(c, ops, max_len, min_len, good_quality, use_max_metric, fp_mode);
}

/* mag_sweep followed by mag_reconcile_edge_flags(flag_mask), with the exchange of the part-boundary edge words overlapped with
   the element sweep: it starts on a side stream when the edge kernel has finished and the compute stream waits for it at the
   end.  Same results as the two calls in sequence. */
int mag_sweep_reconciled(mag_ctx* c, uint32_t ops, double max_len, double min_len, double good_quality, int use_max_metric, int fp_mode,
                         int32_t flag_mask)
{
  CHECK_CTX(c);
  c->overlap_mask = flag_mask;
  int rc = mag_sweep(c, ops, max_len, min_len, good_quality, use_max_metric, fp_mode);
  c->overlap_mask = 0;
  int rc2 = magc_overlap_end(c);
  return rc ? rc : rc2;
}

int mag_get_edge_lengths(mag_ctx* c, double* out)
{
  CHECK_CTX(c);
  int rc = download(c, out, c->d_len, (size_t)c->ne);
  if (rc) return rc;
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  return MAG_OK;
}
int mag_get_qualities(mag_ctx* c, double* out)
{
  CHECK_CTX(c);
  int rc = download(c, out, c->d_qual, (size_t)(c->np + c->npy + c->nt + c->ntri));
  if (rc) return rc;
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  return MAG_OK;
}
int mag_get_flags(mag_ctx* c, int32_t* edge_flags, int32_t* elem_flags)
{
  CHECK_CTX(c);
  int rc;
  if ((rc = magi_materialize_flags(c))) return rc;
  if (edge_flags && (rc = download(c, edge_flags, c->d_edge_flags, (size_t)c->ne))) return rc;
  if (elem_flags && (rc = download(c, elem_flags, c->d_elem_flags, (size_t)(c->np + c->npy + c->nt + c->ntri)))) return rc;
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  return MAG_OK;
}
int mag_get_layer_ok(mag_ctx* c, int32_t* ok, int32_t* codes)
{
  CHECK_CTX(c);
  int rc;
  if (ok && (rc = download(c, ok, c->d_layer_ok, (size_t)(c->np + c->npy)))) return rc;
  if (codes && (rc = download(c, codes, c->d_layer_codes, (size_t)(c->np + c->npy)))) return rc;
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  return MAG_OK;
}

int mag_get_stats(mag_ctx* c, mag_stats* out)
{
  CHECK_CTX(c);
  if (!out) return mag_fail(c, MAG_ERR_ARG, "mag_get_stats: null out");
  MAG_CUDA(c, cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(MagDevStats), cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  return mag_stats_from_dev(c, *c->h_stats, out);
}

int mag_timing_begin(mag_ctx* c, int max_sweeps)
{
  CHECK_CTX(c);
  if (max_sweeps < 0) return mag_fail(c, MAG_ERR_ARG, "mag_timing_begin: negative slot count");
  while ((int)c->tev.size() < 4 * max_sweeps) {
    cudaEvent_t e;
    MAG_CUDA(c, cudaEventCreate(&e));
    c->tev.push_back(e);
  }
  c->t_slots = max_sweeps;
  c->t_used = 0;
  return MAG_OK;
}
int mag_timing_read(mag_ctx* c, float* ms, int* n_out)
{
  CHECK_CTX(c);
  if (!ms || !n_out) return mag_fail(c, MAG_ERR_ARG, "mag_timing_read: null argument");
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < c->t_used; ++i)
    for (int k = 0; k < 3; ++k)
      MAG_CUDA(c, cudaEventElapsedTime(&ms[3 * i + k], c->tev[(size_t)4 * i + k], c->tev[(size_t)4 * i + k + 1]));
  *n_out = c->t_used;
  c->t_slots = c->t_used = 0;
  return MAG_OK;
}
int64_t mag_launch_count(const mag_ctx* c) { return c ? c->n_launches : -1; }

int mag_get_row_layout(mag_ctx* c, int which, int64_t* counts, int32_t* anchor, int32_t* slice_off, int32_t* slots)
{
  CHECK_CTX(c);
  if (!counts || (which != 0 && which != 1)) return mag_fail(c, MAG_ERR_ARG, "mag_get_row_layout: bad argument");
  if (c->legacy_sweep) return mag_fail(c, MAG_ERR_ARG, "mag_get_row_layout: MAG_LEGACY_SWEEP=1 runs without the row layout");
  int rc;
  if (!c->schedule_valid) {
    if ((rc = magk_build_schedule(c))) return rc;
    c->schedule_valid = true;
  }
  const MagRows& r = which ? c->trows : c->erows;
  counts[0] = r.n_rows; counts[1] = r.n_slices; counts[2] = r.n_slots;
  if (anchor && (rc = download(c, anchor, r.d_anchor, (size_t)r.n_slices * 32))) return rc;
  if (slice_off && r.n_slices && (rc = download(c, slice_off, r.d_slice_off, (size_t)r.n_slices + 1))) return rc;
  if (slots && (rc = download(c, slots, r.d_slots, (size_t)r.n_slots * (which ? 4 : 2)))) return rc;
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  if (slots && which == 1 && c->winners_valid)      // bits 29-30 of a tet slot's index hold its max-Jacobian vertex (mag_lean.cuh)
    for (int64_t i = 0; i < r.n_slots; ++i)
      if (slots[4 * i + 3] >= 0) slots[4 * i + 3] &= (1 << 29) - 1;
  return MAG_OK;
}

// host-side construction of the logM field, operation for operation as the reference
// (apf::Matrix product order, apfMatrix.h:94-106; libm log)
int mag_set_metric_logm_from_frames(mag_ctx* c, const double* h, const double* R, int variant, double* out_logM)
{
  CHECK_CTX(c);
  if (!h || !R) return mag_fail(c, MAG_ERR_ARG, "mag_set_metric_logm_from_frames: null array");
  std::vector<double> tmp;
  double* M = out_logM;
  if (!M) { tmp.resize((size_t)c->nv * 9); M = tmp.data(); }
  for (int64_t v = 0; v < c->nv; ++v) {
    const double* r = R + 9 * v;
    double s[3], T[3][3];
    for (int i = 0; i < 3; ++i) s[i] = variant ? -2 * log(h[3 * v + i]) : log(1 / h[3 * v + i] / h[3 * v + i]);
    // T = R * diag(s): r_ij = R_i0*S_0j; += R_i1*S_1j; += R_i2*S_2j with the zeros of S kept
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double a = r[3 * i + 0] * (j == 0 ? s[0] : 0.0);
        a += r[3 * i + 1] * (j == 1 ? s[1] : 0.0);
        a += r[3 * i + 2] * (j == 2 ? s[2] : 0.0);
        T[i][j] = a;
      }
    // M = T * transpose(R)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double a = T[i][0] * r[3 * j + 0];
        a += T[i][1] * r[3 * j + 1];
        a += T[i][2] * r[3 * j + 2];
        M[9 * v + 3 * i + j] = a;
      }
  }
  int rc = mag_set_metric_logm(c, M);
  if (rc) return rc;
  MAG_CUDA(c, cudaStreamSynchronize(c->stream)); // tmp dies here
  return MAG_OK;
}

int mag_get_near_threshold(mag_ctx* c, int which, int64_t* idx, int64_t cap, int64_t* n)
{
  CHECK_CTX(c);
  if (!n || (which != 0 && which != 1)) return mag_fail(c, MAG_ERR_ARG, "mag_get_near_threshold: bad argument");
  MAG_CUDA(c, cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(MagDevStats), cudaMemcpyDeviceToHost, c->stream));
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  int64_t total = (int64_t)(which ? c->h_stats->n_near_elem : c->h_stats->n_near_edge);
  *n = total;
  int64_t m = total < cap ? total : cap;
  if (idx && m > 0) {
    // the device list is int32 (MDS ids are int); widen on the host
    std::vector<int32_t> tmp((size_t)m);
    MAG_CUDA(c, cudaMemcpyAsync(tmp.data(), which ? c->d_near_elem : c->d_near_edge, (size_t)m * 4, cudaMemcpyDeviceToHost, c->stream));
    MAG_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < m; ++i) idx[i] = tmp[(size_t)i];
  }
  return MAG_OK;
}

} // extern "C"
