// mag_host.cu -- mag_sweep_host: export + sweep + results of one part in ONE call, streamed.
//
// What the drop-in adapter does once per MeshAdapt iteration -- export the (changed) mesh, the size field and the flag
// words, sweep, read lengths / qualities / flags back -- is bound by the host link, not by the kernels (n = 203 part:
// 2.7 GB up, 1.3 GB down at ~55 GB/s against 2.5 ms of kernels).  Done as separate calls the downloads can only start
// after the last upload.  This entry point cuts the edges and the tets into slices and runs three streams:
//     upload stream     vertex data | edge slice 0 | edge slice 1 | ... | tet slice 0 | tet slice 1 | ...
//     compute stream                  pack, vertex pass | k_edges(slice 0) | ...        | k_tets(slice 0) | ...
//     download stream                                     | lengths+flags(slice 0) | ...   | qualities+flags(slice 0) | ...
// so the device->host traffic rides under the host->device traffic (PCIe is full duplex) and the call costs about the
// upload time alone.  Results are bit-identical to mag_set_mesh + mag_set_metric_* + mag_set_flags + mag_sweep + getters
// (the slices only change the order in which entities are visited); the part stays resident afterwards.
#include "mag_internal.h"
#include <algorithm>
#include <cmath>

int magk_pack(mag_ctx* c);
int magk_init_stats(mag_ctx* c);
int magk_vertex_pass(mag_ctx* c);
int magk_edges_range(mag_ctx* c, uint32_t ops, double max_len, double min_len, int fp_mode, int64_t first, int64_t n);
int magk_tets_range(mag_ctx* c, uint32_t ops, double good_q, int use_max, int fp_mode, int64_t first, int64_t n);
int magk_length_sum(mag_ctx* c);
int magk_check_conn(mag_ctx* c, int32_t* d_conn, int64_t n);
int magk_build_schedule(mag_ctx* c);
int magk_sweep(mag_ctx* c, uint32_t ops, double max_len, double min_len, double good_q, int use_max, int fp_mode);
int magk_marks_expand(mag_ctx* c, cudaStream_t s, const uint8_t* d_bytes, int32_t* d_words, int64_t n);
int magk_marks_compress(mag_ctx* c, cudaStream_t s, const int32_t* d_words, uint8_t* d_bytes, int64_t n);

namespace {

// event pool: events are reused across calls (cudaEventDisableTiming: ordering only)
struct Events {
  mag_ctx* c;
  size_t next = 0;
  int get(cudaEvent_t* e)
  {
    if (next == c->pipe_ev.size()) {
      cudaEvent_t ev;
      MAG_CUDA(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      c->pipe_ev.push_back(ev);
    }
    *e = c->pipe_ev[next++];
    return MAG_OK;
  }
};

// `later` waits for everything enqueued on `earlier` so far
int chain(mag_ctx* c, Events& ev, cudaStream_t earlier, cudaStream_t later)
{
  cudaEvent_t e;
  int rc = ev.get(&e);
  if (rc) return rc;
  MAG_CUDA(c, cudaEventRecord(e, earlier));
  MAG_CUDA(c, cudaStreamWaitEvent(later, e, 0));
  return MAG_OK;
}

int reserve_mark_bytes(mag_ctx* c)
{
  const int64_t nel = c->np + c->npy + c->nt + c->ntri;
  if (!c->d_edge_bytes && c->ne) MAG_CUDA(c, cudaMalloc((void**)&c->d_edge_bytes, (size_t)c->ne));
  if (!c->d_elem_bytes && nel) MAG_CUDA(c, cudaMalloc((void**)&c->d_elem_bytes, (size_t)nel));
  return MAG_OK;
}
// every stream of the pipeline is idle when an entry point returns, also on the error paths: the caller may free or
// unpin its host buffers right away
int drain(mag_ctx* c, int rc)
{
  cudaError_t e1 = c->s_up ? cudaStreamSynchronize(c->s_up) : cudaSuccess;
  cudaError_t e2 = cudaStreamSynchronize(c->stream);
  cudaError_t e3 = c->s_down ? cudaStreamSynchronize(c->s_down) : cudaSuccess;
  if (rc) return rc;
  cudaError_t e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
  if (e != cudaSuccess) return mag_fail(c, MAG_ERR_CUDA, "pipeline drain: %s", cudaGetErrorString(e));
  return MAG_OK;
}

template <class T>
int copy_async(mag_ctx* c, T* dst, const T* src, size_t count, cudaMemcpyKind kind, cudaStream_t s)
{
  if (count) MAG_CUDA(c, cudaMemcpyAsync(dst, src, count * sizeof(T), kind, s));
  return MAG_OK;
}

} // namespace

extern "C" int mag_sweep_host(mag_ctx* c, const mag_host_part* in, const mag_host_result* out, uint32_t ops, double max_len,
                              double min_len, double good_quality, int use_max_metric, int fp_mode, mag_stats* stats)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  if (!in || !out) return mag_fail(c, MAG_ERR_ARG, "mag_sweep_host: null argument");
  if (fp_mode != MAG_FP_STRICT && fp_mode != MAG_FP_FAST) return mag_fail(c, MAG_ERR_ARG, "mag_sweep_host: bad fp_mode %d", fp_mode);
  if (ops & ~(uint32_t)((MAG_OP_ALL & ~MAG_OP_LAYER_CHECK) | MAG_OP_LENGTH_SUM))
    return mag_fail(c, MAG_ERR_ARG, "mag_sweep_host: op bits 0x%x not supported here (tet parts; use the resident API for layer checks)", ops);
  if ((ops & MAG_OP_LENGTH_SUM) && !(ops & MAG_OP_LENGTHS)) return mag_fail(c, MAG_ERR_ARG, "mag_sweep_host: MAG_OP_LENGTH_SUM needs MAG_OP_LENGTHS");
  const int64_t nv = in->nv, ne = in->ne, nt = in->nt;
  if ((nv > 0 && !in->xyz) || (ne > 0 && !in->edge_v) || (nt > 0 && !in->tet_v))
    return mag_fail(c, MAG_ERR_ARG, "mag_sweep_host: null array with non-zero count");
  if (nv == 0 && (ne > 0 || nt > 0)) return mag_fail(c, MAG_ERR_ARG, "mag_sweep_host: entities without vertices");
  size_t na = 0, nb = 0;
  switch (in->kind) {
    case MAG_KIND_IDENTITY: break;
    case MAG_KIND_ISO: na = (size_t)nv; break;
    case MAG_KIND_ANISO: na = (size_t)nv * 3; nb = (size_t)nv * 9; break;
    case MAG_KIND_LOGM: nb = (size_t)nv * 9; break;
    default: return mag_fail(c, MAG_ERR_ARG, "mag_sweep_host: bad size-field kind %d", in->kind);
  }
  if ((na && !in->field_a) || (nb && !in->field_b)) return mag_fail(c, MAG_ERR_ARG, "mag_sweep_host: null size-field array");
  int rc;
  if (!c->s_up) MAG_CUDA(c, cudaStreamCreateWithFlags(&c->s_up, cudaStreamNonBlocking));
  if (!c->s_down) MAG_CUDA(c, cudaStreamCreateWithFlags(&c->s_down, cudaStreamNonBlocking));
  cudaStream_t s_cmp = c->stream, s_up = c->s_up, s_down = c->s_down;
  Events ev{c};

  if ((rc = magi_reshape(c, 3, nv, ne, nt, 0, 0, 0, in->edge_owned != nullptr, in->elem_owned != nullptr))) return rc;
  c->v2t_valid = false;
  if ((rc = magi_reserve_metric(c, in->kind, na, nb))) return rc;
  c->kind = in->kind;
  c->uniform_refiner = false;
  c->edge_flags_zero = c->elem_flags_zero = false;   // every slice below is uploaded or zeroed explicitly
  c->tet_words_zero = false;
  c->last_ops = ops;
  c->last_fp_mode = fp_mode;
  // everything queued from here on is drained before the call returns, also when a step fails (the caller may free or
  // unpin its buffers right after an error)
  auto run = [&]() -> int {
  int rc;
  // the device arrays may still be read by work queued earlier on the compute stream
  if ((rc = chain(c, ev, s_cmp, s_up))) return rc;

  // ---- vertex data, then pack + per-vertex transforms
  if ((rc = copy_async(c, c->d_xyz, in->xyz, (size_t)nv * 3, cudaMemcpyHostToDevice, s_up)) ||
      (rc = copy_async(c, c->d_ma, in->field_a, na, cudaMemcpyHostToDevice, s_up)) ||
      (rc = copy_async(c, c->d_mb, in->field_b, nb, cudaMemcpyHostToDevice, s_up)))
    return rc;
  if ((rc = chain(c, ev, s_up, s_cmp))) return rc;
  if (nv && (rc = magk_pack(c))) return rc;
  c->vertex_pass_valid = false;
  if (nv && (ops & (MAG_OP_QUALITIES | MAG_OP_MARK_BAD)) && (rc = magk_vertex_pass(c))) return rc;
  if ((rc = magk_init_stats(c))) return rc;   // after the vertex pass: its eigen-solver failures are folded in

  const bool do_edges = ne && (ops & (MAG_OP_LENGTHS | MAG_OP_MARK_SPLIT | MAG_OP_MARK_COLLAPSE));
  int64_t slice = in->slice_entities > 0 ? in->slice_entities : (int64_t)4 << 20;
  slice = std::max<int64_t>(61440, (slice + 61439) / 61440 * 61440); // a whole number of edge chunks (8192) and tet chunks (7680)

  // ---- edges
  for (int64_t e0 = 0; e0 < ne; e0 += slice) {
    const int64_t n = std::min(slice, ne - e0);
    if ((rc = copy_async(c, c->d_edge_v + 2 * e0, in->edge_v + 2 * e0, (size_t)n * 2, cudaMemcpyHostToDevice, s_up))) return rc;
    if (in->edge_flags && (rc = copy_async(c, c->d_edge_flags + e0, in->edge_flags + e0, (size_t)n, cudaMemcpyHostToDevice, s_up))) return rc;
    if (in->edge_owned && (rc = copy_async(c, c->d_edge_owned + e0, in->edge_owned + e0, (size_t)n, cudaMemcpyHostToDevice, s_up))) return rc;
    if ((rc = chain(c, ev, s_up, s_cmp))) return rc;
    if (!in->edge_flags) MAG_CUDA(c, cudaMemsetAsync(c->d_edge_flags + e0, 0, (size_t)n * 4, s_cmp));
    if ((rc = magk_check_conn(c, c->d_edge_v + 2 * e0, 2 * n))) return rc;   // out-of-range ids: counted, made harmless, reported with the statistics
    if ((rc = magk_edges_range(c, ops, max_len, min_len, fp_mode, e0, n))) return rc;
    if ((rc = chain(c, ev, s_cmp, s_down))) return rc;
    if (out->edge_lengths && (ops & MAG_OP_LENGTHS) &&
        (rc = copy_async(c, out->edge_lengths + e0, c->d_len + e0, (size_t)n, cudaMemcpyDeviceToHost, s_down)))
      return rc;
    if (out->edge_flags && (rc = copy_async(c, out->edge_flags + e0, c->d_edge_flags + e0, (size_t)n, cudaMemcpyDeviceToHost, s_down))) return rc;
  }
  if (do_edges && (ops & MAG_OP_LENGTH_SUM) && (rc = magk_length_sum(c))) return rc;

  // ---- tets
  for (int64_t t0 = 0; t0 < nt; t0 += slice) {
    const int64_t n = std::min(slice, nt - t0);
    if ((rc = copy_async(c, c->d_tet_v + 4 * t0, in->tet_v + 4 * t0, (size_t)n * 4, cudaMemcpyHostToDevice, s_up))) return rc;
    if (in->elem_flags && (rc = copy_async(c, c->d_elem_flags + t0, in->elem_flags + t0, (size_t)n, cudaMemcpyHostToDevice, s_up))) return rc;
    if (in->elem_owned && (rc = copy_async(c, c->d_elem_owned + t0, in->elem_owned + t0, (size_t)n, cudaMemcpyHostToDevice, s_up))) return rc;
    if ((rc = chain(c, ev, s_up, s_cmp))) return rc;
    if (!in->elem_flags) MAG_CUDA(c, cudaMemsetAsync(c->d_elem_flags + t0, 0, (size_t)n * 4, s_cmp));
    if ((rc = magk_check_conn(c, c->d_tet_v + 4 * t0, 4 * n))) return rc;
    if ((rc = magk_tets_range(c, ops, good_quality, use_max_metric, fp_mode, t0, n))) return rc;
    if ((rc = chain(c, ev, s_cmp, s_down))) return rc;
    if (out->qualities && (ops & MAG_OP_QUALITIES) &&
        (rc = copy_async(c, out->qualities + t0, c->d_qual + t0, (size_t)n, cudaMemcpyDeviceToHost, s_down)))
      return rc;
    if (out->elem_flags && (rc = copy_async(c, out->elem_flags + t0, c->d_elem_flags + t0, (size_t)n, cudaMemcpyDeviceToHost, s_down))) return rc;
  }

  // ---- statistics; everything has landed when the three streams are idle
  mag_stats local;
  return mag_get_stats(c, stats ? stats : &local);      // synchronizes the compute stream
  };
  return drain(c, run());
}


extern "C" int mag_set_mark_bytes(mag_ctx* c, const uint8_t* edge_marks, const uint8_t* elem_marks)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  const int64_t nel = c->np + c->npy + c->nt + c->ntri;
  int rc;
  if ((rc = reserve_mark_bytes(c))) return rc;
  if (edge_marks && c->ne) {
    MAG_CUDA(c, cudaMemcpyAsync(c->d_edge_bytes, edge_marks, (size_t)c->ne, cudaMemcpyHostToDevice, c->stream));
    if ((rc = magk_marks_expand(c, c->stream, c->d_edge_bytes, c->d_edge_flags, c->ne))) return rc;
  }
  c->edge_flags_zero = edge_marks == nullptr;
  if (elem_marks && nel) {
    MAG_CUDA(c, cudaMemcpyAsync(c->d_elem_bytes, elem_marks, (size_t)nel, cudaMemcpyHostToDevice, c->stream));
    if ((rc = magk_marks_expand(c, c->stream, c->d_elem_bytes, c->d_elem_flags, nel))) return rc;
  }
  c->elem_flags_zero = elem_marks == nullptr;
  c->tet_words_zero = elem_marks == nullptr;
  return MAG_OK;
}

extern "C" int mag_get_mark_bytes(mag_ctx* c, uint8_t* edge_marks, uint8_t* elem_marks)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  const int64_t nel = c->np + c->npy + c->nt + c->ntri;
  int rc;
  if ((rc = reserve_mark_bytes(c)) || (rc = magi_materialize_flags(c))) return rc;
  if (edge_marks && c->ne) {
    if ((rc = magk_marks_compress(c, c->stream, c->d_edge_flags, c->d_edge_bytes, c->ne))) return rc;
    MAG_CUDA(c, cudaMemcpyAsync(edge_marks, c->d_edge_bytes, (size_t)c->ne, cudaMemcpyDeviceToHost, c->stream));
  }
  if (elem_marks && nel) {
    if ((rc = magk_marks_compress(c, c->stream, c->d_elem_flags, c->d_elem_bytes, nel))) return rc;
    MAG_CUDA(c, cudaMemcpyAsync(elem_marks, c->d_elem_bytes, (size_t)nel, cudaMemcpyDeviceToHost, c->stream));
  }
  MAG_CUDA(c, cudaStreamSynchronize(c->stream));
  return MAG_OK;
}

// One sweep of a part whose CONNECTIVITY is already resident (mag_set_mesh / mag_sweep_host earlier): what changes between the
// sweeps of one MeshAdapt iteration, or between the time steps of a solver that re-evaluates its size field on a fixed
// mesh, is the vertex data and the flag words -- 120 B / vertex + 1 B / entity up and 1 B / entity down instead of the
// 2.7 GB + 1.3 GB of the full export (n = 203 part: 1.13 GB + 0.11 GB).
//     upload stream     xyz | size field | mark bytes
//     compute stream                      pack, vertex pass | expand | edge sweep | compress | element sweep | compress
//     download stream                                                               edge bytes (+ lengths) | element bytes (+ qualities)
extern "C" int mag_resweep_host(mag_ctx* c, const mag_host_update* in, const mag_host_marks* out, uint32_t ops, double max_len,
                                double min_len, double good_quality, int use_max_metric, int fp_mode, mag_stats* stats)
{
  if (!c) return MAG_ERR_ARG;
  MAG_CUDA(c, cudaSetDevice(c->device));
  if (!in || !out) return mag_fail(c, MAG_ERR_ARG, "mag_resweep_host: null argument");
  if (fp_mode != MAG_FP_STRICT && fp_mode != MAG_FP_FAST) return mag_fail(c, MAG_ERR_ARG, "mag_resweep_host: bad fp_mode %d", fp_mode);
  if (ops & ~(uint32_t)(MAG_OP_ALL | MAG_OP_LENGTH_SUM)) return mag_fail(c, MAG_ERR_ARG, "mag_resweep_host: unknown op bits 0x%x", ops);
  if ((ops & MAG_OP_LENGTH_SUM) && !(ops & MAG_OP_LENGTHS)) return mag_fail(c, MAG_ERR_ARG, "mag_resweep_host: MAG_OP_LENGTH_SUM needs MAG_OP_LENGTHS");
  if (c->nv == 0 || !c->d_xyz) return mag_fail(c, MAG_ERR_ARG, "mag_resweep_host: no part is resident (mag_set_mesh / mag_sweep_host first)");
  const int64_t nv = c->nv, ne = c->ne, nel = c->np + c->npy + c->nt + c->ntri;
  size_t na = 0, nb = 0;
  if (in->kind >= 0) {
    switch (in->kind) {
      case MAG_KIND_IDENTITY: break;
      case MAG_KIND_ISO: na = (size_t)nv; break;
      case MAG_KIND_ANISO: na = (size_t)nv * 3; nb = (size_t)nv * 9; break;
      case MAG_KIND_LOGM: nb = (size_t)nv * 9; break;
      default: return mag_fail(c, MAG_ERR_ARG, "mag_resweep_host: bad size-field kind %d", in->kind);
    }
    if ((na && !in->field_a) || (nb && !in->field_b)) return mag_fail(c, MAG_ERR_ARG, "mag_resweep_host: null size-field array");
  } else if (c->kind == MAG_KIND_NONE) return mag_fail(c, MAG_ERR_ARG, "mag_resweep_host: no size field is resident and none was given");
  int rc;
  if (!c->s_up) MAG_CUDA(c, cudaStreamCreateWithFlags(&c->s_up, cudaStreamNonBlocking));
  if (!c->s_down) MAG_CUDA(c, cudaStreamCreateWithFlags(&c->s_down, cudaStreamNonBlocking));
  cudaStream_t s_cmp = c->stream, s_up = c->s_up, s_down = c->s_down;
  Events ev{c};
  if ((rc = reserve_mark_bytes(c))) return rc;
  if (in->kind >= 0 && (rc = magi_reserve_metric(c, in->kind, na, nb))) return rc;
  if (!c->schedule_valid) {   // the part came through mag_sweep_host, which defers the device layout of the whole-part sweeps
    if ((rc = magk_build_schedule(c))) return rc;
    c->schedule_valid = true;
  }
  c->last_ops = ops;
  c->last_fp_mode = fp_mode;
  auto run = [&]() -> int {
    int rc;
    if ((rc = chain(c, ev, s_cmp, s_up))) return rc;   // earlier work on the compute stream may still read these arrays
    const bool new_vertex_data = in->xyz || in->kind >= 0;
    if (in->xyz && (rc = copy_async(c, c->d_xyz, in->xyz, (size_t)nv * 3, cudaMemcpyHostToDevice, s_up))) return rc;
    if (in->kind >= 0) {
      if ((rc = copy_async(c, c->d_ma, in->field_a, na, cudaMemcpyHostToDevice, s_up)) ||
          (rc = copy_async(c, c->d_mb, in->field_b, nb, cudaMemcpyHostToDevice, s_up)))
        return rc;
      c->kind = in->kind;
      c->uniform_refiner = false;
    }
    if ((rc = chain(c, ev, s_up, s_cmp))) return rc;
    if (in->edge_marks && (rc = copy_async(c, c->d_edge_bytes, in->edge_marks, (size_t)ne, cudaMemcpyHostToDevice, s_up))) return rc;
    if (in->elem_marks && (rc = copy_async(c, c->d_elem_bytes, in->elem_marks, (size_t)nel, cudaMemcpyHostToDevice, s_up))) return rc;
    if (new_vertex_data) {
      c->vertex_pass_valid = false;
      if ((rc = magk_pack(c))) return rc;
      if ((ops & (MAG_OP_QUALITIES | MAG_OP_MARK_BAD | MAG_OP_LAYER_CHECK)) && (rc = magk_vertex_pass(c))) return rc;
    }
    if ((rc = chain(c, ev, s_up, s_cmp))) return rc;
    if (in->edge_marks && (rc = magk_marks_expand(c, s_cmp, c->d_edge_bytes, c->d_edge_flags, ne))) return rc;
    if (in->elem_marks && (rc = magk_marks_expand(c, s_cmp, c->d_elem_bytes, c->d_elem_flags, nel))) return rc;
    c->edge_flags_zero = in->edge_marks == nullptr;
    c->elem_flags_zero = in->elem_marks == nullptr;
    c->tet_words_zero = in->elem_marks == nullptr;
    if ((rc = magk_init_stats(c))) return rc;
    const uint32_t edge_ops = ops & (MAG_OP_LENGTHS | MAG_OP_MARK_SPLIT | MAG_OP_MARK_COLLAPSE | MAG_OP_LENGTH_SUM);
    const uint32_t elem_ops = ops & (MAG_OP_QUALITIES | MAG_OP_MARK_BAD | MAG_OP_LAYER_CHECK);
    // edges first: their bytes (and lengths) travel back while the elements are evaluated
    if (edge_ops && (rc = magk_sweep(c, edge_ops, max_len, min_len, good_quality, use_max_metric, fp_mode))) return rc;
    if (out->edge_marks && ne) {
      if ((rc = magi_materialize_flags(c))) return rc;
      if ((rc = magk_marks_compress(c, s_cmp, c->d_edge_flags, c->d_edge_bytes, ne))) return rc;
    }
    if ((rc = chain(c, ev, s_cmp, s_down))) return rc;
    if (out->edge_marks && (rc = copy_async(c, out->edge_marks, c->d_edge_bytes, (size_t)ne, cudaMemcpyDeviceToHost, s_down))) return rc;
    if (out->edge_lengths && (ops & MAG_OP_LENGTHS) && (rc = copy_async(c, out->edge_lengths, c->d_len, (size_t)ne, cudaMemcpyDeviceToHost, s_down))) return rc;
    if (elem_ops && (rc = magk_sweep(c, elem_ops, max_len, min_len, good_quality, use_max_metric, fp_mode))) return rc;
    if (out->elem_marks && nel) {
      if ((rc = magi_materialize_flags(c))) return rc;
      if ((rc = magk_marks_compress(c, s_cmp, c->d_elem_flags, c->d_elem_bytes, nel))) return rc;
    }
    if ((rc = chain(c, ev, s_cmp, s_down))) return rc;
    if (out->elem_marks && (rc = copy_async(c, out->elem_marks, c->d_elem_bytes, (size_t)nel, cudaMemcpyDeviceToHost, s_down))) return rc;
    if (out->qualities && (ops & MAG_OP_QUALITIES) && (rc = copy_async(c, out->qualities, c->d_qual, (size_t)nel, cudaMemcpyDeviceToHost, s_down))) return rc;
    mag_stats local;
    return mag_get_stats(c, stats ? stats : &local);      // synchronizes the compute stream
  };
  return drain(c, run());
}

// ma::getLinearQualitiesInMetricSpace's selection and post-processing (ma/maStats.cc:12-31) on the host, with the host's
// libm exactly as the reference: owned simplex elements in iteration order, cbrt of the quality (2-D: signed sqrt).
extern "C" int mag_linear_qualities(int dim, int64_t n, const double* qualities, const uint8_t* keep, double* out, int64_t* n_out)
{
  if ((dim != 2 && dim != 3) || n < 0 || (n && (!qualities || !out)) || !n_out) return MAG_ERR_ARG;
  int64_t k = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (keep && !keep[i]) continue;
    const double lq = qualities[i];
    out[k++] = dim == 2 ? ((lq > 0) ? std::sqrt(lq) : -std::sqrt(-lq)) : cbrt(lq);
  }
  *n_out = k;
  return MAG_OK;
}
