/* mag.h -- C ABI of the B200-native MeshAdapt marking / quality sweep.
 *
 * "mag" = MeshAdapt on GPU.  One opaque context per (CUDA device, mesh part).
 * The library replaces, for whole-mesh sweeps only, the following SCOREC/core
 * entry points (paths relative to the reference tree):
 *
 *   ma::SizeField::measure / shouldSplit / shouldCollapse   ma/maSize.h:30-54, ma/maSize.cc:208-224
 *   AnisoSizeField / LogAnisoSizeField / IsoSizeField::getTransform   ma/maSize.cc:395-413, 506-522, 581-616
 *   ma::measureElementQuality (tets, mean ratio cubed)      ma/maShape.h:24-26, ma/maQuality.cc:139-182
 *   ma::measureTriQuality (2-D meshes)                      ma/maQuality.cc:110-136
 *   ma::markEdgesToSplit                                    ma/maRefine.h:52,  ma/maRefine.cc:395-400
 *   ma::markEdgesToCollapse                                 ma/maCoarsen.cc:287-292
 *   ma::markBadQuality / getMinQuality                      ma/maShape.cc:132-169
 *   ma::markEntities (skip / set-flag / owned-count rules)  ma/maAdapt.cc:293-324
 *   ma::getMaximumEdgeLength                                ma/maSize.cc:673-691
 *   ma::isPrismOk / isPyramidOk (layer element validity)    ma/maQuality.cc:490-560
 *   ma::checkFlagConsistency / syncFlag (part boundaries)   ma/maAdapt.cc:226-256, 498-520
 *   PCU Add<long> / Min<double> / Max<double> on this path  ma/maAdapt.cc:323, ma/maShape.cc:168, ma/maSize.cc:689
 *
 * Conventions: every call returns 0 on success or a MAG_ERR_* code and never
 * throws or aborts across the ABI; mag_last_error() gives the message.  The
 * caller owns every host buffer; the library copies what it needs into
 * device-resident arrays.  One host thread per context.  Plain pointers and
 * sizes only.  There is NO CPU fallback: without a CUDA device mag_create fails.
 *
 * Entity order is the caller's (the adapter uses the reference's m->begin(d)
 * iteration order).  Vertex ids are 0-based int32.  Dimension-3 entities are
 * ordered prisms, pyramids, tets, as MDS iterates them (mds/mds.h:16-26 type
 * order WEDGE, PYRAMID, TET); "element" arrays (flags, owned, qualities) are the
 * concatenation [prisms | pyramids | tets].
 */
#ifndef MAG_H
#define MAG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mag_ctx mag_ctx;

/* error codes */
enum {
  MAG_OK = 0,
  MAG_ERR_CUDA = 1,        /* CUDA runtime error (message has the call site) */
  MAG_ERR_ARG = 2,         /* bad argument / call order */
  MAG_ERR_FLAG_STATE = 3,  /* an entity already carried the true-flag on entry: the reference asserts (maAdapt.cc:308) */
  MAG_ERR_EIGEN = 4,       /* eigenQR did not converge in 100 iterations / zero Wilkinson denominator: the reference asserts (apfMatrix.cc:76, mthQR.cc:228) */
  MAG_ERR_NONSIMPLEX = 5,  /* a prism/pyramid reached markBadQuality without OK_QUALITY: the reference dereferences a null table entry (maQuality.cc:169-182) */
  MAG_ERR_NCCL = 6,
  MAG_ERR_INCONSISTENT = 7 /* part-boundary copies disagree on a flag: the reference asserts (maRefine.cc:430, maCoarsen.cc:305) */
};

/* ma flag bits, identical to ma/maAdapt.h:17-37 */
enum {
  MAG_SPLIT = 1 << 0, MAG_DONT_SPLIT = 1 << 1, MAG_COLLAPSE = 1 << 2, MAG_DONT_COLLAPSE = 1 << 3,
  MAG_CHECKED = 1 << 4, MAG_BAD_QUALITY = 1 << 5, MAG_OK_QUALITY = 1 << 6, MAG_SNAP = 1 << 7,
  MAG_DONT_SNAP = 1 << 8, MAG_DONT_SWAP = 1 << 9, MAG_LAYER = 1 << 10, MAG_LAYER_BASE = 1 << 11,
  MAG_LAYER_TOP = 1 << 12, MAG_DIAGONAL_1 = 1 << 13, MAG_DIAGONAL_2 = 1 << 14, MAG_LAYER_UNSNAP = 1 << 15,
  MAG_DONT_MOVE = 1 << 16, MAG_NEED_NOT_SPLIT = 1 << 17, MAG_NEED_NOT_COLLAPSE = 1 << 18
};

/* what one mag_sweep call does (bit mask) */
enum {
  MAG_OP_LENGTHS = 1 << 0,       /* measure() of EVERY edge -> mag_get_edge_lengths (ma::getEdgeLengthsInMetricSpace, maStats.cc:33-45) */
  MAG_OP_MARK_SPLIT = 1 << 1,    /* ma::markEdgesToSplit */
  MAG_OP_MARK_COLLAPSE = 1 << 2, /* ma::markEdgesToCollapse */
  MAG_OP_QUALITIES = 1 << 3,     /* measureElementQuality of EVERY tet + min (ma::getMinQuality, maStats.cc:12-31 before the cbrt) */
  MAG_OP_MARK_BAD = 1 << 4,      /* ma::markBadQuality */
  MAG_OP_LAYER_CHECK = 1 << 5,   /* ma::isPrismOk / isPyramidOk on every prism / pyramid (ma::checkLayerShape, maLayer.cc:166-192) */
  MAG_OP_ALL = 63,
  MAG_OP_LENGTH_SUM = 1 << 6     /* additionally reduce the sum of the measured lengths over EVERY edge of the part (owned or not, as
                                    ma::getAverageEdgeLength does, ma/maSize.cc:654-671) into mag_stats.sum_length; the average is
                                    sum_length / ne, or the two summed over the parts (needs MAG_OP_LENGTHS; not part of MAG_OP_ALL:
                                    no reference mark needs it) */
};

/* arithmetic mode */
enum {
  MAG_FP_STRICT = 0, /* the reference's operation order, no FMA contraction, IEEE div/sqrt: bit-identical lengths, qualities, flags
                        (LogAniso: exp() is CUDA's, within 1 ulp of glibc -> values within 1e-12, flags identical outside the listed band) */
  MAG_FP_FAST = 1,   /* algebraically equivalent, FMA-contracted evaluation (values within 1e-12 relative); every entity whose value
                        lands within 1e-12 relative of a threshold is re-evaluated in strict arithmetic ON THE DEVICE and listed,
                        so flags and counts stay identical to MAG_FP_STRICT */
  MAG_FP_FAST_LISTED = 2 /* mag_sweep / mag_sweep_reconciled only: MAG_FP_FAST without the re-evaluation of near-threshold EDGES -- an
                        edge within 1e-12 relative of MAXLENGTH / MINLENGTH is decided by the fast value and listed
                        (mag_get_near_threshold), which is exactly the exception the parity rule of this path allows ("edges whose
                        reference metric length lies within 1e-12 relative of the threshold ... must be listed").  Same lengths and
                        qualities as MAG_FP_FAST bit for bit, same flags outside the list; counts may differ from the reference's
                        by at most the length of the list.  Worth it where whole edge families sit ON a threshold (a lattice whose
                        size field is a multiple of the spacing): the lattice benchmark's 8.4 M z edges cost 0.39 ms of strict
                        re-evaluation per sweep, 2.0 ms in a LogAniso field.  Elements are re-evaluated as in MAG_FP_FAST. */
};

typedef struct mag_stats {
  int64_t n_split;          /* owned edges newly marked SPLIT      (return value of ma::markEdgesToSplit) */
  int64_t n_collapse;       /* owned edges newly marked COLLAPSE   (ma::markEdgesToCollapse) */
  int64_t n_bad;            /* owned elements newly marked BAD_QUALITY (ma::markBadQuality) */
  int64_t n_edges_evaluated;/* edges whose predicate was evaluated (not skipped by DONT_ / NEED_NOT_ flags) */
  int64_t n_elems_evaluated;
  int64_t n_near_threshold; /* entities within 1e-12 relative of a threshold (see mag_get_near_threshold) */
  int64_t n_layer_unsafe;   /* prisms / pyramids failing isPrismOk / isPyramidOk (ma::checkLayerShape count) */
  int64_t n_flag_mismatch;  /* part-boundary copies that disagreed before owner reconciliation (must be 0) */
  double min_quality;       /* ma::getMinQuality: min over all tets, initial value 1.0 (valid after MAG_OP_QUALITIES) */
  double max_length;        /* ma::getMaximumEdgeLength: max over owned edges, initial 0.0 (valid after MAG_OP_LENGTHS) */
  double sum_length;        /* MAG_OP_LENGTH_SUM: sum over all edges of the part (fixed tree order, not the reference's serial order:
                               equal to 1e-13 relative, not bit for bit); else 0 */
} mag_stats;

/* ---- lifetime ---- */
int mag_create(mag_ctx** out, int device);
void mag_destroy(mag_ctx* c);
const char* mag_last_error(const mag_ctx* c); /* c may be NULL: last error of a failed mag_create */
/* adopt a caller-owned cudaStream_t (e.g. the caller's timing stream); NULL restores the context's own stream */
int mag_set_stream(mag_ctx* c, void* cuda_stream);
int mag_synchronize(mag_ctx* c);

/* ---- one-time export of a part (replaces per-entity apf::Mesh2 access: apfMDS.cc:309-318, 347-351, 288-298) ----
   xyz [nv][3]; edge_v [ne][2]; tet_v [nt][4]; prism_v [np][6]; pyr_v [npy][5] (any count may be 0, pointer NULL);
   edge_owned [ne] / elem_owned [np+npy+nt] bytes, NULL = every entity is owned (serial mesh). */
int mag_set_mesh(mag_ctx* c, int64_t nv, const double* xyz,
                 int64_t ne, const int32_t* edge_v,
                 int64_t nt, const int32_t* tet_v,
                 int64_t np, const int32_t* prism_v,
                 int64_t npy, const int32_t* pyr_v,
                 const uint8_t* edge_owned, const uint8_t* elem_owned);
/* a 2-D part: the elements are triangles (ma::measureTriQuality, ma/maQuality.cc:110-136; the max-"Jacobian" vertex is
   chosen by |row0 x row1| of Q_v, apf/apfVectorElement.cc:75-84).  tri_v [ntri][3]; element arrays (flags, owned, qualities)
   have ntri entries; good_quality defaults to 0.2 in 2-D (ma/maInput.cc:32-46, the caller passes it). */
int mag_set_mesh_2d(mag_ctx* c, int64_t nv, const double* xyz,
                    int64_t ne, const int32_t* edge_v,
                    int64_t ntri, const int32_t* tri_v,
                    const uint8_t* edge_owned, const uint8_t* elem_owned);
/* moved vertices, same connectivity (apf::Mesh2::setPoint) */
int mag_set_coords(mag_ctx* c, const double* xyz);

/* ---- size field (vertex nodes, linear Lagrange) ---- */
int mag_set_metric_identity(mag_ctx* c);                                   /* ma::IdentitySizeField */
int mag_set_metric_uniform_refiner(mag_ctx* c);                            /* ma::UniformRefiner (ma/maSize.h:75-85): identity measure,
                                                                              shouldSplit constant true, shouldCollapse false */
int mag_set_metric_iso(mag_ctx* c, const double* size /*[nv]*/);           /* IsoSizeField / IsoUserField */
int mag_set_metric_aniso(mag_ctx* c, const double* h /*[nv][3]*/,
                         const double* R /*[nv][9] row-major, frame vectors in columns*/); /* AnisoSizeField */
int mag_set_metric_logm(mag_ctx* c, const double* logM /*[nv][9] row-major*/);           /* LogAnisoSizeField's ma_logM field */

/* ---- incoming "ma_flags" words; NULL = all zero (maAdapt.cc:80-88 getFlags default) ---- */
int mag_set_flags(mag_ctx* c, const int32_t* edge_flags /*[ne]*/, const int32_t* elem_flags /*[np+npy+nt]*/);

/* ma::clearFlagFromDimension (ma/maAdapt.cc:139-147) on the resident "ma_flags" words: clears the bits of flag on every
   edge (dimension 1) or every element (dimension = the mesh dimension).  ma::unMarkBadQuality (ma/maShape.cc:138-150) is
   mag_clear_flag(c, 3, MAG_BAD_QUALITY).  Asynchronous on the context's stream. */
int mag_clear_flag(mag_ctx* c, int dimension, int32_t flag);

/* ma::resetLayer (ma/maLayer.cc:94-103) on the resident flag words, what ma::Adapt's constructor does before any mark:
   markLayerElements (:11-48) -- every prism / pyramid, and every element i with user_layer_tag[i] != 0
   (Input::userDefinedLayerTagName; [np+npy+nt] host array or NULL), puts MAG_LAYER on itself and on the edges of its
   closure; with part-boundary links set (mag_set_edge_links) LAYER is then OR-ed across the copies of shared edges
   (syncFlag, :45-46) -- and freezeLayer (:51-71): LAYER edges get DONT_COLLAPSE | DONT_SPLIT | DONT_SWAP, LAYER elements
   OK_QUALITY, so that the three marks skip the boundary layer.  Bits are OR-ed into the resident words (mag_set_flags /
   mag_clear_flag first for a fresh start).  *n_layer_elements (may be NULL) = this part's count (the reference sums it
   over the parts).  Vertex and face flag words are not held by the library.  Synchronous. */
int mag_reset_layer(mag_ctx* c, const int32_t* user_layer_tag, int64_t* n_layer_elements);

/* ---- the sweep.  max_len / min_len = ma::MAXLENGTH / MINLENGTH (1.5 / 0.5, maSize.h:26-27);
   good_quality = ma::Input::goodQuality; use_max_metric = measureElementQuality's useMax (default true, maShape.h:26).
   Asynchronous on the context's stream; results are read with the getters below (which synchronize). ---- */
int mag_sweep(mag_ctx* c, uint32_t ops, double max_len, double min_len, double good_quality,
              int use_max_metric, int fp_mode);

/* ---- results ---- */
int mag_get_edge_lengths(mag_ctx* c, double* out /*[ne]*/);
int mag_get_qualities(mag_ctx* c, double* out /*[np+npy+nt], prisms/pyramids = 0*/);
int mag_get_flags(mag_ctx* c, int32_t* edge_flags, int32_t* elem_flags); /* either may be NULL */
int mag_get_layer_ok(mag_ctx* c, int32_t* ok /*[np+npy]*/, int32_t* codes /*[np+npy]*/);
int mag_get_stats(mag_ctx* c, mag_stats* out); /* this part only; also reports deferred MAG_ERR_FLAG_STATE / EIGEN / NONSIMPLEX */
/* entities whose value lay within 1e-12 relative of a threshold in the last sweep.
   which: 0 = edges (vs max_len or min_len), 1 = elements (vs good_quality).  Writes up to cap indices, returns the total in *n. */
int mag_get_near_threshold(mag_ctx* c, int which, int64_t* idx, int64_t cap, int64_t* n);

/* ---- export + sweep + results of one tet part in ONE call, streamed (what the adapter does once per MeshAdapt
   iteration, ma/maAdapt.cc:293-324 callers in maRefine.cc / maCoarsen.cc / maShape.cc).  Equivalent to
   mag_set_mesh + mag_set_metric_* + mag_set_flags + mag_sweep + mag_get_edge_lengths / qualities / flags / stats, with
   bit-identical results, but the edges and tets travel in slices so that the device->host copies of finished slices
   overlap the host->device copies of the next ones; the call returns when every output has landed.  Host buffers should
   be page-locked (cudaHostAlloc / cudaHostRegister) for the copies to be asynchronous.  The part stays resident: the
   getters, mag_sweep, mag_reconcile_edge_flags ... work on it afterwards.  Prisms / pyramids are not supported here. */
typedef struct mag_host_part {
  int64_t nv; const double* xyz;           /* [nv][3] */
  int64_t ne; const int32_t* edge_v;       /* [ne][2] */
  int64_t nt; const int32_t* tet_v;        /* [nt][4] */
  const uint8_t* edge_owned;               /* [ne] or NULL */
  const uint8_t* elem_owned;               /* [nt] or NULL */
  int kind;                                /* 0 identity, 1 iso, 2 aniso, 3 logm */
  const double* field_a;                   /* iso: size[nv]; aniso: h[nv][3]; else NULL */
  const double* field_b;                   /* aniso: R[nv][9]; logm: logM[nv][9]; else NULL */
  const int32_t* edge_flags;               /* incoming ma_flags words, NULL = 0 */
  const int32_t* elem_flags;
  int64_t slice_entities;                  /* entities per slice, 0 = default (4 Mi); rounded up to whole work chunks */
} mag_host_part;
typedef struct mag_host_result {           /* any pointer may be NULL */
  double* edge_lengths;                    /* [ne]  (MAG_OP_LENGTHS) */
  double* qualities;                       /* [nt]  (MAG_OP_QUALITIES) */
  int32_t* edge_flags;                     /* [ne] */
  int32_t* elem_flags;                     /* [nt] */
} mag_host_result;
int mag_sweep_host(mag_ctx* c, const mag_host_part* in, const mag_host_result* out, uint32_t ops, double max_len,
                   double min_len, double good_quality, int use_max_metric, int fp_mode, mag_stats* stats /* may be NULL */);

/* ---- the same for a part whose connectivity is ALREADY resident: only what changes between the sweeps of one MeshAdapt
   iteration (or between the time steps of a solver that re-evaluates its size field on a fixed mesh) crosses the host
   link -- vertex coordinates, the size field, and the flag words as one MARK BYTE per entity:
     bit 0 SPLIT, 1 DONT_SPLIT, 2 COLLAPSE, 3 DONT_COLLAPSE (= bits 0-3 of the ma_flags word, ma/maAdapt.h:17-37),
     bit 4 NEED_NOT_SPLIT, 5 NEED_NOT_COLLAPSE, 6 BAD_QUALITY, 7 OK_QUALITY
   i.e. exactly the bits markEntities reads (its allFalseFlags / the true-flag assertion) or writes for the three marks
   (ma/maAdapt.cc:293-324, maRefine.cc:385-400, maCoarsen.cc:277-292, maShape.cc:122-136); the caller keeps the other bits
   of its words and merges:  word = (word & ~MAG_MARK_WORD_MASK) | MAG_MARK_TO_WORD(byte). ---- */
enum {
  MAG_MARK_SPLIT = 1, MAG_MARK_DONT_SPLIT = 2, MAG_MARK_COLLAPSE = 4, MAG_MARK_DONT_COLLAPSE = 8,
  MAG_MARK_NEED_NOT_SPLIT = 16, MAG_MARK_NEED_NOT_COLLAPSE = 32, MAG_MARK_BAD_QUALITY = 64, MAG_MARK_OK_QUALITY = 128
};
#define MAG_MARK_WORD_MASK (MAG_SPLIT | MAG_DONT_SPLIT | MAG_COLLAPSE | MAG_DONT_COLLAPSE | MAG_NEED_NOT_SPLIT | MAG_NEED_NOT_COLLAPSE | MAG_BAD_QUALITY | MAG_OK_QUALITY)
#define MAG_MARK_TO_WORD(b) ((int32_t)(((uint32_t)(b) & 0xFu) | (((uint32_t)(b) & 0x30u) << 13) | (((uint32_t)(b) & 0xC0u) >> 1)))
#define MAG_WORD_TO_MARK(w) ((uint8_t)(((uint32_t)(w) & 0xFu) | (((uint32_t)(w) >> 13) & 0x30u) | (((uint32_t)(w) << 1) & 0xC0u)))
/* resident flag words := MAG_MARK_TO_WORD(byte) (all other bits zero); NULL = all zero, like mag_set_flags */
int mag_set_mark_bytes(mag_ctx* c, const uint8_t* edge_marks /*[ne]*/, const uint8_t* elem_marks /*[np+npy+nt]*/);
/* MAG_WORD_TO_MARK of the resident flag words; either pointer may be NULL.  Synchronous. */
int mag_get_mark_bytes(mag_ctx* c, uint8_t* edge_marks, uint8_t* elem_marks);
typedef struct mag_host_update {
  const double* xyz;                       /* [nv][3] moved vertices (apf::Mesh2::setPoint), NULL = unchanged */
  int kind;                                /* size field given here: 0 identity, 1 iso, 2 aniso, 3 logm; -1 = unchanged */
  const double* field_a;                   /* as in mag_host_part */
  const double* field_b;
  const uint8_t* edge_marks;               /* incoming mark bytes [ne], NULL = all zero */
  const uint8_t* elem_marks;               /* [np+npy+nt], NULL = all zero */
} mag_host_update;
typedef struct mag_host_marks {            /* any pointer may be NULL */
  uint8_t* edge_marks;                     /* [ne] outgoing mark bytes */
  uint8_t* elem_marks;                     /* [np+npy+nt] */
  double* edge_lengths;                    /* [ne], opt-in (MAG_OP_LENGTHS) */
  double* qualities;                       /* [np+npy+nt], opt-in (MAG_OP_QUALITIES) */
} mag_host_marks;
/* update + sweep + marks of the resident part in one streamed call (uploads | kernels | downloads on three streams; the
   edge bytes travel back while the elements are evaluated).  All of MAG_OP_ALL is supported (prisms / pyramids included).
   Results are bit-identical to mag_set_coords + mag_set_metric_* + mag_set_mark_bytes + mag_sweep + mag_get_mark_bytes.
   Returns when every output has landed and every stream is idle, also on errors. */
int mag_resweep_host(mag_ctx* c, const mag_host_update* in, const mag_host_marks* out, uint32_t ops, double max_len,
                     double min_len, double good_quality, int use_max_metric, int fp_mode, mag_stats* stats /* may be NULL */);

/* ---- the sweeps either side of the marking path, over the same resident part ----
   Predictive load-balance weight of every element, ma::getElementWeights (ma/maBalance.cc:83-97):
   SizeField::getWeight = measure(element) / parentMeasure (ma/maSize.cc:147-156,225-229; tets: 4-point Gauss rule
   apf/apfIntegrate.cc:328-342 with getTransform at every point), clamped to [w_min, w_max] as clampForIterations does
   (w_max = 2^(dim * refinesLeft), w_min = 4^(-coarsensLeft), maBalance.cc:41-52; pass +-HUGE_VAL for the raw weight).
   out [np+npy+nt] host (may be NULL: the weights stay on the device); prisms / pyramids get 0 (the reference weighs
   a prism by its base triangle, which needs that face's own vertex order).  On a 2-D part (mag_set_mesh_2d) the elements
   are triangles: 3-point rule apf/apfIntegrate.cc:146-159, dV = |row0(J Q) x row1(J Q)|, parent measure 1/2, out [ntri].
   Synchronous. */
int mag_element_weights(mag_ctx* c, double w_max, double w_min, int fp_mode, double* out);
/* The prisms of the resident part as ma::getElementWeight weighs them (ma/maBalance.cc:21-81): getSizeWeight measures a prism's
   BASE TRIANGLE (its first face, :31-37: SizeField::getWeight(face) = measure(triangle) / (1/2), 3-point rule as for a 2-D
   part), then clampForIterations to [w_min, w_max], clampForLayerPermissions (ma::Input::shouldRefineLayer / shouldCoarsenLayer:
   without the permission the weight is raised / lowered to 1) and accountForTets (shouldTurnLayerToTets: x 3).
   base_v [np][3] host: the first face of every prism in the FACE entity's own vertex order (getDownward(prism, 2)[0], then
   getDownward(face, 0)) -- the order the reference's integrator walks, which decides the last bit.  out [np] host (may be NULL);
   the values also replace the zeros mag_element_weights left for the prisms in the device copy.  Pyramids stay with the
   reference (sparse by construction, maBalance.cc:27-29).  Synchronous. */
int mag_prism_weights(mag_ctx* c, const int32_t* base_v /*[np][3]*/, double w_max, double w_min, int should_refine_layer,
                      int should_coarsen_layer, int should_turn_layer_to_tets, int fp_mode, double* out /*[np]*/);
/* Batch form of ma::getWorstQuality / hasWorseQuality (ma/maQuality.cc:184-226) for many candidate cavities at once
   (collapse / swap / snap operators: ma/maCollapse.cc:37-113, ma/maEdgeSwap.cc:598-740, ma/maSnapper.cc:407,573).
   Cavity k = the tets tet_v[offsets[k] .. offsets[k+1]) given by their four vertex ids: existing vertices of the resident
   part, but the tets need not exist in the mesh (operators evaluate would-be elements).  offsets [ncav+1], offsets[0] = 0,
   no empty cavity (the reference asserts n > 0).  worst [ncav] = smallest measureElementQuality of the cavity;
   qualities [offsets[ncav]] (may be NULL) = every element's quality; hasWorseQuality(cavity, q) == worst < q.
   The per-vertex transforms of the last sweep are reused while mesh and size field are unchanged.  Synchronous. */
int mag_cavity_quality(mag_ctx* c, int64_t ncav, const int64_t* offsets, const int32_t* tet_v /*[.][4]*/,
                       int use_max_metric, int fp_mode, double* worst, double* qualities);
/* The quality test of MANY edge-collapse candidates at once, from the resident part alone -- the consumer of the cavity batch
   inside ma::coarsen (ma/maCoarsen.cc:163-169 picks an independent set of vertices, then every member runs
   ma::Collapse::tryBothDirections, ma/maCollapse.cc:88-113).  Candidate k collapses one end of edge edges[k] (index in the
   resident edge array) onto the other: which_end[k] = 0 collapses the edge's first vertex, 1 its second (Collapse::vertToCollapse).
   As Collapse::computeElementSets / rebuildElements do (:353-383): tets around that vertex that contain the edge vanish, every
   other tet around it is rebuilt with the vertex replaced by the one that is kept (vertex order unchanged).
     new_worst [ncand] = ma::getWorstQuality of the rebuilt tets (+inf if there are none; the reference asserts there are)
     old_worst [ncand] = Collapse::getOldQuality(): the worst of ALL tets around the collapsing vertex (:425-433)
     n_keep    [ncand] = how many tets are rebuilt (may be NULL)
   The collapse passes the reference's test iff !(new_worst < min(goodQuality, max(old_worst, validQuality))) (:93-99 with
   ma::hasWorseQuality, maQuality.cc:200-226).  Strict arithmetic: bit-identical to what the reference computes on the mesh it
   really rebuilds.  Tets only: candidates next to layer elements are the caller's to exclude (their edges carry DONT_COLLAPSE,
   ma/maLayer.cc:75-103).  The vertex -> tet incidence is built on the device on first use after a mag_set_mesh and kept. */
int mag_collapse_quality(mag_ctx* c, int64_t ncand, const int32_t* edges, const uint8_t* which_end, int use_max_metric, int fp_mode,
                         double* new_worst, double* old_worst, int32_t* n_keep);
/* ShortEdgeFixer::shouldApply (ma/maShape.cc:188-219), the classification sweep ma::fixElementShapes runs over the
   BAD_QUALITY elements right after markBadQuality: tet_edges [nt][6] = the edge indices of every tet in
   getDownward(tet, 1) order; the lengths are those of the last MAG_OP_LENGTHS sweep (resident).  For every tet carrying
   BAD_QUALITY: if max / min edge length < max_edge_ratio (ma::Input::maximumEdgeRatio) the resident BAD_QUALITY bit is
   cleared (not a short-edge case), otherwise short_edge[element] = the first shortest edge (the one ShortEdgeRemover is
   given).  short_edge [np+npy+nt] (may be NULL), -1 where nothing is to be removed.  Synchronous. */
int mag_short_edge_test(mag_ctx* c, const int32_t* tet_edges, double max_edge_ratio, int32_t* short_edge,
                        int64_t* n_cleared, int64_t* n_short);
/* ma::getSliverCode / ma::matchSliver (ma/maShape.cc:35-120), the classification LargeAngleTetFixer applies to the
   BAD_QUALITY tets after markBadQuality: codes[element] = the bit code (bits 0-2 / 3-5: signs and near-zero area
   coordinates of vertex 3 projected onto the first face; bit 6 + bits 7-8 / 9-11: the edge projection used when the first
   face is itself worse than good_quality^2), match[element] = {rotation, code_index} of matchSliver's tables, {-1,-1} = no
   match.  face0_v [nt][3] = the vertices of getDownward(tet, 2)[0] in THAT FACE's own order (measureTriQuality walks the
   face entity; NULL = the tet's own (v0, v1, v2), which can differ from the reference in the last bit of the face quality).
   only_bad != 0: only tets whose resident flag word carries BAD_QUALITY are classified.  codes [np+npy+nt], match
   [np+npy+nt][2]; unclassified and layer elements get 0 and {-1,-1}.  Always evaluated in the reference's operation
   order.  Synchronous. */
int mag_sliver_codes(mag_ctx* c, const int32_t* face0_v, double good_quality, int only_bad, int32_t* codes, int32_t* match);
/* Size-field transfer to the vertices that will split the SPLIT-marked edges (ma::makeSplitVert, ma/maRefine.cc:129-151;
   SizeField::interpolate, ma/maSize.cc:414-429,523-534): for every edge whose resident flag word carries MAG_SPLIT, in
   edge order, the edge index, the position of the new vertex (xi = 0) and the size-field values it receives --
   iso: field_a = size[n]; aniso: field_a = h[n][3], field_b = R[n][9] (orthogonalised frame); logm: field_b = logM[n][9].
   *n receives the number of SPLIT edges; with all output pointers NULL the call only counts.  cap = capacity of the
   output arrays in vertices. */
int mag_split_vertices(mag_ctx* c, int fp_mode, int64_t cap, int64_t* n, int32_t* edge_idx, double* xyz,
                       double* field_a, double* field_b);

/* Host-side post-processing of ma::getLinearQualitiesInMetricSpace (ma/maStats.cc:12-31), the vector ma::stats returns and
   measureAnisoStats tabulates (test/measureAnisoStats.cc:217-243): for every element with keep[i] != 0 (owned and simplex;
   NULL = all), in order, cbrt(quality) on a 3-D mesh, the signed square root on a 2-D one, through the host's libm as the
   reference does.  out holds up to n values; *n_out = how many were written.  No context, no device work. */
int mag_linear_qualities(int dim, int64_t n, const double* qualities, const uint8_t* keep, double* out, int64_t* n_out);

/* the logM vertex field from sizes + frames, computed on the host exactly as the reference does (libm log):
   variant 0 = LogAnisoSizeField::init from fields, log(1/h/h)   (ma/maSize.cc:491-499)
   variant 1 = LogMEval from a user function, -2*log(h)          (ma/maSize.cc:343-346)
   then uploaded like mag_set_metric_logm.  out_logM (may be NULL) receives the [nv][9] field. */
int mag_set_metric_logm_from_frames(mag_ctx* c, const double* h, const double* R, int variant, double* out_logM);

/* ---- instrumentation (bench.py): per-sweep device times from CUDA events recorded on the context's stream around
   the vertex pass, the edge kernels and the element kernels.  mag_timing_begin(c, n) arms n slots; each mag_sweep fills
   the next one; mag_timing_read synchronizes and writes up to n rows {vertex_ms, edge_ms, elem_ms}; returns rows in *n_out. */
int mag_timing_begin(mag_ctx* c, int max_sweeps);
int mag_timing_read(mag_ctx* c, float* ms /*[max_sweeps][3]*/, int* n_out);
/* number of CUDA kernels this context has launched so far */
int64_t mag_launch_count(const mag_ctx* c);
/* The device layout the whole-part sweeps run over (diagnostics and tests; built by mag_set_mesh, or by the first mag_sweep
   after mag_sweep_host): every edge (which = 0) / tet (which = 1) is filed under its first vertex, the anchor; rows of one
   anchor's entities are cut into slices of 32 rows stored slot-major (see core_b200/csrc/mag_rows.cuh).
   counts[3] = {rows, slices, slots}; anchor [32 * slices] (-1 = padding row); slice_off [slices + 1];
   slots [slots][2] = {other vertex | not-owned << 31, edge index} or [slots][4] = {v1 | not-owned << 31, v2, v3, tet index},
   index -1 = empty slot.  Array pointers may be NULL (counts only). */
int mag_get_row_layout(mag_ctx* c, int which, int64_t* counts, int32_t* anchor, int32_t* slice_off, int32_t* slots);

/* ---- PUMI's native mesh format (.smb, mds/mds_smb.c:120-600) straight into these arrays, without building the mesh database:
   entity order = m->begin(d) order of the loaded mesh, element vertices as getDownward(e, 0, .) returns them (derived with
   MDS's own rule, mds/mds.c:496-508,634-670).  Host code: no device is needed to read a file.  SURVEY 8f row 4. ---- */
typedef struct mag_smb mag_smb;            /* opaque: owns every array it hands out */
typedef struct mag_smb_arrays {
  int dim, version, nparts;
  int64_t nv, ne, ntri, nquad, nt, np, npy, nhex;
  const double* xyz;                       /* [nv][3] */
  const int32_t* edge_v;                   /* [ne][2] */
  const int32_t* tri_v;                    /* [ntri][3] (the elements of a 2-D mesh; the faces of a 3-D one) */
  const int32_t* tet_v;                    /* [nt][4] */
  const int32_t* prism_v;                  /* [np][6] */
  const int32_t* pyr_v;                    /* [npy][5] */
} mag_smb_arrays;
/* *out is set also on failure (MAG_ERR_ARG): mag_smb_last_error(*out) says why; free it either way */
int mag_smb_read(const char* path, mag_smb** out);
void mag_smb_free(mag_smb* s);
const char* mag_smb_last_error(const mag_smb* s);
int mag_smb_get(const mag_smb* s, mag_smb_arrays* out);
/* dense [nv][components] values of a double-valued vertex tag / field (apf keeps the vertex nodes of field <name> in the tag
   <name>_ver, apf/apfTagData.cc): what mag_set_metric_* takes */
int mag_smb_vertex_field(mag_smb* s, const char* name, int* components, const double** values);
/* mag_set_mesh / mag_set_mesh_2d with the file's arrays (every entity owned) */
int mag_set_mesh_smb(mag_ctx* c, const mag_smb* s);

/* ---- multi-GPU: one part per GPU, NCCL over NVLink (replaces PCU on this path only) ---- */
#define MAG_UNIQUE_ID_BYTES 128
int mag_comm_unique_id(void* out_id /*[MAG_UNIQUE_ID_BYTES]*/);
int mag_comm_init(mag_ctx* c, int nranks, int rank, const void* unique_id);
/* part-boundary links in the layout of struct mds_links (mds/mds_net.h:33-38): for each peer part k, idx[k][0..n[k]) are the
   local indices of the edges shared with that peer, both sides listing them in the same order.  peer_owns (may be NULL, and
   each peer_owns[k] may be NULL) marks entries whose owner is the peer's copy (apfPM.cc:109-126, evaluated by the caller). */
int mag_set_edge_links(mag_ctx* c, int npeers, const int32_t* peer, const int64_t* n, const int32_t* const* idx,
                       const uint8_t* const* peer_owns);
/* ma::checkFlagConsistency(a, 1, flag) for every bit of flag_mask: copies exchange their bits and compare; the number of
   disagreeing copies goes to mag_stats.n_flag_mismatch (the reference asserts it is 0) and, where peer_owns says so, the
   owner's bits overwrite the local ones. */
int mag_reconcile_edge_flags(mag_ctx* c, int32_t flag_mask);
/* mag_sweep + mag_reconcile_edge_flags(flag_mask) in one call, the exchange overlapped with the element sweep (an edge's marks are
   final when the edge sweep ends: the exchange starts there, on a side stream, and is waited for at the end).  Same results as
   the two calls in sequence; without links or communicator it is mag_sweep. */
int mag_sweep_reconciled(mag_ctx* c, uint32_t ops, double max_len, double min_len, double good_quality, int use_max_metric, int fp_mode,
                         int32_t flag_mask);
/* ma::checkFlagConsistency as the reference uses it (maRefine.cc:430, maCoarsen.cc:305): copies are compared, nothing is
   repaired; any disagreement -> MAG_ERR_INCONSISTENT (the reference asserts).  *n_mismatch (may be NULL) = disagreeing local
   copies.  Synchronous. */
int mag_check_edge_flag_consistency(mag_ctx* c, int32_t flag_mask, int64_t* n_mismatch);
/* ma::syncFlag semantics: OR the bits of flag_mask across all copies */
int mag_sync_edge_flags(mag_ctx* c, int32_t flag_mask);
/* global statistics: sums of counts, min of min_quality, max of max_length over all parts */
int mag_allreduce_stats(mag_ctx* c, mag_stats* global);

#ifdef __cplusplus
}
#endif
#endif /* MAG_H */
