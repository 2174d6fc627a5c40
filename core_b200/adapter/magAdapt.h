// magAdapt.h -- C++ drop-in adapter between SCOREC/core's MeshAdapt and the mag C ABI (include/mag.h).
//
// Host side of the boundary, written in the reference's own language against the reference's own headers
// (ma/maSize.h, ma/maAdapt.h, ma/maShapeHandler.h, apf/apfMesh2.h, mds/apfMDS.h).  It replaces, for whole-mesh sweeps:
//
//   ma::makeSizeField(...)            ma/maSize.h:108-113   ->  mag::makeSizeField(...)  (same overloads, + CUDA device)
//   ma::markEdgesToSplit(Adapt*)      ma/maRefine.cc:395    ->  mag::markEdgesToSplit(Adapt*)
//   ma::markEdgesToCollapse(Adapt*)   ma/maCoarsen.cc:287   ->  mag::markEdgesToCollapse(Adapt*)
//   ma::markBadQuality(Adapt*)        ma/maShape.cc:132     ->  mag::markBadQuality(Adapt*)
//   ma::getMinQuality(Adapt*)         ma/maShape.cc:152     ->  mag::getMinQuality(Adapt*)
//   ma::getMaximumEdgeLength(m, sf)   ma/maSize.cc:673      ->  mag::getMaximumEdgeLength(m, sf)
//   ma::getEdgeLengthsInMetricSpace / getLinearQualitiesInMetricSpace   ma/maStats.cc:12-45  ->  mag::... (same vectors)
//   ma::stats(m, sf, el, lq, inMetric)  ma/maStats.cc:115-134  ->  mag::stats (both vectors from one sweep)
//   ma::getElementWeights(Adapt*)     ma/maBalance.cc:83    ->  mag::getElementWeights(Adapt*)  (same "ma_weight" tag)
//   ma::getSliverCode / matchSliver   ma/maShape.cc:35-120  ->  mag::getSliverCodes(Adapt*, ...)  (every tet in one sweep)
//   ma::Collapse::tryBothDirections' quality test   ma/maCollapse.cc:88-113  ->  mag::collapseQualities (many candidates at once)
//   ma::getShapeHandler(Adapt*)       ma/maShapeHandler.cc  ->  mag::shapeHandler  (an ma::ShapeHandlerFunction for Input::shapeHandler)
//
// 3-D meshes (tets, with prisms / pyramids as layer elements) and 2-D meshes (triangles: ma::measureTriQuality).
// mag::GpuSizeField IS an ma::SizeField: it can be put in ma::Input::sizeField and the UNMODIFIED reference keeps
// working -- the per-entity virtuals (measure / shouldSplit / shouldCollapse) are answered from the last device sweep
// while the mesh is unchanged, and delegated to the wrapped reference size field (the rest of the reference running as
// before) for entities created since.  A sweep over the whole mesh by the unmodified ma::markEntities loop is detected
// after kSweepDetect consecutive per-entity queries without an intervening mesh change and turned into ONE device
// sweep.  Error convention: like the reference, failures abort through PCU-style fail-fast (mag::fail prints the
// mag_last_error text and calls abort()).
#ifndef MAG_ADAPT_H
#define MAG_ADAPT_H

#include <maSize.h>
#include <maInput.h>
#include <maTables.h>
#include <vector>

struct mag_ctx;
namespace ma { class Adapt; class ShapeHandler; }

namespace mag {

struct Export;   /* the arrays of one MDS export (magAdapt.cc) */

/* Part-boundary edge lists in the layout mag_set_edge_links takes (struct mds_links, mds/mds_net.h:33-38): for every peer
   part the export indices of the edges shared with it, in the SAME order on both sides, and whether the peer's copy is the
   owner (apf::Sharing::getOwner -- apfPM.cc:109-126 for the default NormalSharing). */
struct EdgeLinks {
  std::vector<int> peer;
  std::vector<std::vector<int> > idx;
  std::vector<std::vector<unsigned char> > peerOwns;
};
/* Built from apf::Sharing alone, no communication: both sides sort a peer's list by the entity pointer of the copy that
   lives on the LOWER-numbered part (its own pointer there, the remote pointer of apf::Copy on the other side; MDS pointers
   grow with the entity index, mds/apfMDS.cc fromEnt/toEnt).  edges = the exported edges in export order; self = this
   part's id (m->getPCU()->Self()). */
void buildEdgeLinks(apf::Sharing* sh, int self, const std::vector<ma::Entity*>& edges, EdgeLinks& out);

int exportSelfCheck(ma::Mesh* m, apf::Field* sizes, apf::Field* frames, ma::Tag* flags, int threads, double* times);

class GpuSizeField : public ma::SizeField
{
  public:
    ~GpuSizeField();
    /* ma::SizeField */
    double measure(ma::Entity* e);
    bool shouldSplit(ma::Entity* edge);
    bool shouldCollapse(ma::Entity* edge);
    void interpolate(apf::MeshElement* parent, ma::Vector const& xi, ma::Entity* newVert);
    void getTransform(apf::MeshElement* e, ma::Vector const& xi, ma::Matrix& t);
    double getWeight(ma::Entity* e);
    void onRefine(ma::Entity* parent, ma::EntityArray& newEntities);
    void onCavity(ma::EntityArray& oldElements, ma::EntityArray& newEntities);
    int getTransferDimension();
    bool hasNodesOn(int dimension);

    /* the CONNECTIVITY changed behind our back without any SizeField callback (a count-preserving edit made by the caller's own
       code).  Moved vertices (apf::Mesh2::setPoint, snapping) and in-place edits of the size fields need no call: every
       bulk sweep and every sweep start re-reads the vertices and compares a hash with what the device holds. */
    void invalidate() { dirty = true; topoValid = false; }
    /* MAG_FP_STRICT (default), MAG_FP_FAST, or MAG_FP_FAST_LISTED (near-threshold edges listed instead of re-evaluated: the marks
       may then differ from the reference's on exactly those edges), see include/mag.h */
    void setArithmetic(int fp_mode) { fpMode = fp_mode; invalidate(); }
    /* host threads for the read-only part of the MDS walk of an export (getDownward / isOwned of every edge and element, the
       "ma_flags" tag reads); default 1 = the reference's own single-threaded access pattern.  MDS adjacency queries do not
       write, so more threads are safe while nothing else touches the mesh (no other thread of the caller does during a
       MeshAdapt call); opt-in because the reference makes no such promise in writing. */
    void setExportThreads(int n) { exportThreads = n < 1 ? 1 : n; }
    /* re-export the mesh + field and run one full device sweep now */
    void refresh(double goodQuality = -1);
    /* Several parts (mesh->getPCU()->Peers() > 1): the flag words of part-boundary edges are compared across parts with NCCL
       where the reference calls ma::checkFlagConsistency (maRefine.cc:430, maCoarsen.cc:305 -> mag_check_edge_flag_consistency)
       and the counts / min / max are reduced with mag_allreduce_stats instead of PCU.  Set up on the first export: the NCCL
       unique id travels from part 0 over PCU, the lists come from buildEdgeLinks.  sharing: whose view of ownership to use
       (NULL = apf::getSharing(mesh), which the adapter then owns). */
    void setSharing(apf::Sharing* sharing) { userSharing = sharing; }
    bool multiPart() const { return commReady; }

    ma::Mesh* mesh;
    ma::SizeField* wrapped;  /* the reference size field built from the same inputs (owned) */
    mag_ctx* ctx;

  private:
    friend struct Access;
    friend int exportSelfCheck(ma::Mesh* m, apf::Field* sizes, apf::Field* frames, ma::Tag* flags, int threads, double* times);
    GpuSizeField();
    enum { kSweepDetect = 4096 };
    int kind;        /* 1 iso, 2 aniso, 3 logm */
    int logVariant;  /* 0 fields (maSize.cc:491-499), 1 user function (maSize.cc:343-346) */
    int fpMode;
    bool dirty;
    /* the device holds the export of the mesh as it is now: set by an export, cleared by every callback a mesh-modifying
       operator makes (interpolate / getTransform / getWeight / onRefine / onCavity), by any out-of-order per-entity query,
       by a change of the entity counts and by invalidate().  While it holds, consecutive bulk sweeps (split, collapse, bad
       quality ... of one MeshAdapt iteration) share ONE export instead of walking MDS again. */
    bool topoValid;
    Export* exported;
    int exportThreads;
    long streak;     /* consecutive per-entity sweep-like queries since the last mesh change */
    double lastGoodQuality;
    apf::Field* fSizes; apf::Field* fFrames; apf::Field* fIso;
    ma::AnisotropicFunction* fnAniso; ma::IsotropicFunction* fnIso;
    std::vector<int> edgeSlot, tetSlot;           /* MDS index -> position in the exported arrays (-1: not exported) */
    std::vector<int> vertSlot;                    /* MDS index of a vertex -> exported vertex id */
    std::vector<double> lengths, qualities;       /* host copies of the last sweep */
    std::vector<int> edgeFlags, elemFlags;        /* flags of the last sweep run on zero incoming words */
    std::vector<int> flagScratch[4];              /* flag words in transit of the bulk marks (buffers kept between sweeps) */
    long nNonSimplex;
    int lastDim, lastId;  /* last per-entity query: dimension and MDS index */
  apf::Sharing* userSharing;   /* not owned */
  apf::Sharing* ownSharing;    /* from apf::getSharing, owned */
  bool commReady;              /* mag_comm_init done for this context */
  unsigned long long snapshotHash;   /* vertex hash (coordinates + field) of the export the snapshot was swept from */
    bool serve(ma::Entity* e, int dim, int& slot);
};

/* same overloads as ma::makeSizeField (ma/maSize.h:108-113) plus the CUDA device ordinal */
GpuSizeField* makeSizeField(ma::Mesh* m, apf::Field* sizes, apf::Field* frames, bool logInterpolation = false, int device = 0);
GpuSizeField* makeSizeField(ma::Mesh* m, ma::AnisotropicFunction* f, bool logInterpolation = false, int device = 0);
GpuSizeField* makeSizeField(ma::Mesh* m, apf::Field* size, int device = 0);
GpuSizeField* makeSizeField(ma::Mesh* m, ma::IsotropicFunction* f, int device = 0);

/* bulk replacements of the reference sweeps; a->sizeField must be a GpuSizeField (otherwise they abort) */
long markEdgesToSplit(ma::Adapt* a);
long markEdgesToCollapse(ma::Adapt* a);
int markBadQuality(ma::Adapt* a);
double getMinQuality(ma::Adapt* a);
double getMaximumEdgeLength(ma::Mesh* m, ma::SizeField* sf);
void getEdgeLengthsInMetricSpace(ma::Mesh* m, ma::SizeField* sf, std::vector<double>& lengths);
void getLinearQualitiesInMetricSpace(ma::Mesh* m, ma::SizeField* sf, std::vector<double>& qualities);
/* ma::stats (ma/maStats.cc:115-134): both vectors from ONE device sweep when inMetric, the reference's own physical-space
   loops otherwise */
void stats(ma::Mesh* m, ma::SizeField* sf, std::vector<double>& edgeLengths, std::vector<double>& linearQualities, bool inMetric);

/* ma::getElementWeights (ma/maBalance.cc:83-97): creates and fills the "ma_weight" element tag; the caller destroys it */
ma::Tag* getElementWeights(ma::Adapt* a);

/* ma::getSliverCode / ma::matchSliver (ma/maShape.cc:35-120) of every tet in one device sweep: codes[i] / matches[i] belong to
   the i-th element of m->begin(3) (layer elements: 0 and {-1,-1}).  The first face of every tet is exported in the face
   entity's own vertex order, which is what the reference's measureTriQuality walks. */
void getSliverCodes(ma::Adapt* a, std::vector<int>& codes, std::vector<ma::CodeMatch>& matches);

/* Exports read MDS's own arrays (struct mds / mds_apf / mds_tag) when the mesh is an MDS mesh -- see magAdapt.cc, "the direct
   export route"; off = every entity through apf::Mesh2's public interface, as in round 1.  Process-wide; the environment
   variable MAG_ADAPTER_PUBLIC_API_ONLY also switches it off. */
void setDirectMds(bool on);
/* where the host time of the bulk sweeps went since the last reset (seconds, process-wide): the MDS export, the re-validation
   of the vertices of a kept export, uploads (mesh / coordinates / field), reading the incoming flag words, the device calls
   of a sweep (flags up, sweep, statistics, flags down), writing changed flag words back, a snapshot refresh (sweep + downloads) */
struct Profile { double export_s, revalidate_s, upload_s, flags_in_s, device_s, flags_out_s, refresh_s; };
Profile& profile();
/* Test hook, host only (no device is touched): exports m through both routes (Aniso field from sizes + frames) and compares every
   array, the slot tables, the change detection of moved vertices / edited field values and -- flags != NULL -- the direct read of
   an int tag.  Returns a bit mask of what differed (0 = identical); times[0] / times[1] = seconds of the public / direct export. */
int exportSelfCheck(ma::Mesh* m, apf::Field* sizes, apf::Field* frames, ma::Tag* flags, int threads, double* times);

/* The quality test of ma::Collapse::tryBothDirections (ma/maCollapse.cc:88-113) for MANY edge-collapse candidates in one device
   call (mag_collapse_quality): candidate i collapses vertsToCollapse[i], an end of edges[i], onto the other end.  newWorst[i] =
   the worst quality of the elements the collapse would rebuild (what hasWorseQuality looks at), oldWorst[i] =
   Collapse::getOldQuality().  Nothing is built in the mesh.  A coarsening pass evaluates the members of one independent set
   (ma/maCoarsen.cc:163-169) -- or every marked edge, both directions -- at once and only runs the topological collapse for the
   candidates that pass  !(newWorst < min(goodQuality, max(oldWorst, validQuality))). */
void collapseQualities(ma::Adapt* a, const std::vector<ma::Entity*>& edges, const std::vector<ma::Entity*>& vertsToCollapse,
                       std::vector<double>& newWorst, std::vector<double>& oldWorst);

/* ma::ShapeHandlerFunction: in->shapeHandler = mag::shapeHandler; getQuality(e) is then served from the device sweep */
ma::ShapeHandler* shapeHandler(ma::Adapt* a);

}

#endif
