// magAdapt.cc -- see magAdapt.h.  Builds against the reference's headers; everything numeric happens behind the
// C ABI of include/mag.h on the GPU.  The wrapped reference ma::SizeField is only consulted for entities that are
// not part of a device sweep (cavity operators on freshly created entities) -- the rest of the reference running
// as before, not a fallback of the sweep.
#include "magAdapt.h"
#include "../../include/mag.h"
#include <maAdapt.h>
#include <maShapeHandler.h>
#include <maShape.h>
#include <maStats.h>
#include <apfMDS.h>
#include <apfMesh2.h>
#include <apfShape.h>
#include <apf.h>
#include <pcu_util.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <thread>
#include <functional>
#include <algorithm>
#include <map>
#include <stdint.h>
#include <PCU.h>

namespace ma {
/* external linkage in the reference, declared in no installed header (maBalance.cc:74-81) */
double getElementWeight(Adapt* a, Entity* e);
}

namespace mag {

static void fail(mag_ctx* c, const char* what, int rc)
{
  fprintf(stderr, "mag adapter: %s failed (%d): %s\n", what, rc, mag_last_error(c));
  abort(); /* the reference's convention: PCU_ALWAYS_ASSERT -> abort (pcu_util.h:44-69) */
}
#define MAG_DO(c, call) do { int rc_ = (call); if (rc_) fail((c), #call, rc_); } while (0)

/* static chunks of [0, n) over nthreads host threads (the calling thread takes the first chunk) */
static void parallelFor(size_t n, int nthreads, const std::function<void(size_t, size_t)>& body)
{
  if (nthreads <= 1 || n < 4096) { body(0, n); return; }
  std::vector<std::thread> pool;
  const size_t per = (n + (size_t)nthreads - 1) / (size_t)nthreads;
  for (int t = 1; t < nthreads; ++t) {
    const size_t b = per * (size_t)t, e = b + per < n ? b + per : n;
    if (b < e) pool.push_back(std::thread(body, b, e));
  }
  body(0, per < n ? per : n);
  for (size_t i = 0; i < pool.size(); ++i) pool[i].join();
}

/* FNV-1a over the bytes of the per-vertex arrays: what the device holds of the coordinates and the size field */
static unsigned long long hashBytes(unsigned long long h, const void* p, size_t n)
{
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

void buildEdgeLinks(apf::Sharing* sh, int self, const std::vector<ma::Entity*>& edges, EdgeLinks& out)
{
  struct Item { uintptr_t key; int idx; unsigned char peerOwns; };
  std::map<int, std::vector<Item> > lists;   /* peers in increasing order */
  apf::CopyArray copies;
  for (size_t i = 0; i < edges.size(); ++i) {
    ma::Entity* e = edges[i];
    if (!sh->isShared(e)) continue;
    const int owner = sh->getOwner(e);
    copies.setSize(0);
    sh->getCopies(e, copies);
    for (size_t k = 0; k < copies.getSize(); ++k) {
      const int p = copies[k].peer;
      if (p == self) continue;
      Item it;
      it.key = self < p ? (uintptr_t)e : (uintptr_t)copies[k].entity;
      it.idx = (int)i;
      it.peerOwns = owner == p ? 1 : 0;
      lists[p].push_back(it);
    }
  }
  out.peer.clear(); out.idx.clear(); out.peerOwns.clear();
  for (std::map<int, std::vector<Item> >::iterator l = lists.begin(); l != lists.end(); ++l) {
    std::vector<Item>& v = l->second;
    std::sort(v.begin(), v.end(), [](const Item& a, const Item& b) { return a.key < b.key; });
    out.peer.push_back(l->first);
    out.idx.push_back(std::vector<int>(v.size()));
    out.peerOwns.push_back(std::vector<unsigned char>(v.size()));
    for (size_t i = 0; i < v.size(); ++i) { out.idx.back()[i] = v[i].idx; out.peerOwns.back()[i] = v[i].peerOwns; }
  }
}

struct Export {
  unsigned long long vertHash;            /* hash of xyz / ma / mb as uploaded */
  std::vector<double> xyz, ma, mb;
  std::vector<int> edge_v, tet_v, prism_v, pyr_v, tri_v;   /* tri_v: the elements of a 2-D mesh */
  std::vector<unsigned char> edge_owned, elem_owned;
  std::vector<ma::Entity*> edges, elems; /* iteration order; elems = prisms | pyramids | tets */
  bool logm_direct;                       /* mb holds the reference's ma_logM field, not frames */
};

struct Access {
  static GpuSizeField* create() { return new GpuSizeField(); }
  static void init(GpuSizeField* g, int kind, int logVariant, apf::Field* sizes, apf::Field* frames, apf::Field* iso,
                   ma::AnisotropicFunction* fa, ma::IsotropicFunction* fi)
  {
    g->kind = kind; g->logVariant = logVariant;
    g->fSizes = sizes; g->fFrames = frames; g->fIso = iso; g->fnAniso = fa; g->fnIso = fi;
  }
  static double lengthAt(GpuSizeField* g, size_t k) { return g->lengths[k]; }
  static void resetOrder(GpuSizeField* g) { g->lastDim = g->lastId = -1; }
  static size_t exportedElems(GpuSizeField* g) { return g->exported ? g->exported->elems.size() : 0; }
  static bool isDirty(GpuSizeField* g) { return g->dirty; }
  static int fpMode(GpuSizeField* g) { return g->fpMode; }
  static int vertSlotOf(GpuSizeField* g, ma::Entity* v) { return g->vertSlot[apf::getMdsIndex(g->mesh, v)]; }
  static long nonSimplex(GpuSizeField* g) { return g->nNonSimplex; }
  static int tetSlotOf(GpuSizeField* g, ma::Entity* e) { return g->tetSlot[apf::getMdsIndex(g->mesh, e)]; }
  static double qualityOf(GpuSizeField* g, ma::Entity* e) { return g->qualities[g->tetSlot[apf::getMdsIndex(g->mesh, e)]]; }
  static bool serveQuality(GpuSizeField* g, ma::Entity* e, double goodQuality, double& q)
  {
    int slot;
    g->lastGoodQuality = goodQuality;
    if (!g->serve(e, g->mesh->getDimension(), slot)) return false;
    q = g->qualities[slot];
    return true;
  }

  static unsigned long long vertexHash(const Export& x)
  {
    unsigned long long h = 1469598103934665603ull;
    h = hashBytes(h, x.xyz.data(), x.xyz.size() * sizeof(double));
    h = hashBytes(h, x.ma.data(), x.ma.size() * sizeof(double));
    return hashBytes(h, x.mb.data(), x.mb.size() * sizeof(double));
  }

  /* the per-vertex part of an export: coordinates and size-field values in m->begin(0) order, vertex ids compacted to 0..nv-1 */
  static void exportVertices(GpuSizeField* g, Export& x)
  {
    ma::Mesh* m = g->mesh;
    size_t nv = m->count(0);
    std::vector<int>& vslot = g->vertSlot;
    vslot.assign(vslot.size(), -1);
    x.xyz.resize(3 * nv);
    /* LogAnisoSizeField built from fields keeps its own "ma_logM" vertex field and is the only thing it updates when
       vertices are created (interpolate / onRefine / onCavity, maSize.cc:523-561): the sizes and frames it was built from go
       stale for new vertices, so the export reads the logM field itself */
    apf::Field* logM = (g->kind == 3 && !g->fnAniso) ? m->findField("ma_logM") : 0;
    if (g->kind == 3 && !g->fnAniso && !logM) { fprintf(stderr, "mag adapter: the wrapped LogAnisoSizeField has no ma_logM field\n"); abort(); }
    x.logm_direct = logM != 0;
    if (g->kind == 1) x.ma.resize(nv);
    else if (logM) x.mb.resize(9 * nv);
    else { x.ma.resize(3 * nv); x.mb.resize(9 * nv); }
    apf::MeshIterator* it = m->begin(0);
    ma::Entity* e;
    int k = 0;
    while ((e = m->iterate(it))) {
      int id = apf::getMdsIndex(m, e);
      if ((size_t)id >= vslot.size()) vslot.resize((size_t)id + 1 + vslot.size() / 2, -1);
      vslot[id] = k;
      ma::Vector p;
      m->getPoint(e, 0, p);
      x.xyz[3 * k] = p[0]; x.xyz[3 * k + 1] = p[1]; x.xyz[3 * k + 2] = p[2];
      if (g->kind == 1) {
        x.ma[k] = g->fnIso ? g->fnIso->getValue(e) : apf::getScalar(g->fIso, e, 0);
      } else if (logM) {
        ma::Matrix M;
        apf::getMatrix(logM, e, 0, M);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) x.mb[9 * k + 3 * i + j] = M[i][j];
      } else {
        ma::Matrix R; ma::Vector h;
        if (g->fnAniso) g->fnAniso->getValue(e, R, h);
        else { apf::getVector(g->fSizes, e, 0, h); apf::getMatrix(g->fFrames, e, 0, R); }
        for (int i = 0; i < 3; ++i) x.ma[3 * k + i] = h[i];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) x.mb[9 * k + 3 * i + j] = R[i][j];
      }
      ++k;
    }
    m->end(it);
    x.vertHash = vertexHash(x);
  }

  /* one pass over the mesh in m->begin(d) order (mds.c:745-777) */
  static void exportMesh(GpuSizeField* g, Export& x)
  {
    ma::Mesh* m = g->mesh;
    const int mdim = m->getDimension();
    if (mdim != 3 && mdim != 2) { fprintf(stderr, "mag adapter: only 2D and 3D meshes are supported\n"); abort(); }
    exportVertices(g, x);
    std::vector<int>& vslot = g->vertSlot;
    apf::MeshIterator* it;
    ma::Entity* e;
    int k;
    size_t ne = m->count(1);
    x.edge_v.resize(2 * ne); x.edge_owned.resize(ne); x.edges.resize(ne);
    g->edgeSlot.assign(g->edgeSlot.size(), -1);
    it = m->begin(1); k = 0;
    while ((e = m->iterate(it))) {    /* the iteration itself stays on one thread; the adjacency queries below do not */
      x.edges[k] = e;
      int id = apf::getMdsIndex(m, e);
      if ((size_t)id >= g->edgeSlot.size()) g->edgeSlot.resize((size_t)id + 1 + g->edgeSlot.size() / 2, -1);
      g->edgeSlot[id] = k;
      ++k;
    }
    m->end(it);
    parallelFor(ne, g->exportThreads, [&](size_t b, size_t en) {
      for (size_t i = b; i < en; ++i) {
        apf::Downward dv;
        m->getDownward(x.edges[i], 0, dv);
        x.edge_v[2 * i] = vslot[apf::getMdsIndex(m, dv[0])];
        x.edge_v[2 * i + 1] = vslot[apf::getMdsIndex(m, dv[1])];
        x.edge_owned[i] = m->isOwned(x.edges[i]) ? 1 : 0;
      }
    });
    /* dimension 3 iterates prisms, pyramids, tets (MDS type order, mds.h:16-26); on a 2-D mesh the elements are the
       triangles (ma::measureTriQuality), kept in the same "te" list and slot table */
    std::vector<ma::Entity*> pr, py, te;
    it = m->begin(mdim);
    while ((e = m->iterate(it))) {
      int t = m->getType(e);
      if (mdim == 3 && t == apf::Mesh::PRISM) pr.push_back(e);
      else if (mdim == 3 && t == apf::Mesh::PYRAMID) py.push_back(e);
      else if (t == (mdim == 3 ? apf::Mesh::TET : apf::Mesh::TRIANGLE)) te.push_back(e);
      else { fprintf(stderr, "mag adapter: element type %d is not supported\n", t); abort(); }
    }
    m->end(it);
    x.elems.clear();
    x.elems.insert(x.elems.end(), pr.begin(), pr.end());
    x.elems.insert(x.elems.end(), py.begin(), py.end());
    x.elems.insert(x.elems.end(), te.begin(), te.end());
    g->nNonSimplex = (long)(pr.size() + py.size());
    x.elem_owned.resize(x.elems.size());
    auto conn = [&](std::vector<ma::Entity*>& v, int n, std::vector<int>& out) {
      out.resize(v.size() * n);
      parallelFor(v.size(), g->exportThreads, [&](size_t b, size_t en) {
        for (size_t i = b; i < en; ++i) {
          apf::Downward dv;
          m->getDownward(v[i], 0, dv);
          for (int j = 0; j < n; ++j) out[i * n + j] = vslot[apf::getMdsIndex(m, dv[j])];
        }
      });
    };
    conn(pr, 6, x.prism_v); conn(py, 5, x.pyr_v);
    if (mdim == 3) conn(te, 4, x.tet_v); else conn(te, 3, x.tri_v);
    g->tetSlot.assign(g->tetSlot.size(), -1);
    for (size_t i = 0; i < te.size(); ++i) {
      int id = apf::getMdsIndex(m, te[i]);
      if ((size_t)id >= g->tetSlot.size()) g->tetSlot.resize((size_t)id + 1 + g->tetSlot.size() / 2, -1);
      g->tetSlot[id] = (int)(g->nNonSimplex + i);
    }
    parallelFor(x.elems.size(), g->exportThreads, [&](size_t b, size_t en) {
      for (size_t i = b; i < en; ++i) x.elem_owned[i] = m->isOwned(x.elems[i]) ? 1 : 0;
    });
  }

  static void upload(GpuSizeField* g, Export& x)
  {
    mag_ctx* c = g->ctx;
    if (g->mesh->getDimension() == 2)
      MAG_DO(c, mag_set_mesh_2d(c, (int64_t)(x.xyz.size() / 3), x.xyz.data(), (int64_t)x.edges.size(), x.edge_v.data(),
                                (int64_t)(x.tri_v.size() / 3), x.tri_v.data(), x.edge_owned.data(), x.elem_owned.data()));
    else
    MAG_DO(c, mag_set_mesh(c, (int64_t)(x.xyz.size() / 3), x.xyz.data(), (int64_t)x.edges.size(), x.edge_v.data(),
                           (int64_t)(x.tet_v.size() / 4), x.tet_v.data(), (int64_t)(x.prism_v.size() / 6), x.prism_v.data(),
                           (int64_t)(x.pyr_v.size() / 5), x.pyr_v.data(), x.edge_owned.data(), x.elem_owned.data()));
    uploadField(g, x);
    uploadLinks(g, x);
  }
  /* several parts: NCCL communicator (once per context; the unique id goes from part 0 to the others over PCU) and the
     part-boundary edge lists of this export */
  static void uploadLinks(GpuSizeField* g, Export& x)
  {
    pcu::PCU* P = g->mesh->getPCU();
    if (!P || P->Peers() <= 1) return;
    mag_ctx* c = g->ctx;
    if (!g->commReady) {
      char id[MAG_UNIQUE_ID_BYTES];
      if (P->Self() == 0) MAG_DO(c, mag_comm_unique_id(id));
      P->Begin();
      if (P->Self() == 0)
        for (int r = 1; r < P->Peers(); ++r) P->Pack(r, id, sizeof(id));
      P->Send();
      while (P->Receive()) P->Unpack(id, sizeof(id));
      MAG_DO(c, mag_comm_init(c, P->Peers(), P->Self(), id));
      g->commReady = true;
    }
    apf::Sharing* sh = g->userSharing;
    if (!sh) {
      if (!g->ownSharing) g->ownSharing = apf::getSharing(g->mesh);
      sh = g->ownSharing;
    }
    EdgeLinks L;
    buildEdgeLinks(sh, P->Self(), x.edges, L);
    std::vector<int64_t> n(L.peer.size());
    std::vector<const int32_t*> idx(L.peer.size());
    std::vector<const uint8_t*> own(L.peer.size());
    for (size_t k = 0; k < L.peer.size(); ++k) { n[k] = (int64_t)L.idx[k].size(); idx[k] = L.idx[k].data(); own[k] = L.peerOwns[k].data(); }
    MAG_DO(c, mag_set_edge_links(c, (int)L.peer.size(), L.peer.data(), n.data(), idx.data(), own.data()));
  }
  static void uploadField(GpuSizeField* g, Export& x)
  {
    mag_ctx* c = g->ctx;
    if (g->kind == 1) MAG_DO(c, mag_set_metric_iso(c, x.ma.data()));
    else if (g->kind == 2) MAG_DO(c, mag_set_metric_aniso(c, x.ma.data(), x.mb.data()));
    else if (x.logm_direct) MAG_DO(c, mag_set_metric_logm(c, x.mb.data()));
    else MAG_DO(c, mag_set_metric_logm_from_frames(c, x.ma.data(), x.mb.data(), g->logVariant, 0));
    MAG_DO(c, mag_synchronize(c));
  }

  /* the export of the mesh as it is now, on the host (entity lists) and on the device: reused while topoValid holds */
  static Export& ensureExported(GpuSizeField* g)
  {
    ma::Mesh* m = g->mesh;
    const int dim = m->getDimension();
    if (g->exported && g->topoValid && g->exported->edges.size() == m->count(1) && g->exported->elems.size() == m->count(dim) &&
        g->exported->xyz.size() == 3 * m->count(0)) {
      /* the connectivity is trusted (no callback, no out-of-order query, same counts); coordinates and field values are
         not -- snapping (ma/ma.cc:37, setPoint) and in-place edits of the size fields reach no callback.  The vertex walk is a
         fourteenth of an export: redo it, and if anything moved, send the per-vertex arrays again (connectivity stays) */
      Export& x = *g->exported;
      const unsigned long long before = x.vertHash;
      exportVertices(g, x);
      if (x.vertHash != before) {
        MAG_DO(g->ctx, mag_set_coords(g->ctx, x.xyz.data()));
        uploadField(g, x);
        g->dirty = true;
      }
      return x;
    }
    if (!g->exported) g->exported = new Export;
    *g->exported = Export();
    exportMesh(g, *g->exported);
    upload(g, *g->exported);
    g->topoValid = true;
    return *g->exported;
  }

  static int readFlags(ma::Mesh* m, ma::Tag* tag, ma::Entity* e)
  {
    /* ma::getFlags (maAdapt.cc:80-88): 0 when the entity has no tag */
    if (!m->hasTag(e, tag)) return 0;
    int f;
    m->getIntTag(e, tag, &f);
    return f;
  }

  /* one device sweep with the incoming "ma_flags" words of the Adapt; writes the changed words back.  Returns stats. */
  static mag_stats sweepWithAdaptFlags(GpuSizeField* g, ma::Adapt* a, unsigned ops)
  {
    Export& x = ensureExported(g);
    ma::Mesh* m = g->mesh;
    /* only the dimension the sweep works on has its flag words read, sent and written back */
    const bool on_edges = ops & (MAG_OP_MARK_SPLIT | MAG_OP_MARK_COLLAPSE), on_elems = ops & MAG_OP_MARK_BAD;
    std::vector<int> ef(on_edges ? x.edges.size() : 0), lf(on_elems ? x.elems.size() : 0);
    parallelFor(ef.size(), g->exportThreads, [&](size_t b, size_t en) {
      for (size_t i = b; i < en; ++i) ef[i] = readFlags(m, a->flagsTag, x.edges[i]);
    });
    parallelFor(lf.size(), g->exportThreads, [&](size_t b, size_t en) {
      for (size_t i = b; i < en; ++i) lf[i] = readFlags(m, a->flagsTag, x.elems[i]);
    });
    mag_ctx* c = g->ctx;
    MAG_DO(c, mag_set_flags(c, on_edges ? ef.data() : 0, on_elems ? lf.data() : 0));
    MAG_DO(c, mag_sweep(c, ops, ma::MAXLENGTH, ma::MINLENGTH, a->input->goodQuality, 1, g->fpMode));
    mag_stats st;
    if (g->commReady) {
      /* where the reference calls checkFlagConsistency (maRefine.cc:430, maCoarsen.cc:305; it asserts): every copy of a shared
         edge must have come out with the same mark on every part */
      if (on_edges) MAG_DO(c, mag_check_edge_flag_consistency(c, (ops & MAG_OP_MARK_SPLIT ? MAG_SPLIT : 0) | (ops & MAG_OP_MARK_COLLAPSE ? MAG_COLLAPSE : 0), 0));
      MAG_DO(c, mag_allreduce_stats(c, &st));   /* owned counts summed, min / max reduced: PCU Add / Min / Max of the reference */
    } else
    MAG_DO(c, mag_get_stats(c, &st)); /* MAG_ERR_FLAG_STATE here == the reference's assert at maAdapt.cc:308 */
    if (on_edges || on_elems) {
      std::vector<int> ef2(ef.size()), lf2(lf.size());
      MAG_DO(c, mag_get_flags(c, on_edges ? ef2.data() : 0, on_elems ? lf2.data() : 0));
      for (size_t i = 0; i < ef.size(); ++i) if (ef2[i] != ef[i]) ma::setFlags(a, x.edges[i], ef2[i]);
      for (size_t i = 0; i < lf.size(); ++i) if (lf2[i] != lf[i]) ma::setFlags(a, x.elems[i], lf2[i]);
    }
    g->dirty = true; /* the per-entity snapshot (zero incoming flags) was not refreshed by this sweep */
    g->lastDim = g->lastId = -1;
    return st;
  }
};

GpuSizeField::GpuSizeField()
  : mesh(0), wrapped(0), ctx(0), kind(0), logVariant(0), fpMode(MAG_FP_STRICT), dirty(true), topoValid(false), exported(0), exportThreads(1), streak(0), lastGoodQuality(-1),
    fSizes(0), fFrames(0), fIso(0), fnAniso(0), fnIso(0), nNonSimplex(0), lastDim(-1), lastId(-1), userSharing(0), ownSharing(0), commReady(false), snapshotHash(0)
{
}

GpuSizeField::~GpuSizeField()
{
  if (ctx) mag_destroy(ctx);
  delete exported;
  delete ownSharing;
  delete wrapped; /* like the reference: an AnisoSizeField destroys the fields it was built from (maSize.cc:385-389) */
}

struct Maker {
  static GpuSizeField* make(ma::Mesh* m, ma::SizeField* wrapped, int kind, int logVariant, apf::Field* sizes,
                            apf::Field* frames, apf::Field* iso, ma::AnisotropicFunction* fa, ma::IsotropicFunction* fi,
                            int device);
};

/* ------------------------------------------------------------------ per-entity service */
bool GpuSizeField::serve(ma::Entity* e, int dim, int& slot)
{
  /* A whole-mesh sweep by the unmodified reference (markEntities, getMaximumEdgeLength, ma::stats) shows up as a long run
     of per-entity queries in mesh iteration order (increasing MDS index of one type) with no size-field callback and no
     change of the entity counts in between.  Only such a run is answered from a device sweep: kSweepDetect queries into
     it the mesh is exported and swept ONCE, and the rest of the run is served from that snapshot.  Anything else -- an
     out-of-order query, a callback, a changed count: a cavity operator at work -- goes to the wrapped reference field. */
  const int id = apf::getMdsIndex(mesh, e);
  const bool in_order = (dim == lastDim && id > lastId);
  /* a whole-mesh loop of the reference starts at the first entity of m->begin(dim): a query for that entity opens a sweep at
     once (meshes of any size; the first 4095 queries no longer go to the CPU).  A loop whose first entities are skipped by
     their flags is still recognised by its length (kSweepDetect in-order queries). */
  bool first = false;
  if (!in_order) {
    apf::MeshIterator* it = mesh->begin(dim);
    ma::Entity* e0 = mesh->iterate(it);
    mesh->end(it);
    first = (e0 == e);
  }
  lastDim = dim;
  lastId = id;
  if (!in_order && !first) { dirty = true; streak = 0; topoValid = false; return false; }
  if (!dirty && ((long)mesh->count(1) != (long)lengths.size() || (long)mesh->count(mesh->getDimension()) != (long)qualities.size())) dirty = true;
  if (first) {
    /* sweep start: the snapshot is reused only if the mesh is still what it was taken from (ensureExported re-reads the
       vertices and compares) */
    if (!dirty) {
      Export& x = Access::ensureExported(this);
      if (x.vertHash != snapshotHash) dirty = true;
    }
    if (dirty) refresh(lastGoodQuality);
  } else if (dirty) {
    /* in order, but not (yet) known to be a sweep: the wrapped field answers, and the device export is no longer trusted
       either -- a count-preserving operator (swap, snap) may be at work without a callback reaching us */
    if (++streak < kSweepDetect) { topoValid = false; return false; }
    refresh(lastGoodQuality);
  }
  const std::vector<int>& map = dim == 1 ? edgeSlot : tetSlot;
  if ((size_t)id >= map.size()) return false;
  slot = map[id];
  return slot >= 0;
}

void GpuSizeField::refresh(double goodQuality)
{
  Export& x = Access::ensureExported(this);
  mag_ctx* c = ctx;
  MAG_DO(c, mag_set_flags(c, 0, 0));
  unsigned ops = MAG_OP_LENGTHS | MAG_OP_MARK_SPLIT | MAG_OP_MARK_COLLAPSE | MAG_OP_QUALITIES;
  /* non-simplex elements never reach a quality predicate here: only tets are marked */
  MAG_DO(c, mag_sweep(c, ops, ma::MAXLENGTH, ma::MINLENGTH, goodQuality < 0 ? 0.0 : goodQuality, 1, fpMode));
  mag_stats st;
  MAG_DO(c, mag_get_stats(c, &st));
  lengths.resize(x.edges.size());
  qualities.resize(x.elems.size());
  edgeFlags.resize(x.edges.size());
  elemFlags.resize(x.elems.size());
  MAG_DO(c, mag_get_edge_lengths(c, lengths.data()));
  MAG_DO(c, mag_get_qualities(c, qualities.data()));
  MAG_DO(c, mag_get_flags(c, edgeFlags.data(), elemFlags.data()));
  lastGoodQuality = goodQuality;
  dirty = false;
  streak = 0;
  snapshotHash = x.vertHash;
}

double GpuSizeField::measure(ma::Entity* e)
{
  int slot;
  if (mesh->getType(e) == apf::Mesh::EDGE && serve(e, 1, slot)) return lengths[slot];
  return wrapped->measure(e);
}
bool GpuSizeField::shouldSplit(ma::Entity* edge)
{
  int slot;
  if (serve(edge, 1, slot)) return (edgeFlags[slot] & MAG_SPLIT) != 0;
  return wrapped->shouldSplit(edge);
}
bool GpuSizeField::shouldCollapse(ma::Entity* edge)
{
  int slot;
  if (serve(edge, 1, slot)) return (edgeFlags[slot] & MAG_COLLAPSE) != 0;
  return wrapped->shouldCollapse(edge);
}
/* everything below is a mesh-modification-time call: the snapshot is no longer trusted afterwards */
void GpuSizeField::interpolate(apf::MeshElement* parent, ma::Vector const& xi, ma::Entity* newVert)
{
  dirty = true; streak = 0; topoValid = false;
  wrapped->interpolate(parent, xi, newVert);
}
void GpuSizeField::getTransform(apf::MeshElement* e, ma::Vector const& xi, ma::Matrix& t)
{
  dirty = true; streak = 0; topoValid = false;
  wrapped->getTransform(e, xi, t);
}
double GpuSizeField::getWeight(ma::Entity* e)
{
  dirty = true; streak = 0; topoValid = false;
  return wrapped->getWeight(e);
}
void GpuSizeField::onRefine(ma::Entity* parent, ma::EntityArray& newEntities)
{
  dirty = true; streak = 0; topoValid = false;
  wrapped->onRefine(parent, newEntities);
}
void GpuSizeField::onCavity(ma::EntityArray& oldElements, ma::EntityArray& newEntities)
{
  dirty = true; streak = 0; topoValid = false;
  wrapped->onCavity(oldElements, newEntities);
}
int GpuSizeField::getTransferDimension() { return wrapped->getTransferDimension(); }
bool GpuSizeField::hasNodesOn(int dimension) { return wrapped->hasNodesOn(dimension); }

/* ------------------------------------------------------------------ construction */
GpuSizeField* Maker::make(ma::Mesh* m, ma::SizeField* wrapped, int kind, int logVariant, apf::Field* sizes,
                          apf::Field* frames, apf::Field* iso, ma::AnisotropicFunction* fa, ma::IsotropicFunction* fi,
                          int device)
{
  GpuSizeField* g = Access::create();
  g->mesh = m;
  g->wrapped = wrapped;
  Access::init(g, kind, logVariant, sizes, frames, iso, fa, fi);
  mag_ctx* c = 0;
  int rc = mag_create(&c, device);
  if (rc) fail(0, "mag_create", rc);
  g->ctx = c;
  return g;
}

GpuSizeField* makeSizeField(ma::Mesh* m, apf::Field* sizes, apf::Field* frames, bool logInterpolation, int device)
{
  /* LogAnisoSizeField::init builds ma_logM from the two fields (maSize.cc:491-499): variant 0 */
  return Maker::make(m, ma::makeSizeField(m, sizes, frames, logInterpolation), logInterpolation ? 3 : 2, 0, sizes, frames, 0, 0, 0, device);
}
GpuSizeField* makeSizeField(ma::Mesh* m, ma::AnisotropicFunction* f, bool logInterpolation, int device)
{
  /* LogMEval evaluates -2 log(h) per vertex (maSize.cc:343-346): variant 1 */
  return Maker::make(m, ma::makeSizeField(m, f, logInterpolation), logInterpolation ? 3 : 2, 1, 0, 0, 0, f, 0, device);
}
GpuSizeField* makeSizeField(ma::Mesh* m, apf::Field* size, int device)
{
  return Maker::make(m, ma::makeSizeField(m, size), 1, 0, 0, 0, size, 0, 0, device);
}
GpuSizeField* makeSizeField(ma::Mesh* m, ma::IsotropicFunction* f, int device)
{
  return Maker::make(m, ma::makeSizeField(m, f), 1, 0, 0, 0, 0, 0, f, device);
}

/* ------------------------------------------------------------------ bulk sweeps */
static GpuSizeField* gpuField(ma::SizeField* sf)
{
  GpuSizeField* g = dynamic_cast<GpuSizeField*>(sf);
  if (!g) { fprintf(stderr, "mag adapter: the size field is not a mag::GpuSizeField\n"); abort(); }
  return g;
}

long markEdgesToSplit(ma::Adapt* a)
{
  GpuSizeField* g = gpuField(a->sizeField);
  mag_stats st = Access::sweepWithAdaptFlags(g, a, MAG_OP_MARK_SPLIT);
  return g->multiPart() ? (long)st.n_split : a->mesh->getPCU()->Add<long>((long)st.n_split); /* maAdapt.cc:323 */
}
long markEdgesToCollapse(ma::Adapt* a)
{
  GpuSizeField* g = gpuField(a->sizeField);
  mag_stats st = Access::sweepWithAdaptFlags(g, a, MAG_OP_MARK_COLLAPSE);
  return g->multiPart() ? (long)st.n_collapse : a->mesh->getPCU()->Add<long>((long)st.n_collapse);
}
int markBadQuality(ma::Adapt* a)
{
  GpuSizeField* g = gpuField(a->sizeField);
  mag_stats st = Access::sweepWithAdaptFlags(g, a, MAG_OP_MARK_BAD);
  return (int)(g->multiPart() ? (long)st.n_bad : a->mesh->getPCU()->Add<long>((long)st.n_bad));
}
double getMinQuality(ma::Adapt* a)
{
  GpuSizeField* g = gpuField(a->sizeField);
  mag_stats st = Access::sweepWithAdaptFlags(g, a, MAG_OP_QUALITIES);
  return g->multiPart() ? st.min_quality : a->mesh->getPCU()->Min<double>(st.min_quality); /* maShape.cc:168 */
}
double getMaximumEdgeLength(ma::Mesh* m, ma::SizeField* sf)
{
  GpuSizeField* g = gpuField(sf);
  g->refresh(-1);
  Access::resetOrder(g);
  mag_stats st;
  if (g->multiPart()) { MAG_DO(g->ctx, mag_allreduce_stats(g->ctx, &st)); return st.max_length; }
  MAG_DO(g->ctx, mag_get_stats(g->ctx, &st));
  return m->getPCU()->Max<double>(st.max_length); /* maSize.cc:689 */
}
void getEdgeLengthsInMetricSpace(ma::Mesh* m, ma::SizeField* sf, std::vector<double>& out)
{
  GpuSizeField* g = gpuField(sf);
  g->refresh(-1);
  Access::resetOrder(g);
  /* ma::stats keeps owned edges only, iteration order (maStats.cc:33-45) */
  out.clear();
  apf::MeshIterator* it = m->begin(1);
  ma::Entity* e;
  size_t k = 0;
  while ((e = m->iterate(it))) { if (m->isOwned(e)) out.push_back(Access::lengthAt(g, k)); ++k; }
  m->end(it);
}
void getLinearQualitiesInMetricSpace(ma::Mesh* m, ma::SizeField* sf, std::vector<double>& out)
{
  GpuSizeField* g = gpuField(sf);
  g->refresh(-1);
  Access::resetOrder(g);
  /* owned simplex elements, cbrt of the mean ratio cubed (maStats.cc:12-31) */
  out.clear();
  const int dim = m->getDimension();
  apf::MeshIterator* it = m->begin(dim);
  ma::Entity* e;
  while ((e = m->iterate(it))) {
    if (m->getType(e) != (dim == 3 ? apf::Mesh::TET : apf::Mesh::TRIANGLE) || !m->isOwned(e)) continue;
    const double lq = Access::qualityOf(g, e);
    out.push_back(dim == 2 ? ((lq > 0) ? sqrt(lq) : -sqrt(-lq)) : cbrt(lq));   /* maStats.cc:24-28 */
  }
  m->end(it);
}

void stats(ma::Mesh* m, ma::SizeField* sf, std::vector<double>& edgeLengths, std::vector<double>& linearQualities, bool inMetric)
{
  edgeLengths.clear();
  linearQualities.clear();
  if (!inMetric) { ma::stats(m, sf, edgeLengths, linearQualities, false); return; }
  GpuSizeField* g = gpuField(sf);
  g->refresh(-1);                                   /* one export + one sweep serve both vectors */
  Access::resetOrder(g);
  const int dim = m->getDimension();
  apf::MeshIterator* it = m->begin(dim);            /* qualities first, as getStatsInMetricSpace does (maStats.cc:95-103) */
  ma::Entity* e;
  while ((e = m->iterate(it))) {
    if (m->getType(e) != (dim == 3 ? apf::Mesh::TET : apf::Mesh::TRIANGLE) || !m->isOwned(e)) continue;
    const double lq = Access::qualityOf(g, e);
    linearQualities.push_back(dim == 2 ? ((lq > 0) ? sqrt(lq) : -sqrt(-lq)) : cbrt(lq));
  }
  m->end(it);
  it = m->begin(1);
  size_t k = 0;
  while ((e = m->iterate(it))) { if (m->isOwned(e)) edgeLengths.push_back(Access::lengthAt(g, k)); ++k; }
  m->end(it);
}

/* ma::getElementWeights (maBalance.cc:83-97): the tag "ma_weight" on every element.  Tets come from one device sweep
   (mag_element_weights, clamped like clampForIterations with the Adapt's iteration counters); layer elements go
   through the reference's own per-entity path (their weight is a base-triangle measure, maBalance.cc:31-37). */
ma::Tag* getElementWeights(ma::Adapt* a)
{
  GpuSizeField* g = gpuField(a->sizeField);
  ma::Mesh* m = a->mesh;
  g->refresh(a->input->goodQuality);        /* re-validates (or redoes) the export: never a stale device copy */
  Access::resetOrder(g);
  const int dim = m->getDimension();
  if (Access::exportedElems(g) != m->count(dim)) { fprintf(stderr, "mag adapter: export out of step with the mesh\n"); abort(); }
  const double w_max = pow(2.0, dim * (a->refinesLeft)), w_min = pow(4.0, -(a->coarsensLeft));
  std::vector<double> w(m->count(dim));
  MAG_DO(g->ctx, mag_element_weights(g->ctx, w_max, w_min, Access::fpMode(g), w.data()));
  ma::Tag* weights = m->createDoubleTag("ma_weight", 1);
  apf::MeshIterator* it = m->begin(dim);
  ma::Entity* e;
  while ((e = m->iterate(it))) {
    const int t = m->getType(e);
    double weight = (t == apf::Mesh::TET || (dim == 2 && t == apf::Mesh::TRIANGLE)) ? w[Access::tetSlotOf(g, e)] : ma::getElementWeight(a, e);
    m->setDoubleTag(e, weights, &weight);
  }
  m->end(it);
  return weights;
}

/* ma::getSliverCode / matchSliver (maShape.cc:35-120) for every tet of the mesh in one device sweep */
void getSliverCodes(ma::Adapt* a, std::vector<int>& codes, std::vector<ma::CodeMatch>& matches)
{
  GpuSizeField* g = gpuField(a->sizeField);
  ma::Mesh* m = a->mesh;
  g->refresh(a->input->goodQuality);
  Access::resetOrder(g);
  if (Access::exportedElems(g) != m->count(3)) { fprintf(stderr, "mag adapter: export out of step with the mesh\n"); abort(); }
  const size_t nel = m->count(3), nns = (size_t)Access::nonSimplex(g);
  std::vector<int> face0(3 * (nel - nns)), match(2 * nel);
  codes.assign(nel, 0);
  apf::MeshIterator* it = m->begin(3);
  ma::Entity* e;
  while ((e = m->iterate(it))) {
    if (m->getType(e) != apf::Mesh::TET) continue;
    const size_t t = (size_t)Access::tetSlotOf(g, e) - nns;
    apf::Downward fs, fv;
    m->getDownward(e, 2, fs);
    m->getDownward(fs[0], 0, fv);
    for (int i = 0; i < 3; ++i) face0[3 * t + i] = Access::vertSlotOf(g, fv[i]);
  }
  m->end(it);
  MAG_DO(g->ctx, mag_sliver_codes(g->ctx, face0.data(), a->input->goodQuality, 0, codes.data(), match.data()));
  matches.resize(nel);
  for (size_t i = 0; i < nel; ++i) { matches[i].rotation = match[2 * i]; matches[i].code_index = match[2 * i + 1]; }
}

/* ------------------------------------------------------------------ shape handler */
class GpuShapeHandler : public ma::ShapeHandler
{
  public:
    GpuShapeHandler(ma::Adapt* a_) : a(a_), inner(ma::getShapeHandler(a_)) {}
    ~GpuShapeHandler() { delete inner; }
    double getQuality(ma::Entity* e)
    {
      GpuSizeField* g = dynamic_cast<GpuSizeField*>(a->sizeField);
      double q;
      const int t = a->mesh->getType(e);
      if (g && (t == apf::Mesh::TET || (t == apf::Mesh::TRIANGLE && a->mesh->getDimension() == 2))) {
        if (Access::serveQuality(g, e, a->input->goodQuality, q)) return q;
        /* LinearHandler::getQuality (maShapeHandler.cc:27-30) on the wrapped reference field */
        if (a->mesh->getShape()->getOrder() == 1) return ma::measureElementQuality(a->mesh, g->wrapped, e);
      }
      return inner->getQuality(e);
    }
    /* SolutionTransfer interface: the linear handler's behaviour, unchanged */
    bool hasNodesOn(int dimension) { return inner->hasNodesOn(dimension); }
    void onVertex(apf::MeshElement* parent, ma::Vector const& xi, ma::Entity* vert) { inner->onVertex(parent, xi, vert); }
    void onRefine(ma::Entity* parent, ma::EntityArray& newEntities) { inner->onRefine(parent, newEntities); }
    void onCavity(ma::EntityArray& oldElements, ma::EntityArray& newEntities) { inner->onCavity(oldElements, newEntities); }
    int getTransferDimension() { return inner->getTransferDimension(); }
  private:
    ma::Adapt* a;
    ma::ShapeHandler* inner;
};

ma::ShapeHandler* shapeHandler(ma::Adapt* a) { return new GpuShapeHandler(a); }

}
