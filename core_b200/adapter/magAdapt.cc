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
#include <typeinfo>
#include <atomic>
#include <chrono>
#include <cstring>
#include <string>
#include <mds_apf.h>   /* struct mds_apf / struct mds / struct mds_tag: MDS's own arrays (the direct export route) */

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

/* ------------------------------------------------------------------ MDS's own arrays (the direct export route)
   apf::Mesh2's public interface answers one entity per virtual call (getDownward of a tet walks faces -> edges -> vertices with
   set intersections: 200 ns; 30 ns per tag access).  MDS itself is a handful of flat arrays (mds/mds.h:32-43: one-level-down ids
   per type, a free list whose MDS_LIVE entries are the live entities; mds/mds_apf.h:31-43: point[][3]; mds/mds_tag.h:16-23: one
   value array + one presence bitmap per tag and type), so the adapter reads those when the mesh is an MDS mesh.  The class that
   owns the struct (apf::MeshMDS) is local to mds/apfMDS.cc; its first own member, right after the apf::Mesh2 base, is the
   `mds_apf* mesh` pointer (mds/apfMDS.cc:774).  directMds() takes it from there and believes it only after the dynamic type,
   the dimension, the model pointer, every entity count and the first vertex's coordinates agree with what the public interface
   says; otherwise (another Mesh2 implementation, a layout this was not compiled against) the public-interface walk is used.
   An in-tree build would replace the lookup by a one-line accessor next to apf::getMdsIndex. */
static double nowSeconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static Profile g_profile = {0, 0, 0, 0, 0, 0, 0};
Profile& profile() { return g_profile; }
struct Lap {   /* adds the time between construction and destruction to one field of the profile */
  double& acc; double t0;
  explicit Lap(double& a) : acc(a), t0(nowSeconds()) {}
  ~Lap() { acc += nowSeconds() - t0; }
};
static bool g_directMds = true;
static unsigned long long g_generation = 0;   /* stamps the vertex data of an export (Export::vertGen) */
void setDirectMds(bool on) { g_directMds = on; }

static inline ma::Entity* toEnt(mds_id id) { return reinterpret_cast<ma::Entity*>(((char*)1) + id); }   /* mds/apfMDS.cc:98-102 */
static inline mds_id idOf(int type, mds_id index) { return index * MDS_TYPES + type; }                  /* mds/mds.c:291-313 */
static inline mds_id indexOf(mds_id id) { return id / MDS_TYPES; }
static inline int typeOf(mds_id id) { return id % MDS_TYPES; }

static mds_apf* directMds(ma::Mesh* m)
{
  if (!g_directMds || getenv("MAG_ADAPTER_PUBLIC_API_ONLY")) return 0;
  if (!strstr(typeid(*m).name(), "MeshMDS")) return 0;
  mds_apf* M = *reinterpret_cast<mds_apf**>(reinterpret_cast<char*>(m) + sizeof(apf::Mesh2));
  if (!M) return 0;
  const int d = m->getDimension();
  if (M->mds.d != d || M->user_model != m->getModel()) return 0;
  long cnt[4] = {0, 0, 0, 0};
  for (int t = 0; t < MDS_TYPES; ++t) cnt[mds_dim[t]] += M->mds.n[t];
  for (int k = 0; k <= d; ++k) if ((size_t)cnt[k] != m->count(k)) return 0;
  apf::MeshIterator* it = m->begin(0);
  ma::Entity* e = m->iterate(it);
  m->end(it);
  if (e) {
    ma::Vector p;
    m->getPoint(e, 0, p);
    const double* q = M->point[apf::getMdsIndex(m, e)];
    if (p[0] != q[0] || p[1] != q[1] || p[2] != q[2]) return 0;
  }
  return M;
}

/* the tag apf keeps the vertex nodes of field f in (apf/apfTagData.cc:60-80: "<name>_ver"); NULL when the field is not tag-backed
   (frozen into an array, apf/apfArrayData.cc) or not laid out as comps doubles per vertex */
static const mds_tag* fieldVertexTag(ma::Mesh* m, apf::Field* f, int comps)
{
  if (!f) return 0;
  const std::string name = std::string(apf::getName(f)) + "_ver";
  const mds_tag* tg = reinterpret_cast<const mds_tag*>(m->findTag(name.c_str()));
  if (!tg || tg->user_type != apf::Mesh::DOUBLE || tg->bytes != comps * (int)sizeof(double)) return 0;
  if (!tg->data[MDS_VERTEX] || !tg->has[MDS_VERTEX]) return 0;
  return tg;
}
static inline bool tagHas(const unsigned char* has, mds_id i) { return (has[i / 8] >> (i % 8)) & 1; }   /* mds/mds_tag.c:105-119 */

/* mds/mds.c:73-136: entity i of dimension d-2 of an element is what its (d-1)-faces conv[2i], conv[2i+1] have in common */
static const int kT10[3][2] = {{2, 0}, {0, 1}, {1, 2}};
static const int kTet21[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 3}, {1, 2}, {2, 3}};
static const int kTet10[4][2] = {{2, 0}, {0, 1}, {1, 2}, {3, 4}};
static const int kW21[9][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 3}, {1, 2}, {2, 3}, {1, 4}, {2, 4}, {3, 4}};
static const int kW10[6][2] = {{0, 2}, {0, 1}, {1, 2}, {6, 8}, {6, 7}, {7, 8}};
static const int kP21[8][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {1, 4}, {1, 2}, {2, 3}, {3, 4}};
static const int kP10[5][2] = {{0, 3}, {0, 1}, {1, 2}, {2, 3}, {4, 5}};
/* mds/mds.c:496-508 common_down: the first entry of a that b holds too */
static inline mds_id commonId(const mds_id* a, int na, const mds_id* b, int nb)
{
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j)
      if (a[i] == b[j]) return a[i];
  return MDS_NONE;
}
/* the tet case of elementVertices written out (mds/mds.c T21 / T10, :80-90): four triangles -> six edges -> four vertices */
static inline bool tetVertices(const struct mds* s, mds_id index, const int* vslot, int* out)
{
  const mds_id* f = s->down[2][MDS_TETRAHEDRON] + 4 * (size_t)index;
  const mds_id* te = s->down[1][MDS_TRIANGLE];
  const mds_id* a = te + 3 * (size_t)indexOf(f[0]);
  const mds_id* b = te + 3 * (size_t)indexOf(f[1]);
  const mds_id* c = te + 3 * (size_t)indexOf(f[2]);
  const mds_id* d = te + 3 * (size_t)indexOf(f[3]);
  const mds_id e0 = commonId(a, 3, b, 3), e1 = commonId(a, 3, c, 3), e2 = commonId(a, 3, d, 3), e3 = commonId(b, 3, d, 3), e4 = commonId(b, 3, c, 3);
  if ((e0 | e1 | e2 | e3 | e4) < 0) return false;
  const mds_id* ev = s->down[0][MDS_EDGE];
  const mds_id *p0 = ev + 2 * (size_t)indexOf(e0), *p1 = ev + 2 * (size_t)indexOf(e1), *p2 = ev + 2 * (size_t)indexOf(e2),
               *p3 = ev + 2 * (size_t)indexOf(e3), *p4 = ev + 2 * (size_t)indexOf(e4);
  const mds_id v0 = commonId(p2, 2, p0, 2), v1 = commonId(p0, 2, p1, 2), v2 = commonId(p1, 2, p2, 2), v3 = commonId(p3, 2, p4, 2);
  if ((v0 | v1 | v2 | v3) < 0) return false;
  out[0] = vslot[indexOf(v0)]; out[1] = vslot[indexOf(v1)]; out[2] = vslot[indexOf(v2)]; out[3] = vslot[indexOf(v3)];
  return true;
}
/* downward vertices of element `index` of MDS type t (tet / wedge / pyramid) in getDownward(e, 0, .) order, as export vertex
   ids; false if the adjacency does not close (never on a valid mesh) */
static bool elementVertices(const struct mds* s, int t, mds_id index, const int (*p21)[2], int nedges, const int (*p10)[2], int nverts,
                            const int* vslot, int* out)
{
  const int nfaces = mds_degree[t][2];
  const mds_id* faces = s->down[2][t] + (size_t)index * nfaces;
  const mds_id* fe[5];
  int fk[5];
  for (int j = 0; j < nfaces; ++j) {
    const int ft = typeOf(faces[j]);
    fk[j] = mds_degree[ft][1];
    fe[j] = s->down[1][ft] + (size_t)indexOf(faces[j]) * fk[j];
  }
  mds_id el_e[9];
  for (int k = 0; k < nedges; ++k)
    if ((el_e[k] = commonId(fe[p21[k][0]], fk[p21[k][0]], fe[p21[k][1]], fk[p21[k][1]])) == MDS_NONE) return false;
  const mds_id* ev = s->down[0][MDS_EDGE];
  for (int k = 0; k < nverts; ++k) {
    const mds_id v = commonId(ev + 2 * (size_t)indexOf(el_e[p10[k][0]]), 2, ev + 2 * (size_t)indexOf(el_e[p10[k][1]]), 2);
    if (v == MDS_NONE) return false;
    out[k] = vslot[indexOf(v)];
  }
  return true;
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
  unsigned long long vertGen;             /* counts the changes of xyz / ma / mb (what the device holds of the vertices) */
  std::vector<double> xyz, ma, mb;
  std::vector<mds_id> vertId, edgeId, elemId;   /* direct route: MDS ids of the exported entities (empty otherwise) */
  mds_apf* mds;                           /* direct route: MDS's own struct; NULL = public-interface walk */
  bool vertsFresh;                        /* verticesUnchanged() found a change and already stored the new values */
  Export() : vertGen(0), mds(0), vertsFresh(false), logm_direct(false) {}
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
  /* the arithmetic of the calls beside the marking sweep: MAG_FP_FAST_LISTED is a mode of mag_sweep only */
  static int fpMode(GpuSizeField* g) { return g->fpMode == MAG_FP_FAST_LISTED ? MAG_FP_FAST : g->fpMode; }
  static int vertSlotOf(GpuSizeField* g, ma::Entity* v) { return g->vertSlot[apf::getMdsIndex(g->mesh, v)]; }
  static long nonSimplex(GpuSizeField* g) { return g->nNonSimplex; }
  static int tetSlotOf(GpuSizeField* g, ma::Entity* e) { return g->tetSlot[apf::getMdsIndex(g->mesh, e)]; }
  static int edgeSlotOf(GpuSizeField* g, ma::Entity* e)
  {
    const size_t id = (size_t)apf::getMdsIndex(g->mesh, e);
    return id < g->edgeSlot.size() ? g->edgeSlot[id] : -1;
  }
  static double qualityOf(GpuSizeField* g, ma::Entity* e) { return g->qualities[g->tetSlot[apf::getMdsIndex(g->mesh, e)]]; }
  static bool serveQuality(GpuSizeField* g, ma::Entity* e, double goodQuality, double& q)
  {
    int slot;
    g->lastGoodQuality = goodQuality;
    if (!g->serve(e, g->mesh->getDimension(), slot)) return false;
    q = g->qualities[slot];
    return true;
  }

  /* the per-vertex part of an export: coordinates and size-field values in m->begin(0) order, vertex ids compacted to 0..nv-1 */
  static void exportVertices(GpuSizeField* g, Export& x)
  {
    if (x.mds && exportVerticesDirect(g, x)) return;
    ma::Mesh* m = g->mesh;
    size_t nv = m->count(0);
    std::vector<int>& vslot = g->vertSlot;
    vslot.assign(vslot.size(), -1);
    x.vertId.clear();
    x.xyz.resize(3 * nv);
    /* LogAnisoSizeField built from fields keeps its own "ma_logM" vertex field and is the only thing it updates when
       vertices are created (interpolate / onRefine / onCavity, maSize.cc:523-561): the sizes and frames it was built from go
       stale for new vertices, so the export reads the logM field itself */
    apf::Field* logM = (g->kind == 3 && !g->fnAniso) ? m->findField("ma_logM") : 0;
    if (g->kind == 3 && !g->fnAniso && !logM) { fprintf(stderr, "mag adapter: the wrapped LogAnisoSizeField has no ma_logM field\n"); abort(); }
    x.logm_direct = logM != 0;
    if (g->kind == 1) { x.ma.resize(nv); x.mb.clear(); }
    else if (logM) { x.ma.clear(); x.mb.resize(9 * nv); }
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
      fieldAt(g, logM, e, k, x);
      ++k;
    }
    m->end(it);
  }
  /* size-field values of one vertex through the public interface (user functions, frozen fields) */
  static void fieldAt(GpuSizeField* g, apf::Field* logM, ma::Entity* e, size_t k, Export& x)
  {
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
  }
  /* the same arrays from MDS's own: live vertices in index order (= m->begin(0) order, mds.c:745-777), coordinates out of
     point[][3], tag-backed fields out of their tag arrays (row-major components, apf/apfFieldData.cc) */
  static bool exportVerticesDirect(GpuSizeField* g, Export& x)
  {
    ma::Mesh* m = g->mesh;
    const mds_apf* M = x.mds;
    const mds_id endv = M->mds.end[MDS_VERTEX];
    const mds_id* live = M->mds.free[MDS_VERTEX];
    const size_t nv = (size_t)M->mds.n[MDS_VERTEX];
    apf::Field* logM = (g->kind == 3 && !g->fnAniso) ? m->findField("ma_logM") : 0;
    if (g->kind == 3 && !g->fnAniso && !logM) { fprintf(stderr, "mag adapter: the wrapped LogAnisoSizeField has no ma_logM field\n"); abort(); }
    x.logm_direct = logM != 0;
    const mds_tag *ta = 0, *tb = 0;         /* tags behind x.ma / x.mb; NULL: per-vertex calls */
    if (g->kind == 1) { if (!g->fnIso) ta = fieldVertexTag(m, g->fIso, 1); }
    else if (logM) tb = fieldVertexTag(m, logM, 9);
    else if (!g->fnAniso) {
      ta = fieldVertexTag(m, g->fSizes, 3); tb = fieldVertexTag(m, g->fFrames, 9);
      if (!ta || !tb) ta = tb = 0;
    }
    std::vector<int>& vslot = g->vertSlot;
    vslot.assign(std::max(vslot.size(), (size_t)endv), -1);
    x.vertId.resize(nv);
    x.xyz.resize(3 * nv);
    if (g->kind == 1) { x.ma.resize(nv); x.mb.clear(); }
    else if (logM) { x.ma.clear(); x.mb.resize(9 * nv); }
    else { x.ma.resize(3 * nv); x.mb.resize(9 * nv); }
    const int ca = g->kind == 1 ? 1 : 3;
    size_t k = 0;
    for (mds_id i = 0; i < endv; ++i) {
      if (live[i] != MDS_LIVE) continue;
      if (k >= nv) return false;
      vslot[i] = (int)k;
      x.vertId[k] = idOf(MDS_VERTEX, i);
      memcpy(&x.xyz[3 * k], M->point[i], 3 * sizeof(double));
      if (ta) {
        if (!tagHas(ta->has[MDS_VERTEX], i)) return false;
        memcpy(&x.ma[ca * k], ta->data[MDS_VERTEX] + (size_t)ta->bytes * i, ta->bytes);
      }
      if (tb) {
        if (!tagHas(tb->has[MDS_VERTEX], i)) return false;
        memcpy(&x.mb[9 * k], tb->data[MDS_VERTEX] + (size_t)tb->bytes * i, tb->bytes);
      }
      if (!ta && !tb) fieldAt(g, logM, toEnt(idOf(MDS_VERTEX, i)), k, x);
      ++k;
    }
    return k == nv;
  }
  /* are the coordinates and field values of x still what the mesh holds?  (snapping, ma/ma.cc:37, and in-place edits of the
     size fields reach no size-field callback.)  The direct route compares in place; the public route re-reads into a copy. */
  static bool verticesUnchanged(GpuSizeField* g, Export& x)
  {
    ma::Mesh* m = g->mesh;
    if (x.mds) {
      const mds_apf* M = x.mds;
      if ((size_t)M->mds.n[MDS_VERTEX] != x.vertId.size()) return false;
      apf::Field* logM = (g->kind == 3 && !g->fnAniso) ? m->findField("ma_logM") : 0;
      const mds_tag *ta = 0, *tb = 0;
      if (g->kind == 1) { if (!g->fnIso) ta = fieldVertexTag(m, g->fIso, 1); }
      else if (x.logm_direct) tb = fieldVertexTag(m, logM, 9);
      else if (!g->fnAniso) { ta = fieldVertexTag(m, g->fSizes, 3); tb = fieldVertexTag(m, g->fFrames, 9); if (!ta || !tb) ta = tb = 0; }
      if (ta || tb) {
        const mds_id* live = M->mds.free[MDS_VERTEX];
        const int ca = g->kind == 1 ? 1 : 3;
        const size_t nv = x.vertId.size();
        for (size_t k = 0; k < nv; ++k) {
          const mds_id i = indexOf(x.vertId[k]);
          if (i >= M->mds.end[MDS_VERTEX] || live[i] != MDS_LIVE) return false;
          if (memcmp(&x.xyz[3 * k], M->point[i], 3 * sizeof(double))) return false;
          if (ta && (!tagHas(ta->has[MDS_VERTEX], i) || memcmp(&x.ma[ca * k], ta->data[MDS_VERTEX] + (size_t)ta->bytes * i, ta->bytes))) return false;
          if (tb && (!tagHas(tb->has[MDS_VERTEX], i) || memcmp(&x.mb[9 * k], tb->data[MDS_VERTEX] + (size_t)tb->bytes * i, tb->bytes))) return false;
        }
        return true;
      }
    }
    Export y;
    y.mds = x.mds;
    exportVertices(g, y);                     /* rewrites the slot table: identical when nothing changed */
    if (y.xyz == x.xyz && y.ma == x.ma && y.mb == x.mb && y.logm_direct == x.logm_direct) return true;
    x.xyz.swap(y.xyz); x.ma.swap(y.ma); x.mb.swap(y.mb); x.vertId.swap(y.vertId); x.logm_direct = y.logm_direct;
    x.vertsFresh = true;
    return false;
  }

  /* one pass over the mesh in m->begin(d) order (mds.c:745-777) */
  static void exportMesh(GpuSizeField* g, Export& x)
  {
    ma::Mesh* m = g->mesh;
    const int mdim = m->getDimension();
    if (mdim != 3 && mdim != 2) { fprintf(stderr, "mag adapter: only 2D and 3D meshes are supported\n"); abort(); }
    x.mds = directMds(m);
    if (x.mds && exportMeshDirect(g, x)) return;
    x.mds = 0;
    x.edgeId.clear(); x.elemId.clear();
    x.tet_v.clear(); x.tri_v.clear();
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

  /* the same export out of MDS's own arrays: live entities in index order per type; edge -> vertices is the one-level-down
     array itself, element -> vertices is derived from the one-level-down arrays with MDS's own rule (elementVertices) */
  static bool exportMeshDirect(GpuSizeField* g, Export& x)
  {
    ma::Mesh* m = g->mesh;
    const mds_apf* M = x.mds;
    const struct mds* s = &M->mds;
    const int mdim = m->getDimension();
    if (!exportVerticesDirect(g, x)) return false;
    const int* vslot = g->vertSlot.data();
    /* element types this path handles; anything else aborts like the walk does */
    const int bad3[] = {MDS_HEXAHEDRON}, bad2[] = {MDS_QUADRILATERAL};
    if (mdim == 3 && s->n[bad3[0]]) { fprintf(stderr, "mag adapter: element type %d is not supported\n", (int)apf::Mesh::HEX); abort(); }
    if (mdim == 2 && s->n[bad2[0]]) { fprintf(stderr, "mag adapter: element type %d is not supported\n", (int)apf::Mesh::QUAD); abort(); }
    const bool onePart = !m->getPCU() || m->getPCU()->Peers() <= 1;
    /* owned = not shared, or shared and m->isOwned (apfMDS.cc:288-298: the owner recorded in the entity's partition-model entity) */
    auto owned = [&](mds_id id) -> unsigned char {
      if (onePart || !mds_get_copies(const_cast<mds_net*>(&M->remotes), id)) return 1;
      return m->isOwned(toEnt(id)) ? 1 : 0;
    };
    /* live indices of one type, in order */
    auto liveOf = [&](int t, std::vector<mds_id>& ids) {
      const mds_id* fr = s->free[t];
      const mds_id end = s->end[t];
      ids.reserve(ids.size() + (size_t)s->n[t]);
      for (mds_id i = 0; i < end; ++i) if (fr[i] == MDS_LIVE) ids.push_back(idOf(t, i));
    };
    x.edgeId.clear();
    liveOf(MDS_EDGE, x.edgeId);
    const size_t ne = x.edgeId.size();
    if (ne != m->count(1)) return false;
    x.edge_v.resize(2 * ne); x.edge_owned.resize(ne); x.edges.resize(ne);
    g->edgeSlot.assign(std::max(g->edgeSlot.size(), (size_t)s->end[MDS_EDGE]), -1);
    const mds_id* ev = s->down[0][MDS_EDGE];
    for (size_t k = 0; k < ne; ++k) {
      const mds_id id = x.edgeId[k], i = indexOf(id);
      x.edges[k] = toEnt(id);
      g->edgeSlot[i] = (int)k;
      x.edge_v[2 * k] = vslot[indexOf(ev[2 * (size_t)i])];
      x.edge_v[2 * k + 1] = vslot[indexOf(ev[2 * (size_t)i + 1])];
      x.edge_owned[k] = owned(id);
    }
    x.elemId.clear();
    x.prism_v.clear(); x.pyr_v.clear(); x.tet_v.clear(); x.tri_v.clear();
    bool ok = true;
    size_t nsimplex = 0, first_simplex = 0;
    if (mdim == 3) {
      liveOf(MDS_WEDGE, x.elemId);
      const size_t np = x.elemId.size();
      liveOf(MDS_PYRAMID, x.elemId);
      const size_t npy = x.elemId.size() - np;
      liveOf(MDS_TETRAHEDRON, x.elemId);
      const size_t nt = x.elemId.size() - np - npy;
      g->nNonSimplex = (long)(np + npy);
      x.prism_v.resize(6 * np); x.pyr_v.resize(5 * npy); x.tet_v.resize(4 * nt);
      const mds_id* ids = x.elemId.data();
      for (size_t k = 0; k < np && ok; ++k) ok = elementVertices(s, MDS_WEDGE, indexOf(ids[k]), kW21, 9, kW10, 6, vslot, &x.prism_v[6 * k]);
      for (size_t k = 0; k < npy && ok; ++k) ok = elementVertices(s, MDS_PYRAMID, indexOf(ids[np + k]), kP21, 8, kP10, 5, vslot, &x.pyr_v[5 * k]);
      std::atomic<int> open_tets(0);
      parallelFor(nt, g->exportThreads, [&](size_t b, size_t en) {
        bool good = true;
        for (size_t k = b; k < en; ++k) good &= tetVertices(s, indexOf(ids[np + npy + k]), vslot, &x.tet_v[4 * k]);
        if (!good) open_tets.fetch_add(1);
      });
      ok = ok && open_tets.load() == 0;
      nsimplex = nt; first_simplex = np + npy;
    } else {
      liveOf(MDS_TRIANGLE, x.elemId);
      const size_t ntri = x.elemId.size();
      g->nNonSimplex = 0;
      x.tri_v.resize(3 * ntri);
      const mds_id* te = s->down[1][MDS_TRIANGLE];
      for (size_t k = 0; k < ntri && ok; ++k) {
        const mds_id* e3 = te + 3 * (size_t)indexOf(x.elemId[k]);
        for (int j = 0; j < 3; ++j) {
          const mds_id v = commonId(ev + 2 * (size_t)indexOf(e3[kT10[j][0]]), 2, ev + 2 * (size_t)indexOf(e3[kT10[j][1]]), 2);
          if (v == MDS_NONE) { ok = false; break; }
          x.tri_v[3 * k + j] = vslot[indexOf(v)];
        }
      }
      nsimplex = ntri; first_simplex = 0;
    }
    if (!ok || x.elemId.size() != m->count(mdim)) return false;
    const size_t nel = x.elemId.size();
    x.elems.resize(nel); x.elem_owned.resize(nel);
    for (size_t k = 0; k < nel; ++k) { x.elems[k] = toEnt(x.elemId[k]); x.elem_owned[k] = owned(x.elemId[k]); }
    const int st = mdim == 3 ? MDS_TETRAHEDRON : MDS_TRIANGLE;
    /* keyed like the walk's table, by apf::getMdsIndex (apfMDS.cc:1259-1271): live entities of the lower types of the
       dimension + index within the type */
    size_t key0 = 0;
    for (int t = 0; t < st; ++t) if (mds_dim[t] == mdim) key0 += (size_t)s->n[t];
    g->tetSlot.assign(std::max(g->tetSlot.size(), key0 + (size_t)s->end[st]), -1);
    for (size_t k = 0; k < nsimplex; ++k) g->tetSlot[key0 + (size_t)indexOf(x.elemId[first_simplex + k])] = (int)(first_simplex + k);
    return true;
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
      bool same;
      { Lap lap(g_profile.revalidate_s); same = verticesUnchanged(g, x); }
      if (!same) {
        Lap lap(g_profile.upload_s);
        if (!x.vertsFresh) exportVertices(g, x);
        x.vertsFresh = false;
        x.vertGen = ++g_generation;
        MAG_DO(g->ctx, mag_set_coords(g->ctx, x.xyz.data()));
        uploadField(g, x);
        g->dirty = true;
      }
      return x;
    }
    if (!g->exported) g->exported = new Export;
    g->exported->vertsFresh = false;        /* the arrays are overwritten in place: a re-export during an adapt allocates nothing */
    { Lap lap(g_profile.export_s); exportMesh(g, *g->exported); }
    g->exported->vertGen = ++g_generation;
    { Lap lap(g_profile.upload_s); upload(g, *g->exported); }
    g->topoValid = true;
    return *g->exported;
  }

  static void lengthsOnly(GpuSizeField* g)
  {
    ensureExported(g);
    mag_ctx* c = g->ctx;
    MAG_DO(c, mag_set_flags(c, 0, 0));
    MAG_DO(c, mag_sweep(c, MAG_OP_LENGTHS, ma::MAXLENGTH, ma::MINLENGTH, 0.0, 1, g->fpMode));
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
    double t_lap = nowSeconds();
    std::vector<int>& ef = g->flagScratch[0]; std::vector<int>& lf = g->flagScratch[1];     /* kept: no page faults per sweep */
    ef.resize(on_edges ? x.edges.size() : 0); lf.resize(on_elems ? x.elems.size() : 0);
    /* direct route: "ma_flags" is one int array + one presence bitmap per entity type (mds/mds_tag.h:16-23; ma::getFlags
       answers 0 for an entity without the tag, maAdapt.cc:80-88) */
    const mds_tag* ftag = x.mds ? reinterpret_cast<const mds_tag*>(a->flagsTag) : 0;
    if (ftag && (ftag->bytes != (int)sizeof(int) || ftag->user_type != apf::Mesh::INT)) ftag = 0;
    auto directWord = [&](mds_id id) -> int {
      const int t = typeOf(id);
      const mds_id i = indexOf(id);
      return (ftag->has[t] && tagHas(ftag->has[t], i)) ? reinterpret_cast<const int*>(ftag->data[t])[i] : 0;
    };
    if (ftag) {
      parallelFor(ef.size(), g->exportThreads, [&](size_t b, size_t en) { for (size_t i = b; i < en; ++i) ef[i] = directWord(x.edgeId[i]); });
      parallelFor(lf.size(), g->exportThreads, [&](size_t b, size_t en) { for (size_t i = b; i < en; ++i) lf[i] = directWord(x.elemId[i]); });
    } else {
    parallelFor(ef.size(), g->exportThreads, [&](size_t b, size_t en) {
      for (size_t i = b; i < en; ++i) ef[i] = readFlags(m, a->flagsTag, x.edges[i]);
    });
    parallelFor(lf.size(), g->exportThreads, [&](size_t b, size_t en) {
      for (size_t i = b; i < en; ++i) lf[i] = readFlags(m, a->flagsTag, x.elems[i]);
    });
    }
    g_profile.flags_in_s += nowSeconds() - t_lap; t_lap = nowSeconds();
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
      std::vector<int>& ef2 = g->flagScratch[2]; std::vector<int>& lf2 = g->flagScratch[3];
      ef2.resize(ef.size()); lf2.resize(lf.size());
      MAG_DO(c, mag_get_flags(c, on_edges ? ef2.data() : 0, on_elems ? lf2.data() : 0));
      g_profile.device_s += nowSeconds() - t_lap; t_lap = nowSeconds();
      /* a changed word goes back through ma::setFlags (setIntTag), or straight into the tag arrays: setIntTag is
         mds_give_tag + one store (apfMDS.cc:448-467), and mds_give_tag on a type whose arrays exist is the presence bit
         (mds_tag.c:121-138).  The first word of a type that has no arrays yet goes through setIntTag, which allocates them. */
      auto put = [&](ma::Entity* e, mds_id id, int word) {
        if (ftag) {
          const int t = typeOf(id);
          const mds_id i = indexOf(id);
          if (ftag->has[t]) {
            ftag->has[t][i / 8] |= (unsigned char)(1 << (i % 8));
            reinterpret_cast<int*>(ftag->data[t])[i] = word;
            return;
          }
        }
        ma::setFlags(a, e, word);
      };
      for (size_t i = 0; i < ef.size(); ++i) if (ef2[i] != ef[i]) put(x.edges[i], ftag ? x.edgeId[i] : 0, ef2[i]);
      for (size_t i = 0; i < lf.size(); ++i) if (lf2[i] != lf[i]) put(x.elems[i], ftag ? x.elemId[i] : 0, lf2[i]);
      g_profile.flags_out_s += nowSeconds() - t_lap;
    } else g_profile.device_s += nowSeconds() - t_lap;
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
      if (x.vertGen != snapshotHash) dirty = true;
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
  Lap lap(g_profile.refresh_s);
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
  snapshotHash = x.vertGen;
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
  Access::lengthsOnly(g);      /* one edge sweep, only the statistics come back (the per-entity snapshot is neither used nor replaced) */
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
  /* layer prisms: weighed by their base triangle (maBalance.cc:31-37), walked in the face's own vertex order; they are the
     first elements of the export (m->begin(3) order: prisms, pyramids, tets) */
  std::vector<int> base;
  apf::MeshIterator* it;
  ma::Entity* e;
  if (dim == 3) {
    it = m->begin(3);
    while ((e = m->iterate(it)) && m->getType(e) == apf::Mesh::PRISM) {
      apf::Downward fs, fv;
      m->getDownward(e, 2, fs);
      m->getDownward(fs[0], 0, fv);
      for (int i = 0; i < 3; ++i) base.push_back(Access::vertSlotOf(g, fv[i]));
    }
    m->end(it);
  }
  if (!base.empty())
    MAG_DO(g->ctx, mag_prism_weights(g->ctx, base.data(), w_max, w_min, a->input->shouldRefineLayer, a->input->shouldCoarsenLayer,
                                     a->input->shouldTurnLayerToTets, Access::fpMode(g), w.data()));
  const size_t nprism = base.size() / 3;
  ma::Tag* weights = m->createDoubleTag("ma_weight", 1);
  it = m->begin(dim);
  size_t k = 0;
  while ((e = m->iterate(it))) {
    const int t = m->getType(e);
    double weight;
    if (t == apf::Mesh::TET || (dim == 2 && t == apf::Mesh::TRIANGLE)) weight = w[Access::tetSlotOf(g, e)];
    else if (t == apf::Mesh::PRISM && k < nprism) weight = w[k];
    else weight = ma::getElementWeight(a, e);    /* pyramids: sparse by construction (maBalance.cc:27-29) */
    m->setDoubleTag(e, weights, &weight);
    ++k;
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

/* ma::Collapse's quality test (maCollapse.cc:88-113) for many candidates in one device call: see mag_collapse_quality */
void collapseQualities(ma::Adapt* a, const std::vector<ma::Entity*>& edges, const std::vector<ma::Entity*>& vertsToCollapse,
                       std::vector<double>& newWorst, std::vector<double>& oldWorst)
{
  GpuSizeField* g = gpuField(a->sizeField);
  ma::Mesh* m = a->mesh;
  if (edges.size() != vertsToCollapse.size()) { fprintf(stderr, "mag adapter: collapseQualities: one vertex per edge\n"); abort(); }
  Access::ensureExported(g);
  Access::resetOrder(g);
  const size_t n = edges.size();
  std::vector<int> slot(n);
  std::vector<unsigned char> end(n);
  for (size_t i = 0; i < n; ++i) {
    slot[i] = Access::edgeSlotOf(g, edges[i]);
    if (slot[i] < 0) { fprintf(stderr, "mag adapter: collapseQualities: edge %zu is not part of the export\n", i); abort(); }
    apf::Downward v;
    m->getDownward(edges[i], 0, v);
    if (v[0] != vertsToCollapse[i] && v[1] != vertsToCollapse[i]) { fprintf(stderr, "mag adapter: collapseQualities: vertex %zu is not an end of its edge\n", i); abort(); }
    end[i] = v[0] == vertsToCollapse[i] ? 0 : 1;
  }
  newWorst.resize(n);
  oldWorst.resize(n);
  MAG_DO(g->ctx, mag_collapse_quality(g->ctx, (int64_t)n, slot.data(), end.data(), 1, Access::fpMode(g), newWorst.data(), oldWorst.data(), 0));
}

/* ------------------------------------------------------------------ export self-check (host only, no device) */
int exportSelfCheck(ma::Mesh* m, apf::Field* sizes, apf::Field* frames, ma::Tag* flags, int threads, double* times)
{
  GpuSizeField* g[2];
  Export x[2];
  const bool was = g_directMds;
  for (int r = 0; r < 2; ++r) {
    g[r] = Access::create();
    g[r]->mesh = m;
    Access::init(g[r], 2, 0, sizes, frames, 0, 0, 0);
    g[r]->setExportThreads(threads);
    g_directMds = r == 1;
    double t0 = nowSeconds();
    Access::exportMesh(g[r], x[r]);
    times[r] = nowSeconds() - t0;
    t0 = nowSeconds();
    Access::exportMesh(g[r], x[r]);         /* again into the same buffers: what a re-export during an adapt costs */
    times[5 + r] = nowSeconds() - t0;
  }
  g_directMds = was;
  int bad = 0;
  if (!x[1].mds) bad |= 1;                                  /* the direct route was not taken */
  if (x[0].xyz != x[1].xyz || x[0].ma != x[1].ma || x[0].mb != x[1].mb) bad |= 2;
  if (x[0].edge_v != x[1].edge_v || x[0].edges != x[1].edges || x[0].edge_owned != x[1].edge_owned) bad |= 4;
  if (x[0].tet_v != x[1].tet_v || x[0].prism_v != x[1].prism_v || x[0].pyr_v != x[1].pyr_v || x[0].tri_v != x[1].tri_v) bad |= 8;
  if (x[0].elems != x[1].elems || x[0].elem_owned != x[1].elem_owned) bad |= 16;
  /* the slot tables agree wherever the walk filled them */
  for (int r = 0; r < 1; ++r) {
    for (size_t i = 0; i < g[0]->edgeSlot.size(); ++i) if (g[0]->edgeSlot[i] >= 0 && (i >= g[1]->edgeSlot.size() || g[1]->edgeSlot[i] != g[0]->edgeSlot[i])) bad |= 32;
    for (size_t i = 0; i < g[0]->tetSlot.size(); ++i) if (g[0]->tetSlot[i] >= 0 && (i >= g[1]->tetSlot.size() || g[1]->tetSlot[i] != g[0]->tetSlot[i])) bad |= 1024;
    for (size_t i = 0; i < g[0]->vertSlot.size(); ++i) if (g[0]->vertSlot[i] >= 0 && (i >= g[1]->vertSlot.size() || g[1]->vertSlot[i] != g[0]->vertSlot[i])) bad |= 2048;
  }
  if (g[0]->nNonSimplex != g[1]->nNonSimplex) bad |= 4096;
  /* unchanged vertices are recognised, a moved vertex and an edited field value are noticed (both routes) */
  for (int r = 0; r < 2; ++r) {
    g_directMds = r == 1;
    if (!Access::verticesUnchanged(g[r], x[r])) bad |= 64;
    apf::MeshIterator* it = m->begin(0);
    ma::Entity* v = m->iterate(it);
    m->end(it);
    ma::Vector p, q;
    m->getPoint(v, 0, p);
    q = p; q[0] += 0.125;
    m->setPoint(v, 0, q);
    if (Access::verticesUnchanged(g[r], x[r])) bad |= 128;
    m->setPoint(v, 0, p);
    x[r].vertsFresh = false;
    Access::exportVertices(g[r], x[r]);
    if (!Access::verticesUnchanged(g[r], x[r])) bad |= 64;
    ma::Vector h;
    apf::getVector(sizes, v, 0, h);
    ma::Vector h2 = h; h2[1] *= 2;
    apf::setVector(sizes, v, 0, h2);
    if (Access::verticesUnchanged(g[r], x[r])) bad |= 256;
    apf::setVector(sizes, v, 0, h);
    x[r].vertsFresh = false;
    Access::exportVertices(g[r], x[r]);
  }
  g_directMds = was;
  /* flag words: the tag arrays against ma-style getIntTag */
  if (flags && x[1].mds) {
    const mds_tag* ftag = reinterpret_cast<const mds_tag*>(flags);
    for (size_t i = 0; i < x[1].edgeId.size(); ++i) {
      const mds_id id = x[1].edgeId[i];
      const int t = typeOf(id);
      const int w = (ftag->has[t] && tagHas(ftag->has[t], indexOf(id))) ? reinterpret_cast<const int*>(ftag->data[t])[indexOf(id)] : 0;
      if (w != Access::readFlags(m, flags, x[0].edges[i])) { bad |= 512; break; }
    }
  }
  for (int r = 0; r < 2; ++r) { g[r]->mesh = 0; delete g[r]; }
  return bad;
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
