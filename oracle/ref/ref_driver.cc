/* ref_driver.cc -- TEST INFRASTRUCTURE (never linked into the product).
 *
 * A thin C interface over the UNMODIFIED SCOREC/core reference (built by
 * oracle/ref/Makefile into oracle/_ref/libscorec_ref.a) so tests, smoke() and
 * bench.py's cpu_baseline leg can run the reference's own implementation of the
 * hot path -- ma::SizeField::measure, ma::measureElementQuality,
 * ma::markEdgesToSplit / markEdgesToCollapse / markBadQuality, ma::getMinQuality
 * -- on a mesh + size field given as flat arrays, and dump the mesh in the
 * reference's own iteration order.  It calls only public/in-tree reference
 * entry points; it contains no restated arithmetic.
 *
 * Entity order everywhere: m->begin(d) iteration order; downward vertices in
 * m->getDownward(e,0,..) order; vertex ids = apf::getMdsIndex (creation order).
 */
#include <apf.h>
#include <apfMesh2.h>
#include <apfMDS.h>
#include <apfBox.h>
#include <apfShape.h>
#include <gmi_null.h>
#include <gmi_mesh.h>
#include <lionPrint.h>
#include <PCU.h>
#include <ma.h>
#include <maSize.h>
#include <maAdapt.h>
#include <maShape.h>
#include <maShapeHandler.h>
#include <maRefine.h>
#include <maStats.h>
#include <maLayer.h>
#include <cstdint>
#include <cstring>
#include <vector>
#include <chrono>

namespace ma {
/* external linkage in the reference but no header declares them
 * (maCoarsen.cc:287, maShape.cc:132,138,152) */
long markEdgesToCollapse(Adapt* a);
int markBadQuality(Adapt* a);
void unMarkBadQuality(Adapt* a);
double getMinQuality(Adapt* a);
/* maBalance.cc:74-81, external linkage, no header */
double getElementWeight(Adapt* a, Entity* e);
/* maShape.cc:35-120, external linkage; maShape.h:66 declares matchSliver with a Mesh* first argument that no definition has */
int getSliverCode(Adapt* a, Entity* tet);
CodeMatch matchSliver(Adapt* a, Entity* tet);
}

namespace {

pcu::PCU* g_pcu = 0;

void ensure_pcu()
{
  if (g_pcu) return;
  int argc = 0; char** argv = 0;
  pcu::Init(&argc, &argv);
  g_pcu = new pcu::PCU;
  lion_set_verbosity(0);
  gmi_register_null();
  gmi_register_mesh();
}

struct ArrayAniso : public ma::AnisotropicFunction {
  apf::Mesh2* m; const double* h; const double* R;
  void getValue(ma::Entity* v, ma::Matrix& r, ma::Vector& hh)
  {
    int i = apf::getMdsIndex(m, v);
    for (int a = 0; a < 3; ++a) hh[a] = h[3*i+a];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) r[a][b] = R[9*i+3*a+b];
  }
};
struct ArrayIso : public ma::IsotropicFunction {
  apf::Mesh2* m; const double* s;
  double getValue(ma::Entity* v) { return s[apf::getMdsIndex(m, v)]; }
};

struct Ref {
  apf::Mesh2* m;
  ma::SizeField* sf;      /* owned by us unless input owns it */
  ma::Input* in;
  ma::Adapt* a;
  apf::Field *f_iso, *f_h, *f_R;
  std::vector<double> fn_h, fn_R, fn_s;
  ArrayAniso fa; ArrayIso fi;
  Ref(): m(0), sf(0), in(0), a(0), f_iso(0), f_h(0), f_R(0) {}
};

void drop_sizefield(Ref* r)
{
  if (r->a) { delete r->a; r->a = 0; }
  if (r->in) { delete r->in; r->in = 0; }
  if (r->sf) { delete r->sf; r->sf = 0; }
  if (r->f_iso) { apf::destroyField(r->f_iso); r->f_iso = 0; }
  if (r->f_h) { apf::destroyField(r->f_h); r->f_h = 0; }
  if (r->f_R) { apf::destroyField(r->f_R); r->f_R = 0; }
}

double now()
{
  return std::chrono::duration<double>(
      std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

/* for ref_shape_shim.cc (same library) */
ma::Adapt* refo_internal_adapt(void* hd) { return hd ? ((Ref*)hd)->a : 0; }
apf::Mesh2* refo_internal_mesh(void* hd) { return hd ? ((Ref*)hd)->m : 0; }

extern "C" {

void* refo_box(int nx, int ny, int nz, double wx, double wy, double wz)
{
  ensure_pcu();
  Ref* r = new Ref;
  r->m = apf::makeMdsBox(nx, ny, nz, wx, wy, wz, true, g_pcu);
  return r;
}

/* general mesh from arrays through apf::buildElement (null model) */
void* refo_build(int64_t nv, const double* xyz,
                 int64_t ntet, const int32_t* tet_v,
                 int64_t nprism, const int32_t* prism_v,
                 int64_t npyr, const int32_t* pyr_v)
{
  ensure_pcu();
  Ref* r = new Ref;
  gmi_model* g = gmi_load(".null");
  apf::Mesh2* m = apf::makeEmptyMdsMesh(g, 3, false, g_pcu);
  std::vector<apf::MeshEntity*> v(nv);
  apf::ModelEntity* me = m->findModelEntity(3, 0);
  for (int64_t i = 0; i < nv; ++i) {
    v[i] = m->createVert(me);
    m->setPoint(v[i], 0, apf::Vector3(xyz[3*i], xyz[3*i+1], xyz[3*i+2]));
  }
  /* creation order: prisms, pyramids, tets -- each type has its own MDS
     index space, iteration order over dimension 3 is prisms, pyramids, tets
     regardless */
  apf::MeshEntity* ev[8];
  for (int64_t i = 0; i < nprism; ++i) {
    for (int j = 0; j < 6; ++j) ev[j] = v[prism_v[6*i+j]];
    apf::buildElement(m, me, apf::Mesh::PRISM, ev);
  }
  for (int64_t i = 0; i < npyr; ++i) {
    for (int j = 0; j < 5; ++j) ev[j] = v[pyr_v[5*i+j]];
    apf::buildElement(m, me, apf::Mesh::PYRAMID, ev);
  }
  for (int64_t i = 0; i < ntet; ++i) {
    for (int j = 0; j < 4; ++j) ev[j] = v[tet_v[4*i+j]];
    apf::buildElement(m, me, apf::Mesh::TET, ev);
  }
  m->acceptChanges();
  r->m = m;
  return r;
}

void* refo_load(const char* model, const char* smb)
{
  ensure_pcu();
  Ref* r = new Ref;
  r->m = apf::loadMdsMesh(model, smb, g_pcu);
  return r;
}

void refo_free(void* h)
{
  Ref* r = (Ref*)h;
  drop_sizefield(r);
  r->m->destroyNative();
  apf::destroyMesh(r->m);
  delete r;
}

/* out: nv, nedge, ntri, nquad, ntet, nhex, nprism, npyramid (apf type order) */
void refo_counts(void* h, int64_t* out)
{
  Ref* r = (Ref*)h;
  for (int t = 0; t < apf::Mesh::TYPES; ++t) out[t] = 0;
  for (int d = 0; d <= r->m->getDimension(); ++d) {
    apf::MeshIterator* it = r->m->begin(d);
    apf::MeshEntity* e;
    while ((e = r->m->iterate(it))) out[r->m->getType(e)]++;
    r->m->end(it);
  }
}

/* any pointer may be null.  elem_type[i] = apf type of the i-th dimension-3
   entity in iteration order; elem_v is padded to 8 ints per element. */
void refo_export(void* h, double* xyz, int32_t* edge_v, int32_t* elem_type,
                 int32_t* elem_v)
{
  Ref* r = (Ref*)h; apf::Mesh2* m = r->m;
  apf::MeshIterator* it; apf::MeshEntity* e;
  if (xyz) {
    it = m->begin(0);
    while ((e = m->iterate(it))) {
      apf::Vector3 p; m->getPoint(e, 0, p);
      int i = apf::getMdsIndex(m, e);
      xyz[3*i] = p[0]; xyz[3*i+1] = p[1]; xyz[3*i+2] = p[2];
    }
    m->end(it);
  }
  if (edge_v) {
    it = m->begin(1); int64_t k = 0;
    while ((e = m->iterate(it))) {
      apf::Downward dv; m->getDownward(e, 0, dv);
      edge_v[2*k] = apf::getMdsIndex(m, dv[0]);
      edge_v[2*k+1] = apf::getMdsIndex(m, dv[1]);
      ++k;
    }
    m->end(it);
  }
  if (elem_type || elem_v) {
    it = m->begin(m->getDimension()); int64_t k = 0;
    while ((e = m->iterate(it))) {
      apf::Downward dv; int n = m->getDownward(e, 0, dv);
      if (elem_type) elem_type[k] = m->getType(e);
      if (elem_v) {
        for (int j = 0; j < 8; ++j) elem_v[8*k+j] = -1;
        for (int j = 0; j < n; ++j) elem_v[8*k+j] = apf::getMdsIndex(m, dv[j]);
      }
      ++k;
    }
    m->end(it);
  }
}

void refo_set_coords(void* h, const double* xyz)
{
  Ref* r = (Ref*)h; apf::Mesh2* m = r->m;
  apf::MeshIterator* it = m->begin(0); apf::MeshEntity* e;
  while ((e = m->iterate(it))) {
    int i = apf::getMdsIndex(m, e);
    m->setPoint(e, 0, apf::Vector3(xyz[3*i], xyz[3*i+1], xyz[3*i+2]));
  }
  m->end(it);
}

/* ---- size fields.  kind: 0 identity, 1 iso scalar field (IsoUserField),
   2 sizes+frames fields linear (AnisoSizeField::init), 3 sizes+frames fields
   log (LogAnisoSizeField::init), 4 anisotropic user function linear,
   5 anisotropic user function log (LogMEval), 6 isotropic user function,
   7 ma::UniformRefiner (maSize.h:75-85: identity measure, shouldSplit constant true). */
int refo_set_sizefield(void* hd, int kind, const double* hs, const double* R)
{
  Ref* r = (Ref*)hd; apf::Mesh2* m = r->m;
  drop_sizefield(r);
  int64_t nv = m->count(0);
  apf::MeshIterator* it; apf::MeshEntity* e;
  if (kind == 0) {
    r->sf = new ma::IdentitySizeField(m);
  } else if (kind == 1) {
    r->f_iso = apf::createFieldOn(m, "refo_size", apf::SCALAR);
    it = m->begin(0);
    while ((e = m->iterate(it)))
      apf::setScalar(r->f_iso, e, 0, hs[apf::getMdsIndex(m, e)]);
    m->end(it);
    r->sf = ma::makeSizeField(m, r->f_iso);
  } else if (kind == 2 || kind == 3) {
    r->f_h = apf::createFieldOn(m, "refo_sizes", apf::VECTOR);
    r->f_R = apf::createFieldOn(m, "refo_frames", apf::MATRIX);
    it = m->begin(0);
    while ((e = m->iterate(it))) {
      int i = apf::getMdsIndex(m, e);
      apf::setVector(r->f_h, e, 0, apf::Vector3(hs + 3*i));
      apf::Matrix3x3 M;
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) M[a][b] = R[9*i+3*a+b];
      apf::setMatrix(r->f_R, e, 0, M);
    }
    m->end(it);
    r->sf = ma::makeSizeField(m, r->f_h, r->f_R, kind == 3);
    /* ~AnisoSizeField destroys hField/rField even when they came in through
       init() (maSize.cc:385-389); LogAnisoSizeField only owns its logM field */
    if (kind == 2) r->f_h = r->f_R = 0;
  } else if (kind == 4 || kind == 5) {
    r->fn_h.assign(hs, hs + 3*nv); r->fn_R.assign(R, R + 9*nv);
    r->fa.m = m; r->fa.h = r->fn_h.data(); r->fa.R = r->fn_R.data();
    r->sf = ma::makeSizeField(m, &r->fa, kind == 5);
  } else if (kind == 6) {
    r->fn_s.assign(hs, hs + nv);
    r->fi.m = m; r->fi.s = r->fn_s.data();
    r->sf = ma::makeSizeField(m, &r->fi);
  } else if (kind == 7) {
    r->sf = new ma::UniformRefiner(m);
  } else return 1;
  return 0;
}

/* the logM vertex field the reference built (kinds 3 and 5), row-major 9/vertex */
int refo_get_logm(void* hd, double* out)
{
  Ref* r = (Ref*)hd; apf::Mesh2* m = r->m;
  apf::Field* f = m->findField("ma_logM");
  if (!f) return 1;
  apf::MeshIterator* it = m->begin(0); apf::MeshEntity* e;
  while ((e = m->iterate(it))) {
    apf::Matrix3x3 M; apf::getMatrix(f, e, 0, M);
    int i = apf::getMdsIndex(m, e);
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) out[9*i+3*a+b] = M[a][b];
  }
  m->end(it);
  return 0;
}

/* ma::SizeField::measure on every edge, iteration order.  returns seconds */
double refo_lengths(void* hd, double* out)
{
  Ref* r = (Ref*)hd; apf::Mesh2* m = r->m;
  double t0 = now();
  apf::MeshIterator* it = m->begin(1); apf::MeshEntity* e; int64_t k = 0;
  while ((e = m->iterate(it))) out[k++] = r->sf->measure(e);
  m->end(it);
  return now() - t0;
}

/* ma::measureElementQuality on every simplex element of the mesh dimension;
   non-simplex entries get NaN-free sentinel 0 and are flagged by elem_type. */
double refo_qualities(void* hd, double* out, int useMax)
{
  Ref* r = (Ref*)hd; apf::Mesh2* m = r->m;
  double t0 = now();
  apf::MeshIterator* it = m->begin(m->getDimension()); apf::MeshEntity* e; int64_t k = 0;
  while ((e = m->iterate(it))) {
    if (apf::isSimplex(m->getType(e)))
      out[k] = ma::measureElementQuality(m, r->sf, e, useMax != 0);
    else out[k] = 0;
    ++k;
  }
  m->end(it);
  return now() - t0;
}

/* SizeField::getTransform at every vertex (xi=0), 9 doubles per vertex */
void refo_vertex_transforms(void* hd, double* out)
{
  Ref* r = (Ref*)hd; apf::Mesh2* m = r->m;
  apf::MeshIterator* it = m->begin(0); apf::MeshEntity* e;
  while ((e = m->iterate(it))) {
    apf::MeshElement* me = apf::createMeshElement(m, e);
    ma::Matrix Q; r->sf->getTransform(me, ma::Vector(0,0,0), Q);
    apf::destroyMeshElement(me);
    int i = apf::getMdsIndex(m, e);
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) out[9*i+3*a+b] = Q[a][b];
  }
  m->end(it);
}

/* SizeField::getTransform at an edge point xi (for unit checks) */
void refo_edge_transform(void* hd, int64_t edge, double xi, double* out)
{
  Ref* r = (Ref*)hd; apf::Mesh2* m = r->m;
  apf::MeshEntity* e = apf::getMdsEntity(m, 1, (int)edge);
  apf::MeshElement* me = apf::createMeshElement(m, e);
  ma::Matrix Q; r->sf->getTransform(me, ma::Vector(xi,0,0), Q);
  apf::destroyMeshElement(me);
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) out[3*a+b] = Q[a][b];
}

/* ---- marking through a real ma::Adapt ----
   which bit0: markEdgesToSplit, bit1: markEdgesToCollapse, bit2: markBadQuality,
   bit3: getMinQuality.  Incoming flag words (may be null = all zero) are OR-ed
   onto whatever Adapt's constructor set (LAYER closure flags, maLayer.cc:11-71).
   counts[0..2] = returned counts, minq = getMinQuality, times[0..3] seconds.
   edge/elem flags out = full "ma_flags" words after the marks. */
int refo_mark(void* hd, int which, double goodQuality,
              const int32_t* edge_flags_in, const int32_t* elem_flags_in,
              int32_t* edge_flags_out, int32_t* elem_flags_out,
              int64_t* counts, double* minq, double* times)
{
  Ref* r = (Ref*)hd; apf::Mesh2* m = r->m;
  if (!r->sf) return 1;
  if (r->a) { delete r->a; r->a = 0; }
  if (r->in) { delete r->in; r->in = 0; }
  /* configureIdentity(m, sf) = defaults + our size field, not owned
     (maInput.cc:287-304); avoids configure()'s extra getMaximumEdgeLength sweep */
  r->in = ma::makeAdvanced(ma::configureIdentity(m, r->sf));
  if (goodQuality >= 0) r->in->goodQuality = goodQuality;
  r->a = new ma::Adapt(r->in);
  ma::Adapt* a = r->a;
  apf::MeshIterator* it; apf::MeshEntity* e; int64_t k;
  if (edge_flags_in) {
    it = m->begin(1); k = 0;
    while ((e = m->iterate(it))) { if (edge_flags_in[k]) ma::setFlag(a, e, edge_flags_in[k]); ++k; }
    m->end(it);
  }
  if (elem_flags_in) {
    it = m->begin(m->getDimension()); k = 0;
    while ((e = m->iterate(it))) { if (elem_flags_in[k]) ma::setFlag(a, e, elem_flags_in[k]); ++k; }
    m->end(it);
  }
  for (int i = 0; i < 3; ++i) counts[i] = -1;
  for (int i = 0; i < 4; ++i) times[i] = 0;
  *minq = 0;
  double t;
  if (which & 1) { t = now(); counts[0] = ma::markEdgesToSplit(a); times[0] = now() - t; }
  if (which & 2) { t = now(); counts[1] = ma::markEdgesToCollapse(a); times[1] = now() - t; }
  if (which & 4) { t = now(); counts[2] = ma::markBadQuality(a); times[2] = now() - t; }
  if (which & 8) { t = now(); *minq = ma::getMinQuality(a); times[3] = now() - t; }
  if (edge_flags_out) {
    it = m->begin(1); k = 0;
    while ((e = m->iterate(it))) edge_flags_out[k++] = ma::getFlags(a, e);
    m->end(it);
  }
  if (elem_flags_out) {
    it = m->begin(m->getDimension()); k = 0;
    while ((e = m->iterate(it))) elem_flags_out[k++] = ma::getFlags(a, e);
    m->end(it);
  }
  return 0;
}

/* ma::isPrismOk / isPyramidOk per non-simplex element (maQuality.cc:490-560).
   ok[i] in {0,1}; codes[i] = prism good-diagonal mask or pyramid good
   rotation; simplex entries get ok=1, code=0. */
void refo_layer_ok(void* hd, int32_t* ok, int32_t* codes)
{
  Ref* r = (Ref*)hd; apf::Mesh2* m = r->m;
  apf::MeshIterator* it = m->begin(m->getDimension()); apf::MeshEntity* e; int64_t k = 0;
  while ((e = m->iterate(it))) {
    int code = 0; int good = 1; int t = m->getType(e);
    if (t == apf::Mesh::PRISM) good = ma::isPrismOk(m, e, &code) ? 1 : 0;
    else if (t == apf::Mesh::PYRAMID) good = ma::isPyramidOk(m, e, &code) ? 1 : 0;
    ok[k] = good; codes[k] = code; ++k;
  }
  m->end(it);
}

/* ma::getMaximumEdgeLength / getAverageEdgeLength (maSize.cc:654-691) */
double refo_max_edge_length(void* hd)
{
  Ref* r = (Ref*)hd;
  return ma::getMaximumEdgeLength(r->m, r->sf);
}
double refo_avg_edge_length(void* hd)
{
  Ref* r = (Ref*)hd;
  return ma::getAverageEdgeLength(r->m);
}

/* ---- predictive element weights (maBalance.cc:74-97): raw[i] = SizeField::getWeight(element i) and
   clamped[i] = ma::getElementWeight(a, element i) with the Adapt's refinesLeft / coarsensLeft set as given.
   Simplex elements only; others get 0.  Either output may be null. */
int refo_weights(void* hd, int refinesLeft, int coarsensLeft, double* raw, double* clamped)
{
  Ref* r = (Ref*)hd; apf::Mesh2* m = r->m;
  if (!r->sf) return 1;
  if (clamped) {
    if (r->a) { delete r->a; r->a = 0; }
    if (r->in) { delete r->in; r->in = 0; }
    r->in = ma::makeAdvanced(ma::configureIdentity(m, r->sf));
    r->a = new ma::Adapt(r->in);
    r->a->refinesLeft = refinesLeft;
    r->a->coarsensLeft = coarsensLeft;
  }
  apf::MeshIterator* it = m->begin(m->getDimension()); apf::MeshEntity* e; int64_t k = 0;
  while ((e = m->iterate(it))) {
    bool simplex = apf::isSimplex(m->getType(e));
    if (raw) raw[k] = simplex ? r->sf->getWeight(e) : 0;
    if (clamped) clamped[k] = simplex ? ma::getElementWeight(r->a, e) : 0;
    ++k;
  }
  m->end(it);
  return 0;
}

/* ---- the same for layer elements (maBalance.cc:21-81): a prism is weighed by its BASE TRIANGLE (raw = getWeight(first face),
   getSizeWeight :31-37), a pyramid by itself; clamped = ma::getElementWeight (clampForIterations, clampForLayerPermissions with
   the Input's defaults, accountForTets).  base_v [nelem][3] = the prism's first face in the FACE's own vertex order (-1 for
   other elements).  Simplex elements get what refo_weights gives them. */
int refo_layer_weights(void* hd, int refinesLeft, int coarsensLeft, double* raw, double* clamped, int32_t* base_v)
{
  Ref* r = (Ref*)hd; apf::Mesh2* m = r->m;
  if (!r->sf) return 1;
  if (r->a) { delete r->a; r->a = 0; }
  if (r->in) { delete r->in; r->in = 0; }
  r->in = ma::makeAdvanced(ma::configureIdentity(m, r->sf));
  r->a = new ma::Adapt(r->in);
  r->a->refinesLeft = refinesLeft;
  r->a->coarsensLeft = coarsensLeft;
  apf::MeshIterator* it = m->begin(m->getDimension()); apf::MeshEntity* e; int64_t k = 0;
  while ((e = m->iterate(it))) {
    const int t = m->getType(e);
    for (int i = 0; i < 3; ++i) base_v[3 * k + i] = -1;
    if (t == apf::Mesh::PRISM) {
      apf::Downward fs, fv;
      m->getDownward(e, 2, fs);
      m->getDownward(fs[0], 0, fv);
      for (int i = 0; i < 3; ++i) base_v[3 * k + i] = (int32_t)apf::getMdsIndex(m, fv[i]);
      raw[k] = r->sf->getWeight(fs[0]);
    } else raw[k] = r->sf->getWeight(e);
    clamped[k] = ma::getElementWeight(r->a, e);
    ++k;
  }
  m->end(it);
  return 0;
}

/* ---- ma::stats(m, sf, edgeLengths, linearQualities, true) (maStats.cc:115-134): the two vectors measureAnisoStats writes
   to its tables.  The caller sizes the outputs for every edge / element; the counts come back in n[2]. */
int refo_stats(void* hd, double* el, double* lq, int64_t* n)
{
  Ref* r = (Ref*)hd;
  if (!r->sf) return 1;
  std::vector<double> e, q;
  ma::stats(r->m, r->sf, e, q, true);
  n[0] = (int64_t)e.size(); n[1] = (int64_t)q.size();
  for (size_t i = 0; i < e.size(); ++i) el[i] = e[i];
  for (size_t i = 0; i < q.size(); ++i) lq[i] = q[i];
  return 0;
}

/* ---- ma::getSliverCode / ma::matchSliver (maShape.cc:35-120) of every tet, with goodQuality as given (< 0: default), and the
   vertices of each tet's first face (getDownward(tet, 2)[0]) in the FACE's own order, which is what measureTriQuality
   walks.  Non-tet elements get code 0 / match -1,-1 / face -1. */
int refo_sliver_codes(void* hd, double goodQuality, int32_t* codes, int32_t* match, int32_t* face0_v)
{
  Ref* r = (Ref*)hd; apf::Mesh2* m = r->m;
  if (!r->sf) return 1;
  if (r->a) { delete r->a; r->a = 0; }
  if (r->in) { delete r->in; r->in = 0; }
  r->in = ma::makeAdvanced(ma::configureIdentity(m, r->sf));
  if (goodQuality >= 0) r->in->goodQuality = goodQuality;
  r->a = new ma::Adapt(r->in);
  apf::MeshIterator* it = m->begin(m->getDimension()); apf::MeshEntity* e; int64_t k = 0;
  while ((e = m->iterate(it))) {
    if (m->getType(e) == apf::Mesh::TET) {
      codes[k] = ma::getSliverCode(r->a, e);
      ma::CodeMatch cm = ma::matchSliver(r->a, e);
      match[2 * k] = cm.rotation; match[2 * k + 1] = cm.code_index;
      apf::Downward fs, fv;
      m->getDownward(e, 2, fs);
      m->getDownward(fs[0], 0, fv);
      for (int i = 0; i < 3; ++i) face0_v[3 * k + i] = (int32_t)apf::getMdsIndex(m, fv[i]);
    } else {
      codes[k] = 0; match[2 * k] = match[2 * k + 1] = -1;
      for (int i = 0; i < 3; ++i) face0_v[3 * k + i] = -1;
    }
    ++k;
  }
  m->end(it);
  return 0;
}

/* ---- what ma::makeSplitVert (maRefine.cc:129-151) gives the vertex that splits each listed edge: its position
   (apf::mapLocalToGlobal at xi = 0) and the size-field values SizeField::interpolate writes on it.  A scratch vertex
   is created for each edge, read back and destroyed again.  Size-field kinds 2 (fields "refo_sizes"/"refo_frames":
   out_a = h[3], out_b = R[9]) and 3 / 5 (field "ma_logM": out_b = logM[9]). */
int refo_split_vertices(void* hd, int64_t n, const int64_t* edges, double* out_xyz, double* out_a, double* out_b)
{
  Ref* r = (Ref*)hd; apf::Mesh2* m = r->m;
  if (!r->sf) return 1;
  apf::Field* fh = m->findField("refo_sizes");
  apf::Field* fR = m->findField("refo_frames");
  apf::Field* fM = m->findField("ma_logM");
  if (!(fh && fR) && !fM) return 2;
  ma::Vector xi(0, 0, 0);
  for (int64_t i = 0; i < n; ++i) {
    apf::MeshEntity* e = apf::getMdsEntity(m, 1, (int)edges[i]);
    apf::MeshElement* me = apf::createMeshElement(m, e);
    ma::Vector point;
    apf::mapLocalToGlobal(me, xi, point);
    apf::MeshEntity* v = m->createVert(m->toModel(e));
    m->setPoint(v, 0, point);
    r->sf->interpolate(me, xi, v);
    for (int c = 0; c < 3; ++c) out_xyz[3*i+c] = point[c];
    if (fM) {
      apf::Matrix3x3 M; apf::getMatrix(fM, v, 0, M);
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) out_b[9*i+3*a+b] = M[a][b];
    } else {
      apf::Vector3 h; apf::getVector(fh, v, 0, h);
      apf::Matrix3x3 M; apf::getMatrix(fR, v, 0, M);
      for (int c = 0; c < 3; ++c) out_a[3*i+c] = h[c];
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) out_b[9*i+3*a+b] = M[a][b];
    }
    apf::destroyMeshElement(me);
    m->destroy(v);
  }
  return 0;
}

/* writes the mesh in the reference's native format (mds_write_smb, mds/mds_smb.c:640-690); with one part the file is
   <path minus ".smb">0.smb.  Fields present on the mesh travel as tags. */
void refo_write_smb(void* hd, const char* path)
{
  Ref* r = (Ref*)hd;
  r->m->writeNative(path);
}
/* vertex fields "sizes" (VECTOR) and "frames" (MATRIX) stored on the mesh without building a size field */
void refo_store_fields(void* hd, const double* h, const double* R)
{
  Ref* r = (Ref*)hd; apf::Mesh2* m = r->m;
  apf::Field* fh = apf::createFieldOn(m, "sizes", apf::VECTOR);
  apf::Field* fR = apf::createFieldOn(m, "frames", apf::MATRIX);
  apf::MeshIterator* it = m->begin(0); apf::MeshEntity* e;
  while ((e = m->iterate(it))) {
    int i = apf::getMdsIndex(m, e);
    apf::setVector(fh, e, 0, apf::Vector3(h + 3*i));
    apf::Matrix3x3 M;
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) M[a][b] = R[9*i+3*a+b];
    apf::setMatrix(fR, e, 0, M);
  }
  m->end(it);
}

/* apf::eigen on one 3x3 (mth::eigenQR), for the known-answer vectors of
   test/eigen_test.cc.  vecs: eigenvector j in row j. returns 3 */
int refo_eigen(const double* A, double* vals, double* vecs)
{
  apf::Matrix3x3 M;
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) M[a][b] = A[3*a+b];
  apf::Vector3 ev[3]; double l[3];
  int n = apf::eigen(M, ev, l);
  for (int j = 0; j < 3; ++j) { vals[j] = l[j]; for (int b = 0; b < 3; ++b) vecs[3*j+b] = ev[j][b]; }
  return n;
}

} // extern "C"
