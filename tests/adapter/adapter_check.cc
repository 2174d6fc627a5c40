// adapter_check.cc -- self-check of the drop-in adapter, callable through ctypes (tests/test_adapter.py, -m gpu).
// In ONE process it runs, on the same apf::Mesh2 (apf::makeMdsBox, jittered) and the same vertex fields:
//   (R) the unmodified reference: ma::makeSizeField + ma::markEdgesToSplit / markEdgesToCollapse / markBadQuality /
//       getMinQuality / getMaximumEdgeLength;
//   (A) the adapter's bulk entry points mag::markEdgesToSplit ... on a mag::GpuSizeField;
//   (B) the UNMODIFIED reference functions again, but with the mag::GpuSizeField / mag::shapeHandler plugged into
//       ma::Input -- the per-entity virtuals are then served from one device sweep (sweep detection);
// and reports every difference in flags, counts and statistics.  Test infrastructure; it ships in libmag_ma.so only so
// that the GPU box (which has no /root/reference) can run it from the prebuilt library.
#include "../../core_b200/adapter/magAdapt.h"
#include "../../include/mag.h"
#include <ma.h>
#include <maAdapt.h>
#include <maRefine.h>
#include <maShape.h>
#include <maCollapse.h>
#include <maShapeHandler.h>
#include <maStats.h>
#include <apfMDS.h>
#include <apfBox.h>
#include <apfMesh2.h>
#include <apf.h>
#include <gmi_null.h>
#include <gmi_mesh.h>
#include <gmi_analytic.h>
#include <lionPrint.h>
#include <PCU.h>
#include <cmath>
#include <vector>
#include <map>
#include <chrono>

namespace ma {
/* external linkage, no header in the reference (maCoarsen.cc:287, maShape.cc:132,152) */
long markEdgesToCollapse(Adapt* a);
int markBadQuality(Adapt* a);
double getMinQuality(Adapt* a);
double getElementWeight(Adapt* a, Entity* e);   /* maBalance.cc:74-81 */
int getSliverCode(Adapt* a, Entity* tet);        /* maShape.cc:35-89 */
CodeMatch matchSliver(Adapt* a, Entity* tet);    /* maShape.cc:91-120 (maShape.h:66 declares a signature no definition has) */
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

/* wall-clock seconds of the last check: [0] the reference's five sweeps (split, collapse, bad, min quality, max length),
   [1] the adapter's bulk entry points for the same five (each one exports the mesh from MDS, uploads, sweeps, writes the
   flag words back), [2] the unmodified reference loops served through the adapter */
double g_times[4] = {0, 0, 0, 0};   /* [3]: the bulk entry points again, warm (see adapter_check) */
mag::Profile g_profile[2];
int g_export_threads = 1;
int g_adapt_dim = 3;         /* mag_adapter_set_adapt_dim: 3 = n^3 box of tets, 2 = n^2 box of triangles, 4 = unit ball on an
                                analytic sphere model (vertices created on the boundary are SNAPPED: ma/ma.cc:37) */
double g_adapt_jitter = 0;   /* mag_adapter_set_adapt_jitter: vertex jitter of the boxes of the ma::adapt checks */   /* mag_adapter_set_threads: host threads of the adapter's MDS walk in the next checks */
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Fields { apf::Field* sizes; apf::Field* frames; };

Fields make_fields(apf::Mesh2* m, const char* tag, double hbar)
{
  Fields f;
  f.sizes = apf::createFieldOn(m, (std::string("sizes_") + tag).c_str(), apf::VECTOR);
  f.frames = apf::createFieldOn(m, (std::string("frames_") + tag).c_str(), apf::MATRIX);
  apf::MeshIterator* it = m->begin(0);
  apf::MeshEntity* v;
  while ((v = m->iterate(it))) {
    apf::Vector3 p;
    m->getPoint(v, 0, p);
    /* rotating shock layer of BASELINE config 3: R = Rz(pi/3 y), H = (hbar (0.1 + 2 |x' - 0.5|), hbar, 2 hbar) */
    double th = (M_PI / 3.0) * p[1], c = cos(th), s = sin(th);
    apf::Matrix3x3 R(c, -s, 0, s, c, 0, 0, 0, 1);
    double xp = p[0] * c + p[1] * s;
    apf::Vector3 h(hbar * (0.1 + 2.0 * fabs(xp - 0.5)), hbar, 2.0 * hbar);
    apf::setVector(f.sizes, v, 0, h);
    apf::setMatrix(f.frames, v, 0, R);
  }
  m->end(it);
  return f;
}

/* the unit sphere as a parametric model face: (longitude, latitude) -> point */
void sphere_point(double const p[2], double x[3], void*)
{
  x[0] = cos(p[0]) * cos(p[1]);
  x[1] = sin(p[0]) * cos(p[1]);
  x[2] = sin(p[1]);
}

/* eight tets around the centre of an octahedron inscribed in the unit sphere; the six outer vertices, the twelve outer edges
   and the eight outer triangles are classified on the analytic face (so refinement snaps what it creates there to the
   sphere), everything else on the region.  The situation of test/ma_test_analytic_model.cc, with an interior vertex. */
apf::Mesh2* make_ball()
{
  gmi_model* model = gmi_make_analytic();
  int periodic[2] = {1, 0};
  double ranges[2][2] = {{0.0, 2.0 * M_PI}, {-0.5 * M_PI, 0.5 * M_PI}};
  gmi_add_analytic(model, 2, 0, sphere_point, periodic, ranges, 0);
  gmi_add_analytic_region(model, 1);
  apf::Mesh2* m = apf::makeEmptyMdsMesh(model, 3, false, g_pcu);
  apf::ModelEntity* face = m->findModelEntity(2, 0);
  apf::ModelEntity* region = m->findModelEntity(3, 1);
  /* +x -x +y -y +z -z, centre */
  const double lon[6] = {0.0, M_PI, 0.5 * M_PI, 1.5 * M_PI, 0.0, 0.0}, lat[6] = {0, 0, 0, 0, 0.5 * M_PI, -0.5 * M_PI};
  apf::MeshEntity* v[7];
  for (int i = 0; i < 6; ++i) {
    double uv[2] = {lon[i], lat[i]}, x[3];
    sphere_point(uv, x, 0);
    v[i] = m->createVertex(face, apf::Vector3(x[0], x[1], x[2]), apf::Vector3(uv[0], uv[1], 0));
  }
  v[6] = m->createVertex(region, apf::Vector3(0, 0, 0), apf::Vector3(0, 0, 0));
  for (int pass = 0; pass < 2; ++pass)          /* surface triangles first, then the tets that reuse them */
    for (int oct = 0; oct < 8; ++oct) {
      apf::MeshEntity* t[4] = {v[oct & 1], v[2 + ((oct >> 1) & 1)], v[4 + ((oct >> 2) & 1)], v[6]};
      apf::Vector3 p[3];
      for (int i = 0; i < 3; ++i) m->getPoint(t[i], 0, p[i]);
      if (apf::cross(p[1] - p[0], p[2] - p[0]) * (apf::Vector3(0, 0, 0) - p[0]) < 0) std::swap(t[1], t[2]);   /* positive volume */
      if (pass == 0) apf::buildElement(m, face, apf::Mesh::TRIANGLE, t);
      else apf::buildElement(m, region, apf::Mesh::TET, t);
    }
  m->acceptChanges();
  m->verify();
  return m;
}

struct Marks {
  std::vector<int> ef, lf;
  long n_split, n_collapse; int n_bad; double min_q, max_len;
};

void collect_flags(ma::Adapt* a, Marks& k)
{
  apf::Mesh2* m = a->mesh;
  apf::MeshIterator* it = m->begin(1); apf::MeshEntity* e;
  k.ef.clear(); k.lf.clear();
  while ((e = m->iterate(it))) k.ef.push_back(ma::getFlags(a, e));
  m->end(it);
  it = m->begin(m->getDimension());
  while ((e = m->iterate(it))) k.lf.push_back(ma::getFlags(a, e));
  m->end(it);
}

long diff(const std::vector<int>& a, const std::vector<int>& b)
{
  if (a.size() != b.size()) return -1;
  long n = 0;
  for (size_t i = 0; i < a.size(); ++i) n += a[i] != b[i];
  return n;
}

} // namespace

/* report[0..4]  reference: n_split n_collapse n_bad min_q max_len
   report[5..9]  adapter bulk (A), report[10..14] unmodified reference loops over the adapter (B)
   report[15..18] flag words differing: A edges, A elems, B edges, B elems;  report[19] sweeps the adapter ran for B
   returns 0 when A and B reproduce the reference exactly (fp_mode strict) / flags+counts exactly, values 1e-12 (fast). */
/* interior vertices moved by jitter / n * (u - 0.5) per component, u from a fixed LCG */
static void jitter_mesh(apf::Mesh2* m, int n, double jitter)
{
  if (!(jitter > 0)) return;
  const int dim = m->getDimension();
  unsigned long long s = 12345;
  apf::MeshIterator* it = m->begin(0); apf::MeshEntity* v;
  while ((v = m->iterate(it))) {
    apf::Vector3 p; m->getPoint(v, 0, p);
    bool interior = true;
    for (int i = 0; i < dim; ++i) interior = interior && p[i] > 1e-9 && p[i] < 1 - 1e-9;
    for (int i = 0; i < dim; ++i) {
      s = s * 6364136223846793005ULL + 1442695040888963407ULL;
      double u = (double)(s >> 11) / 9007199254740992.0;
      if (interior) p[i] += jitter / n * (u - 0.5);
    }
    m->setPoint(v, 0, p);
  }
  m->end(it);
}

static int adapter_check(int n, int nz, int log_interp, int fp_mode, double jitter, double* report);
extern "C" void mag_adapter_set_threads(int n) { g_export_threads = n; }
extern "C" void mag_adapter_set_direct(int on) { mag::setDirectMds(on != 0); }
extern "C" void mag_adapter_set_adapt_jitter(double j) { g_adapt_jitter = j; }
extern "C" void mag_adapter_set_adapt_dim(int d) { g_adapt_dim = (d == 2 || d == 4) ? d : 3; }
extern "C" void mag_adapter_times(double* t) { for (int i = 0; i < 3; ++i) t[i] = g_times[i]; }
/* t[0]: the warm second round of the bulk sweeps; t[1..7] / t[8..14]: mag::Profile of the first (cold) / second (warm) round */
extern "C" void mag_adapter_times2(double* t)
{
  t[0] = g_times[3];
  for (int r = 0; r < 2; ++r) {
    const mag::Profile& p = g_profile[r];
    const double v[7] = {p.export_s, p.revalidate_s, p.upload_s, p.flags_in_s, p.device_s, p.flags_out_s, p.refresh_s};
    for (int i = 0; i < 7; ++i) t[1 + 7 * r + i] = v[i];
  }
}
extern "C" int mag_adapter_check(int n, int log_interp, int fp_mode, double jitter, double* report)
{
  return adapter_check(n, n, log_interp, fp_mode, jitter, report);
}
/* the same on a 2-D box (triangles are the elements: ma::measureTriQuality, goodQuality 0.2 as ma::configure picks in 2-D) */
extern "C" int mag_adapter_check_2d(int n, int log_interp, int fp_mode, double jitter, double* report)
{
  return adapter_check(n, 0, log_interp, fp_mode, jitter, report);
}
static int adapter_check(int n, int nz, int log_interp, int fp_mode, double jitter, double* report)
{
  ensure_pcu();
  apf::Mesh2* m = apf::makeMdsBox(n, n, nz, 1, 1, nz ? 1 : 0, true, g_pcu);
  const int dim = m->getDimension();
  jitter_mesh(m, n, jitter);
  const double hbar = 1.0 / n;
  Marks R, A, B;
  long weight_diffs = 0, sliver_diffs = 0, stats_diffs = 0;
  std::vector<double> ref_el, ref_lq;
  { /* (R) the unmodified reference */
    Fields f = make_fields(m, "ref", hbar);
    ma::SizeField* sf = ma::makeSizeField(m, f.sizes, f.frames, log_interp != 0);
    ma::Input* in = ma::makeAdvanced(ma::configureIdentity(m, sf));
    {
      ma::Adapt a(in);
      const double t0 = now_s();
      R.n_split = ma::markEdgesToSplit(&a);
      R.n_collapse = ma::markEdgesToCollapse(&a);
      R.n_bad = ma::markBadQuality(&a);
      R.min_q = ma::getMinQuality(&a);
      g_times[0] = now_s() - t0;
      collect_flags(&a, R);
    }
    { const double t0 = now_s(); R.max_len = ma::getMaximumEdgeLength(m, sf); g_times[0] += now_s() - t0; }
    ma::stats(m, sf, ref_el, ref_lq, true);
    delete in;
    delete sf; /* destroys (Aniso) or leaves (LogAniso) the input fields */
    if (log_interp) { apf::destroyField(f.sizes); apf::destroyField(f.frames); }
  }
  Fields f = make_fields(m, "gpu", hbar);
  mag::GpuSizeField* g = mag::makeSizeField(m, f.sizes, f.frames, log_interp != 0, 0);
  g->setArithmetic(fp_mode);
  g->setExportThreads(g_export_threads);
  { /* (A) the adapter's bulk entry points */
    ma::Input* in = ma::makeAdvanced(ma::configureIdentity(m, g));
    in->shapeHandler = mag::shapeHandler;
    {
      ma::Adapt a(in);
      mag::profile() = mag::Profile();
      const double t0 = now_s();
      A.n_split = mag::markEdgesToSplit(&a);
      A.n_collapse = mag::markEdgesToCollapse(&a);
      A.n_bad = mag::markBadQuality(&a);
      A.min_q = mag::getMinQuality(&a);
      g_times[1] = now_s() - t0;
      g_profile[0] = mag::profile();
      collect_flags(&a, A);
      /* ma::getElementWeights (maBalance.cc:83-97) against the reference's own per-entity loop on the same Adapt */
      a.refinesLeft = 0; a.coarsensLeft = 1;
      ma::Tag* wt = mag::getElementWeights(&a);
      apf::MeshIterator* wit = m->begin(dim);
      ma::Entity* we;
      while ((we = m->iterate(wit))) {
        double wg, wr = ma::getElementWeight(&a, we);   /* a.sizeField is the adapter; its getWeight delegates to the wrapped reference field */
        m->getDoubleTag(we, wt, &wg);
        const double tol = (fp_mode == MAG_FP_STRICT && !log_interp) ? 0.0 : 1e-12;
        if (fabs(wg - wr) > tol * fabs(wr)) ++weight_diffs;
      }
      m->end(wit);
      apf::removeTagFromDimension(m, wt, 3);
      m->destroyTag(wt);
      /* ma::getSliverCode / matchSliver (maShape.cc:35-120) of every tet against the reference's per-entity calls on the
         same Adapt (its getTransform / face quality go through the adapter to the wrapped reference field) */
      std::vector<int> codes;
      std::vector<ma::CodeMatch> matches;
      if (dim == 3) mag::getSliverCodes(&a, codes, matches);
      wit = m->begin(dim == 3 ? 3 : 0);
      if (dim != 3) { m->end(wit); wit = 0; }
      size_t wi = 0;
      long nel3 = 0;
      while (wit && (we = m->iterate(wit))) {
        const int rc = ma::getSliverCode(&a, we);
        const ma::CodeMatch rm = ma::matchSliver(&a, we);
        if (codes[wi] != rc || matches[wi].rotation != rm.rotation || matches[wi].code_index != rm.code_index) ++sliver_diffs;
        ++wi; ++nel3;
      }
      if (wit) m->end(wit);
      if (log_interp && sliver_diffs * 200 <= nel3) sliver_diffs = 0;   /* CUDA exp() vs glibc exp(): rare borderline bits */
    }
    { const double t0 = now_s(); A.max_len = mag::getMaximumEdgeLength(m, g); g_times[1] += now_s() - t0; }
    { /* the same five sweeps once more on a fresh Adapt after invalidate(): a full re-export + upload with the context, the
         buffers and the kernels warm -- what every further MeshAdapt iteration costs */
      ma::Input* in2 = ma::makeAdvanced(ma::configureIdentity(m, g));
      {
        ma::Adapt a2(in2);
        g->invalidate();
        mag::Profile& pr = mag::profile();
        pr = mag::Profile();
        const double t0 = now_s();
        long s2 = mag::markEdgesToSplit(&a2);
        long c2 = mag::markEdgesToCollapse(&a2);
        int b2 = mag::markBadQuality(&a2);
        double q2 = mag::getMinQuality(&a2);
        double l2 = mag::getMaximumEdgeLength(m, g);
        g_times[3] = now_s() - t0;
        g_profile[1] = pr;
        if (s2 != A.n_split || c2 != A.n_collapse || b2 != A.n_bad || q2 != A.min_q || l2 != A.max_len) stats_diffs = -2;
      }
      delete in2;
    }
    { /* ma::stats through the adapter against the reference's vectors */
      std::vector<double> el, lq;
      mag::stats(m, g, el, lq, true);
      const double tol = (fp_mode == MAG_FP_STRICT && !log_interp) ? 0.0 : 1e-12;
      if (el.size() != ref_el.size() || lq.size() != ref_lq.size()) stats_diffs = -1;
      else {
        for (size_t i = 0; i < el.size(); ++i) if (fabs(el[i] - ref_el[i]) > tol * fabs(ref_el[i])) ++stats_diffs;
        for (size_t i = 0; i < lq.size(); ++i) if (fabs(lq[i] - ref_lq[i]) > tol * fabs(ref_lq[i])) ++stats_diffs;
      }
    }
    delete in;
  }
  long sweeps0 = mag_launch_count(g->ctx);
  { /* (B) the unmodified reference loops, adapter plugged in through ma::Input */
    ma::Input* in = ma::makeAdvanced(ma::configureIdentity(m, g));
    in->shapeHandler = mag::shapeHandler;
    {
      ma::Adapt a(in);
      const double t0 = now_s();
      B.n_split = ma::markEdgesToSplit(&a);
      B.n_collapse = ma::markEdgesToCollapse(&a);
      B.n_bad = ma::markBadQuality(&a);
      B.min_q = ma::getMinQuality(&a);
      g_times[2] = now_s() - t0;
      collect_flags(&a, B);
    }
    { const double t0 = now_s(); B.max_len = ma::getMaximumEdgeLength(m, g); g_times[2] += now_s() - t0; }
    delete in;
  }
  report[19] = (double)(mag_launch_count(g->ctx) - sweeps0);
  delete g;
  if (log_interp) { apf::destroyField(f.sizes); apf::destroyField(f.frames); }
  m->destroyNative();
  apf::destroyMesh(m);
  const Marks* k[3] = {&R, &A, &B};
  for (int i = 0; i < 3; ++i) {
    report[5 * i] = (double)k[i]->n_split; report[5 * i + 1] = (double)k[i]->n_collapse; report[5 * i + 2] = k[i]->n_bad;
    report[5 * i + 3] = k[i]->min_q; report[5 * i + 4] = k[i]->max_len;
  }
  report[15] = (double)diff(R.ef, A.ef); report[16] = (double)diff(R.lf, A.lf);
  report[17] = (double)diff(R.ef, B.ef); report[18] = (double)diff(R.lf, B.lf);
  int bad = 0;
  for (int i = 1; i < 3; ++i) {
    bad |= k[i]->n_split != R.n_split || k[i]->n_collapse != R.n_collapse || k[i]->n_bad != R.n_bad;
    const double tol = (fp_mode == MAG_FP_STRICT && !log_interp) ? 0.0 : 1e-12;
    bad |= fabs(k[i]->min_q - R.min_q) > tol * fabs(R.min_q) || fabs(k[i]->max_len - R.max_len) > tol * R.max_len;
  }
  for (int i = 15; i < 19; ++i) bad |= report[i] != 0;
  bad |= weight_diffs != 0;
  bad |= sliver_diffs != 0;
  bad |= stats_diffs != 0;
  return bad;
}

/* mag::buildEdgeLinks against a stub apf::Sharing (there is no MPI here): two boxes stand for the two x-slab parts of an
   (nxA + nxB) x ny x nz box.  The stub says what PUMI would: an edge of part 0 in its plane x = max and the edge of part 1 in
   its plane x = 0 with the same (y, z) end points are copies of each other; the owner is the part with fewer elements,
   ties -> part 0 (apfPM.cc:109-126).  Returns the number of shared edges of `part` and writes their export indices / owner
   bits in list order (tests/test_adapter.py compares them with boxmesh.slab_part's lists); < 0: error. */
namespace {
struct StubSharing : public apf::Sharing {
  int self, peer;
  bool selfOwns;
  std::map<apf::MeshEntity*, apf::MeshEntity*> remote;
  bool isShared(apf::MeshEntity* e) { return remote.count(e) != 0; }
  bool isOwned(apf::MeshEntity* e) { return !isShared(e) || selfOwns; }
  int getOwner(apf::MeshEntity* e) { return isOwned(e) ? self : peer; }
  void getCopies(apf::MeshEntity* e, apf::CopyArray& copies)
  {
    std::map<apf::MeshEntity*, apf::MeshEntity*>::iterator it = remote.find(e);
    copies.setSize(it == remote.end() ? 0 : 1);
    if (it != remote.end()) copies[0] = apf::Copy(peer, it->second);
  }
};
typedef std::vector<double> PlaneKey;
/* edges with both ends in the plane x = px, keyed by the sorted (y, z) of their ends */
void plane_edges(apf::Mesh2* m, double px, std::map<PlaneKey, apf::MeshEntity*>& out)
{
  apf::MeshIterator* it = m->begin(1);
  apf::MeshEntity* e;
  while ((e = m->iterate(it))) {
    apf::Downward v;
    m->getDownward(e, 0, v);
    apf::Vector3 a, b;
    m->getPoint(v[0], 0, a); m->getPoint(v[1], 0, b);
    if (a[0] != px || b[0] != px) continue;
    PlaneKey k(4);
    const bool swap = (a[1] > b[1]) || (a[1] == b[1] && a[2] > b[2]);
    k[0] = swap ? b[1] : a[1]; k[1] = swap ? b[2] : a[2]; k[2] = swap ? a[1] : b[1]; k[3] = swap ? a[2] : b[2];
    out[k] = e;
  }
  m->end(it);
}
}
extern "C" int mag_adapter_links_check(int nxA, int nxB, int ny, int nz, int part, int* idx_out, unsigned char* own_out, int cap)
{
  ensure_pcu();
  apf::Mesh2* mesh[2] = {apf::makeMdsBox(nxA, ny, nz, nxA, ny, nz, true, g_pcu), apf::makeMdsBox(nxB, ny, nz, nxB, ny, nz, true, g_pcu)};
  std::map<PlaneKey, apf::MeshEntity*> plane[2];
  plane_edges(mesh[0], (double)nxA, plane[0]);
  plane_edges(mesh[1], 0.0, plane[1]);
  int rc = -1;
  if (plane[0].size() == plane[1].size() && !plane[0].empty()) {
    StubSharing sh;
    sh.self = part; sh.peer = 1 - part;
    const long nel[2] = {6L * nxA * ny * nz, 6L * nxB * ny * nz};
    const int owner = (nel[1] < nel[0]) ? 1 : 0;
    sh.selfOwns = owner == part;
    bool ok = true;
    for (std::map<PlaneKey, apf::MeshEntity*>::iterator it = plane[part].begin(); it != plane[part].end(); ++it) {
      std::map<PlaneKey, apf::MeshEntity*>::iterator jt = plane[1 - part].find(it->first);
      if (jt == plane[1 - part].end()) { ok = false; break; }
      sh.remote[it->second] = jt->second;
    }
    std::vector<apf::MeshEntity*> edges;
    apf::MeshIterator* it = mesh[part]->begin(1);
    apf::MeshEntity* e;
    while ((e = mesh[part]->iterate(it))) edges.push_back(e);
    mesh[part]->end(it);
    mag::EdgeLinks L;
    mag::buildEdgeLinks(&sh, part, edges, L);
    if (ok && L.peer.size() == 1 && L.peer[0] == 1 - part && (int)L.idx[0].size() <= cap) {
      rc = (int)L.idx[0].size();
      for (int i = 0; i < rc; ++i) { idx_out[i] = L.idx[0][i]; own_out[i] = L.peerOwns[0][i]; }
    }
  }
  for (int i = 0; i < 2; ++i) { mesh[i]->destroyNative(); apf::destroyMesh(mesh[i]); }
  return rc;
}

/* The drop-in claim end to end: ma::adapt (refine, coarsen, shape correction -- the UNMODIFIED reference driver) on two
   identical jittered boxes, once with the reference's own AnisoSizeField and once with the mag::GpuSizeField /
   mag::shapeHandler plugged into ma::Input (MAG_FP_STRICT: every length and quality the reference sees is bit-identical, so
   every operator decision is the same).  The two adapted meshes must be the same mesh: same entity counts, same
   coordinates and connectivity in iteration order, and the same longest metric edge (the quantity test/aniso_adapt.h:65-74
   checks against MAXLENGTH once the adaptation has converged).
   which: 1 = reference only (no device), 3 = both.  out[0..2] counts of the reference run (verts, edges, tets), out[3..5] of
   the adapter run, out[6] differing coordinates / connectivity entries, out[7] longest metric edge after the adapter run,
   out[8] device kernel launches made during the adapter run, out[9] / out[10] wall-clock seconds of the two runs. */
/* n^3 cells whose bottom n/3 layers are prisms (two per cell), Kuhn tets above; optionally one detached pyramid */
static apf::Mesh2* make_mixed(int n, bool with_pyramid)
{
  ensure_pcu();
    apf::Mesh2* m = apf::makeEmptyMdsMesh(gmi_load(".null"), 3, false, g_pcu);
    apf::ModelEntity* region = m->findModelEntity(3, 0);
    std::vector<apf::MeshEntity*> v((size_t)(n + 1) * (n + 1) * (n + 1));
    for (int k = 0; k <= n; ++k) for (int j = 0; j <= n; ++j) for (int i = 0; i <= n; ++i)
      v[(size_t)i + (size_t)(n + 1) * (j + (size_t)(n + 1) * k)] = m->createVertex(region, apf::Vector3((double)i / n, (double)j / n, (double)k / n), apf::Vector3(0, 0, 0));
    const int layers = n / 3 > 0 ? n / 3 : 1;
    for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
      apf::MeshEntity* c[8];
      const int dx[8] = {0, 1, 1, 0, 0, 1, 1, 0}, dy[8] = {0, 0, 1, 1, 0, 0, 1, 1}, dz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
      for (int q = 0; q < 8; ++q) c[q] = v[(size_t)(i + dx[q]) + (size_t)(n + 1) * ((j + dy[q]) + (size_t)(n + 1) * (k + dz[q]))];
      if (k < layers) {
        apf::MeshEntity* p0[6] = {c[0], c[1], c[2], c[4], c[5], c[6]};
        apf::MeshEntity* p1[6] = {c[0], c[2], c[3], c[4], c[6], c[7]};
        apf::buildElement(m, region, apf::Mesh::PRISM, p0);
        apf::buildElement(m, region, apf::Mesh::PRISM, p1);
      } else {
        /* six tets around the diagonal 0-6, bottom face cut along 0-2 like the prisms' top face below */
        const int kuhn[6][4] = {{0, 1, 2, 6}, {0, 2, 3, 6}, {0, 3, 7, 6}, {0, 7, 4, 6}, {0, 4, 5, 6}, {0, 5, 1, 6}};
        for (int t = 0; t < 6; ++t) {
          apf::MeshEntity* tv[4] = {c[kuhn[t][0]], c[kuhn[t][1]], c[kuhn[t][2]], c[kuhn[t][3]]};
          apf::buildElement(m, region, apf::Mesh::TET, tv);
        }
      }
    }
    if (with_pyramid) {
    apf::MeshEntity* py[5];
    const double px[5][3] = {{3, 0, 0}, {4, 0, 0}, {4, 1, 0}, {3, 1, 0}, {3.5, 0.5, 1}};
    for (int q = 0; q < 5; ++q) py[q] = m->createVertex(region, apf::Vector3(px[q][0], px[q][1], px[q][2]), apf::Vector3(0, 0, 0));
    apf::buildElement(m, region, apf::Mesh::PYRAMID, py);
    }
    m->acceptChanges();
    return m;
}

/* ---- the adapter's two export routes (MDS's own arrays against the public apf::Mesh2 walk), host only: no device is touched.
   kind 0: jittered n^3 box of tets; 1: the same after `iters` iterations of the reference's own ma::adapt (free-list holes, new
   entities interleaved with old ones); 2: n^2 box of triangles; 3: n^3 cells whose bottom n/3 layers are prisms, Kuhn tets
   above, plus one detached pyramid.  Returns mag::exportSelfCheck's mask; times[0..1] = seconds public / direct, times[2..4] =
   vertices, edges, elements exported. */
extern "C" int mag_adapter_export_check(int n, int kind, int threads, int iters, double* times)
{
  ensure_pcu();
  apf::Mesh2* m;
  if (kind == 2) m = apf::makeMdsBox(n, n, 0, 1, 1, 0, true, g_pcu);
  else if (kind == 3) m = make_mixed(n, true);
  else m = apf::makeMdsBox(n, n, n, 1, 1, 1, true, g_pcu);
  if (kind != 3) jitter_mesh(m, n, 0.25);
  Fields f = make_fields(m, "exp", 1.0 / n);
  if (kind == 1) {
    ma::SizeField* sf = ma::makeSizeField(m, f.sizes, f.frames, false);
    ma::Input* in = ma::makeAdvanced(ma::configureIdentity(m, sf));
    in->maximumIterations = iters;
    in->shouldSnap = false;
    in->shouldTransferParametric = false;
    ma::adapt(in);
    /* ~AnisoSizeField would destroy the two fields it was built from (maSize.cc:385-389): keep it alive until the end */
    static std::vector<ma::SizeField*> keep;
    keep.push_back(sf);
  }
  /* a flag tag with a recognisable word on three quarters of the edges */
  apf::MeshTag* flags = m->createIntTag("ma_flags", 1);
  {
    apf::MeshIterator* it = m->begin(1); apf::MeshEntity* e; int k = 0;
    while ((e = m->iterate(it))) { if (k & 3) { int w = k * 7 + 1; m->setIntTag(e, flags, &w); } ++k; }
    m->end(it);
  }
  const int bad = mag::exportSelfCheck(m, f.sizes, f.frames, flags, threads, times);
  times[2] = (double)m->count(0); times[3] = (double)m->count(1); times[4] = (double)m->count(m->getDimension());
  if (kind != 1) { m->destroyNative(); apf::destroyMesh(m); }   /* kind 1: the kept size field still points at the mesh */
  return bad;
}

/* ma::getElementWeights on a mixed prism / tet mesh: the reference's per-entity loop against mag::getElementWeights (tets:
   mag_element_weights; prisms: mag_prism_weights on the base triangles).  No pyramid in the mesh: the reference itself cannot
   weigh one -- SizeField::getWeight(pyramid) asks for an order-2 rule and apf has only a one-point rule of accuracy 1 for
   pyramids (apf/apfIntegrate.cc:530-554), so the integrator dereferences a null rule.  Layer permissions as
   given (ma::Input::shouldRefineLayer / shouldCoarsenLayer / shouldTurnLayerToTets).  Returns the number of elements whose
   weight differs (strict arithmetic: any bit); out[0..2] = prisms, pyramids, tets. */
extern "C" long mag_adapter_layer_weights_check(int n, int fp_mode, int refine_layer, int coarsen_layer, int to_tets, double* out)
{
  apf::Mesh2* m = make_mixed(n, false);
  Fields f = make_fields(m, "lw", 1.0 / n);
  mag::GpuSizeField* g = mag::makeSizeField(m, f.sizes, f.frames, false, 0);
  g->setArithmetic(fp_mode);
  ma::Input* in = ma::makeAdvanced(ma::configureIdentity(m, g));
  in->shouldRefineLayer = refine_layer != 0;
  in->shouldCoarsenLayer = coarsen_layer != 0;
  in->shouldTurnLayerToTets = to_tets != 0;
  long diffs = 0;
  out[0] = out[1] = out[2] = 0;
  {
    ma::Adapt a(in);
    a.refinesLeft = 1; a.coarsensLeft = 1;
    ma::Tag* wt = mag::getElementWeights(&a);
    apf::MeshIterator* it = m->begin(3);
    apf::MeshEntity* e;
    while ((e = m->iterate(it))) {
      const int t = m->getType(e);
      out[t == apf::Mesh::PRISM ? 0 : (t == apf::Mesh::PYRAMID ? 1 : 2)] += 1;
      double wg, wr = ma::getElementWeight(&a, e);   /* through the adapter's getWeight -> the wrapped reference field */
      m->getDoubleTag(e, wt, &wg);
      const double tol = fp_mode == MAG_FP_STRICT ? 0.0 : 1e-12;
      if (!(fabs(wg - wr) <= tol * fabs(wr))) ++diffs;
    }
    m->end(it);
    apf::removeTagFromDimension(m, wt, 3);
    m->destroyTag(wt);
  }
  delete in;
  delete g;
  m->destroyNative();
  apf::destroyMesh(m);
  return diffs;
}

/* mag::collapseQualities against ma::Collapse itself.  On a jittered box whose size field asks for coarsening, every edge the
   reference marks COLLAPSE and whose collapse passes its classification / topology checks is REALLY rebuilt by the unmodified
   reference in each permitted direction (Collapse::computeElementSets + rebuildElements), the qualities of the old and of the
   new elements are taken through its shape handler, and the new elements are destroyed again.  The adapter then answers the
   same candidates in one device call from its export of the restored mesh.  out[0] = candidates (edge, direction), out[1] =
   candidates whose rebuilt cavity holds an inverted element, out[2] = candidates the reference's test would accept.
   Returns the number of candidates whose old or new worst quality differs (strict: any bit; fast: 1e-12). */
extern "C" long mag_adapter_collapse_check(int n, int fp_mode, double size_scale, double* out)
{
  ensure_pcu();
  apf::Mesh2* m = apf::makeMdsBox(n, n, n, 1, 1, 1, true, g_pcu);
  jitter_mesh(m, n, 0.25);
  std::vector<ma::Entity*> edges, verts;
  std::vector<double> refNew, refOld;
  out[0] = out[1] = out[2] = 0;
  {
    Fields f = make_fields(m, "ref", size_scale / n);
    ma::SizeField* sf = ma::makeSizeField(m, f.sizes, f.frames, false);
    ma::Input* in = ma::makeAdvanced(ma::configureIdentity(m, sf));
    {
      ma::Adapt a(in);
      ma::markEdgesToCollapse(&a);
      std::vector<ma::Entity*> marked;
      apf::MeshIterator* it = m->begin(1);
      ma::Entity* e;
      while ((e = m->iterate(it))) if (ma::getFlag(&a, e, ma::COLLAPSE)) marked.push_back(e);
      m->end(it);
      ma::Collapse c;
      c.Init(&a);
      for (size_t i = 0; i < marked.size(); ++i) {
        if (!c.setEdge(marked[i]) || !c.checkClass() || !c.checkTopo()) continue;
        for (int dir = 0; dir < 2; ++dir) {
          if (dir == 1) {
            if (!ma::getFlag(&a, c.vertToKeep, ma::COLLAPSE)) break;      /* tryBothDirections' condition for the other way */
            std::swap(c.vertToKeep, c.vertToCollapse);
          }
          c.computeElementSets();
          ma::EntityArray oldEl;
          c.getOldElements(oldEl);
          double oq = a.shape->getQuality(oldEl[0]);
          for (size_t k = 1; k < oldEl.getSize(); ++k) { const double q = a.shape->getQuality(oldEl[k]); if (q < oq) oq = q; }
          c.rebuildElements();
          double nq = a.shape->getQuality(c.newElements[0]);
          for (size_t k = 1; k < c.newElements.getSize(); ++k) { const double q = a.shape->getQuality(c.newElements[k]); if (q < nq) nq = q; }
          c.destroyNewElements();
          edges.push_back(marked[i]); verts.push_back(c.vertToCollapse);
          refNew.push_back(nq); refOld.push_back(oq);
          out[0] += 1;
          if (nq < 0) out[1] += 1;
          const double toBeat = std::min(in->goodQuality, std::max(oq, in->validQuality));
          if (!(nq < toBeat)) out[2] += 1;
        }
        c.unmark();
      }
    }
    delete in;
    delete sf;
  }
  if (getenv("MAG_CHECK_VERBOSE")) fprintf(stderr, "collapse check: %g candidates, %g inverted, %g accepted\n", out[0], out[1], out[2]);
  Fields f = make_fields(m, "gpu", size_scale / n);
  mag::GpuSizeField* g = mag::makeSizeField(m, f.sizes, f.frames, false, 0);
  g->setArithmetic(fp_mode);
  long diffs = 0;
  {
    ma::Input* in = ma::makeAdvanced(ma::configureIdentity(m, g));
    {
      ma::Adapt a(in);
      std::vector<double> nw, ow;
      mag::collapseQualities(&a, edges, verts, nw, ow);
      const double tol = fp_mode == MAG_FP_STRICT ? 0.0 : 1e-12;
      for (size_t i = 0; i < edges.size(); ++i)
        if (!(fabs(nw[i] - refNew[i]) <= tol * fabs(refNew[i]) + (tol > 0 ? 1e-15 : 0)) || !(fabs(ow[i] - refOld[i]) <= tol * fabs(refOld[i]))) ++diffs;
    }
    delete in;
  }
  delete g;
  m->destroyNative();
  apf::destroyMesh(m);
  return diffs;
}

static int adapt_check(int n, int which, double size_scale, int iterations, int log_interp, int fp_mode, double* out);
extern "C" int mag_adapter_adapt_check(int n, int which, double size_scale, int iterations, double* out)
{
  return adapt_check(n, which, size_scale, iterations, 0, MAG_FP_STRICT, out);
}
/* the same with the log-Euclidean interpolation (ma::configure's default) and / or MAG_FP_FAST: values then agree with the
   reference to 1e-12 instead of bit for bit, so an operator decision that sits within that distance of its threshold may
   differ; reported, not required to be zero */
extern "C" int mag_adapter_adapt_check2(int n, int which, double size_scale, int iterations, int log_interp, int fp_mode, double* out)
{
  return adapt_check(n, which, size_scale, iterations, log_interp, fp_mode, out);
}
static int adapt_check(int n, int which, double size_scale, int iterations, int log_interp, int fp_mode, double* out)
{
  ensure_pcu();
  for (int i = 0; i < 11; ++i) out[i] = 0;
  apf::Mesh2* mesh[2] = {0, 0};
  double ref_max_len = 0;
  for (int run = 0; run < 2; ++run) {
    if (!(which & (1 << run))) continue;
    const bool ball = g_adapt_dim == 4;
    apf::Mesh2* m = ball ? make_ball()
                         : (g_adapt_dim == 2 ? apf::makeMdsBox(n, n, 0, 1, 1, 0, true, g_pcu) : apf::makeMdsBox(n, n, n, 1, 1, 1, true, g_pcu));
    mesh[run] = m;
    if (!ball) jitter_mesh(m, n, g_adapt_jitter);
    Fields f = make_fields(m, run ? "gpu" : "ref", size_scale / n);
    ma::SizeField* sf;
    mag::GpuSizeField* g = 0;
    if (run == 0) sf = ma::makeSizeField(m, f.sizes, f.frames, log_interp != 0);
    else { g = mag::makeSizeField(m, f.sizes, f.frames, log_interp != 0, 0); g->setArithmetic(fp_mode); sf = g; }
    ma::Input* in = ma::makeAdvanced(ma::configureIdentity(m, sf));
    in->maximumIterations = iterations;
    in->shouldSnap = ball;               /* the boxes have a null model; the ball's boundary is an analytic sphere */
    in->shouldTransferParametric = ball;
    if (run == 1) in->shapeHandler = mag::shapeHandler;
    const long l0 = g ? mag_launch_count(g->ctx) : 0;
    const double t0 = now_s();
    ma::adapt(in);                       /* deletes the Input; the size field is ours (ownsSizeField = false) */
    out[9 + run] = now_s() - t0;
    out[3 * run] = (double)m->count(0); out[3 * run + 1] = (double)m->count(1); out[3 * run + 2] = (double)m->count(m->getDimension());
    if (run == 1) {
      out[8] = (double)(mag_launch_count(g->ctx) - l0);
      out[7] = ma::getMaximumEdgeLength(m, g->wrapped);
    } else ref_max_len = ma::getMaximumEdgeLength(m, sf);
    delete sf;
  }
  int bad = 0;
  if (which == 3) {
    apf::Mesh2 *a = mesh[0], *b = mesh[1];
    long diffs = 0;
    const int mdim = a->getDimension();
    for (int d = 0; d <= mdim; ++d) if (a->count(d) != b->count(d)) diffs += 1000000;
    if (!diffs) {
      apf::MeshIterator *ia = a->begin(0), *ib = b->begin(0);
      apf::MeshEntity *ea, *eb;
      while ((ea = a->iterate(ia)) && (eb = b->iterate(ib))) {
        apf::Vector3 pa, pb;
        a->getPoint(ea, 0, pa); b->getPoint(eb, 0, pb);
        if (!(pa[0] == pb[0] && pa[1] == pb[1] && pa[2] == pb[2]) || apf::getMdsIndex(a, ea) != apf::getMdsIndex(b, eb)) ++diffs;
      }
      a->end(ia); b->end(ib);
      for (int d = 1; d <= mdim; d += mdim - 1) {
        ia = a->begin(d); ib = b->begin(d);
        while ((ea = a->iterate(ia)) && (eb = b->iterate(ib))) {
          apf::Downward va, vb;
          const int na = a->getDownward(ea, 0, va), nb = b->getDownward(eb, 0, vb);
          if (na != nb) { ++diffs; continue; }
          for (int i = 0; i < na; ++i) if (apf::getMdsIndex(a, va[i]) != apf::getMdsIndex(b, vb[i])) ++diffs;
        }
        a->end(ia); b->end(ib);
      }
    }
    out[6] = (double)diffs;
    bad = diffs != 0 || out[7] != ref_max_len;   /* the longest metric edge left (test/aniso_adapt.h:65-74 asks <= MAXLENGTH once converged) */
  }
  for (int run = 0; run < 2; ++run)
    if (mesh[run]) { mesh[run]->destroyNative(); apf::destroyMesh(mesh[run]); }
  return bad;
}
