// mag_smb.cu -- PUMI's native mesh format (.smb) straight into the flat arrays mag_set_mesh takes.  Host code.
//
// The reference loads an .smb by rebuilding the whole MDS database entity by entity (mds/mds_smb.c:562-600 read_smb,
// mds_create_entity per entity), and an adapter would then walk it again to export.  This reader skips the database: it parses
// the file and derives element -> vertex connectivity with the rules MDS itself uses, so the arrays come out in the reference's
// own order (entity index order = m->begin(d) order of a freshly loaded mesh, downward vertices as getDownward(e, 0, .) returns
// them).  core_b200/smb.py is the same reader in numpy (tests compare the two on files the reference wrote).
//
// File layout (mds/mds_smb.c; all integers unsigned 32-bit, everything big-endian, pcu/pcu_io.c:238-241):
//   header   magic, version (<= 6), dim, number of parts                                   (:120-133)
//   counts   entities per type in SMB order VERT EDGE TRI QUAD HEX PRIS PYR TET           (:25-35, :577)
//   conn     for every type but VERT: the ONE-LEVEL-DOWN adjacency (edge: 2 vertices, triangle: 3 edges, quad: 4 edges,
//            tet: 4 triangles, prism: tri + 3 quads + tri, pyramid: quad + 4 tris), indices within the down type (:158-183)
//   points   3 doubles per vertex, then (version >= 2) 2 parametric doubles per vertex     (:586-594)
//   remotes  part-boundary vertex links (struct mds_links, :95-113)
//   class    (model id, model dim) per entity of every type                                (:230-250)
//   tags     n headers {type int|double, components, name\0}; then per entity type, per tag: ids + values (:257-300, :448-473)
// Lower adjacencies (mds/mds.c:634-670 step_down / convert_down with the `convs` tables :62-180): entity i of dimension d-2 of
// an element is the entity its (d-1)-dimensional faces conv[2i] and conv[2i+1] have in common (mds.c:496-508 common_down).
#include "mag_internal.h"
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>

namespace {

enum { SMB_VERT, SMB_EDGE, SMB_TRI, SMB_QUAD, SMB_HEX, SMB_PRIS, SMB_PYR, SMB_TET, SMB_TYPES };
const int kDownDegree[SMB_TYPES] = {0, 2, 3, 4, 6, 5, 5, 4};

// mds/mds.c:62-132: pairs of (d-1)-faces whose common (d-2)-entity is entity i
const int T10[3][2] = {{2, 0}, {0, 1}, {1, 2}};
const int TET21[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 3}, {1, 2}, {2, 3}};
const int TET10[4][2] = {{2, 0}, {0, 1}, {1, 2}, {3, 4}};
const int W21[9][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 3}, {1, 2}, {2, 3}, {1, 4}, {2, 4}, {3, 4}};
const int W10[6][2] = {{0, 2}, {0, 1}, {1, 2}, {6, 8}, {6, 7}, {7, 8}};
const int P21[8][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {1, 4}, {1, 2}, {2, 3}, {3, 4}};
const int P10[5][2] = {{0, 3}, {0, 1}, {1, 2}, {2, 3}, {4, 5}};

struct Cursor {
  const unsigned char* p;
  size_t n, o;
  bool ok;
  bool need(size_t k) { if (o + k > n) ok = false; return ok; }
  uint32_t u4()
  {
    if (!need(4)) return 0;
    uint32_t v = ((uint32_t)p[o] << 24) | ((uint32_t)p[o + 1] << 16) | ((uint32_t)p[o + 2] << 8) | p[o + 3];
    o += 4;
    return v;
  }
  double f8()
  {
    if (!need(8)) return 0;
    uint64_t v = 0;
    for (int i = 0; i < 8; ++i) v = (v << 8) | p[o + i];
    o += 8;
    double d;
    memcpy(&d, &v, 8);
    return d;
  }
  std::string str()
  {
    std::string s;
    while (need(1) && p[o]) s.push_back((char)p[o++]);
    if (ok) ++o;
    return s;
  }
};

// the one entry two short lists have in common (first of a that occurs in b); -1: none
int common(const int32_t* a, int ka, const int32_t* b, int kb)
{
  for (int i = 0; i < ka; ++i)
    for (int j = 0; j < kb; ++j)
      if (a[i] == b[j]) return a[i];
  return -1;
}

}  // namespace

struct mag_smb {
  int dim, version, nparts;
  int64_t count[SMB_TYPES];
  std::vector<double> xyz;
  std::vector<int32_t> edge_v, tri_e, quad_e, tri_v, tet_v, prism_v, pyr_v;
  struct Tag { int components; bool is_double; std::vector<int32_t> ids; std::vector<double> values; };
  std::map<std::string, Tag> vertex_tags;          // double-valued vertex tags as stored
  std::map<std::string, std::vector<double> > dense;   // dense [nv][components] copies handed out by mag_smb_vertex_field
  std::string err;
};

extern "C" {

const char* mag_smb_last_error(const mag_smb* s) { return s ? s->err.c_str() : "null handle"; }

void mag_smb_free(mag_smb* s) { delete s; }

int mag_smb_read(const char* path, mag_smb** out)
{
  if (!path || !out) return MAG_ERR_ARG;
  *out = nullptr;
  mag_smb* s = new mag_smb;
  *out = s;                          // kept on failure too, so that the caller can read the error text (and must free it)
  auto fail = [&](const std::string& m) { s->err = m; return MAG_ERR_ARG; };
  FILE* f = fopen(path, "rb");
  if (!f) return fail(std::string("cannot open ") + path);
  std::vector<unsigned char> data;
  {
    unsigned char buf[1 << 16];
    size_t k;
    while ((k = fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + k);
    fclose(f);
  }
  Cursor r{data.data(), data.size(), 0, true};
  r.u4();                                           // magic
  s->version = (int)r.u4();
  s->dim = (int)r.u4();
  s->nparts = (int)r.u4();
  if (!r.ok || s->version > 6) return fail("not an .smb file of version <= 6");
  for (int t = 0; t < SMB_TYPES; ++t) s->count[t] = r.u4();
  const int64_t nv = s->count[SMB_VERT];
  if (nv >= MAG_MAX_ENTITIES) return fail("too many vertices");
  std::vector<int32_t> conn[SMB_TYPES];
  for (int t = 1; t < SMB_TYPES; ++t) {
    const size_t n = (size_t)s->count[t] * kDownDegree[t];
    if (!r.need(4 * n)) return fail("truncated connectivity");
    conn[t].resize(n);
    for (size_t i = 0; i < n; ++i) conn[t][i] = (int32_t)r.u4();
  }
  if (!r.need((size_t)nv * 24)) return fail("truncated coordinates");
  s->xyz.resize((size_t)nv * 3);
  for (size_t i = 0; i < s->xyz.size(); ++i) s->xyz[i] = r.f8();
  if (s->version >= 2) { if (!r.need((size_t)nv * 16)) return fail("truncated parametric coordinates"); r.o += (size_t)nv * 16; }
  {  // remotes: skipped (the vertex links of a multi-part file; the sweep's part-boundary lists are edge lists)
    const uint32_t npeers = r.u4();
    std::vector<uint32_t> cnt(npeers);
    for (uint32_t i = 0; i < npeers; ++i) r.u4();
    for (uint32_t i = 0; i < npeers; ++i) cnt[i] = r.u4();
    for (uint32_t i = 0; i < npeers; ++i) { if (!r.need(4 * (size_t)cnt[i])) return fail("truncated remotes"); r.o += 4 * (size_t)cnt[i]; }
  }
  for (int t = 0; t < SMB_TYPES; ++t) { if (!r.need(8 * (size_t)s->count[t])) return fail("truncated classification"); r.o += 8 * (size_t)s->count[t]; }
  const uint32_t ntags = r.u4();
  struct Head { int type, comps; std::string name; };
  std::vector<Head> heads(ntags);
  for (uint32_t i = 0; i < ntags; ++i) { heads[i].type = (int)r.u4(); heads[i].comps = (int)r.u4(); heads[i].name = r.str(); }
  for (int t = 0; t < SMB_TYPES && r.ok; ++t) {
    std::vector<uint32_t> sizes(ntags);
    for (uint32_t i = 0; i < ntags; ++i) sizes[i] = r.u4();
    for (uint32_t i = 0; i < ntags && r.ok; ++i) {
      const size_t cnt = sizes[i], nvals = cnt * (size_t)heads[i].comps;
      const bool keep = t == SMB_VERT && heads[i].type != 0 && cnt;
      if (!r.need(4 * cnt + (heads[i].type == 0 ? 4 : 8) * nvals)) return fail("truncated tag data");
      if (keep) {
        mag_smb::Tag& tag = s->vertex_tags[heads[i].name];
        tag.components = heads[i].comps;
        tag.is_double = true;
        tag.ids.resize(cnt);
        for (size_t k = 0; k < cnt; ++k) tag.ids[k] = (int32_t)r.u4();
        tag.values.resize(nvals);
        for (size_t k = 0; k < nvals; ++k) tag.values[k] = r.f8();
      } else {
        r.o += 4 * cnt + (heads[i].type == 0 ? 4 : 8) * nvals;
      }
    }
  }
  if (!r.ok) return fail("truncated file");

  // ---- element -> vertex connectivity, MDS's own derivation
  s->edge_v = conn[SMB_EDGE];
  s->tri_e = conn[SMB_TRI];
  s->quad_e = conn[SMB_QUAD];
  const int32_t* ev = s->edge_v.data();
  const int64_t ne = s->count[SMB_EDGE], ntri = s->count[SMB_TRI], nquad = s->count[SMB_QUAD];
  for (size_t i = 0; i < s->edge_v.size(); ++i)
    if (s->edge_v[i] < 0 || s->edge_v[i] >= nv) return fail("edge with a vertex index out of range");
  for (size_t i = 0; i < s->tri_e.size(); ++i)
    if (s->tri_e[i] < 0 || s->tri_e[i] >= ne) return fail("triangle with an edge index out of range");
  for (size_t i = 0; i < s->quad_e.size(); ++i)
    if (s->quad_e[i] < 0 || s->quad_e[i] >= ne) return fail("quad with an edge index out of range");
  s->tri_v.resize((size_t)ntri * 3);
  for (int64_t i = 0; i < ntri; ++i)
    for (int j = 0; j < 3; ++j) {
      const int32_t* ea = ev + 2 * s->tri_e[3 * i + T10[j][0]];
      const int32_t* eb = ev + 2 * s->tri_e[3 * i + T10[j][1]];
      if ((s->tri_v[3 * i + j] = common(ea, 2, eb, 2)) < 0) return fail("triangle whose edges share no vertex");
    }
  // faces of an element: (which face array, degree); the element's edges from its faces, its vertices from its edges
  auto element = [&](const std::vector<int32_t>& faces, int nfaces, const int* face_is_quad, const int (*p21)[2], int nedges,
                     const int (*p10)[2], int nverts, std::vector<int32_t>& out) -> bool {
    const int64_t n = (int64_t)faces.size() / nfaces;
    out.resize((size_t)n * nverts);
    for (int64_t i = 0; i < n; ++i) {
      const int32_t* fe[5];
      int fk[5];
      for (int j = 0; j < nfaces; ++j) {
        const int32_t fidx = faces[(size_t)i * nfaces + j];
        if (face_is_quad[j]) { if (fidx < 0 || fidx >= nquad) return false; fe[j] = s->quad_e.data() + 4 * (size_t)fidx; fk[j] = 4; }
        else { if (fidx < 0 || fidx >= ntri) return false; fe[j] = s->tri_e.data() + 3 * (size_t)fidx; fk[j] = 3; }
      }
      int32_t el_e[9];
      for (int k = 0; k < nedges; ++k)
        if ((el_e[k] = common(fe[p21[k][0]], fk[p21[k][0]], fe[p21[k][1]], fk[p21[k][1]])) < 0) return false;
      for (int k = 0; k < nverts; ++k)
        if ((out[(size_t)i * nverts + k] = common(ev + 2 * el_e[p10[k][0]], 2, ev + 2 * el_e[p10[k][1]], 2)) < 0) return false;
    }
    return true;
  };
  const int tet_faces[4] = {0, 0, 0, 0}, prism_faces[5] = {0, 1, 1, 1, 0}, pyr_faces[5] = {1, 0, 0, 0, 0};   // mds.c W2, P2
  if (!element(conn[SMB_TET], 4, tet_faces, TET21, 6, TET10, 4, s->tet_v)) return fail("tet whose faces do not close");
  if (!element(conn[SMB_PRIS], 5, prism_faces, W21, 9, W10, 6, s->prism_v)) return fail("prism whose faces do not close");
  if (!element(conn[SMB_PYR], 5, pyr_faces, P21, 8, P10, 5, s->pyr_v)) return fail("pyramid whose faces do not close");
  return MAG_OK;
}

int mag_smb_get(const mag_smb* s, mag_smb_arrays* a)
{
  if (!s || !a) return MAG_ERR_ARG;
  a->dim = s->dim; a->version = s->version; a->nparts = s->nparts;
  a->nv = s->count[SMB_VERT]; a->ne = s->count[SMB_EDGE]; a->ntri = s->count[SMB_TRI]; a->nquad = s->count[SMB_QUAD];
  a->nt = s->count[SMB_TET]; a->np = s->count[SMB_PRIS]; a->npy = s->count[SMB_PYR]; a->nhex = s->count[SMB_HEX];
  a->xyz = s->xyz.data();
  a->edge_v = s->edge_v.data(); a->tri_v = s->tri_v.data(); a->tet_v = s->tet_v.data();
  a->prism_v = s->prism_v.data(); a->pyr_v = s->pyr_v.data();
  return MAG_OK;
}

/* apf stores the vertex nodes of field <name> in the tag <name>_ver (apf/apfTagData.cc) */
int mag_smb_vertex_field(mag_smb* s, const char* name, int* components, const double** values)
{
  if (!s || !name || !components || !values) return MAG_ERR_ARG;
  std::map<std::string, mag_smb::Tag>::const_iterator it = s->vertex_tags.find(name);
  if (it == s->vertex_tags.end()) it = s->vertex_tags.find(std::string(name) + "_ver");
  if (it == s->vertex_tags.end()) { s->err = std::string("no double-valued vertex tag ") + name; return MAG_ERR_ARG; }
  const mag_smb::Tag& t = it->second;
  const size_t nv = (size_t)s->count[SMB_VERT];
  if (t.ids.size() != nv) { s->err = std::string("tag ") + name + " is not set on every vertex"; return MAG_ERR_ARG; }
  std::vector<double>& d = s->dense[it->first];
  if (d.empty()) {
    d.resize(nv * (size_t)t.components);
    for (size_t k = 0; k < nv; ++k) {
      if (t.ids[k] < 0 || (size_t)t.ids[k] >= nv) { s->err = "tag entry with a vertex index out of range"; d.clear(); return MAG_ERR_ARG; }
      memcpy(&d[(size_t)t.ids[k] * t.components], &t.values[k * (size_t)t.components], sizeof(double) * (size_t)t.components);
    }
  }
  *components = t.components;
  *values = d.data();
  return MAG_OK;
}

/* the file's part becomes the resident part of the context: mag_set_mesh (3-D) / mag_set_mesh_2d with the arrays as read */
int mag_set_mesh_smb(mag_ctx* c, const mag_smb* s)
{
  if (!c || !s) return MAG_ERR_ARG;
  if (s->count[SMB_HEX]) return mag_fail(c, MAG_ERR_ARG, "mag_set_mesh_smb: hexahedra are not supported");
  if (s->dim == 2) {
    if (s->count[SMB_QUAD]) return mag_fail(c, MAG_ERR_ARG, "mag_set_mesh_smb: quadrilateral elements are not supported");
    return mag_set_mesh_2d(c, s->count[SMB_VERT], s->xyz.data(), s->count[SMB_EDGE], s->edge_v.data(), s->count[SMB_TRI],
                           s->tri_v.data(), nullptr, nullptr);
  }
  if (s->dim != 3) return mag_fail(c, MAG_ERR_ARG, "mag_set_mesh_smb: mesh dimension %d", s->dim);
  return mag_set_mesh(c, s->count[SMB_VERT], s->xyz.data(), s->count[SMB_EDGE], s->edge_v.data(), s->count[SMB_TET], s->tet_v.data(),
                      s->count[SMB_PRIS], s->prism_v.data(), s->count[SMB_PYR], s->pyr_v.data(), nullptr, nullptr);
}

}  // extern "C"
