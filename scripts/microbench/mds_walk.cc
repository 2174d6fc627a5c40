// mds_walk.cc -- where the host time of an MDS export goes (CPU only, no device): the public apf::Mesh2 calls the adapter
// makes per entity against reading MDS's arrays directly (mds/mds.h struct mds, mds/mds_tag.h struct mds_tag).
// Build: see scripts/microbench/Makefile.mds (links oracle/_ref/libscorec_ref.a, needs /root/reference).
#include <apfMDS.h>
#include <apfBox.h>
#include <apfMesh2.h>
#include <apf.h>
#include <gmi_null.h>
#include <gmi_mesh.h>
#include <lionPrint.h>
#include <PCU.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
extern "C" {
#include <mds_apf.h>
#include <mds_tag.h>
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv)
{
  int n = argc > 1 ? atoi(argv[1]) : 48;
  pcu::Init(&argc, &argv);
  pcu::PCU* P = new pcu::PCU;
  lion_set_verbosity(0);
  gmi_register_null();
  gmi_register_mesh();
  apf::Mesh2* m = apf::makeMdsBox(n, n, n, 1, 1, 1, true, P);
  apf::Field* sizes = apf::createFieldOn(m, "sizes", apf::VECTOR);
  apf::Field* frames = apf::createFieldOn(m, "frames", apf::MATRIX);
  apf::MeshTag* flags = m->createIntTag("ma_flags", 1);
  apf::MeshIterator* it;
  apf::MeshEntity* e;
  it = m->begin(0);
  while ((e = m->iterate(it))) { apf::setVector(sizes, e, 0, apf::Vector3(1, 2, 3)); apf::setMatrix(frames, e, 0, apf::Matrix3x3(1, 0, 0, 0, 1, 0, 0, 0, 1)); }
  m->end(it);
  const size_t nv = m->count(0), ne = m->count(1), nf = m->count(2), nt = m->count(3);
  printf("n=%d nv=%zu ne=%zu nf=%zu nt=%zu\n", n, nv, ne, nf, nt);
  double t0, t1;
  std::vector<apf::MeshEntity*> edges(ne), tets(nt), verts(nv);
  t0 = now_s();
  { size_t k = 0; it = m->begin(1); while ((e = m->iterate(it))) edges[k++] = e; m->end(it); }
  { size_t k = 0; it = m->begin(3); while ((e = m->iterate(it))) tets[k++] = e; m->end(it); }
  { size_t k = 0; it = m->begin(0); while ((e = m->iterate(it))) verts[k++] = e; m->end(it); }
  t1 = now_s(); printf("iterate all (v,e,t)            %8.2f ms\n", 1e3 * (t1 - t0));
  std::vector<int> ev(2 * ne), tv(4 * nt);
  t0 = now_s();
  for (size_t i = 0; i < ne; ++i) { apf::Downward dv; m->getDownward(edges[i], 0, dv); ev[2 * i] = apf::getMdsIndex(m, dv[0]); ev[2 * i + 1] = apf::getMdsIndex(m, dv[1]); }
  t1 = now_s(); printf("getDownward(edge,0)            %8.2f ms  %.1f ns/edge\n", 1e3 * (t1 - t0), 1e9 * (t1 - t0) / ne);
  t0 = now_s();
  for (size_t i = 0; i < nt; ++i) { apf::Downward dv; m->getDownward(tets[i], 0, dv); for (int j = 0; j < 4; ++j) tv[4 * i + j] = apf::getMdsIndex(m, dv[j]); }
  t1 = now_s(); printf("getDownward(tet,0)             %8.2f ms  %.1f ns/tet\n", 1e3 * (t1 - t0), 1e9 * (t1 - t0) / nt);
  t0 = now_s();
  long own = 0;
  for (size_t i = 0; i < ne; ++i) own += m->isOwned(edges[i]);
  for (size_t i = 0; i < nt; ++i) own += m->isOwned(tets[i]);
  t1 = now_s(); printf("isOwned(edges+tets)            %8.2f ms  (%ld)\n", 1e3 * (t1 - t0), own);
  std::vector<double> xyz(3 * nv), h(3 * nv), R(9 * nv);
  t0 = now_s();
  for (size_t i = 0; i < nv; ++i) { apf::Vector3 p; m->getPoint(verts[i], 0, p); xyz[3 * i] = p[0]; xyz[3 * i + 1] = p[1]; xyz[3 * i + 2] = p[2]; }
  t1 = now_s(); printf("getPoint(verts)                %8.2f ms  %.1f ns/v\n", 1e3 * (t1 - t0), 1e9 * (t1 - t0) / nv);
  t0 = now_s();
  for (size_t i = 0; i < nv; ++i) {
    apf::Vector3 hh; apf::Matrix3x3 RR;
    apf::getVector(sizes, verts[i], 0, hh); apf::getMatrix(frames, verts[i], 0, RR);
    for (int a = 0; a < 3; ++a) h[3 * i + a] = hh[a];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) R[9 * i + 3 * a + b] = RR[a][b];
  }
  t1 = now_s(); printf("getVector+getMatrix(verts)     %8.2f ms  %.1f ns/v\n", 1e3 * (t1 - t0), 1e9 * (t1 - t0) / nv);
  t0 = now_s();
  unsigned long long hs = 1469598103934665603ull;
  { const unsigned char* b = (const unsigned char*)R.data(); for (size_t i = 0; i < R.size() * 8; ++i) { hs ^= b[i]; hs *= 1099511628211ull; } }
  { const unsigned char* b = (const unsigned char*)h.data(); for (size_t i = 0; i < h.size() * 8; ++i) { hs ^= b[i]; hs *= 1099511628211ull; } }
  { const unsigned char* b = (const unsigned char*)xyz.data(); for (size_t i = 0; i < xyz.size() * 8; ++i) { hs ^= b[i]; hs *= 1099511628211ull; } }
  t1 = now_s(); printf("FNV-1a bytes of xyz+h+R        %8.2f ms  (%llx)\n", 1e3 * (t1 - t0), hs);
  /* flags through the public tag API */
  t0 = now_s();
  for (size_t i = 0; i < ne; ++i) { int f = (int)(i & 7); m->setIntTag(edges[i], flags, &f); }
  t1 = now_s(); printf("setIntTag(edges)               %8.2f ms  %.1f ns/edge\n", 1e3 * (t1 - t0), 1e9 * (t1 - t0) / ne);
  std::vector<int> ef(ne);
  t0 = now_s();
  for (size_t i = 0; i < ne; ++i) { int f = 0; if (m->hasTag(edges[i], flags)) m->getIntTag(edges[i], flags, &f); ef[i] = f; }
  t1 = now_s(); printf("hasTag+getIntTag(edges)        %8.2f ms  %.1f ns/edge\n", 1e3 * (t1 - t0), 1e9 * (t1 - t0) / ne);
  /* direct: struct mds_tag */
  struct mds_tag* tg = (struct mds_tag*)flags;
  std::vector<int> ef2(ne);
  t0 = now_s();
  {
    const int* data = (const int*)tg->data[MDS_EDGE];
    const unsigned char* has = tg->has[MDS_EDGE];
    for (size_t i = 0; i < ne; ++i) {
      const int idx = apf::getMdsIndex(m, edges[i]);
      ef2[i] = (has && (has[idx / 8] & (1 << (idx % 8)))) ? data[idx] : 0;
    }
  }
  t1 = now_s(); printf("direct tag read(edges)         %8.2f ms  %.1f ns/edge  same=%d\n", 1e3 * (t1 - t0), 1e9 * (t1 - t0) / ne, (int)(ef == ef2));
  /* one-level downward through the public API */
  std::vector<apf::MeshEntity*> faces(nf);
  { size_t k = 0; it = m->begin(2); while ((e = m->iterate(it))) faces[k++] = e; m->end(it); }
  std::vector<int> fe(3 * nf), tf(4 * nt);
  t0 = now_s();
  for (size_t i = 0; i < nf; ++i) { apf::Downward d; m->getDownward(faces[i], 1, d); for (int j = 0; j < 3; ++j) fe[3 * i + j] = apf::getMdsIndex(m, d[j]); }
  for (size_t i = 0; i < nt; ++i) { apf::Downward d; m->getDownward(tets[i], 2, d); for (int j = 0; j < 4; ++j) tf[4 * i + j] = apf::getMdsIndex(m, d[j]); }
  t1 = now_s(); printf("one-level down (tri->e, tet->f) %7.2f ms\n", 1e3 * (t1 - t0));
  m->destroyNative();
  apf::destroyMesh(m);
  return 0;
}
