// ref_shape_shim.cc -- test infrastructure: reaches the one class of this path that the reference keeps local to a source
// file, ma::ShortEdgeFixer (ma/maShape.cc:170-241), by compiling that source file INTO this translation unit, where it lies
// in the reference tree (nothing is copied), with its private members opened so the edge the fixer picked can be read back.
// Linked into oracle/_ref/libref_oracle.so next to ref_driver.cc; the archive's own maShape object is then not pulled in
// (every symbol it would define is already defined here).
#include <vector>
#include <map>
#include <set>
#include <string>
#include <sstream>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdint.h>
#define private public
#define protected public
#include REFO_MASHAPE_CC
#undef private
#undef protected
#include <apfMDS.h>

/* ref_driver.cc */
ma::Adapt* refo_internal_adapt(void* hd);
apf::Mesh2* refo_internal_mesh(void* hd);

/* ShortEdgeFixer::shouldApply on every element of the mesh, in iteration order, on the flag words the last refo_mark left
   on the Adapt (BAD_QUALITY marked).  short_edge[i] = MDS index of the edge the fixer would hand to its ShortEdgeRemover, -1
   when shouldApply returned false; elem_flags_out = the flag words afterwards (BAD_QUALITY cleared where the edge-length
   ratio is below maximumEdgeRatio).  Returns the number of elements, < 0 on error. */
extern "C" int64_t refo_short_edge(void* hd, double maximumEdgeRatio, int32_t* short_edge, int32_t* elem_flags_out)
{
  ma::Adapt* a = refo_internal_adapt(hd);
  apf::Mesh2* m = refo_internal_mesh(hd);
  if (!a || !m) return -1;
  const_cast<ma::Input*>(a->input)->maximumEdgeRatio = maximumEdgeRatio;
  ma::ShortEdgeFixer fixer(a);
  apf::MeshIterator* it = m->begin(m->getDimension());
  apf::MeshEntity* e;
  int64_t k = 0;
  while ((e = m->iterate(it))) {
    const bool apply = m->getType(e) == apf::Mesh::TET && fixer.shouldApply(e);
    short_edge[k] = apply ? (int32_t)apf::getMdsIndex(m, fixer.remover.edge) : -1;
    ++k;
  }
  m->end(it);
  it = m->begin(m->getDimension());
  k = 0;
  while ((e = m->iterate(it))) elem_flags_out[k++] = ma::getFlags(a, e);
  m->end(it);
  return k;
}
