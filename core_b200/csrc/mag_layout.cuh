// mag_layout.cuh -- device-side layout of the per-vertex arrays, shared by the kernel translation units.
#pragma once
#include "mag_internal.h"
#include <cuda_runtime.h>

namespace maglay {

// ------------------------------------------------------------------ vertex arrays: blocked chunk-major layout
// Every per-vertex array is stored in blocks of kVB = 32 vertices; inside a block the K 16-byte chunks of the record
// form K planes of 32 double2:   chunk k of vertex v lives at ((double2*)base)[(v / 32) * 32 K + 32 k + v % 32].
// Neighbouring threads work on neighbouring entities, whose vertex ids are close in any locality-preserving numbering
// (grid order for box meshes), so one warp-wide LDG.128 of chunk k touches a few cache lines instead of one line per
// lane as 96-byte-strided records would (the L1 data pipe, not HBM, is what these kernels saturate first), and all K
// chunks of a vertex sit at compile-time offsets from one address (one address computation per vertex, one DRAM page).
// Arrays are padded to a whole number of blocks (vpad()).
//   d_vedge  Aniso   K=6 {x,y} {z,h0} {h1,h2} {R00,R10} {R20,R01} {R11,R21}   (frame columns 0 and 1; column 2 is
//                    overwritten by orthogonalizeR before use, maSize.cc:94-121)
//            LogAniso K=6 {x,y} {z,M00} {M01,M02} {M10,M11} {M12,M20} {M21,M22}
//            Iso / Identity K=2 {x,y} {z,s}
//   d_vpos   K=2 {x,y} {z,det Q_v}
//   d_vq     K=5 {Q00,Q01} {Q02,Q10} {Q11,Q12} {Q20,Q21} {Q22,det Q_v}
constexpr int kVB = MAG_VBLOCK;
// Ownership rides in the connectivity: the sign bit of an entity's FIRST vertex id means "not owned by this part"
// (k_fold_owned, at export time), so the owned-only counters of markEntities (maAdapt.cc:316) cost no extra load.
constexpr int32_t kVidMask = 0x7fffffff;
template <int K>
__device__ __forceinline__ const double2* chunk_ptr(const double* __restrict__ base, int k, int64_t v)
{
  return reinterpret_cast<const double2*>(base) + ((size_t)(v / kVB) * (size_t)(K * kVB) + (size_t)(k * kVB) + (size_t)(v % kVB));
}
// hot loops: vertex ids are non-negative int32, so the block offset is one 32x32->64 multiply-add
template <int K>
__device__ __forceinline__ const double2* chunk_ptr(const double* __restrict__ base, int k, int32_t v)
{
  const unsigned u = (unsigned)v;
  return reinterpret_cast<const double2*>(base) + ((size_t)(u / kVB) * (size_t)(K * kVB) + (size_t)(k * kVB + (int)(u % kVB)));
}
template <int K>
__device__ __forceinline__ double2* chunk_ptr_w(double* __restrict__ base, int k, int64_t v)
{
  return reinterpret_cast<double2*>(base) + ((size_t)(v / kVB) * (size_t)(K * kVB) + (size_t)(k * kVB) + (size_t)(v % kVB));
}
struct Rec12 { double v[12]; };
__device__ __forceinline__ Rec12 load_rec12(const double* __restrict__ base, int32_t vid)
{
  Rec12 r;
#pragma unroll
  for (int i = 0; i < 6; ++i) { double2 t = __ldg(chunk_ptr<6>(base, i, vid)); r.v[2 * i] = t.x; r.v[2 * i + 1] = t.y; }
  return r;
}
__device__ __forceinline__ void load_rec4(const double* __restrict__ base, int32_t vid, double out[4])
{
  double2 a = __ldg(chunk_ptr<2>(base, 0, vid)), b = __ldg(chunk_ptr<2>(base, 1, vid));
  out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
}


} // namespace maglay
