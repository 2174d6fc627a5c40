/* ma_oracle.h -- TEST INFRASTRUCTURE (see ma_oracle.c).  CPU restatement of the
 * SCOREC/core MeshAdapt marking / quality arithmetic; never used by the product. */
#ifndef MA_ORACLE_H
#define MA_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* metric kinds.  (ma, mb) argument pairs:
   IDENTITY: (NULL, NULL)            ma::IdentitySizeField      maSize.cc:49-92
   ISO:      (s[nv], NULL)           IsoSizeField/IsoUserField  maSize.cc:581-616
   ANISO:    (h[nv][3], R[nv][9])    AnisoSizeField             maSize.cc:364-453
   LOGM:     (NULL, logM[nv][9])     LogAnisoSizeField          maSize.cc:455-579
   R and logM row-major, frame vectors in the COLUMNS of R. */
enum { MAO_IDENTITY = 0, MAO_ISO = 1, MAO_ANISO = 2, MAO_LOGM = 3 };

double mao_det3(const double A[3][3]);
int mao_eigen_qr(const double a[3][3], double l[3][3], double q[3][3], int* iters);
int mao_eigen(const double A[3][3], double vecs[3][3], double vals[3]);
void mao_transform_aniso(const double h[3], const double R[3][3], double Q[3][3]);
int mao_transform_logm(const double logM[3][3], double Q[3][3]);
void mao_logm_from_frame(int variant, const double h[3], const double R[3][3], double out[3][3]);

double mao_edge_length(int kind, const double* xyz, const double* ma, const double* mb,
                       const int32_t* ev, int* status);
int mao_vertex_transform(int kind, const double* ma, const double* mb, int32_t v, double Q[3][3]);
double mao_tet_quality(int kind, const double* xyz, const double* ma, const double* mb,
                       const int32_t* tv, int use_max, int* status);
double mao_tri_quality(int kind, const double* xyz, const double* ma, const double* mb,
                       const int32_t* tv, int use_max, int* status);
int mao_tri_qualities(int kind, const double* xyz, const double* ma, const double* mb,
                      int64_t nt, const int32_t* tri_v, int use_max, double* out);
int mao_prism_ok(const double* xyz, const int32_t* pv, int* good_codes);
int mao_pyramid_ok(const double* xyz, const int32_t* pv, int* good_rotation);

double mao_tet_weight(int kind, const double* xyz, const double* ma, const double* mb, const int32_t* tv, int* status);
double mao_clamp(double x, double max, double min);
int mao_sliver_code(int kind, const double* xyz, const double* ma, const double* mb, const int32_t* tv,
                    const int32_t* face0_v, double good_quality, int* status);
void mao_match_sliver(int code, int* rotation, int* code_index);
int mao_sliver_codes(int kind, const double* xyz, const double* ma, const double* mb, int64_t nt, const int32_t* tet_v,
                     const int32_t* face0_v, double good_quality, int32_t* codes, int32_t* match);
double mao_tri_weight(int kind, const double* xyz, const double* ma, const double* mb, const int32_t* tv, int* status);
int mao_tri_weights(int kind, const double* xyz, const double* ma, const double* mb,
                    int64_t nt, const int32_t* tri_v, double w_max, double w_min, double* out);
int mao_tet_weights(int kind, const double* xyz, const double* ma, const double* mb,
                    int64_t nt, const int32_t* tet_v, double w_max, double w_min, double* out);
void mao_split_vertex(int kind, const double* xyz, const double* ma, const double* mb, const int32_t* ev,
                      double* out_xyz, double* out_a, double* out_b);
void mao_split_vertices(int kind, const double* xyz, const double* ma, const double* mb, int64_t n,
                        const int32_t* edge_v, double* out_xyz, double* out_a, double* out_b);

int mao_edge_lengths(int kind, const double* xyz, const double* ma, const double* mb,
                     int64_t ne, const int32_t* edge_v, double* out);
int mao_tet_qualities(int kind, const double* xyz, const double* ma, const double* mb,
                      int64_t nt, const int32_t* tet_v, int use_max, double* out);
void mao_vertex_transforms(int kind, const double* ma, const double* mb, int64_t nv, double* Q9);
int64_t mao_mark_entities(int64_t n, const double* value, int cmp, double thr,
                          int32_t* flags, const uint8_t* owned,
                          int32_t true_flag, int32_t set_false_flag, int32_t all_false_flags);
double mao_min_quality(int64_t n, const double* q);
double mao_max_length(int64_t n, const double* len, const uint8_t* owned);

#ifdef __cplusplus
}
#endif
#endif
