/* ORACLE — test infrastructure only (see icp_oracle.cpp header).  C ABI so tests reach it by ctypes. */
#ifndef LS_ORACLE_ICP_H_
#define LS_ORACLE_ICP_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lso_icp_params {
  int max_iterations;   /* CounterTransformationChecker.maxIterationCount   icp_default.yaml:22-23 */
  float trim_ratio;     /* TrimmedDistOutlierFilter.ratio                   icp_default.yaml:14-16 */
  int use_differential; /* DifferentialTransformationChecker on/off         icp_default.yaml:24-27 */
  float min_diff_rot;   /* minDiffRotErr  [rad] */
  float min_diff_trans; /* minDiffTransErr [m]  */
  int smooth_length;    /* smoothLength */
  int num_threads;      /* OpenMP threads over the query loop (libnabo's optional mode) */
} lso_icp_params;

typedef struct lso_icp_stats {
  int iterations;
  int converged;        /* stopped by the differential checker */
  int max_iter_reached; /* stopped by the counter (not an error) */
  int last_kept;        /* matches with weight 1 in the last iteration */
  float last_limit;     /* trimmed-distance limit (squared metres) of the last iteration */
  float used_ratio;     /* last_kept / n  (upstream pointUsedRatio) */
} lso_icp_stats;

void lso_default_params(lso_icp_params* p);
void lso_mean(const float* ref4, int m, float mu[3]);
void lso_nn_brute(const float* q3, int n, const float* ref_xyz3, int m, int32_t* ids, float* d2);
void lso_nn_kdtree(const float* q3, int n, const float* ref_xyz3, int m, int32_t* ids, float* d2, int num_threads);
float lso_trim_limit(const float* d2, int n, float ratio, int* n_finite_out);
void lso_transform_points(const float T[16], const float* in4, int n, float* out4);
void lso_transform_cloud(const float T[16], const float* in4, const float* nin, int nstride, int n,
                         float* out4, float* nout3);
int lso_check_rigid(const float T[16]);
void lso_correct_rigid(const float Tin[16], float Tout[16]);
void lso_normal_equations(const float* step4, int n, const float* ref_c3, const float* nrm, int nstride,
                          const int32_t* ids, const float* d2, float limit, double A[36], double b[6],
                          int* kept_out, double* A_f64, double* b_f64);
int lso_solve_step(const double A[36], const double b[6], float T_step[16], double x_out[6]);
void lso_mat4_mul(const float* A, const float* B, float* C);
void lso_sincos(double x, double* s, double* c);
void lso_knn_self(const float* feat4, int n, int k, int32_t* ids, float* d2, int num_threads);
void lso_knn_normals(const float* feat4, int n, int k, float* out3, int num_threads);
int lso_icp(const float* reading4, int n, const float* ref4, const float* ref_normals, int nstride, int m,
            const float T0[16], const lso_icp_params* prm, float T_out[16], lso_icp_stats* stats,
            int32_t* ids_hist, float* d2_last, float* T_iter_hist);

#ifdef __cplusplus
}
#endif
#endif
