// Pose-graph solve on the device (K5 linearise, K6 Gauss-Newton update), C ABI ls_pg_* of include/ls_b200.h.
//
// Replaces what IncrementalEstimator asks of gtsam::ISAM2 (reference laser_slam/src/incremental_estimator.cpp:
// 151-163 estimate, 165-266 estimateAndRemove, 268-291 registerPrior): factors are the ExpressionFactor<SE3>s built
// by LaserTrack::makeMeasurementFactor / makeRelativeMeasurementFactor (reference laser_slam/src/laser_track.cpp:
// 431-458) with Diagonal or Robust(Cauchy(1)) noise (laser_track.cpp:37-64).  Conventions are those of
// oracle/posegraph_oracle.py ([translation; rotation-vector] tangent, decoupled Local, t += dt, R <- R Exp(dr)).
//
// Structure exploited: a laser_slam graph is, per track, a CHAIN (prior on the first pose, odometry + ICP factors
// between consecutive poses) plus a few loop closures.  Poses are ordered track by track, so
//   H = H_c + U^T U,   H_c block-tridiagonal (chain + priors),   U = whitened Jacobians of the "extra" factors
// and the update solves H d = -g exactly through the Woodbury identity:
//   y = H_c^-1 (-g),  Z = H_c^-1 U^T,  (I + U Z) w = U y,  d = y - Z w,
// with H_c^-1 applied to all 6 #LC + 1 right-hand sides at once by block cyclic reduction (log2 P parallel levels).
// All arithmetic is float64; every sum that enters the solve has a fixed order (deterministic); only the reported cost
// statistic is an atomic sum.  Not HBM-bound (a few MB per iteration,
// SURVEY.md §8d): the cost is latency -- kernel launches and dependent levels -- not bandwidth.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/ls_b200.h"

namespace {

struct FactorDev {
  int type, robust, fix_a, extra;  // extra: index among border factors or -1
  int ia, ib;                      // pose indices (ia = -1: factor does not depend on node a)
  int chain;                       // 1: couples consecutive poses ia+1 == ib of one track
  int pad;
  double meas[7], sigma[6], fixed_a[7];
};

// ---------------------------------------------------------------- small dense helpers (row-major 3x3 / 6x6)
__device__ __forceinline__ void quat_to_R(const double* q, double* R) {
  const double n = 1.0 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double w = q[0] * n, x = q[1] * n, y = q[2] * n, z = q[3] * n;
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
__device__ __forceinline__ void mat3_tmul(const double* A, const double* B, double* C) {  // A^T B
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
__device__ __forceinline__ void so3_log(const double* R, double* w) {
  const double vx = 0.5 * (R[7] - R[5]), vy = 0.5 * (R[2] - R[6]), vz = 0.5 * (R[3] - R[1]);
  const double s = sqrt(vx * vx + vy * vy + vz * vz);
  const double c = 0.5 * (R[0] + R[4] + R[8] - 1.0);
  const double th = atan2(s, c);
  const double scale = s < 1e-8 ? 1.0 + th * th / 6.0 : th / s;
  w[0] = vx * scale; w[1] = vy * scale; w[2] = vz * scale;
}
__device__ __forceinline__ void jr_inv(const double* p, double* J) {
  const double th2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
  const double th = sqrt(th2);
  const double c = th < 1e-6 ? 1.0 / 12.0 : 1.0 / th2 - (1.0 + cos(th)) / (2.0 * th * sin(th));
  const double K[9] = {0, -p[2], p[1], p[2], 0, -p[0], -p[1], p[0], 0};
  double K2[9];
  mat3_mul(K, K, K2);
  for (int i = 0; i < 9; ++i) J[i] = 0.5 * K[i] + c * K2[i];
  J[0] += 1.0; J[4] += 1.0; J[8] += 1.0;
}

// ---------------------------------------------------------------- K5: linearise every factor
__global__ void pg_linearize_kernel(int F, const FactorDev* __restrict__ fac, const double* __restrict__ poses,
                                    double* __restrict__ Ja, double* __restrict__ Jb, double* __restrict__ r,
                                    double* cost) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const FactorDev& fc = fac[f];
  double ja[36], jb[36], res[6];
  for (int i = 0; i < 36; ++i) { ja[i] = 0.0; jb[i] = 0.0; }
  double Rm[9];
  quat_to_R(fc.meas, Rm);
  const double* tm = fc.meas + 4;
  if (fc.type == 0) {  // prior on node b
    const double* X = poses + 7 * (size_t)fc.ib;
    double R[9], RE[9], w[3], Ji[9];
    quat_to_R(X, R);
    const double d[3] = {X[4] - tm[0], X[5] - tm[1], X[6] - tm[2]};
    for (int i = 0; i < 3; ++i) res[i] = Rm[i] * d[0] + Rm[3 + i] * d[1] + Rm[6 + i] * d[2];
    mat3_tmul(Rm, R, RE);
    so3_log(RE, w);
    jr_inv(w, Ji);
    for (int i = 0; i < 3; ++i) {
      res[3 + i] = w[i];
      for (int j = 0; j < 3; ++j) {
        jb[6 * i + j] = Rm[3 * j + i];
        jb[6 * (3 + i) + 3 + j] = Ji[3 * i + j];
      }
    }
  } else {
    const double* A = fc.fix_a ? fc.fixed_a : poses + 7 * (size_t)fc.ia;
    const double* B = poses + 7 * (size_t)fc.ib;
    double Ra[9], Rb[9], RmtRat[9], RE[9], w[3], Ji[9], tmp[9];
    quat_to_R(A, Ra);
    quat_to_R(B, Rb);
    const double d[3] = {B[4] - A[4], B[5] - A[5], B[6] - A[6]};
    double v[3];
    for (int i = 0; i < 3; ++i) v[i] = Ra[i] * d[0] + Ra[3 + i] * d[1] + Ra[6 + i] * d[2];
    const double u[3] = {v[0] - tm[0], v[1] - tm[1], v[2] - tm[2]};
    for (int i = 0; i < 3; ++i) res[i] = Rm[i] * u[0] + Rm[3 + i] * u[1] + Rm[6 + i] * u[2];
    // RmtRat = Rm^T Ra^T
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) RmtRat[3 * i + j] = Rm[i] * Ra[3 * j] + Rm[3 + i] * Ra[3 * j + 1] + Rm[6 + i] * Ra[3 * j + 2];
    mat3_mul(RmtRat, Rb, RE);
    so3_log(RE, w);
    jr_inv(w, Ji);
    const double Kv[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
    double RmtKv[9], RbtRa[9], JiRbtRa[9];
    mat3_tmul(Rm, Kv, RmtKv);
    mat3_tmul(Rb, Ra, RbtRa);
    mat3_mul(Ji, RbtRa, JiRbtRa);
    (void)tmp;
    for (int i = 0; i < 3; ++i) {
      res[3 + i] = w[i];
      for (int j = 0; j < 3; ++j) {
        jb[6 * i + j] = RmtRat[3 * i + j];
        jb[6 * (3 + i) + 3 + j] = Ji[3 * i + j];
        if (!fc.fix_a) {
          ja[6 * i + j] = -RmtRat[3 * i + j];
          ja[6 * i + 3 + j] = RmtKv[3 * i + j];
          ja[6 * (3 + i) + 3 + j] = -JiRbtRa[3 * i + j];
        }
      }
    }
  }
  double e2 = 0.0;
  for (int i = 0; i < 6; ++i) {
    const double inv = 1.0 / fc.sigma[i];
    res[i] *= inv;
    e2 += res[i] * res[i];
    for (int j = 0; j < 6; ++j) { ja[6 * i + j] *= inv; jb[6 * i + j] *= inv; }
  }
  double c = 0.5 * e2;
  if (fc.robust) {
    const double sw = sqrt(1.0 / (1.0 + e2));
    c = 0.5 * log1p(e2);
    for (int i = 0; i < 6; ++i) res[i] *= sw;
    for (int i = 0; i < 36; ++i) { ja[i] *= sw; jb[i] *= sw; }
  }
  for (int i = 0; i < 36; ++i) { Ja[36 * (size_t)f + i] = ja[i]; Jb[36 * (size_t)f + i] = jb[i]; }
  for (int i = 0; i < 6; ++i) r[6 * (size_t)f + i] = res[i];
  atomicAdd(cost, c);
}

// ---------------------------------------------------------------- K5b: block-tridiagonal H_c and gradient
// thread per pose; incident factors through a CSR list (fixed order => deterministic sums)
__global__ void pg_assemble_kernel(int P, const int* __restrict__ inc_ptr, const int* __restrict__ inc_fac,
                                   const FactorDev* __restrict__ fac, const double* __restrict__ Ja,
                                   const double* __restrict__ Jb, const double* __restrict__ r,
                                   const int* __restrict__ damp, double* __restrict__ D, double* __restrict__ Bsub,
                                   double* __restrict__ g) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P) return;
  double d[36], b[36], gg[6];
  for (int i = 0; i < 36; ++i) { d[i] = 0.0; b[i] = 0.0; }
  for (int i = 0; i < 6; ++i) gg[i] = 0.0;
  for (int e = inc_ptr[k]; e < inc_ptr[k + 1]; ++e) {
    const int f = inc_fac[e];
    const FactorDev& fc = fac[f];
    const bool self_is_b = (fc.ib == k);
    const double* Js = (self_is_b ? Jb : Ja) + 36 * (size_t)f;
    const double* Jo = (self_is_b ? Ja : Jb) + 36 * (size_t)f;
    const double* rf = r + 6 * (size_t)f;
    for (int i = 0; i < 6; ++i) {
      double s = 0.0;
      for (int m = 0; m < 6; ++m) s += Js[6 * m + i] * rf[m];
      gg[i] += s;
    }
    if (fc.extra >= 0) continue;  // border factors live in U, not in H_c
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) {
        double s = 0.0;
        for (int m = 0; m < 6; ++m) s += Js[6 * m + i] * Js[6 * m + j];
        d[6 * i + j] += s;
      }
    if (fc.chain && self_is_b) {  // H[k][k-1] = Jb^T Ja
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
          double s = 0.0;
          for (int m = 0; m < 6; ++m) s += Js[6 * m + i] * Jo[6 * m + j];
          b[6 * i + j] += s;
        }
    }
  }
  if (damp[k]) {
    // Gauge damping on the first pose of a track that has no prior / fixed-node factor of its own (its prior
    // was removed when the track got linked, incremental_estimator.cpp:212-237): added to H only, never to g,
    // so the Gauss-Newton fixed point is unchanged while the chain block becomes positive definite.
    for (int i = 0; i < 3; ++i) { d[7 * i] += 1.0; d[7 * (3 + i)] += 4.0; }  // sigma 1 m / 0.5 rad: << any real factor
  }
  for (int i = 0; i < 36; ++i) { D[36 * (size_t)k + i] = d[i]; Bsub[36 * (size_t)k + i] = b[i]; }
  for (int i = 0; i < 6; ++i) g[6 * (size_t)k + i] = gg[i];
}

// ---------------------------------------------------------------- K6a': block cyclic reduction of H_c
// The sequential block sweeps above cost 2 x P dependent steps per right-hand side (75 ms per Gauss-Newton iteration at
// 5000 poses and 1201 right-hand sides).  Odd-even (cyclic) reduction solves the same block-tridiagonal SPD system in
// log2(P) levels, every level fully parallel over nodes and right-hand sides: level l (stride s = 2^l, nodes numbered
// m = 1..P) eliminates the nodes m = s (mod 2s) into their neighbours m +- s, which keep a Schur complement and a coupling
// of stride 2s; the back-substitution walks the levels down again.  Tracks are just zero couplings.
//   Dc[m]      current diagonal block of node m (final once the node is eliminated)
//   Di[m]      its inverse, formed at the node's elimination level
//   Ll[l][j]   coupling H^(l)[m][m - s] of node m = j << l at level l  (the right coupling is the transpose of the right
//              neighbour's left one: the matrix stays symmetric)
__device__ __forceinline__ bool inv6_spd(const double* A, double* X) {  // X = A^-1 through Cholesky; false: not positive definite
  double L[36];
  for (int i = 0; i < 36; ++i) L[i] = 0.0;
  for (int j = 0; j < 6; ++j) {
    double s = A[6 * j + j];
    for (int m = 0; m < j; ++m) s -= L[6 * j + m] * L[6 * j + m];
    if (!(s > 0.0)) return false;
    const double d = sqrt(s), id = 1.0 / d;
    L[6 * j + j] = d;
    for (int i = j + 1; i < 6; ++i) {
      double v = A[6 * i + j];
      for (int m = 0; m < j; ++m) v -= L[6 * i + m] * L[6 * j + m];
      L[6 * i + j] = v * id;
    }
  }
  for (int c = 0; c < 6; ++c) {  // solve L L^T x = e_c
    double y[6];
    for (int i = 0; i < 6; ++i) {
      double v = (i == c) ? 1.0 : 0.0;
      for (int m = 0; m < i; ++m) v -= L[6 * i + m] * y[m];
      y[i] = v / L[6 * i + i];
    }
    for (int i = 5; i >= 0; --i) {
      double v = y[i];
      for (int m = i + 1; m < 6; ++m) v -= L[6 * m + i] * X[6 * m + c];
      X[6 * i + c] = v / L[6 * i + i];
    }
  }
  return true;
}
__device__ __forceinline__ void mm6(const double* A, const double* B, double* C) {  // C = A B
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0.0;
      for (int m = 0; m < 6; ++m) s += A[6 * i + m] * B[6 * m + j];
      C[6 * i + j] = s;
    }
}
__device__ __forceinline__ void mmt6(const double* A, const double* B, double* C) {  // C = A B^T
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0.0;
      for (int m = 0; m < 6; ++m) s += A[6 * i + m] * B[6 * j + m];
      C[6 * i + j] = s;
    }
}
__device__ __forceinline__ void mtm6(const double* A, const double* B, double* C) {  // C = A^T B
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double s = 0.0;
      for (int m = 0; m < 6; ++m) s += A[6 * m + i] * B[6 * m + j];
      C[6 * i + j] = s;
    }
}

__global__ void cr_init_kernel(int P, const double* __restrict__ D, const double* __restrict__ Bsub, double* __restrict__ Dc,
                               double* __restrict__ L0) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P) return;
  for (int i = 0; i < 36; ++i) { Dc[36 * (size_t)k + i] = D[36 * (size_t)k + i]; L0[36 * (size_t)(k + 1) + i] = Bsub[36 * (size_t)k + i]; }
}
// nodes eliminated at this level: invert their (final) diagonal
__global__ void cr_invert_kernel(int P, int s, const double* __restrict__ Dc, double* __restrict__ Di, int* fail) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = (long long)s + (long long)t * 2 * s;  // m = s (mod 2s)
  if (m > P) return;
  double X[36];
  if (!inv6_spd(Dc + 36 * (size_t)(m - 1), X)) {
    *fail = 1;
    for (int i = 0; i < 36; ++i) X[i] = (i % 7 == 0) ? 1.0 : 0.0;
  }
  for (int i = 0; i < 36; ++i) Di[36 * (size_t)(m - 1) + i] = X[i];
}
// nodes kept at this level: Schur complement and the coupling of stride 2s
__global__ void cr_reduce_kernel(int P, int s, int lshift, double* __restrict__ Dc, const double* __restrict__ Di,
                                 const double* __restrict__ Lcur, double* __restrict__ Lnext) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = (long long)(t + 1) * 2 * s;  // m = 0 (mod 2s)
  if (m > P) return;
  double A[36], W[36], T[36];
  for (int i = 0; i < 36; ++i) A[i] = Dc[36 * (size_t)(m - 1) + i];
  const double* Lm = Lcur + 36 * (size_t)(m >> lshift);             // H[m][m-s]
  mm6(Lm, Di + 36 * (size_t)(m - s - 1), W);                        // W = Lm Dinv[m-s]   (m - s >= s >= 1 always)
  mmt6(W, Lm, T);
  for (int i = 0; i < 36; ++i) A[i] -= T[i];
  double* Ln = Lnext + 36 * (size_t)(m >> (lshift + 1));
  if (m - 2 * s >= 1) {
    mm6(W, Lcur + 36 * (size_t)((m - s) >> lshift), T);             // - W H[m-s][m-2s]
    for (int i = 0; i < 36; ++i) Ln[i] = -T[i];
  } else {
    for (int i = 0; i < 36; ++i) Ln[i] = 0.0;
  }
  if (m + s <= P) {
    const double* Lp = Lcur + 36 * (size_t)((m + s) >> lshift);     // H[m+s][m]
    mtm6(Lp, Di + 36 * (size_t)(m + s - 1), W);                     // V = Lp^T Dinv[m+s]
    mm6(W, Lp, T);
    for (int i = 0; i < 36; ++i) A[i] -= T[i];
  }
  for (int i = 0; i < 36; ++i) Dc[36 * (size_t)(m - 1) + i] = A[i];
}

// right-hand sides: Z[(k*6+i)*ncol + c] = -g (c == 0) or the c-th column of U^T
__global__ void cr_rhs_kernel(int P, int ncol, const FactorDev* __restrict__ fac, const int* __restrict__ extra_fac,
                              const double* __restrict__ Ja, const double* __restrict__ Jb, const double* __restrict__ g,
                              double* __restrict__ Z) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  if (c == 0) {
    for (int r = blockIdx.y; r < 6 * P; r += gridDim.y) Z[(size_t)r * ncol] = -g[r];
    return;
  }
  if (blockIdx.y != 0) return;
  const int f = extra_fac[(c - 1) / 6], row = (c - 1) % 6;
  const int ia = fac[f].ia, ib = fac[f].ib;
  for (int i = 0; i < 6; ++i) {
    if (ia >= 0) Z[((size_t)ia * 6 + i) * ncol + c] = Ja[36 * (size_t)f + 6 * row + i];
    Z[((size_t)ib * 6 + i) * ncol + c] = Jb[36 * (size_t)f + 6 * row + i];
  }
}
// unit right-hand sides of the marginals: column c = e_(pose[c / 6], c % 6)
__global__ void cr_unit_rhs_kernel(int ncol, const int* __restrict__ qpos, double* __restrict__ X) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncol) X[((size_t)qpos[c / 6] * 6 + c % 6) * ncol + c] = 1.0;
}
// forward, eliminated nodes: t = Dinv b (in place)
__global__ void cr_fwd_elim_kernel(int P, int s, int ncol, const double* __restrict__ Di, double* __restrict__ Z) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = (long long)s + (long long)blockIdx.y * 2 * s;
  if (c >= ncol || m > P) return;
  const double* X = Di + 36 * (size_t)(m - 1);
  double b[6], t[6];
  for (int i = 0; i < 6; ++i) b[i] = Z[((size_t)(m - 1) * 6 + i) * ncol + c];
  for (int i = 0; i < 6; ++i) {
    double v = 0.0;
    for (int j = 0; j < 6; ++j) v += X[6 * i + j] * b[j];
    t[i] = v;
  }
  for (int i = 0; i < 6; ++i) Z[((size_t)(m - 1) * 6 + i) * ncol + c] = t[i];
}
// forward, kept nodes: b -= H[m][m-s] t[m-s] + H[m][m+s] t[m+s]
__global__ void cr_fwd_keep_kernel(int P, int s, int lshift, int ncol, const double* __restrict__ Lcur, double* __restrict__ Z) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = (long long)(blockIdx.y + 1) * 2 * s;
  if (c >= ncol || m > P) return;
  double b[6];
  for (int i = 0; i < 6; ++i) b[i] = Z[((size_t)(m - 1) * 6 + i) * ncol + c];
  {
    const double* Lm = Lcur + 36 * (size_t)(m >> lshift);
    double t[6];
    for (int i = 0; i < 6; ++i) t[i] = Z[((size_t)(m - s - 1) * 6 + i) * ncol + c];
    for (int i = 0; i < 6; ++i) {
      double v = 0.0;
      for (int j = 0; j < 6; ++j) v += Lm[6 * i + j] * t[j];
      b[i] -= v;
    }
  }
  if (m + s <= P) {
    const double* Lp = Lcur + 36 * (size_t)((m + s) >> lshift);
    double t[6];
    for (int i = 0; i < 6; ++i) t[i] = Z[((size_t)(m + s - 1) * 6 + i) * ncol + c];
    for (int i = 0; i < 6; ++i) {
      double v = 0.0;
      for (int j = 0; j < 6; ++j) v += Lp[6 * j + i] * t[j];
      b[i] -= v;
    }
  }
  for (int i = 0; i < 6; ++i) Z[((size_t)(m - 1) * 6 + i) * ncol + c] = b[i];
}
// backward, nodes eliminated at this level: x = t - Dinv (H[m][m-s] x[m-s] + H[m][m+s] x[m+s])
__global__ void cr_bwd_kernel(int P, int s, int lshift, int ncol, const double* __restrict__ Di, const double* __restrict__ Lcur,
                              double* __restrict__ Z) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const long long m = (long long)s + (long long)blockIdx.y * 2 * s;
  if (c >= ncol || m > P) return;
  double r[6] = {0, 0, 0, 0, 0, 0};
  if (m - s >= 1) {
    const double* Lm = Lcur + 36 * (size_t)(m >> lshift);
    double x[6];
    for (int i = 0; i < 6; ++i) x[i] = Z[((size_t)(m - s - 1) * 6 + i) * ncol + c];
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) r[i] += Lm[6 * i + j] * x[j];
  }
  if (m + s <= P) {
    const double* Lp = Lcur + 36 * (size_t)((m + s) >> lshift);
    double x[6];
    for (int i = 0; i < 6; ++i) x[i] = Z[((size_t)(m + s - 1) * 6 + i) * ncol + c];
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) r[i] += Lp[6 * j + i] * x[j];
  }
  const double* X = Di + 36 * (size_t)(m - 1);
  for (int i = 0; i < 6; ++i) {
    double v = 0.0;
    for (int j = 0; j < 6; ++j) v += X[6 * i + j] * r[j];
    Z[((size_t)(m - 1) * 6 + i) * ncol + c] -= v;
  }
}

// ---------------------------------------------------------------- K6c: S = I + U Z (padded to n16), rhs = U y
__global__ void pg_border_kernel(int n, int n16, int ncol, const FactorDev* __restrict__ fac,
                                 const int* __restrict__ extra_fac, const double* __restrict__ Ja,
                                 const double* __restrict__ Jb, const double* __restrict__ Z, double* __restrict__ S,
                                 double* __restrict__ rhs) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;  // 0..n16 (col n16 => rhs)
  const int rowi = blockIdx.y;
  if (col > n16 || rowi >= n16) return;
  double v = 0.0;
  if (rowi < n && (col < n || col == n16)) {
    const int f = extra_fac[rowi / 6], rr = rowi % 6;
    const int zc = (col == n16) ? 0 : col + 1;
    const double* ja = Ja + 36 * (size_t)f + 6 * rr;
    const double* jb = Jb + 36 * (size_t)f + 6 * rr;
    const int ia = fac[f].ia, ib = fac[f].ib;
    if (ia >= 0)
      for (int i = 0; i < 6; ++i) v += ja[i] * Z[((size_t)ia * 6 + i) * ncol + zc];
    for (int i = 0; i < 6; ++i) v += jb[i] * Z[((size_t)ib * 6 + i) * ncol + zc];
  }
  if (col == n16) {
    rhs[rowi] = v;
  } else {
    if (rowi == col) v += 1.0;  // identity (also on the padding so the padded matrix stays SPD)
    S[(size_t)rowi * n16 + col] = v;
  }
}

// ---------------------------------------------------------------- K6d: dense SPD solve (right-looking, 16-wide panels)
constexpr int NB = 16;
__global__ void dense_panel_kernel(int n16, int p, double* __restrict__ A) {
  __shared__ double Lpp[NB][NB + 1];
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += blockDim.x) Lpp[e / NB][e % NB] = A[(size_t)(p + e / NB) * n16 + p + e % NB];
  __syncthreads();
  if (tid == 0) {
    for (int j = 0; j < NB; ++j) {
      double s = Lpp[j][j];
      for (int m = 0; m < j; ++m) s -= Lpp[j][m] * Lpp[j][m];
      const double dd = sqrt(s > 0.0 ? s : 1.0);
      Lpp[j][j] = dd;
      for (int i = j + 1; i < NB; ++i) {
        double v = Lpp[i][j];
        for (int m = 0; m < j; ++m) v -= Lpp[i][m] * Lpp[j][m];
        Lpp[i][j] = v / dd;
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int i = e / NB, j = e % NB;
    A[(size_t)(p + i) * n16 + p + j] = (j <= i) ? Lpp[i][j] : 0.0;
  }
  // rows below the diagonal block: L[i, p:p+NB] = A[i, p:p+NB] * Lpp^-T
  for (int i = p + NB + tid; i < n16; i += blockDim.x) {
    double row[NB];
    for (int j = 0; j < NB; ++j) {
      double v = A[(size_t)i * n16 + p + j];
      for (int m = 0; m < j; ++m) v -= row[m] * Lpp[j][m];
      row[j] = v / Lpp[j][j];
    }
    for (int j = 0; j < NB; ++j) A[(size_t)i * n16 + p + j] = row[j];
  }
}

__global__ void dense_update_kernel(int n16, int p, double* __restrict__ A) {
  // trailing update of the lower triangle: A[i][j] -= sum_k L[i][p+k] L[j][p+k], tiles of 16x16
  const int bi = blockIdx.y + (p / NB) + 1, bj = blockIdx.x + (p / NB) + 1;
  if (bj > bi || bi * NB >= n16) return;
  __shared__ double Li[NB][NB + 1], Lj[NB][NB + 1];
  const int ty = threadIdx.y, tx = threadIdx.x;
  Li[ty][tx] = A[(size_t)(bi * NB + ty) * n16 + p + tx];
  Lj[ty][tx] = A[(size_t)(bj * NB + ty) * n16 + p + tx];
  __syncthreads();
  double s = 0.0;
  for (int k = 0; k < NB; ++k) s += Li[ty][k] * Lj[tx][k];
  A[(size_t)(bi * NB + ty) * n16 + bj * NB + tx] -= s;
}

// single CTA: forward then backward substitution with the factor in A's lower triangle
__global__ void dense_solve_kernel(int n16, const double* __restrict__ A, double* __restrict__ x) {
  __shared__ double xj;
  for (int j = 0; j < n16; ++j) {
    if (threadIdx.x == 0) { x[j] = x[j] / A[(size_t)j * n16 + j]; xj = x[j]; }
    __syncthreads();
    for (int i = j + 1 + threadIdx.x; i < n16; i += blockDim.x) x[i] -= A[(size_t)i * n16 + j] * xj;
    __syncthreads();
  }
  for (int j = n16 - 1; j >= 0; --j) {
    if (threadIdx.x == 0) { x[j] = x[j] / A[(size_t)j * n16 + j]; xj = x[j]; }
    __syncthreads();
    for (int i = threadIdx.x; i < j; i += blockDim.x) x[i] -= A[(size_t)j * n16 + i] * xj;
    __syncthreads();
  }
}

// ---------------------------------------------------------------- K6e: d = y - Z w, retract (warp per pose)
__global__ void pg_update_kernel(int P, int ncol, const double* __restrict__ Z, const double* __restrict__ w,
                                 double* __restrict__ poses, unsigned long long* dmax_bits) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= P) return;
  double d[6];
  for (int i = 0; i < 6; ++i) {
    const double* zr = Z + ((size_t)k * 6 + i) * ncol;
    double s = 0.0;
    for (int c = 1 + lane; c < ncol; c += 32) s += zr[c] * w[c - 1];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    d[i] = zr[0] - s;
  }
  if (lane != 0) return;
  double* X = poses + 7 * (size_t)k;
  double m = 0.0;
  for (int i = 0; i < 6; ++i) m = fmax(m, fabs(d[i]));
  atomicMax(dmax_bits, (unsigned long long)__double_as_longlong(m));
  X[4] += d[0]; X[5] += d[1]; X[6] += d[2];
  const double th = sqrt(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
  const double half = 0.5 * th;
  const double sc = th < 1e-8 ? 0.5 - th * th / 48.0 : sin(half) / th;
  const double dq[4] = {cos(half), sc * d[3], sc * d[4], sc * d[5]};
  const double q[4] = {X[0], X[1], X[2], X[3]};
  double o[4] = {q[0] * dq[0] - q[1] * dq[1] - q[2] * dq[2] - q[3] * dq[3],
                 q[0] * dq[1] + q[1] * dq[0] + q[2] * dq[3] - q[3] * dq[2],
                 q[0] * dq[2] - q[1] * dq[3] + q[2] * dq[0] + q[3] * dq[1],
                 q[0] * dq[3] + q[1] * dq[2] - q[2] * dq[1] + q[3] * dq[0]};
  const double n = 1.0 / sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
  for (int i = 0; i < 4; ++i) X[i] = o[i] * n;
}

// ---------------------------------------------------------------- marginal covariances (gtsam::Marginals)
// Sigma_kk = (H^-1)_kk at the current estimate, for a chunk of requested poses.  Column c = 6*q + j is the unit vector
// e_(pose[q], j): X = H_c^-1 E through the cyclic-reduction levels (cr_unit_rhs_kernel + cr_solve), then the border correction
// through the same Woodbury identity as the update: Sigma = X - Z S^-1 (U X), of which only the 6x6 block of rows
// pose[q] is wanted.  X layout as Z: X[(k*6+i)*ncol + c].
// T = U X for the border rows (n of them, padded to n16 with zeros): T[row*ncol + c].  X is zero outside the column's track.
__global__ void pg_border_rhs_kernel(int n, int n16, int ncol, const FactorDev* __restrict__ fac, const int* __restrict__ extra_fac,
                                     const double* __restrict__ Ja, const double* __restrict__ Jb, const int* __restrict__ tb,
                                     const int* __restrict__ te, const double* __restrict__ X, double* __restrict__ T) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int rowi = blockIdx.y;
  if (c >= ncol || rowi >= n16) return;
  double v = 0.0;
  if (rowi < n) {
    const int f = extra_fac[rowi / 6], rr = rowi % 6;
    const int ia = fac[f].ia, ib = fac[f].ib;
    const int k0 = tb[c / 6], k1 = te[c / 6];
    const double* ja = Ja + 36 * (size_t)f + 6 * rr;
    const double* jb = Jb + 36 * (size_t)f + 6 * rr;
    if (ia >= k0 && ia < k1)
      for (int i = 0; i < 6; ++i) v += ja[i] * X[((size_t)ia * 6 + i) * ncol + c];
    if (ib >= k0 && ib < k1)
      for (int i = 0; i < 6; ++i) v += jb[i] * X[((size_t)ib * 6 + i) * ncol + c];
  }
  T[(size_t)rowi * ncol + c] = v;
}

// many right-hand sides against the dense factor (thread per column; W[row*ncol + c], coalesced across the threads)
__global__ void dense_solve_multi_kernel(int n16, int ncol, const double* __restrict__ A, double* __restrict__ W) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  for (int j = 0; j < n16; ++j) {
    double v = W[(size_t)j * ncol + c];
    for (int m = 0; m < j; ++m) v -= A[(size_t)j * n16 + m] * W[(size_t)m * ncol + c];
    W[(size_t)j * ncol + c] = v / A[(size_t)j * n16 + j];
  }
  for (int j = n16 - 1; j >= 0; --j) {
    double v = W[(size_t)j * ncol + c];
    for (int m = j + 1; m < n16; ++m) v -= A[(size_t)m * n16 + j] * W[(size_t)m * ncol + c];
    W[(size_t)j * ncol + c] = v / A[(size_t)j * n16 + j];
  }
}

// cov[q][i][j] = X[(pose q, i)][6q + j] - sum_r Z[(pose q, i)][1 + r] * W[r][6q + j]
__global__ void pg_marginal_block_kernel(int nq, int ncol, int n, int ncolZ, const int* __restrict__ qpos,
                                         const double* __restrict__ X, const double* __restrict__ Z, const double* __restrict__ W,
                                         double* __restrict__ cov) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nq * 36) return;
  const int q = e / 36, i = (e % 36) / 6, j = e % 6;
  const int kp = qpos[q], c = 6 * q + j;
  double v = X[((size_t)kp * 6 + i) * ncol + c];
  const double* zr = Z + ((size_t)kp * 6 + i) * ncolZ + 1;
  for (int r = 0; r < n; ++r) v -= zr[r] * W[(size_t)r * ncol + c];
  cov[e] = v;
}

struct HostFactor {
  ls_factor f;
  bool active;
};

}  // namespace

struct ls_pg {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  // graph (host)
  std::vector<uint64_t> keys;        // insertion order
  std::vector<uint32_t> tracks;
  std::vector<double> poses;         // 7 per pose, insertion order
  std::unordered_map<uint64_t, int> key_index;
  std::vector<HostFactor> factors;
  uint64_t launches = 0;
  // device buffers (grown on demand)
  size_t capF = 0, capP = 0, capZ = 0, capS = 0, capInc = 0, capE = 0;
  FactorDev* d_fac = nullptr;
  double *d_poses = nullptr, *d_Ja = nullptr, *d_Jb = nullptr, *d_r = nullptr, *d_D = nullptr, *d_B = nullptr,
         *d_g = nullptr, *d_Z = nullptr, *d_S = nullptr, *d_rhs = nullptr, *d_cost = nullptr;
  int *d_inc_ptr = nullptr, *d_inc_fac = nullptr, *d_extra = nullptr, *d_track_begin = nullptr, *d_fail = nullptr,
      *d_damp = nullptr;
  unsigned long long* d_dmax = nullptr;
  // marginals scratch
  size_t capX = 0, capW = 0, capQ = 0;
  double *d_X = nullptr, *d_W = nullptr, *d_cov = nullptr;
  double *d_Dc = nullptr, *d_Di = nullptr, *d_Ll = nullptr;  // cyclic reduction: current diagonals, inverses, per-level couplings
  size_t capCR = 0;
  int *d_qpos = nullptr, *d_qtb = nullptr, *d_qte = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
};

namespace {

int pg_fail(ls_pg* pg, int code, const char* msg) {
  if (pg) pg->err = msg;
  return code;
}

#define PGCU(call)                                                               \
  do {                                                                           \
    cudaError_t e_ = (call);                                                     \
    if (e_ != cudaSuccess) return pg_fail(pg, LS_ERR_CUDA, cudaGetErrorString(e_)); \
  } while (0)

template <typename T>
int grow(ls_pg* pg, T** p, size_t count) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  PGCU(cudaMalloc((void**)p, count * sizeof(T)));
  return LS_OK;
}

}  // namespace

extern "C" {

int ls_pg_create(int device, ls_pg** out) {
  if (!out) return LS_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return LS_ERR_CUDA;  // no CPU fallback
  if (device < 0 || device >= count) return LS_ERR_ARG;
  ls_pg* pg = new ls_pg();
  pg->device = device;
  cudaSetDevice(device);
  if (cudaStreamCreateWithFlags(&pg->stream, cudaStreamNonBlocking) != cudaSuccess) { delete pg; return LS_ERR_CUDA; }
  if (cudaMalloc((void**)&pg->d_cost, sizeof(double)) != cudaSuccess || cudaMalloc((void**)&pg->d_fail, sizeof(int)) != cudaSuccess ||
      cudaMalloc((void**)&pg->d_dmax, sizeof(unsigned long long)) != cudaSuccess || cudaEventCreate(&pg->e0) != cudaSuccess ||
      cudaEventCreate(&pg->e1) != cudaSuccess) {
    ls_pg_destroy(pg);
    return LS_ERR_NOMEM;
  }
  *out = pg;
  return LS_OK;
}

void ls_pg_destroy(ls_pg* pg) {
  if (!pg) return;
  cudaSetDevice(pg->device);
  if (pg->stream) cudaStreamSynchronize(pg->stream);
  void* bufs[] = {pg->d_fac, pg->d_poses, pg->d_Ja, pg->d_Jb, pg->d_r, pg->d_D, pg->d_B, pg->d_g, pg->d_Z,
                  pg->d_S, pg->d_rhs, pg->d_cost, pg->d_inc_ptr, pg->d_inc_fac, pg->d_extra, pg->d_track_begin, pg->d_fail,
                  pg->d_dmax, pg->d_damp, pg->d_X, pg->d_W, pg->d_cov, pg->d_qpos, pg->d_qtb, pg->d_qte, pg->d_Dc, pg->d_Di, pg->d_Ll};
  for (void* b : bufs)
    if (b) cudaFree(b);
  if (pg->e0) cudaEventDestroy(pg->e0);
  if (pg->e1) cudaEventDestroy(pg->e1);
  if (pg->stream) cudaStreamDestroy(pg->stream);
  delete pg;
}

const char* ls_pg_last_error(const ls_pg* pg) { return pg ? pg->err.c_str() : "null graph"; }
uint64_t ls_pg_launch_count(const ls_pg* pg) { return pg ? pg->launches : 0; }
int ls_pg_num_poses(const ls_pg* pg) { return pg ? (int)pg->keys.size() : LS_ERR_ARG; }
int ls_pg_num_factors(const ls_pg* pg) {
  if (!pg) return LS_ERR_ARG;
  int n = 0;
  for (const auto& f : pg->factors) n += f.active ? 1 : 0;
  return n;
}

int ls_pg_add_poses(ls_pg* pg, const uint64_t* keys, const uint32_t* track_ids, const double* poses7, int n) {
  if (!pg || !keys || !poses7 || n < 0) return pg_fail(pg, LS_ERR_ARG, "bad argument");
  for (int i = 0; i < n; ++i) {
    if (pg->key_index.count(keys[i])) return pg_fail(pg, LS_ERR_ARG, "duplicate pose key");
    pg->key_index[keys[i]] = (int)pg->keys.size();
    pg->keys.push_back(keys[i]);
    pg->tracks.push_back(track_ids ? track_ids[i] : 0u);
    pg->poses.insert(pg->poses.end(), poses7 + 7 * (size_t)i, poses7 + 7 * (size_t)i + 7);
  }
  return LS_OK;
}

int ls_pg_set_poses(ls_pg* pg, const uint64_t* keys, const double* poses7, int n) {
  if (!pg || !keys || !poses7 || n < 0) return pg_fail(pg, LS_ERR_ARG, "bad argument");
  for (int i = 0; i < n; ++i) {
    auto it = pg->key_index.find(keys[i]);
    if (it == pg->key_index.end()) return pg_fail(pg, LS_ERR_ARG, "unknown pose key");
    std::memcpy(&pg->poses[7 * (size_t)it->second], poses7 + 7 * (size_t)i, 7 * sizeof(double));
  }
  return LS_OK;
}

int ls_pg_add_factors(ls_pg* pg, const ls_factor* f, int n, uint64_t* out_indices) {
  if (!pg || !f || n < 0) return pg_fail(pg, LS_ERR_ARG, "bad argument");
  for (int i = 0; i < n; ++i) {
    if (f[i].type != LS_FACTOR_PRIOR && f[i].type != LS_FACTOR_BETWEEN) return pg_fail(pg, LS_ERR_ARG, "bad factor type");
    for (int s = 0; s < 6; ++s)
      if (!(f[i].sigma[s] > 0.0)) return pg_fail(pg, LS_ERR_ARG, "sigma must be positive");
    if (out_indices) out_indices[i] = pg->factors.size();
    pg->factors.push_back(HostFactor{f[i], true});
  }
  return LS_OK;
}

int ls_pg_remove_factors(ls_pg* pg, const uint64_t* idx, int n) {
  if (!pg || !idx || n < 0) return pg_fail(pg, LS_ERR_ARG, "bad argument");
  for (int i = 0; i < n; ++i) {
    if (idx[i] >= pg->factors.size() || !pg->factors[idx[i]].active) return pg_fail(pg, LS_ERR_ARG, "bad factor index");
    pg->factors[idx[i]].active = false;
  }
  return LS_OK;
}

int ls_pg_get_poses(const ls_pg* pg, uint64_t* out_keys, double* out_poses7, int* n) {
  if (!pg || !n) return LS_ERR_ARG;
  *n = (int)pg->keys.size();
  if (out_keys) std::memcpy(out_keys, pg->keys.data(), pg->keys.size() * sizeof(uint64_t));
  if (out_poses7) std::memcpy(out_poses7, pg->poses.data(), pg->poses.size() * sizeof(double));
  return LS_OK;
}

}  // extern "C"

namespace {
// gn_iters Gauss-Newton iterations, then -- when n_mk > 0 -- the marginal covariances of the poses mkeys[0..n_mk) at
// the resulting estimate (one more linearisation, no update).
int pg_run(ls_pg* pg, int gn_iters, ls_pg_stats* stats, const uint64_t* mkeys, int n_mk, double* cov_out) {
  if (stats) std::memset(stats, 0, sizeof(*stats));
  const int P = (int)pg->keys.size();
  if (P == 0 || (gn_iters == 0 && n_mk == 0)) return LS_OK;
  PGCU(cudaSetDevice(pg->device));
  // ---- ordering: track by track, insertion order inside a track (= time order in laser_slam)
  std::map<uint32_t, std::vector<int>> by_track;
  for (int i = 0; i < P; ++i) by_track[pg->tracks[i]].push_back(i);
  std::vector<int> order, pos_of(P), track_begin;
  for (auto& kv : by_track) {
    track_begin.push_back((int)order.size());
    for (int i : kv.second) { pos_of[i] = (int)order.size(); order.push_back(i); }
  }
  track_begin.push_back(P);
  const int n_tracks = (int)track_begin.size() - 1;
  std::vector<int> track_of_pos(P);
  for (int t = 0; t < n_tracks; ++t)
    for (int k = track_begin[t]; k < track_begin[t + 1]; ++k) track_of_pos[k] = t;
  // ---- factor table
  std::vector<FactorDev> fd;
  std::vector<int> extra_fac;
  std::vector<char> anchored(n_tracks, 0);
  for (const auto& hf : pg->factors) {
    if (!hf.active) continue;
    const ls_factor& f = hf.f;
    FactorDev d;
    std::memset(&d, 0, sizeof(d));
    d.type = f.type;
    d.robust = f.robust;
    d.fix_a = (f.type == LS_FACTOR_BETWEEN) ? f.fix_a : 0;
    d.extra = -1;
    auto ita = pg->key_index.find(f.key_a), itb = pg->key_index.find(f.key_b);
    if (f.type == LS_FACTOR_PRIOR) {
      if (ita == pg->key_index.end()) return pg_fail(pg, LS_ERR_ARG, "prior on unknown key");
      d.ia = -1;
      d.ib = pos_of[ita->second];
      anchored[track_of_pos[d.ib]] = 1;
    } else {
      if (itb == pg->key_index.end() || (!d.fix_a && ita == pg->key_index.end()))
        return pg_fail(pg, LS_ERR_ARG, "between factor on unknown key");
      d.ib = pos_of[itb->second];
      d.ia = d.fix_a ? -1 : pos_of[ita->second];
      if (d.fix_a) {
        anchored[track_of_pos[d.ib]] = 1;
      } else if (d.ib == d.ia + 1 && track_of_pos[d.ia] == track_of_pos[d.ib]) {
        d.chain = 1;  // consecutive poses of one track: part of the block-tridiagonal H_c
      }
      if (!d.fix_a && !d.chain) {  // everything else (loop closures, reversed pairs) goes to the border
        if (d.ia == d.ib) return pg_fail(pg, LS_ERR_ARG, "between factor with identical nodes");
        d.extra = (int)extra_fac.size();
        extra_fac.push_back((int)fd.size());
      }
    }
    std::memcpy(d.meas, f.meas, sizeof(d.meas));
    std::memcpy(d.sigma, f.sigma, sizeof(d.sigma));
    std::memcpy(d.fixed_a, f.fixed_a, sizeof(d.fixed_a));
    fd.push_back(d);
  }
  const int F = (int)fd.size();
  if (F == 0) return LS_OK;
  // a track without prior / fixed-node factor must at least be tied to the rest by a border factor
  std::vector<int> damp(P, 0);
  {
    std::vector<char> linked(n_tracks, 0);
    for (int fi : extra_fac) {
      if (fd[fi].ia >= 0) linked[track_of_pos[fd[fi].ia]] = 1;
      linked[track_of_pos[fd[fi].ib]] = 1;
    }
    for (int t = 0; t < n_tracks; ++t)
      if (!anchored[t]) {
        if (!linked[t])
          return pg_fail(pg, LS_ERR_STATE, "a track has no prior, no fixed-node factor and no link to another track: "
                                           "the graph has a free gauge");
        damp[track_begin[t]] = 1;
      }
  }
  const int E = (int)extra_fac.size();
  const int n = 6 * E, n16 = ((n + NB - 1) / NB) * NB, ncol = n + 1;
  // incidence lists
  std::vector<int> inc_ptr(P + 1, 0), inc_fac;
  for (const auto& d : fd) { if (d.ia >= 0) ++inc_ptr[d.ia + 1]; ++inc_ptr[d.ib + 1]; }
  for (int k = 0; k < P; ++k) inc_ptr[k + 1] += inc_ptr[k];
  inc_fac.resize(inc_ptr[P]);
  {
    std::vector<int> cur(inc_ptr.begin(), inc_ptr.end() - 1);
    for (int f = 0; f < F; ++f) {
      if (fd[f].ia >= 0) inc_fac[cur[fd[f].ia]++] = f;
      inc_fac[cur[fd[f].ib]++] = f;
    }
  }
  std::vector<double> hp(7 * (size_t)P);
  for (int k = 0; k < P; ++k) std::memcpy(&hp[7 * (size_t)k], &pg->poses[7 * (size_t)order[k]], 7 * sizeof(double));
  // ---- device buffers
  int rc;
  if ((size_t)F > pg->capF) {
    const size_t cap = (size_t)F + F / 4 + 64;
    if ((rc = grow(pg, &pg->d_fac, cap)) || (rc = grow(pg, &pg->d_Ja, cap * 36)) || (rc = grow(pg, &pg->d_Jb, cap * 36)) ||
        (rc = grow(pg, &pg->d_r, cap * 6)))
      return rc;
    pg->capF = cap;
  }
  if ((size_t)P > pg->capP) {
    const size_t cap = (size_t)P + P / 4 + 64;
    if ((rc = grow(pg, &pg->d_poses, cap * 7)) || (rc = grow(pg, &pg->d_D, cap * 36)) || (rc = grow(pg, &pg->d_B, cap * 36)) ||
        (rc = grow(pg, &pg->d_g, cap * 6)) ||
        (rc = grow(pg, &pg->d_inc_ptr, cap + 1)) || (rc = grow(pg, &pg->d_track_begin, cap + 1)) ||
        (rc = grow(pg, &pg->d_damp, cap + 1)))
      return rc;
    pg->capP = cap;
  }
  // cyclic-reduction levels: level l holds one coupling per node m = j << l, j = 1 .. P >> l
  std::vector<size_t> lvl_off;
  size_t lvl_total = 0;
  int n_levels = 0;
  for (int l = 0; (1ll << l) <= (long long)P; ++l) {
    lvl_off.push_back(lvl_total);
    lvl_total += ((size_t)P >> l) + 1;
    ++n_levels;
  }
  lvl_off.push_back(lvl_total);
  lvl_total += 2;  // the "next level" slot the top level's (empty) reduction would write
  if (lvl_total > pg->capCR || (size_t)P > pg->capCR) {
    const size_t cap = lvl_total + lvl_total / 4 + 64;
    if ((rc = grow(pg, &pg->d_Dc, cap * 36)) || (rc = grow(pg, &pg->d_Di, cap * 36)) || (rc = grow(pg, &pg->d_Ll, cap * 36))) return rc;
    pg->capCR = cap;
  }
  if (inc_fac.size() > pg->capInc) {
    if ((rc = grow(pg, &pg->d_inc_fac, inc_fac.size() * 2 + 64))) return rc;
    pg->capInc = inc_fac.size() * 2 + 64;
  }
  if ((size_t)E + 1 > pg->capE) {
    if ((rc = grow(pg, &pg->d_extra, (size_t)E * 2 + 64))) return rc;
    pg->capE = (size_t)E * 2 + 64;
  }
  const size_t needZ = (size_t)P * 6 * ncol;
  if (needZ > pg->capZ) {
    if ((rc = grow(pg, &pg->d_Z, needZ + needZ / 4))) return rc;
    pg->capZ = needZ + needZ / 4;
  }
  const size_t needS = (size_t)n16 * n16 + n16 + 16;
  if (needS > pg->capS) {
    if ((rc = grow(pg, &pg->d_S, needS * 2)) || (rc = grow(pg, &pg->d_rhs, (size_t)n16 * 2 + 64))) return rc;
    pg->capS = needS * 2;
  }
  cudaStream_t st = pg->stream;
  PGCU(cudaMemcpyAsync(pg->d_fac, fd.data(), (size_t)F * sizeof(FactorDev), cudaMemcpyHostToDevice, st));
  PGCU(cudaMemcpyAsync(pg->d_poses, hp.data(), hp.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  PGCU(cudaMemcpyAsync(pg->d_inc_ptr, inc_ptr.data(), inc_ptr.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  PGCU(cudaMemcpyAsync(pg->d_inc_fac, inc_fac.data(), inc_fac.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  if (E) PGCU(cudaMemcpyAsync(pg->d_extra, extra_fac.data(), (size_t)E * sizeof(int), cudaMemcpyHostToDevice, st));
  PGCU(cudaMemcpyAsync(pg->d_track_begin, track_begin.data(), track_begin.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  PGCU(cudaMemcpyAsync(pg->d_damp, damp.data(), (size_t)P * sizeof(int), cudaMemcpyHostToDevice, st));
  PGCU(cudaMemsetAsync(pg->d_fail, 0, sizeof(int), st));
  cudaEvent_t e0 = pg->e0, e1 = pg->e1;
  cudaEventRecord(e0, st);
  double cost_first = 0.0, cost_last = 0.0, dmax_last = 0.0;
  // every right-hand side of `buf` (6P x ncols, column index fastest) through the factored levels, in place
  auto cr_solve = [&](double* buf, int ncols) {
    for (int l = 0; l < n_levels; ++l) {
      const long long sl = 1ll << l;
      const int n_elim = (int)((P - sl) / (2 * sl)) + 1, n_keep = (int)(P / (2 * sl));
      cr_fwd_elim_kernel<<<dim3((ncols + 127) / 128, n_elim), 128, 0, st>>>(P, (int)sl, ncols, pg->d_Di, buf);
      if (n_keep > 0)
        cr_fwd_keep_kernel<<<dim3((ncols + 127) / 128, n_keep), 128, 0, st>>>(P, (int)sl, l, ncols, pg->d_Ll + 36 * lvl_off[l], buf);
      pg->launches += n_keep > 0 ? 2 : 1;
    }
    for (int l = n_levels - 1; l >= 0; --l) {
      const long long sl = 1ll << l;
      const int n_elim = (int)((P - sl) / (2 * sl)) + 1;
      cr_bwd_kernel<<<dim3((ncols + 127) / 128, n_elim), 128, 0, st>>>(P, (int)sl, l, ncols, pg->d_Di, pg->d_Ll + 36 * lvl_off[l], buf);
      ++pg->launches;
    }
  };
  const int n_pass = gn_iters + (n_mk > 0 ? 1 : 0);  // the last pass of a marginals request only linearises and factors
  for (int it = 0; it < n_pass; ++it) {
    const bool update = it < gn_iters;
    PGCU(cudaMemsetAsync(pg->d_cost, 0, sizeof(double), st));
    PGCU(cudaMemsetAsync(pg->d_dmax, 0, sizeof(unsigned long long), st));
    pg_linearize_kernel<<<(F + 127) / 128, 128, 0, st>>>(F, pg->d_fac, pg->d_poses, pg->d_Ja, pg->d_Jb, pg->d_r, pg->d_cost);
    pg_assemble_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, pg->d_inc_ptr, pg->d_inc_fac, pg->d_fac, pg->d_Ja, pg->d_Jb, pg->d_r,
                                                        pg->d_damp, pg->d_D, pg->d_B, pg->d_g);
    // H_c^-1 [ -g | U^T ] by block cyclic reduction (K6a'): factor the levels, then every right-hand side through them
    cr_init_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, pg->d_D, pg->d_B, pg->d_Dc, pg->d_Ll + 36 * lvl_off[0]);
    for (int l = 0; l < n_levels; ++l) {
      const long long sl = 1ll << l;
      const int n_elim = (int)((P - sl) / (2 * sl)) + 1, n_keep = (int)(P / (2 * sl));
      cr_invert_kernel<<<(n_elim + 63) / 64, 64, 0, st>>>(P, (int)sl, pg->d_Dc, pg->d_Di, pg->d_fail);
      if (n_keep > 0)
        cr_reduce_kernel<<<(n_keep + 63) / 64, 64, 0, st>>>(P, (int)sl, l, pg->d_Dc, pg->d_Di, pg->d_Ll + 36 * lvl_off[l],
                                                            pg->d_Ll + 36 * lvl_off[l + 1]);
      pg->launches += n_keep > 0 ? 2 : 1;
    }
    PGCU(cudaMemsetAsync(pg->d_Z, 0, (size_t)P * 6 * ncol * sizeof(double), st));
    cr_rhs_kernel<<<dim3((ncol + 127) / 128, 256), 128, 0, st>>>(P, ncol, pg->d_fac, pg->d_extra, pg->d_Ja, pg->d_Jb, pg->d_g, pg->d_Z);
    cr_solve(pg->d_Z, ncol);
    pg->launches += 4;
    if (E) {
      pg_border_kernel<<<dim3((n16 + 1 + 127) / 128, n16), 128, 0, st>>>(n, n16, ncol, pg->d_fac, pg->d_extra, pg->d_Ja, pg->d_Jb,
                                                                        pg->d_Z, pg->d_S, pg->d_rhs);
      ++pg->launches;
      for (int p = 0; p < n16; p += NB) {
        dense_panel_kernel<<<1, 256, 0, st>>>(n16, p, pg->d_S);
        const int rem = (n16 - p - NB) / NB;
        if (rem > 0) dense_update_kernel<<<dim3(rem, rem), dim3(NB, NB), 0, st>>>(n16, p, pg->d_S);
        pg->launches += rem > 0 ? 2 : 1;
      }
      if (update) {
        dense_solve_kernel<<<1, 1024, 0, st>>>(n16, pg->d_S, pg->d_rhs);
        ++pg->launches;
      }
    }
    if (!update) break;
    pg_update_kernel<<<(P * 32 + 255) / 256, 256, 0, st>>>(P, ncol, pg->d_Z, pg->d_rhs, pg->d_poses, pg->d_dmax);
    ++pg->launches;
    if (it == 0 || it == gn_iters - 1) {
      double c;
      unsigned long long dm;
      PGCU(cudaMemcpyAsync(&c, pg->d_cost, sizeof(double), cudaMemcpyDeviceToHost, st));
      PGCU(cudaMemcpyAsync(&dm, pg->d_dmax, sizeof(dm), cudaMemcpyDeviceToHost, st));
      PGCU(cudaStreamSynchronize(st));
      if (it == 0) cost_first = c;
      cost_last = c;
      std::memcpy(&dmax_last, &dm, sizeof(double));
    }
  }
  // ---- marginal covariances at the estimate just linearised: chunks of requested poses
  if (n_mk > 0) {
    constexpr int kChunk = 64;
    std::vector<int> qpos(n_mk), qtb(n_mk), qte(n_mk);
    for (int q = 0; q < n_mk; ++q) {
      auto it = pg->key_index.find(mkeys[q]);
      if (it == pg->key_index.end()) return pg_fail(pg, LS_ERR_ARG, "marginal of an unknown key");
      qpos[q] = pos_of[it->second];
      qtb[q] = track_begin[track_of_pos[qpos[q]]];
      qte[q] = track_begin[track_of_pos[qpos[q]] + 1];
    }
    const int ncm = 6 * kChunk;
    const size_t needX = (size_t)P * 6 * ncm, needW = (size_t)(n16 > 0 ? n16 : 1) * ncm;
    if (needX > pg->capX) { if ((rc = grow(pg, &pg->d_X, needX))) return rc; pg->capX = needX; }
    if (needW > pg->capW) { if ((rc = grow(pg, &pg->d_W, needW))) return rc; pg->capW = needW; }
    if ((size_t)n_mk > pg->capQ) {
      const size_t cap = (size_t)n_mk + 64;
      if ((rc = grow(pg, &pg->d_qpos, cap)) || (rc = grow(pg, &pg->d_qtb, cap)) || (rc = grow(pg, &pg->d_qte, cap)) ||
          (rc = grow(pg, &pg->d_cov, cap * 36)))
        return rc;
      pg->capQ = cap;
    }
    PGCU(cudaMemcpyAsync(pg->d_qpos, qpos.data(), (size_t)n_mk * sizeof(int), cudaMemcpyHostToDevice, st));
    PGCU(cudaMemcpyAsync(pg->d_qtb, qtb.data(), (size_t)n_mk * sizeof(int), cudaMemcpyHostToDevice, st));
    PGCU(cudaMemcpyAsync(pg->d_qte, qte.data(), (size_t)n_mk * sizeof(int), cudaMemcpyHostToDevice, st));
    for (int q0 = 0; q0 < n_mk; q0 += kChunk) {
      const int nq = n_mk - q0 < kChunk ? n_mk - q0 : kChunk, nc = 6 * nq;
      PGCU(cudaMemsetAsync(pg->d_X, 0, (size_t)P * 6 * nc * sizeof(double), st));
      cr_unit_rhs_kernel<<<(nc + 127) / 128, 128, 0, st>>>(nc, pg->d_qpos + q0, pg->d_X);
      ++pg->launches;
      cr_solve(pg->d_X, nc);
      if (E) {
        pg_border_rhs_kernel<<<dim3((nc + 127) / 128, n16), 128, 0, st>>>(n, n16, nc, pg->d_fac, pg->d_extra, pg->d_Ja, pg->d_Jb,
                                                                          pg->d_qtb + q0, pg->d_qte + q0, pg->d_X, pg->d_W);
        dense_solve_multi_kernel<<<(nc + 63) / 64, 64, 0, st>>>(n16, nc, pg->d_S, pg->d_W);
        pg->launches += 2;
      }
      pg_marginal_block_kernel<<<(nq * 36 + 127) / 128, 128, 0, st>>>(nq, nc, E ? n : 0, ncol, pg->d_qpos + q0, pg->d_X, pg->d_Z,
                                                                      pg->d_W, pg->d_cov + 36 * (size_t)q0);
      ++pg->launches;
    }
    PGCU(cudaMemcpyAsync(cov_out, pg->d_cov, (size_t)n_mk * 36 * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  cudaEventRecord(e1, st);
  int fail = 0;
  PGCU(cudaMemcpyAsync(&fail, pg->d_fail, sizeof(int), cudaMemcpyDeviceToHost, st));
  PGCU(cudaMemcpyAsync(hp.data(), pg->d_poses, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGCU(cudaStreamSynchronize(st));
  PGCU(cudaGetLastError());
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (fail) return pg_fail(pg, LS_ERR_CONVERGENCE, "chain block not positive definite (under-constrained graph)");
  for (int k = 0; k < P; ++k)
    for (int i = 0; i < 7; ++i)
      if (!std::isfinite(hp[7 * (size_t)k + i])) return pg_fail(pg, LS_ERR_CONVERGENCE, "non-finite pose after update");
  for (int k = 0; k < P; ++k) std::memcpy(&pg->poses[7 * (size_t)order[k]], &hp[7 * (size_t)k], 7 * sizeof(double));
  if (stats) {
    stats->iterations = gn_iters;
    stats->n_poses = P;
    stats->n_factors = F;
    stats->n_border = E;
    stats->cost_first = cost_first;
    stats->cost_last = cost_last;
    stats->last_step_max = dmax_last;
    stats->device_ms = ms;
  }
  return LS_OK;
}
}  // namespace

extern "C" {

int ls_pg_optimize(ls_pg* pg, int gn_iters, ls_pg_stats* stats) {
  if (!pg || gn_iters < 0) return pg_fail(pg, LS_ERR_ARG, "bad argument");
  return pg_run(pg, gn_iters, stats, nullptr, 0, nullptr);
}

int ls_pg_marginals(ls_pg* pg, const uint64_t* keys, int n, double* out_cov36) {
  if (!pg || !keys || !out_cov36 || n < 0) return pg_fail(pg, LS_ERR_ARG, "bad argument");
  if (n == 0) return LS_OK;
  return pg_run(pg, 0, nullptr, keys, n, out_cov36);
}

}  // extern "C"
