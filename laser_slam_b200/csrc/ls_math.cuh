// Scalar math shared by the CUDA kernels (and, compiled for the host with -ffp-contract=off, by the
// tests' CPU simulation of the search -- never by the product's run-time path).
//
// Every float/double operation here must be individually rounded: the library is compiled with
// nvcc -fmad=false so nothing is contracted into an FMA.  That, plus order-independent integer
// reductions elsewhere, is what makes the whole registration bit-reproducible.
//
// Replaces, on the reference's path, the arithmetic of libpointmatcher's
// PointToPlaneErrorMinimizer / RigidTransformation / TransformationCheckers that
// PointMatcher::ICP::compute runs (reference laser_slam/src/laser_track.cpp:496,
// laser_slam/configurations/icp_default.yaml:18-27).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define LS_HD __host__ __device__ __forceinline__
#define LS_HDN __host__ __device__
#else
#define LS_HD inline
#define LS_HDN inline
#endif

namespace ls {

// ---- float32 rigid transform, column-major T: x' = ((r00 x + r01 y) + r02 z) + tx ---------------
LS_HD void xform_point(const float* T, float x, float y, float z, float& ox, float& oy, float& oz) {
  float a, b, c, s;
  a = T[0] * x; b = T[4] * y; c = T[8] * z; s = a + b; s = s + c; ox = s + T[12];
  a = T[1] * x; b = T[5] * y; c = T[9] * z; s = a + b; s = s + c; oy = s + T[13];
  a = T[2] * x; b = T[6] * y; c = T[10] * z; s = a + b; s = s + c; oz = s + T[14];
}

// rotation only (normals descriptor): n' = (r00 x + r01 y) + r02 z
LS_HD void rotate_vec(const float* T, float x, float y, float z, float& ox, float& oy, float& oz) {
  float a, b, c, s;
  a = T[0] * x; b = T[4] * y; c = T[8] * z; s = a + b; ox = s + c;
  a = T[1] * x; b = T[5] * y; c = T[9] * z; s = a + b; oy = s + c;
  a = T[2] * x; b = T[6] * y; c = T[10] * z; s = a + b; oz = s + c;
}

// C = A*B, C(i,j) = (((a_i0 b_0j + a_i1 b_1j) + a_i2 b_2j) + a_i3 b_3j); C may alias A or B.
LS_HDN void mat4_mul(const float* A, const float* B, float* C) {
  float tmp[16];
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) {
      float s = A[i] * B[j * 4];
      float t = A[4 + i] * B[j * 4 + 1];
      s = s + t;
      t = A[8 + i] * B[j * 4 + 2];
      s = s + t;
      t = A[12 + i] * B[j * 4 + 3];
      s = s + t;
      tmp[j * 4 + i] = s;
    }
  for (int i = 0; i < 16; ++i) C[i] = tmp[i];
}

// squared distance, x -> y -> z, float32, no FMA (libnabo leaf arithmetic)
LS_HD float dist2(float qx, float qy, float qz, float px, float py, float pz) {
  const float dx = qx - px, dy = qy - py, dz = qz - pz;
  const float a = dx * dx, b = dy * dy, c = dz * dz;
  const float s = a + b;
  return s + c;
}

// ---- deterministic double sin/cos ---------------------------------------------------------------
LS_HDN void det_sincos(double x, double* s_out, double* c_out) {
  const double two_over_pi = 0.63661977236758134308;
  const double pio2_hi = 1.57079632673412561417e+00;
  const double pio2_lo = 6.07710050650619224932e-11;
  const double kf = floor(x * two_over_pi + 0.5);
  double r = x - kf * pio2_hi;
  r = r - kf * pio2_lo;
  const double z = r * r;
  double ps = -1.0 / 355687428096000.0;
  ps = ps * z + 1.0 / 1307674368000.0;
  ps = ps * z - 1.0 / 6227020800.0;
  ps = ps * z + 1.0 / 39916800.0;
  ps = ps * z - 1.0 / 362880.0;
  ps = ps * z + 1.0 / 5040.0;
  ps = ps * z - 1.0 / 120.0;
  ps = ps * z + 1.0 / 6.0;
  const double sr = r - r * (z * ps);
  double pc = -1.0 / 6402373705728000.0;
  pc = pc * z + 1.0 / 20922789888000.0;
  pc = pc * z - 1.0 / 87178291200.0;
  pc = pc * z + 1.0 / 479001600.0;
  pc = pc * z - 1.0 / 3628800.0;
  pc = pc * z + 1.0 / 40320.0;
  pc = pc * z - 1.0 / 720.0;
  pc = pc * z + 1.0 / 24.0;
  pc = pc * z - 0.5;
  const double cr = 1.0 + z * pc;
  const long long k = (long long)kf;
  switch ((int)(k & 3)) {
    case 0: *s_out = sr; *c_out = cr; break;
    case 1: *s_out = cr; *c_out = -sr; break;
    case 2: *s_out = -sr; *c_out = -cr; break;
    default: *s_out = -cr; *c_out = sr; break;
  }
}

LS_HD bool is_finite_d(double v) { return v == v && fabs(v) <= 1.7976931348623157e308; }
LS_HD bool is_finite_f(float v) { return v == v && fabsf(v) <= 3.402823466e38f; }

// ---- 6x6 SPD solve: column Cholesky, reciprocal pivots; false => rank deficient ------------------
LS_HDN bool chol6(const double* A, const double* b, double* x) {
  double L[36], inv[6], y[6];
  for (int i = 0; i < 36; ++i) L[i] = 0.0;
  for (int j = 0; j < 6; ++j) {
    double s = A[j * 6 + j];
    for (int k = 0; k < j; ++k) {
      const double t = L[j * 6 + k] * L[j * 6 + k];
      s = s - t;
    }
    if (!(s > 1e-10 * A[j * 6 + j]) || !is_finite_d(s)) return false;
    const double d = sqrt(s);
    L[j * 6 + j] = d;
    inv[j] = 1.0 / d;
    for (int i = j + 1; i < 6; ++i) {
      double v = A[i * 6 + j];
      for (int k = 0; k < j; ++k) {
        const double t = L[i * 6 + k] * L[j * 6 + k];
        v = v - t;
      }
      L[i * 6 + j] = v * inv[j];
    }
  }
  for (int i = 0; i < 6; ++i) {
    double v = b[i];
    for (int k = 0; k < i; ++k) {
      const double t = L[i * 6 + k] * y[k];
      v = v - t;
    }
    y[i] = v * inv[i];
  }
  for (int i = 5; i >= 0; --i) {
    double v = y[i];
    for (int k = i + 1; k < 6; ++k) {
      const double t = L[k * 6 + i] * x[k];
      v = v - t;
    }
    x[i] = v * inv[i];
  }
  return true;
}

// minimum-norm solution through a cyclic Jacobi eigen-decomposition (rank-deficient fallback)
LS_HDN void jacobi_pinv_solve6(const double* Ain, const double* b, double* x) {
  double a[36], v[36];
  for (int i = 0; i < 36; ++i) { a[i] = Ain[i]; v[i] = 0.0; }
  for (int i = 0; i < 6; ++i) v[i * 6 + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int p = 0; p < 6; ++p) {
      const double t = a[p * 6 + p] * a[p * 6 + p];
      dg = dg + t;
      for (int q = p + 1; q < 6; ++q) {
        const double u = a[p * 6 + q] * a[p * 6 + q];
        off = off + u;
      }
    }
    if (!(off > 1e-40 * dg)) break;
    for (int p = 0; p < 5; ++p)
      for (int q = p + 1; q < 6; ++q) {
        const double apq = a[p * 6 + q];
        if (apq == 0.0) continue;
        const double theta = (a[q * 6 + q] - a[p * 6 + p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 6; ++k) {
          const double akp = a[k * 6 + p], akq = a[k * 6 + q];
          a[k * 6 + p] = c * akp - s * akq;
          a[k * 6 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 6; ++k) {
          const double apk = a[p * 6 + k], aqk = a[q * 6 + k];
          a[p * 6 + k] = c * apk - s * aqk;
          a[q * 6 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 6; ++k) {
          const double vkp = v[k * 6 + p], vkq = v[k * 6 + q];
          v[k * 6 + p] = c * vkp - s * vkq;
          v[k * 6 + q] = s * vkp + c * vkq;
        }
      }
  }
  double lmax = 0.0;
  for (int i = 0; i < 6; ++i) lmax = fmax(lmax, fabs(a[i * 6 + i]));
  for (int i = 0; i < 6; ++i) x[i] = 0.0;
  for (int k = 0; k < 6; ++k) {
    const double lam = a[k * 6 + k];
    if (!(lam > 1e-10 * lmax)) continue;
    double proj = 0.0;
    for (int i = 0; i < 6; ++i) {
      const double t = v[i * 6 + k] * b[i];
      proj = proj + t;
    }
    const double coef = proj / lam;
    for (int i = 0; i < 6; ++i) {
      const double t = v[i * 6 + k] * coef;
      x[i] = x[i] + t;
    }
  }
}

// (rotation vector, translation) -> float 4x4 column-major; theta == 0 or NaN => rotation = I
LS_HDN void step_matrix(const double* x, float* T) {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double n2 = x[0] * x[0];
  double t1 = x[1] * x[1];
  n2 = n2 + t1;
  t1 = x[2] * x[2];
  n2 = n2 + t1;
  const double th = sqrt(n2);
  if (th > 0.0 && is_finite_d(th)) {
    const double ux = x[0] / th, uy = x[1] / th, uz = x[2] / th;
    double s, c;
    det_sincos(th, &s, &c);
    const double sx = s * ux, sy = s * uy, sz = s * uz;
    const double omc = 1.0 - c;
    const double cx = omc * ux, cy = omc * uy, cz = omc * uz;
    double tmp;
    tmp = cx * uy; R[1] = tmp - sz; R[3] = tmp + sz;
    tmp = cx * uz; R[2] = tmp + sy; R[6] = tmp - sy;
    tmp = cy * uz; R[5] = tmp - sx; R[7] = tmp + sx;
    tmp = cx * ux; R[0] = tmp + c;
    tmp = cy * uy; R[4] = tmp + c;
    tmp = cz * uz; R[8] = tmp + c;
  }
  for (int r = 0; r < 3; ++r) {
    for (int cc = 0; cc < 3; ++cc) T[cc * 4 + r] = (float)R[r * 3 + cc];
    T[12 + r] = (float)x[3 + r];
  }
  T[3] = T[7] = T[11] = 0.f;
  T[15] = 1.f;
}

// quaternion (w,x,y,z) from the 3x3 block of a column-major float 4x4 (Eigen's conversion)
LS_HDN void quat_from_T(const float* T, double* q) {
  double m[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) m[r][c] = (double)T[c * 4 + r];
  double t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (m[2][1] - m[1][2]) * t;
    q[2] = (m[0][2] - m[2][0]) * t;
    q[3] = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    q[1 + i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[k][j] - m[j][k]) * t;
    q[1 + j] = (m[j][i] + m[i][j]) * t;
    q[1 + k] = (m[k][i] + m[i][k]) * t;
  }
}

LS_HDN double quat_angular_distance(const double* a, const double* b) {
  const double w = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
  const double x = -a[0] * b[1] + a[1] * b[0] - a[2] * b[3] + a[3] * b[2];
  const double y = -a[0] * b[2] + a[1] * b[3] + a[2] * b[0] - a[3] * b[1];
  const double z = -a[0] * b[3] - a[1] * b[2] + a[2] * b[1] + a[3] * b[0];
  return 2.0 * atan2(sqrt(x * x + y * y + z * z), fabs(w));
}

// Unit eigenvector of the smallest eigenvalue of a symmetric 3x3 (row-major) by cyclic Jacobi in double,
// fixed sweep order; ties between eigenvalues -> lowest index.  Surface normal of a neighbourhood covariance.
LS_HDN void smallest_eigvec3(const double* C, double* n) {
  double a[9], v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; ++i) a[i] = C[i];
  for (int sweep = 0; sweep < 40; ++sweep) {
    const double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
    const double dg = a[0] * a[0] + a[4] * a[4] + a[8] * a[8];
    if (!(off > 1e-40 * dg)) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = a[p * 3 + q];
        if (apq == 0.0) continue;
        const double theta = (a[q * 3 + q] - a[p * 3 + p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k * 3 + p], akq = a[k * 3 + q];
          a[k * 3 + p] = c * akp - s * akq;
          a[k * 3 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p * 3 + k], aqk = a[q * 3 + k];
          a[p * 3 + k] = c * apk - s * aqk;
          a[q * 3 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k * 3 + p], vkq = v[k * 3 + q];
          v[k * 3 + p] = c * vkp - s * vkq;
          v[k * 3 + q] = s * vkp + c * vkq;
        }
      }
  }
  int m = 0;
  if (a[4] < a[m * 4]) m = 1;
  if (a[8] < a[m * 4]) m = 2;
  const double x = v[m], y = v[3 + m], z = v[6 + m];
  const double inv = 1.0 / sqrt(x * x + y * y + z * z);
  n[0] = x * inv;
  n[1] = y * inv;
  n[2] = z * inv;
}

// RigidTransformation::checkParameters / correctParameters (reference common.hpp:136-149 path)
LS_HDN int check_rigid(const float* T) {
  const float a = T[0], b = T[4], c = T[8], d = T[1], e = T[5], f = T[9], g = T[2], h = T[6], i = T[10];
  const float det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  return fabsf(1.0f - det) <= 1e-3f ? 1 : 0;
}

LS_HDN void correct_rigid(const float* Tin, float* Tout) {
  float c0[3] = {Tin[0], Tin[1], Tin[2]}, c1[3] = {Tin[4], Tin[5], Tin[6]};
  const float n0 = sqrtf(c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2]);
  for (int k = 0; k < 3; ++k) c0[k] /= n0;
  const float d = c0[0] * c1[0] + c0[1] * c1[1] + c0[2] * c1[2];
  for (int k = 0; k < 3; ++k) c1[k] -= d * c0[k];
  const float n1 = sqrtf(c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2]);
  for (int k = 0; k < 3; ++k) c1[k] /= n1;
  const float c2[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2],
                       c0[0] * c1[1] - c0[1] * c1[0]};
  for (int k = 0; k < 16; ++k) Tout[k] = Tin[k];
  for (int k = 0; k < 3; ++k) { Tout[k] = c0[k]; Tout[4 + k] = c1[k]; Tout[8 + k] = c2[k]; }
}

}  // namespace ls
