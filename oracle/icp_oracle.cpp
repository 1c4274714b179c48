// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under laser_slam_b200/ may include, link or call this.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// PARITY UNPINNED: the reference's own tests hold no golden vector for this path
// (reference laser_slam/test/test_empty.cpp:3-5 is ASSERT_TRUE(true)) and the arithmetic lives in
// un-vendored third-party libraries (reference dependencies.rosinstall:23-28: ethz-asl/libnabo and
// ethz-asl/libpointmatcher, no version pin) that are absent from this container.  This file is a
// CPU restatement of their published algorithm as driven by the reference call sites
//   icp_.compute(last_scan.scan, sub_map, T0)         reference laser_slam/src/laser_track.cpp:496
//   icp_.compute(sub_map_b, sub_map_a, initial_guess)   reference laser_slam/src/incremental_estimator.cpp:108
// with the chain of reference laser_slam/configurations/icp_default.yaml:9-27
//   KDTreeMatcher{knn 1, epsilon 0}  TrimmedDistOutlierFilter{ratio}  PointToPlaneErrorMinimizer
//   CounterTransformationChecker + DifferentialTransformationChecker
// and is cross-checked only against independent implementations (brute force, scipy cKDTree,
// numpy lstsq) in tests/.  Where upstream arithmetic is not reproducible (Eigen GEMM summation
// order, traversal-order tie-breaks, libm sin/cos) the oracle FIXES a definition; each is marked
// [DEFINED] below and listed in oracle/README.md.  The CUDA path follows these definitions, which
// is what makes correspondence indices and poses bit-comparable.
//
// Build: g++ -O2 -march=native -ffp-contract=off -fopenmp -shared -fPIC  (see oracle/Makefile).
// -ffp-contract=off is REQUIRED: every float/double operation below is individually rounded.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "icp_oracle.h"

namespace {

// ------------------------------------------------------------------------------------------------
// [DEFINED] float32 rigid transform of a point (upstream: Eigen 4x4 * 4xN product, order
// unspecified).  x' = ((r00*x + r01*y) + r02*z) + tx, each op rounded to float32, w == 1.
// T is column-major (Eigen default; T[c*4+r]).
inline void xform_point(const float* T, float x, float y, float z, float* out) {
  for (int r = 0; r < 3; ++r) {
    float a = T[0 * 4 + r] * x;
    float b = T[1 * 4 + r] * y;
    float c = T[2 * 4 + r] * z;
    float s = a + b;
    s = s + c;
    s = s + T[3 * 4 + r];
    out[r] = s;
  }
}

// [DEFINED] float32 4x4 product C = A*B, C(i,j) = (((a_i0 b_0j + a_i1 b_1j) + a_i2 b_2j) + a_i3 b_3j).
inline void mat4_mul(const float* A, const float* B, float* C) {
  float tmp[16];
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) {
      float s = A[0 * 4 + i] * B[j * 4 + 0];
      float t = A[1 * 4 + i] * B[j * 4 + 1];
      s = s + t;
      t = A[2 * 4 + i] * B[j * 4 + 2];
      s = s + t;
      t = A[3 * 4 + i] * B[j * 4 + 3];
      s = s + t;
      tmp[j * 4 + i] = s;
    }
  std::memcpy(C, tmp, sizeof(tmp));
}

// libnabo leaf arithmetic: dist = sum_i (q_i - p_i)^2 accumulated
// x -> y -> z in float32, no FMA.
inline float dist2(const float* q, const float* p) {
  float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
  float a = dx * dx;
  float b = dy * dy;
  float c = dz * dz;
  float s = a + b;
  return s + c;
}

// ------------------------------------------------------------------------------------------------
// kd-tree after libnabo KDTREE_LINEAR_HEAP: leaf when <= bucket points,
// split dimension = largest extent, split position = median by nth_element, cut value = coordinate
// of the first point of the right part.  [DEFINED] ties: lowest reference index wins (libnabo's
// winner depends on traversal order); implemented as lexicographic (d2, index) minimisation with a
// far-child visit whenever plane_d2 <= best (so equal-distance candidates are never pruned).
struct KdNode {
  int dim;        // -1 => leaf
  float cut;
  int left, right;  // children (internal) or [begin,end) into perm (leaf)
};

struct KdTree {
  const float* pts = nullptr;  // xyz, stride 3, centred reference
  std::vector<int> perm;
  std::vector<KdNode> nodes;
  static constexpr int kBucket = 8;

  int build(int begin, int end) {
    const int id = (int)nodes.size();
    nodes.push_back(KdNode{-1, 0.f, begin, end});
    const int count = end - begin;
    if (count <= kBucket) return id;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = begin; i < end; ++i)
      for (int a = 0; a < 3; ++a) {
        const float v = pts[3 * perm[i] + a];
        lo[a] = std::min(lo[a], v);
        hi[a] = std::max(hi[a], v);
      }
    int dim = 0;
    float ext = hi[0] - lo[0];
    for (int a = 1; a < 3; ++a)
      if (hi[a] - lo[a] > ext) { ext = hi[a] - lo[a]; dim = a; }
    if (!(ext > 0.f)) return id;  // all points identical: keep as (large) leaf
    const int right_count = count / 2, left_count = count - right_count;
    const float* P = pts;
    std::nth_element(perm.begin() + begin, perm.begin() + begin + left_count, perm.begin() + end,
                     [P, dim](int a, int b) {
                       const float va = P[3 * a + dim], vb = P[3 * b + dim];
                       return va < vb || (va == vb && a < b);
                     });
    const float cut = pts[3 * perm[begin + left_count] + dim];
    const int l = build(begin, begin + left_count);
    const int r = build(begin + left_count, end);
    nodes[id].dim = dim;
    nodes[id].cut = cut;
    nodes[id].left = l;
    nodes[id].right = r;
    return id;
  }

  void init(const float* centred_xyz, int m) {
    pts = centred_xyz;
    perm.resize(m);
    for (int i = 0; i < m; ++i) perm[i] = i;
    nodes.clear();
    nodes.reserve(m / 2 + 16);
    if (m > 0) build(0, m);
  }

  void search(int node, const float* q, float& best, int& best_id) const {
    const KdNode& nd = nodes[node];
    if (nd.dim < 0) {
      for (int i = nd.left; i < nd.right; ++i) {
        const int id = perm[i];
        const float d = dist2(q, pts + 3 * id);
        if (d < best || (d == best && id < best_id)) { best = d; best_id = id; }
      }
      return;
    }
    const float diff = q[nd.dim] - nd.cut;
    const int near = diff < 0.f ? nd.left : nd.right;
    const int far = diff < 0.f ? nd.right : nd.left;
    search(near, q, best, best_id);
    const float pd = diff * diff;
    if (pd <= best) search(far, q, best, best_id);
  }

  // K nearest, ascending by (d2, index) -- the lexicographic order that makes ties reproducible
  void knn_rec(int node, const float* q, int k, float* bd, int* bi) const {
    const KdNode& nd = nodes[node];
    if (nd.dim < 0) {
      for (int i = nd.left; i < nd.right; ++i) {
        const int id = perm[i];
        const float d = dist2(q, pts + 3 * id);
        if (d < bd[k - 1] || (d == bd[k - 1] && id < bi[k - 1])) {
          int j = k - 1;
          while (j > 0 && (d < bd[j - 1] || (d == bd[j - 1] && id < bi[j - 1]))) {
            bd[j] = bd[j - 1];
            bi[j] = bi[j - 1];
            --j;
          }
          bd[j] = d;
          bi[j] = id;
        }
      }
      return;
    }
    const float diff = q[nd.dim] - nd.cut;
    const int near = diff < 0.f ? nd.left : nd.right;
    const int far = diff < 0.f ? nd.right : nd.left;
    knn_rec(near, q, k, bd, bi);
    const float pd = diff * diff;
    if (pd <= bd[k - 1]) knn_rec(far, q, k, bd, bi);
  }
  void knn(const float* q, int k, float* bd, int* bi) const {
    for (int j = 0; j < k; ++j) { bd[j] = INFINITY; bi[j] = std::numeric_limits<int>::max(); }
    if (!nodes.empty()) knn_rec(0, q, k, bd, bi);
  }

  void nn(const float* q, int* id, float* d2) const {
    float best = INFINITY;
    int bid = -1;  // libnabo: unfound = -1 / +inf
    if (!nodes.empty()) search(0, q, best, bid);
    *id = bid;
    *d2 = best;
  }
};

// ------------------------------------------------------------------------------------------------
// [DEFINED] deterministic double sin/cos (libm's last-ulp behaviour differs between hosts and from
// CUDA's): Cody-Waite reduction by pi/2 + Taylor polynomials in r^2, Horner, every op rounded.
inline void det_sincos(double x, double* s_out, double* c_out) {
  const double two_over_pi = 0.63661977236758134308;
  const double pio2_hi = 1.57079632673412561417e+00;  // first 33 bits of pi/2
  const double pio2_lo = 6.07710050650619224932e-11;  // pi/2 - pio2_hi
  double kf = std::floor(x * two_over_pi + 0.5);
  double r = x - kf * pio2_hi;
  r = r - kf * pio2_lo;
  const double z = r * r;
  // sin(r) = r * (1 + z*(S1 + z*(S2 + ...)))   up to r^17
  double ps = -1.0 / 355687428096000.0;          // -1/17!
  ps = ps * z + 1.0 / 1307674368000.0;           // +1/15!
  ps = ps * z - 1.0 / 6227020800.0;              // -1/13!
  ps = ps * z + 1.0 / 39916800.0;                // +1/11!
  ps = ps * z - 1.0 / 362880.0;                  // -1/9!
  ps = ps * z + 1.0 / 5040.0;                    // +1/7!
  ps = ps * z - 1.0 / 120.0;                     // -1/5!
  ps = ps * z + 1.0 / 6.0;                       // +1/3!  (sign folded below)
  double sr = r - r * (z * ps);
  // cos(r) = 1 - z/2 + z^2/24 - ...            up to r^18
  double pc = -1.0 / 6402373705728000.0;         // -1/18!
  pc = pc * z + 1.0 / 20922789888000.0;          // +1/16!
  pc = pc * z - 1.0 / 87178291200.0;             // -1/14!
  pc = pc * z + 1.0 / 479001600.0;               // +1/12!
  pc = pc * z - 1.0 / 3628800.0;                 // -1/10!
  pc = pc * z + 1.0 / 40320.0;                   // +1/8!
  pc = pc * z - 1.0 / 720.0;                     // -1/6!
  pc = pc * z + 1.0 / 24.0;                      // +1/4!
  pc = pc * z - 0.5;                             // -1/2!
  double cr = 1.0 + z * pc;
  long long k = (long long)kf;
  switch (k & 3) {
    case 0: *s_out = sr; *c_out = cr; break;
    case 1: *s_out = cr; *c_out = -sr; break;
    case 2: *s_out = -sr; *c_out = -cr; break;
    default: *s_out = -cr; *c_out = sr; break;
  }
}

// ------------------------------------------------------------------------------------------------
// 6x6 solve (upstream PointToPlaneErrorMinimizer: A.llt().solve(b), rank-revealing fallback).
// [DEFINED] double Cholesky with a fixed operation order; pivot loss > 1e10 => minimum-norm
// solution through a cyclic Jacobi eigen-decomposition.
bool chol6(const double A[36], const double b[6], double x[6]) {
  // [DEFINED] column Cholesky; each pivot is inverted once (inv = 1/sqrt(s)) and later divisions
  // are multiplications by inv, so the device needs 6 sqrt + 6 div only.
  double L[36], inv[6];
  for (int i = 0; i < 36; ++i) L[i] = 0.0;
  for (int j = 0; j < 6; ++j) {
    double s = A[j * 6 + j];
    for (int k = 0; k < j; ++k) {
      double t = L[j * 6 + k] * L[j * 6 + k];
      s = s - t;
    }
    if (!(s > 1e-10 * A[j * 6 + j]) || !std::isfinite(s)) return false;
    const double d = std::sqrt(s);
    L[j * 6 + j] = d;
    inv[j] = 1.0 / d;
    for (int i = j + 1; i < 6; ++i) {
      double v = A[i * 6 + j];
      for (int k = 0; k < j; ++k) {
        double t = L[i * 6 + k] * L[j * 6 + k];
        v = v - t;
      }
      L[i * 6 + j] = v * inv[j];
    }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double v = b[i];
    for (int k = 0; k < i; ++k) {
      double t = L[i * 6 + k] * y[k];
      v = v - t;
    }
    y[i] = v * inv[i];
  }
  for (int i = 5; i >= 0; --i) {
    double v = y[i];
    for (int k = i + 1; k < 6; ++k) {
      double t = L[k * 6 + i] * x[k];
      v = v - t;
    }
    x[i] = v * inv[i];
  }
  return true;
}

void jacobi_pinv_solve6(const double Ain[36], const double b[6], double x[6]) {
  double a[36], v[36];
  std::memcpy(a, Ain, sizeof(a));
  for (int i = 0; i < 36; ++i) v[i] = 0.0;
  for (int i = 0; i < 6; ++i) v[i * 6 + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int p = 0; p < 6; ++p) {
      double t = a[p * 6 + p] * a[p * 6 + p];
      dg = dg + t;
      for (int q = p + 1; q < 6; ++q) {
        double u = a[p * 6 + q] * a[p * 6 + q];
        off = off + u;
      }
    }
    if (!(off > 1e-40 * dg)) break;
    for (int p = 0; p < 5; ++p)
      for (int q = p + 1; q < 6; ++q) {
        const double apq = a[p * 6 + q];
        if (apq == 0.0) continue;
        const double theta = (a[q * 6 + q] - a[p * 6 + p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 6; ++k) {  // columns p,q of a
          const double akp = a[k * 6 + p], akq = a[k * 6 + q];
          a[k * 6 + p] = c * akp - s * akq;
          a[k * 6 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 6; ++k) {  // rows p,q of a
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
  for (int i = 0; i < 6; ++i) lmax = std::max(lmax, std::fabs(a[i * 6 + i]));
  for (int i = 0; i < 6; ++i) x[i] = 0.0;
  for (int k = 0; k < 6; ++k) {
    const double lam = a[k * 6 + k];
    if (!(lam > 1e-10 * lmax)) continue;
    double proj = 0.0;
    for (int i = 0; i < 6; ++i) {
      double t = v[i * 6 + k] * b[i];
      proj = proj + t;
    }
    const double coef = proj / lam;
    for (int i = 0; i < 6; ++i) {
      double t = v[i * 6 + k] * coef;
      x[i] = x[i] + t;
    }
  }
}

// x (rotation vector, translation) -> float 4x4, after Eigen AngleAxis::toRotationMatrix
// (upstream: AngleAxis(||x0-2||, x0-2/||.||); NaN => rotation block := identity).
void step_matrix(const double x[6], float T[16]) {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double n2 = x[0] * x[0];
  double t1 = x[1] * x[1];
  n2 = n2 + t1;
  t1 = x[2] * x[2];
  n2 = n2 + t1;
  const double th = std::sqrt(n2);
  if (th > 0.0 && std::isfinite(th)) {
    const double ux = x[0] / th, uy = x[1] / th, uz = x[2] / th;
    double s, c;
    det_sincos(th, &s, &c);
    const double sx = s * ux, sy = s * uy, sz = s * uz;
    const double omc = 1.0 - c;
    const double cx = omc * ux, cy = omc * uy, cz = omc * uz;
    double tmp;
    tmp = cx * uy; R[0 * 3 + 1] = tmp - sz; R[1 * 3 + 0] = tmp + sz;
    tmp = cx * uz; R[0 * 3 + 2] = tmp + sy; R[2 * 3 + 0] = tmp - sy;
    tmp = cy * uz; R[1 * 3 + 2] = tmp - sx; R[2 * 3 + 1] = tmp + sx;
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

// Quaternion (w,x,y,z) from the 3x3 block of a float 4x4, after Eigen's Quaternion(Matrix3).
void quat_from_T(const float T[16], double q[4]) {
  double m[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) m[r][c] = (double)T[c * 4 + r];
  double t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
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
    t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    q[1 + i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[k][j] - m[j][k]) * t;
    q[1 + j] = (m[j][i] + m[i][j]) * t;
    q[1 + k] = (m[k][i] + m[i][k]) * t;
  }
}

double quat_angular_distance(const double a[4], const double b[4]) {
  // d = a * conj(b); angle = 2*atan2(|vec d|, |d.w|)   (Eigen >= 3.3 formulation)
  const double w = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
  const double x = -a[0] * b[1] + a[1] * b[0] - a[2] * b[3] + a[3] * b[2];
  const double y = -a[0] * b[2] + a[1] * b[3] + a[2] * b[0] - a[3] * b[1];
  const double z = -a[0] * b[3] - a[1] * b[2] + a[2] * b[1] + a[3] * b[0];
  return 2.0 * std::atan2(std::sqrt(x * x + y * y + z * z), std::fabs(w));
}

inline long long quantise(float v, float scale) { return std::llrint((double)(v * scale)); }

}  // namespace

extern "C" {

void lso_default_params(lso_icp_params* p) {
  // reference laser_slam/configurations/icp_default.yaml:14-27
  p->max_iterations = 40;
  p->trim_ratio = 0.75f;
  p->use_differential = 1;
  p->min_diff_rot = 0.001f;
  p->min_diff_trans = 0.01f;
  p->smooth_length = 4;
  p->num_threads = 1;
}

// [DEFINED] reference mean: exact order-independent fixed-point sum (2^-24 m resolution), then
// mean = float(double(sum) / (M * 2^24)).  Upstream: float32 rowwise().sum()/N (SIMD order).
void lso_mean(const float* ref4, int m, float mu[3]) {
  long long s[3] = {0, 0, 0};
  for (int i = 0; i < m; ++i)
    for (int a = 0; a < 3; ++a) s[a] += std::llrint((double)ref4[4 * i + a] * 16777216.0);
  for (int a = 0; a < 3; ++a) mu[a] = (float)((double)s[a] / ((double)m * 16777216.0));
}

// Brute-force exact NN with the defined tie-break; ref_xyz3 stride 3.
void lso_nn_brute(const float* q3, int n, const float* ref_xyz3, int m, int32_t* ids, float* d2) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    float best = INFINITY;
    int bid = -1;
    for (int j = 0; j < m; ++j) {
      const float d = dist2(q3 + 3 * i, ref_xyz3 + 3 * j);
      if (d < best) { best = d; bid = j; }  // ascending j + strict '<' == lowest index on ties
    }
    ids[i] = bid;
    d2[i] = best;
  }
}

void lso_nn_kdtree(const float* q3, int n, const float* ref_xyz3, int m, int32_t* ids, float* d2,
                   int num_threads) {
  KdTree tree;
  tree.init(ref_xyz3, m);
#pragma omp parallel for schedule(static) num_threads(num_threads > 0 ? num_threads : 1)
  for (int i = 0; i < n; ++i) tree.nn(q3 + 3 * i, &ids[i], &d2[i]);
}

// TrimmedDistOutlierFilter limit: nth_element at (size_t)(n * ratio) over
// the finite distances, evaluated in float32 like upstream's `values.size() * quantile`.
float lso_trim_limit(const float* d2, int n, float ratio, int* n_finite_out) {
  std::vector<float> v;
  v.reserve(n);
  for (int i = 0; i < n; ++i)
    if (std::isfinite(d2[i])) v.push_back(d2[i]);
  if (n_finite_out) *n_finite_out = (int)v.size();
  if (v.empty()) return -1.f;
  size_t k = (size_t)((float)v.size() * ratio);
  if (k >= v.size()) k = v.size() - 1;
  std::nth_element(v.begin(), v.begin() + k, v.end());
  return v[k];
}

void lso_transform_points(const float T[16], const float* in4, int n, float* out4) {
  for (int i = 0; i < n; ++i) {
    float o[3];
    xform_point(T, in4[4 * i], in4[4 * i + 1], in4[4 * i + 2], o);
    out4[4 * i] = o[0];
    out4[4 * i + 1] = o[1];
    out4[4 * i + 2] = o[2];
    out4[4 * i + 3] = in4[4 * i + 3];
  }
}

// RigidTransformation::compute on a cloud with a `normals` descriptor:
// features' = T*features, normals' = R*normals ([DEFINED] same rounded order, no translation).
void lso_transform_cloud(const float T[16], const float* in4, const float* nin, int nstride, int n,
                         float* out4, float* nout3) {
  lso_transform_points(T, in4, n, out4);
  if (!nin || !nout3) return;
  for (int i = 0; i < n; ++i) {
    const float x = nin[(size_t)i * nstride], y = nin[(size_t)i * nstride + 1], z = nin[(size_t)i * nstride + 2];
    for (int r = 0; r < 3; ++r) {
      float a = T[0 * 4 + r] * x;
      float b = T[1 * 4 + r] * y;
      float c = T[2 * 4 + r] * z;
      float s = a + b;
      nout3[3 * i + r] = s + c;
    }
  }
}

// RigidTransformation::checkParameters / correctParameters, as used by
// correctTransformationMatrix (reference laser_slam/include/laser_slam/common.hpp:136-149).
int lso_check_rigid(const float T[16]) {
  const float a = T[0], b = T[4], c = T[8], d = T[1], e = T[5], f = T[9], g = T[2], h = T[6], i = T[10];
  const float det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  return std::fabs(1.0f - det) <= 1e-3f ? 1 : 0;
}

void lso_correct_rigid(const float Tin[16], float Tout[16]) {
  // Gram-Schmidt on the rotation columns: col0 normalised, col1 made orthogonal to col0 and
  // normalised, col2 = col0 x col1; translation kept.
  float c0[3] = {Tin[0], Tin[1], Tin[2]}, c1[3] = {Tin[4], Tin[5], Tin[6]};
  float n0 = std::sqrt(c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2]);
  for (float& v : c0) v /= n0;
  float d = c0[0] * c1[0] + c0[1] * c1[1] + c0[2] * c1[2];
  for (int k = 0; k < 3; ++k) c1[k] -= d * c0[k];
  float n1 = std::sqrt(c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2]);
  for (float& v : c1) v /= n1;
  const float c2[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2],
                       c0[0] * c1[1] - c0[1] * c1[0]};
  std::memcpy(Tout, Tin, 16 * sizeof(float));
  for (int k = 0; k < 3; ++k) { Tout[k] = c0[k]; Tout[4 + k] = c1[k]; Tout[8 + k] = c2[k]; }
}

// PointToPlaneErrorMinimizer normal equations for one iteration (SURVEY.md §8c spec 4).
//   step4: transformed reading (float4), ref_c: centred reference xyz (stride 3), nrm: normals.
// Outputs A (row-major 6x6), b, kept count.  [DEFINED] accumulation: each float32 product is
// quantised to 2^-22 and summed as int64 (order independent); A_f64/b_f64 (optional) hold the plain
// double accumulation of the same float32 per-point terms for the tolerance report.
void lso_normal_equations(const float* step4, int n, const float* ref_c3, const float* nrm, int nstride,
                          const int32_t* ids, const float* d2, float limit, double A[36], double b[6],
                          int* kept_out, double* A_f64, double* b_f64) {
  long long Ai[21], bi[6];
  double Ad[21], bd[6];
  for (auto& v : Ai) v = 0;
  for (auto& v : bi) v = 0;
  for (auto& v : Ad) v = 0.0;
  for (auto& v : bd) v = 0.0;
  int kept = 0;
  const float scale = 4194304.0f;  // 2^22
  for (int i = 0; i < n; ++i) {
    if (!(d2[i] <= limit)) continue;
    const int id = ids[i];
    if (id < 0) continue;
    ++kept;
    const float sx = step4[4 * i], sy = step4[4 * i + 1], sz = step4[4 * i + 2];
    const float* q = ref_c3 + 3 * (size_t)id;
    const float nx = nrm[(size_t)id * nstride], ny = nrm[(size_t)id * nstride + 1], nz = nrm[(size_t)id * nstride + 2];
    float f[6];
    {
      float a = sy * nz, c = sz * ny;
      f[0] = a - c;
      a = sz * nx; c = sx * nz;
      f[1] = a - c;
      a = sx * ny; c = sy * nx;
      f[2] = a - c;
    }
    f[3] = nx; f[4] = ny; f[5] = nz;
    const float dx = sx - q[0], dy = sy - q[1], dz = sz - q[2];
    float e = dx * nx;
    float t = dy * ny;
    e = e + t;
    t = dz * nz;
    e = e + t;
    int k = 0;
    for (int r = 0; r < 6; ++r)
      for (int c = r; c < 6; ++c, ++k) {
        const float p = f[r] * f[c];
        Ai[k] += quantise(p, scale);
        Ad[k] += (double)p;
      }
    for (int r = 0; r < 6; ++r) {
      const float p = f[r] * e;
      bi[r] += quantise(p, scale);
      bd[r] += (double)p;
    }
  }
  int k = 0;
  for (int r = 0; r < 6; ++r)
    for (int c = r; c < 6; ++c, ++k) {
      const double v = (double)Ai[k] / 4194304.0;
      A[r * 6 + c] = v;
      A[c * 6 + r] = v;
      if (A_f64) { A_f64[r * 6 + c] = Ad[k]; A_f64[c * 6 + r] = Ad[k]; }
    }
  for (int r = 0; r < 6; ++r) {
    b[r] = -((double)bi[r] / 4194304.0);
    if (b_f64) b_f64[r] = -bd[r];
  }
  if (kept_out) *kept_out = kept;
}

// Solve + build the float step matrix.  Returns 0 ok, 1 if the result is not finite.
int lso_solve_step(const double A[36], const double b[6], float T_step[16], double x_out[6]) {
  double x[6];
  if (!chol6(A, b, x)) jacobi_pinv_solve6(A, b, x);
  for (int i = 0; i < 6; ++i)
    if (!std::isfinite(x[i])) return 1;
  step_matrix(x, T_step);
  if (x_out) std::memcpy(x_out, x, sizeof(x));
  return 0;
}

void lso_mat4_mul(const float* A, const float* B, float* C) { mat4_mul(A, B, C); }

// [DEFINED] eigenvector of the smallest eigenvalue of a symmetric 3x3 (row-major): cyclic Jacobi in double with
// a fixed sweep order; equal eigenvalues -> lowest index.
static void smallest_eigvec3(const double C[9], double n[3]) {
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
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
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
  const double inv = 1.0 / std::sqrt(x * x + y * y + z * z);
  n[0] = x * inv; n[1] = y * inv; n[2] = z * inv;
}

// Exact k nearest neighbours (self included, ties by index) of every point of a cloud within the cloud itself,
// on the coordinates centred like ICP::compute centres a reference.  ids: n*k, d2: n*k (ascending).
void lso_knn_self(const float* feat4, int n, int k, int32_t* ids, float* d2, int num_threads) {
  float mu[3];
  lso_mean(feat4, n, mu);
  std::vector<float> c(3 * (size_t)n);
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) c[3 * (size_t)i + a] = feat4[4 * (size_t)i + a] - mu[a];
  KdTree tree;
  tree.init(c.data(), n);
#pragma omp parallel for schedule(static) num_threads(num_threads > 0 ? num_threads : 1)
  for (int i = 0; i < n; ++i) tree.knn(&c[3 * (size_t)i], k, d2 + (size_t)i * k, ids + (size_t)i * k);
}

// Surface normals (upstream: SurfaceNormalDataPointsFilter + orientation towards the sensor,
// icp_default.yaml:5-7).  [DEFINED]: neighbourhood = exact k-NN (self included) on the centred float32 cloud;
// mean and covariance accumulated in double in (d2, index) order; eigenvector of the smallest eigenvalue (Jacobi);
// flipped so that it points towards the scan-frame origin; fewer than 3 neighbours -> zero normal.
void lso_knn_normals(const float* feat4, int n, int k, float* out3, int num_threads) {
  float mu[3];
  lso_mean(feat4, n, mu);
  std::vector<float> c(3 * (size_t)n);
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a) c[3 * (size_t)i + a] = feat4[4 * (size_t)i + a] - mu[a];
  KdTree tree;
  tree.init(c.data(), n);
#pragma omp parallel for schedule(static) num_threads(num_threads > 0 ? num_threads : 1)
  for (int i = 0; i < n; ++i) {
    float bd[64];
    int bi[64];
    tree.knn(&c[3 * (size_t)i], k, bd, bi);
    int cnt = 0;
    double mx = 0.0, my = 0.0, mz = 0.0;
    for (int j = 0; j < k; ++j) {
      if (bi[j] == std::numeric_limits<int>::max()) break;
      mx = mx + (double)c[3 * (size_t)bi[j]];
      my = my + (double)c[3 * (size_t)bi[j] + 1];
      mz = mz + (double)c[3 * (size_t)bi[j] + 2];
      ++cnt;
    }
    float nn[3] = {0.f, 0.f, 0.f};
    if (cnt >= 3) {
      const double inv = 1.0 / (double)cnt;
      mx = mx * inv; my = my * inv; mz = mz * inv;
      double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int j = 0; j < cnt; ++j) {
        const double dx = (double)c[3 * (size_t)bi[j]] - mx, dy = (double)c[3 * (size_t)bi[j] + 1] - my,
                     dz = (double)c[3 * (size_t)bi[j] + 2] - mz;
        C[0] = C[0] + dx * dx; C[1] = C[1] + dx * dy; C[2] = C[2] + dx * dz;
        C[4] = C[4] + dy * dy; C[5] = C[5] + dy * dz; C[8] = C[8] + dz * dz;
      }
      C[3] = C[1]; C[6] = C[2]; C[7] = C[5];
      double nv[3];
      smallest_eigvec3(C, nv);
      const double ox = (double)c[3 * (size_t)i] + (double)mu[0], oy = (double)c[3 * (size_t)i + 1] + (double)mu[1],
                   oz = (double)c[3 * (size_t)i + 2] + (double)mu[2];
      const double dot = nv[0] * ox + (nv[1] * oy + nv[2] * oz);
      const double sgn = dot > 0.0 ? -1.0 : 1.0;
      nn[0] = (float)(sgn * nv[0]); nn[1] = (float)(sgn * nv[1]); nn[2] = (float)(sgn * nv[2]);
    }
    out3[3 * (size_t)i] = nn[0]; out3[3 * (size_t)i + 1] = nn[1]; out3[3 * (size_t)i + 2] = nn[2];
  }
}
void lso_sincos(double x, double* s, double* c) { det_sincos(x, s, c); }

// PointMatcher::ICP::compute with identity reading/reference filters
// (normals are an input; the random filters of icp_default.yaml:1-7 are not reproducible, A.8).
// Returns 0 ok, 1 = ConvergenceError (no point to minimise / NaN), <0 argument error.
int lso_icp(const float* reading4, int n, const float* ref4, const float* ref_normals, int nstride, int m,
            const float T0[16], const lso_icp_params* prm, float T_out[16], lso_icp_stats* stats,
            int32_t* ids_hist /* opt, max_iterations*n */, float* d2_last /* opt, n */,
            float* T_iter_hist /* opt, max_iterations*16 */) {
  if (!reading4 || !ref4 || !ref_normals || !T0 || !prm || !T_out || n < 0 || m < 0 || nstride < 3) return -1;
  lso_icp_stats st;
  std::memset(&st, 0, sizeof(st));
  std::memcpy(T_out, T0, 16 * sizeof(float));
  if (n == 0 || m == 0) { if (stats) *stats = st; return 1; }
  const int nthreads = prm->num_threads > 0 ? prm->num_threads : 1;

  // (2) centre the reference
  float mu[3];
  lso_mean(ref4, m, mu);
  std::vector<float> refc(3 * (size_t)m);
  for (int i = 0; i < m; ++i)
    for (int a = 0; a < 3; ++a) refc[3 * (size_t)i + a] = ref4[4 * (size_t)i + a] - mu[a];
  // (3) matcher->init
  KdTree tree;
  tree.init(refc.data(), m);
  // (5) T_refMean_dataIn = T_refIn_refMean^-1 * T0  == [R0 | t0 - mu]
  float T_pre[16];
  std::memcpy(T_pre, T0, sizeof(T_pre));
  for (int a = 0; a < 3; ++a) T_pre[12 + a] = T0[12 + a] - mu[a];
  std::vector<float> rd(4 * (size_t)n), step(4 * (size_t)n);
  lso_transform_points(T_pre, reading4, n, rd.data());

  float T_iter[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  std::vector<int32_t> ids(n);
  std::vector<float> d2(n);
  // checker state (Appendix A.6)
  std::vector<double> quat_hist, trans_hist;
  {
    double q[4];
    quat_from_T(T_iter, q);
    quat_hist.insert(quat_hist.end(), q, q + 4);
    trans_hist.insert(trans_hist.end(), {0.0, 0.0, 0.0});
  }
  int iter = 0, rc = 0;
  bool iterate = true;
  while (iterate) {
    lso_transform_points(T_iter, rd.data(), n, step.data());
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int i = 0; i < n; ++i) tree.nn(&step[4 * (size_t)i], &ids[i], &d2[i]);
    if (ids_hist) std::memcpy(ids_hist + (size_t)iter * n, ids.data(), sizeof(int32_t) * n);
    int n_finite = 0;
    const float limit = lso_trim_limit(d2.data(), n, prm->trim_ratio, &n_finite);
    if (n_finite == 0) { rc = 1; break; }
    double A[36], b[6];
    int kept = 0;
    lso_normal_equations(step.data(), n, refc.data(), ref_normals, nstride, ids.data(), d2.data(), limit, A, b,
                         &kept, nullptr, nullptr);
    st.last_kept = kept;
    st.last_limit = limit;
    if (kept == 0) { rc = 1; break; }
    float T_step[16];
    if (lso_solve_step(A, b, T_step, nullptr)) { rc = 1; break; }
    mat4_mul(T_step, T_iter, T_iter);
    if (T_iter_hist) std::memcpy(T_iter_hist + (size_t)iter * 16, T_iter, sizeof(T_iter));
    ++iter;
    // checkers
    if (iter >= prm->max_iterations) { iterate = false; st.max_iter_reached = 1; }
    if (prm->use_differential) {
      double q[4];
      quat_from_T(T_iter, q);
      quat_hist.insert(quat_hist.end(), q, q + 4);
      trans_hist.insert(trans_hist.end(), {(double)T_iter[12], (double)T_iter[13], (double)T_iter[14]});
      const size_t cnt = quat_hist.size() / 4;
      if ((int)cnt > prm->smooth_length) {
        double mr = 0.0, mt = 0.0;
        for (size_t i = cnt - 1; i >= cnt - prm->smooth_length; --i) {
          mr += std::fabs(quat_angular_distance(&quat_hist[4 * i], &quat_hist[4 * (i - 1)]));
          const double dx = trans_hist[3 * i] - trans_hist[3 * (i - 1)],
                       dy = trans_hist[3 * i + 1] - trans_hist[3 * (i - 1) + 1],
                       dz = trans_hist[3 * i + 2] - trans_hist[3 * (i - 1) + 2];
          mt += std::sqrt(dx * dx + dy * dy + dz * dz);
        }
        mr /= prm->smooth_length;
        mt /= prm->smooth_length;
        if (std::isnan(mr) || std::isnan(mt)) { rc = 1; break; }
        if (mr < (double)prm->min_diff_rot && mt < (double)prm->min_diff_trans) { iterate = false; st.converged = 1; }
      }
    }
  }
  st.iterations = iter;
  st.used_ratio = n > 0 ? (float)st.last_kept / (float)n : 0.f;
  if (stats) *stats = st;
  if (d2_last) std::memcpy(d2_last, d2.data(), sizeof(float) * n);
  if (rc != 0) return rc;
  // (7) T_refIn_refMean * T_iter * T_refMean_dataIn, evaluated left to right
  float T_mean[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, mu[0], mu[1], mu[2], 1};
  float tmp[16];
  mat4_mul(T_mean, T_iter, tmp);
  mat4_mul(tmp, T_pre, T_out);
  for (int i = 0; i < 16; ++i)
    if (!std::isfinite(T_out[i])) { std::memcpy(T_out, T0, 16 * sizeof(float)); return 1; }
  return 0;
}

}  // extern "C"
