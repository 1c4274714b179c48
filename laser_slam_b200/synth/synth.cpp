// Synthetic Velodyne-shaped workload generator.
//
// The reference ships no point clouds (reference .gitignore:23-26 excludes *.pcd/*.vtk/*.bag), so
// every parity / benchmark input is produced here: a closed street-canyon scene ray-cast from an
// HDL-64- or VLS-128-shaped sensor that moves along a seeded trajectory.  Output layout is the
// libpointmatcher DataPoints one the reference hands to LaserTrack (reference
// laser_slam/include/laser_slam/common.hpp:14-15,113-120): `features` column-major 4xN float
// (x,y,z,1) and a `normals` descriptor 3xN float.  Points are expressed in the SENSOR frame.
//
// CPU only, single implementation (Python reaches it through ctypes) so that the oracle, the CUDA
// path and the bench all see byte-identical inputs.  Counter-based SplitMix64 -> Box-Muller RNG:
//   seed(sequence s, scan k, stream) = 0x5EED000000000000 + (s<<32) + (k<<8) + stream.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

constexpr double kPi = 3.14159265358979323846;

struct Rng {
  uint64_t state;
  explicit Rng(uint64_t seed) : state(seed) {}
  uint64_t next() {
    uint64_t z = (state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
  double normal() {
    double u1 = uniform(), u2 = uniform();
    if (u1 < 1e-300) u1 = 1e-300;
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * kPi * u2);
  }
};

inline uint64_t make_seed(uint32_t seq, uint32_t scan, uint32_t stream) {
  return 0x5EED000000000000ull + ((uint64_t)seq << 32) + ((uint64_t)scan << 8) + stream;
}

struct V3 { double x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

struct M3 { double m[9]; };  // row-major
inline V3 mul(const M3& R, V3 v) {
  return {R.m[0] * v.x + R.m[1] * v.y + R.m[2] * v.z, R.m[3] * v.x + R.m[4] * v.y + R.m[5] * v.z,
          R.m[6] * v.x + R.m[7] * v.y + R.m[8] * v.z};
}
inline V3 mulT(const M3& R, V3 v) {
  return {R.m[0] * v.x + R.m[3] * v.y + R.m[6] * v.z, R.m[1] * v.x + R.m[4] * v.y + R.m[7] * v.z,
          R.m[2] * v.x + R.m[5] * v.y + R.m[8] * v.z};
}
inline M3 mul(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
inline M3 rot_zyx(double yaw, double pitch, double roll) {
  const double cy = std::cos(yaw), sy = std::sin(yaw), cp = std::cos(pitch), sp = std::sin(pitch),
               cr = std::cos(roll), sr = std::sin(roll);
  M3 Rz{{cy, -sy, 0, sy, cy, 0, 0, 0, 1}}, Ry{{cp, 0, sp, 0, 1, 0, -sp, 0, cp}},
      Rx{{1, 0, 0, 0, cr, -sr, 0, sr, cr}};
  return mul(mul(Rz, Ry), Rx);
}

struct Pose { M3 R; V3 t; };  // T_world_sensor

// ---------------------------------------------------------------- scene -------------------------
struct Box { V3 lo, hi; };
struct Cyl { double cx, cy, r, z0, z1; };

struct Scene {
  double zg = -1.73, zc = 20.0, half_w = 12.0, alcove_depth = 8.0, alcove_period = 20.0,
         alcove_half = 3.0, half_len = 150.0;
  std::vector<Box> free_boxes;  // union = free space; index 0 is the corridor
  std::vector<Box> boxes;
  std::vector<Cyl> cyls;
};

Scene build_scene(uint32_t scene_seed, double length_m) {
  Scene sc;
  sc.half_len = 0.5 * length_m;
  sc.free_boxes.push_back({{-sc.half_w, -sc.half_len, sc.zg}, {sc.half_w, sc.half_len, sc.zc}});
  const int n_alc = (int)std::floor(sc.half_len / sc.alcove_period);
  for (int k = -n_alc; k <= n_alc; ++k) {
    const double yc = k * sc.alcove_period + 7.0;  // offset so alcoves are not mirror-symmetric in y
    if (yc - sc.alcove_half <= -sc.half_len + 1.0 || yc + sc.alcove_half >= sc.half_len - 1.0) continue;
    // alcoves alternate sides so the canyon is not mirror-symmetric in x either
    if ((k & 1) == 0)
      sc.free_boxes.push_back({{sc.half_w, yc - sc.alcove_half, sc.zg},
                               {sc.half_w + sc.alcove_depth, yc + sc.alcove_half, sc.zc}});
    else
      sc.free_boxes.push_back({{-sc.half_w - sc.alcove_depth, yc - sc.alcove_half, sc.zg},
                               {-sc.half_w, yc + sc.alcove_half, sc.zc}});
  }
  // obstacles: 8 boxes + 4 cylinders per 100 m of street (24 + 12 for the canonical 300 m scene),
  // kept in the outer lanes (|x| in [6.5, 10.5]) so the sensor path (|x| < 3) never enters one.
  const int n_seg = std::max(1, (int)std::lround(length_m / 100.0));
  for (int s = 0; s < n_seg; ++s) {
    Rng rng(make_seed(0xFFFFu, scene_seed * 4096u + (uint32_t)s, 7));
    const double y_lo = -sc.half_len + s * 100.0;
    for (int b = 0; b < 8; ++b) {
      const double side = (rng.uniform() < 0.5) ? -1.0 : 1.0;
      const double cx = side * (6.5 + 4.0 * rng.uniform());
      const double cy = y_lo + 4.0 + 92.0 * rng.uniform();
      const double hx = 0.4 + 0.8 * rng.uniform(), hy = 0.5 + 1.5 * rng.uniform(),
                   hz = 0.8 + 2.5 * rng.uniform();
      sc.boxes.push_back({{cx - hx, cy - hy, sc.zg}, {cx + hx, cy + hy, sc.zg + hz}});
    }
    for (int c = 0; c < 4; ++c) {
      const double side = (rng.uniform() < 0.5) ? -1.0 : 1.0;
      const double cx = side * (6.5 + 4.0 * rng.uniform());
      const double cy = y_lo + 4.0 + 92.0 * rng.uniform();
      sc.cyls.push_back({cx, cy, 0.15 + 0.35 * rng.uniform(), sc.zg, sc.zg + 3.0 + 4.0 * rng.uniform()});
    }
  }
  return sc;
}

inline bool inside(const Box& b, V3 p, double eps) {
  return p.x >= b.lo.x - eps && p.x <= b.hi.x + eps && p.y >= b.lo.y - eps && p.y <= b.hi.y + eps &&
         p.z >= b.lo.z - eps && p.z <= b.hi.z + eps;
}

// Exit parameter of ray o + t d from box b (o inside b).  Returns t and the inward face normal.
inline double exit_box(const Box& b, V3 o, V3 d, V3* n_in) {
  double t = 1e300;
  V3 n{0, 0, 0};
  const double lo[3] = {b.lo.x, b.lo.y, b.lo.z}, hi[3] = {b.hi.x, b.hi.y, b.hi.z};
  const double oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
  for (int a = 0; a < 3; ++a) {
    if (dd[a] > 0) {
      const double ta = (hi[a] - oo[a]) / dd[a];
      if (ta < t) { t = ta; n = {0, 0, 0}; (a == 0 ? n.x : a == 1 ? n.y : n.z) = -1.0; }
    } else if (dd[a] < 0) {
      const double ta = (lo[a] - oo[a]) / dd[a];
      if (ta < t) { t = ta; n = {0, 0, 0}; (a == 0 ? n.x : a == 1 ? n.y : n.z) = 1.0; }
    }
  }
  *n_in = n;
  return t;
}

// First entry of ray into a solid axis-aligned box (o outside); t<0 if none.
inline double enter_box(const Box& b, V3 o, V3 d, V3* n_out) {
  double t0 = 0.0, t1 = 1e300;
  V3 n{0, 0, 0};
  const double lo[3] = {b.lo.x, b.lo.y, b.lo.z}, hi[3] = {b.hi.x, b.hi.y, b.hi.z};
  const double oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
  for (int a = 0; a < 3; ++a) {
    if (dd[a] == 0.0) {
      if (oo[a] < lo[a] || oo[a] > hi[a]) return -1.0;
      continue;
    }
    double ta = (lo[a] - oo[a]) / dd[a], tb = (hi[a] - oo[a]) / dd[a];
    double sgn = -1.0;
    if (ta > tb) { std::swap(ta, tb); sgn = 1.0; }
    if (ta > t0) { t0 = ta; n = {0, 0, 0}; (a == 0 ? n.x : a == 1 ? n.y : n.z) = sgn; }
    if (tb < t1) t1 = tb;
    if (t0 > t1) return -1.0;
  }
  if (t0 <= 0.0) return -1.0;
  *n_out = n;
  return t0;
}

inline double enter_cyl(const Cyl& c, V3 o, V3 d, V3* n_out) {
  double best = -1.0;
  // side
  const double ox = o.x - c.cx, oy = o.y - c.cy;
  const double a = d.x * d.x + d.y * d.y;
  if (a > 1e-18) {
    const double b = ox * d.x + oy * d.y, cc = ox * ox + oy * oy - c.r * c.r;
    const double disc = b * b - a * cc;
    if (disc > 0) {
      const double t = (-b - std::sqrt(disc)) / a;
      if (t > 0) {
        const double z = o.z + t * d.z;
        if (z >= c.z0 && z <= c.z1) {
          best = t;
          *n_out = {(ox + t * d.x) / c.r, (oy + t * d.y) / c.r, 0.0};
        }
      }
    }
  }
  // top cap
  if (d.z < 0 && o.z > c.z1) {
    const double t = (c.z1 - o.z) / d.z;
    const double x = ox + t * d.x, y = oy + t * d.y;
    if (x * x + y * y <= c.r * c.r && (best < 0 || t < best)) { best = t; *n_out = {0, 0, 1}; }
  }
  return best;
}

// Nearest hit along a ray starting inside free space.  Returns range (d is unit) and the unit
// surface normal oriented against the ray (toward the sensor).
double cast(const Scene& sc, V3 o, V3 d, V3* normal) {
  // walk through the union of free boxes
  double t_acc = 0.0;
  V3 p = o, n_wall{0, 0, 0};
  int cur = -1;
  for (size_t i = 0; i < sc.free_boxes.size(); ++i)
    if (inside(sc.free_boxes[i], p, 0.0)) { cur = (int)i; break; }
  if (cur < 0) cur = 0;
  for (int hop = 0; hop < 8; ++hop) {
    V3 n;
    const double t = exit_box(sc.free_boxes[cur], p, d, &n);
    t_acc += t;
    p = p + t * d;
    n_wall = n;
    // does the exit point open into another free box?
    int nxt = -1;
    const V3 probe = p + 1e-6 * d;
    if (cur == 0) {
      for (size_t i = 1; i < sc.free_boxes.size(); ++i)
        if (inside(sc.free_boxes[i], probe, 0.0)) { nxt = (int)i; break; }
    } else if (inside(sc.free_boxes[0], probe, 0.0)) {
      nxt = 0;
    }
    if (nxt < 0) break;
    cur = nxt;
  }
  double best = t_acc;
  V3 nb = n_wall;
  for (const Box& b : sc.boxes) {
    V3 n;
    const double t = enter_box(b, o, d, &n);
    if (t > 0 && t < best) { best = t; nb = n; }
  }
  for (const Cyl& c : sc.cyls) {
    V3 n;
    const double t = enter_cyl(c, o, d, &n);
    if (t > 0 && t < best) { best = t; nb = n; }
  }
  if (dot(nb, d) > 0) nb = -1.0 * nb;
  *normal = nb;
  return best;
}

// --------------------------------------------------------------- trajectory ---------------------
// Pose k of sequence s: forward along world +y at 0.8 +- 0.1 m/scan, sinusoidal heading (yaw rate
// <= 1.5 deg/scan), roll/pitch <= 0.3 deg, z bounce <= 2 cm.  Closed form in k (no accumulation of
// noise) except for the along-track distance, which is a seeded prefix sum.
struct Traj {
  std::vector<Pose> truth, odom;
};

Traj make_traj(uint32_t seq, int n_poses, double y_start, double speed) {
  Traj tr;
  tr.truth.resize(n_poses);
  tr.odom.resize(n_poses);
  Rng rs(make_seed(seq, 0, 1));
  const double period = 60.0 + 10.0 * rs.uniform();
  const double amp = (1.5 * kPi / 180.0) * period / (2.0 * kPi) * 0.9;  // max yaw rate 1.35 deg/scan
  const double ph = 2.0 * kPi * rs.uniform(), ph2 = 2.0 * kPi * rs.uniform(), ph3 = 2.0 * kPi * rs.uniform();
  double x = -1.0 + 2.0 * rs.uniform(), y = y_start;
  for (int k = 0; k < n_poses; ++k) {
    Rng rk(make_seed(seq, (uint32_t)k, 2));
    const double theta = amp * std::sin(2.0 * kPi * k / period + ph);
    const double yaw = 0.5 * kPi + theta;
    const double pitch = (0.3 * kPi / 180.0) * std::sin(2.0 * kPi * k / 17.0 + ph2);
    const double roll = (0.3 * kPi / 180.0) * std::sin(2.0 * kPi * k / 23.0 + ph3);
    const double z = 0.02 * std::sin(2.0 * kPi * k / 11.0 + ph2);
    tr.truth[k].R = rot_zyx(yaw, pitch, roll);
    tr.truth[k].t = {x, y, z};
    const double v = speed + (speed * 0.125) * (2.0 * rk.uniform() - 1.0);
    x += v * std::cos(yaw);
    y += v * std::sin(yaw);
  }
  // odometry = truth o noise(sigma_t = 0.03 m, sigma_r = 0.2 deg), composed incrementally
  tr.odom[0] = tr.truth[0];
  for (int k = 1; k < n_poses; ++k) {
    Rng rk(make_seed(seq, (uint32_t)k, 3));
    // relative truth
    M3 Rrel;
    {
      const M3& A = tr.truth[k - 1].R;
      const M3& B = tr.truth[k].R;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
          Rrel.m[3 * i + j] = A.m[i] * B.m[j] + A.m[3 + i] * B.m[3 + j] + A.m[6 + i] * B.m[6 + j];
    }
    V3 trel = mulT(tr.truth[k - 1].R, tr.truth[k].t - tr.truth[k - 1].t);
    const double sr = 0.2 * kPi / 180.0;
    M3 N = rot_zyx(sr * rk.normal(), sr * rk.normal(), sr * rk.normal());
    trel = trel + V3{0.03 * rk.normal(), 0.03 * rk.normal(), 0.03 * rk.normal()};
    Rrel = mul(Rrel, N);
    tr.odom[k].R = mul(tr.odom[k - 1].R, Rrel);
    tr.odom[k].t = tr.odom[k - 1].t + mul(tr.odom[k - 1].R, trel);
  }
  return tr;
}

void pose_to_colmajor(const Pose& P, double* T16) {
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T16[c * 4 + r] = P.R.m[3 * r + c];
    T16[12 + r] = (r == 0 ? P.t.x : r == 1 ? P.t.y : P.t.z);
  }
  T16[3] = T16[7] = T16[11] = 0.0;
  T16[15] = 1.0;
}

}  // namespace

extern "C" {

// sensor: 0 = HDL-64-shaped (64 rings x 2048 = 131072 rays), 1 = VLS-128-shaped (128 x 2048).
int ls_synth_num_rays(int sensor) { return (sensor == 1 ? 128 : 64) * 2048; }

// Poses of sequence `seq` as column-major 4x4 doubles (T_world_sensor), truth and odometry.
void ls_synth_trajectory(uint32_t seq, int n_poses, double y_start, double speed, double* truth16,
                         double* odom16) {
  Traj tr = make_traj(seq, n_poses, y_start, speed);
  for (int k = 0; k < n_poses; ++k) {
    if (truth16) pose_to_colmajor(tr.truth[k], truth16 + 16 * k);
    if (odom16) pose_to_colmajor(tr.odom[k], odom16 + 16 * k);
  }
}

// One scan.  features: 4*N floats (x,y,z,1 per point, i.e. DataPoints::features.data());
// normals: 3*N floats (DataPoints descriptor "normals").  Point index = ring*2048 + azimuth_step.
// T_ws16: column-major T_world_sensor (double) of the pose the scan is taken from.
// Returns N.
int ls_synth_scan(int sensor, uint32_t scene_seed, double scene_length_m, uint32_t seq, uint32_t scan_idx,
                  const double* T_ws16, double range_sigma, float* features, float* normals) {
  static thread_local Scene sc;
  static thread_local uint32_t sc_seed = 0xFFFFFFFFu;
  static thread_local double sc_len = -1.0;
  if (sc_seed != scene_seed || sc_len != scene_length_m) {
    sc = build_scene(scene_seed, scene_length_m);
    sc_seed = scene_seed;
    sc_len = scene_length_m;
  }
  Pose P;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) P.R.m[3 * r + c] = T_ws16[c * 4 + r];
  P.t = {T_ws16[12], T_ws16[13], T_ws16[14]};
  const int n_rings = (sensor == 1) ? 128 : 64, n_az = 2048;
  std::vector<double> elev(n_rings);
  if (sensor == 1) {
    for (int r = 0; r < n_rings; ++r) elev[r] = (15.0 - 40.0 * r / 127.0) * kPi / 180.0;
  } else {
    for (int r = 0; r < 32; ++r) elev[r] = (2.0 - r / 3.0) * kPi / 180.0;            // +2 .. -8.33
    for (int r = 0; r < 32; ++r) elev[32 + r] = (-8.83 - 0.5 * r) * kPi / 180.0;    // -8.83 .. -24.33
  }
  Rng rng(make_seed(seq, scan_idx, 4));
  int i = 0;
  for (int r = 0; r < n_rings; ++r) {
    const double ce = std::cos(elev[r]), se = std::sin(elev[r]);
    for (int j = 0; j < n_az; ++j, ++i) {
      const double az = 2.0 * kPi * j / n_az;
      const V3 ds{ce * std::cos(az), ce * std::sin(az), se};
      const V3 dw = mul(P.R, ds);
      V3 nw;
      double range = cast(sc, P.t, dw, &nw);
      range += range_sigma * rng.normal();
      if (range < 0.05) range = 0.05;
      const V3 ps = range * ds;
      const V3 ns = mulT(P.R, nw);
      features[4 * i + 0] = (float)ps.x;
      features[4 * i + 1] = (float)ps.y;
      features[4 * i + 2] = (float)ps.z;
      features[4 * i + 3] = 1.0f;
      // unit float32 normal
      float nx = (float)ns.x, ny = (float)ns.y, nz = (float)ns.z;
      const float inv = 1.0f / std::sqrt(nx * nx + ny * ny + nz * nz);
      normals[3 * i + 0] = nx * inv;
      normals[3 * i + 1] = ny * inv;
      normals[3 * i + 2] = nz * inv;
    }
  }
  return i;
}

}  // extern "C"
