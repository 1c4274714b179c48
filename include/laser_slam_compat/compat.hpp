// Minimal stand-ins for the third-party types that appear in laser_slam's public API
// (reference laser_slam/include/laser_slam/common.hpp:6-20,87-133; SURVEY.md §8b lists the members used).
// They exist ONLY because Eigen, libpointmatcher, GTSAM, minkindr and mincurves are absent from this build
// environment; with the real libraries present these few types are what an adapter would map 1:1
// (INTEGRATION.md).  Header-only, no dependencies.
#ifndef LASER_SLAM_COMPAT_HPP_
#define LASER_SLAM_COMPAT_HPP_

#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ls_b200.h"

// ------------------------------------------------------------------------------------------------ curves::Time
namespace curves {
typedef int64_t Time;  // nanoseconds (mincurves)
}

// ------------------------------------------------------------------------------------------------ kindr::minimal
namespace kindr {
namespace minimal {

typedef std::array<double, 3> Position;
typedef std::array<double, 16> Matrix4d;  // column-major, like Eigen::Matrix4d::data()

// double-precision unit quaternion (w, x, y, z), Hamilton convention
class RotationQuaternion {
 public:
  RotationQuaternion() : q_{1, 0, 0, 0} {}
  RotationQuaternion(double w, double x, double y, double z) : q_{w, x, y, z} {}
  double w() const { return q_[0]; }
  double x() const { return q_[1]; }
  double y() const { return q_[2]; }
  double z() const { return q_[3]; }
  std::array<double, 9> getRotationMatrix() const {  // row-major 3x3
    const double w = q_[0], x = q_[1], y = q_[2], z = q_[3];
    return {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
  }
  // SO3::constructAndRenormalize(R): nearest unit quaternion of an approximately orthonormal matrix
  static RotationQuaternion constructAndRenormalize(const std::array<double, 9>& m) {
    double q[4];
    const double t = m[0] + m[4] + m[8];
    if (t > 0) {
      const double s = std::sqrt(t + 1.0) * 2;
      q[0] = 0.25 * s; q[1] = (m[7] - m[5]) / s; q[2] = (m[2] - m[6]) / s; q[3] = (m[3] - m[1]) / s;
    } else if (m[0] > m[4] && m[0] > m[8]) {
      const double s = std::sqrt(1.0 + m[0] - m[4] - m[8]) * 2;
      q[0] = (m[7] - m[5]) / s; q[1] = 0.25 * s; q[2] = (m[1] + m[3]) / s; q[3] = (m[2] + m[6]) / s;
    } else if (m[4] > m[8]) {
      const double s = std::sqrt(1.0 + m[4] - m[0] - m[8]) * 2;
      q[0] = (m[2] - m[6]) / s; q[1] = (m[1] + m[3]) / s; q[2] = 0.25 * s; q[3] = (m[5] + m[7]) / s;
    } else {
      const double s = std::sqrt(1.0 + m[8] - m[0] - m[4]) * 2;
      q[0] = (m[3] - m[1]) / s; q[1] = (m[2] + m[6]) / s; q[2] = (m[5] + m[7]) / s; q[3] = 0.25 * s;
    }
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double sgn = q[0] >= 0 ? 1.0 : -1.0;
    return RotationQuaternion(sgn * q[0] / n, sgn * q[1] / n, sgn * q[2] / n, sgn * q[3] / n);
  }
  RotationQuaternion inverse() const { return RotationQuaternion(q_[0], -q_[1], -q_[2], -q_[3]); }
  RotationQuaternion operator*(const RotationQuaternion& o) const {
    const double* a = q_.data();
    const double* b = o.q_.data();
    RotationQuaternion r(a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                         a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                         a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                         a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]);
    const double n = std::sqrt(r.q_[0] * r.q_[0] + r.q_[1] * r.q_[1] + r.q_[2] * r.q_[2] + r.q_[3] * r.q_[3]);
    for (double& v : r.q_) v /= n;
    return r;
  }
  Position rotate(const Position& v) const {
    const std::array<double, 9> R = getRotationMatrix();
    return {R[0] * v[0] + R[1] * v[1] + R[2] * v[2], R[3] * v[0] + R[4] * v[1] + R[5] * v[2],
            R[6] * v[0] + R[7] * v[1] + R[8] * v[2]};
  }

 private:
  std::array<double, 4> q_;
};

// T_a_b: maps coordinates of frame b into frame a (reference common.hpp:97-110)
template <typename Scalar>
class QuatTransformationTemplate {
 public:
  typedef RotationQuaternion Rotation;
  typedef kindr::minimal::Position Position;
  QuatTransformationTemplate() : p_{0, 0, 0} {}
  QuatTransformationTemplate(const Rotation& q, const Position& p) : q_(q), p_(p) {}
  const Rotation& getRotation() const { return q_; }
  const Position& getPosition() const { return p_; }
  QuatTransformationTemplate inverse() const {
    const Rotation qi = q_.inverse();
    const Position t = qi.rotate(p_);
    return QuatTransformationTemplate(qi, Position{-t[0], -t[1], -t[2]});
  }
  QuatTransformationTemplate operator*(const QuatTransformationTemplate& o) const {
    const Position t = q_.rotate(o.p_);
    return QuatTransformationTemplate(q_ * o.q_, Position{t[0] + p_[0], t[1] + p_[1], t[2] + p_[2]});
  }
  Matrix4d getTransformationMatrix() const {
    const std::array<double, 9> R = q_.getRotationMatrix();
    Matrix4d T{};
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) T[c * 4 + r] = R[3 * r + c];
      T[12 + r] = p_[r];
    }
    T[15] = 1.0;
    return T;
  }
  // {qw,qx,qy,qz,tx,ty,tz}: the pose layout of the C ABI
  void toArray7(double* out) const {
    out[0] = q_.w(); out[1] = q_.x(); out[2] = q_.y(); out[3] = q_.z();
    out[4] = p_[0]; out[5] = p_[1]; out[6] = p_[2];
  }
  static QuatTransformationTemplate fromArray7(const double* a) {
    return QuatTransformationTemplate(Rotation(a[0], a[1], a[2], a[3]), Position{a[4], a[5], a[6]});
  }

 private:
  Rotation q_;
  Position p_;
};

}  // namespace minimal
}  // namespace kindr

// ------------------------------------------------------------------------------------------------ PointMatcher<float>
// DataPoints memory layout as libpointmatcher: `features` column-major (dim+1) x N, `descriptors` D x N.
template <typename T>
struct PointMatcher {
  struct ConvergenceError : std::runtime_error {
    explicit ConvergenceError(const std::string& m) : std::runtime_error(m) {}
  };
  // 4x4, column-major (Eigen default), data() is what the C ABI takes
  struct TransformationParameters {
    std::array<T, 16> m;
    TransformationParameters() { setIdentity(); }
    void setIdentity() { m.fill(T(0)); m[0] = m[5] = m[10] = m[15] = T(1); }
    T& operator()(int r, int c) { return m[c * 4 + r]; }
    T operator()(int r, int c) const { return m[c * 4 + r]; }
    T* data() { return m.data(); }
    const T* data() const { return m.data(); }
    template <typename S>
    static TransformationParameters cast(const std::array<S, 16>& src) {
      TransformationParameters t;
      for (int i = 0; i < 16; ++i) t.m[i] = static_cast<T>(src[i]);
      return t;
    }
  };
  struct Label {
    std::string text;
    size_t span;
  };
  struct DataPoints {
    std::vector<T> features;     // 4 x N, column-major: x,y,z,1 per point
    std::vector<T> descriptors;  // D x N, column-major
    std::vector<Label> featureLabels, descriptorLabels;
    size_t descriptorDim = 0;
    size_t getNbPoints() const { return features.size() / 4; }
    bool descriptorExists(const std::string& name) const {
      for (const auto& l : descriptorLabels)
        if (l.text == name) return true;
      return false;
    }
    // row offset of a descriptor inside a descriptor column, or -1
    int descriptorOffset(const std::string& name) const {
      size_t off = 0;
      for (const auto& l : descriptorLabels) {
        if (l.text == name) return (int)off;
        off += l.span;
      }
      return -1;
    }
    void concatenate(const DataPoints& o) {
      if (getNbPoints() == 0) { *this = o; return; }
      if (o.descriptorDim != descriptorDim) throw std::runtime_error("DataPoints::concatenate: descriptor mismatch");
      features.insert(features.end(), o.features.begin(), o.features.end());
      descriptors.insert(descriptors.end(), o.descriptors.begin(), o.descriptors.end());
    }
    // convenience: cloud with a 3-row "normals" descriptor
    static DataPoints fromArrays(const T* feat4, const T* normals3, size_t n) {
      DataPoints d;
      d.features.assign(feat4, feat4 + 4 * n);
      d.featureLabels = {{"x", 1}, {"y", 1}, {"z", 1}, {"pad", 1}};
      if (normals3) {
        d.descriptors.assign(normals3, normals3 + 3 * n);
        d.descriptorLabels = {{"normals", 3}};
        d.descriptorDim = 3;
      }
      return d;
    }
  };
};

// ------------------------------------------------------------------------------------------------ gtsam
namespace gtsam {

typedef uint64_t Key;

// The only factor kinds laser_slam builds are ExpressionFactor<SE3> priors and relative-pose factors
// (reference laser_track.cpp:431-458); a graph is therefore a list of ls_factor records.
class NonlinearFactorGraph {
 public:
  void push_back(const ls_factor& f) { factors_.push_back(f); }
  bool empty() const { return factors_.empty(); }
  size_t size() const { return factors_.size(); }
  void clear() { factors_.clear(); }
  const ls_factor& at(size_t i) const { return factors_.at(i); }
  const std::vector<ls_factor>& factors() const { return factors_; }

 private:
  std::vector<ls_factor> factors_;
};

class Values {
 public:
  typedef kindr::minimal::QuatTransformationTemplate<double> SE3;
  void clear() { v_.clear(); }
  void insert(Key k, const SE3& T) {
    if (!v_.emplace(k, T).second) throw std::runtime_error("Values::insert: key already exists");
  }
  bool exists(Key k) const { return v_.count(k) != 0; }
  const SE3& at(Key k) const { return v_.at(k); }
  size_t size() const { return v_.size(); }
  bool empty() const { return v_.empty(); }
  std::map<Key, SE3>::const_iterator begin() const { return v_.begin(); }
  std::map<Key, SE3>::const_iterator end() const { return v_.end(); }

 private:
  std::map<Key, SE3> v_;
};

// noiseModel::Diagonal::Sigmas / Robust::Create(Cauchy(1), Diagonal)
struct NoiseModel {
  std::array<double, 6> sigmas;
  bool cauchy;
};

}  // namespace gtsam

#endif  // LASER_SLAM_COMPAT_HPP_
