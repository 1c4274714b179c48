// Velodyne packet assembler / de-skew, the ROS-free core of
// VelodyneAssemblerRos::pclCallback (reference sensor_drivers/velodyne_assembler/src/velodyne_assembler_ros.cpp:57-143).
//
// The reference transforms every packet on the CPU as it arrives and transforms the whole assembled cloud once more
// before publishing.  Here the host only keeps the books -- which packets belong to the revolution and the float32 4x4
// that takes each of them to the revolution's start -- and the points are moved once per revolution by one kernel
// (ls_deskew_revolution): out = T_final (x) (T_packet (x) p), the same two float32 transforms in the same order.
#ifndef LASER_SLAM_VELODYNE_ASSEMBLER_HPP_
#define LASER_SLAM_VELODYNE_ASSEMBLER_HPP_

#include <cmath>
#include <stdexcept>
#include <vector>

#include "laser_slam/common.hpp"

namespace laser_slam {

class VelodyneAssembler {
 public:
  typedef PointMatcher::TransformationParameters Matrix4;

  // T_sensor_base: static offset of the sensor w.r.t. the vehicle (reference :37-52); naive_assembling ignores the
  // vehicle motion (:81-82); cuda_device: the GPU that moves the points.
  explicit VelodyneAssembler(const Matrix4& T_sensor_base = Matrix4(), bool naive_assembling = false, int cuda_device = 0)
      : T_sensor_base_(T_sensor_base), T_base_sensor_(rigidInverse(T_sensor_base)), naive_(naive_assembling), device_(cuda_device) {}

  // One packet cloud in the sensor frame with the vehicle pose T_fixed_base at its stamp (what tf returns, :84-95).
  // Returns true when this packet starts a new revolution; *revolution then holds the finished one, expressed in the
  // sensor frame at its LAST packet (:105-108), and *stamp_ns that packet's stamp (:110).
  bool addPacket(const DataPoints& cloud_in, const Matrix4& T_fixed_base_current, Time stamp_ns, DataPoints* revolution,
                 Time* stamp_out_ns) {
    const size_t n = cloud_in.getNbPoints();
    if (n == 0) return false;  // :78
    const Matrix4 T_cur = naive_ ? Matrix4() : T_fixed_base_current;
    const Matrix4 T_basePrevious_baseCurrent = multiply(rigidInverse(T_fixed_basePrevious_), T_cur);  // :95-96
    T_fixed_basePrevious_ = T_cur;
    const double current_azimuth_rad = std::atan2((double)cloud_in.features(1, 0), (double)cloud_in.features(0, 0));  // :101
    bool published = false;
    if ((last_azimuth_rad_ > kStartAngleRad && current_azimuth_rad <= kStartAngleRad) || !initialized_) {  // :102-103
      if (initialized_) {
        if (revolution == NULL) throw std::runtime_error("VelodyneAssembler: null output");
        finish(rigidInverse(T_sensorStart_sensorCurrent_), revolution);
        if (stamp_out_ns != NULL) *stamp_out_ns = last_stamp_;
        published = true;
      }
      points_.clear();
      offsets_.assign(1, 0);
      T_packets_.clear();
      initialized_ = true;
      T_sensorStart_sensorCurrent_ = Matrix4();  // :121
      append(cloud_in, T_sensorStart_sensorCurrent_);   // the first packet is taken as it is (:118)
    } else {
      const Matrix4 T_sensorPrevious_sensorCurrent = multiply(multiply(T_sensor_base_, T_basePrevious_baseCurrent), T_base_sensor_);  // :124-125
      T_sensorStart_sensorCurrent_ = multiply(T_sensorStart_sensorCurrent_, T_sensorPrevious_sensorCurrent);                          // :128
      append(cloud_in, T_sensorStart_sensorCurrent_);  // :131-134
    }
    last_azimuth_rad_ = current_azimuth_rad;
    last_stamp_ = stamp_ns;
    return published;
  }

  size_t pointsInProgress() const { return points_.size() / 4; }
  size_t packetsInProgress() const { return T_packets_.size() / 16; }

  // float32 helpers, every operation rounded in a fixed order (oracle/__init__.py restates them)
  static Matrix4 multiply(const Matrix4& A, const Matrix4& B) {
    Matrix4 C;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        float s = A(i, 0) * B(0, j);
        float t = A(i, 1) * B(1, j);
        s = s + t;
        t = A(i, 2) * B(2, j);
        s = s + t;
        t = A(i, 3) * B(3, j);
        C(i, j) = s + t;
      }
    return C;
  }
  // [DEFINED] rigid inverse [R^T, -(R^T t)] where the reference calls Eigen's general 4x4 inverse (:95, :107)
  static Matrix4 rigidInverse(const Matrix4& T) {
    Matrix4 O;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) O(i, j) = T(j, i);
    for (int i = 0; i < 3; ++i) {
      float a = T(0, i) * T(0, 3);
      float b = T(1, i) * T(1, 3);
      float c = T(2, i) * T(2, 3);
      float s = a + b;
      s = s + c;
      O(i, 3) = -s;
    }
    return O;
  }

 private:
  static constexpr double kStartAngleRad = 1.57079632679489661923;  // M_PI / 2 (:100)

  void append(const DataPoints& cloud, const Matrix4& T) {
    const size_t n = cloud.getNbPoints();
    const float* f = cloud.features.data();
    points_.insert(points_.end(), f, f + 4 * n);
    offsets_.push_back((int)(points_.size() / 4));
    T_packets_.insert(T_packets_.end(), T.data(), T.data() + 16);
  }

  void finish(const Matrix4& T_final, DataPoints* out) const {
    const size_t m = points_.size() / 4;
    std::vector<float> moved(4 * (m ? m : 1));
    const int rc = ls_deskew_revolution(device_, points_.data(), offsets_.data(), (int)(offsets_.size() - 1), T_packets_.data(),
                                        T_final.data(), moved.data());
    if (rc != LS_OK) throw std::runtime_error("ls_deskew_revolution failed (no usable CUDA device? there is no CPU fallback)");
    *out = DataPoints::fromArrays(moved.data(), NULL, m);
  }

  Matrix4 T_sensor_base_, T_base_sensor_;
  bool naive_ = false;
  int device_ = 0;
  Matrix4 T_fixed_basePrevious_, T_sensorStart_sensorCurrent_;
  bool initialized_ = false;
  double last_azimuth_rad_ = 0.0;
  Time last_stamp_ = 0;
  std::vector<float> points_;     // packets of the revolution in progress, concatenated, untouched
  std::vector<int> offsets_{0};   // packet k = points [offsets_[k], offsets_[k+1])
  std::vector<float> T_packets_;  // 16 floats per packet, column-major
};

}  // namespace laser_slam

#endif  // LASER_SLAM_VELODYNE_ASSEMBLER_HPP_
