// laser_slam core value types -- same names and fields as reference
// laser_slam/include/laser_slam/common.hpp:14-20,83-133,263-269 (the CSV / Clock helpers of that header are
// support code outside the path and are not reproduced).
#ifndef LASER_SLAM_COMMON_HPP_
#define LASER_SLAM_COMMON_HPP_

#include <map>
#include <vector>

#include "laser_slam_compat/compat.hpp"

namespace laser_slam {

typedef ::PointMatcher<float> PointMatcher;
typedef PointMatcher::DataPoints DataPoints;
typedef kindr::minimal::QuatTransformationTemplate<double> SE3;
typedef SE3::Rotation SO3;
typedef curves::Time Time;
typedef gtsam::Key Key;

/// \brief Pose type including absolute transformation and time stamp (reference common.hpp:87-94).
struct Pose {
  SE3 T_w;
  curves::Time time_ns = 0;
  Key key = 0;
};

/// \brief RelativePose type (reference common.hpp:97-110).
struct RelativePose {
  SE3 T_a_b;
  curves::Time time_a_ns = 0;
  curves::Time time_b_ns = 0;
  Key key_a = 0;
  Key key_b = 0;
  unsigned int track_id_a = 0;
  unsigned int track_id_b = 0;
};

/// \brief LaserScan type (reference common.hpp:113-120).
struct LaserScan {
  DataPoints scan;
  curves::Time time_ns = 0;
  Key key = 0;
};

typedef std::vector<double> Covariance;
typedef std::vector<Pose> PoseVector;
typedef std::vector<RelativePose> RelativePoseVector;
typedef std::map<Time, SE3> Trajectory;

// correctTransformationMatrix (reference common.hpp:136-149): project onto a rigid transform if the
// 3x3 block fails RigidTransformation::checkParameters.
inline void correctTransformationMatrix(PointMatcher::TransformationParameters* T) {
  if (!ls_check_rigid(T->data())) {
    PointMatcher::TransformationParameters fixed;
    ls_correct_rigid(T->data(), fixed.data());
    *T = fixed;
  }
}

// convertTransformationMatrixToSE3 (reference common.hpp:263-269): float 4x4 -> double, renormalised rotation.
inline SE3 convertTransformationMatrixToSE3(const PointMatcher::TransformationParameters& T) {
  std::array<double, 9> R;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[3 * r + c] = (double)T(r, c);
  return SE3(SO3::constructAndRenormalize(R), SE3::Position{(double)T(0, 3), (double)T(1, 3), (double)T(2, 3)});
}

}  // namespace laser_slam

#endif  // LASER_SLAM_COMMON_HPP_
