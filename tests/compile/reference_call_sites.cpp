// Compile-only check of the drop-in boundary: every member of laser_slam::LaserTrack /
// IncrementalEstimator and of the third-party types in their signatures is used here the way the reference's own callers
// use it -- laser_slam_ros/src/laser_slam_worker.cpp:47,133-173,197,263-272,519 (the ROS worker),
// laser_slam/src/laser_track.cpp:14-64,137-146,172,229,262-265,424-428,431-458,485-499 and
// laser_slam/src/incremental_estimator.cpp:17-20,56-60,108,117-125,156-161,258-277 (the library's use of
// libpointmatcher / GTSAM / minkindr / mincurves).  tests/test_abi.py compiles this translation unit against include/
// (g++ -fsyntax-only); nothing here runs.
#include <fstream>
#include <memory>
#include <sstream>
#include <vector>

#include "laser_slam/incremental_estimator.hpp"
#include "laser_slam/laser_track.hpp"

namespace laser_slam {  // the reference's sources live in this namespace: PointMatcher, SE3, ... resolve as they do there

// ---- the worker's scan callback and accessors -----------------------------------------------------------------------
void worker_call_sites(std::shared_ptr<IncrementalEstimator> incremental_estimator, unsigned int worker_id, const Pose& new_pose,
                       const LaserScan& new_scan) {
  std::shared_ptr<LaserTrack> laser_track = incremental_estimator->getLaserTrack(worker_id);       // worker.cpp:47
  gtsam::NonlinearFactorGraph new_factors;
  gtsam::Values new_values;
  bool is_prior = false;
  laser_track->processPoseAndLaserScan(new_pose, new_scan, &new_factors, &new_values, &is_prior);  // :133, :158
  if (laser_track->getNumScans() > 1u) {                                                           // :140
    Pose current = laser_track->getCurrentPose();                                                  // :141
    (void)laser_track->evaluate(laser_track->getMinTime());                                        // :145-148
    (void)laser_track->getMaxTime();
    (void)current;
  }
  gtsam::Values result;
  if (is_prior) result = incremental_estimator->registerPrior(new_factors, new_values, worker_id); // :167
  else result = incremental_estimator->estimate(new_factors, new_values, new_scan.time_ns);        // :169
  laser_track->updateFromGTSAMValues(result);                                                      // :173
  DataPoints local_cloud;
  laser_track->getLocalCloudInWorldFrame(laser_track->getMaxTime(), &local_cloud);                 // :197
  for (const auto& track : incremental_estimator->getAllLaserTracks()) {                           // :263-272
    Trajectory trajectory;
    track->getTrajectory(&trajectory);
    const std::vector<LaserScan>& scans = track->getLaserScans();
    (void)scans;
  }
  Trajectory odometry;
  laser_track->getOdometryTrajectory(&odometry);                                                   // :519
  // SegMatch side
  RelativePose loop_closure;
  incremental_estimator->processLoopClosure(loop_closure);
  DataPoints sub_map;
  laser_track->buildSubMapAroundTime(laser_track->getMaxTime(), 3u, &sub_map);
  gtsam::Expression<SE3> leaf = laser_track->getValueExpression(laser_track->getMaxTime());
  std::map<Time, double> times;
  laser_track->getScanMatchingTimes(&times);
  laser_track->saveTrajectory("trajectory.csv");
  laser_track->printTrajectory();
  (void)leaf;
  (void)incremental_estimator->getCurrentPose(worker_id);
}

// ---- libpointmatcher as laser_track.cpp / incremental_estimator.cpp use it -------------------------------------------
void pointmatcher_call_sites(const LaserTrackParams& params, DataPoints* scan, const DataPoints& reading, const DataPoints& sub_map) {
  PointMatcher::ICP icp_;
  std::ifstream ifs_icp_configurations(params.icp_configuration_file.c_str());                    // laser_track.cpp:14-21
  if (ifs_icp_configurations.good()) icp_.loadFromYaml(ifs_icp_configurations);
  else icp_.setDefault();
  std::ifstream ifs_input_filters(params.icp_input_filters_file.c_str());                          // :22-30
  PointMatcher::DataPointsFilters input_filters_(ifs_input_filters);
  input_filters_.apply(*scan);                                                                     // :146
  std::shared_ptr<PointMatcher::Transformation> rigid_transformation_ =
      std::shared_ptr<PointMatcher::Transformation>(PointMatcher::get().REG(Transformation).create("RigidTransformation"));  // :33
  PointMatcher::TransformationParameters transformation_matrix;
  correctTransformationMatrix(&transformation_matrix);                                             // :262, common.hpp:136-149
  if (!rigid_transformation_->checkParameters(transformation_matrix))
    transformation_matrix = rigid_transformation_->correctParameters(transformation_matrix);
  DataPoints moved = rigid_transformation_->compute(reading, transformation_matrix);               // :265, :485
  DataPoints concatenated = sub_map;
  concatenated.concatenate(moved);                                                                 // :485
  const size_t n_points = concatenated.getNbPoints() + (size_t)concatenated.features.cols();
  const size_t dim = (size_t)concatenated.features.rows();
  const float x0 = concatenated.features(0, 0);
  (void)n_points; (void)dim; (void)x0;
  PointMatcher::TransformationParameters icp_solution = transformation_matrix;
  try {
    icp_solution = icp_.compute(reading, concatenated, transformation_matrix);                     // :496, incremental_estimator.cpp:108
  } catch (PointMatcher::ConvergenceError error) {                                                 // :497-502
    icp_solution = transformation_matrix;
  }
  SE3 T_a_b = convertTransformationMatrixToSE3(icp_solution);                                      // :515, common.hpp:263-269
  (void)T_a_b;
}

// ---- GTSAM / minkindr as laser_track.cpp / incremental_estimator.cpp use them -----------------------------------------
void gtsam_call_sites(const LaserTrackParams& params, LaserTrack* track, const RelativePose& m, const Pose& prior) {
  using namespace gtsam;
  noiseModel::Base::shared_ptr odometry_noise = noiseModel::Diagonal::Sigmas(params.odometry_noise_model);   // laser_track.cpp:37-64
  noiseModel::Base::shared_ptr icp_noise =
      noiseModel::Robust::Create(noiseModel::mEstimator::Cauchy::Create(1), noiseModel::Diagonal::Sigmas(params.icp_noise_model));
  Expression<SE3> T_w_b(track->getValueExpression(m.time_b_ns));                                   // :431-451
  Expression<SE3> T_w_a(track->getValueExpression(m.time_a_ns));
  Expression<SE3> T_a_w(kindr::minimal::inverse(T_w_a));
  Expression<SE3> relative(kindr::minimal::compose(T_a_w, T_w_b));
  ExpressionFactor<SE3> relative_factor(icp_noise, m.T_a_b, relative);
  Expression<SE3> frozen(track->evaluate(m.time_a_ns));                                            // :440-444 (fix_first_node)
  ExpressionFactor<SE3> frozen_factor(odometry_noise, m.T_a_b, kindr::minimal::compose(kindr::minimal::inverse(frozen), T_w_b));
  ExpressionFactor<SE3> prior_factor(odometry_noise, prior.T_w, track->getValueExpression(prior.time_ns));  // :453-458
  NonlinearFactorGraph graph;
  graph.push_back(relative_factor);
  graph.push_back(frozen_factor);
  graph.push_back(prior_factor);
  track->appendPriorFactors(track->getMinTime(), &graph);                                          // :339-409
  track->appendOdometryFactors(track->getMinTime(), track->getMaxTime(), odometry_noise, &graph);
  track->appendICPFactors(track->getMinTime(), track->getMaxTime(), icp_noise, &graph);
  track->appendLoopClosureFactors(track->getMinTime(), track->getMaxTime(), icp_noise, &graph);
  KeySet keys = graph.keys();
  Values values;
  track->initializeGTSAMValues(keys, &values);                                                     // :411-414
  track->updateCovariancesFromGTSAMValues(graph, values);                                          // :421-429
  Marginals marginals(graph, values);
  Marginals::Matrix6 covariance = marginals.marginalCovariance(*keys.begin());
  (void)covariance;
  ISAM2Params isam2_params;                                                                        // incremental_estimator.cpp:17-20
  isam2_params.setRelinearizeSkip(1);
  isam2_params.setRelinearizeThreshold(0.001);
  ISAM2 isam2(isam2_params);
  ISAM2Result update_result = isam2.update(graph, values);                                         // :156-161
  update_result.print();
  isam2.update();
  isam2.update();
  std::vector<size_t> remove_indices(update_result.newFactorsIndices.begin(), update_result.newFactorsIndices.begin() + 1);
  isam2.update(NonlinearFactorGraph(), Values(), remove_indices);                                  // :258
  Values estimate = isam2.calculateEstimate();
  (void)estimate;
  SE3 T(SO3(1.0, 0.0, 0.0, 0.0), SE3::Position{0.0, 100.0, 0.0});                                  // laser_track.cpp:167-169
  SE3 inv = T.inverse() * T;
  (void)inv.getTransformationMatrix();
  (void)inv.getPosition();
  (void)inv.getRotation().w();
  std::vector<Covariance> covariances;
  track->getCovariances(&covariances);
}

}  // namespace laser_slam

int main() {
  laser_slam::IncrementalEstimator default_constructed;  // incremental_estimator.hpp:21
  (void)default_constructed;
  return 0;
}
