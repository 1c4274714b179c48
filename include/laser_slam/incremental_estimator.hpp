// IncrementalEstimator -- public interface of reference laser_slam/include/laser_slam/incremental_estimator.hpp:17-81;
// the iSAM2 object is replaced by a device pose graph (ls_pg_*).
#ifndef LASER_SLAM_INCREMENTAL_ESTIMATOR_HPP_
#define LASER_SLAM_INCREMENTAL_ESTIMATOR_HPP_

#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "laser_slam/common.hpp"
#include "laser_slam/laser_track.hpp"
#include "laser_slam/parameters.hpp"

namespace laser_slam {

class IncrementalEstimator {
 public:
  IncrementalEstimator() {}
  explicit IncrementalEstimator(const EstimatorParams& parameters, unsigned int n_laser_slam_workers = 1u);
  ~IncrementalEstimator();
  IncrementalEstimator(const IncrementalEstimator&) = delete;
  IncrementalEstimator& operator=(const IncrementalEstimator&) = delete;

  void processLoopClosure(const RelativePose& loop_closure);
  Pose getCurrentPose(unsigned int laser_track_id = 0u) const;
  std::shared_ptr<LaserTrack> getLaserTrack(unsigned int laser_track_id);
  std::vector<std::shared_ptr<LaserTrack> > getAllLaserTracks();

  gtsam::Values estimate(const gtsam::NonlinearFactorGraph& new_factors, const gtsam::Values& new_values,
                         laser_slam::Time timestamp_ns = 0u);
  gtsam::Values estimateAndRemove(const gtsam::NonlinearFactorGraph& new_factors,
                                  const gtsam::NonlinearFactorGraph& new_associations_factors,
                                  const gtsam::Values& new_values, const std::vector<unsigned int>& affected_worker_ids,
                                  laser_slam::Time timestamp_ns = 0u);
  gtsam::Values registerPrior(const gtsam::NonlinearFactorGraph& new_factors, const gtsam::Values& new_values,
                              const unsigned int worker_id);

  // (new) One scan callback of SEVERAL workers at once: poses[i] / scans[i] belong to worker worker_ids[i].  Equivalent to
  // calling getLaserTrack(worker_ids[i])->processPoseAndLaserScan(...) for every i (reference
  // laser_slam_ros/src/laser_slam_worker.cpp:133,158), but the scan-to-sub-map registrations of the step run as ONE
  // batched launch on the estimator's device context (ls_icp_register_submap_batch): the tracks are independent, so
  // the results are the same bits.  Outputs are indexed like the inputs.
  void processPosesAndLaserScans(const std::vector<unsigned int>& worker_ids, const std::vector<Pose>& poses,
                                 const std::vector<LaserScan>& scans, std::vector<gtsam::NonlinearFactorGraph>* new_factors,
                                 std::vector<gtsam::Values>* new_values, std::vector<bool>* is_prior);
  // (new) The same in two halves, for callers that have something to do while the GPU registers: begin() stages every
  // track's registration and launches, end() waits and finishes the callbacks.  In between only prefetchLaserScans may be
  // called: it uploads scans that a later begin() will be handed (recognised by their time stamps), so the next step's
  // host-to-device copies overlap this step's ICP iterations.  Purely an overlap device: results never change.
  void beginPosesAndLaserScans(const std::vector<unsigned int>& worker_ids, const std::vector<Pose>& poses,
                               const std::vector<LaserScan>& scans);
  void endPosesAndLaserScans(std::vector<gtsam::NonlinearFactorGraph>* new_factors, std::vector<gtsam::Values>* new_values,
                             std::vector<bool>* is_prior);
  void prefetchLaserScans(const std::vector<unsigned int>& worker_ids, const std::vector<LaserScan>& scans);
  ls_ctx* trackContext() const { return track_ctx_; }

  const ls_pg_stats& getLastSolveStats() const { return last_stats_; }

 private:
  gtsam::Values updateGraph(const gtsam::NonlinearFactorGraph& factors, const gtsam::Values& values,
                            const std::vector<uint64_t>& remove, std::vector<uint64_t>* new_indices);
  unsigned int n_laser_slam_workers_ = 0u;
  mutable std::recursive_mutex full_class_mutex_;
  // a step between beginPosesAndLaserScans and endPosesAndLaserScans
  struct PendingStep {
    bool open = false, inflight = false;
    std::vector<unsigned int> worker_ids;
    std::vector<LaserTrack::PendingIcp> pending;
    std::vector<size_t> active;
    std::vector<float> T_outs;
    std::vector<ls_icp_stats> stats;
    std::vector<int> statuses;
  } step_;
  std::vector<std::shared_ptr<LaserTrack> > laser_tracks_;
  ls_pg* graph_ = nullptr;
  // device side of the lidar odometry, shared by all tracks so that one launch can serve them all
  ls_ctx* track_ctx_ = nullptr;
  ls_map* track_ring_ = nullptr;
  int track_ring_capacity_ = 0, track_ring_max_pts_ = 0;
  ls_ctx* icp_ctx_ = nullptr;  // loop-closure ICP (sub-map <-> sub-map)
  ls_icp_params icp_params_;
  gtsam::noiseModel::Base::shared_ptr loop_closure_noise_model_, first_association_noise_model_;
  std::unordered_map<unsigned int, size_t> factor_indices_to_remove_;
  std::vector<std::vector<unsigned int> > linked_workers_;
  EstimatorParams params_;
  ls_pg_stats last_stats_;
};

}  // namespace laser_slam

#endif  // LASER_SLAM_INCREMENTAL_ESTIMATOR_HPP_
