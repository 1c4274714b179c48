// C hooks over the C++ host layer so that tests (ctypes) can drive LaserTrack / IncrementalEstimator exactly the
// way the ROS worker does (reference laser_slam_ros/src/laser_slam_worker.cpp:124-173): processPoseAndLaserScan,
// then registerPrior or estimate, then updateFromGTSAMValues.  Not part of the drop-in boundary.
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "laser_slam/incremental_estimator.hpp"
#include "laser_slam/velodyne_assembler.hpp"

using namespace laser_slam;

namespace {
struct Handle {
  std::unique_ptr<IncrementalEstimator> est;
  std::string err;
  std::vector<unsigned int> step_ids;   // the step between lsh_begin_batch and lsh_end_batch
  std::vector<int64_t> step_times;
};
LaserScan make_scan(const float* feat4, const float* normals3, int n, int64_t time_ns, bool view) {
  LaserScan s;
  s.scan = view ? DataPoints::viewOfArrays(feat4, normals3, (size_t)n) : DataPoints::fromArrays(feat4, normals3, (size_t)n);
  s.time_ns = time_ns;
  return s;
}
template <typename F>
int guarded(Handle* h, F f) {
  try {
    return f();
  } catch (const laser_slam::PointMatcher::ConvergenceError& e) {
    h->err = e.what();
    return LS_ERR_CONVERGENCE;
  } catch (const std::exception& e) {
    h->err = e.what();
    return LS_ERR_STATE;
  }
}
}  // namespace

extern "C" {

void* lsh_create(int n_workers, int nscan_in_sub_map, int use_icp_factors, int use_odom_factors, int robust_icp, int device,
                 int do_icp_step_on_loop_closures, int loop_closures_sub_maps_radius, const char* icp_yaml_path, char* err, int errlen) {
  try {
    EstimatorParams p;
    p.laser_track_params.nscan_in_sub_map = nscan_in_sub_map;
    p.laser_track_params.use_icp_factors = use_icp_factors != 0;
    p.laser_track_params.use_odom_factors = use_odom_factors != 0;
    p.laser_track_params.add_m_estimator_on_icp = robust_icp != 0;
    p.laser_track_params.cuda_device = device;
    p.laser_track_params.icp_configuration_file = icp_yaml_path ? icp_yaml_path : "";
    p.do_icp_step_on_loop_closures = do_icp_step_on_loop_closures != 0;
    p.loop_closures_sub_maps_radius = loop_closures_sub_maps_radius;
    Handle* h = new Handle();
    h->est.reset(new IncrementalEstimator(p, (unsigned int)n_workers));
    return h;
  } catch (const std::exception& e) {
    if (err && errlen > 0) std::strncpy(err, e.what(), (size_t)errlen - 1), err[errlen - 1] = 0;
    return nullptr;
  }
}

void lsh_destroy(void* hv) { delete static_cast<Handle*>(hv); }
const char* lsh_last_error(void* hv) { return static_cast<Handle*>(hv)->err.c_str(); }

// One scan callback.  pose7 = odometry pose T_w (qw,qx,qy,qz,tx,ty,tz).  out_icp7 (may be NULL) receives the ICP
// T_a_b of this step (identity for the first scan); out_stats (may be NULL) the device-side ICP statistics.
int lsh_step(void* hv, int worker, int64_t time_ns, const double* pose7, const float* feat4, const float* normals3, int n,
             double* out_icp7, ls_icp_stats* out_stats) {
  Handle* h = static_cast<Handle*>(hv);
  return guarded(h, [&]() {
    std::shared_ptr<LaserTrack> track = h->est->getLaserTrack((unsigned int)worker);
    Pose pose;
    pose.T_w = SE3::fromArray7(pose7);
    pose.time_ns = time_ns;
    LaserScan scan;
    scan.scan = DataPoints::fromArrays(feat4, normals3, (size_t)n);
    scan.time_ns = time_ns;
    gtsam::NonlinearFactorGraph new_factors;
    gtsam::Values new_values;
    bool is_prior = false;
    track->processPoseAndLaserScan(pose, scan, &new_factors, &new_values, &is_prior);
    gtsam::Values result = is_prior ? h->est->registerPrior(new_factors, new_values, (unsigned int)worker)
                                    : h->est->estimate(new_factors, new_values, time_ns);
    track->updateFromGTSAMValues(result);
    if (out_icp7) {
      SE3 T;
      if (!track->getIcpTransformations().empty() && !is_prior) T = track->getIcpTransformations().back().T_a_b;
      T.toArray7(out_icp7);
    }
    if (out_stats) *out_stats = track->getLastIcpStats();
    return LS_OK;
  });
}

// The scan callbacks of `n_workers` workers at once (IncrementalEstimator::processPosesAndLaserScans: one batched
// launch for all registrations), each followed by registerPrior / estimate and updateFromGTSAMValues exactly as
// lsh_step does.  pose7: 7 doubles per worker; feat4 / normals3: per-worker host pointers; n: points per worker.
// out_icp7 (may be NULL): 7 doubles per worker.  with_estimator == 0 skips the pose-graph update (odometry only).
int lsh_step_batch(void* hv, int n_workers, const int* workers, const int64_t* times_ns, const double* pose7, const float* const* feat4,
                   const float* const* normals3, const int* n, int with_estimator, double* out_icp7, ls_icp_stats* out_stats) {
  // with_estimator bit 1 (value 2): the DataPoints are VIEWS of the caller's arrays (DataPoints::viewOfArrays), which
  // must then stay valid for the life of the estimator -- the tracks keep every scan.  Otherwise they are copies.
  const bool views = (with_estimator & 2) != 0;
  with_estimator &= 1;
  Handle* h = static_cast<Handle*>(hv);
  return guarded(h, [&]() {
    std::vector<unsigned int> ids;
    std::vector<Pose> poses((size_t)n_workers);
    std::vector<LaserScan> scans((size_t)n_workers);
    for (int i = 0; i < n_workers; ++i) {
      ids.push_back((unsigned int)workers[i]);
      poses[i].T_w = SE3::fromArray7(pose7 + 7 * (size_t)i);
      poses[i].time_ns = times_ns[i];
      scans[i].scan = views ? DataPoints::viewOfArrays(feat4[i], normals3[i], (size_t)n[i]) : DataPoints::fromArrays(feat4[i], normals3[i], (size_t)n[i]);
      scans[i].time_ns = times_ns[i];
    }
    std::vector<gtsam::NonlinearFactorGraph> nf;
    std::vector<gtsam::Values> nv;
    std::vector<bool> prior;
    h->est->processPosesAndLaserScans(ids, poses, scans, &nf, &nv, &prior);
    for (int i = 0; i < n_workers; ++i) {
      std::shared_ptr<LaserTrack> track = h->est->getLaserTrack(ids[i]);
      if (with_estimator) {
        gtsam::Values result = prior[i] ? h->est->registerPrior(nf[i], nv[i], ids[i]) : h->est->estimate(nf[i], nv[i], times_ns[i]);
        track->updateFromGTSAMValues(result);
      }
      if (out_icp7) {
        SE3 T;
        if (!track->getIcpTransformations().empty() && !prior[i]) T = track->getIcpTransformations().back().T_a_b;
        T.toArray7(out_icp7 + 7 * (size_t)i);
      }
      if (out_stats) out_stats[i] = track->getLastIcpStats();
    }
    return LS_OK;
  });
}

// lsh_step_batch in two halves (IncrementalEstimator::beginPosesAndLaserScans / endPosesAndLaserScans) with the prefetch
// hint in between.  flags bit 1 (value 2): DataPoints are views of the caller's arrays.
int lsh_begin_batch(void* hv, int n_workers, const int* workers, const int64_t* times_ns, const double* pose7, const float* const* feat4,
                    const float* const* normals3, const int* n, int flags) {
  Handle* h = static_cast<Handle*>(hv);
  return guarded(h, [&]() {
    std::vector<Pose> poses((size_t)n_workers);
    std::vector<LaserScan> scans((size_t)n_workers);
    h->step_ids.clear();
    h->step_times.assign(times_ns, times_ns + n_workers);
    for (int i = 0; i < n_workers; ++i) {
      h->step_ids.push_back((unsigned int)workers[i]);
      poses[i].T_w = SE3::fromArray7(pose7 + 7 * (size_t)i);
      poses[i].time_ns = times_ns[i];
      scans[i] = make_scan(feat4[i], normals3[i], n[i], times_ns[i], (flags & 2) != 0);
    }
    h->est->beginPosesAndLaserScans(h->step_ids, poses, scans);
    return LS_OK;
  });
}

int lsh_prefetch(void* hv, int n_workers, const int* workers, const int64_t* times_ns, const float* const* feat4,
                 const float* const* normals3, const int* n, int flags) {
  Handle* h = static_cast<Handle*>(hv);
  return guarded(h, [&]() {
    std::vector<unsigned int> ids;
    std::vector<LaserScan> scans((size_t)n_workers);
    for (int i = 0; i < n_workers; ++i) {
      ids.push_back((unsigned int)workers[i]);
      scans[i] = make_scan(feat4[i], normals3[i], n[i], times_ns[i], (flags & 2) != 0);
    }
    h->est->prefetchLaserScans(ids, scans);
    return LS_OK;
  });
}

int lsh_end_batch(void* hv, int with_estimator, double* out_icp7, ls_icp_stats* out_stats) {
  Handle* h = static_cast<Handle*>(hv);
  return guarded(h, [&]() {
    std::vector<gtsam::NonlinearFactorGraph> nf;
    std::vector<gtsam::Values> nv;
    std::vector<bool> prior;
    h->est->endPosesAndLaserScans(&nf, &nv, &prior);
    for (size_t i = 0; i < h->step_ids.size(); ++i) {
      std::shared_ptr<LaserTrack> track = h->est->getLaserTrack(h->step_ids[i]);
      if (with_estimator & 1) {
        gtsam::Values result = prior[i] ? h->est->registerPrior(nf[i], nv[i], h->step_ids[i]) : h->est->estimate(nf[i], nv[i], h->step_times[i]);
        track->updateFromGTSAMValues(result);
      }
      if (out_icp7) {
        SE3 T;
        if (!track->getIcpTransformations().empty() && !prior[i]) T = track->getIcpTransformations().back().T_a_b;
        T.toArray7(out_icp7 + 7 * i);
      }
      if (out_stats) out_stats[i] = track->getLastIcpStats();
    }
    return LS_OK;
  });
}

int lsh_loop_closure(void* hv, int track_a, int64_t time_a, int track_b, int64_t time_b, const double* w_T_a_b7) {
  Handle* h = static_cast<Handle*>(hv);
  return guarded(h, [&]() {
    RelativePose lc;
    lc.T_a_b = SE3::fromArray7(w_T_a_b7);
    lc.time_a_ns = time_a;
    lc.time_b_ns = time_b;
    lc.track_id_a = (unsigned int)track_a;
    lc.track_id_b = (unsigned int)track_b;
    h->est->processLoopClosure(lc);
    return LS_OK;
  });
}

// Trajectory of one track: times and poses (7 doubles each); returns the number of nodes (<= cap written).
int lsh_trajectory(void* hv, int worker, int64_t* times, double* poses7, int cap) {
  Handle* h = static_cast<Handle*>(hv);
  return guarded(h, [&]() {
    Trajectory traj;
    h->est->getLaserTrack((unsigned int)worker)->getTrajectory(&traj);
    int i = 0;
    for (const auto& kv : traj) {
      if (i < cap) {
        if (times) times[i] = kv.first;
        if (poses7) kv.second.toArray7(poses7 + 7 * (size_t)i);
      }
      ++i;
    }
    return i;
  });
}

int lsh_num_scans(void* hv, int worker) {
  Handle* h = static_cast<Handle*>(hv);
  return guarded(h, [&]() { return (int)h->est->getLaserTrack((unsigned int)worker)->getNumScans(); });
}

// LaserTrack::buildSubMapAroundTime; returns the number of points (features4 / normals3 sized by the caller).
int lsh_build_submap(void* hv, int worker, int64_t time_ns, int radius, float* features4, float* normals3, int cap_points) {
  Handle* h = static_cast<Handle*>(hv);
  return guarded(h, [&]() {
    DataPoints sub;
    h->est->getLaserTrack((unsigned int)worker)->buildSubMapAroundTime(time_ns, (unsigned int)radius, &sub);
    const int m = (int)sub.getNbPoints();
    if (m <= cap_points) {
      std::memcpy(features4, sub.features.data(), sizeof(float) * 4 * (size_t)m);
      if (normals3) std::memcpy(normals3, sub.descriptors.data(), sizeof(float) * 3 * (size_t)m);
    }
    return m;
  });
}


// ---- laser_slam::VelodyneAssembler (include/laser_slam/velodyne_assembler.hpp) for the tests
void* lsh_assembler_create(const float* T_sensor_base16, int naive, int device) {
  VelodyneAssembler::Matrix4 T;
  if (T_sensor_base16) std::memcpy(T.data(), T_sensor_base16, 16 * sizeof(float));
  return new VelodyneAssembler(T, naive != 0, device);
}
void lsh_assembler_destroy(void* a) { delete static_cast<VelodyneAssembler*>(a); }
// returns 1 when a revolution was completed by this packet (out4 then holds *m_out points, at most cap), 0 if not, < 0 on error
int lsh_assembler_add_packet(void* av, const float* pts4, int n, const float* T_fixed_base16, int64_t stamp_ns, float* out4, int cap,
                             int* m_out, int64_t* stamp_out) {
  try {
    VelodyneAssembler* a = static_cast<VelodyneAssembler*>(av);
    VelodyneAssembler::Matrix4 T;
    std::memcpy(T.data(), T_fixed_base16, 16 * sizeof(float));
    DataPoints in = DataPoints::viewOfArrays(pts4, NULL, (size_t)n), rev;
    Time st = 0;
    if (!a->addPacket(in, T, (Time)stamp_ns, &rev, &st)) return 0;
    const int m = (int)rev.getNbPoints();
    if (m > cap) return LS_ERR_ARG;
    std::memcpy(out4, static_cast<const DataPoints&>(rev).features.data(), sizeof(float) * 4 * (size_t)m);
    *m_out = m;
    if (stamp_out) *stamp_out = (int64_t)st;
    return 1;
  } catch (const std::exception&) {
    return LS_ERR_STATE;
  }
}
}  // extern "C"
