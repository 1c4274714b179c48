// LaserTrack over the B200 C ABI.  Control flow follows reference laser_slam/src/laser_track.cpp (cited per
// function); the heavy steps call include/ls_b200.h instead of libpointmatcher:
//   laser_scans_ copies + RigidTransformation::compute + concatenate  ->  ls_map_push_scan / device assembly
//   icp_.compute                                                       ->  ls_icp_register_submap
// Scans must arrive with a "normals" descriptor: the reference computes it in its input / reference filters
// (icp_default.yaml:5-7), which are upstream of this path.
#include "laser_slam/laser_track.hpp"

#include <atomic>
#include <chrono>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace laser_slam {

namespace {

#define LS_CHECK(cond, msg)                                             \
  do {                                                                  \
    if (!(cond)) throw std::logic_error(std::string("CHECK failed: ") + (msg)); \
  } while (0)

std::atomic<uint64_t> g_key_counter{1};

std::string readFile(const std::string& path) {
  std::ifstream ifs(path.c_str());
  if (!ifs.good()) return std::string();
  std::stringstream ss;
  ss << ifs.rdbuf();
  return ss.str();
}

PointMatcher::TransformationParameters toFloatMatrix(const SE3& T) {
  return PointMatcher::TransformationParameters::cast(T.getTransformationMatrix());
}

void throwOnError(ls_ctx* ctx, int rc, const char* what) {
  if (rc < 0) throw std::runtime_error(std::string(what) + ": " + ls_b200_last_error(ctx));
}

}  // namespace

// reference laser_track.cpp:10-65
LaserTrack::LaserTrack(const LaserTrackParams& parameters, unsigned int laser_track_id)
    : laser_track_id_(laser_track_id), params_(parameters) {
  // ICP chain: the YAML the reference hands to icp_.loadFromYaml, or libpointmatcher's setDefault() values
  //  when the file cannot be opened (reference :14-21).
  ls_icp_default_params(&icp_params_);
  const std::string yaml = readFile(params_.icp_configuration_file);
  if (!yaml.empty()) {
    if (ls_icp_params_from_yaml(yaml.c_str(), &icp_params_) != LS_OK)
      throw std::runtime_error("unsupported ICP chain in " + params_.icp_configuration_file);
  } else {
    icp_params_.trim_ratio = 0.85f;
    icp_params_.min_diff_rot = 0.001f;
    icp_params_.min_diff_trans = 0.001f;
    icp_params_.smooth_length = 3;
  }
  // reference :24-30 is fatal when the input-filter file cannot be opened; an empty name means "no filters".
  if (!params_.icp_input_filters_file.empty() && readFile(params_.icp_input_filters_file).empty())
    throw std::runtime_error("Could not open ICP input filters configuration file.");
  // noise models (reference :36-64)
  odometry_noise_model_ = gtsam::NoiseModel{params_.odometry_noise_model, params_.add_m_estimator_on_odom};
  icp_noise_model_ = gtsam::NoiseModel{params_.icp_noise_model, params_.add_m_estimator_on_icp};
  prior_noise_model_ = gtsam::NoiseModel{{{1e-7, 1e-7, 1e-7, 1e-7, 1e-7, 1e-7}}, false};
  std::memset(&last_icp_stats_, 0, sizeof(last_icp_stats_));
  const int rc = ls_b200_init(params_.cuda_device, &ctx_);
  if (rc != LS_OK) throw std::runtime_error("ls_b200_init failed: no usable CUDA device (no CPU fallback)");
}

LaserTrack::~LaserTrack() {
  if (map_) ls_map_destroy(map_);
  if (ctx_) ls_b200_destroy(ctx_);
}

// reference :67-73
void LaserTrack::processPose(const Pose& pose) {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  pose_measurements_.push_back(pose);
}

// reference :75-120 (older twin of processPoseAndLaserScan without factor output)
void LaserTrack::processLaserScan(const LaserScan& in_scan) {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  const Pose pose = findPose(in_scan.time_ns);  // registered earlier through processPose
  for (auto it = pose_measurements_.begin(); it != pose_measurements_.end(); ++it)
    if (it->time_ns == in_scan.time_ns) { pose_measurements_.erase(it); break; }  // re-appended just below
  processPoseAndLaserScan(pose, in_scan, NULL, NULL, NULL);
}

// reference :122-231
void LaserTrack::processPoseAndLaserScan(const Pose& pose, const LaserScan& in_scan, gtsam::NonlinearFactorGraph* newFactors,
                                         gtsam::Values* newValues, bool* is_prior) {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  const auto t_start = std::chrono::steady_clock::now();
  if (newFactors != NULL) LS_CHECK(newFactors->empty(), "newFactors must be empty on entry");
  if (newValues != NULL) newValues->clear();
  LS_CHECK(in_scan.scan.descriptorExists("normals"), "scans must carry a 'normals' descriptor");

  LaserScan scan = in_scan;  // the reference copies too (:143); filters would run on the copy
  pose_measurements_.push_back(pose);

  if (trajectory_.empty()) {
    scan.key = extendTrajectory(scan.time_ns, findPose(scan.time_ns).T_w);
    findPose(scan.time_ns).key = scan.key;
    laser_scans_.push_back(scan);
    if (newFactors != NULL) {
      Pose prior_pose = pose;
      prior_pose.key = scan.key;
      prior_pose.time_ns = scan.time_ns;
      if (params_.force_priors)  // reference :165-169
        prior_pose.T_w = SE3(SO3(1.0, 0.0, 0.0, 0.0), SE3::Position{0.0, kDistanceBetweenPriorPoses_m * laser_track_id_, 0.0});
      newFactors->push_back(makeMeasurementFactor(prior_pose, prior_noise_model_));
    }
    if (is_prior != NULL) *is_prior = true;
  } else {
    const Time t_last = trajectory_.rbegin()->first;
    const SE3 last_pose_measurement = findPose(t_last).T_w;
    const SE3 new_pose_measurement = findPose(scan.time_ns).T_w;
    RelativePose relative_measurement;
    relative_measurement.T_a_b = last_pose_measurement.inverse() * new_pose_measurement;
    relative_measurement.time_a_ns = t_last;
    relative_measurement.key_a = findPose(t_last).key;
    relative_measurement.time_b_ns = scan.time_ns;
    relative_measurement.track_id_a = relative_measurement.track_id_b = laser_track_id_;
    // extend the trajectory by odometry (reference :192)
    scan.key = extendTrajectory(scan.time_ns, trajectory_.rbegin()->second.value * relative_measurement.T_a_b);
    findPose(scan.time_ns).key = scan.key;
    laser_scans_.push_back(scan);
    relative_measurement.key_b = scan.key;
    odometry_measurements_.push_back(relative_measurement);
    if (params_.use_icp_factors) computeICPTransformations();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
    scan_matching_times_.emplace(scan.time_ns, ms);
    if (newFactors != NULL) {
      if (params_.use_odom_factors) newFactors->push_back(makeRelativeMeasurementFactor(relative_measurement, odometry_noise_model_));
      if (params_.use_icp_factors && !icp_transformations_.empty())
        newFactors->push_back(makeRelativeMeasurementFactor(icp_transformations_.back(), icp_noise_model_));
    }
    if (is_prior != NULL) *is_prior = false;
  }
  if (newValues != NULL) newValues->insert(scan.key, pose.T_w);  // reference :228-230
}

void LaserTrack::getLastPointCloud(DataPoints* out_point_cloud) const {  // stub in the reference too (:233-237)
  LS_CHECK(out_point_cloud != NULL, "null output");
}
void LaserTrack::getPointCloudOfTimeInterval(const std::pair<Time, Time>&, DataPoints* out_point_cloud) const {
  LS_CHECK(out_point_cloud != NULL, "null output");
  *out_point_cloud = DataPoints();  // reference :239-245
}

// reference :247-266
void LaserTrack::getLocalCloudInWorldFrame(const Time& timestamp_ns, DataPoints* out_point_cloud) const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  LS_CHECK(out_point_cloud != NULL, "null output");
  const size_t idx = scanIndexAtTime(timestamp_ns);
  PointMatcher::TransformationParameters T = toFloatMatrix(evaluate(timestamp_ns));
  correctTransformationMatrix(&T);
  assembleSubMap({idx}, {T}, out_point_cloud);
}

void LaserTrack::getTrajectory(Trajectory* trajectory) const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  LS_CHECK(trajectory != NULL, "null output");
  trajectory->clear();
  for (const auto& kv : trajectory_) trajectory->emplace(kv.first, kv.second.value);
}
const std::vector<LaserScan>& LaserTrack::getLaserScans() const { return laser_scans_; }
void LaserTrack::getCovariances(std::vector<Covariance>* out) const {
  LS_CHECK(out != NULL, "null output");
  *out = covariances_;
}
Pose LaserTrack::getCurrentPose() const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  Pose p;
  if (!trajectory_.empty()) {
    p.time_ns = trajectory_.rbegin()->first;
    p.T_w = trajectory_.rbegin()->second.value;
    p.key = trajectory_.rbegin()->second.key;
  }
  return p;
}
Pose LaserTrack::getPreviousPose() const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  Pose p;
  if (trajectory_.size() > 1u) {
    auto it = trajectory_.rbegin();
    ++it;
    p.time_ns = it->first;
    p.T_w = it->second.value;
    p.key = it->second.key;
  }
  return p;
}
void LaserTrack::getOdometryTrajectory(Trajectory* trajectory) const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  LS_CHECK(trajectory != NULL, "null output");
  trajectory->clear();
  for (const auto& pose : pose_measurements_) trajectory->emplace(pose.time_ns, pose.T_w);
}
Time LaserTrack::getMinTime() const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  LS_CHECK(!trajectory_.empty(), "empty trajectory");
  return trajectory_.begin()->first;
}
Time LaserTrack::getMaxTime() const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  LS_CHECK(!trajectory_.empty(), "empty trajectory");
  return trajectory_.rbegin()->first;
}
void LaserTrack::getLaserScansTimes(std::vector<curves::Time>* out_times_ns) const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  LS_CHECK(out_times_ns != NULL, "null output");
  out_times_ns->clear();
  for (const auto& s : laser_scans_) out_times_ns->push_back(s.time_ns);
}
size_t LaserTrack::getNumScans() const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  return laser_scans_.size();
}

// reference :339-344 (curves::DiscreteSE3Curve::addPriorFactors: prior at the node value)
void LaserTrack::appendPriorFactors(const Time& prior_time_ns, gtsam::NonlinearFactorGraph* graph) const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  LS_CHECK(graph != NULL, "null graph");
  auto it = trajectory_.find(prior_time_ns);
  LS_CHECK(it != trajectory_.end(), "no trajectory node at the prior time");
  Pose p;
  p.T_w = it->second.value;
  p.time_ns = prior_time_ns;
  p.key = it->second.key;
  graph->push_back(makeMeasurementFactor(p, prior_noise_model_));
}
// reference :346-361
void LaserTrack::appendOdometryFactors(const Time& tmin, const Time& tmax, const gtsam::NoiseModel& noise,
                                       gtsam::NonlinearFactorGraph* graph) const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  LS_CHECK(graph != NULL, "null graph");
  for (const auto& m : odometry_measurements_)
    if (m.time_a_ns >= tmin && m.time_b_ns <= tmax) graph->push_back(makeRelativeMeasurementFactor(m, noise));
}
namespace {
template <typename MakeFn>
void appendWindowed(const RelativePoseVector& v, const Time& tmin, const Time& tmax, gtsam::NonlinearFactorGraph* graph, MakeFn make) {
  for (const auto& m : v) {
    if (m.time_b_ns >= tmin && m.time_b_ns <= tmax) {  // second node inside the window
      const bool a_inside = m.time_a_ns >= tmin && m.time_a_ns <= tmax;
      graph->push_back(make(m, !a_inside));  // first node outside -> frozen (fix_first_node)
    }
  }
}
}  // namespace
// reference :363-384
void LaserTrack::appendICPFactors(const Time& tmin, const Time& tmax, const gtsam::NoiseModel& noise,
                                  gtsam::NonlinearFactorGraph* graph) const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  LS_CHECK(graph != NULL, "null graph");
  appendWindowed(icp_transformations_, tmin, tmax, graph,
                 [&](const RelativePose& m, bool fix) { return makeRelativeMeasurementFactor(m, noise, fix); });
}
// reference :386-409
void LaserTrack::appendLoopClosureFactors(const Time& tmin, const Time& tmax, const gtsam::NoiseModel& noise,
                                          gtsam::NonlinearFactorGraph* graph) const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  LS_CHECK(graph != NULL, "null graph");
  appendWindowed(loop_closures_, tmin, tmax, graph,
                 [&](const RelativePose& m, bool fix) { return makeRelativeMeasurementFactor(m, noise, fix); });
}

// reference :411-419
void LaserTrack::initializeGTSAMValues(const std::vector<Key>& keys, gtsam::Values* values) const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  LS_CHECK(values != NULL, "null values");
  for (const auto& kv : trajectory_)
    for (Key k : keys)
      if (kv.second.key == k && !values->exists(k)) values->insert(k, kv.second.value);
}
void LaserTrack::updateFromGTSAMValues(const gtsam::Values& values) {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  for (auto& kv : trajectory_)
    if (values.exists(kv.second.key)) kv.second.value = values.at(kv.second.key);
}

// reference :431-451
ls_factor LaserTrack::makeRelativeMeasurementFactor(const RelativePose& m, const gtsam::NoiseModel& noise, bool fix_first_node) const {
  ls_factor f;
  std::memset(&f, 0, sizeof(f));
  f.type = LS_FACTOR_BETWEEN;
  f.robust = noise.cauchy ? 1 : 0;
  f.fix_a = fix_first_node ? 1 : 0;
  f.key_a = m.key_a;
  f.key_b = m.key_b;
  m.T_a_b.toArray7(f.meas);
  for (int i = 0; i < 6; ++i) f.sigma[i] = noise.sigmas[i];
  SE3 fixed;  // constant T_w_a when the first node is frozen (reference :440-444)
  if (fix_first_node) fixed = evaluate(m.time_a_ns);
  fixed.toArray7(f.fixed_a);
  return f;
}
// reference :453-458
ls_factor LaserTrack::makeMeasurementFactor(const Pose& pose_measurement, const gtsam::NoiseModel& noise) const {
  ls_factor f;
  std::memset(&f, 0, sizeof(f));
  f.type = LS_FACTOR_PRIOR;
  f.robust = noise.cauchy ? 1 : 0;
  f.key_a = f.key_b = getValueKey(pose_measurement.time_ns);
  pose_measurement.T_w.toArray7(f.meas);
  for (int i = 0; i < 6; ++i) f.sigma[i] = noise.sigmas[i];
  SE3().toArray7(f.fixed_a);
  return f;
}

// reference :460-464
void LaserTrack::computeICPTransformations() {
  if (getNumScans() > 1u) localScanToSubMap();
}

uint64_t LaserTrack::residentScan(size_t index) const {
  auto it = resident_.find(index);
  if (it != resident_.end() && ls_map_scan_size(map_, it->second) >= 0) return it->second;
  const DataPoints& c = laser_scans_[index].scan;
  const int off = c.descriptorOffset("normals");
  LS_CHECK(off >= 0, "scan without normals");
  uint64_t id = 0;
  const int rc = ls_map_push_scan(map_, c.features.data(), c.descriptors.data() + off, (int)c.descriptorDim, (int)c.getNbPoints(), &id);
  throwOnError(ctx_, rc, "ls_map_push_scan");
  resident_[index] = id;
  return id;
}

// reference :466-519
void LaserTrack::localScanToSubMap() {
  const size_t n = laser_scans_.size();
  const LaserScan& last_scan = laser_scans_[n - 1u];
  RelativePose icp_transformation;
  icp_transformation.time_b_ns = last_scan.time_ns;
  icp_transformation.time_a_ns = laser_scans_[n - 2u].time_ns;
  icp_transformation.track_id_a = icp_transformation.track_id_b = laser_track_id_;

  // device ring large enough for the sub-map + the reading; (re)created when a larger scan shows up
  size_t max_pts = 0;
  for (size_t i = (n > 16 ? n - 16 : 0); i < n; ++i) max_pts = std::max(max_pts, laser_scans_[i].scan.getNbPoints());
  const int want_cap = std::max(8, params_.nscan_in_sub_map + 3);
  if (!map_ || (int)max_pts > map_max_pts_ || want_cap > map_capacity_) {
    if (map_) ls_map_destroy(map_);
    map_ = nullptr;
    resident_.clear();
    map_max_pts_ = (int)(max_pts + max_pts / 4 + 1024);
    map_capacity_ = want_cap;
    throwOnError(ctx_, ls_map_create(ctx_, map_capacity_, map_max_pts_, &map_), "ls_map_create");
  }

  // the last (nscan_in_sub_map - 1) scans expressed in the frame of the second-last scan (reference :474-486)
  const SE3 T_w_to_second_last_scan = evaluate(laser_scans_[n - 2u].time_ns);
  std::vector<size_t> part_index{n - 2u};
  std::vector<PointMatcher::TransformationParameters> part_T(1);  // identity: scan n-2 verbatim (:476)
  const size_t n_prev = std::min(n - 2u, size_t(params_.nscan_in_sub_map > 0 ? params_.nscan_in_sub_map - 1 : 0));
  for (size_t i = 0u; i < n_prev; ++i) {
    const size_t idx = n - 3u - i;
    PointMatcher::TransformationParameters T = toFloatMatrix(T_w_to_second_last_scan.inverse() * evaluate(laser_scans_[idx].time_ns));
    correctTransformationMatrix(&T);
    part_index.push_back(idx);
    part_T.push_back(T);
  }
  // initial guess from the trajectory (reference :488-491)
  const SE3 initial_guess = evaluate(icp_transformation.time_a_ns).inverse() * evaluate(icp_transformation.time_b_ns);
  const PointMatcher::TransformationParameters T0 = toFloatMatrix(initial_guess);
  PointMatcher::TransformationParameters icp_solution = T0;

  // upload what is not resident yet (normally only the newest scan), reading last so it cannot evict a part
  std::vector<uint64_t> part_ids;
  for (size_t idx : part_index) part_ids.push_back(residentScan(idx));
  const uint64_t reading_id = residentScan(n - 1u);
  for (size_t k = 0; k < part_index.size(); ++k) part_ids[k] = residentScan(part_index[k]);
  std::vector<float> T_flat;
  for (const auto& T : part_T) T_flat.insert(T_flat.end(), T.data(), T.data() + 16);

  const int rc = ls_icp_register_submap(ctx_, &icp_params_, map_, reading_id, (int)part_ids.size(), part_ids.data(), T_flat.data(),
                                        T0.data(), icp_solution.data(), &last_icp_stats_, NULL, NULL, NULL);
  if (rc == LS_ERR_CONVERGENCE) {
    icp_solution = T0;  // PointMatcher::ConvergenceError is swallowed: keep the initial guess (reference :495-502)
  } else {
    throwOnError(ctx_, rc, "ls_icp_register_submap");
  }
  icp_transformation.T_a_b = convertTransformationMatrixToSE3(icp_solution);
  icp_transformation.key_a = findPose(icp_transformation.time_a_ns).key;
  icp_transformation.key_b = findPose(icp_transformation.time_b_ns).key;
  icp_transformations_.push_back(icp_transformation);
}

// reference :521-555 (reverse linear scan for an exact time stamp)
const Pose& LaserTrack::findPose(const Time& timestamp_ns) const {
  LS_CHECK(!pose_measurements_.empty(), "Cannot register the scan as no pose was registered.");
  for (auto it = pose_measurements_.rbegin(); it != pose_measurements_.rend(); ++it)
    if (it->time_ns == timestamp_ns) return *it;
  throw std::logic_error("CHECK failed: The requested time does not exist in the pose measurements.");
}
Pose& LaserTrack::findPose(const Time& timestamp_ns) {
  return const_cast<Pose&>(static_cast<const LaserTrack*>(this)->findPose(timestamp_ns));
}
// reference :557-571
Pose LaserTrack::findNearestPose(const Time& timestamp_ns) const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  Pose pose;
  pose.time_ns = timestamp_ns;
  pose.T_w = evaluate(timestamp_ns);
  pose.key = Key();
  return pose;
}
// reference :573-582; keys are unique across tracks (mincurves' process-wide key generator) and carry the track id
Key LaserTrack::extendTrajectory(const Time& timestamp_ns, const SE3& value) {
  LS_CHECK(trajectory_.empty() || timestamp_ns > trajectory_.rbegin()->first, "trajectory must be extended forward in time");
  const Key key = ((Key)laser_track_id_ << 48) | (g_key_counter.fetch_add(1) & 0xFFFFFFFFFFFFull);
  trajectory_.emplace(timestamp_ns, Node{value, key});
  return key;
}
size_t LaserTrack::scanIndexAtTime(const curves::Time& time_ns) const {  // reference :584-600
  for (size_t i = 0; i < laser_scans_.size(); ++i)
    if (laser_scans_[i].time_ns == time_ns) return i;
  throw std::logic_error("CHECK failed: Could not find the scan.");
}
Key LaserTrack::getValueKey(const curves::Time& time_ns) const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  auto it = trajectory_.find(time_ns);
  LS_CHECK(it != trajectory_.end(), "no trajectory node at that time");
  return it->second.key;
}
SE3 LaserTrack::evaluate(const curves::Time& time_ns) const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  auto it = trajectory_.find(time_ns);
  LS_CHECK(it != trajectory_.end(), "no trajectory node at that time (only exact node times are evaluated)");
  return it->second.value;
}
void LaserTrack::getScanMatchingTimes(std::map<Time, double>* out) const {
  LS_CHECK(out != NULL, "null output");
  *out = scan_matching_times_;
}
void LaserTrack::saveTrajectory(const std::string& filename) const {  // curves::saveCurveTimesAndValues
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  std::ofstream out(filename.c_str());
  out.precision(17);
  for (const auto& kv : trajectory_) {
    double a[7];
    kv.second.value.toArray7(a);
    out << kv.first << "," << a[4] << "," << a[5] << "," << a[6] << "," << a[0] << "," << a[1] << "," << a[2] << "," << a[3] << "\n";
  }
}

void LaserTrack::assembleSubMap(const std::vector<size_t>& scan_indices,
                                const std::vector<PointMatcher::TransformationParameters>& Ts, DataPoints* out) const {
  LS_CHECK(!scan_indices.empty() && scan_indices.size() == Ts.size(), "bad sub-map description");
  // a private ring sized for this request (loop-closure sub-maps can be wider than the rolling window)
  size_t max_pts = 0, total = 0;
  for (size_t idx : scan_indices) {
    max_pts = std::max(max_pts, laser_scans_[idx].scan.getNbPoints());
    total += laser_scans_[idx].scan.getNbPoints();
  }
  ls_map* tmp = nullptr;
  throwOnError(ctx_, ls_map_create(ctx_, (int)std::max<size_t>(2, scan_indices.size()), (int)std::max<size_t>(1, max_pts), &tmp),
               "ls_map_create");
  std::vector<uint64_t> ids;
  std::vector<float> T_flat;
  try {
    for (size_t k = 0; k < scan_indices.size(); ++k) {
      const DataPoints& c = laser_scans_[scan_indices[k]].scan;
      const int off = c.descriptorOffset("normals");
      LS_CHECK(off >= 0, "scan without normals");
      uint64_t id = 0;
      throwOnError(ctx_, ls_map_push_scan(tmp, c.features.data(), c.descriptors.data() + off, (int)c.descriptorDim,
                                          (int)c.getNbPoints(), &id), "ls_map_push_scan");
      ids.push_back(id);
      T_flat.insert(T_flat.end(), Ts[k].data(), Ts[k].data() + 16);
    }
    std::vector<float> feat(4 * std::max<size_t>(1, total)), nrm(3 * std::max<size_t>(1, total));
    int m = 0;
    throwOnError(ctx_, ls_map_assemble(ctx_, tmp, (int)ids.size(), ids.data(), T_flat.data(), feat.data(), nrm.data(), &m),
                 "ls_map_assemble");
    *out = DataPoints::fromArrays(feat.data(), nrm.data(), (size_t)m);
  } catch (...) {
    ls_map_destroy(tmp);
    throw;
  }
  ls_map_destroy(tmp);
}

// reference :602-651: the centre scan verbatim, then up to `radius` scans before it (decreasing time stamps) and after
// it (increasing), each re-expressed in the centre scan's frame
void LaserTrack::describeSubMapAroundTime(const curves::Time& time_ns, const unsigned int sub_maps_radius,
                                          std::vector<size_t>* scan_indices,
                                          std::vector<PointMatcher::TransformationParameters>* Ts) const {
  const SE3 T_w_a = evaluate(time_ns);
  const size_t centre = scanIndexAtTime(time_ns);
  scan_indices->assign(1, centre);
  Ts->assign(1, PointMatcher::TransformationParameters());
  auto add = [&](size_t i) {
    PointMatcher::TransformationParameters T = toFloatMatrix(T_w_a.inverse() * evaluate(laser_scans_[i].time_ns));
    correctTransformationMatrix(&T);
    scan_indices->push_back(i);
    Ts->push_back(T);
  };
  for (unsigned int i = 1; i <= sub_maps_radius && centre >= i; ++i) add(centre - i);
  for (unsigned int i = 1; i <= sub_maps_radius && centre + i < laser_scans_.size(); ++i) add(centre + i);
}

void LaserTrack::buildSubMapAroundTime(const curves::Time& time_ns, const unsigned int sub_maps_radius, DataPoints* submap_out) const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  LS_CHECK(submap_out != NULL, "null output");
  std::vector<size_t> idx;
  std::vector<PointMatcher::TransformationParameters> Ts;
  describeSubMapAroundTime(time_ns, sub_maps_radius, &idx, &Ts);
  assembleSubMap(idx, Ts, submap_out);
}

void LaserTrack::stageSubMapAroundTime(const curves::Time& time_ns, const unsigned int sub_maps_radius, ls_ctx* ctx,
                                       ls_map** ring_out, std::vector<uint64_t>* ids_out, std::vector<float>* T_parts_out) const {
  std::lock_guard<std::recursive_mutex> lock(full_laser_track_mutex_);
  LS_CHECK(ctx != NULL && ring_out != NULL && ids_out != NULL && T_parts_out != NULL, "null argument");
  std::vector<size_t> idx;
  std::vector<PointMatcher::TransformationParameters> Ts;
  describeSubMapAroundTime(time_ns, sub_maps_radius, &idx, &Ts);
  size_t max_pts = 1;
  for (size_t i : idx) max_pts = std::max(max_pts, laser_scans_[i].scan.getNbPoints());
  ls_map* ring = nullptr;
  throwOnError(ctx, ls_map_create(ctx, (int)std::max<size_t>(2, idx.size()), (int)max_pts, &ring), "ls_map_create");
  ids_out->clear();
  T_parts_out->clear();
  try {
    for (size_t k = 0; k < idx.size(); ++k) {
      const DataPoints& c = laser_scans_[idx[k]].scan;
      const int off = c.descriptorOffset("normals");
      LS_CHECK(off >= 0, "scan without normals");
      uint64_t id = 0;
      throwOnError(ctx, ls_map_push_scan(ring, c.features.data(), c.descriptors.data() + off, (int)c.descriptorDim,
                                         (int)c.getNbPoints(), &id), "ls_map_push_scan");
      ids_out->push_back(id);
      T_parts_out->insert(T_parts_out->end(), Ts[k].data(), Ts[k].data() + 16);
    }
  } catch (...) {
    ls_map_destroy(ring);
    throw;
  }
  *ring_out = ring;
}

}  // namespace laser_slam
