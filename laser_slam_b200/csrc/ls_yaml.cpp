// Host-side reader for the libpointmatcher ICP chain YAML the reference loads with
// icp_.loadFromYaml (reference laser_slam/src/laser_track.cpp:14-21,
// laser_slam/src/incremental_estimator.cpp:52-60).  Only the modules of
// laser_slam/configurations/icp_default.yaml are understood; yaml-cpp is not available, so this is a
// small indentation-insensitive "module name / key: value" scanner, enough for that file's grammar.
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>

#include "../../include/ls_b200.h"

namespace {

std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) ++a;
  while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

}  // namespace

// hash32: two rounds of a multiply-xorshift mixer over (index, salt); the oracle restates it (oracle/__init__.py keep_mask)
extern "C" int ls_keep_point(uint32_t index, uint32_t salt, float prob) {
  if (!(prob < 1.0f)) return 1;
  if (!(prob > 0.0f)) return 0;
  uint32_t h = index * 0x9E3779B1u + salt * 0x85EBCA77u + 0x165667B1u;
  h ^= h >> 15; h *= 0x2C1B3C6Du;
  h ^= h >> 12; h *= 0x297A2D39u;
  h ^= h >> 15;
  return (double)h < (double)prob * 4294967296.0 ? 1 : 0;
}

extern "C" int ls_icp_params_from_yaml(const char* yaml_text, ls_icp_params* p) {
  if (!yaml_text || !p) return LS_ERR_ARG;
  ls_icp_default_params(p);
  bool saw_counter = false, saw_diff = false, saw_trim = false, saw_checkers = false;
  std::string section, module;
  std::istringstream in(yaml_text);
  std::string raw;
  while (std::getline(in, raw)) {
    const size_t hash = raw.find('#');
    if (hash != std::string::npos) raw = raw.substr(0, hash);
    std::string line = trim(raw);
    if (line.empty()) continue;
    const bool top_level = !std::isspace((unsigned char)raw[0]) && raw[0] != '-';
    if (!line.empty() && line[0] == '-') line = trim(line.substr(1));
    std::string key = line, val;
    const size_t colon = line.find(':');
    if (colon != std::string::npos) {
      key = trim(line.substr(0, colon));
      val = trim(line.substr(colon + 1));
    }
    if (top_level) {
      section = key;
      module.clear();
      if (section == "transformationCheckers") saw_checkers = true;
      if (section == "errorMinimizer" && !val.empty()) module = val;
      if (section == "matcher" && !val.empty()) module = val;
      if (section == "errorMinimizer" && !val.empty() && val != "PointToPlaneErrorMinimizer") return LS_ERR_ARG;
      continue;
    }
    // module names end in a known suffix and carry no value (or an empty one)
    const bool is_module = val.empty() && (key.find("Matcher") != std::string::npos || key.find("Filter") != std::string::npos ||
                                           key.find("Minimizer") != std::string::npos ||
                                           key.find("Checker") != std::string::npos || key.find("Inspector") != std::string::npos ||
                                           key.find("Logger") != std::string::npos);
    if (is_module) {
      module = key;
      if (section == "matcher" && module != "KDTreeMatcher") return LS_ERR_ARG;
      if (section == "errorMinimizer" && module != "PointToPlaneErrorMinimizer") return LS_ERR_ARG;
      if (section == "outlierFilters") {
        if (module != "TrimmedDistOutlierFilter") return LS_ERR_ARG;
        saw_trim = true;
      }
      if (section == "readingDataPointsFilters" || section == "referenceDataPointsFilters") {
        ++p->unapplied_modules;  // reported, applied by the caller (ls_keep_point / ls_estimate_normals)
        if (section == "referenceDataPointsFilters" && module.find("SurfaceNormal") != std::string::npos && p->reference_normals_knn == 0)
          p->reference_normals_knn = 5;  // libpointmatcher's default knn of the surface-normal filters
      }
      if (module == "CounterTransformationChecker") saw_counter = true;
      if (module == "DifferentialTransformationChecker") saw_diff = true;
      continue;
    }
    if (val.empty()) continue;
    const double num = std::atof(val.c_str());
    if (section == "readingDataPointsFilters" || section == "referenceDataPointsFilters") {
      const bool reading = section == "readingDataPointsFilters";
      if (module == "RandomSamplingDataPointsFilter" && key == "prob" && reading) p->reading_sampling_prob = (float)num;
      if ((module == "SamplingSurfaceNormalDataPointsFilter" || module == "SurfaceNormalDataPointsFilter") && !reading) {
        if (key == "knn") p->reference_normals_knn = (int)num;
        if (key == "ratio") p->reference_sampling_ratio = (float)num;
      }
      continue;
    }
    if (module == "KDTreeMatcher") {
      if (key == "knn" && (int)num != 1) return LS_ERR_ARG;         // only 1-NN is built
      if (key == "epsilon" && num != 0.0) return LS_ERR_ARG;        // only the exact search is built
      if (key == "maxDist") return LS_ERR_ARG;                      // unbounded search only
    } else if (module == "TrimmedDistOutlierFilter") {
      if (key == "ratio") p->trim_ratio = (float)num;
    } else if (module == "CounterTransformationChecker") {
      if (key == "maxIterationCount") p->max_iterations = (int)num;
    } else if (module == "DifferentialTransformationChecker") {
      if (key == "minDiffRotErr") p->min_diff_rot = (float)num;
      if (key == "minDiffTransErr") p->min_diff_trans = (float)num;
      if (key == "smoothLength") p->smooth_length = (int)num;
    }
  }
  if (!saw_trim) p->trim_ratio = 1.0f;  // no outlier filter: every match has weight 1
  if (saw_checkers) {
    p->use_differential = saw_diff ? 1 : 0;
    if (!saw_counter) p->max_iterations = 40;  // libpointmatcher CounterTransformationChecker default
  }
  if (p->max_iterations < 1 || !(p->trim_ratio > 0.f) || p->trim_ratio > 1.f) return LS_ERR_ARG;
  if (p->use_differential && (p->smooth_length < 1 || p->smooth_length > 15)) return LS_ERR_ARG;
  return LS_OK;
}
