// TEST INFRASTRUCTURE ONLY -- stand-in for nuslam/ekf.hpp (package `nuslam`, absent): prm.cpp only takes its random
// engine for PRM sampling, which the occupancy-grid path never calls.
#pragma once
#include <random>
#include <cmath>
#include <algorithm>
#include <iostream>
#include <stdexcept>
namespace nuslam {
inline std::mt19937& get_random() {
  static std::mt19937 mt{0};
  return mt;
}
}  // namespace nuslam
