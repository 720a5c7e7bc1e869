// tests/cpp/controller_check.cpp -- host-logic check of mppi::Controller (include/mppi.hpp) without a GPU: the engine
// is a deterministic stand-in, the poses come from stdin ("x y theta" per line, or "x y qx qy qz qw"), one line per
// callback goes to stdout.  tests/test_cpp_controller.py compares it with motion_planning_b200.Controller driven by the
// same stand-in (that mirror is itself pinned to the unmodified reference node).
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

#include "mppi.hpp"

struct StubEngine {   // setGoal / reset / step, like mppi::MPPI
  mppi::State goal{{0, 0, 0}};
  int steps = 0, resets = 0;
  void setGoal(const mppi::State& g) { goal = g; }
  void reset() { ++resets; }
  mppi::Control step(const mppi::State& x) {
    ++steps;
    return mppi::Control{{0.5 + 0.25 * (goal[0] - x[0]) - 0.125 * x[2], -0.75 + 0.5 * (goal[1] - x[1]) + 0.0625 * goal[2]}};
  }
};

int main(int argc, char** argv) {
  // usage: controller_check [x y]...            waypoint list (none = parallel park)
  //        controller_check track L [x y]...    follow the polyline with look-ahead L
  mppi::Controller<StubEngine>::Waypoints wps;
  const bool track = argc > 2 && std::string(argv[1]) == "track";
  for (int i = track ? 3 : 1; i + 1 < argc; i += 2) wps.push_back({std::atof(argv[i]), std::atof(argv[i + 1])});
  StubEngine eng;
  mppi::Controller<StubEngine> node(eng, track ? mppi::Controller<StubEngine>::Waypoints{} : wps);
  if (track) node.trackPath(wps, std::atof(argv[2]));
  std::string line;
  while (std::getline(std::cin, line)) {
    std::istringstream is(line);
    double v[6];
    int n = 0;
    while (n < 6 && (is >> v[n])) ++n;
    if (n == 0) continue;
    const mppi::Twist tw = (n == 6) ? node.posCb(v[0], v[1], v[2], v[3], v[4], v[5]) : node.posCb(v[0], v[1], v[2]);
    std::printf("%zu %d %d %.17g %.17g %.17g %.17g %.17g %d %d\n", node.idx(), (int)node.init(), (int)node.done(), tw.vx, tw.wz,
                node.goal()[0], node.goal()[1], node.goal()[2], eng.steps, eng.resets);
  }
  return 0;
}
