// examples/ros_control_node.cpp -- what the C++ side of a ROS control node looks like on top of
// mppi::MPPI (include/mppi.hpp).  ROS is not available in this image, so the odometry source is a
// simulated diff-drive (the reference's own fake-encoder setup, control/launch/mppi_pentagon.launch:5-9)
// and cmd_vel is printed; the body of the loop is exactly Controller.pos_cb (control/src/mppi:327-386):
// waypoint state machine -> mppi.step(x0) -> wheelsToTwist -> publish.
//
//   g++ -std=c++17 -Iinclude examples/ros_control_node.cpp -Lmotion_planning_b200/lib -lmppi_b200 \
//       -Wl,-rpath,$PWD/motion_planning_b200/lib -o examples/ros_control_node
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mppi.hpp"

int main(int argc, char** argv) {
  const int K = argc > 1 ? std::atoi(argv[1]) : 4096, T = argc > 2 ? std::atoi(argv[2]) : 64;
  const int max_iters = argc > 3 ? std::atoi(argv[3]) : 400;
  // waypoints: control/config/waypoints.yaml:1 (pentagon); empty list = parallel park (parallel.yaml:1)
  const std::vector<std::array<double, 2>> waypoints = {{1, 0}, {2, 1}, {1, 2}, {0, 2}, {0, 0}};
  const double thresh = 0.05;                               // control/src/mppi:62
  mppi::DiffDrive robot;                                    // control/src/mppi:18-20
  mppi::MPPI ctl(robot, mppi::QuadraticCost(), K, T);
  mppi::State x{{0.0, 0.0, 0.0}};
  size_t idx = 0;
  auto goal_for = [&](size_t i) {                           // control/src/mppi:346-352
    return mppi::State{{waypoints[i][0], waypoints[i][1], std::atan2(waypoints[i][1] - x[1], waypoints[i][0] - x[0])}};
  };
  mppi::State goal = goal_for(idx);
  ctl.setGoal(goal);
  int reached = 0;
  for (int it = 0; it < max_iters; ++it) {
    if (std::hypot(x[0] - goal[0], x[1] - goal[1]) <= thresh) {   // control/src/mppi:339-340,357-373
      ++reached;
      idx = (idx + 1 >= waypoints.size()) ? 0 : idx + 1;
      ctl.reset();
      goal = goal_for(idx);
      ctl.setGoal(goal);
      std::printf("WAYPOINT REACHED at iter %d -> next (%.1f, %.1f)\n", it, goal[0], goal[1]);
      if (reached == static_cast<int>(waypoints.size())) break;
    }
    const mppi::Control u = ctl.step(x);                      // control/src/mppi:341,379
    double vx, wz;
    mppi::MPPI::wheelsToTwist(u, robot.wheel_radius, robot.wheel_base, vx, wz);   // :382
    if (it % 50 == 0) std::printf("iter %4d  x=(%.3f %.3f %.3f)  cmd_vel: vx=%.4f wz=%.4f\n", it, x[0], x[1], x[2], vx, wz);
    x = ctl.predictedNextState();                             // simulated odometry = the model itself
  }
  std::printf("final state (%.3f %.3f %.3f), waypoints reached: %d\n", x[0], x[1], x[2], reached);
  return 0;
}
