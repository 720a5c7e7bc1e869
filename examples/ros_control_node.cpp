// examples/ros_control_node.cpp -- what the C++ side of a ROS control node looks like on top of
// mppi::MPPI + mppi::Controller (include/mppi.hpp).  ROS is not available in this image, so the odometry source is a
// simulated diff-drive (the reference's own fake-encoder setup, control/launch/mppi_pentagon.launch:5-9: here the
// engine's predicted next state is fed back as the next odometry sample) and cmd_vel is printed.  With ROS the body of
// the loop is the odometry callback: node.posCb(x, y, qx, qy, qz, qw) -> Twist -> cmd_vel (control/src/mppi:327-386).
//
//   g++ -std=c++17 -Iinclude examples/ros_control_node.cpp -Lmotion_planning_b200/lib -lmppi_b200 -Wl,-rpath,$PWD/motion_planning_b200/lib -o examples/ros_control_node
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "mppi.hpp"

int main(int argc, char** argv) {
  const int K = argc > 1 ? std::atoi(argv[1]) : 4096, T = argc > 2 ? std::atoi(argv[2]) : 64;
  const int max_iters = argc > 3 ? std::atoi(argv[3]) : 400;
  // waypoints: control/config/waypoints.yaml:1 (pentagon); an empty list = parallel park (parallel.yaml:1)
  const mppi::Controller<>::Waypoints waypoints = {{1, 0}, {2, 1}, {1, 2}, {0, 2}, {0, 0}};
  mppi::DiffDrive robot;                                    // control/src/mppi:18-20
  // 4th argument "user": the same robot as a caller-supplied ODE functor -- the reference's registerODE / model= hook
  // (control/include/control/rk4.hpp:32,58; control/src/mppi:62,66) -- compiled at run time for this GPU
  //               "kinematic": the robot as a kinematic functor (speed and yaw rate from the wheel speeds), which runs the
  //               headline mixed-precision pipeline
  const bool user = argc > 4 && !std::strcmp(argv[4], "user");
  const bool kinematic = argc > 4 && !std::strcmp(argv[4], "kinematic");
  std::unique_ptr<mppi::MPPI> engine_ptr;
  if (kinematic) {
    const double r = robot.wheel_radius, L = robot.wheel_base, wmax = 6.35492;
    mppi::UserKinematics dyn(
        "template <typename R> __device__ void mppi_user_speed_yaw(const R u[2], R* speed, R* yaw_rate) {\n"
        "  *speed = R(0.5 * 0.033) * (u[0] + u[1]);\n"
        "  *yaw_rate = R(0.033 / 0.16) * (u[1] - u[0]);\n"
        "}\n",
        /*speed bound=*/r * wmax, /*yaw-rate bound=*/r / L * 2 * wmax);
    engine_ptr.reset(new mppi::MPPI(dyn, mppi::QuadraticCost(), K, T));   // Options().precision = MIXED
    std::printf("dynamics: user-supplied kinematic functor (NVRTC), precision mixed\n");
  } else if (user) {
    mppi::UserDynamics dyn(
        "template <typename R> __device__ void mppi_user_ode(const R x[3], const R u[2], R xdot[3]) {\n"
        "  const R r = R(0.033), L = R(0.16);\n"
        "  xdot[0] = (r / R(2.0)) * cos(x[2]) * (u[0] + u[1]);\n"
        "  xdot[1] = (r / R(2.0)) * sin(x[2]) * (u[0] + u[1]);\n"
        "  xdot[2] = (r / L) * (u[1] - u[0]);\n"
        "}\n");
    engine_ptr.reset(new mppi::MPPI(dyn, mppi::QuadraticCost(), K, T));
    std::printf("dynamics: user-supplied ODE functor (NVRTC)\n");
  } else {
    engine_ptr.reset(new mppi::MPPI(robot, mppi::QuadraticCost(), K, T));
  }
  mppi::MPPI& engine = *engine_ptr;
  mppi::Controller<> node(engine, waypoints, /*thresh=*/0.05, robot.wheel_radius, robot.wheel_base);
  mppi::State x{{0.0, 0.0, 0.0}};
  int reached = 0;
  size_t last_idx = node.idx();
  for (int it = 0; it < max_iters; ++it) {
    const mppi::Twist tw = node.posCb(x[0], x[1], x[2]);      // one odometry sample -> one MPPI step -> one Twist
    if (node.idx() != last_idx) {
      ++reached;
      last_idx = node.idx();
      std::printf("WAYPOINT REACHED at iter %d -> next (%.1f, %.1f)\n", it, node.goal()[0], node.goal()[1]);
      if (reached == static_cast<int>(waypoints.size())) break;
    }
    if (it % 50 == 0) std::printf("iter %4d  x=(%.3f %.3f %.3f)  cmd_vel: vx=%.4f wz=%.4f\n", it, x[0], x[1], x[2], tw.vx, tw.wz);
    // simulated odometry = the model itself (callbacks that only (re)initialise do not move the robot)
    if (node.stepped()) x = engine.predictedNextState();
  }
  std::printf("final state (%.3f %.3f %.3f), waypoints reached: %d\n", x[0], x[1], x[2], reached);
  return 0;
}
