// mppi.hpp -- header-only C++ facade `mppi::MPPI` over the C ABI (include/mppi_b200.h).
//
// BASELINE.json's north_star asks for "the reference's mppi::MPPI C++ API surface (construct with
// diff-drive/bicycle dynamics functor, cost functor, K samples, T horizon; mppi.step(x0) -> u_t)".
// The reference has no such class (its MPPI is the Python class control/src/mppi:61-213, SURVEY.md
// section 0), so this facade is NEW; it follows the reference's C++ conventions instead:
//   - model objects constructed like control::models::DiffDrive(wheel_radius, wheel_base,
//     abs_max_wheel_vel)                    (control/include/control/Models.hpp:38-42)
//   - 0-success return codes + out-parameters available through the C ABI underneath
//                                           (control/include/control/TrajMPC.hpp:72,110-122)
//   - no Eigen dependency (std::array), so it builds in a plain catkin C++17 package
//     (control/CMakeLists.txt:5-9).
// On the device the dynamics and cost "functors" are compile-time kernel template arguments: the tag types
// below select the built-in instantiations, UserDynamics / UserKinematics / UserCost hand a caller's functor
// over as CUDA text that is compiled at run time.  A host std::function cannot run inside the rollout kernel;
// mppi::RK4 at the end of this file is the reference's host-side integrator interface (registerODE / solve)
// for code that simulates a plant beside the controller.
#ifndef MPPI_HPP_
#define MPPI_HPP_

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "mppi_b200.h"

namespace mppi {

using State = std::array<double, 3>;    // x, y, theta
using Control = std::array<double, 2>;  // (u_l, u_r) wheel rad/s, or (v, delta) for the bicycle

class Error : public std::runtime_error {
 public:
  Error(mppi_status s, const std::string& where)
      : std::runtime_error(where + ": status " + std::to_string(static_cast<int>(s)) + " (" + mppi_last_error() + ")"), status(s) {}
  mppi_status status;
};

// ---- dynamics functors (device-side: template tags) ------------------------------------------------
struct DiffDrive {   // dd_dynamics + rk4, control/src/mppi:23-30,39-54; ctor order as Models.hpp:38-42
  double wheel_radius = 0.033, wheel_base = 0.16, abs_max_wheel_vel = 6.35492;
  DiffDrive() = default;
  DiffDrive(double r, double L, double umax) : wheel_radius(r), wheel_base(L), abs_max_wheel_vel(umax) {}
  void apply(mppi_params& p) const {
    p.model = MPPI_MODEL_DIFF_DRIVE;
    p.wheel_radius = wheel_radius;
    p.wheel_base = wheel_base;
    p.u_max[0] = p.u_max[1] = abs_max_wheel_vel;
  }
};

struct Bicycle {     // NEW (BASELINE.json config 3): kinematic bicycle, u = (v, delta)
  double wheel_base = 0.16, max_speed = 0.22, max_steer = 0.6;
  double speed_noise_std = 0.08, steer_noise_std = 0.25;
  void apply(mppi_params& p) const {
    p.model = MPPI_MODEL_BICYCLE;
    p.wheel_base = wheel_base;
    p.u_max[0] = max_speed;
    p.u_max[1] = max_steer;
    p.noise_std[0] = speed_noise_std;
    p.noise_std[1] = steer_noise_std;
  }
};

struct UnicycleEuler {   // unicycle_dynamics + euler, control/src/mppi:33-36,57-58
  double max_speed = 0.22, max_yaw_rate = 2.84;
  void apply(mppi_params& p) const {
    p.model = MPPI_MODEL_UNICYCLE_EULER;
    p.u_max[0] = max_speed;
    p.u_max[1] = max_yaw_rate;
  }
};

// A caller-supplied model: the ODE functor `ode(x, u, xdot_out)` of the reference's C++ library (control/include/control/
// rk4.hpp:32,58; registerODE) -- on the device a functor is a template argument, so it is handed over as CUDA text and
// compiled for sm_100a at run time (mppi_create_user):
//   template <typename R> __device__ void mppi_user_ode(const R x[3], const R u[2], R xdot[3]) { ... }
struct UserDynamics {
  std::string ode_source;
  int integrator = 0;          // 0: RK4 with the control held (control/src/mppi:39-50), 1: explicit Euler (:57-58)
  bool wrap_theta = true;      // control/src/mppi:52-53
  double u_max0 = 6.35492, u_max1 = 6.35492;
  double noise_std0 = 0.9, noise_std1 = 0.9;
  int kind = 0;                // mppi_user_model.kind: 0 ODE functor, 1 kinematic functor (UserKinematics below)
  double speed_max = 0.0, yaw_rate_max = 0.0;
  UserDynamics() = default;
  explicit UserDynamics(std::string src) : ode_source(std::move(src)) {}
  void apply(mppi_params& p) const {
    p.model = MPPI_MODEL_USER;
    p.u_max[0] = u_max0;
    p.u_max[1] = u_max1;
    p.noise_std[0] = noise_std0;
    p.noise_std[1] = noise_std1;
  }
};

// A caller-supplied KINEMATIC model: forward speed and yaw rate as functions of the controls,
//   template <typename R> __device__ void mppi_user_speed_yaw(const R u[2], R* speed, R* yaw_rate) { ... }
// for xdot = speed cos(theta), ydot = speed sin(theta), thetadot = yaw_rate -- the family dd_dynamics / unicycle_dynamics
// (control/src/mppi:23-36) belong to.  The functor is dropped into the built-in kernels: every precision works, MIXED included
// (give speed_max / yaw_rate_max, bounds of |speed| and |yaw_rate| over the clipped controls).  integrator 0 = rk4 with the theta
// wrap (:39-54), 1 = Euler without (:57-58).
struct UserKinematics : UserDynamics {
  UserKinematics(std::string src, double speed_bound, double yaw_rate_bound, int integrator_ = 0) : UserDynamics(std::move(src)) {
    kind = 1;
    integrator = integrator_;
    wrap_theta = integrator_ == 0;
    speed_max = speed_bound;
    yaw_rate_max = yaw_rate_bound;
  }
};

// ---- cost functor ------------------------------------------------------------------------------------
struct OccupancyGrid {   // nav_msgs/OccupancyGrid layout as published by map/src/viz_grid.cpp:109-137
  std::vector<int8_t> cells;   // row-major, idx = ix + iy * width, values 0 / 50 / 100
  int width = 0, height = 0;
  double resolution = 1.0, origin_x = 0.0, origin_y = 0.0;
  double weight = 0.0;         // running cost += weight * cell / 100
};

struct QuadraticCost {   // get_cost + terminal cost, control/src/mppi:69-73,165-171,180-184
  std::array<double, 3> Q{{1e3, 1e3, 0.0}};
  std::array<double, 4> R{{1.0, 0.0, 0.0, 1.0}};
  std::array<double, 3> P1{{1e3, 1e3, 1e3}};
  OccupancyGrid grid;    // optional (weight == 0 -> none)
  void apply(mppi_params& p) const {
    for (int i = 0; i < 3; ++i) {
      p.q[i] = Q[i];
      p.p1[i] = P1[i];
    }
    for (int i = 0; i < 4; ++i) p.r[i] = R[i];
  }
};

// A caller-supplied cost functor (with UserDynamics): CUDA text defining
//   template <typename R> __device__ R mppi_user_running_cost(const R x[3], const R goal[3], const R u_nom[2], const R eps[2], int t);
//   template <typename R> __device__ R mppi_user_terminal_cost(const R x[3], const R goal[3]);
// in place of get_cost (control/src/mppi:180-184) and the terminal cost (:165-171)
struct UserCost {
  std::string source;
  OccupancyGrid grid;    // the occupancy-grid term stays available on top of a user cost
  UserCost() = default;
  explicit UserCost(std::string src) : source(std::move(src)) {}
  void apply(mppi_params&) const {}
};

struct Options {
  mppi_precision precision = MPPI_PRECISION_MIXED;
  mppi_weighting weighting = MPPI_WEIGHT_COST_TO_GO;
  double sigma = 0.9;      // sig = sigma * I, control/src/mppi:88
  double lambda = 1e-3;    // control/src/mppi:89
  uint64_t seed = 0;
  int device = 0;
};

class MPPI {
 public:
  // construct with a dynamics functor, a cost functor, K samples and T horizon
  template <typename Dynamics, typename Cost = QuadraticCost>
  MPPI(const Dynamics& dyn, const Cost& cost, int K, int T, const Options& opt = Options()) : T_(T) {
    mppi_params p;
    check(mppi_default_params(&p), "mppi_default_params");
    p.K = K;
    p.T = T;
    p.precision = opt.precision;
    p.weighting = opt.weighting;
    p.sig[0] = p.sig[3] = opt.sigma;
    p.noise_std[0] = p.noise_std[1] = opt.sigma;
    p.lambda = opt.lambda;
    p.seed = opt.seed;
    p.device = opt.device;
    dyn.apply(p);
    cost.apply(p);
    constexpr bool user_dyn = std::is_base_of<UserDynamics, Dynamics>::value, user_cost = std::is_same<Cost, UserCost>::value;
    static_assert(user_dyn || !user_cost, "a UserCost functor needs UserDynamics (the kernels are instantiated for the pair)");
    if constexpr (user_dyn) {
      std::string text = dyn.ode_source;
      if constexpr (user_cost) text += "\n" + cost.source;
      mppi_user_model um = {};
      um.source = text.c_str();
      um.integrator = dyn.integrator;
      um.wrap_theta = dyn.wrap_theta ? 1 : 0;
      um.has_cost = user_cost ? 1 : 0;
      um.kind = dyn.kind;
      um.speed_max = dyn.speed_max;
      um.yaw_rate_max = dyn.yaw_rate_max;
      // MIXED needs a kinematic functor with stated bounds and the built-in cost; otherwise F64 (the default of user models)
      if (p.precision == MPPI_PRECISION_MIXED && !(dyn.kind == 1 && !user_cost && dyn.speed_max > 0)) p.precision = MPPI_PRECISION_F64;
      check(mppi_create_user(&p, &um, &h_), "mppi_create_user");
    } else {
      check(mppi_create(&p, &h_), "mppi_create");
    }
    if (cost.grid.weight != 0.0 && !cost.grid.cells.empty()) {
      const mppi_status st = mppi_set_grid(h_, cost.grid.cells.data(), cost.grid.width, cost.grid.height, cost.grid.resolution,
                                           cost.grid.origin_x, cost.grid.origin_y, cost.grid.weight);
      if (st != MPPI_OK) {   // the destructor does not run for a constructor that throws: release the engine here
        mppi_destroy(h_);
        h_ = nullptr;
        check(st, "mppi_set_grid");
      }
    }
  }
  MPPI(const MPPI&) = delete;
  MPPI& operator=(const MPPI&) = delete;
  MPPI(MPPI&& o) noexcept : h_(o.h_), T_(o.T_), x_next_(o.x_next_) { o.h_ = nullptr; }
  ~MPPI() {
    if (h_) mppi_destroy(h_);
  }

  void setGoal(const State& g) { check(mppi_set_goal(h_, g.data()), "mppi_set_goal"); }
  void reset() { check(mppi_reset(h_), "mppi_reset"); }   // MPPI.initialize, control/src/mppi:79-83

  // mppi.step(x0) -> u_t   (= MPPI.get_path; u_t = uvec[-1], control/src/mppi:85-102,379)
  Control step(const State& x0) {
    Control u;
    check(mppi_step(h_, x0.data(), u.data(), x_next_.data()), "mppi_step");
    return u;
  }
  const State& predictedNextState() const { return x_next_; }   // what get_path returns, :94,102

  std::vector<double> nominal() const {   // latest_uvec (2,T) row-major
    std::vector<double> U(2 * T_);
    check(mppi_get_nominal(h_, U.data()), "mppi_get_nominal");
    return U;
  }
  void setNominal(const std::vector<double>& U) {
    if (static_cast<int>(U.size()) != 2 * T_) throw std::invalid_argument("nominal must have 2*T entries");
    check(mppi_set_nominal(h_, U.data()), "mppi_set_nominal");
  }

  // incremental map update (Grid::update_grid, map/src/map/grid.cpp:155-199): overwrite rows y0.., columns x0.. of the
  // resident occupancy grid with the row-major (height, width) patch; asynchronous, ordered before the next step
  void updateGrid(const std::vector<int8_t>& patch, int x0, int y0, int width, int height) {
    if (static_cast<long long>(patch.size()) != static_cast<long long>(width) * height)
      throw std::invalid_argument("patch must have width*height cells");
    check(mppi_update_grid(h_, patch.data(), x0, y0, width, height), "mppi_update_grid");
  }

  // twist for cmd_vel, Controller.wheelsToTwist (control/src/mppi:319-325)
  static void wheelsToTwist(const Control& u, double wheel_radius, double wheel_base, double& vx, double& wz) {
    vx = wheel_radius * (u[0] + u[1]) / 2.0;
    wz = wheel_radius * (-u[0] + u[1]) / wheel_base;
  }

  mppi_handle handle() const { return h_; }

 private:
  static void check(mppi_status s, const char* where) {
    if (s != MPPI_OK) throw Error(s, where);
  }
  mppi_handle h_ = nullptr;
  int T_ = 0;
  State x_next_{{0, 0, 0}};
};

// ---- the caller: Controller (control/src/mppi:296-389) for a C++ node -------------------------------------
// Same state machine as the reference's rospy node and as motion_planning_b200.Controller (the Python mirror that is
// tested against the unmodified reference): one odometry sample in, one MPPI step, one Twist out; waypoint list
// (cyclic) or parallel park.  `Engine` needs setGoal(State), reset(), step(State) -> Control; it defaults to
// mppi::MPPI and is a template parameter so that the host logic can be exercised without a GPU.
struct Twist {
  double vx = 0.0, wz = 0.0;   // geometry_msgs/Twist linear.x, angular.z (control/src/mppi:383-385)
};

// yaw of tf.transformations.euler_from_quaternion(q)[2] (call site control/src/mppi:333-334): the two matrix entries of
// tf's quaternion_matrix the yaw needs, the quaternion scaled by sqrt(2 / |q|^2)
inline double yawFromQuaternion(double x, double y, double z, double w) {
  const double nq = x * x + y * y + z * z + w * w;
  if (nq < 8.881784197001252e-16) return 0.0;
  const double s = std::sqrt(2.0 / nq);
  x *= s;
  y *= s;
  z *= s;
  w *= s;
  return std::atan2(x * y + z * w, 1.0 - y * y - z * z);
}

template <typename Engine = MPPI>
class Controller {
 public:
  using Waypoints = std::vector<std::array<double, 2>>;
  // waypoints empty = parallel park (control/src/mppi:305-309); thresh = MPPI.thresh (:62,74)
  explicit Controller(Engine& engine, Waypoints waypoints = {}, double thresh = 0.05, double wheel_radius = 0.033,
                      double wheel_base = 0.16)
      : mppi_(engine), waypoints_(std::move(waypoints)), thresh_(thresh), r_(wheel_radius), L_(wheel_base) {
    parallel_park_ = waypoints_.empty();
  }

  // planner hand-off: a new vertex list (e.g. the path of global_planner's trace_path); re-initialises towards its head
  void setWaypoints(Waypoints waypoints) {
    waypoints_ = std::move(waypoints);
    tracking_ = false;
    parallel_park_ = waypoints_.empty();
    idx_ = 0;
    init_ = true;
    done_ = false;
  }

  // planner hand-off, tracking variant (NEW; same semantics as motion_planning_b200.Controller.track_path): follow the
  // polyline with a moving look-ahead goal -- the point `lookahead` metres of arc length beyond the robot's projection onto
  // the path, heading along the path -- instead of stopping at every vertex; the nominal sequence is kept between samples
  void trackPath(Waypoints path, double lookahead = 0.3) {
    track_ = std::move(path);
    lookahead_ = lookahead;
    progress_ = 0.0;
    tracking_ = !track_.empty();
    parallel_park_ = false;
    idx_ = 0;
    init_ = true;
    done_ = false;
  }

  // look-ahead point of a polyline: (x, y, heading of the path there); s_proj in/out = arc length of the projection (monotone)
  static State lookaheadGoal(const Waypoints& p, double px, double py, double lookahead, double& s_proj) {
    const size_t n = p.size();
    if (n == 1) return State{{p[0][0], p[0][1], 0.0}};
    std::vector<double> len(n - 1), cum(n, 0.0);
    for (size_t i = 0; i + 1 < n; ++i) {
      len[i] = std::hypot(p[i + 1][0] - p[i][0], p[i + 1][1] - p[i][1]);
      cum[i + 1] = cum[i] + len[i];
    }
    const double s_min = s_proj;
    double best = 1e300, sp = s_min;
    for (size_t i = 0; i + 1 < n; ++i) {
      if (len[i] == 0.0 || cum[i + 1] < s_min) continue;
      const double ex = p[i + 1][0] - p[i][0], ey = p[i + 1][1] - p[i][1];
      double u = ((px - p[i][0]) * ex + (py - p[i][1]) * ey) / (len[i] * len[i]);
      u = std::min(1.0, std::max(u, std::max(0.0, (s_min - cum[i]) / len[i])));
      const double d = std::hypot(px - (p[i][0] + u * ex), py - (p[i][1] + u * ey));
      if (d < best - 1e-12) {
        best = d;
        sp = cum[i] + u * len[i];
      }
    }
    s_proj = sp;
    const double s_goal = std::min(sp + lookahead, cum[n - 1]);
    size_t i = 0;
    while (i + 2 < n && cum[i + 1] <= s_goal) ++i;       // last segment whose start is <= s_goal
    while (len[i] == 0.0 && i > 0) --i;
    const double u = len[i] > 0.0 ? (s_goal - cum[i]) / len[i] : 0.0;
    const double ex = p[i + 1][0] - p[i][0], ey = p[i + 1][1] - p[i][1];
    return State{{p[i][0] + u * ex, p[i][1] + u * ey, std::atan2(ey, ex)}};
  }

  // Controller.pos_cb, control/src/mppi:328-386
  Twist posCb(double x, double y, double theta) {
    if (tracking_) return trackCb(x, y, theta);
    start_ = State{{x, y, theta}};                                        // :335
    stepped_ = false;
    if (parallel_park_) goal_ = State{{0.0, -1.0, 0.0}};                  // :337
    if (std::hypot(start_[0] - goal_[0], start_[1] - goal_[1]) > thresh_ && !init_) {
      mppi_.setGoal(goal_);
      last_u_ = mppi_.step(start_);                                       // :341 -- the hot path
      stepped_ = true;
      done_ = false;
    } else if (init_) {                                                   // :344-354
      initialize();
      if (!parallel_park_) goal_ = goalTowards(waypoints_[idx_]);
      init_ = false;
    } else {
      if (!parallel_park_) {                                              // :357-373: next waypoint, cyclic
        idx_ = (idx_ + 1 >= waypoints_.size()) ? 0 : idx_ + 1;
        initialize();
        goal_ = goalTowards(waypoints_[idx_]);
      } else {
        done_ = true;                                                     // :375
      }
    }
    const Control u = done_ ? Control{{0.0, 0.0}} : last_u_;              // :378-381
    Twist tw;                                                             // wheelsToTwist, :320-326
    tw.vx = r_ * (u[0] + u[1]) / 2.0;
    tw.wz = r_ * (-u[0] + u[1]) / L_;
    return tw;
  }
  Twist posCb(double x, double y, double qx, double qy, double qz, double qw) {   // nav_msgs/Odometry pose
    return posCb(x, y, yawFromQuaternion(qx, qy, qz, qw));
  }

  size_t idx() const { return idx_; }
  bool init() const { return init_; }
  bool done() const { return done_; }
  bool stepped() const { return stepped_; }   // did the last posCb run an MPPI step (or only (re)initialise / stop)?
  bool parallelPark() const { return parallel_park_; }
  const State& goal() const { return goal_; }

 private:
  Twist trackCb(double x, double y, double theta) {
    start_ = State{{x, y, theta}};
    stepped_ = false;
    goal_ = lookaheadGoal(track_, x, y, lookahead_, progress_);
    const std::array<double, 2>& end = track_.back();
    if (init_) {
      initialize();
      init_ = false;
    } else if (std::hypot(x - end[0], y - end[1]) > thresh_) {
      mppi_.setGoal(goal_);
      last_u_ = mppi_.step(start_);
      stepped_ = true;
      done_ = false;
    } else {
      done_ = true;
    }
    const Control u = done_ ? Control{{0.0, 0.0}} : last_u_;
    Twist tw;
    tw.vx = r_ * (u[0] + u[1]) / 2.0;
    tw.wz = r_ * (-u[0] + u[1]) / L_;
    return tw;
  }
  void initialize() {            // MPPI.initialize (:79-83): uvec = [[0, 0]]
    mppi_.reset();
    last_u_ = Control{{0.0, 0.0}};
  }
  State goalTowards(const std::array<double, 2>& w) const {   // :347-352
    return State{{w[0], w[1], std::atan2(w[1] - start_[1], w[0] - start_[0])}};
  }
  Engine& mppi_;
  Waypoints waypoints_, track_;
  double lookahead_ = 0.3, progress_ = 0.0;
  bool tracking_ = false;
  double thresh_, r_, L_;
  bool parallel_park_ = true, init_ = true, done_ = false, stepped_ = false;
  size_t idx_ = 0;
  State start_{{0, 0, 0}}, goal_{{0, 0, 0}};
  Control last_u_{{0, 0}};
};

// ---- host-side integrator (NOT on the engine's path) ------------------------------------------------
// The interface of the reference's control::RK4 (control/include/control/rk4.hpp:19-62) without Eigen: a fixed-step classic
// Runge-Kutta over a registered ODE `ode(x, u, xdot_out)`, the control held over each step.  For host code around the
// controller -- a simulated plant that feeds odometry back, a planner that previews a control sequence; the rollouts of MPPI
// itself never come through here (they run in the kernels, from the same ODE handed over as text: UserDynamics).
//   solve(x0, U, horizon): floor(horizon / dt) steps, step i uses column i of U; returns the states AFTER each step
//   (control/src/control/rk4.cpp:56-84); the stage order is that of rk4.cpp:115-138
template <size_t NX, size_t NU>
class RK4 {
 public:
  using X = std::array<double, NX>;
  using U = std::array<double, NU>;
  using Ode = std::function<void(const X&, const U&, X&)>;
  explicit RK4(double dt) : dt_(dt) {}
  void registerODE(Ode ode) { ode_ = std::move(ode); }
  void integrate(X& x, const U& u) const {   // one step, in place
    if (!ode_) throw std::logic_error("mppi::RK4: no ODE registered");
    X k1{}, k2{}, k3{}, k4{}, tmp{};
    ode_(x, u, k1);
    for (size_t i = 0; i < NX; ++i) tmp[i] = x[i] + dt_ * (0.5 * k1[i]);
    ode_(tmp, u, k2);
    for (size_t i = 0; i < NX; ++i) tmp[i] = x[i] + dt_ * (0.5 * k2[i]);
    ode_(tmp, u, k3);
    for (size_t i = 0; i < NX; ++i) tmp[i] = x[i] + dt_ * k3[i];
    ode_(tmp, u, k4);
    for (size_t i = 0; i < NX; ++i) x[i] += (dt_ / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
  }
  std::vector<X> solve(const X& x0, const std::vector<U>& u, double horizon) const {
    const size_t n = static_cast<size_t>(horizon / dt_);
    if (u.size() < n) throw std::invalid_argument("mppi::RK4::solve: the control signal is shorter than horizon / dt");
    std::vector<X> traj;
    traj.reserve(n);
    X x = x0;
    for (size_t i = 0; i < n; ++i) {
      integrate(x, u[i]);
      traj.push_back(x);
    }
    return traj;
  }
  double dt() const { return dt_; }

 private:
  double dt_;
  Ode ode_;
};

// the reference's diff-drive ODE (dd_dynamics, control/src/mppi:23-30) for mppi::RK4<3, 2>::registerODE
inline RK4<3, 2>::Ode diffDriveOde(const DiffDrive& robot) {
  const double half_r = 0.5 * robot.wheel_radius, r_over_L = robot.wheel_radius / robot.wheel_base;
  return [half_r, r_over_L](const State& x, const Control& u, State& xdot) {
    xdot[0] = half_r * std::cos(x[2]) * (u[0] + u[1]);
    xdot[1] = half_r * std::sin(x[2]) * (u[0] + u[1]);
    xdot[2] = r_over_L * (u[1] - u[0]);
  };
}

}  // namespace mppi
#endif  // MPPI_HPP_
