// tests/cpp/rk4_check.cpp -- mppi::RK4 (include/mppi.hpp), the reference's host-side integrator interface: integrate a control
// sequence read from stdin over the diff-drive ODE and print the trajectory at full precision.
//   stdin:  dt horizon n, x0 (3), then n controls (2 each);  stdout: one state per line
#include <cstdio>
#include <vector>

#include "mppi.hpp"

int main() {
  double dt, horizon;
  int n;
  if (std::scanf("%lf %lf %d", &dt, &horizon, &n) != 3) return 2;
  mppi::State x0;
  if (std::scanf("%lf %lf %lf", &x0[0], &x0[1], &x0[2]) != 3) return 2;
  std::vector<mppi::Control> u(n);
  for (auto& c : u)
    if (std::scanf("%lf %lf", &c[0], &c[1]) != 2) return 2;
  mppi::RK4<3, 2> rk4(dt);
  try {
    rk4.solve(x0, u, horizon);
    return 3;                                    // no ODE registered: must have thrown
  } catch (const std::logic_error&) {
  }
  rk4.registerODE(mppi::diffDriveOde(mppi::DiffDrive()));
  for (const auto& x : rk4.solve(x0, u, horizon)) std::printf("%.17g %.17g %.17g\n", x[0], x[1], x[2]);
  return 0;
}
