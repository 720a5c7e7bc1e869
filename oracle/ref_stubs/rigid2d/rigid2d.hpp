// TEST INFRASTRUCTURE ONLY -- minimal stand-in for the third-party header rigid2d/rigid2d.hpp (package `rigid2d` of
// nuturtle.rosinstall:1-6, absent from /root/reference), holding only what map/src/map/{map,grid,prm}.cpp use:
// the POD Vector2D and almost_equal with its published default epsilon.  Lets oracle/build_map_ref.py compile the
// reference's own map sources where they lie; nothing here is reference code.
#pragma once
#include <cmath>
namespace rigid2d {
struct Vector2D {
  double x = 0.0, y = 0.0;
  Vector2D() {}
  Vector2D(double x_, double y_) : x(x_), y(y_) {}
};
constexpr bool almost_equal(double d1, double d2, double epsilon = 1.0e-12) {
  return (d1 - d2 < 0 ? d2 - d1 : d1 - d2) < epsilon;
}
}  // namespace rigid2d
