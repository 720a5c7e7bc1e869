// TEST INFRASTRUCTURE ONLY -- C entry points over the reference's own map::Grid (compiled from /root/reference/map/src/map/
// {map,grid,prm}.cpp where they lie by oracle/build_map_ref.py), so that oracle/map_grid.py can be checked cell by cell
// against the real thing.  What viz_grid.cpp:84-137 does, without ROS: scale the obstacle vertices, Grid(obstacles, inflate),
// build_map(resolution), occupancy_grid(map).
#include <cstdint>
#include <cstring>
#include <vector>

#include "map/grid.hpp"

extern "C" {

// obstacles: flat (x, y) pairs, nverts[i] vertices for obstacle i.  Returns the number of cells (W*H), writes W, H, origin.
// cells may be NULL to query the size first.
int mapref_build(const double* xy, const int* nverts, int nobs, double scale, double resolution, double inflate,
                 int8_t* cells, int ncap, int* W, int* H, double* origin_xy) {
  std::vector<map::Obstacle> obs;
  int k = 0;
  for (int i = 0; i < nobs; ++i) {
    map::Obstacle o;
    for (int j = 0; j < nverts[i]; ++j, ++k) {
      rigid2d::Vector2D v(xy[2 * k], xy[2 * k + 1]);
      v.x /= scale;     // map/src/viz_grid.cpp:88-90
      v.y /= scale;
      o.vertices.push_back(v);
    }
    obs.push_back(o);
  }
  map::Grid grid(obs, inflate);
  grid.build_map(resolution);
  std::vector<int8_t> m;
  grid.occupancy_grid(m);
  auto dims = grid.return_grid_dimensions();
  auto bounds = grid.return_map_bounds();
  *W = dims.at(0);
  *H = dims.at(1);
  origin_xy[0] = bounds.at(0).x;
  origin_xy[1] = bounds.at(0).y;
  if (cells && (int)m.size() <= ncap) std::memcpy(cells, m.data(), m.size());
  return (int)m.size();
}

// the incrementally revealed grid: reveal around the cells of a path (grid indices), one update_grid per path point
// (grid.cpp:155-173), then fake_occupancy_grid (:182-199)
int mapref_reveal(const double* xy, const int* nverts, int nobs, double scale, double resolution, double inflate,
                  const int* path_ixiy, int npath, int visibility, int8_t* fake, int ncap) {
  std::vector<map::Obstacle> obs;
  int k = 0;
  for (int i = 0; i < nobs; ++i) {
    map::Obstacle o;
    for (int j = 0; j < nverts[i]; ++j, ++k) o.vertices.push_back(rigid2d::Vector2D(xy[2 * k] / scale, xy[2 * k + 1] / scale));
    obs.push_back(o);
  }
  map::Grid grid(obs, inflate);
  grid.build_map(resolution);
  auto cells = grid.return_grid();
  auto dims = grid.return_grid_dimensions();
  for (int p = 0; p < npath; ++p) {
    const int rmj = map::grid2rowmajor(path_ixiy[2 * p], path_ixiy[2 * p + 1], dims.at(0));
    grid.update_grid(cells.at(rmj), visibility);
  }
  std::vector<int8_t> m;
  grid.fake_occupancy_grid(m);
  if (fake && (int)m.size() <= ncap) std::memcpy(fake, m.data(), m.size());
  return (int)m.size();
}
}
