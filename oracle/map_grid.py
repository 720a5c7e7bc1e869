"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the map package's occupancy grid, the input format of the
occupancy-grid cost term (BASELINE.json config 4).

Restates, citing the reference lines each function follows (paths relative to /root/reference/):
  * map/src/map/map.cpp:47-76      Map::find_map_extent      -> map_extent
  * map/src/map/prm.cpp:422-461    lineToPoint               -> line_to_point
  * map/src/map/prm.cpp:267-394    not_inside                -> not_inside
  * map/src/map/grid.cpp:17-69     Grid::build_map           -> build_map
  * map/src/map/grid.cpp:126-144   Grid::occupancy_grid      -> occupancy values 0 / 50 / 100
  * map/src/map/grid.cpp:155-237   update_grid / fake_occupancy_grid / get_neighbours -> FakeGrid (incremental reveal)
  * map/src/viz_grid.cpp:50-137    the node that scales the obstacles and publishes nav_msgs/OccupancyGrid

Third-party pieces absent from /root/reference (the `rigid2d` / `nuslam` packages of nuturtle.rosinstall:1-6, Eigen 3):
`rigid2d::almost_equal(a, b)` is restated as |a - b| < 1e-12 (its published default epsilon), `euclidean_distance(dx, dy)`
as sqrt(dx^2 + dy^2), Eigen's `squaredNorm` / `normalized` / `dot` as the plain two-term expressions in double.

PARITY PINNING: `oracle/build_map_ref.py` compiles the reference's OWN map.cpp / grid.cpp / prm.cpp where they lie, against
minimal stand-ins of those third-party headers, into oracle/_ref/libmapref.so; tests/test_map_grid.py checks this restatement
cell by cell against it (and against tests/golden/map_grid_*.npz made from it).

Only tests/, __graft_entry__.smoke() and bench.py's checker legs may import this module; the product never does.
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

FREE, INFLATION, OCCUPIED = 0, 50, 100        # map/src/map/grid.cpp:126-144

# map/config/map.yaml:4-18 (coordinates in "cell" units; the nodes divide them by `scale`, map/src/viz_grid.cpp:88-90)
MAP_YAML_OBSTACLES = [
    [[12., 6.], [14.5, 3.5], [17., 5.5], [17., 8.5], [14., 8.]],
    [[24., 6.], [26., 3.5], [31., 7.5], [24.5, 9.5]],
    [[34., 26.], [10., 26.], [10., 12.], [34., 12.]],
    [[0., 26.], [0., 6.], [4., 6.], [4., 26.]],
    [[4., 32.], [6., 30.], [8., 32.]],
    [[17., 32.], [18., 30.], [19., 32.]],
    [[0., 36.], [0., 32.], [29., 32.], [29., 36.]],
    [[34., 36.], [33., 34.], [34., 32.]],
    [[6., 44.], [2., 43.], [2., 39.], [6., 38.], [8., 41.]],
    [[11., 48.], [17., 41.], [14., 48.]],
    [[30., 48.], [22., 40.], [32., 48.]],
    [[0., 0.], [34., 0.]],
    [[34., 0.], [34., 48.]],
    [[34., 48.], [0., 48.]],
    [[0., 48.], [0., 0.]],
]


def almost_equal(a, b, eps=1.0e-12):
    """rigid2d::almost_equal (third party, absent): |a - b| < eps."""
    return np.abs(a - b) < eps


def scale_obstacles(obstacles, scale):
    """map/src/viz_grid.cpp:84-92: every vertex coordinate is divided by SCALE."""
    return [[(float(x) / scale, float(y) / scale) for x, y in ob] for ob in obstacles]


def map_extent(obstacles):
    """Map::find_map_extent, map/src/map/map.cpp:47-76.  The reference leaves x_min, x_max, y_min uninitialised (:49, only
    y_max = 0); they are taken as 0 here, which is what the published demos rely on (all coordinates are >= 0).  Note the
    `else if`: a vertex that raises the maximum is never considered for the minimum."""
    x_min = x_max = y_min = y_max = 0.0
    for ob in obstacles:
        for (x, y) in ob:
            if x > x_max:
                x_max = x
            elif x < x_min:
                x_min = x
            if y > y_max:
                y_max = y
            elif y < y_min:
                y_min = y
    return (x_min, y_min), (x_max, y_max)


def arange_accumulate(start, stop, step):
    """map::arange<double>, map/include/map/grid.hpp:119-126: `for (v = start; v < stop; v += step)` -- the cell origins
    are ACCUMULATED sums, not i * step (so the cell count at an exactly divisible extent depends on rounding)."""
    out = []
    v = float(start)
    while v < stop:
        out.append(v)
        v += step
    return np.array(out, dtype=np.float64)


def line_to_point(A, B, P):
    """lineToPoint, map/src/map/prm.cpp:422-461.  A, B: (2,) segment end points; P: (..., 2) points.  Returns (u, D):
    u = position of the foot point along A->B (:433-434), D = signed distance, > 0 on the left of A->B (:442-459)."""
    ax, ay = A
    bx, by = B
    px, py = P[..., 0], P[..., 1]
    ex, ey = bx - ax, by - ay
    u = ((px - ax) * ex + (py - ay) * ey) / (ex * ex + ey * ey)
    nx, ny = -ey, ex                                   # inward (left-hand) normal, :442
    nn = np.sqrt(nx * nx + ny * ny)
    nx, ny = nx / nn, ny / nn                          # Eigen normalized(), :452
    D = (px - ax) * nx + (py - ay) * ny                # :455-459
    return u, D


def not_inside(P, obstacles, inflate):
    """not_inside, map/src/map/prm.cpp:267-394, for an array of points P (..., 2) at once.  The reference returns `false`
    early in several branches and otherwise ANDs `!on_all_left` over the obstacles; no branch has a side effect, so the
    result is  free = !(any early return) && !(any obstacle with on_all_left)  -- evaluated here without the early exits."""
    shape = P.shape[:-1]
    early = np.zeros(shape, dtype=bool)
    not_free = np.zeros(shape, dtype=bool)
    px, py = P[..., 0], P[..., 1]
    for ob in obstacles:
        on_all_left = np.ones(shape, dtype=bool)
        n = len(ob)
        for i in range(n):
            A, B = ob[i], ob[(i + 1) % n]              # :283-303 (last vertex pairs with vertex 0)
            u, D = line_to_point(A, B, P)
            in_win = ((u >= 0.0) & (u <= 1.0)) | almost_equal(u, 0.0) | almost_equal(u, 1.0)
            dB = np.sqrt((B[0] - px) ** 2 + (B[1] - py) ** 2)          # euclidean_distance (third party, absent)
            dA = np.sqrt((A[0] - px) ** 2 + (A[1] - py) ** 2)
            neg = D < 0.0                                                # :307
            zero = ~neg & almost_equal(D, 0.0)
            # D < 0, inside the segment window: outside the inflation band -> not inside this obstacle (:312-318)
            on_all_left &= ~(neg & in_win & (D < -inflate))
            # D < 0 (or D ~ 0 off the segment, :352-380): closest to an end point
            for beyond, dist in (((u > 1.0), dB), ((u < 0.0), dA)):     # :321-343, :355-379
                m = (neg & ~in_win | zero & ~in_win) & beyond
                on_all_left &= ~(m & (dist > inflate))
                early |= m & ~(dist > inflate)
            # on an edge, within the segment: disqualified immediately (:346-351)
            early |= zero & in_win
        not_free |= on_all_left                                          # :383-387
    return ~(early | not_free)


def build_map(obstacles, resolution, inflate):
    """Grid::build_map + Grid::occupancy_grid, map/src/map/grid.cpp:17-69,126-144.  Returns (cells int8 (H, W) row-major with
    idx = x + y * W (:251-266), resolution, origin = map_min (map/src/viz_grid.cpp:112-114))."""
    (x_min, y_min), (x_max, y_max) = map_extent(obstacles)
    xc = arange_accumulate(x_min, x_max, resolution)                     # :20
    yc = arange_accumulate(y_min, y_max, resolution)                     # :21
    off = resolution / 2.0                                               # Cell::Cell, :7-14
    cx, cy = np.meshgrid(xc + off, yc + off)                             # row i = y, column j = x (:24-27)
    P = np.stack([cx, cy], axis=-1)
    occ = ~not_inside(P, obstacles, 0.0)                                 # :33
    infl = ~occ & ~not_inside(P, obstacles, inflate)                     # :37
    g = np.zeros(cx.shape, dtype=np.int8)
    g[infl] = INFLATION
    g[occ] = OCCUPIED
    return g, float(resolution), np.array([x_min, y_min])


def reference_demo_grid(scale=5.0, resolution=0.06, inflate=0.1):
    """The grid viz_grid publishes on `grid_map` for map/launch/viz_map.launch:52-57 (map.yaml, scale 5, resolution 0.06,
    inflate 0.1) -- BASELINE.json config 4's "map pkg grid"."""
    return build_map(scale_obstacles(MAP_YAML_OBSTACLES, scale), resolution, inflate)


class FakeGrid(object):
    """The robot's incrementally revealed view of the map: Grid::fake_grid starts all Free (grid.cpp:62-66) and
    Grid::update_grid(cc, visibility) copies the true cell types of the (2 v + 1)^2 - 1 neighbours of the current cell into it
    (grid.cpp:155-173, get_neighbours :201-237: the centre cell itself is skipped, cells outside the grid are ignored)."""

    def __init__(self, true_cells):
        self.true = np.asarray(true_cells, dtype=np.int8)
        self.fake = np.zeros_like(self.true)

    def update(self, ix, iy, visibility):
        """Returns the patch (x0, y0, w, h, cells) that changed-or-not region covers: the neighbourhood clipped to the grid."""
        H, W = self.true.shape
        x0, x1 = max(0, ix - visibility), min(W - 1, ix + visibility)
        y0, y1 = max(0, iy - visibility), min(H - 1, iy + visibility)
        keep = self.fake[iy, ix] if (0 <= ix < W and 0 <= iy < H) else None
        self.fake[y0:y1 + 1, x0:x1 + 1] = self.true[y0:y1 + 1, x0:x1 + 1]
        if keep is not None:
            self.fake[iy, ix] = keep                                     # (0, 0) is skipped, :217-220
        return x0, y0, x1 - x0 + 1, y1 - y0 + 1, np.ascontiguousarray(self.fake[y0:y1 + 1, x0:x1 + 1])

    def occupancy(self):
        """Grid::fake_occupancy_grid, grid.cpp:182-199."""
        return self.fake.copy()
