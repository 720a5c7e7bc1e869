"""CPU: pin oracle/map_grid.py (NumPy restatement of the map package's occupancy grid: map/src/map/grid.cpp:17-69,126-237,
map/src/map/prm.cpp:267-394,422-461, map/src/map/map.cpp:47-76) against the reference's OWN sources -- the committed golden
grids made from oracle/_ref/libmapref.so (tests/golden/make_map_golden.py) and, when that library is present, live."""
import os

import numpy as np
import pytest

from oracle import map_grid, map_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "map_grid_reference.npz")


def test_demo_and_planner_grids_match_golden():
    g = np.load(GOLD)
    cells, res, org = map_grid.reference_demo_grid()                       # map/launch/viz_map.launch:52-57
    assert cells.shape == g["demo"].shape == (161, 114)                    # 9.6 / 0.06 accumulates to 161 rows, not 160
    assert np.array_equal(cells, g["demo"]) and res == float(g["demo_res"]) and np.array_equal(org, g["demo_origin"])
    assert set(np.unique(cells)) == {0, 50, 100}                           # grid.cpp:126-144
    cells2, _, _ = map_grid.build_map(map_grid.scale_obstacles(map_grid.MAP_YAML_OBSTACLES, 10.0), 0.1, 0.1)
    assert np.array_equal(cells2, g["planner"])                            # global_planner/launch/incremental.launch:56-65


def test_incremental_reveal_matches_golden():
    g = np.load(GOLD)
    fg = map_grid.FakeGrid(g["planner"])
    assert not fg.occupancy().any()                                        # fake grid starts all Free, grid.cpp:62-66
    for ix, iy in g["reveal_path"]:
        x0, y0, w, h, patch = fg.update(int(ix), int(iy), int(g["reveal_visibility"]))
        assert patch.shape == (h, w)
    assert np.array_equal(fg.occupancy(), g["reveal_fake"])


@pytest.mark.skipif(not map_ref.available(), reason="oracle/_ref/libmapref.so not built here")
@pytest.mark.parametrize("seed", range(8))
def test_restatement_matches_live_reference_on_random_maps(seed):
    """random convex polygons + walls, random scale / resolution / inflation: every cell equals the reference's."""
    rng = np.random.RandomState(seed)
    obstacles = []
    for _ in range(rng.randint(2, 7)):
        c = rng.uniform(4, 40, size=2)
        n = rng.randint(3, 7)
        ang = np.sort(rng.uniform(0, 2 * np.pi, size=n))
        r = rng.uniform(1.5, 6.0)
        obstacles.append([[float(c[0] + r * np.cos(a)), float(c[1] + r * np.sin(a))] for a in ang])   # CCW
    obstacles += [[[0., 0.], [44., 0.]], [[44., 0.], [44., 44.]], [[44., 44.], [0., 44.]], [[0., 44.], [0., 0.]]]
    scale = float(rng.choice([4.0, 5.0, 10.0]))
    res = float(rng.choice([0.05, 0.06, 0.1, 0.13]))
    inflate = float(rng.choice([0.0, 0.05, 0.1, 0.25]))
    want, _, worg = map_ref.build(obstacles, scale, res, inflate)
    got, _, gorg = map_grid.build_map(map_grid.scale_obstacles(obstacles, scale), res, inflate)
    assert got.shape == want.shape and np.array_equal(gorg, worg)
    assert np.array_equal(got, want), "%d cells differ" % int((got != want).sum())
    path = np.stack([rng.randint(0, want.shape[1], size=12), rng.randint(0, want.shape[0], size=12)], axis=1)
    vis = int(rng.randint(1, 4))
    fake = map_ref.reveal(obstacles, scale, res, inflate, path, vis)
    fg = map_grid.FakeGrid(got)
    for ix, iy in path:
        fg.update(int(ix), int(iy), vis)
    assert np.array_equal(fg.occupancy(), fake)
