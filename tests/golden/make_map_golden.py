"""Generates tests/golden/map_grid_*.npz from the reference's OWN map::Grid (oracle/_ref/libmapref.so, compiled from
/root/reference/map/src/map/{map,grid,prm}.cpp by oracle/build_map_ref.py).  Run in the build container:

    python -m oracle.build_map_ref && python tests/golden/make_map_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import map_grid, map_ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # (a) the demo grid of map/launch/viz_map.launch:52-57 -- BASELINE config 4's grid
    g, res, org = map_ref.build(map_grid.MAP_YAML_OBSTACLES, 5.0, 0.06, 0.1)
    # (b) the planner demos' grid, global_planner/launch/incremental.launch:56-65 (scale 10, resolution 0.1, inflate 0.1)
    g2, res2, org2 = map_ref.build(map_grid.MAP_YAML_OBSTACLES, 10.0, 0.1, 0.1)
    # (c) incremental reveal along a diagonal path with visibility 2 (Grid::update_grid, grid.cpp:155-173)
    path = np.array([[5 + i, 8 + i] for i in range(20)], dtype=np.int32)
    fake = map_ref.reveal(map_grid.MAP_YAML_OBSTACLES, 10.0, 0.1, 0.1, path, 2)
    np.savez_compressed(os.path.join(HERE, "map_grid_reference.npz"), demo=g, demo_res=res, demo_origin=org,
                        planner=g2, planner_res=res2, planner_origin=org2, reveal_path=path, reveal_visibility=2, reveal_fake=fake)
    print("demo", g.shape, "planner", g2.shape, "revealed cells", int((fake != 0).sum()))


if __name__ == "__main__":
    main()
