"""debug: per-lap launch counts by kernel class for the shock workload"""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench, runko_b200 as rb
from runko_b200._lib import check
L = rb.lib(); check(L.b2p_init(0))
conf, tpg, gb, desc, ns, ppc = bench.named_workload("shock", 1, 512)
tile = conf.n_cells_per_tile
grid = rb.Grid(conf); Lx = conf.n_tiles[0] * tile[0]
shock = bench.ShockState(conf, Lx)
wall = rb.reflector_wall(walloc=shock.walloc)
tiles = []
for i in range(tpg[0]):
    for j in range(tpg[1]):
        for k in range(tpg[2]):
            t = rb.PicTile((i, j, k), conf); t.register_reflector_wall(wall); grid.add_tile(t); tiles.append(t)
for sp in range(2):
    grid.inject_drifting_stripe(sp, 4, 1e-5, 5.0, -1, shock.walloc, shock.injloc, seed=42)
nk = L.b2p_profile_num_classes(); names = [L.b2p_profile_class_name(k).decode() for k in range(nk)]
for lap in range(12):
    check(L.b2p_profile_enable(1))
    grid.step_shock(lap, n_filter_passes=4); shock.inject_front(grid, lap, conf.cfl, 1000)
    a, b, c = np.zeros(nk), np.zeros(nk, np.uint64), np.zeros(nk)
    check(L.b2p_profile_report(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p)))
    check(L.b2p_profile_enable(0))
    print(lap, {names[k]: int(b[k]) for k in range(nk) if b[k]}, flush=True)
