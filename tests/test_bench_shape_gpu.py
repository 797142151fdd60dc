"""Parity AT THE BENCHMARKED SHAPE: tiles of 64^3 cells, 2 species x 16 ppc (the density at which the counting sort's
batch / walk paths and the deposit's warp aggregation actually run in bench.py), 2x1x1 periodic tiles, the lap of
projects/pic-turbulence/pic.py phase by phase against the multi-threaded CPU oracle:

  lap 0   push -> positions / momenta bit-exact; pack -> subregion ends and the AoS payload bit-exact; exchange + append
          -> containers bit-exact slot by slot; sort (freshly injected, unsorted input) -> bit-exact order; deposit +
          field phase -> B bit-exact, E / J within 1e-5
  laps 1-4 whole laps on both sides (the containers drift away from the sorted order)
  lap 5   push, pack, exchange, then the GPU containers are re-seeded with the oracle's so that both sides sort
          IDENTICAL drifted input at mean 16 particles per cell -> bit-exact order again; fields within 1e-3.
"""
import os

import numpy as np
import pytest

import runko_b200 as rb
from oracle.oracle import OracleGrid
from util import DEAD, assert_bits_equal, pic_conf

pytestmark = pytest.mark.gpu


def _juttner(rng, n, theta=0.3):
    u = np.empty(n)
    todo = np.arange(n)
    while todo.size:
        x = rng.random((4, todo.size))
        uu, eta = -theta * np.log(x[0] * x[1] * x[2]), -theta * np.log(x[0] * x[1] * x[2] * x[3])
        ok = eta * eta - uu * uu > 1.0
        u[todo[ok]] = uu[ok]
        todo = todo[~ok]
    mu, phi = 2.0 * rng.random(n) - 1.0, 2.0 * np.pi * rng.random(n)
    st = np.sqrt(np.maximum(0.0, 1.0 - mu * mu))
    return np.stack([u * st * np.cos(phi), u * st * np.sin(phi), u * mu])


def _compare_containers(org, tiles):
    for (i, j, k), tile in tiles.items():
        t = org.cid(i, j, k)
        for sp in range(2):
            o = org.get_particles(t, sp, alive_only=False)
            g = tile.get_particles(sp, alive_only=False)
            assert_bits_equal(g[6], o[6], f"tile {(i, j, k)} species {sp} ids (slot by slot)")
            alive = o[6] != DEAD
            for c in range(6):
                assert_bits_equal(g[c][alive], o[c][alive], f"tile {(i, j, k)} species {sp} comp {c}")


def _fields_close(org, tiles, which, tol):
    for (i, j, k), tile in tiles.items():
        o = org.get_fields(org.cid(i, j, k), with_halo=True)
        g = tile.get_fields_f32(with_halo=True)
        for name, a, b in zip("EBJ", g, o):
            if name in which:
                if tol == 0:
                    assert_bits_equal(a, b, f"{name} {(i, j, k)}")
                else:
                    assert np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(b)), 1e-30), (name, (i, j, k))


def test_bench_shape_lap_parity():
    n, ppc, cfl = 64, 16, 0.45
    oppc = 2 * ppc
    q0 = -(cfl ** 2) / (0.5 * oppc * 2.0)                         # bench.py make_conf
    conf = pic_conf(n_tiles=(2, 1, 1), n_cells=(n, n, n), cfl=cfl, q0=q0, q1=abs(q0), current_filter="binomial2")
    binit = float(np.sqrt((1.0 + 1.5 * 0.3) * oppc * abs(q0) * cfl ** 2 * 10.0))
    rng = np.random.default_rng(5)
    org = OracleGrid(conf)
    grid = rb.Grid(conf)
    tiles = {}
    ii, jj, kk = np.meshgrid(*(np.arange(n, dtype=np.float64),) * 3, indexing="ij")
    corner = np.stack([ii.ravel(), jj.ravel(), kk.ravel()])
    B = np.zeros((3, n + 6, n + 6, n + 6), np.float32)
    B[2] = binit
    for i in range(2):
        tile = rb.PicTile((i, 0, 0), conf)
        t = org.cid(i, 0, 0)
        org.set_fields(t, None, B, None, with_halo=True)
        tile.set_fields_f32(None, B, None, with_halo=True)
        # slot p = round * Ncells + cell (pic.py:141-156 / k_inject_thermal): an UNSORTED container, species 1 on top of species 0
        pos = np.concatenate([corner + rng.random(corner.shape) for _ in range(ppc)], axis=1) + np.array([i * n, 0, 0], np.float64)[:, None]
        for sp in range(2):
            vel = _juttner(rng, pos.shape[1])
            org.inject(t, sp, *pos, *vel)
            tile._inject_arrays(sp, pos, vel)
        grid.add_tile(tile)
        tiles[(i, 0, 0)] = tile
    threads = max(2, os.cpu_count() or 2)
    M = rb.comm_mode
    for m in (M.emf_E, M.emf_B):
        org.local_communication(m.value)
        grid.local_communication(m)

    def migrate():
        for name in ("push_half_b",):
            org.phase(name, threads=threads); grid.phase(name)
        org.local_communication(M.emf_B.value); grid.local_communication(M.emf_B)
        org.phase("push_particles", threads=threads); grid.phase("push_particles")
        _compare_containers(org, tiles)                            # the push: bit-exact positions / momenta
        org.phase("pack_outgoing_particles", threads=threads); grid.phase("pack_outgoing_particles")
        for (i, j, k), tile in tiles.items():
            ob, oe = org.get_outgoing(org.cid(i, j, k))
            gb, ge = tile.get_outgoing()
            assert_bits_equal(ge, oe, "subregion_particle_ends_")
            for f in ("pos", "vel", "id"):
                assert_bits_equal(gb[f], ob[f], "outgoing " + f)
        org.local_communication(M.pic_particle.value); grid.local_communication(M.pic_particle)
        _compare_containers(org, tiles)                            # append + periodic wrap

    def rest_of_lap():
        org.phase("deposit_current", threads=threads); grid.phase("deposit_current")
        for m in (M.emf_J_exchange, M.emf_J):
            org.local_communication(m.value); grid.local_communication(m)
        org.phase("filter_current", threads=threads); grid.phase("filter_current")
        org.local_communication(M.emf_J.value); grid.local_communication(M.emf_J)
        for _ in range(2):
            org.phase("filter_current", threads=threads); grid.phase("filter_current")
        org.phase("push_half_b", threads=threads); grid.phase("push_half_b")
        org.local_communication(M.emf_B.value); grid.local_communication(M.emf_B)
        org.phase("push_e", threads=threads); grid.phase("push_e")
        org.phase("add_current", threads=threads); grid.phase("add_current")
        org.local_communication(M.emf_E.value); grid.local_communication(M.emf_E)

    # ---- lap 0
    migrate()
    org.phase("sort_particles", threads=threads); grid.phase("sort_particles")
    _compare_containers(org, tiles)                                # sort of unsorted input at 16 ppc
    rest_of_lap()
    _fields_close(org, tiles, "B", 0)
    _fields_close(org, tiles, "EJ", 1e-5)
    # ---- laps 1-4
    for lap in range(1, 5):
        org.step_pic(lap, threads=threads)
        grid.step_pic(lap)
    # ---- lap 5: the sort of a container that drifted for five laps
    for name in ("push_half_b",):
        org.phase(name, threads=threads); grid.phase(name)
    org.local_communication(M.emf_B.value); grid.local_communication(M.emf_B)
    for name in ("push_particles", "pack_outgoing_particles"):
        org.phase(name, threads=threads); grid.phase(name)
    org.local_communication(M.pic_particle.value); grid.local_communication(M.pic_particle)
    n_alive = 0
    for (i, j, k), tile in tiles.items():                          # identical drifted input on both sides
        for sp in range(2):
            o = org.get_particles(org.cid(i, j, k), sp, alive_only=False)
            g_ids = tile.get_ids(sp)
            assert abs(len(g_ids) - int(np.count_nonzero(o[6] != DEAD))) <= 64   # same population up to rare boundary flips (1e-7 J noise)
            tile.set_particles_raw(sp, *o)
            n_alive += int(np.count_nonzero(o[6] != DEAD))
    assert n_alive == 2 * 2 * ppc * n ** 3                          # nothing lost over five laps of migration
    org.phase("sort_particles", threads=threads); grid.phase("sort_particles")
    _compare_containers(org, tiles)
    keys = tiles[(0, 0, 0)].sort_keys(0).astype(np.int64)
    assert np.all(np.diff(keys) >= 0)
    rest_of_lap()
    _fields_close(org, tiles, "EBJ", 1e-3)
