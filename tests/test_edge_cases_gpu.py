"""Edge cases of the particle path on the CUDA side against the oracle: containers that are empty, entirely
dead or hold a single particle; particles sitting exactly on the tile faces (the >= / < tie-breaks of
pic/particle.c++:228-238); three species; a container that outgrows its capacity class while receiving
arrivals; whole laps on the smallest legal tile (3^3 cells)."""
import numpy as np
import pytest

import runko_b200 as rb
from oracle.oracle import OracleGrid
from util import DEAD, assert_bits_equal, pic_conf, random_lattice, random_particles

pytestmark = pytest.mark.gpu


def same_containers(org, t, tile, nsp=2):
    for sp in range(nsp):
        o = org.get_particles(t, sp, alive_only=False)
        g = tile.get_particles(sp, alive_only=False)
        assert_bits_equal(g[6], o[6], f"species {sp} ids")
        alive = o[6] != DEAD
        for c in range(6):
            assert_bits_equal(g[c][alive], o[c][alive], f"species {sp} comp {c}")


def lap_ops(org, t, tile, ops):
    for op in ops:
        org.tile_op(t, op)
        getattr(tile, op)()


@pytest.mark.parametrize("kind", ["empty", "all-dead", "single", "single-dead-tail"])
def test_degenerate_containers(kind):
    rng = np.random.default_rng(1)
    n = (6, 5, 7)
    conf = pic_conf(n_cells=n)
    org, tile = OracleGrid(conf), rb.PicTile((0, 0, 0), conf)
    E, B, J = (random_lattice(rng, n) for _ in range(3))
    org.set_fields(0, E, B, J, with_halo=True)
    tile.set_fields_f32(E, B, J, with_halo=True)
    m = {"empty": 0, "all-dead": 300, "single": 1, "single-dead-tail": 40}[kind]
    for sp in range(2):
        pos, vel, ids = random_particles(rng, m, tile.mins, tile.maxs)
        if kind == "all-dead":
            ids[:] = DEAD
        if kind == "single-dead-tail":
            ids[1:] = DEAD
        org.set_particles(0, sp, *pos, *vel, ids)
        tile.set_particles_raw(sp, *pos, *vel, ids)
    for _ in range(2):
        lap_ops(org, 0, tile, ["push_particles", "pack_outgoing_particles", "sort_particles", "deposit_current"])
        same_containers(org, 0, tile)
        oJ, gJ = org.get_fields(0, with_halo=True)[2], tile.get_fields_f32(with_halo=True)[2]
        assert np.max(np.abs(gJ - oJ)) <= 1e-5 * max(np.max(np.abs(oJ)), 1e-30)
    for sp in range(2):
        ok, on = org.kinetic_energy(0, sp)
        gk, gn = tile.kinetic_energy(sp)
        assert gn == on and abs(gk - ok) <= 1e-9 * max(ok, 1e-30)


def test_particles_exactly_on_tile_faces():
    """x == mins stays (>=), x == maxs leaves (<): every combination of face / edge / corner positions of the
    middle tile of a 3^3 grid, with zero velocity and zero fields so the push leaves them where they are."""
    n = (4, 5, 6)
    conf = pic_conf(n_tiles=(3, 3, 3), n_cells=n)
    org = OracleGrid(conf)
    t = org.cid(1, 1, 1)
    tile = rb.PicTile((1, 1, 1), conf)
    axes = [[tile.mins[d], 0.5 * (tile.mins[d] + tile.maxs[d]), tile.maxs[d],
             np.nextafter(np.float32(tile.maxs[d]), np.float32(-np.inf)), np.nextafter(np.float32(tile.mins[d]), np.float32(-np.inf))]
            for d in range(3)]
    pts = np.array([(x, y, z) for x in axes[0] for y in axes[1] for z in axes[2]], np.float32).T
    m = pts.shape[1]
    for sp in range(2):
        ids = (np.uint64(sp + 1) << np.uint64(40)) | np.arange(m, dtype=np.uint64)
        org.set_particles(t, sp, *pts, *np.zeros((3, m), np.float32), ids)
        tile.set_particles_raw(sp, *pts, *np.zeros((3, m), np.float32), ids)
    lap_ops(org, t, tile, ["push_particles", "pack_outgoing_particles"])
    obuf, oends = org.get_outgoing(t)
    gbuf, gends = tile.get_outgoing()
    assert_bits_equal(gends, oends, "subregion ends")
    for f in ("pos", "vel", "id"):
        assert_bits_equal(gbuf[f], obuf[f], f"outgoing {f}")
    same_containers(org, t, tile)
    stay = int(np.sum(org.get_particles(t, 0, alive_only=False)[6] != DEAD))
    assert stay == 3 ** 3            # mins, the middle and the last float below maxs stay on every axis


def test_three_species_lap():
    rng = np.random.default_rng(3)
    T, n = (2, 1, 2), (6, 6, 6)
    conf = pic_conf(n_tiles=T, n_cells=n, q0=-0.04, q1=0.03, q2=0.01, m0=1.0, m1=2.0, m2=5.0)
    org, grid, tiles = OracleGrid(conf), rb.Grid(conf), {}
    for i in range(T[0]):
        for k in range(T[2]):
            tile = rb.PicTile((i, 0, k), conf)
            t = org.cid(i, 0, k)
            E, B, J = (random_lattice(rng, n, 0.2) for _ in range(3))
            org.set_fields(t, E, B, J, with_halo=True)
            tile.set_fields_f32(E, B, J, with_halo=True)
            for sp in range(3):
                pos, vel, _ = random_particles(rng, 900 + 100 * sp, tile.mins, tile.maxs, u_scale=0.4)
                org.inject(t, sp, *pos.astype(np.float64), *vel.astype(np.float64))
                tile._inject_arrays(sp, pos.astype(np.float64), vel.astype(np.float64))
            grid.add_tile(tile)
            tiles[(i, 0, k)] = tile
    for m in (1, 2):
        org.local_communication(m)
        grid.local_communication(m)
    org.step_pic(0)
    grid.step_pic(0)
    for (i, j, k), tile in tiles.items():
        same_containers(org, org.cid(i, j, k), tile, nsp=3)
        for name, a, b in zip("EBJ", tile.get_fields_f32(with_halo=True), org.get_fields(org.cid(i, j, k), with_halo=True)):
            assert np.max(np.abs(a - b)) <= 1e-5 * np.max(np.abs(b)), name
    ob, oe, ok, on = org.energies()
    gb, ge, gk, gn = grid.energies()
    assert len(gk) == 3 and np.array_equal(on, gn) and np.allclose(gk, ok, rtol=1e-6)


def test_container_outgrows_its_capacity_while_receiving():
    """Tile (0,0,0) starts almost empty, its x-neighbour streams 300k particles into it: the append reallocates
    the receiving container (capacity classes of 256 Ki slots) without losing order, ids or values."""
    rng = np.random.default_rng(4)
    T, n = (2, 1, 1), (8, 8, 8)
    conf = pic_conf(n_tiles=T, n_cells=n, q0=-1e-4, q1=1e-4, current_filter=None)
    org, grid, tiles = OracleGrid(conf), rb.Grid(conf), {}
    for i in range(2):
        tile = rb.PicTile((i, 0, 0), conf)
        grid.add_tile(tile)
        tiles[i] = tile
    m = 300000
    for sp in range(2):
        pos = np.stack([8.0 + 0.3 * rng.random(m), 8.0 * rng.random(m), 8.0 * rng.random(m)])   # just right of the face
        vel = np.stack([-2.0 - rng.random(m), 0.05 * rng.standard_normal(m), 0.05 * rng.standard_normal(m)])
        org.inject(org.cid(1, 0, 0), sp, *pos, *vel)
        tiles[1]._inject_arrays(sp, pos, vel)
        few = np.stack([4.0 + rng.random(10), 4.0 + rng.random(10), 4.0 + rng.random(10)])
        org.inject(org.cid(0, 0, 0), sp, *few, *np.zeros((3, 10)))
        tiles[0]._inject_arrays(sp, few, np.zeros((3, 10)))
    org.step_pic(0)
    grid.step_pic(0)
    for i in range(2):
        same_containers(org, org.cid(i, 0, 0), tiles[i])      # lap 0: E = 0 going in, bit-exact incl. the appended order
    assert tiles[0].container_size(0) > 262144                # it did outgrow the first capacity class
    org.step_pic(1)
    grid.step_pic(1)
    for i in range(2):                                        # lap 1 sees the 1e-7 deposit-order noise of J in E: same slots, ulps
        for sp in range(2):
            o = org.get_particles(org.cid(i, 0, 0), sp, alive_only=False)
            g = tiles[i].get_particles(sp, alive_only=False)
            assert_bits_equal(g[6], o[6], "ids")
            alive = o[6] != DEAD
            for c in range(6):
                assert np.allclose(g[c][alive], o[c][alive], rtol=1e-5, atol=1e-5)


def test_laps_on_the_smallest_tiles():
    rng = np.random.default_rng(5)
    T, n = (3, 2, 2), (3, 3, 3)
    conf = pic_conf(n_tiles=T, n_cells=n, q0=-0.02, q1=0.02)
    org, grid, tiles = OracleGrid(conf), rb.Grid(conf), {}
    for i in range(T[0]):
        for j in range(T[1]):
            for k in range(T[2]):
                tile = rb.PicTile((i, j, k), conf)
                t = org.cid(i, j, k)
                E, B, J = (random_lattice(rng, n, 0.1) for _ in range(3))
                org.set_fields(t, E, B, J, with_halo=True)
                tile.set_fields_f32(E, B, J, with_halo=True)
                for sp in range(2):
                    pos, vel, _ = random_particles(rng, 200, tile.mins, tile.maxs, u_scale=0.5)
                    org.inject(t, sp, *pos.astype(np.float64), *vel.astype(np.float64))
                    tile._inject_arrays(sp, pos.astype(np.float64), vel.astype(np.float64))
                grid.add_tile(tile)
                tiles[(i, j, k)] = tile
    for m in (1, 2):
        org.local_communication(m)
        grid.local_communication(m)
    org.step_pic(0)
    grid.step_pic(0)
    for (i, j, k), tile in tiles.items():
        same_containers(org, org.cid(i, j, k), tile)
    for lap in range(1, 6):
        org.step_pic(lap)
        grid.step_pic(lap)
    n_o = sum(len(org.get_particles(org.cid(*idx), sp)[6]) for idx in tiles for sp in range(2))
    n_g = sum(len(t.get_particles(sp)[6]) for t in tiles.values() for sp in range(2))
    assert n_o == n_g == 2 * 200 * len(tiles)          # nothing lost or duplicated through six laps of migration
