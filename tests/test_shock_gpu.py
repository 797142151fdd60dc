"""Parity of the pic-shock boundary pieces (BASELINE configs[3]) on the CUDA
path against the CPU oracle — which tests/test_oracle_vs_reference_build.py pins bit for bit against
the reference's own pic/reflector_wall.c++ and YeeLattice::apply_edge_bc:

  * edge boundary conditions: bit-exact;
  * reflector wall: positions, momenta, ids (reflected / parked / untouched) bit-exact; the correction
    current within the stated deposit tolerance (atomic accumulation order);
  * a miniature of the shock lap (projects/pic-shock/pic.py:229-279: faraday + stencil +
    binomial2_unrolled + linear_1st_unrolled, conducting wall + upstream edge BCs, moving injector
    through batch_inject_in_x_stripe) over a row of tiles.
"""
import numpy as np
import pytest

import runko_b200 as rb
from oracle.oracle import OracleGrid
from util import DEAD, assert_bits_equal, pic_conf, random_lattice

pytestmark = pytest.mark.gpu

E_, B_, J_ = rb.comm_mode.emf_E, rb.comm_mode.emf_B, rb.comm_mode.emf_J


@pytest.mark.parametrize("case", [
    dict(direction=0, side=0, position=5.0, Ex=1.5, Ey=-2.0, Ez=0.25, E_components=0b110),
    dict(direction=0, side=1, position=7.0, Bx=0.1, By=0.2, Bz=0.3),
    dict(direction=1, side=0, position=15.5, Jx=3.0, Jz=-1.0, J_components=0b101),
    dict(direction=2, side=1, position=13.0, Ex=9.0, Ey=8.0, Ez=7.0),
    dict(direction=0, side=0, position=100.0, Ex=1.0, Ey=1.0, Ez=1.0),
    dict(direction=0, side=0, position=-3.0, Ex=1.0, Ey=1.0, Ez=1.0),
])
def test_edge_bc_bit_exact(case):
    rng = np.random.default_rng(21)
    n = (10, 11, 13)
    conf = pic_conf(n_tiles=(2, 2, 2), n_cells=n)
    org = OracleGrid(conf)
    t = org.cid(0, 1, 1)
    tile = rb.PicTile((0, 1, 1), conf)

    def load():
        E, B, J = (random_lattice(rng, n) for _ in range(3))
        org.set_fields(t, E, B, J, with_halo=True)
        tile.set_fields_f32(E, B, J, with_halo=True)

    def same(what):
        for a, b, nm in zip(tile.get_fields_f32(with_halo=True), org.get_fields(t, with_halo=True), "EBJ"):
            assert_bits_equal(a, b, what + nm)

    load()
    bc = rb.edge_bc(**case)
    for mode in (E_, B_, J_):
        org.apply_edge_bc(t, bc, mode.value)
        tile.apply_edge_bc(bc, mode)
    same("apply_edge_bc ")
    bc2 = rb.edge_bc(direction=1, side=1, position=14.0, Ex=4.0, Ey=4.0, Ez=4.0, Bx=5.0, By=5.0, Bz=5.0, Jx=6.0, Jy=6.0, Jz=6.0)
    for b in (bc, bc2):
        org.register_edge_bc(t, b)
        tile.register_edge_bc(b)
    load()
    for mode in (J_, E_, B_):
        org.apply_edge_bcs(t, mode.value)
        tile.apply_edge_bcs(mode)
    same("apply_edge_bcs ")


def test_edge_bc_rejects_other_modes():
    tile = rb.PicTile((0, 0, 0), pic_conf(n_cells=(5, 5, 5)))
    with pytest.raises(rb.B2PError):
        tile.apply_edge_bc(rb.edge_bc(direction=0, side=0, position=3.0), rb.comm_mode.pic_particle)


def _same_particles(org, t, tile):
    for sp in range(2):
        o = org.get_particles(t, sp, alive_only=False)
        g = tile.get_particles(sp, alive_only=False)
        assert_bits_equal(g[6], o[6], f"species {sp} ids")
        alive = o[6] != DEAD
        for c in range(6):
            assert_bits_equal(g[c][alive], o[c][alive], f"species {sp} comp {c}")


@pytest.mark.parametrize("beta", [0.0, 0.3])
@pytest.mark.parametrize("fuse", [1, 0])
def test_reflector_wall_parity(beta, fuse):
    """push -> reflect -> deposit -> advance over several laps: particles bit-exact (incl. the parked ones
    turning dead), J — zigzag of the post-reflection particles + the correction lattice — within 1e-5 max|J|.
    The reflected tile must not adopt the current its fused push deposited before the reflection."""
    from runko_b200._lib import check
    check(rb.lib().b2p_set_option(b"fuse_deposit", fuse))
    try:
        rng = np.random.default_rng(31)
        n = (12, 6, 7)
        conf = pic_conf(n_tiles=(2, 1, 1), n_cells=n, q0=-0.7, q1=0.4, cfl=0.45)
        org = OracleGrid(conf)
        t = org.cid(0, 0, 0)
        tile = rb.PicTile((0, 0, 0), conf)
        E, B, J = (random_lattice(rng, n) for _ in range(3))
        org.set_fields(t, E, B, J, with_halo=True)
        tile.set_fields_f32(E, B, J, with_halo=True)
        for sp in range(2):
            m = 6000 + 13 * sp
            pos = np.stack([5.0 + 6.0 * rng.random(m), 6.0 * rng.random(m), 7.0 * rng.random(m)]).astype(np.float32)
            vel = (0.6 * rng.standard_normal((3, m))).astype(np.float32)
            vel[0] -= np.float32(0.8)
            pos[0, :40] = (2.5 + 2.0 * rng.random(40)).astype(np.float32)   # already behind the wall: parked (-> dead)
            ids = (np.uint64(sp + 1) << np.uint64(40)) | np.arange(m, dtype=np.uint64)
            ids[rng.random(m) < 0.05] = DEAD
            org.set_particles(t, sp, *pos, *vel, ids)
            tile.set_particles_raw(sp, *pos, *vel, ids)
        gamma = 1.0 / np.sqrt(1.0 - beta * beta)
        for w in (rb.reflector_wall(walloc=5.0, betawall=beta, gammawall=gamma), rb.reflector_wall(walloc=40.0)):
            org.register_reflector_wall(t, w)
            tile.register_reflector_wall(w)
        dead0 = sum(int(np.sum(org.get_particles(t, sp, alive_only=False)[6] == DEAD)) for sp in range(2))
        for lap in range(4):
            for op in ("push_particles", "reflect_particles"):
                org.tile_op(t, op)
                getattr(tile, op)()
            _same_particles(org, t, tile)
            org.tile_op(t, "deposit_current")
            tile.deposit_current()
            oJ = org.get_fields(t, with_halo=True)[2]
            gJ = tile.get_fields_f32(with_halo=True)[2]
            assert np.max(np.abs(gJ - oJ)) <= 1e-5 * np.max(np.abs(oJ)), f"lap {lap}"
            org.tile_op(t, "advance_reflector_walls")
            tile.advance_reflector_walls()
            assert tile.reflector_walls() == org.reflector_walls(t)
        dead1 = sum(int(np.sum(org.get_particles(t, sp, alive_only=False)[6] == DEAD)) for sp in range(2))
        assert dead1 > dead0                                   # some particles were parked
        assert np.any(org.get_particles(t, 0)[3] > 0.5)        # and some reflected
    finally:
        check(rb.lib().b2p_set_option(b"fuse_deposit", 1))


class _OracleTileFace:
    """the slice of the tile API MovingInjector / batch_inject_in_x_stripe need, on the oracle"""

    def __init__(self, org, t, tile):
        self.org, self.t, self.ref = org, t, tile

    def batch_inject_in_x_stripe(self, sp, pgen, x_left, x_right):
        # the host logic is the product's own (runko_b200.tiles.PicTileHost); only the container differs
        org, t = self.org, self.t

        class Shim(rb.tiles.PicTileHost):
            mins, maxs, n_cells = self.ref.mins, self.ref.maxs, self.ref.n_cells
            global_coordinate_map = self.ref.global_coordinate_map

            def _backend_inject(self_inner, sp_, a):
                org.inject(t, sp_, *a)

        Shim().batch_inject_in_x_stripe(sp, pgen, x_left, x_right)


def test_shock_lap_parity():
    """Miniature of BASELINE configs[3] on a 4x1x2 row of tiles: wall at x = 7 with the conducting edge BC behind
    it, upstream edge BC near the right end, plasma drifting in -x, injector stripe moving in +x."""
    rng = np.random.default_rng(41)
    n_tiles, n = (4, 1, 2), (8, 6, 5)
    kw = {}
    for ax in "xyz":
        kw[f"stencil_{ax}_delta"] = 0.02
        kw[f"stencil_{ax}_beta_p1"] = -0.01
    conf = pic_conf(n_tiles=n_tiles, n_cells=n, cfl=0.45, q0=-0.05, q1=0.05, particle_pusher="faraday",
                    field_propagator="stencil", current_filter="binomial2_unrolled", field_interpolator="linear_1st_unrolled",
                    prealloc_per_species=4000, **kw)
    Lx = n_tiles[0] * n[0]
    walloc, beta = 7.0, 0.5
    wall = rb.reflector_wall(walloc=walloc)
    cbc = rb.edge_bc(direction=0, side=0, position=walloc, E_components=0b110, B_components=0, J_components=0b111)
    ubc = rb.edge_bc(direction=0, side=1, position=Lx - 5.0, Ex=0.0, Ey=0.01, Ez=-0.02, Bx=0.0, By=0.04, Bz=0.02,
                     J_components=0b111)
    org = OracleGrid(conf)
    grid = rb.Grid(conf)
    tiles, faces = {}, []
    for i in range(n_tiles[0]):
        for k in range(n_tiles[2]):
            tile = rb.PicTile((i, 0, k), conf)
            t = org.cid(i, 0, k)
            E, B, J = (random_lattice(rng, n, 0.02) for _ in range(3))
            org.set_fields(t, E, B, J, with_halo=True)
            tile.set_fields_f32(E, B, J, with_halo=True)
            for obj_reg_w, obj_reg_bc in ((lambda w: org.register_reflector_wall(t, w), lambda b: org.register_edge_bc(t, b)),
                                          (tile.register_reflector_wall, tile.register_edge_bc)):
                obj_reg_w(wall)
                obj_reg_bc(cbc)
                obj_reg_bc(ubc)
            grid.add_tile(tile)
            tiles[(i, 0, k)] = tile
            faces.append((tile, _OracleTileFace(org, t, tile)))

    injloc0 = walloc + 9.0

    def make_pgen(seed):
        g = np.random.default_rng(seed)

        def pgen(x, y, z):
            m = len(x)
            pos = (x + g.random(m), y + g.random(m), z + g.random(m))
            vel = (-0.6 + 0.1 * g.standard_normal(m), 0.1 * g.standard_normal(m), 0.1 * g.standard_normal(m))
            return rb.ParticleStateBatch(pos=pos, vel=vel)
        return pgen

    # identical generator streams for both sides
    for side in (0, 1):
        for sp in range(2):
            pg = make_pgen(100 + sp)
            for tile, face in faces:
                for _ in range(2):
                    (tile if side == 0 else face).batch_inject_in_x_stripe(sp, pg, walloc, injloc0)
    # the injection front of projects/pic-shock (runko/moving_injector.py): every lap > 0 the stripe between the drifted
    # previous front and the advanced front is filled, until the front nears the right end of the box
    def stripes(injloc, stride, margin=10.0):
        lap = 0
        while True:
            if lap > 0:
                left, right = max(injloc - beta * stride, walloc), injloc + 1.0 * stride
                if right >= Lx - margin:
                    return
                injloc = right
                yield lap, left, right, injloc
            else:
                yield lap, None, None, injloc
            lap += 1
    fronts = [stripes(injloc0, conf.cfl) for _ in range(2)]
    last_front = [injloc0, injloc0]
    sides = [[tf[0] for tf in faces], [tf[1] for tf in faces]]
    pg_lap = [[make_pgen(200 + sp) for sp in range(2)] for _ in range(2)]

    def oracle_lap(lap, passes=4):
        ph, lc = org.phase, org.local_communication
        ph("push_half_b"); ph("apply_edge_bcs_B"); lc(2)
        ph("push_particles"); ph("reflect_particles"); ph("pack_outgoing_particles"); lc(3)
        if lap % 5 == 0:
            ph("sort_particles")
        ph("deposit_current"); lc(6); lc(0); ph("apply_edge_bcs_J")
        for i in range(passes):
            if i > 0 and i % 3 == 0:
                lc(0)
            ph("filter_current")
        ph("apply_edge_bcs_J")
        ph("push_half_b"); ph("apply_edge_bcs_B"); lc(2)
        ph("push_e"); ph("apply_edge_bcs_E"); ph("add_current"); ph("apply_edge_bcs_E"); lc(1)
        ph("advance_reflector_walls")

    for m in (1, 2):
        org.local_communication(m)
        grid.local_communication(m)
    for lap in range(6):
        oracle_lap(lap)
        grid.step_shock(lap, n_filter_passes=4)
        for side in (0, 1):
            step = next(fronts[side], None)
            if step is None or step[1] is None:
                continue
            _, x_left, x_right, last_front[side] = step
            for tile in sides[side]:
                for sp in range(2):
                    tile.batch_inject_in_x_stripe(sp, pg_lap[side][sp], x_left, x_right)
        if lap == 0:
            for (i, j, k), tile in tiles.items():
                _same_particles(org, org.cid(i, j, k), tile)
    assert last_front[0] == last_front[1] and last_front[0] > injloc0
    n_total = 0
    for (i, j, k), tile in tiles.items():
        t = org.cid(i, j, k)
        for sp in range(2):
            o = org.get_particles(t, sp)
            g = tile.get_particles(sp)
            common, oi, gi = np.intersect1d(o[6], g[6], return_indices=True)
            assert len(common) >= 0.995 * len(o[6])
            n_total += len(common)
            if len(common) == 0:
                assert len(g[6]) == 0
                continue
            for c in range(6):
                d = g[c][gi] - o[c][oi]
                assert np.sqrt(np.mean(d * d)) <= 1e-3 * max(np.sqrt(np.mean(o[c][oi] ** 2)), 1e-2), (i, k, sp, c)
        for name, a, b in zip("EBJ", tile.get_fields_f32(with_halo=True), org.get_fields(t, with_halo=True)):
            assert np.max(np.abs(a - b)) <= 1e-3 * max(np.max(np.abs(b)), 1e-30), (name, i, k)
            if i == 0 and name == "E":
                # conducting half-space behind the wall (interior cells; the x-halo is the periodic
                # neighbour's): Ey = Ez = 0 exactly on both
                assert not np.any(a[1:, 3:3 + 7, 3:-3, 3:-3]) and not np.any(b[1:, 3:3 + 7, 3:-3, 3:-3])
    assert n_total > 1000
