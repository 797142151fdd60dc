"""Known-answer / behavioural tests restated from the reference's own suites
(tests/py/test_emf_fdtd2.py, test_emf_current_filter_binomial2.py,
test_pic_particle_pusher.py, test_pic_current_depositer_zigzag_1st*.py,
test_pic_particle_sorting.py, tests/py-multirank/test_{emf,pic}_simulation.py) so that they
run on BOTH backends: the CPU oracle here and the CUDA path on the B200 box, where
/root/reference does not exist.  Same shapes (10x11x13 tiles), same assertions."""
import itertools

import numpy as np
import pytest

from backends import GRID, TILE
from util import DEAD, emf_conf, pic_conf

N = (10, 11, 13)


def lattice_from(fn, n=N, stagger="E", origin=(0, 0, 0)):
    """emf/tile.c++:198-208: sample fn at the Yee-staggered points of the whole haloed lattice"""
    H = [v + 6 for v in n]
    i, j, k = np.meshgrid(*(np.arange(h, dtype=np.float64) - 3 + o for h, o in zip(H, origin)), indexing="ij")
    h = 0.5
    if stagger == "E":
        pts = ((i + h, j, k), (i, j + h, k), (i, j, k + h))
    else:
        pts = ((i, j + h, k + h), (i + h, j, k + h), (i + h, j + h, k))
    out = np.zeros((3,) + tuple(H), np.float32)
    for c in range(3):
        out[c] = np.broadcast_to(np.asarray(fn(*pts[c])[c], dtype=np.float64), i.shape)
    return out


ZERO = np.zeros((3,) + tuple(v + 6 for v in N), np.float32)
INNER = (slice(None), slice(4, -4), slice(4, -4), slice(4, -4))   # reference tests look at [1:-1] of the interior


# ---- tests/py/test_emf_fdtd2.py:44-186 ---------------------------------------------------------
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_push_half_b_of_linear_E(backend, axis):
    fns = [lambda x, y, z: (0 * x, z, -y), lambda x, y, z: (z, 0 * x, -x), lambda x, y, z: (y, -x, 0 * x)]
    sign = [1, -1, 1][axis]
    t = TILE[backend](emf_conf(n_cells=N, cfl=1))
    t.set_fields(lattice_from(fns[axis], stagger="E"), ZERO, ZERO)
    t.op("push_half_b")
    B = t.get_fields()[1][INNER]
    b = B[axis].flat[0]
    assert sign * b > 0
    assert np.allclose(B[axis], b, atol=1e-5)
    for c in range(3):
        if c != axis:
            assert np.allclose(B[c], 0, atol=1e-5)


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_push_e_of_linear_B(backend, axis):
    fns = [lambda x, y, z: (0 * x, z, -y), lambda x, y, z: (z, 0 * x, -x), lambda x, y, z: (y, -x, 0 * x)]
    sign = [-1, 1, -1][axis]
    t = TILE[backend](emf_conf(n_cells=N, cfl=1))
    t.set_fields(ZERO, lattice_from(fns[axis], stagger="B"), ZERO)
    t.op("push_e")
    E = t.get_fields()[0][INNER]
    e = E[axis].flat[0]
    assert sign * e > 0
    assert np.allclose(E[axis], e, atol=1e-5)


def test_zero_coefficient_stencil_is_fdtd2(backend):
    # tests/py/test_emf_stencil.py: stencil with all extra coefficients zero == fdtd2 to 1e-6
    rng = np.random.default_rng(0)
    E = rng.standard_normal(ZERO.shape).astype(np.float32)
    out = []
    for prop in ("fdtd2", "stencil"):
        t = TILE[backend](emf_conf(n_cells=N, cfl=0.45, field_propagator=prop))
        t.set_fields(E, ZERO, ZERO)
        t.op("push_half_b")
        out.append(t.get_fields()[1])
    assert np.allclose(out[0], out[1], atol=1e-6)


# ---- tests/py/test_emf_current_filter_binomial2.py:49-116 -----------------------------------------
@pytest.mark.parametrize("variant", ["binomial2", "binomial2_unrolled"])
def test_filter_constant_and_checkerboard(backend, variant):
    t = TILE[backend](emf_conf(n_cells=N, current_filter=variant))
    J = np.empty_like(ZERO)
    J[0], J[1], J[2] = 1.0, 2.0, 3.0
    t.set_fields(ZERO, ZERO, J)
    t.op("filter_current")
    out = t.get_fields()[2]
    inner = (slice(None),) + (slice(1, -1),) * 3
    assert np.allclose(out[inner], J[inner], atol=1e-5)            # constant J is invariant
    i, j, k = np.meshgrid(*(np.arange(v + 6) for v in N), indexing="ij")
    cb = ((i + j + k) % 2).astype(np.float32)
    t.set_fields(ZERO, ZERO, np.stack([cb, cb, cb]))
    t.op("filter_current")
    out = t.get_fields()[2][inner]
    assert np.all(out > 0) and np.all(out < 1)                     # checkerboard is smoothed strictly into (0, 1)
    assert np.allclose(out, 0.5, atol=1e-6)                        # (1/4,1/2,1/4)^3 kills the Nyquist mode


# ---- tests/py/test_pic_particle_pusher.py ------------------------------------------------------------
def one_particle(t, sp, pos, vel):
    for s in range(2):
        n = 1 if s == sp else 0
        a = [np.full(n, v, np.float32) for v in (*pos, *vel)]
        t.set_particles(s, *a, np.arange(n, dtype=np.uint64))


@pytest.mark.parametrize("pusher", ["boris", "higuera_cary", "faraday"])
@pytest.mark.parametrize("interp", ["linear_1st", "linear_1st_unrolled"])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_uniform_E_accelerates_along_the_field(backend, pusher, interp, axis):
    cfl = 0.45
    conf = pic_conf(n_cells=N, cfl=cfl, particle_pusher=pusher, field_interpolator=interp, q0=-1.0, m0=1.0, q1=1.0, m1=2.0)
    t = TILE[backend](conf)
    E = np.zeros_like(ZERO)
    E[axis] = 0.01
    t.set_fields(E, ZERO, ZERO)
    for sp, qm in ((0, -1.0), (1, 0.5)):
        one_particle(t, sp, (5.3, 5.6, 6.1), (0.0, 0.0, 0.0))
        t.op("push_particles")
        p = t.get_particles(sp)
        u = np.array([p[3][0], p[4][0], p[5][0]])
        want = np.zeros(3)
        want[axis] = qm * 0.01 / cfl                 # B = 0: u1 = u0 + (q/m) E / c   (pic/particle_boris.h:39-53)
        assert np.allclose(u, want, rtol=1e-4, atol=1e-7), (pusher, u, want)
        x = np.array([p[0][0], p[1][0], p[2][0]])
        assert np.sign(x[axis] - (5.3, 5.6, 6.1)[axis]) == np.sign(qm)


@pytest.mark.parametrize("pusher", ["boris", "higuera_cary", "faraday"])
def test_uniform_B_rotates_without_changing_energy(backend, pusher):
    conf = pic_conf(n_cells=N, particle_pusher=pusher, q0=-1.0, q1=1.0)
    t = TILE[backend](conf)
    B = np.zeros_like(ZERO)
    B[2] = 0.05
    t.set_fields(ZERO, B, ZERO)
    one_particle(t, 1, (5.3, 5.6, 6.1), (0.8, 0.0, 0.1))
    t.op("push_particles")
    p = t.get_particles(1)
    u = np.array([p[3][0], p[4][0], p[5][0]], np.float64)
    assert abs(np.dot(u, u) - (0.8 ** 2 + 0.1 ** 2)) < 1e-5
    assert u[1] < 0 and abs(u[2] - 0.1) < 1e-6       # q > 0, v = +x, B = +z  =>  F = q v x B = -y


@pytest.mark.parametrize("pusher", ["boris", "higuera_cary", "faraday"])
def test_push_is_translation_invariant_across_tiles(backend, pusher):
    """tests/py/test_pic_particle_pusher.py: same local state in tile (0,0,0) and (3,3,3) -> same momenta"""
    rng = np.random.default_rng(3)
    E, B = (0.2 * rng.standard_normal(ZERO.shape).astype(np.float32) for _ in range(2))
    out = []
    for idx in ((0, 0, 0), (3, 3, 3)):
        conf = pic_conf(n_tiles=(4, 4, 4), n_cells=N, particle_pusher=pusher)
        t = TILE[backend](conf, idx)
        t.set_fields(E, B, ZERO)
        one_particle(t, 0, tuple(t.mins[d] + (4.3, 5.6, 6.1)[d] for d in range(3)), (0.3, -0.2, 0.5))
        t.op("push_particles")
        p = t.get_particles(0)
        out.append([p[3][0], p[4][0], p[5][0], p[0][0] - t.mins[0], p[1][0] - t.mins[1], p[2][0] - t.mins[2]])
    assert np.allclose(out[0][:3], out[1][:3], atol=1e-6) and np.allclose(out[0][3:], out[1][3:], atol=1e-4)


# ---- tests/py/test_pic_current_depositer_zigzag_1st{,_atomic}.py ------------------------------------------
@pytest.mark.parametrize("dep", ["zigzag_1st", "zigzag_1st_atomic"])
def test_deposit_known_answers(backend, dep):
    conf = pic_conf(n_cells=N, current_depositer=dep, q0=-0.7, q1=0.7)
    t = TILE[backend](conf)
    t.set_fields(ZERO, ZERO, ZERO)
    # stationary particle -> exactly zero current
    one_particle(t, 0, (5.3, 5.6, 6.1), (0, 0, 0))
    t.op("deposit_current")
    assert not np.any(t.get_fields()[2])
    # single-axis velocity -> only that component, sign = sign(q v); total current = q * displacement
    for axis, v in itertools.product(range(3), (0.5, -0.5)):
        vel = [0.0, 0.0, 0.0]
        vel[axis] = v
        one_particle(t, 0, (5.3, 5.6, 6.1), vel)
        t.op("deposit_current")
        J = t.get_fields()[2].astype(np.float64)
        for c in range(3):
            if c != axis:
                assert not np.any(J[c])
        disp = 0.45 * v / np.sqrt(1 + v * v)
        assert np.isclose(J[axis].sum(), -0.7 * disp, rtol=1e-5)
    # +q and -q on the same trajectory cancel
    n = 50
    rng = np.random.default_rng(5)
    pos = (3 + 6 * rng.random((3, n))).astype(np.float32)
    vel = rng.standard_normal((3, n)).astype(np.float32)
    for sp in range(2):
        t.set_particles(sp, *pos, *vel, np.arange(n, dtype=np.uint64))
    t.op("deposit_current")
    J = t.get_fields()[2]
    assert np.max(np.abs(J)) < 1e-6


def test_deposit_conserves_charge_flux(backend):
    """size-independent property: sum_cells J_c = sum_p q (x2 - x1)_c"""
    conf = pic_conf(n_cells=N, q0=-0.3, q1=0.9)
    t = TILE[backend](conf)
    rng = np.random.default_rng(6)
    tot = np.zeros(3)
    for sp, q in ((0, -0.3), (1, 0.9)):
        n = 3000
        pos = (0.5 + rng.random((3, n)) * (np.array(N)[:, None] - 1.0)).astype(np.float32)
        vel = (0.7 * rng.standard_normal((3, n))).astype(np.float32)
        t.set_particles(sp, *pos, *vel, np.arange(n, dtype=np.uint64))
        u = vel.astype(np.float64)
        tot += q * (0.45 * u / np.sqrt(1 + (u * u).sum(0))).sum(1)
    t.op("deposit_current")
    J = t.get_fields()[2].astype(np.float64)
    assert np.allclose(J.reshape(3, -1).sum(1), tot, rtol=2e-4)


# ---- tests/py/test_pic_particle_sorting.py:25-85 ----------------------------------------------------------
def test_sort_keeps_the_multiset_and_orders_by_cell(backend):
    conf = pic_conf(n_cells=N)
    t = TILE[backend](conf)
    t.op("sort_particles")                                            # empty tile is fine
    rng = np.random.default_rng(7)
    n = 5000
    pos = (rng.random((3, n)) * np.array(N)[:, None]).astype(np.float32)
    pos = np.minimum(pos, np.nextafter(np.array(N, np.float32), 0)[:, None])
    vel = rng.standard_normal((3, n)).astype(np.float32)
    ids = np.arange(n, dtype=np.uint64)
    ids[rng.random(n) < 0.1] = DEAD
    t.set_particles(0, *pos, *vel, ids)
    before = sorted(zip(*[a.tolist() for a in t.get_particles(0, alive_only=True)]))
    for rounds in (1, 3):
        for _ in range(rounds):
            t.op("sort_particles")
        after = t.get_particles(0, alive_only=True)
        assert sorted(zip(*[a.tolist() for a in after])) == before
        k = t.sort_keys(0).astype(np.int64)
        assert np.all(np.diff(k) >= 0)                                # sorted, dead (UINT32_MAX) last
        assert np.count_nonzero(k == 0xFFFFFFFF) == np.count_nonzero(ids == DEAD)
    # stability: equal keys keep container order => a second sort is the identity
    a = t.get_particles(0, alive_only=False)
    t.op("sort_particles")
    b = t.get_particles(0, alive_only=False)
    assert np.array_equal(a[6], b[6])


# ---- tests/py-multirank/test_emf_simulation.py:116-222 (all tiles local here) ----------------------------
def test_J_exchange_multiplicities(backend):
    conf = emf_conf(n_tiles=(2, 2, 2), n_cells=N)
    g = GRID[backend](conf)
    J0 = np.array([1.0, 2.0, 3.0], np.float32)
    for idx in g.tiles():
        J = np.zeros_like(ZERO)
        J[:, 3:-3, 3:-3, 3:-3] = J0[:, None, None, None]
        g.set_fields(idx, ZERO, ZERO, J)
    g.local_communication(0)      # emf_J: halos now hold J0
    g.local_communication(6)      # emf_J_exchange
    nx, ny, nz = N
    a, b, c = nx - 6, ny - 6, nz - 6
    want = {1: a * b * c, 2: 2 * 3 * (a * b + b * c + a * c), 4: 4 * 9 * (a + b + c), 8: 8 * 27}
    for idx in g.tiles():
        J = g.get_fields(idx)[2][:, 3:-3, 3:-3, 3:-3]
        for comp in range(3):
            vals, counts = np.unique(J[comp], return_counts=True)
            assert vals.tolist() == [m * float(J0[comp]) for m in (1, 2, 4, 8)]
            assert counts.tolist() == [want[m] for m in (1, 2, 4, 8)]


def test_halo_sync_keeps_constant_fields_constant(backend):
    # tests/py-multirank/test_emf_simulation.py:53-113
    conf = emf_conf(n_tiles=(2, 2, 2), n_cells=N, cfl=0.45)
    g = GRID[backend](conf)
    for idx in g.tiles():
        F = np.zeros_like(ZERO)
        F[:, 3:-3, 3:-3, 3:-3] = np.array([1, 2, 3], np.float32)[:, None, None, None]
        g.set_fields(idx, F, F, ZERO)
    for _ in range(3):
        g.local_communication(1)
        g.local_communication(2)
        g.phase("push_half_b")
        g.local_communication(2)
        g.phase("push_e")
    for idx in g.tiles():
        E, B, _ = g.get_fields(idx)
        for comp in range(3):
            assert np.all(E[comp, 3:-3, 3:-3, 3:-3] == comp + 1) and np.all(B[comp, 3:-3, 3:-3, 3:-3] == comp + 1)


# ---- tests/py-multirank/test_pic_simulation.py:212-325, 365-437, 576-691 -----------------------------------
def test_particles_migrate_in_all_27_directions_with_periodic_wrap(backend):
    n_tiles = (2, 2, 2)
    conf = pic_conf(n_tiles=n_tiles, n_cells=N, cfl=0.45)
    g = GRID[backend](conf)
    L = np.array(n_tiles) * np.array(N)
    expect = {}
    next_id = 0
    for idx in g.tiles():
        mins = np.array(idx) * np.array(N)
        pos, vel = [], []
        for d in itertools.product((-1, 0, 1), repeat=3):
            # a particle 0.1 cells inside the tile edge facing d, moving outwards fast enough to cross
            p = np.array([mins[a] + (0.1 if d[a] < 0 else (N[a] - 0.1 if d[a] > 0 else N[a] / 2)) for a in range(3)])
            v = np.array([5.0 * d[a] for a in range(3)])
            pos.append(p)
            vel.append(v)
        pos, vel = np.array(pos).T, np.array(vel).T
        for sp in range(2):
            g.inject(idx, sp, pos, vel)
    g.phase("push_particles")
    g.phase("pack_outgoing_particles")
    g.local_communication(3)
    total = 0
    for idx in g.tiles():
        mins = np.array(idx) * np.array(N)
        for sp in range(2):
            x, y, z, ux, uy, uz, ids = g.get_particles(idx, sp, alive_only=True)
            total += len(ids)
            assert len(ids) == 27                                     # one arrival from each direction (+ the stayer)
            assert len(np.unique(ids)) == 27
            P = np.stack([x, y, z])
            assert np.all(P >= mins[:, None]) and np.all(P < (mins + np.array(N))[:, None])   # wrapped into the box
            assert np.all(P >= 0) and np.all(P < L[:, None])
    assert total == 2 * 27 * 8                                        # ids conserved


def test_migration_into_empty_tiles_and_noop(backend):
    # tests/py-multirank/test_pic_simulation.py:100-211, 532-575
    conf = pic_conf(n_tiles=(2, 2, 2), n_cells=N)
    g = GRID[backend](conf)
    n = 500
    rng = np.random.default_rng(9)
    pos = (2 + rng.random((3, n)) * (np.array(N)[:, None] - 4)).astype(np.float64)
    vel = np.zeros((3, n))
    g.inject((0, 0, 0), 0, pos, vel)                                  # everything else is empty
    before = g.get_particles((0, 0, 0), 0, alive_only=True)
    for _ in range(2):
        g.phase("push_particles")
        g.phase("pack_outgoing_particles")
        g.local_communication(3)
    after = g.get_particles((0, 0, 0), 0, alive_only=True)
    for a, b in zip(before, after):
        assert np.array_equal(a, b)
    for idx in g.tiles():
        if idx != (0, 0, 0):
            assert len(g.get_particles(idx, 0, alive_only=True)[6]) == 0


# ---- tests/py/test_pic_reflector_wall.py:29-133, test_emf_edge_bc.py:85-133 (restated) ------------
def _reflector_tile(backend):
    import runko_b200 as rb
    conf = pic_conf(n_tiles=(4, 4, 4), n_cells=N, cfl=1, q0=-1, current_depositer="zigzag_1st", current_filter=None)
    t = TILE[backend](conf)
    t.register_reflector_wall(rb.reflector_wall(walloc=5.0))
    t.set_fields(ZERO, ZERO, ZERO)
    return t


def test_stationary_wall_reflects_ux(backend):
    t = _reflector_tile(backend)
    t.inject(0, ([5.5], [5.5], [6.5]), ([-1.0], [0.0], [0.0]))
    t.op("push_particles")
    t.op("reflect_particles")
    x, y, z, ux, uy, uz, ids = t.get_particles(0, alive_only=True)
    assert len(x) == 1 and ux[0] > 0 and abs(abs(ux[0]) - 1.0) < 1e-4 and x[0] > 5.0
    assert abs(uy[0]) < 1e-5 and abs(uz[0]) < 1e-5


def test_no_current_behind_wall_after_reflect_and_deposit(backend):
    t = _reflector_tile(backend)
    j, k = np.meshgrid(np.arange(N[1]), np.arange(N[2]), indexing="ij")
    m = j.size
    t.inject(0, (np.full(m, 5.5), j.ravel() + 0.5, k.ravel() + 0.5), (np.full(m, -0.5), np.zeros(m), np.zeros(m)))
    for op in ("push_particles", "reflect_particles", "deposit_current"):
        t.op(op)
    J = t.get_fields()[2][:, 3:-3, 3:-3, 3:-3]
    assert np.all(np.abs(J[:, :4]) < 1e-5)          # x-indices 0..3: well behind the wall
    assert np.any(np.abs(J) > 1e-3)                 # and the reflected current itself is there


def test_particle_far_behind_wall_gets_killed(backend):
    t = _reflector_tile(backend)
    t.inject(0, ([2.0], [5.5], [6.5]), ([-0.1], [0.0], [0.0]))
    t.op("push_particles")
    t.op("reflect_particles")
    assert len(t.get_particles(0, alive_only=True)[0]) == 0


@pytest.mark.parametrize("side,position,expect", [(0, 5.0, slice(0, 6)), (1, 7.0, slice(7, 10))])
def test_x_edge_sets_correct_cells(backend, side, position, expect):
    import runko_b200 as rb
    t = TILE[backend](emf_conf(n_cells=N, n_tiles=(4, 4, 4)))
    one = np.ones_like(ZERO)
    t.set_fields(one, ZERO, ZERO)
    t.apply_edge_bc(rb.edge_bc(direction=0, side=side, position=position, Ex=0.0, Ey=0.0, Ez=0.0), rb.comm_mode.emf_E.value)
    E = t.get_fields()[0][:, 3:-3, 3:-3, 3:-3]
    mask = np.zeros(N[0], bool)
    mask[expect] = True
    assert np.all(E[:, mask] == 0) and np.all(E[:, ~mask] == 1)
