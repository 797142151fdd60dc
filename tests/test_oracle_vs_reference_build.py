"""Pins the oracle restatement (oracle/pic_oracle.cpp) BIT FOR BIT against the reference's
own kernel sources compiled here from /root/reference (oracle/Makefile.ref ->
oracle/_ref/libref_kernels.so): the same seeded inputs go through emf::YeeLattice /
pic::ParticleContainer and through the oracle, and every output — fields, particle
containers slot by slot, migration buffers, sort order, and the reference's own
atomic-order deposit — must be identical.  Skipped where the reference build is absent
(it travels to the GPU box as a prebuilt .so; /root/reference itself does not)."""
import numpy as np
import pytest

from oracle import reference_build as rbuild
from oracle.oracle import OracleGrid
from util import DEAD, assert_bits_equal, emf_conf, pic_conf, random_lattice, random_particles


@pytest.fixture(scope="module", autouse=True)
def _ref():
    if not rbuild.available() and not rbuild.build():
        pytest.skip("reference build unavailable (no /root/reference and no prebuilt oracle/_ref/libref_kernels.so)")


SHAPES = [(10, 11, 13), (3, 3, 3), (16, 4, 7)]


def pair(conf, idx=(0, 0, 0)):
    org = OracleGrid(conf)
    return org, org.cid(*idx), rbuild.RefTile(conf, idx)


def load_fields(rng, org, t, ref, n):
    E, B, J = (random_lattice(rng, n) for _ in range(3))
    org.set_fields(t, E, B, J, with_halo=True)
    ref.set_fields(E, B, J)


def same_fields(org, t, ref, what=""):
    for a, b, nm in zip(org.get_fields(t, with_halo=True), ref.get_fields(), "EBJ"):
        assert_bits_equal(a, b, what + nm)


def load_particles(rng, org, t, ref, n, dead_frac=0.05, margin=0.0, u_scale=0.8):
    for sp in range(2):
        pos, vel, ids = random_particles(rng, n + 11 * sp, ref.mins, ref.maxs, dead_frac=dead_frac, tag=sp + 1, margin=margin,
                                         u_scale=u_scale)
        org.set_particles(t, sp, *pos, *vel, ids)
        ref.set_particles(sp, *pos, *vel, ids)


def same_particles(org, t, ref):
    for sp in range(2):
        o = org.get_particles(t, sp, alive_only=False)
        r = ref.get_particles(sp)
        assert_bits_equal(o[6], r[6], "ids")
        alive = r[6] != DEAD
        for c in range(6):
            assert_bits_equal(o[c][alive], r[c][alive], f"species {sp} comp {c}")


@pytest.mark.parametrize("n", SHAPES)
def test_fdtd2(n):
    rng = np.random.default_rng(1)
    org, t, ref = pair(emf_conf(n_cells=n))
    load_fields(rng, org, t, ref, n)
    for op in ("push_half_b", "push_e", "push_half_b", "add_current", "push_e"):
        org.tile_op(t, op)
        ref.op(op)
        same_fields(org, t, ref, op + " ")


@pytest.mark.parametrize("n", SHAPES)
def test_stencil(n):
    rng = np.random.default_rng(2)
    kw = {f"stencil_{ax}_{nm}": float(0.05 * rng.standard_normal()) for ax in "xyz" for nm in
          ("delta", "gamma", "beta_p1", "beta_p2", "beta2_p1", "beta2_p2", "beta3_p1", "beta3_p2", "zeta_p1", "zeta_p2",
           "zeta2_p1", "zeta2_p2", "zeta3_p1", "zeta3_p2")}
    org, t, ref = pair(emf_conf(n_cells=n, field_propagator="stencil", **kw))
    load_fields(rng, org, t, ref, n)
    for op in ("push_half_b", "push_e", "push_half_b"):
        org.tile_op(t, op)
        ref.op(op)
        same_fields(org, t, ref, op + " ")


@pytest.mark.parametrize("variant", ["binomial2", "binomial2_unrolled"])
@pytest.mark.parametrize("n", SHAPES)
def test_filter(variant, n):
    rng = np.random.default_rng(3)
    org, t, ref = pair(emf_conf(n_cells=n, current_filter=variant))
    load_fields(rng, org, t, ref, n)
    for _ in range(3):
        org.tile_op(t, "filter_current")
        ref.op("filter_current")
        same_fields(org, t, ref)


@pytest.mark.parametrize("pusher", ["boris", "higuera_cary", "faraday"])
@pytest.mark.parametrize("interp", ["linear_1st", "linear_1st_unrolled"])
def test_push(pusher, interp):
    rng = np.random.default_rng(4)
    n = (10, 11, 13)
    conf = pic_conf(n_tiles=(2, 3, 2), n_cells=n, particle_pusher=pusher, field_interpolator=interp, m1=3.0)
    org, t, ref = pair(conf, (1, 2, 1))
    load_fields(rng, org, t, ref, n)
    load_particles(rng, org, t, ref, 4000)
    for _ in range(3):
        org.tile_op(t, "push_particles")
        ref.op("push_particles")
        same_particles(org, t, ref)


@pytest.mark.parametrize("dep", ["zigzag_1st_atomic", "zigzag_1st"])
def test_deposit(dep):
    """The CPU reference accumulates serially in container order, so even the atomic variant is
    deterministic here and the oracle must match it bit for bit."""
    rng = np.random.default_rng(5)
    n = (10, 11, 13)
    conf = pic_conf(n_tiles=(2, 2, 2), n_cells=n, current_depositer=dep, q0=-0.7, q1=0.4)
    org, t, ref = pair(conf, (1, 0, 1))
    load_fields(rng, org, t, ref, n)
    load_particles(rng, org, t, ref, 20000)
    org.tile_op(t, "deposit_current")
    ref.op("deposit_current")
    same_fields(org, t, ref)


def test_sort():
    rng = np.random.default_rng(6)
    n = (10, 11, 13)
    conf = pic_conf(n_tiles=(2, 2, 2), n_cells=n)
    org, t, ref = pair(conf, (1, 1, 0))
    load_particles(rng, org, t, ref, 20000, dead_frac=0.1)
    for _ in range(2):
        org.tile_op(t, "sort_particles")
        ref.op("sort_particles")
        same_particles(org, t, ref)


def test_pack_outgoing_and_append():
    rng = np.random.default_rng(7)
    n = (10, 11, 13)
    conf = pic_conf(n_tiles=(3, 3, 3), n_cells=n)
    org, t, ref = pair(conf, (1, 1, 1))
    load_particles(rng, org, t, ref, 20000, margin=0.8)
    org.tile_op(t, "pack_outgoing_particles")
    ref.op("pack_outgoing_particles")
    ob, oe = org.get_outgoing(t)
    rb_, re_ = ref.get_outgoing()
    assert_bits_equal(oe, re_, "subregion ends")
    for f in ("pos", "vel", "id"):
        assert_bits_equal(ob[f], rb_[f], "outgoing " + f)
    same_particles(org, t, ref)
    assert np.count_nonzero(np.diff(np.concatenate([[0], oe[:27].astype(np.int64)]))) >= 24


@pytest.mark.parametrize("n_tiles", [(2, 2, 2), (1, 1, 1), (3, 1, 2)])
def test_whole_laps_on_a_periodic_grid(n_tiles):
    """halo fill, J exchange, migration with periodic wrap, sort, deposit, filters, FDTD: whole
    laps of projects/pic-turbulence/pic.py on every tile, reference kernels vs oracle."""
    rng = np.random.default_rng(8)
    n = (5, 6, 7)
    conf = pic_conf(n_tiles=n_tiles, n_cells=n, q0=-0.05, q1=0.05, current_filter="binomial2")
    org = OracleGrid(conf)
    ref = rbuild.RefGrid(conf)
    for idx, rt in ref.tiles.items():
        t = org.cid(*idx)
        E, B, J = (random_lattice(rng, n, 0.3) for _ in range(3))
        org.set_fields(t, E, B, J, with_halo=True)
        rt.set_fields(E, B, J)
        for sp in range(2):
            pos, vel, ids = random_particles(rng, 4 * int(np.prod(n)), rt.mins, rt.maxs, u_scale=1.5, tag=100 * t + sp)
            org.set_particles(t, sp, *pos, *vel, ids)
            rt.set_particles(sp, *pos, *vel, ids)
    for mode in (1, 2, 6, 0):
        org.local_communication(mode)
        ref.local_communication(mode)
    for idx, rt in ref.tiles.items():
        same_fields(org, org.cid(*idx), rt, f"comm {idx} ")
    for lap in range(6):
        org.step_pic(lap)
        ref.step_pic(lap)
        for idx, rt in ref.tiles.items():
            t = org.cid(*idx)
            same_particles(org, t, rt)
            same_fields(org, t, rt, f"lap {lap} {idx} ")


def test_energies():
    rng = np.random.default_rng(9)
    n = (10, 11, 13)
    org, t, ref = pair(pic_conf(n_cells=n))
    load_fields(rng, org, t, ref, n)
    load_particles(rng, org, t, ref, 5000, dead_frac=0.1)
    ob, oe = org.field_energy(t)
    rb_, re_, rk = ref.energies()
    assert ob == rb_ and oe == re_
    for sp in range(2):
        assert org.kinetic_energy(t, sp)[0] == rk[sp]


# ---------------------------------------------------------------- pic-shock boundary pieces --
def _bc(**kw):
    from runko_b200._abi import EdgeBC
    b = EdgeBC(direction=kw.get("direction", 0), side=kw.get("side", 0), position=kw.get("position", 0.0),
               E_components=kw.get("E_components", 7), B_components=kw.get("B_components", 7),
               J_components=kw.get("J_components", 7))
    for name in "EBJ":
        for c, v in enumerate(kw.get(name, (0.0, 0.0, 0.0))):
            getattr(b, name)[c] = v
    return b


@pytest.mark.parametrize("case", [
    dict(direction=0, side=0, position=5.0, E=(1.5, -2.0, 0.25), E_components=0b110),
    dict(direction=0, side=1, position=7.0, B=(0.1, 0.2, 0.3)),
    dict(direction=1, side=0, position=15.5, J=(3.0, 0.0, -1.0), J_components=0b101),
    dict(direction=2, side=1, position=13.0, E=(9.0, 8.0, 7.0)),
    dict(direction=0, side=0, position=100.0, E=(1.0, 1.0, 1.0)),     # tile fully behind the edge
    dict(direction=0, side=0, position=-3.0, E=(1.0, 1.0, 1.0)),      # tile fully in front: untouched
])
def test_edge_bc(case):
    """emf::Tile::apply_edge_bc(s): the reference's YeeLattice::apply_edge_bc vs the oracle, all three modes."""
    rng = np.random.default_rng(21)
    n = (10, 11, 13)
    conf = pic_conf(n_tiles=(2, 2, 2), n_cells=n)
    org, t, ref = pair(conf, (0, 1, 1))
    load_fields(rng, org, t, ref, n)
    bc = _bc(**case)
    for mode in (1, 2, 0):                       # emf_E, emf_B, emf_J
        org.apply_edge_bc(t, bc, mode)
        ref.apply_edge_bc(bc, mode)
    same_fields(org, t, ref, "apply_edge_bc ")
    org.register_edge_bc(t, bc)
    ref.register_edge_bc(bc)
    bc2 = _bc(direction=1, side=1, position=14.0, E=(4.0, 4.0, 4.0), B=(5.0, 5.0, 5.0), J=(6.0, 6.0, 6.0))
    org.register_edge_bc(t, bc2)
    ref.register_edge_bc(bc2)
    load_fields(rng, org, t, ref, n)
    for mode in (0, 1, 2):
        org.apply_edge_bcs(t, mode)
        ref.apply_edge_bcs(mode)
    same_fields(org, t, ref, "apply_edge_bcs ")


def test_edge_bc_rejects_other_modes():
    conf = pic_conf(n_tiles=(1, 1, 1), n_cells=(5, 5, 5))
    org, t, ref = pair(conf)
    bc = _bc(direction=0, side=0, position=3.0)
    from oracle.oracle import OracleError
    with pytest.raises(OracleError):
        org.apply_edge_bc(t, bc, 3)
    with pytest.raises(rbuild.RefError):
        ref.apply_edge_bc(bc, 3)


@pytest.mark.parametrize("beta", [0.0, 0.3])
@pytest.mark.parametrize("dep", ["zigzag_1st_atomic", "zigzag_1st"])
def test_reflector_wall(beta, dep):
    """pic::Tile::reflect_particles / ParticleContainer::reflect_at_wall (the reference's own, compiled from
    pic/reflector_wall.c++) vs the oracle: particles slot by slot (reflected, parked = dead, untouched), the
    correction current as it lands in J through deposit_current, and the advancing wall — over several laps
    of push -> reflect -> deposit -> advance."""
    from runko_b200._abi import ReflectorWall
    rng = np.random.default_rng(31)
    n = (12, 6, 7)
    conf = pic_conf(n_tiles=(2, 1, 1), n_cells=n, current_depositer=dep, q0=-0.7, q1=0.4, cfl=0.45)
    org, t, ref = pair(conf, (0, 0, 0))
    load_fields(rng, org, t, ref, n)
    # particles in front of the wall, streaming towards it
    for sp in range(2):
        m = 6000 + 13 * sp
        pos = np.stack([5.0 + 6.0 * rng.random(m), 6.0 * rng.random(m), 7.0 * rng.random(m)]).astype(np.float32)
        vel = (0.6 * rng.standard_normal((3, m))).astype(np.float32)
        vel[0] -= np.float32(0.8)
        pos[0, :40] = (2.5 + 2.0 * rng.random(40)).astype(np.float32)   # already behind the wall: parked (-> dead)
        ids = (np.uint64(sp + 1) << np.uint64(40)) | np.arange(m, dtype=np.uint64)
        ids[rng.random(m) < 0.05] = DEAD
        org.set_particles(t, sp, *pos, *vel, ids)
        ref.set_particles(sp, *pos, *vel, ids)
    gamma = 1.0 / np.sqrt(1.0 - beta * beta)
    wall = ReflectorWall(walloc=5.0, betawall=beta, gammawall=gamma)
    far = ReflectorWall(walloc=40.0, betawall=0.0, gammawall=1.0)        # not in this tile: ignored
    for w in (wall, far):
        org.register_reflector_wall(t, w)
        ref.register_reflector_wall(w)
    n_dead0 = sum(int(np.sum(ref.get_particles(sp)[6] == DEAD)) for sp in range(2))
    for lap in range(4):
        for op in ("push_particles", "reflect_particles"):
            org.tile_op(t, op)
            ref.op(op)
        same_particles(org, t, ref)
        org.tile_op(t, "deposit_current")
        ref.op("deposit_current")
        same_fields(org, t, ref, f"lap {lap} ")
        org.tile_op(t, "advance_reflector_walls")
        ref.op("advance_reflector_walls")
        assert org.reflector_walls(t) == ref.reflector_walls()
    # the case exercised all three branches: reflected (ux > 0 near the wall), parked (new dead slots)
    n_dead1 = sum(int(np.sum(ref.get_particles(sp)[6] == DEAD)) for sp in range(2))
    x, _, _, ux, _, _, ids = ref.get_particles(0, alive_only=True)
    assert n_dead1 > n_dead0
    assert np.any(ux > 0.5)


def test_reflector_wall_outside_tile_is_a_no_op():
    from runko_b200._abi import ReflectorWall
    rng = np.random.default_rng(33)
    n = (6, 6, 6)
    conf = pic_conf(n_tiles=(3, 1, 1), n_cells=n)
    org, t, ref = pair(conf, (2, 0, 0))
    load_fields(rng, org, t, ref, n)
    load_particles(rng, org, t, ref, 500)
    w = ReflectorWall(walloc=3.0, betawall=0.0, gammawall=1.0)
    org.register_reflector_wall(t, w)
    ref.register_reflector_wall(w)
    for op in ("reflect_particles", "deposit_current"):
        org.tile_op(t, op)
        ref.op(op)
    same_particles(org, t, ref)
    same_fields(org, t, ref)
