"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py):
 - `-m "not gpu"`: the oracle still reproduces them bit-for-bit (pins the checker);
 - `-m gpu`: the CUDA path, through the C-ABI, reproduces them — bit-exact for every
   elementwise fp32 kernel and all integer / index / byte outputs; deposition within
   1e-5 * max|J| (atomic accumulation order)."""
import ast
import os

import numpy as np
import pytest

from backends import GRID, TILE
from util import DEAD, assert_bits_equal, emf_conf, pic_conf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N = (6, 7, 9)


def load(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    return d, ast.literal_eval(str(d["conf"])) if "conf" in d else None


@pytest.mark.parametrize("name", ["fdtd2", "stencil", "filter_binomial2", "filter_binomial2_unrolled"])
def test_field_kernels(backend, name):
    d, kw = load(name)
    t = TILE[backend](emf_conf(n_cells=N, **kw))
    t.set_fields(d["E"], d["B"], d["J"])
    for op in d["ops"]:
        t.op(str(op))
    E, B, J = t.get_fields()
    assert_bits_equal(E, d["oE"], name + " E")
    assert_bits_equal(B, d["oB"], name + " B")
    assert_bits_equal(J, d["oJ"], name + " J")


def run_particle_case(backend, name):
    d, kw = load(name)
    conf = pic_conf(n_tiles=tuple(int(v) for v in d["n_tiles"]), n_cells=N, q0=-0.7, q1=0.4, m1=3.0, **kw)
    t = TILE[backend](conf, tuple(int(v) for v in d["idx"]))
    t.set_fields(d["E"], d["B"], d["J"])
    for sp in range(2):
        t.set_particles(sp, *d[f"in{sp}"], d[f"in{sp}_id"])
    for op in d["ops"]:
        t.op(str(op))
    return d, t


def check_particles(d, t, sorted_keys=False):
    for sp in range(2):
        p = t.get_particles(sp, alive_only=False)
        assert_bits_equal(p[6], d[f"out{sp}_id"], "ids")
        alive = d[f"out{sp}_id"] != DEAD
        for c in range(6):
            assert_bits_equal(np.asarray(p[c])[alive], d[f"out{sp}"][c][alive], f"species {sp} comp {c}")
        assert_bits_equal(t.sort_keys(sp), d[f"keys{sp}"], "cell keys")


@pytest.mark.parametrize("pusher", ["boris", "higuera_cary", "faraday"])
def test_push(backend, pusher):
    d, t = run_particle_case(backend, "push_" + pusher)
    check_particles(d, t)


def test_sort(backend):
    d, t = run_particle_case(backend, "sort")
    check_particles(d, t)
    k = t.sort_keys(0).astype(np.int64)
    assert np.all(np.diff(k) >= 0)


def test_pack_outgoing(backend):
    d, t = run_particle_case(backend, "pack_outgoing")
    check_particles(d, t)
    buf, ends = t.get_outgoing()
    assert_bits_equal(ends, d["out_ends"], "subregion ends")
    assert_bits_equal(buf["pos"], d["out_pos"], "outgoing pos")
    assert_bits_equal(buf["vel"], d["out_vel"], "outgoing vel")
    assert_bits_equal(buf["id"], d["out_id"], "outgoing id")


@pytest.mark.parametrize("name", ["deposit_atomic", "deposit_sorted"])
def test_deposit(backend, name):
    d, t = run_particle_case(backend, name)
    J = t.get_fields()[2]
    if backend == "oracle":
        assert_bits_equal(J, d["oJ"], "J")
    else:   # stated tolerance: atomic accumulation order differs
        assert np.max(np.abs(J - d["oJ"])) <= 1e-5 * np.max(np.abs(d["oJ"]))


def test_grid_comm_and_laps(backend):
    d, _ = load("grid_laps")
    n_tiles, n = tuple(int(v) for v in d["n_tiles"]), tuple(int(v) for v in d["n_cells"])
    conf = pic_conf(n_tiles=n_tiles, n_cells=n, q0=-0.05, q1=0.05, current_filter="binomial2")
    g = GRID[backend](conf)
    order = [(t % n_tiles[0], (t // n_tiles[0]) % n_tiles[1], t // (n_tiles[0] * n_tiles[1])) for t in range(int(np.prod(n_tiles)))]
    for t, idx in enumerate(order):
        g.set_fields(idx, d[f"t{t}_E"], d[f"t{t}_B"], d[f"t{t}_J"])
        for sp in range(2):
            p = d[f"t{t}_p{sp}"].astype(np.float64)
            g.inject(idx, sp, p[:3], p[3:])
    for mode in (1, 2, 6, 0):
        g.local_communication(mode)
    for t, idx in enumerate(order):
        E, B, J = g.get_fields(idx)
        assert_bits_equal(E, d[f"t{t}_commE"], "halo E")
        assert_bits_equal(B, d[f"t{t}_commB"], "halo B")
        assert_bits_equal(J, d[f"t{t}_commJ"], "J exchange + halo")
    for lap in range(int(d["laps"])):
        g.step_pic(lap)
    exact = backend == "oracle"
    for t, idx in enumerate(order):
        E, B, J = g.get_fields(idx)
        for a, nm in ((E, "E"), (B, "B"), (J, "J")):
            ref = d[f"t{t}_lap{nm}"]
            if exact:
                assert_bits_equal(a, ref, "lap " + nm)
            else:
                assert np.max(np.abs(a - ref)) <= 1e-4 * np.max(np.abs(ref)), nm
        for sp in range(2):
            p = g.get_particles(idx, sp, alive_only=False)
            ids = d[f"t{t}_lap_id{sp}"]
            if exact:
                assert_bits_equal(p[6], ids, "ids after laps")
            else:   # the same particles in the same tiles (boundary flips from 1e-7 J noise are not expected in 2 laps)
                assert np.array_equal(np.sort(p[6]), np.sort(ids))


# ---- pic-shock pieces, antenna, snapshot (tests/golden/make_golden_shock.py) ----------------------
def test_reflector_wall_golden(backend):
    """Fixture = the REFERENCE'S OWN pic/reflector_wall.c++ (oracle/_ref) over three laps of push -> reflect -> deposit
    -> advance with a moving wall: particles slot by slot bit-exact; J bit-exact on the oracle, 1e-5 max|J| on CUDA."""
    import runko_b200 as rb
    d = np.load(os.path.join(GOLD, "shock_reflector.npz"))
    n = (9, 5, 6)
    t = TILE[backend](pic_conf(n_tiles=(2, 1, 1), n_cells=n, q0=-0.7, q1=0.4, cfl=0.45))
    t.set_fields(d["E"], d["B"], d["J"])
    for sp in range(2):
        t.set_particles(sp, *d[f"in{sp}"], d[f"in{sp}_id"])
    w = d["wall"]
    t.register_reflector_wall(rb.reflector_wall(walloc=float(w[0]), betawall=float(w[1]), gammawall=float(w[2])))
    for _ in range(int(d["laps"])):
        for op in ("push_particles", "reflect_particles", "deposit_current", "advance_reflector_walls"):
            t.op(op)
    for sp in range(2):
        p = t.get_particles(sp, alive_only=False)
        assert_bits_equal(p[6], d[f"out{sp}_id"], "ids")
        alive = d[f"out{sp}_id"] != DEAD
        for c in range(6):
            assert_bits_equal(np.asarray(p[c])[alive], d[f"out{sp}"][c][alive], f"species {sp} comp {c}")
    assert int(np.sum(d["out0_id"] == DEAD)) > int(np.sum(d["in0_id"] == DEAD))     # the fixture parks particles
    J = t.get_fields()[2]
    if backend == "oracle":
        assert_bits_equal(J, d["oJ"], "J")
    else:
        assert np.max(np.abs(J - d["oJ"])) <= 1e-5 * np.max(np.abs(d["oJ"]))


def test_edge_bc_golden(backend):
    """Fixture = the reference's own YeeLattice::apply_edge_bc through three registered BCs and all three modes."""
    import runko_b200 as rb
    d = np.load(os.path.join(GOLD, "shock_edge_bc.npz"))
    t = TILE[backend](pic_conf(n_tiles=(2, 2, 2), n_cells=(9, 5, 6)), (0, 1, 1))
    t.set_fields(d["E"], d["B"], d["J"])
    obj = t.g if backend == "oracle" else t.tile
    for kw in ast.literal_eval(str(d["bcs"])):
        bc = rb.edge_bc(**kw)
        obj.register_edge_bc(t.t, bc) if backend == "oracle" else obj.register_edge_bc(bc)
    for mode in (2, 0, 1):
        obj.apply_edge_bcs(t.t, mode) if backend == "oracle" else obj.apply_edge_bcs(mode)
    for a, nm in zip(t.get_fields(), "EBJ"):
        assert_bits_equal(a, d["o" + nm], nm)


def test_antenna_golden(backend):
    import runko_b200 as rb
    d = np.load(os.path.join(GOLD, "antenna.npz"))
    n = (9, 5, 6)
    t = TILE[backend](emf_conf(n_tiles=(2, 2, 1), n_cells=n, cfl=0.45), (1, 0, 0))
    z = np.zeros_like(d["J"])
    t.set_fields(z, z, d["J"])
    sumA = 0.0
    for m in ast.literal_eval(str(d["modes"])):
        if "lap_coeffs" in m:
            m["lap_coeffs"] = eval(m["lap_coeffs"])
        t.register_antenna(rb.antenna_mode(**m))
        sumA += float(np.abs(m["A"]).sum())
    t.op("deposit_antenna_current")
    t.op("deposit_antenna_current")
    J = t.get_fields()[2]
    if backend == "oracle":
        assert_bits_equal(J, d["oJ"], "J")
    else:   # fp32 phases through cosf / sinf: an ulp of the potential between glibc and CUDA (tests/test_antenna.py)
        assert np.max(np.abs(J - d["oJ"])) <= 8e-6 * sumA


def test_snapshot_golden(backend, tmp_path):
    d = np.load(os.path.join(GOLD, "snapshot_inputs.npz"))
    T, n = (2, 1, 2), (4, 6, 4)
    g = GRID[backend](pic_conf(n_tiles=T, n_cells=n))
    for t in range(4):
        idx = (t % T[0], (t // T[0]) % T[1], t // (T[0] * T[1]))
        g.set_fields(idx, d[f"t{t}_E"], d[f"t{t}_B"], d[f"t{t}_J"])
        for sp in range(2):
            p = d[f"t{t}_p{sp}"].astype(np.float64)
            g.inject(idx, sp, p[:3], p[3:])
    (g.g if backend == "oracle" else g.grid).write_fields_snapshot(tmp_path, 9, 2, 2)
    assert open(tmp_path / "flds_9.bin", "rb").read() == open(os.path.join(GOLD, "snapshot_flds_9.bin"), "rb").read()
