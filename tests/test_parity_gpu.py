"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on the same
seeded inputs.  Integer / index / byte outputs and every elementwise fp32 kernel
are required to be BIT-EXACT; only current deposition (atomic accumulation order)
and the reductions are compared within a stated tolerance."""
import numpy as np
import pytest

import runko_b200 as rb
from oracle.oracle import OracleGrid
from util import (DEAD, assert_bits_equal, emf_conf, pic_conf, random_lattice, random_particles, ulp_diff)

pytestmark = pytest.mark.gpu

SHAPES = [(10, 11, 13), (3, 3, 3), (32, 6, 17)]


def make_pair(conf, pic=True):
    org = OracleGrid(conf)
    cls = rb.PicTile if pic else rb.Tile
    tile = cls((0, 0, 0), conf)
    return org, tile


def load_fields(rng, org, tile, n_cells, t=0):
    E, B, J = (random_lattice(rng, n_cells) for _ in range(3))
    org.set_fields(t, E, B, J, with_halo=True)
    tile.set_fields_f32(E, B, J, with_halo=True)
    return E, B, J


def compare_fields(org, tile, t=0, which="EBJ"):
    oE, oB, oJ = org.get_fields(t, with_halo=True)
    gE, gB, gJ = tile.get_fields_f32(with_halo=True)
    if "E" in which:
        assert_bits_equal(gE, oE, "E")
    if "B" in which:
        assert_bits_equal(gB, oB, "B")
    if "J" in which:
        assert_bits_equal(gJ, oJ, "J")


@pytest.mark.parametrize("n_cells", SHAPES)
def test_fdtd2_bit_exact(n_cells):
    rng = np.random.default_rng(1)
    conf = emf_conf(n_cells=n_cells)
    org, tile = make_pair(conf, pic=False)
    load_fields(rng, org, tile, n_cells)
    for op in ("push_half_b", "push_e", "push_half_b", "add_current", "push_e"):
        org.tile_op(0, op)
        getattr(tile, op)()
        compare_fields(org, tile)


@pytest.mark.parametrize("n_cells", SHAPES)
def test_stencil_bit_exact(n_cells):
    rng = np.random.default_rng(2)
    kw = {}
    for ax in "xyz":
        for name in ("delta", "gamma", "beta_p1", "beta_p2", "beta2_p1", "beta2_p2", "beta3_p1", "beta3_p2",
                     "zeta_p1", "zeta_p2", "zeta2_p1", "zeta2_p2", "zeta3_p1", "zeta3_p2"):
            kw[f"stencil_{ax}_{name}"] = float(0.05 * rng.standard_normal())
    conf = emf_conf(n_cells=n_cells, field_propagator="stencil", **kw)
    org, tile = make_pair(conf, pic=False)
    load_fields(rng, org, tile, n_cells)
    for op in ("push_half_b", "push_e", "push_half_b"):
        org.tile_op(0, op)
        getattr(tile, op)()
        compare_fields(org, tile)


@pytest.mark.parametrize("variant", ["binomial2", "binomial2_unrolled"])
@pytest.mark.parametrize("n_cells", SHAPES)
def test_filter_bit_exact(variant, n_cells):
    rng = np.random.default_rng(3)
    conf = emf_conf(n_cells=n_cells, current_filter=variant)
    org, tile = make_pair(conf, pic=False)
    load_fields(rng, org, tile, n_cells)
    for _ in range(3):
        org.tile_op(0, "filter_current")
        tile.filter_current()
        compare_fields(org, tile, which="J")


def test_filter_without_config_is_logic_error():
    tile = rb.Tile((0, 0, 0), emf_conf())
    with pytest.raises(rb.B2PLogicError):
        tile.filter_current()


def load_particles(rng, org, tile, conf, n, dead_frac=0.05, t=0, margin=0.0, u_scale=0.5):
    mins = np.array(tile.mins)
    maxs = np.array(tile.maxs)
    for sp in range(2):
        pos, vel, ids = random_particles(rng, n + 17 * sp, mins, maxs, dead_frac=dead_frac, tag=sp + 1, margin=margin,
                                         u_scale=u_scale)
        org.set_particles(t, sp, *pos, *vel, ids)
        tile.set_particles_raw(sp, *pos, *vel, ids)


def compare_particles(org, tile, t=0, alive_only=False):
    for sp in range(2):
        o = org.get_particles(t, sp, alive_only=alive_only)
        g = tile.get_particles(sp, alive_only=alive_only)
        names = ["x", "y", "z", "ux", "uy", "uz", "id"]
        for a, b, nm in zip(g, o, names):
            if nm != "id" and not alive_only:
                # pos/vel of dead slots are unspecified in the reference (never read)
                alive = o[6] != DEAD
                a, b = a[alive], b[alive]
            assert_bits_equal(a, b, f"species {sp} {nm}")


@pytest.mark.parametrize("pusher", ["boris", "higuera_cary", "faraday"])
@pytest.mark.parametrize("interp", ["linear_1st", "linear_1st_unrolled"])
def test_push_bit_exact(pusher, interp):
    rng = np.random.default_rng(4)
    n_cells = (10, 11, 13)
    conf = pic_conf(n_cells=n_cells, particle_pusher=pusher, field_interpolator=interp, m1=3.0)
    org, tile = make_pair(conf)
    load_fields(rng, org, tile, n_cells)
    load_particles(rng, org, tile, conf, 5000)
    for _ in range(2):
        org.tile_op(0, "push_particles")
        tile.push_particles()
        compare_particles(org, tile)


def test_sort_bit_exact():
    rng = np.random.default_rng(5)
    n_cells = (10, 11, 13)
    conf = pic_conf(n_cells=n_cells)
    org, tile = make_pair(conf)
    load_particles(rng, org, tile, conf, 20000, dead_frac=0.1)
    for sp in range(2):
        assert_bits_equal(tile.sort_keys(sp), org.sort_keys(0, sp), "sort keys")
    for _ in range(2):
        org.tile_op(0, "sort_particles")
        tile.sort_particles()
        compare_particles(org, tile)
    # sortedness + dead slots at the end
    k = tile.sort_keys(0)
    assert np.all(np.diff(k.astype(np.int64)) >= 0)


@pytest.mark.parametrize("case", ["uniform-200k", "thirty-per-cell", "nearly-sorted", "crowded-cell", "very-crowded-cell", "radix-path"])
def test_sort_bit_exact_cases(case):
    """Counting sort, every way k_sort_cells orders a cell's member list: the 16- / 32-input register networks
    (test_sort_bit_exact: 14 per cell), the insertion sort for 33..64 (crowded cases: 40 per cell), the whole block
    ranking a long list (uniform-200k: 132 per cell; crowded-cell: 3000 in one), block ranges too long for the
    shared-memory staging (thirty-per-cell: networks and insertion sorts working in global memory); the hand-over to
    the radix sort when the largest cell exceeds SORT_RADIX_POP (decided from the first probe, then from the previous
    sort's hint), and the radix path forced by option: all must give the oracle's stable order."""
    import ctypes as C
    from runko_b200._lib import check
    rng = np.random.default_rng(55)
    n_cells = (12, 9, 14)
    conf = pic_conf(n_cells=n_cells)
    org, tile = make_pair(conf)
    L = rb.lib()
    if case == "radix-path":
        check(L.b2p_set_option(b"sort_counting", 0))
    try:
        n = 60000 if "crowded" in case else (48000 if case == "thirty-per-cell" else 200000)
        load_particles(rng, org, tile, conf, n, dead_frac=0.07)
        if "crowded" in case:
            # 3000 particles in one cell: ranked by the whole block in k_sort_cells;
            # 9000: over SORT_RADIX_POP -> radix sort for that container
            m = 3000 if case == "crowded-cell" else 9000
            for sp in range(2):
                x, y, z, ux, uy, uz, ids = org.get_particles(0, sp, alive_only=False)
                x[1000:1000 + m] = 5.25 + 0.5 * rng.random(m).astype(np.float32)
                y[1000:1000 + m] = 4.25
                z[1000:1000 + m] = 7.5
                org.set_particles(0, sp, x, y, z, ux, uy, uz, ids)
                tile.set_particles_raw(sp, x, y, z, ux, uy, uz, ids)
        rounds = 1 if case in ("uniform-200k", "radix-path") else 3
        for r in range(rounds):
            org.tile_op(0, "sort_particles")
            tile.sort_particles()
            compare_particles(org, tile)
            if r + 1 < rounds:                       # drift a little, kill a few, sort again
                for _ in range(2):
                    org.tile_op(0, "push_particles")
                    tile.push_particles()
                compare_particles(org, tile)
        k = tile.sort_keys(1)
        assert np.all(np.diff(k.astype(np.int64)) >= 0)
    finally:
        check(L.b2p_set_option(b"sort_counting", 1))


def test_const_division_bit_exact():
    """The pushers' constant-divisor division (particles.cu: DivC) equals the IEEE quotient bit for bit:
    random mantissas over the whole exponent range, values around the range-test thresholds, zeros of
    both signs, denormals, inf/nan, for several divisors incl. ones outside the fast-path range."""
    import ctypes as C
    from runko_b200._lib import check
    rng = np.random.default_rng(77)
    L = rb.lib()
    n = 1 << 22
    bits = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    x = bits.view(np.float32).copy()
    special = np.array([0.0, -0.0, 1e-45, -1e-45, 1e-38, 2.0 ** -100, np.nextafter(np.float32(2.0 ** -100), np.float32(0)),
                        2.0 ** 100, np.nextafter(np.float32(2.0 ** 100), np.float32(np.inf)), np.inf, -np.inf, np.nan, 3.4e38],
                       np.float32)
    x[:special.size] = special
    x[100:100 + (1 << 20)] = (rng.standard_normal(1 << 20) * 0.7).astype(np.float32)     # the physical range
    for c in (0.45, 0.45 / 2, 1.0, 0.999999, 3.0, 1.9999999, 1e-3, 7.7e5, 1e-7, 1e9, float(np.float32(1) / np.float32(3))):
        out, ref = np.empty(n, np.float32), np.empty(n, np.float32)
        check(L.b2p_selfcheck_const_division(x.ctypes.data_as(C.c_void_p), n, C.c_float(c), out.ctypes.data_as(C.c_void_p),
                                             ref.ctypes.data_as(C.c_void_p)))
        host = x / np.float32(c)
        same = (out.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(out) & np.isnan(ref))
        assert same.all(), f"c={c}: {np.count_nonzero(~same)} mismatches vs device IEEE division, e.g. x={x[~same][:4]}"
        ok_host = (ref.view(np.uint32) == host.view(np.uint32)) | (np.isnan(ref) & np.isnan(host))
        assert ok_host.all(), f"c={c}: device IEEE division differs from numpy"


def test_sort_empty_and_single():
    conf = pic_conf()
    tile = rb.PicTile((0, 0, 0), conf)
    tile.sort_particles()
    assert tile.container_size(0) == 0


def test_pack_outgoing_bit_exact():
    rng = np.random.default_rng(6)
    n_cells = (10, 11, 13)
    conf = pic_conf(n_tiles=(3, 3, 3), n_cells=n_cells)
    org = OracleGrid(conf)
    t = org.cid(1, 1, 1)
    tile = rb.PicTile((1, 1, 1), conf)
    load_particles(rng, org, tile, conf, 30000, dead_frac=0.05, t=t, margin=0.8)
    org.tile_op(t, "pack_outgoing_particles")
    tile.pack_outgoing_particles()
    obuf, oends = org.get_outgoing(t)
    gbuf, gends = tile.get_outgoing()
    assert_bits_equal(gends, oends, "subregion ends")
    assert len(obuf) > 1000
    for f in ("pos", "vel", "id"):
        assert_bits_equal(gbuf[f], obuf[f], f"outgoing {f}")
    compare_particles(org, tile, t=t)      # leavers marked dead in place, everything else untouched
    # all 26 directions are populated
    counts = np.diff(np.concatenate([[0], oends.astype(np.int64)]))
    assert np.count_nonzero(counts[:27]) == 26


def test_deposit_tolerance():
    rng = np.random.default_rng(7)
    n_cells = (10, 11, 13)
    for dep in ("zigzag_1st_atomic", "zigzag_1st"):
        conf = pic_conf(n_cells=n_cells, current_depositer=dep, q0=-0.7, q1=0.4)
        org, tile = make_pair(conf)
        load_fields(rng, org, tile, n_cells)
        load_particles(rng, org, tile, conf, 40000, dead_frac=0.05)
        org.tile_op(0, "deposit_current")
        tile.deposit_current()
        oJ = org.get_fields(0, with_halo=True)[2]
        gJ = tile.get_fields_f32(with_halo=True)[2]
        # stated tolerance: per-step J within 1e-5 * max|J| (atomic accumulation order differs)
        assert np.max(np.abs(gJ - oJ)) <= 1e-5 * np.max(np.abs(oJ))
        assert np.max(np.abs(oJ)) > 0


def test_deposit_single_particle_bit_exact():
    """With one particle there is no accumulation-order freedom: J must be bit-identical."""
    n_cells = (10, 11, 13)
    conf = pic_conf(n_cells=n_cells, q0=-0.7)
    org, tile = make_pair(conf)
    rng = np.random.default_rng(8)
    for _ in range(20):
        pos, vel, ids = random_particles(rng, 1, tile.mins, tile.maxs, u_scale=2.0)
        for sp in range(2):
            n = 1 if sp == 0 else 0
            org.set_particles(0, sp, *(p[:n] for p in pos), *(v[:n] for v in vel), ids[:n])
            tile.set_particles_raw(sp, *(p[:n] for p in pos), *(v[:n] for v in vel), ids[:n])
        org.tile_op(0, "deposit_current")
        tile.deposit_current()
        oJ = org.get_fields(0, with_halo=True)[2]
        gJ = tile.get_fields_f32(with_halo=True)[2]
        assert np.array_equal(gJ, oJ)   # +0 == -0 allowed here (0 + x vs x)


def test_energies():
    rng = np.random.default_rng(9)
    n_cells = (10, 11, 13)
    conf = pic_conf(n_cells=n_cells)
    org, tile = make_pair(conf)
    load_fields(rng, org, tile, n_cells)
    load_particles(rng, org, tile, conf, 10000, dead_frac=0.1)
    ob, oe = org.field_energy(0)
    gb, ge = tile.field_energy()
    # reference sums in serial fp32; the GPU sums in fp64: 1e-4 relative
    assert abs(gb - ob) <= 1e-4 * ob and abs(ge - oe) <= 1e-4 * oe
    for sp in range(2):
        ok, on = org.kinetic_energy(0, sp)
        gk, gn = tile.kinetic_energy(sp)
        assert gn == on
        assert abs(gk - ok) <= 1e-9 * ok


def build_grids(conf, rng, ppc=3, u_scale=0.4, fields=True):
    org = OracleGrid(conf)
    grid = rb.Grid(conf)
    tiles = {}
    T = conf.n_tiles
    n_cells = tuple(conf.n_cells_per_tile)
    for i in range(T[0]):
        for j in range(T[1]):
            for k in range(T[2]):
                tile = rb.PicTile((i, j, k), conf)
                t = org.cid(i, j, k)
                if fields:
                    E, B, J = (random_lattice(rng, n_cells, 0.3) for _ in range(3))
                    org.set_fields(t, E, B, J, with_halo=True)
                    tile.set_fields_f32(E, B, J, with_halo=True)
                n = ppc * int(np.prod(n_cells))
                for sp in range(conf.__dict__.get("_nsp", 2)):
                    pos, vel, _ = random_particles(rng, n, tile.mins, tile.maxs, u_scale=u_scale)
                    org.inject(t, sp, *pos.astype(np.float64), *vel.astype(np.float64))
                    tile._inject_arrays(sp, pos.astype(np.float64), vel.astype(np.float64))
                grid.add_tile(tile)
                tiles[(i, j, k)] = tile
    return org, grid, tiles


def compare_all_fields(org, tiles, which="EBJ", exact=True, tol=0.0):
    for (i, j, k), tile in tiles.items():
        t = org.cid(i, j, k)
        o = org.get_fields(t, with_halo=True)
        g = tile.get_fields_f32(with_halo=True)
        for name, a, b in zip("EBJ", g, o):
            if name not in which:
                continue
            if exact:
                assert_bits_equal(a, b, f"tile {(i, j, k)} {name}")
            else:
                assert np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(b)), 1e-30), (name, (i, j, k))


@pytest.mark.parametrize("n_tiles", [(2, 2, 2), (1, 1, 1), (3, 1, 2)])
def test_halo_fill_and_J_exchange_bit_exact(n_tiles):
    rng = np.random.default_rng(10)
    conf = pic_conf(n_tiles=n_tiles, n_cells=(5, 6, 7))
    org, grid, tiles = build_grids(conf, rng, ppc=0)
    for mode in (rb.comm_mode.emf_E, rb.comm_mode.emf_B, rb.comm_mode.emf_J_exchange, rb.comm_mode.emf_J):
        org.local_communication(mode.value)
        grid.local_communication(mode)
        compare_all_fields(org, tiles)


@pytest.mark.parametrize("n_tiles", [(2, 2, 2), (1, 1, 1), (3, 2, 1)])
def test_particle_migration_bit_exact(n_tiles):
    """push -> pack -> local exchange (+ periodic wrap): containers identical, slot by slot."""
    rng = np.random.default_rng(11)
    conf = pic_conf(n_tiles=n_tiles, n_cells=(5, 6, 7), cfl=0.45)
    org, grid, tiles = build_grids(conf, rng, ppc=4, u_scale=3.0)
    total0 = sum(len(t.get_ids(sp)) for t in tiles.values() for sp in range(2))
    for lap in range(3):
        org.phase("push_particles")
        grid.phase("push_particles")
        org.phase("pack_outgoing_particles")
        grid.phase("pack_outgoing_particles")
        for (i, j, k), tile in tiles.items():
            ob, oe = org.get_outgoing(org.cid(i, j, k))
            gb, ge = tile.get_outgoing()
            assert_bits_equal(ge, oe, "ends")
            for f in ("pos", "vel", "id"):
                assert_bits_equal(gb[f], ob[f], f)
        org.local_communication(rb.comm_mode.pic_particle.value)
        grid.local_communication(rb.comm_mode.pic_particle)
        if lap == 1:
            org.phase("sort_particles")
            grid.phase("sort_particles")
        for (i, j, k), tile in tiles.items():
            compare_particles(org, tile, t=org.cid(i, j, k))
    total1 = sum(len(t.get_ids(sp)) for t in tiles.values() for sp in range(2))
    assert total0 == total1   # id conservation across migration


@pytest.mark.parametrize("n_tiles,filt", [((2, 2, 2), "binomial2"), ((1, 1, 1), "binomial2_unrolled")])
def test_full_lap_parity(n_tiles, filt):
    """Whole laps of projects/pic-turbulence/pic.py.  Lap 1: particle state bit-exact (the push
    precedes the deposit), fields to 1e-5; after 10 laps: relative L2 <= 1e-3."""
    rng = np.random.default_rng(12)
    conf = pic_conf(n_tiles=n_tiles, n_cells=(6, 7, 8), current_filter=filt, q0=-0.05, q1=0.05)
    org, grid, tiles = build_grids(conf, rng, ppc=4, u_scale=0.3)
    org.local_communication(1); org.local_communication(2)
    grid.local_communication(rb.comm_mode.emf_E); grid.local_communication(rb.comm_mode.emf_B)
    org.step_pic(0)
    grid.step_pic(0)
    for (i, j, k), tile in tiles.items():
        compare_particles(org, tile, t=org.cid(i, j, k))
    compare_all_fields(org, tiles, which="B")
    compare_all_fields(org, tiles, which="EJ", exact=False, tol=1e-5)
    for lap in range(1, 10):
        org.step_pic(lap)
        grid.step_pic(lap)
    for (i, j, k), tile in tiles.items():
        t = org.cid(i, j, k)
        for sp in range(2):
            o = org.get_particles(t, sp)
            g = tile.get_particles(sp)
            # same particles in the same tiles, up to rare boundary flips from the 1e-7 J noise
            common, oi, gi = np.intersect1d(o[6], g[6], return_indices=True)
            assert len(common) >= 0.999 * len(o[6])
            for c in range(6):
                d = g[c][gi] - o[c][oi]
                assert np.sqrt(np.mean(d * d)) <= 1e-3 * max(np.sqrt(np.mean(o[c][oi] ** 2)), 1e-3)
    compare_all_fields(org, tiles, exact=False, tol=1e-3)
    ob, oe, ok, on = org.energies()
    gb, ge, gk, gn = grid.energies()
    assert abs(gb - ob) <= 1e-4 * ob and abs(ge - oe) <= 1e-3 * max(oe, 1e-12)
    assert np.allclose(gk, ok, rtol=1e-4)


def test_emf_lap_bit_exact():
    rng = np.random.default_rng(13)
    conf = emf_conf(n_tiles=(2, 1, 2), n_cells=(8, 9, 10), cfl=1.0)
    org = OracleGrid(conf)
    grid = rb.Grid(conf)
    tiles = {}
    for i in range(2):
        for k in range(2):
            tile = rb.Tile((i, 0, k), conf)
            E, B, J = (random_lattice(rng, (8, 9, 10)) for _ in range(3))
            org.set_fields(org.cid(i, 0, k), E, B, J, with_halo=True)
            tile.set_fields_f32(E, B, J, with_halo=True)
            grid.add_tile(tile)
            tiles[(i, 0, k)] = tile
    for _ in range(5):
        org.step_emf()
        grid.step_emf()
    compare_all_fields(org, tiles, which="EB")


def test_multigpu_parity_two_ranks():
    """NCCL halo / J-exchange / particle migration between 2 GPUs vs the whole-grid oracle."""
    import os
    import subprocess
    import sys
    try:
        import torch
        ngpu = torch.cuda.device_count()
    except Exception:
        ngpu = 0
    if ngpu < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(here, "run_multigpu_parity.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("fuse", [1, 0])
@pytest.mark.parametrize("sequence", ["push-pack-comm-sort-deposit", "push-deposit", "push-pack-deposit", "push-push-pack-comm-deposit"])
def test_deposit_after_push_matches_reference_per_tile(fuse, sequence):
    """The fused push+deposit (stayers deposited in the push, arrivals on append) must give every
    tile the J the reference's deposit_current gives it — including halo contributions — and fall
    back to a fresh deposit whenever the leavers of that push were not removed."""
    from runko_b200._lib import check
    check(rb.lib().b2p_set_option(b"fuse_deposit", fuse))
    try:
        rng = np.random.default_rng(21)
        conf = pic_conf(n_tiles=(2, 2, 2), n_cells=(5, 6, 7), q0=-0.7, q1=0.4)
        org, grid, tiles = build_grids(conf, rng, ppc=6, u_scale=2.0)
        for step in sequence.split("-"):
            if step == "comm":
                org.local_communication(rb.comm_mode.pic_particle.value)
                grid.local_communication(rb.comm_mode.pic_particle)
            else:
                name = {"push": "push_particles", "pack": "pack_outgoing_particles", "sort": "sort_particles",
                        "deposit": "deposit_current"}[step]
                org.phase(name)
                grid.phase(name)
        for (i, j, k), tile in tiles.items():
            oJ = org.get_fields(org.cid(i, j, k), with_halo=True)[2]
            gJ = tile.get_fields_f32(with_halo=True)[2]
            assert np.max(np.abs(oJ)) > 0
            assert np.max(np.abs(gJ - oJ)) <= 1e-5 * np.max(np.abs(oJ)), (sequence, (i, j, k))
            if "comm" in sequence:
                compare_particles(org, tile, t=org.cid(i, j, k))
    finally:
        check(rb.lib().b2p_set_option(b"fuse_deposit", 1))
