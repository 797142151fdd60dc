"""BASELINE configs[0] and configs[1] as parity cases (bench.py measures configs[4]; configs[2] is
bench.py --workload emf-wave; configs[3]'s pieces are tests/test_shock_gpu.py): the physics set-up of the
reference's drivers on reduced grids, CUDA path vs CPU oracle, compared through the DIAGNOSTICS the drivers
write every lap — the north star's "energy-conservation and dispersion diagnostics agree":

  * configs[0] projects/pic-turbulence/pic.py:23-160: decaying turbulence, guide field + the driver's own B
    perturbation (np.random.seed(2) phases), pair plasma, boris + linear_1st + zigzag_1st_atomic + 3x binomial2,
    one 64^3 tile as shipped, 2 ppc per species, first 100 laps: mean kinetic energy and <B^2>, <E^2> per lap
    (io_average_*) within the stated relative tolerances, and the same evolution of their weighted sum on both.
  * configs[1] projects/pic-beam-instabilities/beam.py:46-233: two counter-streaming cold beams (gamma_b = 3) in
    species 0, thin z (6 cells), no filter, the driver's lap order: growth of the parallel / perpendicular field
    energies (the driver's own store_data() diagnostic) — same e-folding rate on both.
Tolerances are stated where asserted."""
import itertools

import numpy as np
import pytest

import runko_b200 as rb
from oracle.oracle import OracleGrid
from util import Conf

pytestmark = pytest.mark.gpu


def test_config0_turbulence_energy_diagnostics():
    n = (64, 64, 64)
    cfl, ppc = 0.45, 2
    oppc = 2 * ppc
    q0 = -(cfl ** 2) / (0.5 * oppc * 2.0)                       # pic.py:46-50 (gamma = c_omp = 1, m0 = m1 = 1)
    conf = Conf(n_tiles=[1, 1, 1], n_cells_per_tile=list(n), cfl=cfl, field_propagator="fdtd2", m0=1, m1=1, q0=q0, q1=abs(q0),
                particle_pusher="boris", field_interpolator="linear_1st", current_depositer="zigzag_1st_atomic",
                current_filter="binomial2")
    delgam, sigma = 0.3, 10
    m0 = abs(q0)
    binit = np.sqrt((1.0 + 1.5 * delgam) * oppc * m0 * cfl ** 2 * sigma)   # pic.py:64-67
    A0, n_perp, n_par = 0.8 * binit, 2, 3
    np.random.seed(n_perp)                                       # pic.py:94-97
    ph1, ph2, ph3 = (2.0 * np.pi * np.random.rand(n_perp, n_perp, n_par) for _ in range(3))
    kx, ky, kz = (2.0 * np.pi / L for L in n)

    def beta(a, b):
        return np.sqrt(8.0) / (np.sqrt(a ** 2 + b ** 2) * n_perp * np.sqrt(n_par))

    def Bx(x, y, z):                                             # pic.py:109-121
        bx = np.zeros_like(x)
        x, y = y, x
        for a, b in itertools.product(range(1, n_perp + 1), repeat=2):
            for o in range(1, n_par + 1):
                bx += beta(a, b) * a * np.sin(b * kx * x + ph1[a - 1, b - 1, o - 1]) * np.cos(a * ky * y + ph2[a - 1, b - 1, o - 1]) \
                    * np.sin(o * kz * z + ph3[a - 1, b - 1, o - 1])
        return A0 * bx

    def By(x, y, z):                                             # pic.py:123-135
        by = np.zeros_like(x)
        x, y = y, x
        for a, b in itertools.product(range(1, n_perp + 1), repeat=2):
            for o in range(1, n_par + 1):
                by -= beta(a, b) * b * np.cos(b * kx * x + ph1[a - 1, b - 1, o - 1]) * np.sin(a * ky * y + ph2[a - 1, b - 1, o - 1]) \
                    * np.sin(o * kz * z + ph3[a - 1, b - 1, o - 1])
        return A0 * by

    zero = lambda x, y, z: np.zeros_like(x)                      # noqa: E731
    Bz = lambda x, y, z: np.full_like(x, binit)                  # noqa: E731
    tile = rb.PicTile((0, 0, 0), conf)
    tile.batch_set_EBJ(zero, zero, zero, Bx, By, Bz, zero, zero, zero)
    E, B, J = tile.get_fields_f32(with_halo=True)
    org = OracleGrid(conf)
    org.set_fields(0, E, B, J, with_halo=True)
    grid = rb.Grid(conf)
    grid.add_tile(tile)
    rng = np.random.default_rng(42)
    ii, jj, kk = np.meshgrid(*(np.arange(v, dtype=np.float64) for v in n), indexing="ij")
    corner = np.stack([ii.ravel(), jj.ravel(), kk.ravel()])
    pos = np.concatenate([corner + rng.random(corner.shape) for _ in range(ppc)], axis=1)
    for sp in range(2):                                          # positrons on top of electrons (pic.py:141-156)
        vel = 0.55 * rng.standard_normal(pos.shape)              # thermal spread of theta ~ 0.3
        tile._inject_arrays(sp, pos, vel)
        org.inject(0, sp, *pos, *vel)
    for m in (1, 2):                                             # prelude: sync E, B halos (pic.py:177-185)
        org.local_communication(m)
        grid.local_communication(m)
    dev = {"kin": 0.0, "B": 0.0, "E": 0.0}
    tot0 = None
    for lap in range(100):
        org.step_pic(lap, threads=1)
        grid.step_pic(lap)
        if lap % 5 == 4 or lap < 3:
            ob, oe, ok, on = org.energies()
            gb, ge, gk, gn = grid.energies()
            if lap < 3:
                assert np.array_equal(on, gn)               # container sizes incl. dead slots; later a boundary crossing
                                                            # may fall on a neighbouring lap (1e-7 deposit-order noise)
            dev["kin"] = max(dev["kin"], float(np.max(np.abs(gk - ok) / ok)))
            dev["B"] = max(dev["B"], abs(gb - ob) / ob)
            dev["E"] = max(dev["E"], abs(ge - oe) / max(oe, 1e-30))
            # total energy in the diagnostics' own normalisation: field energies + sum_s m_s * kinetic_s
            tot_o, tot_g = ob + oe + m0 * float(ok.sum()), gb + ge + m0 * float(gk.sum())
            if tot0 is None:
                tot0 = (tot_o, tot_g)
    print("config 0, 100 laps: max relative deviation of the per-lap diagnostics", dev,
          "total-energy drift (oracle, cuda)", tot_o / tot0[0] - 1.0, tot_g / tot0[1] - 1.0)
    assert dev["kin"] <= 1e-4 and dev["B"] <= 1e-4      # diagnostics within 1e-4 relative
    assert dev["E"] <= 1e-3                             # <E^2> is four orders smaller and noise-dominated
    assert abs((tot_g / tot0[1]) - (tot_o / tot0[0])) <= 1e-4   # the same energy evolution on both


def test_config1_beam_growth_diagnostics():
    T, n = (2, 2, 1), (48, 20, 6)
    cfl, skin, num_n0, gamma_b = 0.45, 10, 8, 3.0
    v_b = np.sqrt(1 - 1 / gamma_b ** 2)
    num_n = 2 * num_n0
    q = (cfl / skin) ** 2 * gamma_b / num_n                      # beam.py:76-93 (alpha = 1)
    conf = Conf(n_tiles=list(T), n_cells_per_tile=list(n), cfl=cfl, field_propagator="fdtd2", q0=-q, m0=1,
                particle_pusher="boris", field_interpolator="linear_1st", current_depositer="zigzag_1st_atomic",
                current_filter="binomial2")
    org = OracleGrid(conf)
    grid = rb.Grid(conf)
    tiles = {}
    for i, j in itertools.product(range(T[0]), range(T[1])):
        tile = rb.PicTile((i, j, 0), conf)
        t = org.cid(i, j, 0)
        ii, jj, kk = np.meshgrid(*(np.arange(v, dtype=np.float64) for v in n), indexing="ij")
        corner = np.stack([ii.ravel() + i * n[0], jj.ravel() + j * n[1], kk.ravel()])
        for sign in (+1.0, -1.0):                                # both beams in species 0, same seed (beam.py:123-150)
            rng = np.random.default_rng(42)
            pos = np.repeat(corner, num_n0, axis=1) + rng.random((3, corner.shape[1] * num_n0))
            vel = 1e-3 * rng.standard_normal(pos.shape)
            vel[0] += sign * gamma_b * v_b
            tile._inject_arrays(0, pos, vel)
            org.inject(t, 0, *pos, *vel)
        grid.add_tile(tile)
        tiles[(i, j)] = tile

    def lap_of(phase, comm):                                     # beam.py:196-233, filter disabled as shipped
        phase("push_half_b"); comm(2)
        phase("push_particles"); phase("pack_outgoing_particles"); comm(3)
        phase("deposit_current"); comm(6); comm(0)
        phase("push_half_b"); comm(2)
        phase("push_e"); phase("add_current"); comm(1)

    def store_data(get):                                         # beam.py:169-190
        s = np.zeros(4)
        for f in get():
            Ei, Bi = f[0][:, 3:-3, 3:-3, 3:-3].astype(np.float64), f[1][:, 3:-3, 3:-3, 3:-3].astype(np.float64)
            s += [np.sum(Ei[0] ** 2), np.sum(Ei[1] ** 2 + Ei[2] ** 2), np.sum(Bi[0] ** 2), np.sum(Bi[1] ** 2 + Bi[2] ** 2)]
        return s

    series = {"o": [], "g": []}
    laps = 240
    for lap in range(laps):
        if lap % 20 == 0:
            series["o"].append(store_data(lambda: [org.get_fields(org.cid(i, j, 0), with_halo=True) for (i, j) in tiles]))
            series["g"].append(store_data(lambda: [t.get_fields_f32(with_halo=True) for t in tiles.values()]))
        lap_of(lambda p: (lap % 5 == 0 and p == "deposit_current" and org.phase("sort_particles")) or org.phase(p),
               org.local_communication)
        lap_of(lambda p: (lap % 5 == 0 and p == "deposit_current" and grid.phase("sort_particles")) or grid.phase(p),
               lambda m: grid.local_communication(m))
    o, g = np.array(series["o"]), np.array(series["g"])
    # field energy grows by orders of magnitude (the instability); both implementations must show the same curve
    growth = o[-1] / np.maximum(o[1], 1e-300)
    print("config 1: growth of (E_par^2, E_perp^2, B_par^2, B_perp^2) between laps 20 and", laps - 20, growth)
    assert growth[1] > 10.0 or growth[0] > 10.0
    rel = np.abs(g[1:] - o[1:]) / np.maximum(np.abs(o[1:]), 1e-300)
    print("config 1: max relative deviation of the four diagnostics over the run", rel.max(axis=0))
    # e-folding rates between the first and last samples agree to 2 % (exponential amplification of the 1e-7
    # deposit-order noise is allowed for; the curves themselves agree to 5 %)
    rate_o = np.log(o[-1, :2] / o[1, :2])
    rate_g = np.log(g[-1, :2] / g[1, :2])
    assert np.all(np.abs(rate_g - rate_o) <= 0.02 * np.abs(rate_o) + 0.02)
    assert np.all(rel[:, [0, 1, 3]] <= 0.05)
