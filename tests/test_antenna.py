"""Driven-turbulence antenna: emf::Tile::register_antenna / deposit_antenna_current
(src/runko/emf/tile.c++:566-791).  emf/tile.c++ needs corgi and cannot be compiled in this image, so the oracle
is pinned by the reference's own unit tests (tests/py/test_emf_antenna.py, test_emf_antenna_time_evolution.py: 15
cases, run unmodified through tests/refshim by tests/test_reference_suite.py); here the same known answers are
restated for both backends and the CUDA path is compared with the oracle.

Tolerance of the CUDA path: the phases are narrowed to fp32 and go through cosf / sinf, which differ between glibc
and CUDA by an ulp or two of the POTENTIAL (|A| ~ 1); J is a second difference of it, so the comparison is absolute:
|dJ| <= 4e-6 * sum|A| (about 30 ulp of the potential through six differences times cfl)."""
import numpy as np
import pytest

import runko_b200 as rb
from backends import TILE
from util import emf_conf, random_lattice

N = (10, 11, 13)
INNER = (slice(None), slice(3, -3), slice(3, -3), slice(3, -3))


def analytic(n, tile_idx, fn):
    """fn(x) evaluated at the three Yee-staggered positions of every interior cell (global coordinates)"""
    i, j, k = np.meshgrid(*(np.arange(v, dtype=np.float64) + o * v for v, o in zip(n, tile_idx)), indexing="ij")
    return (fn(i + 0.5, j, k), fn(i, j + 0.5, k), fn(i, j, k + 0.5))


def make(backend, idx=(0, 0, 0), n_tiles=(1, 1, 1), cfl=0.45):
    t = TILE[backend](emf_conf(n_tiles=n_tiles, n_cells=N, cfl=cfl), idx)
    zero = np.zeros((3,) + tuple(v + 6 for v in N), np.float32)
    t.set_fields(zero, zero, zero)
    return t


def test_wave_vector_x(backend):
    """test_emf_antenna.py:38-66: A = (1,2,3) exp(i K x)  =>  J = cfl (0, 2, 3) K^2 cos(K x)"""
    t, K, cfl = make(backend), 0.06, 0.45
    t.register_antenna(rb.antenna_mode(A=(1, 2, 3), k=(K, 0, 0)))
    t.op("deposit_antenna_current")
    J = t.get_fields()[2][INNER]
    ax, ay, az = analytic(N, (0, 0, 0), lambda x, y, z: np.cos(K * x))
    assert np.allclose(J[0], 0, atol=1e-5)
    assert np.allclose(J[1], 2 * cfl * K ** 2 * ay, atol=1e-5)
    assert np.allclose(J[2], 3 * cfl * K ** 2 * az, atol=1e-5)


def test_mode_number_uses_the_global_grid(backend):
    """test_emf_antenna.py:194-224: n = (0, Ny, 0) means k_y = 2 pi Ny / L_y of the GLOBAL grid; tile (1,2,0)"""
    T, idx, Ny, cfl = (2, 3, 1), (1, 2, 0), 2, 0.45
    t = make(backend, idx=idx, n_tiles=T)
    K = 2 * np.pi * Ny / (T[1] * N[1])
    K2 = (2 * np.sin(K / 2)) ** 2                 # the discrete curl-curl of a plane wave: (2 sin(K/2))^2 instead of K^2
    t.register_antenna(rb.antenna_mode(A=(1, 2, 3), n=(0, Ny, 0)))
    t.op("deposit_antenna_current")
    J = t.get_fields()[2][INNER]
    ax, ay, az = analytic(N, idx, lambda x, y, z: np.cos(K * y))
    assert np.allclose(J[0], 1 * cfl * K2 * ax, atol=2e-5)
    assert np.allclose(J[1], 0, atol=2e-5)
    assert np.allclose(J[2], 3 * cfl * K2 * az, atol=2e-5)


def test_lap_coeffs_rotate_the_phase_and_run_out(backend):
    """test_emf_antenna_time_evolution.py:107-158: phi[lap] = q^(lap+1), q = p exp(i phi); fourth deposit throws"""
    t, K, cfl, p, ph = make(backend), 0.07, 0.45, 0.9, 0.4
    q = p * np.exp(1j * ph)
    t.register_antenna(rb.antenna_mode(A=(1, 2, 3), k=(0, K, 0), lap_coeffs=[q, q ** 2, q ** 3]))
    total = np.zeros((3,) + N)
    for M in (1, 2, 3):
        t.op("deposit_antenna_current")          # J accumulates: deposit_current adds
        ax, ay, az = analytic(N, (0, 0, 0), lambda x, y, z: np.cos(K * y + M * ph))
        total += np.stack([p ** M * cfl * K ** 2 * ax, 0 * ay, p ** M * 3 * cfl * K ** 2 * az])
        assert np.allclose(t.get_fields()[2][INNER], total, atol=3e-5)
    with pytest.raises(Exception):
        t.op("deposit_antenna_current")


def test_antenna_mode_argument_errors():
    with pytest.raises(rb.B2PError):
        rb.antenna_mode(A=(1, 2, 3))
    with pytest.raises(rb.B2PError):
        rb.antenna_mode(A=(1, 2, 3), k=(1, 0, 0), n=(1, 0, 0))
    with pytest.raises(rb.B2PError):
        rb.antenna_mode(A=(1, 2), k=(1, 0, 0))


@pytest.mark.gpu
def test_cuda_antenna_matches_the_oracle():
    """Five modes of both kinds with lap coefficients on a tile in the middle of a 3x2x2 grid, J preloaded with noise,
    three deposits: whole haloed lattice (the reference also adds the potential itself outside the interior)."""
    from oracle.oracle import OracleGrid
    rng = np.random.default_rng(77)
    conf = emf_conf(n_tiles=(3, 2, 2), n_cells=N, cfl=0.45)
    org = OracleGrid(conf)
    idx = (1, 1, 0)
    t = org.cid(*idx)
    tile = rb.Tile(idx, conf)
    E, B, J = (random_lattice(rng, N, 0.01) for _ in range(3))
    org.set_fields(t, E, B, J, with_halo=True)
    tile.set_fields_f32(E, B, J, with_halo=True)
    sumA = 0.0
    for m in range(5):
        A = rng.standard_normal(3)
        coeffs = (rng.standard_normal(3) + 1j * rng.standard_normal(3)) if m % 2 else None
        kw = dict(k=0.3 * rng.standard_normal(3)) if m < 3 else dict(n=rng.integers(-3, 4, 3))
        for reg in (lambda a: org.register_antenna(t, a), tile.register_antenna):
            reg(rb.antenna_mode(A=A, lap_coeffs=coeffs, **kw))
        sumA += np.abs(A).sum() * (1.0 if coeffs is None else np.abs(coeffs).max())
    for lap in range(3):
        org.tile_op(t, "deposit_antenna_current")
        tile.deposit_antenna_current()
        oJ = org.get_fields(t, with_halo=True)[2]
        gJ = tile.get_fields_f32(with_halo=True)[2]
        assert np.max(np.abs(gJ - oJ)) <= 4e-6 * sumA * (lap + 1), (lap, np.max(np.abs(gJ - oJ)))
        assert np.max(np.abs(oJ[INNER[1:] if False else INNER])) > 1e-3
    with pytest.raises(rb.B2PLogicError):
        tile.deposit_antenna_current()
