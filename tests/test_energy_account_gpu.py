"""The kinetic-energy account (runko_b200/csrc/particles.cuh: KE_SLOTS): a push of every container of a grid leaves the
per-species sum of sqrt(1 + u.u) - 1 (pic/particle.c++:352-377), pack_outgoing_particles subtracts the leavers, the particle
exchange adds the arrivals, so `Grid.energies()` — the reference's per-lap io_average_kinetic_energy — reads the account
instead of every container.  The account must equal the sum over the containers (option energy_cache = 0 forces that sum on
the same state) after whole laps, in the middle of a lap, and it must be dropped by every change it does not follow.
The push keeps the account only when the energies were read since the previous push."""
import numpy as np
import pytest

import runko_b200 as rb
from runko_b200._lib import check
from util import Conf

pytestmark = pytest.mark.gpu


def _grid(n_tiles=(3, 2, 2), edge=8, ppc=6):
    cfl = 0.45
    q0 = -(cfl ** 2) / (0.5 * 2 * ppc * 2.0)
    conf = Conf(n_tiles=list(n_tiles), n_cells_per_tile=[edge] * 3, cfl=cfl, field_propagator="fdtd2", current_filter="binomial2",
                q0=q0, m0=1.0, q1=abs(q0), m1=1.0, particle_pusher="boris", field_interpolator="linear_1st",
                current_depositer="zigzag_1st_atomic")
    grid = rb.Grid(conf)
    tiles = []
    for i in range(n_tiles[0]):
        for j in range(n_tiles[1]):
            for k in range(n_tiles[2]):
                t = rb.PicTile((i, j, k), conf)
                grid.add_tile(t)
                tiles.append(t)
    grid.set_uniform_B(0.0, 0.0, 0.3)
    grid.inject_thermal(ppc, 0.3, seed=7)
    for m in (rb.comm_mode.emf_E, rb.comm_mode.emf_B):
        grid.local_communication(m)
    return conf, grid, tiles


def _both(grid):
    """(kinetic energies as energies() returns them, launches it took, the same from the container sums, launches)"""
    L = rb.lib()
    a0 = L.b2p_launch_count()
    ke = np.array(grid.energies()[2], np.float64)
    a1 = L.b2p_launch_count()
    check(L.b2p_set_option(b"energy_cache", 0))
    try:
        full = np.array(grid.energies()[2], np.float64)
    finally:
        check(L.b2p_set_option(b"energy_cache", 1))
    a2 = L.b2p_launch_count()
    return ke, a1 - a0, full, a2 - a1


def test_account_equals_container_sums_over_laps_and_inside_a_lap():
    conf, grid, tiles = _grid()
    ke, n_acc, full, n_full = _both(grid)
    assert n_acc == n_full                                   # nothing pushed yet: no account, both are container sums
    np.testing.assert_array_equal(ke, full)
    for lap in range(7):                                     # laps 0 and 5 sort
        grid.step_pic(lap)
        ke, n_acc, full, n_full = _both(grid)
        assert n_acc < n_full, "energies() did not use the account after a whole-grid lap"
        np.testing.assert_allclose(ke, full, rtol=2e-7)
    # inside a lap: after the push (leavers still in their old containers), after the pack (leavers dead), after the exchange
    M = rb.comm_mode
    grid.phase("push_half_b"); grid.local_communication(M.emf_B)
    grid.phase("push_particles")
    ke_push, n_acc, full, n_full = _both(grid)
    assert n_acc < n_full
    np.testing.assert_allclose(ke_push, full, rtol=2e-7)
    grid.phase("pack_outgoing_particles")
    ke_pack, n_acc, full, n_full = _both(grid)
    assert n_acc < n_full
    np.testing.assert_allclose(ke_pack, full, rtol=2e-7)
    assert np.all(ke_pack < ke_push)                         # the leavers' energy left the account ...
    grid.local_communication(M.pic_particle)
    ke_comm, n_acc, full, n_full = _both(grid)
    assert n_acc < n_full
    np.testing.assert_allclose(ke_comm, full, rtol=2e-7)
    np.testing.assert_allclose(ke_comm, ke_push, rtol=1e-12)  # ... and came back with the arrivals (periodic grid: same particles)


def test_account_is_dropped_by_changes_it_does_not_follow():
    conf, grid, tiles = _grid(n_tiles=(2, 2, 1))
    # the push keeps the account only for a caller that read the energies since the previous push (the reference's lap
    # does, every lap): a bare lap does not pay for it
    grid.step_pic(1)
    ke, n_acc, full, n_full = _both(grid)
    assert n_acc == n_full
    np.testing.assert_array_equal(ke, full)
    grid.step_pic(2)
    ke, n_acc, full, n_full = _both(grid)
    assert n_acc < n_full
    np.testing.assert_allclose(ke, full, rtol=2e-7)
    # an upload: double every momentum of one container
    x, y, z, ux, uy, uz, ids = tiles[1].get_particles(0, alive_only=False)
    tiles[1].set_particles_raw(0, x, y, z, 2 * ux, 2 * uy, 2 * uz, ids)
    ke2, n_acc, full2, n_full = _both(grid)
    assert n_acc == n_full, "the account survived set_particles"
    np.testing.assert_array_equal(ke2, full2)
    assert ke2[0] > ke[0] * 1.01
    # a push of only some tiles does not restart it
    for t in tiles[:2]:
        t.push_particles()
    ke3, n_acc, full3, n_full = _both(grid)
    assert n_acc == n_full
    np.testing.assert_array_equal(ke3, full3)
    # the per-tile API over all tiles does (runko/simulation.py's `for tile in tiles: tile.push_particles()`)
    for t in tiles:
        t.push_particles()
    ke4, n_acc, full4, n_full = _both(grid)
    assert n_acc < n_full
    np.testing.assert_allclose(ke4, full4, rtol=2e-7)
    # injection drops it again
    tiles[0]._inject_arrays(1, np.array(tiles[0].mins)[:, None] + 4.0, np.full((3, 1), 0.5))
    ke5, n_acc, full5, n_full = _both(grid)
    assert n_acc == n_full
    np.testing.assert_array_equal(ke5, full5)
