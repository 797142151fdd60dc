"""Drives projects/pic-turbulence/pic.py's lap function (pic.py:187-221) through the drop-in modules
runko_cpp_bindings / pycorgi / mpi4py of runko_b200/dropin — i.e. through pybind11 -> the C-ABI -> the CUDA kernels.

Where /root/reference is present the UNMODIFIED reference package does the driving (runko.Configuration,
runko.TileGrid, tile_grid.configure_simulation -> runko.Simulation.for_each_lap); on a box without it (the GPU box)
`MiniSimulation` below restates runko.Simulation's dispatch (runko/simulation.py:235-319): `prtcl_*` / `grid_*` ->
`for tile in local_tiles: getattr(tile, name)()`, `comm_local` -> corgi_grid.local_communication(mode.value),
`comm_external` -> handshake + recv_data / send_data / wait_data, `io_average_*` -> the text-file diagnostics.

Usage: python dropin_lap_driver.py <outdir> <ncells> <ppc> <nlaps>  -> writes <outdir>/initial_state.npz (what the
oracle side of the test loads), <outdir>/average_*.txt (the diagnostics) and prints one JSON line.
TEST INFRASTRUCTURE."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(ROOT, "runko_b200", "dropin"), ROOT]
REF = "/root/reference"
HAVE_REF = os.path.isdir(os.path.join(REF, "runko")) and os.environ.get("DROPIN_DRIVER_NO_REFERENCE") != "1"
if HAVE_REF:
    sys.path.append(REF)

import numpy as np  # noqa: E402


class MiniSimulation:
    """runko.Simulation's lap-function dispatch (runko/simulation.py:209-333), nothing else."""

    def __init__(self, corgi_grid, nt, outdir):
        import runko_cpp_bindings as rcb
        self._g, self._last, self.lap, self._rcb = corgi_grid, nt, 0, rcb
        self._paths = {k: os.path.join(outdir, f"average_{k}.txt") for k in ("kinetic_energy", "B_energy_density", "E_energy_density")}

    def local_tiles(self):
        for cid in self._g.get_local_tiles():
            yield self._g.get_tile(cid)

    def _action(self, method, *vargs):
        rcb = self._rcb
        if method.startswith("prtcl_"):
            name = {"prtcl_push": "push_particles", "prtcl_sort": "sort_particles", "prtcl_pack_outgoing": "pack_outgoing_particles",
                    "prtcl_deposit_current": "deposit_current"}[method]
            for tile in self.local_tiles():
                getattr(tile, name)()
        elif method.startswith("grid_"):
            for tile in self.local_tiles():
                getattr(tile, method[5:])(*vargs)
        elif method.startswith("io_"):
            what = method[3:]
            if what == "average_kinetic_energy":
                rcb.pic.threeD._write_average_kinetic_energy(self.lap, self._paths["kinetic_energy"], self._g)
            elif what == "average_B_energy_density":
                rcb.emf.threeD._write_average_B_energy_density(self.lap, self._paths["B_energy_density"], self._g)
            elif what == "average_E_energy_density":
                rcb.emf.threeD._write_average_E_energy_density(self.lap, self._paths["E_energy_density"], self._g)
            else:
                raise AttributeError(method)
        elif method.startswith("comm_"):
            for mode in vargs:
                if type(mode) is not rcb.tools.comm_mode:
                    raise TypeError("Communications only accept runko.comm_mode arguments.")
                if method[5:] == "local":
                    self._g.local_communication(mode.value)
                else:
                    hs = rcb.tools._virtual_tile_sync_handshake_mode(mode)
                    for m in ([hs] if hs else []) + [mode.value]:
                        self._g.recv_data(m); self._g.send_data(m); self._g.wait_data(m)
        else:
            raise RuntimeError(f"{method} is not supported!")

    class _X:
        def __init__(self, action): self._a = action
        def __getattr__(self, name): return lambda *v: self._a(name, *v)

    def prelude(self, f):
        f(self._X(self._action))

    def for_each_lap(self, f):
        while self.lap < self._last:
            f(self._X(self._action))
            self.lap += 1


def main():
    outdir, n, ppc, nlaps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    os.makedirs(outdir, exist_ok=True)
    if HAVE_REF:
        import runko
        tools, pic = runko.tools, runko.pic
        config = runko.Configuration(None)
    else:
        import runko_cpp_bindings as rcb
        import pycorgi.threeD as pycorgi
        tools, pic = rcb.tools, rcb.pic

        class Configuration:
            def __getattr__(self, name):
                if name.startswith("__"):
                    raise AttributeError(name)
                return None
        config = Configuration()
    # projects/pic-turbulence/pic.py:22-67 at n^3 cells, one tile (config 0 of BASELINE.json)
    cfl, oppc = 0.45, 2 * ppc
    config.n_tiles, config.n_cells_per_tile = [1, 1, 1], [n, n, n]
    config.tile_partitioning, config.catepillar_track_length = "catepillar_track", 1
    config.cfl, config.field_propagator, config.current_filter = cfl, "fdtd2", "binomial2"
    config.m0 = config.m1 = 1
    config.q0 = -(cfl ** 2) / (0.5 * oppc * 2.0)
    config.q1 = abs(config.q0)
    config.particle_pusher, config.field_interpolator, config.current_depositer = "boris", "linear_1st", "zigzag_1st_atomic"
    config.n_laps, config.io_outdir, config.verbose = nlaps, outdir, False
    delgam, sigma = 0.3, 10
    binit = np.sqrt((1.0 + 1.5 * delgam) * oppc * abs(config.q0) * cfl ** 2 * sigma)
    rng_b = np.random.default_rng(2)
    ph = 2.0 * np.pi * rng_b.random(3)
    k = 2.0 * np.pi / n

    zero = lambda x, y, z: np.zeros_like(x)                                      # noqa: E731
    Bx = lambda x, y, z: 0.3 * binit * np.sin(2 * k * y + ph[0]) * np.cos(k * z + ph[1])   # noqa: E731
    By = lambda x, y, z: -0.3 * binit * np.cos(k * x + ph[2]) * np.sin(2 * k * z + ph[0])  # noqa: E731
    Bz = lambda x, y, z: np.full_like(x, binit)                                   # noqa: E731
    rng = np.random.default_rng(42)

    def pgen0(x, y, z):                                                           # pic.py:141-156
        pgen0.pos = x + rng.random(len(x)), y + rng.random(len(x)), z + rng.random(len(x))
        return pic.threeD.ParticleStateBatch(pos=pgen0.pos, vel=tuple(0.55 * rng.standard_normal(len(x)) for _ in range(3)))

    def pgen1(x, y, z):
        return pic.threeD.ParticleStateBatch(pos=pgen0.pos, vel=tuple(0.55 * rng.standard_normal(len(x)) for _ in range(3)))

    if HAVE_REF:
        tile_grid = runko.TileGrid(config)
        indices = list(tile_grid.local_tile_indices())
    else:
        corgi = pycorgi.Grid(1, 1, 1)
        corgi.set_grid_lims(0, n, 0, n, 0, n)
        indices = [(0, 0, 0)]
    tiles = []
    for idx in indices:                                                           # pic.py:162-174
        tile = pic.threeD.Tile(idx, config)
        tile.batch_set_EBJ(zero, zero, zero, Bx, By, Bz, zero, zero, zero)
        for _ in range(ppc):
            tile.batch_inject_to_cells(0, pgen0)
            tile.batch_inject_to_cells(1, pgen1)
        (tile_grid.add_tile if HAVE_REF else corgi.add_tile)(tile, idx)
        tiles.append(tile)
    t0 = tiles[0]
    E, B, J = t0.get_fields_f32(with_halo=True)
    st = {"E": E, "B": B, "J": J}
    for sp in range(2):
        for name, a in zip(("x", "y", "z", "ux", "uy", "uz", "id"), t0.get_particles(sp, alive_only=False)):
            st[f"{name}{sp}"] = a
    np.savez(os.path.join(outdir, "initial_state.npz"), **st)
    if HAVE_REF:
        simulation = tile_grid.configure_simulation(config)                       # -> runko.Simulation, unmodified
    else:
        simulation = MiniSimulation(corgi, nlaps, outdir)
    M = tools.comm_mode

    def sync_EB(x):                                                               # pic.py:177-185
        x.comm_external(M.emf_E, M.emf_B)
        x.comm_local(M.emf_E, M.emf_B)
    simulation.prelude(sync_EB)

    def pic_simulation_step(x):                                                   # pic.py:187-225
        x.grid_push_half_b()
        x.comm_external(M.emf_B); x.comm_local(M.emf_B)
        x.prtcl_push()
        x.prtcl_pack_outgoing()
        x.comm_external(M.pic_particle); x.comm_local(M.pic_particle)
        if simulation.lap % 5 == 0:
            x.prtcl_sort()
        x.prtcl_deposit_current()
        x.comm_external(M.emf_J); x.comm_local(M.emf_J_exchange)
        x.comm_external(M.emf_J); x.comm_local(M.emf_J)
        x.grid_filter_current()
        x.comm_external(M.emf_J); x.comm_local(M.emf_J)
        x.grid_filter_current()
        x.grid_filter_current()
        x.grid_push_half_b()
        x.comm_external(M.emf_B); x.comm_local(M.emf_B)
        x.grid_push_e()
        x.grid_add_current()
        x.comm_external(M.emf_E); x.comm_local(M.emf_E)
        x.io_average_kinetic_energy()
        x.io_average_B_energy_density()
        x.io_average_E_energy_density()
    simulation.for_each_lap(pic_simulation_step)
    import _b200pic
    print(json.dumps({"driver": "runko.Simulation (unmodified reference package)" if HAVE_REF else "MiniSimulation (restated dispatch)",
                      "tile_module": type(t0).__module__, "laps": nlaps, "kernels_launched": int(_b200pic.launch_count())}))


if __name__ == "__main__":
    main()
