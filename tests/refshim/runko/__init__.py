"""A stand-in `runko` package for running the REFERENCE'S OWN unit tests
(/root/reference/tests/py/test_{emf,pic}*.py, unmodified, where they lie) against
this repo.  It exposes exactly the surface those tests touch — `Configuration`,
`emf.threeD.Tile`, `pic.threeD.{Tile, ParticleState, ParticleStateBatch}` — with the
host logic of runko_b200.tiles (the product's mirror of the pybind11 interface) bound to
one of two backends:

  RUNKO_SHIM_BACKEND=oracle (default)  the CPU oracle  -> pins the oracle (CPU, no GPU)
  RUNKO_SHIM_BACKEND=b200              libb200pic.so   -> the CUDA path through the C-ABI

TEST INFRASTRUCTURE ONLY (lives under tests/; may import oracle/).
"""
import os
import sys
import types

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from runko_b200 import _abi  # noqa: E402
from runko_b200 import tiles as _t  # noqa: E402

BACKEND = os.environ.get("RUNKO_SHIM_BACKEND", "oracle")


class Configuration:
    """runko/configuration.py:4-56 without the .ini reader: a missing attribute reads as None."""

    def __init__(self, config_path=None):
        if config_path is not None:
            raise NotImplementedError("the shim only supports Configuration(None)")

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return None


if BACKEND == "b200":
    EmfTile, PicTile = _t.Tile, _t.PicTile
else:
    from oracle.oracle import OracleError, OracleGrid

    class _OracleBacked:
        _need_pic = False
        error_type = RuntimeError

        def __init__(self, tile_grid_idx, config):
            try:
                self._cfg = _abi.make_config(config, need_pic=self._need_pic)
                self._g = OracleGrid(self._cfg)
            except (_abi.ConfigError, OracleError) as e:
                raise RuntimeError(str(e)) from None
            self.index = tuple(int(v) for v in tile_grid_idx)
            for d in range(3):
                if not 0 <= self.index[d] < self._cfg.n_tiles[d]:
                    raise RuntimeError("Trying to create tile outside of configured grid.")
            self._t = self._g.cid(*self.index)
            self.n_cells = tuple(self._cfg.n_cells)
            self.mins = [float(self.index[d] * self.n_cells[d]) for d in range(3)]
            self.maxs = [float((self.index[d] + 1) * self.n_cells[d]) for d in range(3)]

        def _backend_set_fields(self, E, B, J, with_halo):
            self._g.set_fields(self._t, E, B, J, with_halo=with_halo)

        def _backend_get_fields(self, with_halo):
            return self._g.get_fields(self._t, with_halo=with_halo)

        def _op(self, name):
            try:
                self._g.tile_op(self._t, name)
            except OracleError as e:
                raise RuntimeError(str(e)) from None

    class EmfTile(_OracleBacked, _t.EmfTileHost):
        def _ck(self, f, *a):
            try:
                return f(*a)
            except OracleError as e:
                raise RuntimeError(str(e)) from None

        def register_antenna(self, mode): self._ck(self._g.register_antenna, self._t, mode)
        def deposit_antenna_current(self): self._op("deposit_antenna_current")
        def register_edge_bc(self, bc): self._ck(self._g.register_edge_bc, self._t, bc)
        def apply_edge_bcs(self, mode): self._ck(self._g.apply_edge_bcs, self._t, _t._mode(mode))
        def apply_edge_bc(self, bc, mode): self._ck(self._g.apply_edge_bc, self._t, bc, _t._mode(mode))
        def push_half_b(self): self._op("push_half_b")
        def push_e(self): self._op("push_e")
        def add_current(self): self._op("add_current")
        def filter_current(self): self._op("filter_current")

    class PicTile(_t.PicTileHost, EmfTile):
        _need_pic = True

        def get_particles(self, sp, alive_only=True):
            try:
                return tuple(np.array(a) for a in self._g.get_particles(self._t, int(sp), alive_only=alive_only))
            except OracleError as e:
                raise RuntimeError(str(e)) from None

        def _backend_inject(self, sp, a):
            try:
                self._g.inject(self._t, sp, *a)
            except OracleError as e:
                raise RuntimeError(str(e)) from None

        def push_particles(self): self._op("push_particles")
        def deposit_current(self): self._op("deposit_current")
        def sort_particles(self): self._op("sort_particles")
        def pack_outgoing_particles(self): self._op("pack_outgoing_particles")
        def register_reflector_wall(self, wall): self._ck(self._g.register_reflector_wall, self._t, wall)
        def reflect_particles(self): self._op("reflect_particles")
        def advance_reflector_walls(self): self._op("advance_reflector_walls")


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


emf = _module("runko.emf", threeD=_module("runko.emf.threeD", Tile=EmfTile, edge_bc=_t.edge_bc, antenna_mode=_t.antenna_mode))
pic = _module("runko.pic", threeD=_module("runko.pic.threeD", Tile=PicTile, ParticleState=_t.ParticleStateD,
                                          ParticleStateBatch=_t.ParticleStateBatch, reflector_wall=_t.reflector_wall))
tools = _module("runko.tools", comm_mode=_t.comm_mode)

# runko.MovingInjector is pure Python in the reference: its own file is loaded where it lies (this package only
# exists to run the reference's unit tests, which need /root/reference anyway)
import importlib.util as _ilu  # noqa: E402
_mi = "/root/reference/runko/moving_injector.py"
if os.path.exists(_mi):
    _spec = _ilu.spec_from_file_location("runko.moving_injector", _mi)
    _mod = _ilu.module_from_spec(_spec)
    _spec.loader.exec_module(_mod)
    MovingInjector = _mod.MovingInjector
