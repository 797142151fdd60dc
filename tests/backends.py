"""One face over the two implementations the tests compare: the CPU oracle (checker) and
the CUDA path through the C-ABI (product).  Test bodies written against this face run
unchanged on either; the `backend` fixture in conftest.py marks the b200 variant `gpu`."""
import numpy as np


class OracleTile:
    name = "oracle"

    def __init__(self, conf, idx=(0, 0, 0)):
        from oracle.oracle import OracleGrid
        self.g = OracleGrid(conf)
        self.t = self.g.cid(*idx)
        n = self.g.n_cells
        self.mins = [float(idx[d] * n[d]) for d in range(3)]
        self.maxs = [float((idx[d] + 1) * n[d]) for d in range(3)]

    def set_fields(self, E=None, B=None, J=None):
        self.g.set_fields(self.t, E, B, J, with_halo=True)

    def get_fields(self):
        return self.g.get_fields(self.t, with_halo=True)

    def set_particles(self, sp, x, y, z, ux, uy, uz, ids):
        self.g.set_particles(self.t, sp, x, y, z, ux, uy, uz, ids)

    def inject(self, sp, pos, vel):
        self.g.inject(self.t, sp, *pos, *vel)

    def get_particles(self, sp, alive_only=False):
        return self.g.get_particles(self.t, sp, alive_only=alive_only)

    def op(self, name):
        self.g.tile_op(self.t, name)

    def sort_keys(self, sp):
        return self.g.sort_keys(self.t, sp)

    def get_outgoing(self):
        return self.g.get_outgoing(self.t)

    def register_reflector_wall(self, wall):
        self.g.register_reflector_wall(self.t, wall)

    def register_antenna(self, mode):
        self.g.register_antenna(self.t, mode)

    def apply_edge_bc(self, bc, mode):
        self.g.apply_edge_bc(self.t, bc, int(mode))


class B200Tile:
    name = "b200"

    def __init__(self, conf, idx=(0, 0, 0)):
        import runko_b200 as rb
        has_pic = conf.__dict__.get("particle_pusher") is not None
        self.tile = (rb.PicTile if has_pic else rb.Tile)(idx, conf)
        self.mins, self.maxs = self.tile.mins, self.tile.maxs

    def set_fields(self, E=None, B=None, J=None):
        self.tile.set_fields_f32(E, B, J, with_halo=True)

    def get_fields(self):
        return self.tile.get_fields_f32(with_halo=True)

    def set_particles(self, sp, x, y, z, ux, uy, uz, ids):
        self.tile.set_particles_raw(sp, x, y, z, ux, uy, uz, ids)

    def inject(self, sp, pos, vel):
        self.tile._inject_arrays(sp, pos, vel)

    def get_particles(self, sp, alive_only=False):
        return self.tile.get_particles(sp, alive_only=alive_only)

    def op(self, name):
        getattr(self.tile, name)()

    def sort_keys(self, sp):
        return self.tile.sort_keys(sp)

    def get_outgoing(self):
        return self.tile.get_outgoing()

    def register_reflector_wall(self, wall):
        self.tile.register_reflector_wall(wall)

    def register_antenna(self, mode):
        self.tile.register_antenna(mode)

    def apply_edge_bc(self, bc, mode):
        self.tile.apply_edge_bc(bc, int(mode))


class OracleGridB:
    """whole periodic grid on the oracle"""
    name = "oracle"

    def __init__(self, conf):
        from oracle.oracle import OracleGrid
        self.g = OracleGrid(conf)
        self.n_tiles = self.g.n_tiles

    def tiles(self):
        T = self.n_tiles
        return [(i, j, k) for k in range(T[2]) for j in range(T[1]) for i in range(T[0])]

    def set_fields(self, idx, E=None, B=None, J=None):
        self.g.set_fields(self.g.cid(*idx), E, B, J, with_halo=True)

    def get_fields(self, idx):
        return self.g.get_fields(self.g.cid(*idx), with_halo=True)

    def inject(self, idx, sp, pos, vel):
        self.g.inject(self.g.cid(*idx), sp, *pos, *vel)

    def get_particles(self, idx, sp, alive_only=False):
        return self.g.get_particles(self.g.cid(*idx), sp, alive_only=alive_only)

    def local_communication(self, mode):
        self.g.local_communication(int(mode))

    def phase(self, name):
        self.g.phase(name)

    def step_pic(self, lap):
        self.g.step_pic(lap)

    def get_outgoing(self, idx):
        return self.g.get_outgoing(self.g.cid(*idx))


class B200GridB:
    name = "b200"

    def __init__(self, conf):
        import runko_b200 as rb
        self.rb = rb
        self.grid = rb.Grid(conf)
        self.n_tiles = tuple(conf.n_tiles)
        has_pic = conf.__dict__.get("particle_pusher") is not None
        self._tiles = {}
        for idx in self.tiles():
            t = (rb.PicTile if has_pic else rb.Tile)(idx, conf)
            self.grid.add_tile(t)
            self._tiles[idx] = t

    tiles = OracleGridB.tiles

    def set_fields(self, idx, E=None, B=None, J=None):
        self._tiles[idx].set_fields_f32(E, B, J, with_halo=True)

    def get_fields(self, idx):
        return self._tiles[idx].get_fields_f32(with_halo=True)

    def inject(self, idx, sp, pos, vel):
        self._tiles[idx]._inject_arrays(sp, pos, vel)

    def get_particles(self, idx, sp, alive_only=False):
        return self._tiles[idx].get_particles(sp, alive_only=alive_only)

    def local_communication(self, mode):
        self.grid.local_communication(int(mode))

    def phase(self, name):
        self.grid.phase(name)

    def step_pic(self, lap):
        self.grid.step_pic(lap)

    def get_outgoing(self, idx):
        return self._tiles[idx].get_outgoing()


TILE = {"oracle": OracleTile, "b200": B200Tile}
GRID = {"oracle": OracleGridB, "b200": B200GridB}
