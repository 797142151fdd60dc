"""Host-side mirror of the reference's pybind11 tile API for the PIC hot path:
emf.threeD.Tile (src/runko/bindings/pyemf.c++:214-267), pic.threeD.Tile
(src/runko/bindings/pypic.c++:92-139) and the slice of pycorgi.threeD.Grid that
runko/simulation.py drives (external/corgi/pycorgi/pycorgi.c++:79-126,317-374).
Same method names, argument meaning and error behaviour; every call goes through
the C-ABI of libb200pic.so (include/b200pic.h).
"""
import ctypes as C
import enum

import numpy as np

from . import _abi
from ._abi import ConfigError, ParticleState, make_config
from ._lib import B2PError, B2PLogicError, check, lib


class comm_mode(enum.Enum):
    """runko::comm_mode (src/runko/communication_common.h:30-38, bindings/pytools.c++:23-29)."""
    emf_J = 0
    emf_E = 1
    emf_B = 2
    pic_particle = 3
    pic_particle_extra = 4
    emf_J_exchange = 6


_number_of_particles = 5  # not exported by the reference either


def _virtual_tile_sync_handshake_mode(mode):
    """communication_common.h:42-49"""
    return _number_of_particles if mode == comm_mode.pic_particle else None


def _get_gpu_mem_kB():
    return int(lib().b2p_gpu_mem_kB())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class ParticleStateBatch:
    """pic::ParticleStateBatch (pic/tile.h:38-43)."""

    def __init__(self, pos, vel):
        self.pos = tuple(pos)
        self.vel = tuple(vel)


class ParticleStateD:
    """runko::ParticleState<double> as exposed to Python (bindings/pypic.c++:60-64)."""

    def __init__(self, pos, vel):
        self.pos = list(pos)
        self.vel = list(vel)


def edge_bc(*, direction=0, side=0, position=0.0, Ex=0.0, Ey=0.0, Ez=0.0, Bx=0.0, By=0.0, Bz=0.0,
            Jx=0.0, Jy=0.0, Jz=0.0, E_components=0b111, B_components=0b111, J_components=0b111):
    """emf.threeD.edge_bc (bindings/pyemf.c++:171-199; keyword-only like the reference)."""
    bc = _abi.EdgeBC(direction=int(direction), side=int(side), position=float(position),
                     E_components=int(E_components), B_components=int(B_components), J_components=int(J_components))
    bc.E[:] = (Ex, Ey, Ez)
    bc.B[:] = (Bx, By, Bz)
    bc.J[:] = (Jx, Jy, Jz)
    return bc


class antenna_mode:
    """emf.threeD.antenna_mode (bindings/pyemf.c++:102-170): keyword-only A and exactly one of k (wave vector) /
    n (mode number of the global grid), optional complex lap_coeffs consumed one per deposit."""

    def __init__(self, *, A, k=None, n=None, lap_coeffs=None):
        if (k is not None and n is not None) or (k is None and n is None):
            raise B2PError("antenna_mode expects k or n to be defined but not both.")

        def vec3(x):
            x = np.asarray(x, dtype=np.float64)
            if x.ndim != 1:
                raise B2PError("Antenna expects A and k/n to be rank-1 arrays (specifically 3D vectors).")
            if x.shape[0] != 3:
                raise B2PError("Antenna expects A and k/n to be 3D vectors.")
            return x

        self.A = vec3(A)
        self.wave = vec3(k if k is not None else n)
        self.wave_kind = 0 if k is not None else 1
        self.lap_coeffs = None
        if lap_coeffs is not None:
            c = np.asarray(lap_coeffs, dtype=np.complex128)
            if c.ndim != 1:
                raise B2PError("lap_coeffs must be 1D array.")
            self.lap_coeffs = np.ascontiguousarray(np.stack([c.real, c.imag], axis=1))

    def _as_struct(self):
        m = _abi.AntennaMode(wave_kind=self.wave_kind)
        m.A[:] = self.A
        m.wave[:] = self.wave
        if self.lap_coeffs is not None:
            m.n_lap_coeffs = len(self.lap_coeffs)
            # a valid non-NULL pointer even for an empty list: NULL means "no lap_coeffs" (phi = 1)
            self._keep = self.lap_coeffs if len(self.lap_coeffs) else np.zeros((1, 2))
            m.lap_coeffs = self._keep.ctypes.data_as(C.POINTER(C.c_double))
        return m


def reflector_wall(*, walloc=0.0, betawall=0.0, gammawall=1.0):
    """pic.threeD.reflector_wall (bindings/pypic.c++:70-90)."""
    return _abi.ReflectorWall(walloc=float(walloc), betawall=float(betawall), gammawall=float(gammawall))


def _mode(m):
    return m.value if isinstance(m, comm_mode) else int(m)


class EmfTileHost:
    """Backend-independent host logic of emf::Tile<3> (emf/tile.c++:184-335,
    bindings/pyemf.c++:32-78): where the Yee-staggered sample points are, what the
    setters validate and what the getters return.  A backend supplies
    `_backend_set_fields(E, B, J, with_halo)` / `_backend_get_fields(with_halo)` on
    fp32 component-major arrays plus `mins`, `maxs`, `n_cells`.  The CUDA backend is
    `Tile` below; tests/refshim binds the same logic to the CPU oracle."""

    error_type = B2PError

    # -- geometry -------------------------------------------------------------
    def global_coordinate_map(self):
        """emf/tile.h:215-236: (i,j,k) tile-local (fractional) index -> global coordinates."""
        mins, maxs, e = self.mins, self.maxs, self.n_cells

        def m(idx):
            return tuple(mins[d] + (idx[d] / float(e[d])) * (maxs[d] - mins[d]) for d in range(3))
        return m

    def _lattice_shape(self, with_halo):
        n = self.n_cells
        return (3,) + (tuple(v + 6 for v in n) if with_halo else tuple(n))

    # -- field setters / getters ------------------------------------------------
    def _upload(self, E, B, J, with_halo=False):
        arrs = []
        for a in (E, B, J):
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float32)
                if a.shape != self._lattice_shape(with_halo):
                    raise self.error_type("Batch field setter returned array with incorrect shape!")
            arrs.append(a)
        self._backend_set_fields(arrs[0], arrs[1], arrs[2], bool(with_halo))

    def set_EBJ(self, E, B, J):
        """emf/tile.c++:184-224: per-point callables evaluated at the Yee-staggered positions."""
        nx, ny, nz = self.n_cells
        gm = self.global_coordinate_map()
        out = np.empty((3, 3, nx, ny, nz), np.float64)
        for i in range(nx):
            for j in range(ny):
                for k in range(nz):
                    x, y, z = gm((i, j, k))
                    for f, fn in enumerate((E, B, J)):
                        if f == 1:
                            out[f, 0, i, j, k] = fn(x, y + 0.5, z + 0.5)[0]
                            out[f, 1, i, j, k] = fn(x + 0.5, y, z + 0.5)[1]
                            out[f, 2, i, j, k] = fn(x + 0.5, y + 0.5, z)[2]
                        else:
                            out[f, 0, i, j, k] = fn(x + 0.5, y, z)[0]
                            out[f, 1, i, j, k] = fn(x, y + 0.5, z)[1]
                            out[f, 2, i, j, k] = fn(x, y, z + 0.5)[2]
        self._upload(out[0], out[1], out[2])

    def batch_set_EBJ(self, Ex, Ey, Ez, Bx, By, Bz, Jx, Jy, Jz):
        """emf/tile.c++:226-335."""
        nx, ny, nz = self.n_cells
        gm = self.global_coordinate_map()
        ii, jj, kk = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64),
                                 np.arange(nz, dtype=np.float64), indexing="ij")
        x, y, z = gm((ii, jj, kk))
        xp5, yp5, zp5 = gm((ii + 0.5, jj + 0.5, kk + 0.5))
        x, y, z, xp5, yp5, zp5 = (np.ascontiguousarray(a) for a in (x, y, z, xp5, yp5, zp5))

        def shaped(a):
            a = np.asarray(a, dtype=np.float64)
            if a.shape != (nx, ny, nz):
                raise self.error_type("Batch field setter returned array with incorrect shape!")
            return a
        E = np.stack([shaped(Ex(xp5, y, z)), shaped(Ey(x, yp5, z)), shaped(Ez(x, y, zp5))])
        B = np.stack([shaped(Bx(x, yp5, zp5)), shaped(By(xp5, y, zp5)), shaped(Bz(xp5, yp5, z))])
        J = np.stack([shaped(Jx(xp5, y, z)), shaped(Jy(x, yp5, z)), shaped(Jz(x, y, zp5))])
        self._upload(E, B, J)

    def _download(self, with_halo):
        return self._backend_get_fields(bool(with_halo))

    def get_EBJ(self):
        """bindings/pyemf.c++:32-78: owning float64 copies of the fp32 interior."""
        E, B, J = self._download(False)
        return tuple(tuple(np.array(f[c], dtype=np.float64) for c in range(3)) for f in (E, B, J))

    def get_EBJ_with_halo(self):
        E, B, J = self._download(True)
        return tuple(tuple(np.array(f[c], dtype=np.float64) for c in range(3)) for f in (E, B, J))

    # raw fp32 access used by the parity tests
    def get_fields_f32(self, with_halo=True):
        return self._download(with_halo)

    def set_fields_f32(self, E=None, B=None, J=None, with_halo=True):
        self._upload(E, B, J, with_halo)



class Tile(EmfTileHost):
    """emf::Tile<3> — owns E, B, J on the device (emf/tile.h:36-213)."""

    _need_pic = False

    def __init__(self, tile_grid_idx, config):
        try:
            self._cfg = config if isinstance(config, _abi.B2PConfig) else make_config(config, need_pic=self._need_pic)
        except ConfigError as e:
            raise B2PError(str(e)) from None
        idx = (C.c_int32 * 3)(*[int(v) for v in tile_grid_idx])
        h = C.c_void_p()
        check(lib().b2p_tile_create(C.byref(self._cfg), C.byref(idx), C.byref(h)))
        self._h = h
        self.index = tuple(int(v) for v in tile_grid_idx)
        mins, maxs = (C.c_double * 3)(), (C.c_double * 3)()
        check(lib().b2p_tile_bounds(self._h, C.byref(mins), C.byref(maxs)))
        self.mins, self.maxs = list(mins), list(maxs)
        self.n_cells = tuple(self._cfg.n_cells)
        self.cid = self.index[0] + self._cfg.n_tiles[0] * (self.index[1] + self._cfg.n_tiles[1] * self.index[2])
        self.communication = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().b2p_tile_destroy(h)
            except Exception:
                pass
            self._h = None

    @staticmethod
    def canonical_type():
        return Tile

    # -- backend: C-ABI ----------------------------------------------------------
    def _backend_set_fields(self, E, B, J, with_halo):
        check(lib().b2p_tile_set_fields(self._h, _ptr(E), _ptr(B), _ptr(J), int(with_halo)))

    def _backend_get_fields(self, with_halo):
        E, B, J = (np.empty(self._lattice_shape(with_halo), np.float32) for _ in range(3))
        check(lib().b2p_tile_get_fields(self._h, _ptr(E), _ptr(B), _ptr(J), int(with_halo)))
        return E, B, J

    # -- field solver ----------------------------------------------------------
    def push_half_b(self):
        check(lib().b2p_tile_push_half_b(self._h))

    def push_e(self):
        check(lib().b2p_tile_push_e(self._h))

    def add_current(self):
        check(lib().b2p_tile_add_current(self._h))

    def filter_current(self):
        check(lib().b2p_tile_filter_current(self._h))

    def clear_current(self):
        check(lib().b2p_tile_clear_current(self._h))

    def field_energy(self):
        b, e = C.c_double(), C.c_double()
        check(lib().b2p_tile_field_energy(self._h, C.byref(b), C.byref(e)))
        return b.value, e.value

    # -- antenna (emf/tile.c++:566-791) -----------------------------------------------
    def register_antenna(self, mode):
        check(lib().b2p_tile_register_antenna(self._h, C.byref(mode._as_struct())))

    def deposit_antenna_current(self):
        check(lib().b2p_tile_deposit_antenna_current(self._h))

    # -- edge boundary conditions (emf/tile.c++:808-847) ------------------------------
    def register_edge_bc(self, bc):
        check(lib().b2p_tile_register_edge_bc(self._h, C.byref(bc)))

    def apply_edge_bcs(self, mode):
        check(lib().b2p_tile_apply_edge_bcs(self._h, _mode(mode)))

    def apply_edge_bc(self, bc, mode):
        check(lib().b2p_tile_apply_edge_bc(self._h, C.byref(bc), _mode(mode)))


class PicTileHost:
    """Backend-independent host logic of pic::Tile<3> (pic/tile.c++:180-322,
    pic/particle.c++:82-168): injection cell order and validation, getters.  A backend
    supplies `get_particles(sp, alive_only)` and `_backend_inject(sp, six float64 arrays)`."""

    def get_positions(self, sp):
        p = self.get_particles(sp)
        return p[0], p[1], p[2]

    def get_velocities(self, sp):
        p = self.get_particles(sp)
        return p[3], p[4], p[5]

    def get_ids(self, sp):
        return self.get_particles(sp)[6]

    # -- injection ---------------------------------------------------------------
    def _inject_arrays(self, sp, pos, vel):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (*pos, *vel)]
        for v in a:
            if v.ndim != 1:
                raise self.error_type("pic::Tile::batch_inject_in_x_stripe: given batch must be one dimensional.")
            if v.shape[0] != a[0].shape[0]:
                raise self.error_type("pic::Tile::batch_inject_in_x_stripe: batches must have same length.")
        self._backend_inject(int(sp), a)

    def inject(self, sp, particles):
        """pic/tile.c++:207-217"""
        pos = np.array([p.pos for p in particles], dtype=np.float64).reshape(-1, 3)
        vel = np.array([p.vel for p in particles], dtype=np.float64).reshape(-1, 3)
        self._inject_arrays(sp, pos.T, vel.T)

    def inject_to_each_cell(self, sp, pgen):
        """pic/tile.c++:180-205: generator called per cell corner, cells in i->j->k order."""
        nx, ny, nz = self.n_cells
        gm = self.global_coordinate_map()
        out = []
        for i in range(nx):
            for j in range(ny):
                for k in range(nz):
                    out.extend(pgen(*gm((i, j, k))))
        self.inject(sp, out)

    def batch_inject_to_cells(self, sp, pgen):
        self.batch_inject_in_x_stripe(sp, pgen, float(self.mins[0]), float(self.maxs[0]))

    def batch_inject_in_x_stripe(self, sp, pgen, x_left, x_right):
        """pic/tile.c++:235-322"""
        xmin, xmax = float(self.mins[0]), float(self.maxs[0])
        if x_right <= xmin or x_left >= xmax:
            return
        nx, ny, nz = self.n_cells
        dl, dr = x_left - xmin, x_right - xmin
        i_begin = 0 if dl <= 0.0 else min(nx, int(np.floor(dl)))
        i_end = nx if dr >= float(nx) else int(np.ceil(dr))
        if i_begin >= i_end:
            return
        gm = self.global_coordinate_map()
        ii, jj, kk = np.meshgrid(np.arange(i_begin, i_end, dtype=np.float64), np.arange(ny, dtype=np.float64),
                                 np.arange(nz, dtype=np.float64), indexing="ij")
        x, y, z = (np.ascontiguousarray(a.reshape(-1)) for a in gm((ii, jj, kk)))
        batch = pgen(x, y, z)
        self._inject_arrays(sp, batch.pos, batch.vel)



class PicTile(PicTileHost, Tile):
    """pic::Tile<3> (pic/tile.h:52-202)."""

    _need_pic = True

    @staticmethod
    def canonical_type():
        return PicTile

    @property
    def n_species(self):
        return self._cfg.n_species

    # -- backend: C-ABI ----------------------------------------------------------
    def container_size(self, sp):
        n = C.c_uint64()
        check(lib().b2p_tile_container_size(self._h, int(sp), C.byref(n)))
        return n.value

    def get_particles(self, sp, alive_only=True, out=None):
        """Seven arrays x,y,z,ux,uy,uz,id.  `out` (raw container only): caller-owned arrays of at least
        container_size entries — e.g. page-locked with lib().b2p_host_register — filled in place."""
        n = self.container_size(sp)
        if out is not None:
            a, ids = list(out[:6]), out[6]
        else:
            a = [np.empty(n, np.float32) for _ in range(6)]
            ids = np.empty(n, np.uint64)
        m = C.c_uint64()
        check(lib().b2p_tile_get_particles(self._h, int(sp), int(alive_only), *[_ptr(v) for v in a], _ptr(ids), C.byref(m)))
        if out is not None:
            return tuple(v[:m.value] for v in a) + (ids[:m.value],)
        return tuple(np.array(v[:m.value]) for v in a) + (np.array(ids[:m.value]),)

    def _backend_inject(self, sp, a):
        check(lib().b2p_tile_inject(self._h, sp, a[0].shape[0], *[_ptr(v) for v in a]))

    def set_particles_raw(self, sp, x, y, z, ux, uy, uz, ids):
        a = [np.ascontiguousarray(v, dtype=np.float32) for v in (x, y, z, ux, uy, uz)]
        i = np.ascontiguousarray(ids, dtype=np.uint64)
        check(lib().b2p_tile_set_particles(self._h, int(sp), len(i), *[_ptr(v) for v in a], _ptr(i)))

    # -- hot path ------------------------------------------------------------------
    def push_particles(self):
        check(lib().b2p_tile_push_particles(self._h))

    def deposit_current(self):
        check(lib().b2p_tile_deposit_current(self._h))

    def sort_particles(self):
        check(lib().b2p_tile_sort_particles(self._h))

    def pack_outgoing_particles(self):
        check(lib().b2p_tile_pack_outgoing_particles(self._h))

    def sort_keys(self, sp):
        k = np.empty(self.container_size(sp), np.uint32)
        check(lib().b2p_tile_sort_keys(self._h, int(sp), _ptr(k)))
        return k

    def get_outgoing(self):
        n = C.c_uint64()
        ends = np.zeros(27 * self.n_species, np.uint64)
        check(lib().b2p_tile_get_outgoing(self._h, None, 0, _ptr(ends), C.byref(n)))
        buf = np.zeros(n.value, dtype=np.dtype([("pos", np.float32, 3), ("vel", np.float32, 3), ("id", np.uint64)]))
        check(lib().b2p_tile_get_outgoing(self._h, _ptr(buf), n.value, _ptr(ends), C.byref(n)))
        return buf, ends

    def kinetic_energy(self, sp):
        e, n = C.c_double(), C.c_uint64()
        check(lib().b2p_tile_kinetic_energy(self._h, int(sp), C.byref(e), C.byref(n)))
        return e.value, n.value

    # -- reflector wall (pic/reflector_wall.c++:226-297) --------------------------------
    def register_reflector_wall(self, wall):
        check(lib().b2p_tile_register_reflector_wall(self._h, C.byref(wall)))

    def reflect_particles(self):
        check(lib().b2p_tile_reflect_particles(self._h))

    def advance_reflector_walls(self):
        check(lib().b2p_tile_advance_reflector_walls(self._h))

    def reflector_walls(self):
        n = C.c_uint64()
        out = (_abi.ReflectorWall * 16)()
        check(lib().b2p_tile_reflector_walls(self._h, out, 16, C.byref(n)))
        return [(w.walloc, w.betawall, w.gammawall) for w in out[:n.value]]


class Grid:
    """The slice of corgi::Grid<3> the PIC lap uses, batched per phase on the device."""

    def __init__(self, config):
        try:
            self._cfg = config if isinstance(config, _abi.B2PConfig) else make_config(config)
        except ConfigError as e:
            raise B2PError(str(e)) from None
        h = C.c_void_p()
        check(lib().b2p_grid_create(C.byref(self._cfg), C.byref(h)))
        self._h = h
        self._tiles = {}
        self.n_species = self._cfg.n_species

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().b2p_grid_destroy(h)
            except Exception:
                pass
            self._h = None

    def add_tile(self, tile, idx=None):
        check(lib().b2p_grid_add_tile(self._h, tile._h))
        self._tiles[tile.index] = tile   # keep_alive<1,2> (pycorgi.c++:333)

    def get_tile(self, i, j, k):
        return self._tiles[(i, j, k)]

    def get_local_tiles(self):
        return list(self._tiles.values())

    def local_communication(self, mode):
        check(lib().b2p_grid_local_communication(self._h, mode.value if isinstance(mode, comm_mode) else int(mode)))

    def external_communication(self, mode):
        check(lib().b2p_grid_external_communication(self._h, mode.value if isinstance(mode, comm_mode) else int(mode)))

    def phase(self, name):
        check(getattr(lib(), "b2p_grid_" + name)(self._h))

    def apply_edge_bcs(self, mode):
        check(lib().b2p_grid_apply_edge_bcs(self._h, _mode(mode)))

    def write_fields_snapshot(self, prefix, lap, stride=1, nspecies=2):
        """io_emf_snapshot: "<prefix>/flds_<lap>.bin" in the reference's RNKO v3 format (mpiio_fields.c++)."""
        check(lib().b2p_grid_write_fields_snapshot(self._h, str(prefix).encode(), int(lap), int(stride), int(nspecies)))

    def step_pic(self, lap):
        check(lib().b2p_grid_step_pic(self._h, int(lap)))

    def step_shock(self, lap, n_filter_passes=3):
        """One lap of projects/pic-shock/pic.py:229-279 (injector and IO excluded), phase by phase on
        the device: edge BCs after every field update, reflector wall between push and pack."""
        M = comm_mode
        multi = getattr(self, "_multi", False)

        def comm(m, local=None):
            if multi:
                self.external_communication(m)
            self.local_communication(m if local is None else local)

        self.phase("push_half_b"); self.apply_edge_bcs(M.emf_B); comm(M.emf_B)
        self.phase("push_particles"); self.phase("reflect_particles"); self.phase("pack_outgoing_particles")
        comm(M.pic_particle)
        if lap % 5 == 0:
            self.phase("sort_particles")
        self.phase("deposit_current")
        comm(M.emf_J, M.emf_J_exchange); comm(M.emf_J)
        self.apply_edge_bcs(M.emf_J)
        for i in range(n_filter_passes):
            if i > 0 and i % 3 == 0:
                comm(M.emf_J)
            self.phase("filter_current")
        self.apply_edge_bcs(M.emf_J)
        self.phase("push_half_b"); self.apply_edge_bcs(M.emf_B); comm(M.emf_B)
        self.phase("push_e"); self.apply_edge_bcs(M.emf_E)
        self.phase("add_current"); self.apply_edge_bcs(M.emf_E); comm(M.emf_E)
        self.phase("advance_reflector_walls")

    def step_emf(self):
        check(lib().b2p_grid_step_emf(self._h))

    def energies(self):
        b, e = C.c_double(), C.c_double()
        k = np.zeros(max(1, self.n_species), np.float64)
        s = np.zeros(max(1, self.n_species), np.uint64)
        check(lib().b2p_grid_energies(self._h, C.byref(b), C.byref(e), _ptr(k), _ptr(s)))
        return b.value, e.value, k[:self.n_species], s[:self.n_species]

    def alive_counts(self):
        c = np.zeros(max(1, self.n_species), np.uint64)
        check(lib().b2p_grid_alive_counts(self._h, _ptr(c)))
        return c[:self.n_species]

    def inject_thermal(self, ppc, delgam, seed=1):
        check(lib().b2p_grid_inject_thermal(self._h, int(ppc), float(delgam), int(seed)))

    def inject_drifting_stripe(self, sp, ppc, delgam, gamma_drift, dir_sign, x_left, x_right, seed=1):
        check(lib().b2p_grid_inject_drifting_stripe(self._h, int(sp), int(ppc), float(delgam), float(gamma_drift), int(dir_sign),
                                                    float(x_left), float(x_right), int(seed)))

    def set_uniform_B(self, bx, by, bz):
        check(lib().b2p_grid_set_uniform_B(self._h, bx, by, bz))


class MpiioFieldsWriter:
    """emf.threeD.MpiioFieldsWriter (bindings/pyemf.c++:275-281): same constructor arguments; `write(grid, lap)`
    and `write_collective(grid, lap)` both place every local tile with pwrite()."""

    def __init__(self, prefix, Nx, NxMesh, Ny, NyMesh, Nz, NzMesh, stride, nspecies=2):
        self.prefix, self.stride, self.nspecies = str(prefix), int(stride), int(nspecies)
        self.dims = (int(Nx), int(NxMesh), int(Ny), int(NyMesh), int(Nz), int(NzMesh))

    def write(self, grid, lap):
        c = grid._cfg
        if (c.n_tiles[0], c.n_cells[0], c.n_tiles[1], c.n_cells[1], c.n_tiles[2], c.n_cells[2]) != self.dims:
            raise B2PError("MpiioFieldsWriter: grid dimensions differ from the writer's")
        grid.write_fields_snapshot(self.prefix, lap, self.stride, self.nspecies)
        return True

    write_collective = write


def sync():
    check(lib().b2p_sync())
