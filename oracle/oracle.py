"""ctypes front-end of the CPU oracle (oracle/pic_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs — never by runko_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from runko_b200._abi import AntennaMode, B2PConfig, EdgeBC, ParticleState, ReflectorWall, make_config  # struct layouts only

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libpic_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "pic_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


_lib = None
f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C")


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_last_error.restype = C.c_char_p
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(B2PConfig)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_num_tiles.argtypes = [C.c_void_p]
        L.orc_tile_cid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        vp, ci = C.c_void_p, C.c_int
        for name in ("push_half_b", "push_e", "add_current", "filter_current", "clear_current",
                     "push_particles", "deposit_current", "sort_particles", "pack_outgoing_particles",
                     "reflect_particles", "advance_reflector_walls", "deposit_antenna_current"):
            getattr(L, "orc_tile_" + name).argtypes = [vp, ci]
        L.orc_tile_set_fields.argtypes = [vp, ci, vp, vp, vp, ci]
        L.orc_tile_get_fields.argtypes = [vp, ci, vp, vp, vp, ci]
        L.orc_tile_field_energy.argtypes = [vp, ci, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_tile_inject.argtypes = [vp, ci, ci, C.c_uint64] + [vp] * 6
        L.orc_tile_set_particles.argtypes = [vp, ci, ci, C.c_uint64] + [vp] * 7
        L.orc_tile_container_size.argtypes = [vp, ci, ci, C.POINTER(C.c_uint64)]
        L.orc_tile_get_particles.argtypes = [vp, ci, ci, ci] + [vp] * 7 + [C.POINTER(C.c_uint64)]
        L.orc_tile_sort_keys.argtypes = [vp, ci, ci, vp]
        L.orc_tile_get_outgoing.argtypes = [vp, ci, vp, C.c_uint64, vp, C.POINTER(C.c_uint64)]
        L.orc_tile_kinetic_energy.argtypes = [vp, ci, ci, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.orc_tile_interpolate.argtypes = [vp, ci, C.c_uint64, vp, vp, vp, vp]
        L.orc_tile_register_antenna.argtypes = [vp, ci, C.POINTER(AntennaMode)]
        L.orc_tile_register_edge_bc.argtypes = [vp, ci, C.POINTER(EdgeBC)]
        L.orc_tile_apply_edge_bcs.argtypes = [vp, ci, ci]
        L.orc_tile_apply_edge_bc.argtypes = [vp, ci, C.POINTER(EdgeBC), ci]
        L.orc_tile_register_reflector_wall.argtypes = [vp, ci, C.POINTER(ReflectorWall)]
        L.orc_tile_reflector_walls.argtypes = [vp, ci, vp, C.c_uint64, C.POINTER(C.c_uint64)]
        L.orc_local_communication.argtypes = [vp, ci]
        L.orc_grid_phase.argtypes = [vp, C.c_char_p, ci]
        L.orc_step_pic.argtypes = [vp, C.c_int64, ci]
        L.orc_step_emf.argtypes = [vp, ci]
        L.orc_energies.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), vp, vp]
        L.orc_write_fields_snapshot.argtypes = [vp, C.c_char_p, C.c_int32, C.c_int32, C.c_int32]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleGrid:
    """All tiles of the global periodic grid, CPU, single process."""

    def __init__(self, conf):
        self.cfg = conf if isinstance(conf, B2PConfig) else make_config(conf)
        self._L = lib()
        self._g = self._L.orc_create(C.byref(self.cfg))
        if not self._g:
            raise OracleError(self._L.orc_last_error().decode())
        self.n_cells = tuple(self.cfg.n_cells)
        self.n_tiles = tuple(self.cfg.n_tiles)
        self.n_species = self.cfg.n_species

    def __del__(self):
        if getattr(self, "_g", None):
            self._L.orc_destroy(self._g)
            self._g = None

    def _ck(self, rc):
        if rc:
            raise OracleError(self._L.orc_last_error().decode())

    def cid(self, i, j, k):
        return self._L.orc_tile_cid(self._g, i, j, k)

    @property
    def num_tiles(self):
        return self._L.orc_num_tiles(self._g)

    def shape(self, with_halo):
        n = self.n_cells
        return (3,) + (tuple(x + 6 for x in n) if with_halo else tuple(n))

    # fields -----------------------------------------------------------------
    def set_fields(self, t, E=None, B=None, J=None, with_halo=False):
        arrs = []
        for a in (E, B, J):
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float32)
                assert a.shape == self.shape(with_halo), (a.shape, self.shape(with_halo))
            arrs.append(a)
        self._ck(self._L.orc_tile_set_fields(self._g, t, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), int(with_halo)))

    def get_fields(self, t, with_halo=False):
        E, B, J = (np.empty(self.shape(with_halo), np.float32) for _ in range(3))
        self._ck(self._L.orc_tile_get_fields(self._g, t, _p(E), _p(B), _p(J), int(with_halo)))
        return E, B, J

    def tile_op(self, t, name):
        self._ck(getattr(self._L, "orc_tile_" + name)(self._g, t))

    def field_energy(self, t):
        b, e = C.c_double(), C.c_double()
        self._ck(self._L.orc_tile_field_energy(self._g, t, C.byref(b), C.byref(e)))
        return b.value, e.value

    # particles --------------------------------------------------------------
    def inject(self, t, sp, x, y, z, ux, uy, uz):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, y, z, ux, uy, uz)]
        self._ck(self._L.orc_tile_inject(self._g, t, sp, len(a[0]), *[_p(v) for v in a]))

    def set_particles(self, t, sp, x, y, z, ux, uy, uz, ids):
        a = [np.ascontiguousarray(v, dtype=np.float32) for v in (x, y, z, ux, uy, uz)]
        i = np.ascontiguousarray(ids, dtype=np.uint64)
        self._ck(self._L.orc_tile_set_particles(self._g, t, sp, len(i), *[_p(v) for v in a], _p(i)))

    def container_size(self, t, sp):
        n = C.c_uint64()
        self._ck(self._L.orc_tile_container_size(self._g, t, sp, C.byref(n)))
        return n.value

    def get_particles(self, t, sp, alive_only=True):
        n = self.container_size(t, sp)
        a = [np.empty(n, np.float32) for _ in range(6)]
        ids = np.empty(n, np.uint64)
        m = C.c_uint64()
        self._ck(self._L.orc_tile_get_particles(self._g, t, sp, int(alive_only), *[_p(v) for v in a], _p(ids), C.byref(m)))
        return tuple(v[:m.value] for v in a) + (ids[:m.value],)

    def sort_keys(self, t, sp):
        k = np.empty(self.container_size(t, sp), np.uint32)
        self._ck(self._L.orc_tile_sort_keys(self._g, t, sp, _p(k)))
        return k

    def get_outgoing(self, t):
        n = C.c_uint64()
        ends = np.zeros(27 * self.n_species, np.uint64)
        self._ck(self._L.orc_tile_get_outgoing(self._g, t, None, 0, _p(ends), C.byref(n)))
        buf = np.zeros(n.value, dtype=np.dtype([("pos", np.float32, 3), ("vel", np.float32, 3), ("id", np.uint64)]))
        self._ck(self._L.orc_tile_get_outgoing(self._g, t, _p(buf), n.value, _p(ends), C.byref(n)))
        return buf, ends

    def kinetic_energy(self, t, sp):
        e, n = C.c_double(), C.c_uint64()
        self._ck(self._L.orc_tile_kinetic_energy(self._g, t, sp, C.byref(e), C.byref(n)))
        return e.value, n.value

    def interpolate(self, t, x, y, z):
        a = [np.ascontiguousarray(v, dtype=np.float32) for v in (x, y, z)]
        out = np.empty((len(a[0]), 6), np.float32)
        self._ck(self._L.orc_tile_interpolate(self._g, t, len(a[0]), *[_p(v) for v in a], _p(out)))
        return out

    # pic-shock boundary pieces ---------------------------------------------------
    def register_edge_bc(self, t, bc):
        self._ck(self._L.orc_tile_register_edge_bc(self._g, t, C.byref(bc)))

    def apply_edge_bcs(self, t, mode):
        self._ck(self._L.orc_tile_apply_edge_bcs(self._g, t, int(mode)))

    def apply_edge_bc(self, t, bc, mode):
        self._ck(self._L.orc_tile_apply_edge_bc(self._g, t, C.byref(bc), int(mode)))

    def register_antenna(self, t, mode):
        self._ck(self._L.orc_tile_register_antenna(self._g, t, C.byref(mode._as_struct())))

    def register_reflector_wall(self, t, wall):
        self._ck(self._L.orc_tile_register_reflector_wall(self._g, t, C.byref(wall)))

    def reflector_walls(self, t):
        n = C.c_uint64()
        out = (ReflectorWall * 16)()
        self._ck(self._L.orc_tile_reflector_walls(self._g, t, out, 16, C.byref(n)))
        return [(w.walloc, w.betawall, w.gammawall) for w in out[:n.value]]

    # grid -------------------------------------------------------------------
    def local_communication(self, mode):
        self._ck(self._L.orc_local_communication(self._g, int(mode)))

    def phase(self, name, threads=1):
        self._ck(self._L.orc_grid_phase(self._g, name.encode(), threads))

    def step_pic(self, lap, threads=1):
        self._ck(self._L.orc_step_pic(self._g, lap, threads))

    def step_emf(self, threads=1):
        self._ck(self._L.orc_step_emf(self._g, threads))

    def write_fields_snapshot(self, prefix, lap, stride=1, nspecies=2):
        self._ck(self._L.orc_write_fields_snapshot(self._g, str(prefix).encode(), int(lap), int(stride), int(nspecies)))

    def energies(self):
        b, e = C.c_double(), C.c_double()
        k = np.zeros(max(1, self.n_species), np.float64)
        s = np.zeros(max(1, self.n_species), np.uint64)
        self._ck(self._L.orc_energies(self._g, C.byref(b), C.byref(e), _p(k), _p(s)))
        return b.value, e.value, k[:self.n_species], s[:self.n_species]
