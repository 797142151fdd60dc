"""ctypes front-end of oracle/_ref/libref_kernels.so — the REFERENCE'S OWN kernel sources
(emf::YeeLattice, pic::ParticleContainer) compiled from /root/reference by oracle/Makefile.ref.

TEST INFRASTRUCTURE ONLY: used to pin the oracle restatement bit for bit
(tests/test_oracle_vs_reference_build.py) and, optionally, as bench.py's CPU baseline.
"""
import ctypes as C
import itertools
import os
import subprocess

import numpy as np

from runko_b200._abi import B2PConfig, EdgeBC, ReflectorWall, make_config

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_kernels.so")
REF = "/root/reference"


def available():
    return os.path.exists(SO) and cpu_ok()


def cpu_ok():
    """libref_kernels.so is compiled with -mavx2 (oracle/Makefile.ref)."""
    try:
        with open("/proc/cpuinfo") as f:
            return " avx2 " in f.read().replace("\n", " ")
    except OSError:
        return False


def build():
    """Only possible where /root/reference is mounted (the build container)."""
    if not os.path.isdir(REF):
        return False
    subprocess.check_call(["make", "-C", _HERE, "-f", "Makefile.ref", "-j8"], stdout=subprocess.DEVNULL)
    return True


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(SO)
        L.ref_last_error.restype = C.c_char_p
        L.ref_tile_create.restype = C.c_void_p
        L.ref_tile_create.argtypes = [C.POINTER(B2PConfig), C.POINTER(C.c_int32 * 3)]
        L.ref_tile_destroy.argtypes = [C.c_void_p]
        L.ref_tile_destroy.restype = None
        vp = C.c_void_p
        L.ref_tile_set_fields.argtypes = [vp, vp, vp, vp]
        L.ref_tile_get_fields.argtypes = [vp, vp, vp, vp]
        L.ref_tile_set_particles.argtypes = [vp, C.c_int, C.c_uint64] + [vp] * 7
        L.ref_tile_container_size.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64)]
        L.ref_tile_get_particles.argtypes = [vp, C.c_int] + [vp] * 7
        L.ref_tile_op.argtypes = [vp, C.c_char_p]
        L.ref_tile_get_outgoing.argtypes = [vp, vp, C.c_uint64, vp, C.POINTER(C.c_uint64)]
        L.ref_tile_append.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp]
        L.ref_tile_halo.argtypes = [vp, vp, C.POINTER(C.c_int32 * 3), C.c_int]
        L.ref_tile_energies.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), vp]
        L.ref_tile_register_edge_bc.argtypes = [vp, C.POINTER(EdgeBC)]
        L.ref_tile_apply_edge_bc.argtypes = [vp, C.POINTER(EdgeBC), C.c_int]
        L.ref_tile_apply_edge_bcs.argtypes = [vp, C.c_int]
        L.ref_tile_register_reflector_wall.argtypes = [vp, C.POINTER(ReflectorWall)]
        L.ref_tile_reflector_walls.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


class RefError(RuntimeError):
    pass


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


STATE = np.dtype([("pos", np.float32, 3), ("vel", np.float32, 3), ("id", np.uint64)])


class RefTile:
    """One reference emf::YeeLattice + its pic::ParticleContainers."""

    def __init__(self, conf, idx=(0, 0, 0)):
        self.cfg = conf if isinstance(conf, B2PConfig) else make_config(conf)
        self.L = lib()
        i = (C.c_int32 * 3)(*idx)
        self.h = self.L.ref_tile_create(C.byref(self.cfg), C.byref(i))
        if not self.h:
            raise RefError(self.L.ref_last_error().decode())
        self.idx = tuple(idx)
        self.n_cells = tuple(self.cfg.n_cells)
        self.n_species = self.cfg.n_species
        self.mins = [float(idx[d] * self.n_cells[d]) for d in range(3)]
        self.maxs = [float((idx[d] + 1) * self.n_cells[d]) for d in range(3)]

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_tile_destroy(self.h)
            self.h = None

    def _ck(self, rc):
        if rc:
            raise RefError(self.L.ref_last_error().decode())

    def shape(self):
        return (3,) + tuple(n + 6 for n in self.n_cells)

    def set_fields(self, E=None, B=None, J=None):
        a = [None if v is None else np.ascontiguousarray(v, np.float32) for v in (E, B, J)]
        for v in a:
            assert v is None or v.shape == self.shape()
        self._ck(self.L.ref_tile_set_fields(self.h, _p(a[0]), _p(a[1]), _p(a[2])))

    def get_fields(self):
        E, B, J = (np.empty(self.shape(), np.float32) for _ in range(3))
        self._ck(self.L.ref_tile_get_fields(self.h, _p(E), _p(B), _p(J)))
        return E, B, J

    def set_particles(self, sp, x, y, z, ux, uy, uz, ids):
        a = [np.ascontiguousarray(v, np.float32) for v in (x, y, z, ux, uy, uz)]
        i = np.ascontiguousarray(ids, np.uint64)
        self._ck(self.L.ref_tile_set_particles(self.h, sp, len(i), *[_p(v) for v in a], _p(i)))

    def container_size(self, sp):
        n = C.c_uint64()
        self._ck(self.L.ref_tile_container_size(self.h, sp, C.byref(n)))
        return n.value

    def get_particles(self, sp, alive_only=False):
        n = self.container_size(sp)
        a = [np.empty(n, np.float32) for _ in range(6)]
        ids = np.empty(n, np.uint64)
        self._ck(self.L.ref_tile_get_particles(self.h, sp, *[_p(v) for v in a], _p(ids)))
        if alive_only:
            m = ids != np.uint64(0xFFFFFFFFFFFFFFFF)
            return tuple(v[m] for v in a) + (ids[m],)
        return tuple(a) + (ids,)

    def op(self, name):
        self._ck(self.L.ref_tile_op(self.h, name.encode()))

    def get_outgoing(self):
        n = C.c_uint64()
        ends = np.zeros(27 * self.n_species, np.uint64)
        self._ck(self.L.ref_tile_get_outgoing(self.h, None, 0, _p(ends), C.byref(n)))
        buf = np.zeros(n.value, STATE)
        self._ck(self.L.ref_tile_get_outgoing(self.h, _p(buf), n.value, _p(ends), C.byref(n)))
        return buf, ends

    def append(self, sp, spans, wrap=None):
        spans = [np.ascontiguousarray(s, STATE) for s in spans]
        ptrs = (C.c_void_p * len(spans))(*[s.ctypes.data for s in spans])
        counts = np.array([len(s) for s in spans], np.uint64)
        if wrap is None:
            self._ck(self.L.ref_tile_append(self.h, sp, len(spans), ptrs, _p(counts), 0, None, None))
        else:
            lo, hi = (np.asarray(v, np.float32) for v in wrap)
            self._ck(self.L.ref_tile_append(self.h, sp, len(spans), ptrs, _p(counts), 1, _p(lo), _p(hi)))

    def register_edge_bc(self, bc):
        self._ck(self.L.ref_tile_register_edge_bc(self.h, C.byref(bc)))

    def apply_edge_bc(self, bc, mode):
        self._ck(self.L.ref_tile_apply_edge_bc(self.h, C.byref(bc), int(mode)))

    def apply_edge_bcs(self, mode):
        self._ck(self.L.ref_tile_apply_edge_bcs(self.h, int(mode)))

    def register_reflector_wall(self, wall):
        self._ck(self.L.ref_tile_register_reflector_wall(self.h, C.byref(wall)))

    def reflector_walls(self):
        n = C.c_uint64()
        out = (ReflectorWall * 16)()
        self._ck(self.L.ref_tile_reflector_walls(self.h, out, 16, C.byref(n)))
        return [(w.walloc, w.betawall, w.gammawall) for w in out[:n.value]]

    def energies(self):
        b, e = C.c_double(), C.c_double()
        k = np.zeros(max(1, self.n_species))
        self._ck(self.L.ref_tile_energies(self.h, C.byref(b), C.byref(e), _p(k)))
        return b.value, e.value, k[:self.n_species]


class RefGrid:
    """All tiles of a periodic grid on the reference kernels, with corgi's local_communication
    order restated (external/corgi/src/corgi/corgi.h:1697-1718, cellular_automata.h:48-62,
    emf/tile.c++:478-542, pic/tile_communication.c++:121-195)."""

    def __init__(self, conf):
        self.cfg = make_config(conf)
        self.T = tuple(self.cfg.n_tiles)
        self.n = tuple(self.cfg.n_cells)
        self.tiles = {}
        for k, j, i in itertools.product(range(self.T[2]), range(self.T[1]), range(self.T[0])):
            self.tiles[(i, j, k)] = RefTile(self.cfg, (i, j, k))
        self.n_species = self.cfg.n_species

    def neighbour(self, idx, d):
        return tuple((idx[a] + d[a]) % self.T[a] for a in range(3))

    @staticmethod
    def moore():
        for kr, jr, ir in itertools.product((-1, 0, 1), repeat=3):
            if (ir, jr, kr) != (0, 0, 0):
                yield (ir, jr, kr)

    def phase(self, name):
        for t in self.tiles.values():
            t.op(name)

    def local_communication(self, mode):
        L = lib()
        if mode in (0, 1, 2, 6):
            for d in self.moore():
                dd = (C.c_int32 * 3)(*d)
                for idx, t in self.tiles.items():
                    o = self.tiles[self.neighbour(idx, d)]
                    t._ck(L.ref_tile_halo(t.h, o.h, C.byref(dd), mode))
            return
        assert mode == 3
        incoming = {idx: [[] for _ in range(self.n_species)] for idx in self.tiles}
        outs = {idx: t.get_outgoing() for idx, t in self.tiles.items()}
        for d in self.moore():
            inv = ((-d[0] + 1) * 3 + (-d[1] + 1)) * 3 + (-d[2] + 1)
            for idx in self.tiles:
                buf, ends = outs[self.neighbour(idx, d)]
                for sp in range(self.n_species):
                    q = 27 * sp + inv
                    b = 0 if q == 0 else int(ends[q - 1])
                    incoming[idx][sp].append(buf[b:int(ends[q])])
        lo = np.zeros(3, np.float32)
        hi = np.array([self.T[a] * self.n[a] for a in range(3)], np.float32)
        for idx, t in self.tiles.items():
            for sp in range(self.n_species):
                t.append(sp, incoming[idx][sp], wrap=(lo, hi))

    def step_pic(self, lap):
        """projects/pic-turbulence/pic.py:187-221"""
        P, C_ = self.phase, self.local_communication
        P("push_half_b"); C_(2)
        P("push_particles"); P("pack_outgoing_particles"); C_(3)
        if lap % 5 == 0:
            P("sort_particles")
        P("deposit_current")
        C_(6); C_(0)
        if self.cfg.current_filter >= 0:
            P("filter_current"); C_(0); P("filter_current"); P("filter_current")
        P("push_half_b"); C_(2)
        P("push_e"); P("add_current"); C_(1)
