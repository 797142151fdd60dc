"""ctypes mirror of include/b200pic.h (struct layouts and enums only; no compute)."""
import ctypes as C

MAX_SPECIES = 8
HALO = 3
DEAD_ID = 0xFFFFFFFFFFFFFFFF

COMM_EMF_J, COMM_EMF_E, COMM_EMF_B, COMM_PIC_PARTICLE = 0, 1, 2, 3
COMM_PIC_PARTICLE_EXTRA, COMM_NUMBER_OF_PARTICLES, COMM_EMF_J_EXCHANGE = 4, 5, 6

PROPAGATORS = {"fdtd2": 0, "stencil": 1}
FILTERS = {None: -1, "binomial2": 0, "binomial2_unrolled": 1}
PUSHERS = {None: -1, "boris": 0, "higuera_cary": 1, "faraday": 2}
INTERPOLATORS = {"linear_1st": 0, "linear_1st_unrolled": 1}
DEPOSITERS = {"zigzag": 0, "zigzag_1st": 0, "zigzag_1st_atomic": 1}

# emf/tile.c++:112-128: config key suffix -> (row, col) of StencilAxisCoeffs::M
STENCIL_KEYS = {
    "delta": (1, 0), "gamma": (2, 0),
    "beta_p1": (0, 1), "beta_p2": (0, 2), "beta2_p1": (1, 1), "beta2_p2": (1, 2),
    "beta3_p1": (2, 1), "beta3_p2": (2, 2),
    "zeta_p1": (0, 3), "zeta_p2": (0, 4), "zeta2_p1": (1, 3), "zeta2_p2": (1, 4),
    "zeta3_p1": (2, 3), "zeta3_p2": (2, 4),
}


class B2PConfig(C.Structure):
    _fields_ = [
        ("n_tiles", C.c_int32 * 3),
        ("n_cells", C.c_int32 * 3),
        ("cfl", C.c_double),
        ("field_propagator", C.c_int32),
        ("current_filter", C.c_int32),
        ("stencil", ((C.c_float * 5) * 3) * 3),
        ("n_species", C.c_int32),
        ("q", C.c_double * MAX_SPECIES),
        ("m", C.c_double * MAX_SPECIES),
        ("particle_pusher", C.c_int32),
        ("field_interpolator", C.c_int32),
        ("current_depositer", C.c_int32),
        ("prealloc_per_species", C.c_uint64),
    ]


class ParticleState(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("vel", C.c_float * 3), ("id", C.c_uint64)]


assert C.sizeof(ParticleState) == 32


class EdgeBC(C.Structure):
    """b2p_edge_bc == emf::edge_bc (src/runko/emf/edge_bc.h:25-40)"""
    _fields_ = [("direction", C.c_uint8), ("side", C.c_uint8), ("position", C.c_float),
                ("E", C.c_float * 3), ("B", C.c_float * 3), ("J", C.c_float * 3),
                ("E_components", C.c_uint8), ("B_components", C.c_uint8), ("J_components", C.c_uint8)]


class AntennaMode(C.Structure):
    """b2p_antenna_mode == emf::antenna_mode flattened (src/runko/emf/antenna.h:31-45)"""
    _fields_ = [("A", C.c_double * 3), ("wave", C.c_double * 3), ("wave_kind", C.c_int32),
                ("n_lap_coeffs", C.c_uint64), ("lap_coeffs", C.POINTER(C.c_double))]


class ReflectorWall(C.Structure):
    """b2p_reflector_wall == pic::reflector_wall (src/runko/pic/reflector_wall.h:14-25)"""
    _fields_ = [("walloc", C.c_float), ("betawall", C.c_float), ("gammawall", C.c_float)]


class ConfigError(RuntimeError):
    pass


def _get(conf, key):
    """toolbox::ConfigParser::get (tools/config_parser.c++:26-92): reads obj.__dict__;
    a missing key or a None value reads as absent."""
    d = conf if isinstance(conf, dict) else getattr(conf, "__dict__", {})
    return d.get(key, None)


def _typed(conf, key, kind, required):
    v = _get(conf, key)
    if v is None:
        if required:
            raise ConfigError(f"Configuration is missing required key: {key}")
        return None
    if kind == "double":
        if isinstance(v, bool) or not isinstance(v, (int, float)):
            raise ConfigError(f"{key}: unsupported type {type(v).__name__} for value {v!r}")
        return float(v)
    if kind == "string":
        if not isinstance(v, str):
            raise ConfigError(f"{key}: unsupported type {type(v).__name__} for value {v!r}")
        return v
    if kind == "size":
        if isinstance(v, bool) or not isinstance(v, int):
            raise ConfigError(f"{key}: unsupported type {type(v).__name__} for value {v!r}")
        return int(v)
    if kind == "ivec3":
        try:
            vals = [int(x) for x in v]
        except Exception:
            raise ConfigError(f"{key}: unsupported type {type(v).__name__} for value {v!r}")
        if len(vals) != 3:
            raise ConfigError(f"{key}: expected 3 values, got {len(vals)}")
        return vals
    raise AssertionError(kind)


def make_config(conf, need_pic=None):
    """Flatten a runko Configuration-like object (or dict) into a B2PConfig with the
    reference's key names and error behaviour (emf/tile.c++:82-142, pic/tile.c++:28-129)."""
    c = B2PConfig()
    cells = _typed(conf, "n_cells_per_tile", "ivec3", True)
    c.cfl = _typed(conf, "cfl", "double", True)
    prop = _typed(conf, "field_propagator", "string", True)
    if prop not in PROPAGATORS:
        raise ConfigError(f"{prop} is not supported field propagator.")
    c.field_propagator = PROPAGATORS[prop]
    filt = _typed(conf, "current_filter", "string", False)
    if filt is not None and filt not in FILTERS:
        raise ConfigError(f"{filt} is not supported current filter.")
    c.current_filter = FILTERS[filt]
    if prop == "stencil":
        for a, ax in enumerate("xyz"):
            for name, (r, col) in STENCIL_KEYS.items():
                v = _typed(conf, f"stencil_{ax}_{name}", "double", False)
                if v is None:
                    v = _typed(conf, f"stencil_{name}", "double", False)
                c.stencil[a][r][col] = 0.0 if v is None else v
    tiles = _typed(conf, "n_tiles", "ivec3", True)
    for d in range(3):
        c.n_tiles[d] = tiles[d]
        c.n_cells[d] = cells[d]
    pusher = _typed(conf, "particle_pusher", "string", bool(need_pic))
    if need_pic is None:
        need_pic = pusher is not None
    c.n_species = 0
    c.particle_pusher = -1
    if need_pic:
        if pusher not in PUSHERS or pusher is None:
            raise ConfigError(f"{pusher} is not supported particle pusher.")
        c.particle_pusher = PUSHERS[pusher]
        interp = _typed(conf, "field_interpolator", "string", True)
        if interp not in INTERPOLATORS:
            raise ConfigError(f"{interp} is not supported field_interpolator.")
        c.field_interpolator = INTERPOLATORS[interp]
        dep = _typed(conf, "current_depositer", "string", True)
        if dep not in DEPOSITERS:
            raise ConfigError(f"{dep} is not supported current depositer.")
        c.current_depositer = DEPOSITERS[dep]
        n = 0
        while n < MAX_SPECIES:
            q = _typed(conf, f"q{n}", "double", False)
            m = _typed(conf, f"m{n}", "double", False)
            if q is None or m is None:
                break
            c.q[n], c.m[n] = q, m
            n += 1
        c.n_species = n
        pre = _typed(conf, "prealloc_per_species", "size", False)
        c.prealloc_per_species = 0 if pre is None else pre
    return c
