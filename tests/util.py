"""Shared helpers of the parity tests: one config, the same seeded inputs for the
CPU oracle and the CUDA path."""
import numpy as np

DEAD = np.uint64(0xFFFFFFFFFFFFFFFF)


class Conf:
    """Stand-in for runko.Configuration: attributes are read through __dict__."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def pic_conf(n_tiles=(1, 1, 1), n_cells=(10, 11, 13), **kw):
    d = dict(n_tiles=list(n_tiles), n_cells_per_tile=list(n_cells), cfl=0.45, field_propagator="fdtd2",
             current_filter="binomial2", q0=-1.0, m0=1.0, q1=1.0, m1=1.0, particle_pusher="boris",
             field_interpolator="linear_1st", current_depositer="zigzag_1st_atomic")
    d.update(kw)
    return Conf(**d)


def emf_conf(n_tiles=(1, 1, 1), n_cells=(10, 11, 13), **kw):
    d = dict(n_tiles=list(n_tiles), n_cells_per_tile=list(n_cells), cfl=0.45, field_propagator="fdtd2")
    d.update(kw)
    return Conf(**d)


def random_lattice(rng, n_cells, scale=1.0):
    shape = (3,) + tuple(n + 6 for n in n_cells)
    return (scale * rng.standard_normal(shape)).astype(np.float32)


def random_particles(rng, n, mins, maxs, u_scale=0.5, dead_frac=0.0, tag=0, margin=0.0):
    """n particles uniformly inside [mins-margin, maxs+margin), thermal-ish momenta, unique ids."""
    mins = np.asarray(mins, np.float64) - margin
    maxs = np.asarray(maxs, np.float64) + margin
    pos = (mins[:, None] + rng.random((3, n)) * (maxs - mins)[:, None]).astype(np.float32)
    # keep strictly inside after fp32 rounding
    pos = np.minimum(pos, np.nextafter(maxs.astype(np.float32), np.float32(-np.inf))[:, None])
    vel = (u_scale * rng.standard_normal((3, n))).astype(np.float32)
    ids = (np.uint64(tag) << np.uint64(40)) | np.arange(n, dtype=np.uint64)
    if dead_frac > 0:
        dead = rng.random(n) < dead_frac
        ids[dead] = DEAD
    return pos, vel, ids


def ulp_diff(a, b):
    """max distance in units-in-the-last-place between two fp32 arrays"""
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, np.int64(-2**31) - a, a)
    b = np.where(b < 0, np.int64(-2**31) - b, b)
    return int(np.max(np.abs(a - b))) if a.size else 0


def assert_bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.dtype == np.float32:
        same = a.view(np.uint32) == b.view(np.uint32)
        # +0 / -0 are interchangeable nowhere in this contract: require identical bits
    else:
        same = a == b
    if not np.all(same):
        bad = np.argwhere(~same)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {bad.shape[0]} of {a.size} elements differ; first at {i}: {a[i]!r} vs {b[i]!r}")
