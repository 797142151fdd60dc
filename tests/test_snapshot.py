"""Field snapshot writer: "<prefix>/flds_<lap>.bin", the reference's RNKO v3 format
(src/runko/io/snapshots/mpiio_header.h:56-82, mpiio_fields.c++:221-400).

The reference's writer itself needs MPI-IO + corgi and cannot be compiled in this image, so the format is
pinned (i) by the reference's own READER, runko/mpiio_reader.py, imported from /root/reference where that
exists, (ii) by the known-answer cases of tests/py/test_mpiio_fields.py restated here (constant fields,
stride, multi-tile placement, density per coarse cell, species clamp, header fields), on both backends, and
(iii) by byte-for-byte equality of the CUDA path's file with the oracle's on a random multi-tile PIC grid."""
import os
import struct
import sys

import numpy as np
import pytest

from backends import GRID
from util import pic_conf, random_lattice, random_particles

NAMES = ["ex", "ey", "ez", "bx", "by", "bz", "jx", "jy", "jz"]


def read_snapshot(path):
    """own minimal reader of the layout in mpiio_header.h:56-82"""
    raw = open(path, "rb").read()
    (magic, ver, hsize, nf, nx, ny, nz, stride, Nx, Ny, Nz, mx, my, mz, lap, dsize) = struct.unpack_from("<IIIIiiiiiiiiiiiI", raw, 0)
    names = [raw[64 + 16 * f:80 + 16 * f].split(b"\0")[0].decode() for f in range(nf)]
    hdr = dict(magic=magic, version=ver, header_size=hsize, num_fields=nf, nx=nx, ny=ny, nz=nz, stride=stride, Nx=Nx, Ny=Ny, Nz=Nz,
               NxMesh=mx, NyMesh=my, NzMesh=mz, lap=lap, dtype_size=dsize, field_names=names)
    assert len(raw) == hsize + nf * nx * ny * nz * 4
    data = np.frombuffer(raw, np.float32, offset=hsize).reshape(nf, nz, ny, nx)
    return hdr, {n: data[i] for i, n in enumerate(names)}


def const_lattice(n, v):
    a = np.empty((3,) + tuple(x + 6 for x in n), np.float32)
    for c in range(3):
        a[c] = v[c]
    return a


def writer_of(backend, g):
    return (g.g if backend == "oracle" else g.grid).write_fields_snapshot


def test_constant_fields_header_and_values(backend, tmp_path):
    """test_mpiio_fields.py:42-122,331-392: header fields, names, shapes (nz, ny, nx), constant values, lap."""
    n = (10, 11, 13)
    g = GRID[backend](pic_conf(n_tiles=(1, 1, 1), n_cells=n))
    g.set_fields((0, 0, 0), const_lattice(n, (1, 2, 3)), const_lattice(n, (4, 5, 6)), const_lattice(n, (7, 8, 9)))
    writer_of(backend, g)(tmp_path, 42, 1, 2)
    hdr, f = read_snapshot(tmp_path / "flds_42.bin")
    assert hdr["magic"] == 0x524E4B4F and hdr["version"] == 3 and hdr["header_size"] == 512 and hdr["dtype_size"] == 4
    assert (hdr["nx"], hdr["ny"], hdr["nz"], hdr["stride"], hdr["lap"], hdr["num_fields"]) == (10, 11, 13, 1, 42, 11)
    assert (hdr["Nx"], hdr["Ny"], hdr["Nz"], hdr["NxMesh"], hdr["NyMesh"], hdr["NzMesh"]) == (1, 1, 1, 10, 11, 13)
    assert hdr["field_names"] == NAMES + ["n0", "n1"]
    for i, nm in enumerate(NAMES):
        assert f[nm].shape == (13, 11, 10) and np.all(f[nm] == i + 1)
    assert np.all(f["n0"] == 0) and np.all(f["n1"] == 0)


@pytest.mark.parametrize("stride", [2, 4, 8])
def test_stride_samples_EB_and_sums_J(backend, tmp_path, stride):
    """test_mpiio_fields.py:124-189: E,B point samples at (i,j,k)*stride; J summed over stride^3 cells."""
    n = (8, 8, 8)
    rng = np.random.default_rng(3)
    g = GRID[backend](pic_conf(n_tiles=(1, 1, 1), n_cells=n))
    E, B, J = (random_lattice(rng, n) for _ in range(3))
    g.set_fields((0, 0, 0), E, B, J)
    writer_of(backend, g)(tmp_path, 0, stride, 0)
    hdr, f = read_snapshot(tmp_path / "flds_0.bin")
    m = 8 // stride
    assert (hdr["nx"], hdr["ny"], hdr["nz"], hdr["num_fields"]) == (m, m, m, 9)
    inner = (slice(3, -3),) * 3
    for c, nm in enumerate(("ex", "ey", "ez")):
        assert np.array_equal(f[nm], E[c][inner][::stride, ::stride, ::stride].transpose(2, 1, 0))
    for c, nm in enumerate(("bx", "by", "bz")):
        assert np.array_equal(f[nm], B[c][inner][::stride, ::stride, ::stride].transpose(2, 1, 0))
    for c, nm in enumerate(("jx", "jy", "jz")):
        ref = J[c][inner].astype(np.float64).reshape(m, stride, m, stride, m, stride).sum(axis=(1, 3, 5)).transpose(2, 1, 0)
        assert np.allclose(f[nm], ref, rtol=1e-5, atol=1e-5)


def test_multi_tile_placement_and_density(backend, tmp_path):
    """test_mpiio_fields.py:191-238,394-435,477-508: tile (ti,tj,tk) lands at its global offset; density = alive
    particles per coarse cell; species beyond the requested count are not written; request is clamped to 5."""
    n, T = (4, 6, 5), (3, 2, 2)
    g = GRID[backend](pic_conf(n_tiles=T, n_cells=n))
    rng = np.random.default_rng(9)
    expect = np.zeros((2, T[2] * n[2], T[1] * n[1], T[0] * n[0]))
    for idx in g.tiles():
        i, j, k = idx
        g.set_fields(idx, const_lattice(n, (100 * i + 10 * j + k,) * 3), const_lattice(n, (0, 0, 0)), const_lattice(n, (0, 0, 0)))
        mins = [i * n[0], j * n[1], k * n[2]]
        maxs = [mins[d] + n[d] for d in range(3)]
        for sp in range(2):
            pos, vel, _ = random_particles(rng, 300 + 50 * sp, mins, maxs)
            g.inject(idx, sp, pos.astype(np.float64), vel.astype(np.float64))
            ci = np.floor(pos).astype(int)
            np.add.at(expect[sp], (ci[2], ci[1], ci[0]), 1)
    writer_of(backend, g)(tmp_path, 7, 1, 9)               # 9 species requested -> clamped to 5 slots
    hdr, f = read_snapshot(tmp_path / "flds_7.bin")
    assert hdr["num_fields"] == 14 and hdr["field_names"][9:] == ["n0", "n1", "n2", "n3", "n4"]
    for idx in g.tiles():
        i, j, k = idx
        blk = f["ex"][k * n[2]:(k + 1) * n[2], j * n[1]:(j + 1) * n[1], i * n[0]:(i + 1) * n[0]]
        assert np.all(blk == 100 * i + 10 * j + k)
    assert np.array_equal(f["n0"], expect[0]) and np.array_equal(f["n1"], expect[1])
    assert not np.any(f["n2"]) and not np.any(f["n4"])
    writer_of(backend, g)(tmp_path, 8, 2, 1)               # stride 2, one species
    hdr, f = read_snapshot(tmp_path / "flds_8.bin")
    assert hdr["num_fields"] == 10 and (hdr["nx"], hdr["ny"], hdr["nz"]) == (6, 6, 4)
    # stride 2 over a 5-cell z mesh keeps floor(5/2) = 2 coarse cells per tile; the odd last layer is not sampled
    tot = sum(1 for _ in g.tiles())
    assert f["n0"].sum() <= expect[0].sum() and f["n0"].sum() > 0.7 * expect[0].sum() and tot == 12


@pytest.mark.skipif(not os.path.isdir("/root/reference/runko"), reason="/root/reference is not present on this box")
def test_reference_reader_reads_it(tmp_path):
    """The reference's own reader (runko/mpiio_reader.py, loaded from /root/reference without importing the
    runko package) parses a snapshot written by the oracle: header dict and field arrays as its tests expect."""
    import importlib.util
    import types
    pkg = types.ModuleType("runko")
    pkg.__path__ = ["/root/reference/runko"]
    saved = sys.modules.get("runko")
    sys.modules["runko"] = pkg
    try:
        for name in ("mpiio_constants", "mpiio_reader"):
            spec = importlib.util.spec_from_file_location("runko." + name, f"/root/reference/runko/{name}.py")
            mod = importlib.util.module_from_spec(spec)
            sys.modules["runko." + name] = mod
            spec.loader.exec_module(mod)
        reader = sys.modules["runko.mpiio_reader"]
        n, T = (6, 4, 8), (2, 1, 2)
        g = GRID["oracle"](pic_conf(n_tiles=T, n_cells=n))
        rng = np.random.default_rng(5)
        lat = {}
        for idx in g.tiles():
            lat[idx] = [random_lattice(rng, n) for _ in range(3)]
            g.set_fields(idx, *lat[idx])
        g.g.write_fields_snapshot(tmp_path, 3, 2, 2)
        path = tmp_path / "flds_3.bin"
        hdr = reader.read_header(path)
        assert hdr["magic"] == reader.MAGIC and hdr["nx"] == 6 and hdr["ny"] == 2 and hdr["nz"] == 8 and hdr["lap"] == 3
        assert hdr["field_names"][:9] == NAMES
        fields = reader.read_field_snapshot(path)
        own = read_snapshot(path)[1]
        for nm in hdr["field_names"]:
            assert np.array_equal(np.asarray(fields[nm]), own[nm]), nm
        i, j, k = 1, 0, 1
        blk = np.asarray(fields["by"])[k * 4:(k + 1) * 4, 0:2, i * 3:(i + 1) * 3]
        assert np.array_equal(blk, lat[(i, j, k)][1][1][3:-3, 3:-3, 3:-3][::2, ::2, ::2].transpose(2, 1, 0))
    finally:
        for name in ("runko.mpiio_constants", "runko.mpiio_reader"):
            sys.modules.pop(name, None)
        if saved is not None:
            sys.modules["runko"] = saved
        else:
            sys.modules.pop("runko", None)


@pytest.mark.gpu
@pytest.mark.parametrize("stride", [1, 2, 4])
def test_cuda_snapshot_is_byte_identical_to_the_oracle(tmp_path, stride):
    n, T = (8, 12, 4), (2, 2, 3)
    conf = pic_conf(n_tiles=T, n_cells=n)
    go, gb = GRID["oracle"](conf), GRID["b200"](conf)
    rng = np.random.default_rng(17)
    for idx in go.tiles():
        E, B, J = (random_lattice(rng, n) for _ in range(3))
        mins = [idx[d] * n[d] for d in range(3)]
        maxs = [mins[d] + n[d] for d in range(3)]
        for g in (go, gb):
            g.set_fields(idx, E, B, J)
        for sp in range(2):
            pos, vel, _ = random_particles(rng, 700, mins, maxs)
            for g in (go, gb):
                g.inject(idx, sp, pos.astype(np.float64), vel.astype(np.float64))
    (tmp_path / "o").mkdir()
    (tmp_path / "b").mkdir()
    go.g.write_fields_snapshot(tmp_path / "o", 5, stride, 2)
    gb.grid.write_fields_snapshot(tmp_path / "b", 5, stride, 2)
    assert open(tmp_path / "o" / "flds_5.bin", "rb").read() == open(tmp_path / "b" / "flds_5.bin", "rb").read()
