"""The drop-in for the reference's compiled Python modules (runko_b200/dropin: `runko_cpp_bindings`, `pycorgi`,
an `mpi4py` stand-in; compiled half: runko_b200/csrc/pybind/b200_bindings.cpp).

CPU (no GPU needed): the module / class / method surface SURVEY.md §8b lists; the UNMODIFIED reference package
(/root/reference/runko, where present) imports against it, builds a TileGrid, and a tile constructor goes all the way
through pybind11 into the C-ABI (and, without a device, fails loudly — no fallback).
GPU: projects/pic-turbulence/pic.py's lap function driven through the drop-in (by runko.Simulation itself where
the reference is present, else by the restated dispatch of tests/dropin_lap_driver.py) against the CPU oracle."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DROPIN = os.path.join(ROOT, "runko_b200", "dropin")
REF = "/root/reference"


def _run(code, with_ref):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([DROPIN, ROOT] + ([REF] if with_ref else []))
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    return subprocess.run([sys.executable, "-c", code], cwd="/tmp", env=env, capture_output=True, text=True, timeout=600)


def test_dropin_module_surface():
    """exact names the reference's Python layer imports and calls"""
    code = r'''
import runko_cpp_bindings as rcb, pycorgi.threeD as pc, mpi4py.MPI as MPI
from runko_cpp_bindings.tools import _virtual_tile_sync_handshake_mode, comm_mode, _get_gpu_mem_kB
from runko_cpp_bindings.emf.threeD import MpiioFieldsWriter, MpiioParticlesWriter, MpiioSpectraWriter, _write_average_B_energy_density, _write_average_E_energy_density
from runko_cpp_bindings.pic.threeD import _write_average_kinetic_energy
assert {m.name: m.value for m in comm_mode} == dict(emf_J=0, emf_E=1, emf_B=2, pic_particle=3, pic_particle_extra=4, emf_J_exchange=6)
assert _virtual_tile_sync_handshake_mode(comm_mode.pic_particle) and _virtual_tile_sync_handshake_mode(comm_mode.emf_E) is None
emf_methods = "set_EBJ batch_set_EBJ get_EBJ get_EBJ_with_halo push_half_b push_e add_current filter_current global_coordinate_map register_antenna deposit_antenna_current register_edge_bc apply_edge_bcs apply_edge_bc canonical_type virtual_tile_specialization load_metainfo nhood".split()
pic_methods = emf_methods + "get_positions get_velocities get_ids inject_to_each_cell inject batch_inject_to_cells batch_inject_in_x_stripe push_particles pack_outgoing_particles deposit_current sort_particles register_reflector_wall reflect_particles advance_reflector_walls".split()
for m in emf_methods: assert callable(getattr(rcb.emf.threeD.Tile, m)), m
for m in pic_methods: assert callable(getattr(rcb.pic.threeD.Tile, m)), m
for n in "VirtualTile antenna_mode edge_bc".split(): assert hasattr(rcb.emf.threeD, n), n
for n in "VirtualTile ParticleState ParticleStateBatch reflector_wall".split(): assert hasattr(rcb.pic.threeD, n), n
assert "pycorgi" not in rcb.pic.threeD.Tile.__module__ and "pycorgi" in pc.Tile.__module__          # runko/tile_grid.py:145
assert issubclass(rcb.pic.threeD.Tile, rcb.emf.threeD.Tile) and rcb.pic.threeD.Tile.canonical_type() is rcb.pic.threeD.Tile
g = pc.Grid(4, 2, 2)
for m in "rank size master add_tile get_tile get_tile_ids get_local_tiles get_virtual_tiles get_boundary_tiles is_local analyze_boundaries send_tiles recv_tiles send_data recv_data wait_data local_communication bcast_mpi_grid get_mpi_grid set_mpi_grid get_Nx get_Ny get_Nz set_grid_lims id".split():
    assert callable(getattr(g, m)), m
assert (g.get_Nx(), g.get_Ny(), g.get_Nz(), g.id(1, 1, 1), g.rank(), g.size()) == (4, 2, 2, 1 + 4 * (1 + 2), 0, 1)
g.set_mpi_grid(3, 1, 1, 1); assert g.get_mpi_grid(3, 1, 1) == 1
w = MPI.COMM_WORLD
assert (w.Get_rank(), w.Get_size(), w.size) == (0, 1, 1) and w.gather(5) == [5] and w.bcast(7) == 7; w.barrier(); w.Barrier()
import _b200pic
assert _b200pic.sizeof_config == len(bytes(__import__("runko_b200")._abi.B2PConfig()))
print("surface ok")
'''
    r = _run(code, with_ref=False)
    assert r.returncode == 0 and "surface ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "runko")), reason="/root/reference is not present on this box")
def test_unmodified_reference_package_imports_and_reaches_the_c_abi():
    code = r'''
import runko, inspect
assert runko.__file__.startswith("/root/reference/"), runko.__file__
assert runko.pic.threeD.Tile.__module__ == "runko_cpp_bindings.pic.threeD"
conf = runko.Configuration(None)
conf.n_tiles, conf.n_cells_per_tile = [2, 2, 2], [8, 8, 8]
conf.tile_partitioning = "hilbert_curve"
tg = runko.TileGrid(conf)                                     # pycorgi.Grid + runko/balance_grid.py (Hilbert) + bcast_mpi_grid
assert sorted(tg.local_tile_indices()) == [(i, j, k) for i in range(2) for j in range(2) for k in range(2)]
conf.cfl, conf.field_propagator, conf.q0, conf.m0 = "foo", "fdtd2", -1.0, 1.0
try:
    runko.emf.threeD.Tile((0, 0, 0), conf)
    raise SystemExit("bad config accepted")
except RuntimeError as e:                                      # tests/py/test_emf.py:14-31
    import re; assert re.search("cfl.*unsupported type.*foo", str(e)), str(e)
conf.cfl = 0.45
conf.particle_pusher, conf.field_interpolator, conf.current_depositer = "boris", "linear_1st", "zigzag_1st_atomic"
import _b200pic
try:
    t = runko.pic.threeD.Tile((0, 0, 0), conf)
    print("tile constructed on a GPU box:", t.mins, t.maxs)
except RuntimeError as e:
    assert "no usable CUDA device" in str(e) and "no CPU fallback" in str(e), str(e)
    print("no device: constructor failed loudly inside libb200pic.so")
print("reference package ok")
'''
    r = _run(code, with_ref=True)
    assert r.returncode == 0 and "reference package ok" in r.stdout, r.stdout + r.stderr


# The reference's OWN host-side unit tests (unmodified, where they lie) against the drop-in modules: everything below the
# Python package that needs no device — configuration, module layout, TileGrid construction over pycorgi.Grid, virtual tile
# specialisations, the Hilbert tile ordering.  (The tests that compute — test_emf*.py, test_pic*.py — need a GPU, which the
# container holding /root/reference does not have; tests/test_reference_suite.py runs them on the oracle and
# tests/test_kats.py restates their known answers for the CUDA path.)
REF_HOST_TESTS = {
    "test_configuration.py": "4 passed",
    "test_module_meta.py": "1 passed",
    "test_auto_outdir.py": "13 passed",
    "test_auto_tile_grid.py": "1 passed",
    "test_tile_grid.py": "3 passed",
    "test_virtual_tile_specializations.py": "2 passed",
    "test_hilbert.py": "7 passed",
}


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests", "py")), reason="/root/reference is not present on this box")
@pytest.mark.parametrize("fname", sorted(REF_HOST_TESTS))
def test_reference_host_side_unit_tests_pass_on_the_dropin(fname):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([DROPIN, ROOT, REF])
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", os.path.join(REF, "tests", "py", fname)],
                       cwd="/tmp", env=env, capture_output=True, text=True, timeout=600)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert REF_HOST_TESTS[fname] in r.stdout, tail


@pytest.mark.gpu
@pytest.mark.parametrize("use_reference", [True, False])
def test_pic_turbulence_lap_through_the_dropin_matches_the_oracle(use_reference, tmp_path):
    """BASELINE configs[0] physics at 64^3, one tile, 2 x 2 ppc, 10 laps: the per-lap diagnostics the lap function
    writes (io_average_*) agree with the CPU oracle's to 1e-4, particle ids are conserved."""
    if use_reference and not os.path.isdir(os.path.join(REF, "runko")):
        pytest.skip("/root/reference is not present on this box (the restated dispatch runs instead)")
    sys.path.insert(0, ROOT)
    from oracle.oracle import OracleGrid
    from util import Conf
    n, ppc, nlaps = 64, 2, 10
    out = str(tmp_path / "run")
    env = dict(os.environ)
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    if not use_reference:
        env["DROPIN_DRIVER_NO_REFERENCE"] = "1"
    r = subprocess.run([sys.executable, os.path.join(HERE, "dropin_lap_driver.py"), out, str(n), str(ppc), str(nlaps)],
                       cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["tile_module"] == "runko_cpp_bindings.pic.threeD" and info["kernels_launched"] > 0
    assert ("unmodified" in info["driver"]) == use_reference
    st = np.load(os.path.join(out, "initial_state.npz"))
    cfl, oppc = 0.45, 2 * ppc
    q0 = -(cfl ** 2) / (0.5 * oppc * 2.0)
    conf = Conf(n_tiles=[1, 1, 1], n_cells_per_tile=[n, n, n], cfl=cfl, field_propagator="fdtd2", m0=1, m1=1, q0=q0, q1=abs(q0),
                particle_pusher="boris", field_interpolator="linear_1st", current_depositer="zigzag_1st_atomic", current_filter="binomial2")
    org = OracleGrid(conf)
    org.set_fields(0, st["E"], st["B"], st["J"], with_halo=True)
    for sp in range(2):
        org.set_particles(0, sp, *(st[f"{c}{sp}"] for c in ("x", "y", "z", "ux", "uy", "uz", "id")))
    for m in (1, 2):
        org.local_communication(m)
    ref = {"kin": [], "B": [], "E": []}
    for lap in range(nlaps):
        org.step_pic(lap, threads=os.cpu_count() or 1)
        b, e, k, s = org.energies()
        ref["kin"].append(k / s)
        # field energy densities from the oracle's lattices, summed in fp64: the reference's own diagnostic accumulates
        # 262144 nearly equal fp32 terms serially (emf/yee_lattice.c++:383-428), which alone is 3e-4 off the exact sum
        Eo, Bo, _ = org.get_fields(0, with_halo=False)
        ref["B"].append(0.5 * float(np.sum(Bo.astype(np.float64) ** 2)) / n ** 3)
        ref["E"].append(0.5 * float(np.sum(Eo.astype(np.float64) ** 2)) / n ** 3)
        assert abs(b / n ** 3 - ref["B"][-1]) <= 1e-3 * ref["B"][-1]

    def rows(name):
        return np.array([[float(v) for v in line.split()] for line in open(os.path.join(out, name)).read().strip().splitlines()])
    kin, eb, ee = rows("average_kinetic_energy.txt"), rows("average_B_energy_density.txt"), rows("average_E_energy_density.txt")
    assert kin.shape == (nlaps, 3) and eb.shape == (nlaps, 2) and list(kin[:, 0]) == list(range(nlaps))
    # the text files carry 6 significant digits (operator<< of the reference): compare at that precision
    assert np.max(np.abs(kin[:, 1:] - np.array(ref["kin"])) / np.array(ref["kin"])) <= 1e-4
    assert np.max(np.abs(eb[:, 1] - np.array(ref["B"])) / np.array(ref["B"])) <= 1e-4
    assert np.max(np.abs(ee[:, 1] - np.array(ref["E"])) / np.maximum(np.array(ref["E"]), 1e-30)) <= 1e-3
