#!/usr/bin/env python
"""Writes tests/golden/shock_reflector.npz, shock_edge_bc.npz, antenna.npz and snapshot.bin.

Provenance: shock_reflector / shock_edge_bc are outputs of THE REFERENCE'S OWN code — pic/reflector_wall.c++
(ParticleContainer::reflect_at_wall, Tile<3>::reflect_particles / advance_reflector_walls) and
YeeLattice::apply_edge_bc, compiled from /root/reference into oracle/_ref/libref_kernels.so — on seeded inputs
(this script needs that build; the oracle reproduces the same bits, tests/test_oracle_vs_reference_build.py).
antenna.npz and snapshot.bin come from the oracle, whose antenna passes the reference's 15 antenna unit tests and
whose snapshot file is parsed by the reference's own reader (emf/tile.c++ and the MPI-IO writer cannot be compiled here).

    python tests/golden/make_golden_shock.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import reference_build as rbuild  # noqa: E402
from oracle.oracle import OracleGrid  # noqa: E402
from runko_b200 import antenna_mode, edge_bc, reflector_wall  # noqa: E402
from util import DEAD, emf_conf, pic_conf, random_lattice, random_particles  # noqa: E402

N = (9, 5, 6)


def reflector():
    rng = np.random.default_rng(301)
    conf = pic_conf(n_tiles=(2, 1, 1), n_cells=N, q0=-0.7, q1=0.4, cfl=0.45)
    ref = rbuild.RefTile(conf, (0, 0, 0))
    E, B, J = (random_lattice(rng, N) for _ in range(3))
    ref.set_fields(E, B, J)
    out = dict(E=E, B=B, J=J, wall=np.array([4.0, 0.25, 1.0 / np.sqrt(1 - 0.25 ** 2)], np.float32), laps=3)
    for sp in range(2):
        m = 1500 + 7 * sp
        pos = np.stack([4.0 + 5.0 * rng.random(m), 5.0 * rng.random(m), 6.0 * rng.random(m)]).astype(np.float32)
        pos[0, :30] = (2.0 + 1.8 * rng.random(30)).astype(np.float32)
        vel = (0.6 * rng.standard_normal((3, m))).astype(np.float32)
        vel[0] -= np.float32(0.8)
        ids = (np.uint64(sp + 1) << np.uint64(40)) | np.arange(m, dtype=np.uint64)
        ids[rng.random(m) < 0.05] = DEAD
        ref.set_particles(sp, *pos, *vel, ids)
        out[f"in{sp}"], out[f"in{sp}_id"] = np.concatenate([pos, vel]), ids
    from runko_b200._abi import ReflectorWall
    ref.register_reflector_wall(ReflectorWall(*[float(v) for v in out["wall"]]))
    for lap in range(out["laps"]):
        for op in ("push_particles", "reflect_particles", "deposit_current", "advance_reflector_walls"):
            ref.op(op)
    for sp in range(2):
        p = ref.get_particles(sp)
        out[f"out{sp}"], out[f"out{sp}_id"] = np.stack(p[:6]), p[6]
    out["oJ"] = ref.get_fields()[2]
    out["walloc"] = np.float32(ref.reflector_walls()[0][0])
    np.savez_compressed(os.path.join(HERE, "shock_reflector.npz"), **out)


def edge():
    rng = np.random.default_rng(302)
    conf = pic_conf(n_tiles=(2, 2, 2), n_cells=N)
    ref = rbuild.RefTile(conf, (0, 1, 1))
    E, B, J = (random_lattice(rng, N) for _ in range(3))
    ref.set_fields(E, B, J)
    bcs = [dict(direction=0, side=0, position=4.0, E_components=0b110, B_components=0, J_components=0b111),
           dict(direction=0, side=1, position=7.0, Ex=0.5, Ey=0.1, Ez=-0.2, Bx=0.3, By=0.04, Bz=0.02, J_components=0b111),
           dict(direction=2, side=1, position=9.5, Jx=1.0, Jy=2.0, Jz=3.0, E_components=0b001, B_components=0b100)]
    for kw in bcs:
        ref.register_edge_bc(edge_bc(**kw))
    for mode in (2, 0, 1):
        ref.apply_edge_bcs(mode)
    oE, oB, oJ = ref.get_fields()
    np.savez_compressed(os.path.join(HERE, "shock_edge_bc.npz"), E=E, B=B, J=J, oE=oE, oB=oB, oJ=oJ, bcs=np.array(repr(bcs)))


def antenna():
    rng = np.random.default_rng(303)
    conf = emf_conf(n_tiles=(2, 2, 1), n_cells=N, cfl=0.45)
    g = OracleGrid(conf)
    t = g.cid(1, 0, 0)
    J = random_lattice(rng, N, 0.01)
    g.set_fields(t, J=J, with_halo=True)
    modes = [dict(A=[1.0, -0.5, 0.25], k=[0.11, 0.0, 0.23]), dict(A=[0.0, 2.0, 1.0], n=[1, 2, 0], lap_coeffs=[0.5 + 0.5j, -0.25j])]
    for m in modes:
        g.register_antenna(t, antenna_mode(**m))
    g.tile_op(t, "deposit_antenna_current")
    g.tile_op(t, "deposit_antenna_current")
    np.savez_compressed(os.path.join(HERE, "antenna.npz"), J=J, oJ=g.get_fields(t, with_halo=True)[2],
                        modes=np.array(repr([{k: (str(v) if k == "lap_coeffs" else v) for k, v in m.items()} for m in modes])))


def snapshot():
    rng = np.random.default_rng(304)
    T, n = (2, 1, 2), (4, 6, 4)
    conf = pic_conf(n_tiles=T, n_cells=n)
    g = OracleGrid(conf)
    out = {}
    for t in range(g.num_tiles):
        idx = (t % T[0], (t // T[0]) % T[1], t // (T[0] * T[1]))
        E, B, J = (random_lattice(rng, n) for _ in range(3))
        g.set_fields(t, E, B, J, with_halo=True)
        mins = [idx[d] * n[d] for d in range(3)]
        out[f"t{t}_E"], out[f"t{t}_B"], out[f"t{t}_J"] = E, B, J
        for sp in range(2):
            pos, vel, _ = random_particles(rng, 150, mins, [mins[d] + n[d] for d in range(3)])
            g.inject(t, sp, *pos.astype(np.float64), *vel.astype(np.float64))
            out[f"t{t}_p{sp}"] = np.concatenate([pos, vel])
    g.write_fields_snapshot(HERE, 9, 2, 2)
    os.replace(os.path.join(HERE, "flds_9.bin"), os.path.join(HERE, "snapshot_flds_9.bin"))
    np.savez_compressed(os.path.join(HERE, "snapshot_inputs.npz"), **out)


if __name__ == "__main__":
    if not rbuild.available() and not rbuild.build():
        raise SystemExit("needs oracle/_ref/libref_kernels.so (build with /root/reference present)")
    reflector(); edge(); antenna(); snapshot()
    print("written:", sorted(f for f in os.listdir(HERE) if f.startswith(("shock_", "antenna", "snapshot"))))
