#!/usr/bin/env python
"""Regenerates tests/golden/*.npz: seeded inputs and the outputs of the CPU oracle for
every kernel of the PIC-step path.

Provenance: the reference ships no numeric golden vectors for this path and
is C++ (cannot be imported); these vectors were produced by oracle/pic_oracle.cpp at a
revision that (i) passes the reference's own unit tests run unmodified through
tests/refshim (tests/test_reference_suite.py) and (ii) agrees bit-for-bit with the
reference's kernel sources compiled from /root/reference where that build is available
(oracle/_ref, tests/test_oracle_vs_reference_build.py).  They freeze that behaviour so the
GPU box (which has no /root/reference) checks the CUDA path against fixed files.

    python tests/golden/make_golden.py        # rewrites the .npz files next to this script
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.oracle import OracleGrid  # noqa: E402
from util import emf_conf, pic_conf, random_lattice, random_particles  # noqa: E402

N = (6, 7, 9)


def stencil_kw(rng):
    kw = {}
    for ax in "xyz":
        for name in ("delta", "gamma", "beta_p1", "beta_p2", "beta2_p1", "beta2_p2", "beta3_p1", "beta3_p2",
                     "zeta_p1", "zeta_p2", "zeta2_p1", "zeta2_p2", "zeta3_p1", "zeta3_p2"):
            kw[f"stencil_{ax}_{name}"] = float(np.float32(0.05 * rng.standard_normal()))
    return kw


def fields_case(name, conf_kw, ops, seed):
    rng = np.random.default_rng(seed)
    kw = dict(conf_kw)
    if kw.get("field_propagator") == "stencil":
        kw.update(stencil_kw(rng))
    conf = emf_conf(n_cells=N, **kw)
    g = OracleGrid(conf)
    E, B, J = (random_lattice(rng, N) for _ in range(3))
    g.set_fields(0, E, B, J, with_halo=True)
    for op in ops:
        g.tile_op(0, op)
    oE, oB, oJ = g.get_fields(0, with_halo=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), conf=np.array(repr(kw)), ops=np.array(ops), E=E, B=B, J=J,
                        oE=oE, oB=oB, oJ=oJ)


def particle_case(name, conf_kw, ops, seed, n=600, n_tiles=(1, 1, 1), idx=(0, 0, 0), margin=0.0, dead_frac=0.05):
    rng = np.random.default_rng(seed)
    conf = pic_conf(n_tiles=n_tiles, n_cells=N, q0=-0.7, q1=0.4, m1=3.0, **conf_kw)
    g = OracleGrid(conf)
    t = g.cid(*idx)
    E, B, J = (random_lattice(rng, N) for _ in range(3))
    g.set_fields(t, E, B, J, with_halo=True)
    mins = np.array(idx) * np.array(N)
    out = dict(conf=np.array(repr(conf_kw)), ops=np.array(ops), n_tiles=np.array(n_tiles), idx=np.array(idx), E=E, B=B, J=J)
    for sp in range(2):
        pos, vel, ids = random_particles(rng, n + 13 * sp, mins, mins + np.array(N), dead_frac=dead_frac, tag=sp + 1,
                                         margin=margin, u_scale=0.8)
        g.set_particles(t, sp, *pos, *vel, ids)
        out[f"in{sp}"] = np.concatenate([pos, vel]).astype(np.float32)
        out[f"in{sp}_id"] = ids
    for op in ops:
        g.tile_op(t, op)
    for sp in range(2):
        p = g.get_particles(t, sp, alive_only=False)
        out[f"out{sp}"] = np.stack(p[:6]).astype(np.float32)
        out[f"out{sp}_id"] = p[6]
        out[f"keys{sp}"] = g.sort_keys(t, sp)
    out["oJ"] = g.get_fields(t, with_halo=True)[2]
    if "pack_outgoing_particles" in ops:
        buf, ends = g.get_outgoing(t)
        out["out_pos"], out["out_vel"], out["out_id"], out["out_ends"] = buf["pos"], buf["vel"], buf["id"], ends
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def grid_case(name, seed, n_tiles=(2, 1, 2), laps=2):
    """halo fill, J exchange and whole laps on a small multi-tile periodic grid"""
    rng = np.random.default_rng(seed)
    n = (4, 5, 6)
    conf = pic_conf(n_tiles=n_tiles, n_cells=n, q0=-0.05, q1=0.05, current_filter="binomial2")
    g = OracleGrid(conf)
    out = dict(n_tiles=np.array(n_tiles), n_cells=np.array(n), laps=np.array(laps))
    for t in range(g.num_tiles):
        i, j, k = t % n_tiles[0], (t // n_tiles[0]) % n_tiles[1], t // (n_tiles[0] * n_tiles[1])
        E, B, J = (random_lattice(rng, n, 0.3) for _ in range(3))
        g.set_fields(t, E, B, J, with_halo=True)
        out[f"t{t}_E"], out[f"t{t}_B"], out[f"t{t}_J"] = E, B, J
        mins = np.array([i, j, k]) * np.array(n)
        for sp in range(2):
            pos, vel, _ = random_particles(rng, 3 * int(np.prod(n)), mins, mins + np.array(n), u_scale=1.5)
            g.inject(t, sp, *pos.astype(np.float64), *vel.astype(np.float64))
            out[f"t{t}_p{sp}"] = np.concatenate([pos, vel]).astype(np.float32)
    for mode in (1, 2, 6, 0):                       # emf_E, emf_B, emf_J_exchange, emf_J
        g.local_communication(mode)
    for t in range(g.num_tiles):
        f = g.get_fields(t, with_halo=True)
        out[f"t{t}_commE"], out[f"t{t}_commB"], out[f"t{t}_commJ"] = f
    for lap in range(laps):
        g.step_pic(lap)
    for t in range(g.num_tiles):
        f = g.get_fields(t, with_halo=True)
        out[f"t{t}_lapE"], out[f"t{t}_lapB"], out[f"t{t}_lapJ"] = f
        for sp in range(2):
            p = g.get_particles(t, sp, alive_only=False)
            out[f"t{t}_lap_p{sp}"] = np.stack(p[:6]).astype(np.float32)
            out[f"t{t}_lap_id{sp}"] = p[6]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def main():
    fields_case("fdtd2", dict(field_propagator="fdtd2"), ["push_half_b", "push_e", "push_half_b", "add_current", "push_e"], 101)
    fields_case("stencil", dict(field_propagator="stencil"), ["push_half_b", "push_e", "push_half_b"], 102)
    fields_case("filter_binomial2", dict(current_filter="binomial2"), ["filter_current"] * 3, 103)
    fields_case("filter_binomial2_unrolled", dict(current_filter="binomial2_unrolled"), ["filter_current"] * 3, 104)
    for p in ("boris", "higuera_cary", "faraday"):
        particle_case("push_" + p, dict(particle_pusher=p), ["push_particles", "push_particles"], 110)
    particle_case("deposit_atomic", dict(current_depositer="zigzag_1st_atomic"), ["deposit_current"], 120)
    particle_case("deposit_sorted", dict(current_depositer="zigzag_1st"), ["deposit_current"], 121)
    particle_case("sort", {}, ["sort_particles"], 130, n=900, dead_frac=0.15)
    particle_case("pack_outgoing", {}, ["pack_outgoing_particles"], 140, n=900, n_tiles=(3, 3, 3), idx=(1, 1, 1), margin=0.8)
    grid_case("grid_laps", 150)
    tot = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print("golden fixtures written:", tot, "bytes")


if __name__ == "__main__":
    main()
