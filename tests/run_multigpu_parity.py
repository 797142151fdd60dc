"""Multi-GPU parity driver (launched by torchrun, one rank per GPU):
every rank owns a block of tiles, exchanges halos / currents / particles over NCCL and
is compared with the single-process CPU oracle of the whole periodic grid.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_multigpu_parity.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import runko_b200 as rb  # noqa: E402
from oracle.oracle import OracleGrid  # noqa: E402
from runko_b200._lib import check  # noqa: E402
from util import DEAD, assert_bits_equal, pic_conf, random_lattice, random_particles  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    L = rb.lib()
    check(L.b2p_init(local))
    blocks = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    tpb = (2, 2, 1) if world <= 2 else (2, 1, 1)
    n_tiles = tuple(blocks[d] * tpb[d] for d in range(3))
    n_cells = (6, 7, 8)
    conf = pic_conf(n_tiles=n_tiles, n_cells=n_cells, q0=-0.05, q1=0.05, current_filter="binomial2")
    rng = np.random.default_rng(77)          # same stream on every rank
    org = OracleGrid(conf)
    grid = rb.Grid(conf)
    owner = np.zeros(int(np.prod(n_tiles)), np.int32)
    tiles = {}
    for i in range(n_tiles[0]):
        for j in range(n_tiles[1]):
            for k in range(n_tiles[2]):
                cid = org.cid(i, j, k)
                r = (i // tpb[0]) + blocks[0] * ((j // tpb[1]) + blocks[1] * (k // tpb[2]))
                owner[cid] = r
                E, B, J = (random_lattice(rng, n_cells, 0.3) for _ in range(3))
                mins = np.array([i * n_cells[0], j * n_cells[1], k * n_cells[2]], float)
                parts = [random_particles(rng, 4 * int(np.prod(n_cells)), mins, mins + np.array(n_cells), u_scale=1.5) for _ in range(2)]
                org.set_fields(cid, E, B, J, with_halo=True)
                for sp in range(2):
                    org.inject(cid, sp, *parts[sp][0].astype(np.float64), *parts[sp][1].astype(np.float64))
                if r == rank:
                    t = rb.PicTile((i, j, k), conf)
                    t.set_fields_f32(E, B, J, with_halo=True)
                    for sp in range(2):
                        t._inject_arrays(sp, parts[sp][0].astype(np.float64), parts[sp][1].astype(np.float64))
                    grid.add_tile(t)
                    tiles[(i, j, k)] = t
    uid = np.zeros(128, np.uint8)
    if rank == 0:
        check(L.b2p_nccl_unique_id(uid.ctypes.data_as(C.c_void_p)))
    lst = [uid.tobytes()]
    dist.broadcast_object_list(lst, src=0)
    uid = np.frombuffer(lst[0], np.uint8).copy()
    check(L.b2p_grid_comm_init(grid._h, rank, world, uid.ctypes.data_as(C.c_void_p), owner.ctypes.data_as(C.c_void_p)))

    def compare(exact_particles, tolJ):
        for (i, j, k), t in tiles.items():
            cid = org.cid(i, j, k)
            o = org.get_fields(cid, with_halo=True)
            g = t.get_fields_f32(with_halo=True)
            assert_bits_equal(g[1], o[1], f"B {(i, j, k)}") if tolJ == 0 else None
            for name, a, b in zip("EBJ", g, o):
                err = np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)
                assert err <= max(tolJ, 0.0) or np.array_equal(a, b), (name, (i, j, k), err)
            for sp in range(2):
                og = org.get_particles(cid, sp, alive_only=False)
                gg = t.get_particles(sp, alive_only=False)
                if exact_particles:
                    alive = og[6] != DEAD
                    assert_bits_equal(gg[6], og[6], "ids")
                    for c in range(6):
                        assert_bits_equal(gg[c][alive], og[c][alive], f"particle comp {c}")

    # halo + J-exchange semantics in isolation (bit-exact)
    M = rb.comm_mode
    for mode in (M.emf_E, M.emf_B):
        org.local_communication(mode.value)
        grid.external_communication(mode); grid.local_communication(mode)
    org.local_communication(M.emf_J_exchange.value)
    grid.external_communication(M.emf_J); grid.local_communication(M.emf_J_exchange)
    org.local_communication(M.emf_J.value)
    grid.external_communication(M.emf_J); grid.local_communication(M.emf_J)
    compare(True, 0.0)
    # one snapshot file written by all ranks together (pwrite per tile row, header by rank 0) == the oracle's file,
    # byte for byte (fields after the bit-exact halo phase, density of the freshly injected particles)
    import tempfile
    d = [tempfile.mkdtemp(prefix="b2p_snap_") if rank == 0 else None]
    dist.broadcast_object_list(d, src=0)
    dist.barrier()
    grid.write_fields_snapshot(d[0], 3, 2, 2)
    rb.sync()
    dist.barrier()
    if rank == 0:
        os.makedirs(os.path.join(d[0], "o"))
        org.write_fields_snapshot(os.path.join(d[0], "o"), 3, 2, 2)
        a = open(os.path.join(d[0], "flds_3.bin"), "rb").read()
        b = open(os.path.join(d[0], "o", "flds_3.bin"), "rb").read()
        assert a == b, "multi-rank snapshot differs from the oracle's"
    dist.barrier()
    # particle migration across ranks (bit-exact containers)
    for _ in range(2):
        org.phase("push_particles"); grid.phase("push_particles")
        org.phase("pack_outgoing_particles"); grid.phase("pack_outgoing_particles")
        org.local_communication(M.pic_particle.value)
        grid.external_communication(M.pic_particle); grid.local_communication(M.pic_particle)
    compare(True, 0.0)
    # whole laps
    org.step_pic(0); grid.step_pic(0)
    compare(True, 1e-5)
    for lap in range(1, 6):
        org.step_pic(lap); grid.step_pic(lap)
    compare(False, 2e-3)
    n_local = sum(len(t.get_ids(sp)) for t in tiles.values() for sp in range(2))
    tot = [None] * world
    dist.all_gather_object(tot, n_local)
    assert sum(tot) == 2 * 4 * int(np.prod(n_cells)) * int(np.prod(n_tiles)), "particles lost in migration"
    # the kinetic-energy account (particles.cuh: KE_SLOTS) follows particles across ranks: on every rank it equals the sum
    # over that rank's containers (energy_cache = 0 forces the sum)
    grid.energies()                                      # a read: the next push keeps the account
    for lap in range(6, 9):
        grid.step_pic(lap)
        n0 = L.b2p_launch_count()
        ke = np.array(grid.energies()[2])
        n1 = L.b2p_launch_count()
        check(L.b2p_set_option(b"energy_cache", 0))
        full = np.array(grid.energies()[2])
        check(L.b2p_set_option(b"energy_cache", 1))
        n2 = L.b2p_launch_count()
        assert n1 - n0 < n2 - n1, "energies() did not use the account on a multi-rank grid"
        assert np.allclose(ke, full, rtol=2e-7, atol=0.0), (rank, lap, ke, full)
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU parity OK on {world} ranks, tiles {n_tiles}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
