"""World-size-2 CPU emulation of the multi-GPU halo exchange (launched by
tests/test_multirank_gloo.py): each rank owns half of the tiles, gets its exchange plan from
the product's host code (b2p_plan_describe, the same build_plan() comm.cu uses), packs the
slabs in plan order into ONE buffer per peer, ships them over torch.distributed/gloo (standing
in for the grouped ncclSend/ncclRecv), unpacks by recv_key order, and must reproduce what the
single-process oracle computes for the whole periodic grid — for E/B/J halo fill and for the
J exchange.  Validates keys, ordering, slab geometry and the tag-free pairing without a GPU."""
import ctypes as C
import itertools
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import runko_b200 as rb  # noqa: E402
from oracle.oracle import OracleGrid  # noqa: E402
from runko_b200 import _abi  # noqa: E402
from util import pic_conf, random_lattice  # noqa: E402

H = 3


def region(d, n, corresponding):
    """emf/yee_lattice.h:493-517,580-602 along one axis"""
    if d == 0:
        return slice(H, H + n)
    if corresponding:
        return slice(n, n + H) if d == -1 else slice(H, 2 * H)
    return slice(0, H) if d == -1 else slice(H + n, n + 2 * H)


def dirs_of(index):
    return (index // 9 - 1, (index // 3) % 3 - 1, index % 3 - 1)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    n_tiles, n = (4, 2, 1), (4, 5, 6)
    conf = pic_conf(n_tiles=n_tiles, n_cells=n)
    cfg = _abi.make_config(conf)
    T = n_tiles
    owner = np.array([(c % T[0]) // 2 for c in range(T[0] * T[1] * T[2])], np.int32)      # 2x1x1 blocks of 2x2x1 tiles
    rows = np.zeros((1024, 7), np.int64)
    nrow = rb.lib().b2p_plan_describe(C.byref(cfg), owner.ctypes.data_as(C.c_void_p), rank, rows.ctypes.data_as(C.c_void_p), 1024)
    rows = rows[:nrow]
    peer = 1 - rank
    assert set(rows[:, 0]) == {peer}

    rng = np.random.default_rng(5)                     # the same global state on both ranks
    org = OracleGrid(conf)
    pre = {}
    for c in range(org.num_tiles):
        pre[c] = [random_lattice(rng, n) for _ in range(3)]
        org.set_fields(c, *pre[c], with_halo=True)

    def exchange(which, kind):
        """kind 0: my corresponding_subregion(-dir) (feeds the peer's halo fill);
           kind 1: my subregion(dir) (my halo facing the peer; feeds the peer's J exchange)."""
        send = rows[np.argsort(rows[:, 4], kind="stable")]
        recv = rows[np.argsort(rows[:, 5], kind="stable")]
        chunks = []
        for r in send:
            d = dirs_of(int(r[2]))
            src = pre[int(r[1])][which]
            sl = tuple(region(-d[a], n[a], True) if kind == 0 else region(d[a], n[a], False) for a in range(3))
            chunks.append(np.ascontiguousarray(src[(slice(None),) + sl]).ravel())
        sbuf = torch.from_numpy(np.concatenate(chunks))
        rbuf = torch.empty(int(recv[:, 6].sum()), dtype=torch.float32)
        assert sbuf.numel() == rbuf.numel()
        reqs = [dist.isend(sbuf, peer), dist.irecv(rbuf, peer)]
        for q in reqs:
            q.wait()
        out, off = {}, 0
        for r in recv:
            d = dirs_of(int(r[2]))
            dims = tuple(H if d[a] else n[a] for a in range(3))
            size = int(r[6])
            out[(int(r[1]), int(r[2]))] = rbuf[off:off + size].numpy().reshape((3,) + dims)
            off += size
        return out

    mine = [c for c in range(org.num_tiles) if owner[c] == rank]
    # ---- halo fill of E, B, J --------------------------------------------------------------
    got = {c: [a.copy() for a in pre[c]] for c in mine}
    for which, mode in ((0, 1), (1, 2), (2, 0)):
        slabs = exchange(which, 0)
        for c in mine:
            i, j, k = c % T[0], (c // T[0]) % T[1], c // (T[0] * T[1])
            for dk, dj, di in itertools.product((-1, 0, 1), repeat=3):
                if (di, dj, dk) == (0, 0, 0):
                    continue
                d = (di, dj, dk)
                oc = (i + di) % T[0] + T[0] * ((j + dj) % T[1] + T[1] * ((k + dk) % T[2]))
                dst = (slice(None),) + tuple(region(d[a], n[a], False) for a in range(3))
                if owner[oc] == rank:
                    src = (slice(None),) + tuple(region(d[a], n[a], True) for a in range(3))
                    got[c][which][dst] = pre[oc][which][src]
                else:
                    got[c][which][dst] = slabs[(c, ((di + 1) * 3 + (dj + 1)) * 3 + (dk + 1))]
        org.local_communication(mode)
    for c in mine:
        ref = org.get_fields(c, with_halo=True)
        for which in range(3):
            assert np.array_equal(got[c][which], ref[which]), ("halo fill", c, which)
    # ---- J exchange (reads the neighbours' halos as they were BEFORE the exchange) -------------
    for c in range(org.num_tiles):
        pre[c] = list(org.get_fields(c, with_halo=True))
    slabs = exchange(2, 1)
    gotJ = {c: pre[c][2].copy() for c in mine}
    for c in mine:
        i, j, k = c % T[0], (c // T[0]) % T[1], c // (T[0] * T[1])
        for dk, dj, di in itertools.product((-1, 0, 1), repeat=3):      # Moore order kr -> jr -> ir
            if (di, dj, dk) == (0, 0, 0):
                continue
            d = (di, dj, dk)
            oc = (i + di) % T[0] + T[0] * ((j + dj) % T[1] + T[1] * ((k + dk) % T[2]))
            dst = (slice(None),) + tuple(region(-d[a], n[a], True) for a in range(3))
            if owner[oc] == rank:
                src = (slice(None),) + tuple(region(-d[a], n[a], False) for a in range(3))
                gotJ[c][dst] = gotJ[c][dst] + pre[oc][2][src]
            else:
                gotJ[c][dst] = gotJ[c][dst] + slabs[(c, ((di + 1) * 3 + (dj + 1)) * 3 + (dk + 1))]
    org.local_communication(6)
    for c in mine:
        assert np.array_equal(gotJ[c], org.get_fields(c, with_halo=True)[2]), ("J exchange", c)
    dist.barrier()
    if rank == 0:
        print("gloo world-size-2 exchange emulation OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
