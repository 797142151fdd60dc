"""Host-side logic of the drop-in boundary (no GPU): config parsing with the reference's key
names / type rules / error messages, enum values, and the multi-GPU exchange plan."""
import ctypes as C
import itertools
import re

import numpy as np
import pytest

import runko_b200 as rb
from runko_b200 import _abi
from util import Conf, pic_conf


def test_comm_mode_values_match_the_reference():
    # src/runko/communication_common.h:30-38
    M = rb.comm_mode
    assert (M.emf_J.value, M.emf_E.value, M.emf_B.value, M.pic_particle.value, M.pic_particle_extra.value) == (0, 1, 2, 3, 4)
    assert rb._virtual_tile_sync_handshake_mode(M.pic_particle) is not None
    assert rb._virtual_tile_sync_handshake_mode(M.emf_E) is None


def test_unrecognised_config_value_type_fails_like_the_reference():
    # tests/py/test_emf.py:14-31
    class foo:
        pass
    conf = Conf(n_tiles=[2, 3, 4], n_cells_per_tile=[10, 12, 14], field_propagator="fdtd2", cfl=foo())
    with pytest.raises(_abi.ConfigError, match=re.compile("cfl.*unsupported type.*foo", re.I)):
        _abi.make_config(conf)


def test_missing_and_none_keys_read_as_absent():
    conf = Conf(n_tiles=[1, 1, 1], n_cells_per_tile=[4, 4, 4], field_propagator="fdtd2", cfl=0.45, current_filter=None)
    c = _abi.make_config(conf)
    assert c.current_filter == -1 and c.n_species == 0 and c.particle_pusher == -1
    with pytest.raises(_abi.ConfigError, match="cfl"):
        _abi.make_config(Conf(n_tiles=[1, 1, 1], n_cells_per_tile=[4, 4, 4], field_propagator="fdtd2"))
    with pytest.raises(_abi.ConfigError, match="not supported field propagator"):
        _abi.make_config(Conf(n_tiles=[1, 1, 1], n_cells_per_tile=[4, 4, 4], field_propagator="fdtd9", cfl=1))


def test_species_are_the_contiguous_q_m_pairs():
    # pic/tile.c++:52-71: q0/m0, q1/m1, ... until the first gap
    conf = pic_conf(q2=2.0, m2=5.0, q4=1.0, m4=1.0)
    c = _abi.make_config(conf)
    assert c.n_species == 3 and list(c.q[:3]) == [-1.0, 1.0, 2.0] and c.m[2] == 5.0


def test_stencil_keys_fill_the_coefficient_matrix():
    # emf/tile.c++:99-142: axis-specific key wins over the generic one
    conf = Conf(n_tiles=[1, 1, 1], n_cells_per_tile=[4, 4, 4], field_propagator="stencil", cfl=0.45,
                stencil_delta=0.1, stencil_x_delta=0.2, stencil_z_zeta3_p2=-0.3)
    c = _abi.make_config(conf)
    assert c.stencil[0][1][0] == pytest.approx(0.2) and c.stencil[1][1][0] == pytest.approx(0.1)
    assert c.stencil[2][2][4] == pytest.approx(-0.3) and c.stencil[0][2][4] == 0.0


def plan(cfg, owner, rank):
    rows = np.zeros((4096, 7), np.int64)
    n = rb.lib().b2p_plan_describe(C.byref(cfg), owner.ctypes.data_as(C.c_void_p), rank, rows.ctypes.data_as(C.c_void_p), 4096)
    assert 0 <= n <= 4096
    return rows[:n]


@pytest.mark.parametrize("n_tiles,blocks", [((2, 1, 1), (2, 1, 1)), ((4, 2, 2), (2, 2, 1)), ((2, 2, 2), (2, 2, 2)),
                                              ((4, 4, 2), (2, 2, 2)), ((3, 1, 1), (3, 1, 1))])
def test_exchange_plan_pairs_up(n_tiles, blocks):
    """For every rank pair the sender's slab list (sorted by send_key) is the receiver's list
    (sorted by recv_key): same keys, same sizes -> tag-free matched NCCL sends/recvs; and every
    (tile, direction) has either a local or a planned remote neighbour."""
    cfg = _abi.make_config(pic_conf(n_tiles=n_tiles, n_cells=(5, 6, 7)))
    T = n_tiles
    tpb = [T[d] // blocks[d] for d in range(3)]
    owner = np.zeros(T[0] * T[1] * T[2], np.int32)
    for k, j, i in itertools.product(range(T[2]), range(T[1]), range(T[0])):
        owner[i + T[0] * (j + T[1] * k)] = (i // tpb[0]) + blocks[0] * ((j // tpb[1]) + blocks[1] * (k // tpb[2]))
    nr = int(owner.max()) + 1
    plans = [plan(cfg, owner, r) for r in range(nr)]
    for a in range(nr):
        mine = plans[a]
        assert np.all(owner[mine[:, 1]] == a) and np.all(owner[mine[:, 3]] == mine[:, 0])
        # 26 directions per owned tile = local + remote
        for cid in np.flatnonzero(owner == a):
            i, j, k = cid % T[0], (cid // T[0]) % T[1], cid // (T[0] * T[1])
            remote = 0
            for dk, dj, di in itertools.product((-1, 0, 1), repeat=3):
                if (di, dj, dk) == (0, 0, 0):
                    continue
                oc = (i + di) % T[0] + T[0] * ((j + dj) % T[1] + T[1] * ((k + dk) % T[2]))
                remote += owner[oc] != a
            assert remote == np.count_nonzero(mine[:, 1] == cid)
        for b in range(nr):
            if a == b:
                continue
            send = mine[mine[:, 0] == b]
            recv = plans[b][plans[b][:, 0] == a]
            send = send[np.argsort(send[:, 4], kind="stable")]
            recv = recv[np.argsort(recv[:, 5], kind="stable")]
            assert len(send) == len(recv)
            assert np.array_equal(send[:, 4], recv[:, 5]), "send/recv key order differs"
            assert np.array_equal(send[:, 6], recv[:, 6]), "slab sizes differ"
            assert np.array_equal(send[:, 1], recv[:, 3]) and np.array_equal(send[:, 3], recv[:, 1])


def test_sorting_network_header_is_what_the_generator_writes():
    """runko_b200/csrc/sortnet.cuh (the compare-exchange lists k_sort_cells orders a cell's member list with) is generated by
    tools/gen_sortnet.py, which checks the networks against sorted() before writing: the committed header must be its output."""
    import importlib.util
    import os
    import random
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_sortnet", os.path.join(root, "tools", "gen_sortnet.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text = open(os.path.join(root, "runko_b200", "csrc", "sortnet.cuh")).read()
    for n, pairs in ((16, 63), (32, 191)):
        ces = gen.batcher(n)
        assert len(ces) == pairs
        assert " ".join(f"CE({a},{b})" for a, b in ces) in " ".join(text.replace("\\\n", " ").split())
        for _ in range(200):                                  # ... and they sort (ties included)
            v = [random.randrange(40) for _ in range(n)]
            w = v[:]
            for a, b in ces:
                if w[a] > w[b]:
                    w[a], w[b] = w[b], w[a]
            assert w == sorted(v)
