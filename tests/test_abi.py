"""The C-ABI boundary without a GPU: the library loads, exports every symbol the header
declares, the ctypes struct mirrors match the C layouts, compute entry points fail loudly
without a CUDA device, and the product never touches the oracle."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

import runko_b200 as rb
from runko_b200 import _abi
from runko_b200._lib import SYMBOLS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200pic.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2p_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = rb.lib()
    names = declared_functions()
    assert len(names) >= 50
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/b200pic.h but not exported by libb200pic.so"
    assert sorted(SYMBOLS) == names, "runko_b200/_lib.py SYMBOLS is out of sync with include/b200pic.h"


def test_header_cites_the_reference_interfaces():
    src = open(HEADER).read()
    for cite in ("emf/tile.c++:359-375", "pic/tile.c++:326-365", "pic/tile.c++:369-415", "pic/tile_communication.c++:68-96",
                 "corgi.h:1697-1718", "communication_common.h:30-38", "particles_common.h:26-34"):
        assert cite in src, cite


def test_struct_layouts_match_the_header():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "b200pic.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(b2p_config), offsetof(b2p_config, cfl), offsetof(b2p_config, stencil),
         offsetof(b2p_config, n_species), offsetof(b2p_config, q), offsetof(b2p_config, particle_pusher),
         offsetof(b2p_config, prealloc_per_species), sizeof(b2p_particle_state));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        got = [int(v) for v in subprocess.check_output([exe]).split()]
    K = _abi.B2PConfig
    want = [C.sizeof(K), K.cfl.offset, K.stencil.offset, K.n_species.offset, K.q.offset, K.particle_pusher.offset,
            K.prealloc_per_species.offset, C.sizeof(_abi.ParticleState)]
    assert got == want


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point reports an error (never a silent CPU path)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    L = rb.lib()
    assert L.b2p_init(0) != 0
    assert b"no CPU fallback" in L.b2p_last_error()
    from util import pic_conf
    with pytest.raises(rb.B2PError):
        rb.PicTile((0, 0, 0), pic_conf())


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "runko_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "oracle" not in txt.replace("tests/refshim binds the same logic to the CPU oracle", ""), os.path.join(base, f)


def test_plan_describe_is_host_only():
    """b2p_plan_describe needs neither a GPU nor NCCL."""
    import numpy as np
    from util import pic_conf
    cfg = _abi.make_config(pic_conf(n_tiles=(2, 2, 1)))
    owner = np.array([0, 1, 0, 1], np.int32)
    rows = np.zeros((256, 7), np.int64)
    n = rb.lib().b2p_plan_describe(C.byref(cfg), owner.ctypes.data_as(C.c_void_p), 0, rows.ctypes.data_as(C.c_void_p), 256)
    assert n > 0 and n <= 2 * 26
    assert set(rows[:n, 0]) == {1}


def test_missing_library_fails_loudly():
    """No silent fallback when the CUDA library is absent: importing the package's binding with the
    library path pointing nowhere (B2P_LIB, the experiment override, or a deleted libb200pic.so)
    raises ImportError naming the missing file."""
    import subprocess
    import sys
    code = ("import runko_b200\n"
            "try:\n"
            "    runko_b200.lib()\n"
            "except ImportError as e:\n"
            "    assert 'is missing' in str(e) and 'no CPU fallback' in str(e), str(e)\n"
            "    print('raised')\n")
    env = dict(os.environ, B2P_LIB="/nonexistent/libb200pic.so")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=root, timeout=300)
    assert r.returncode == 0 and "raised" in r.stdout, r.stdout + r.stderr
