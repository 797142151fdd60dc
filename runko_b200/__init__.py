"""runko_b200 — B200 (sm_100a) implementation of runko's per-timestep PIC hot path
behind the reference's tile API.  All compute runs in libb200pic.so (hand-written
CUDA, C-ABI in include/b200pic.h); this package is the thin host-side mirror of
the reference's pybind11 interface.  No CPU fallback exists.
"""
from ._lib import B2PError, B2PLogicError, SO_PATH, lib  # noqa: F401
from .tiles import (Grid, MpiioFieldsWriter, ParticleStateBatch, ParticleStateD, PicTile, Tile, antenna_mode, comm_mode, edge_bc, reflector_wall, sync,  # noqa: F401
                    _get_gpu_mem_kB, _virtual_tile_sync_handshake_mode)
