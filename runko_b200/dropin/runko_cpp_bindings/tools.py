"""runko_cpp_bindings.tools (src/runko/bindings/pytools.c++:23-37)"""
import enum

from b200_dropin import _get_gpu_mem_kB, _virtual_tile_sync_handshake_mode, comm_mode  # noqa: F401


class particle(enum.Enum):
    """runko::particle (src/runko/particles_common.h)"""
    electron = 0
    ion = 1
    photon = 2
