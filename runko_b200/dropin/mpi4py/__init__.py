"""Stand-in for `mpi4py` (absent in this image; imported by runko/simulation.py:21, tile_grid.py:8,
runko_logging.py:19, auto_tile_grid.py:4): the handful of COMM_WORLD calls the reference's Python layer makes.
One process per GPU; rank / size come from the launcher's environment (torchrun: RANK / WORLD_SIZE), collectives
from torch.distributed (gloo) when more than one rank runs.  The particle / field data path never goes through this
module — that is NCCL inside libb200pic.so."""
from . import MPI  # noqa: F401
