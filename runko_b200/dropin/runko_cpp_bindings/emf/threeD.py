"""runko_cpp_bindings.emf.threeD (src/runko/bindings/pyemf.c++:214-281)"""
from b200_dropin import EmfTile as Tile  # noqa: F401
from b200_dropin import EmfVirtualTile as VirtualTile  # noqa: F401
from b200_dropin import (MpiioFieldsWriter, MpiioParticlesWriter, MpiioSpectraWriter, _write_average_B_energy_density,  # noqa: F401
                         _write_average_E_energy_density, antenna_mode, edge_bc)
