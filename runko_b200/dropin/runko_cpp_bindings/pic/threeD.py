"""runko_cpp_bindings.pic.threeD (src/runko/bindings/pypic.c++:55-139)"""
from b200_dropin import ParticleState, ParticleStateBatch, _write_average_kinetic_energy, reflector_wall  # noqa: F401
from b200_dropin import PicTile as Tile  # noqa: F401
from b200_dropin import PicVirtualTile as VirtualTile  # noqa: F401
