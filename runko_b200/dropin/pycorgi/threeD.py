"""pycorgi.threeD (external/corgi/pycorgi/pycorgi.c++:317-374)"""
from b200_dropin import CorgiTile as Tile  # noqa: F401
from b200_dropin import Grid  # noqa: F401
