"""Drop-in for `pycorgi` (external/corgi/pycorgi/pycorgi.c++): only the 3-D specialisation runko uses."""
from . import threeD  # noqa: F401
