from . import threeD  # noqa: F401
