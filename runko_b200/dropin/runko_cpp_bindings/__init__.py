"""Drop-in for the reference's compiled module `runko_cpp_bindings` (src/runko/bindings/runko_cpp_bindings.c++:14-40):
submodules `tools`, `emf.threeD`, `pic.threeD`; pycorgi is imported first, as the reference does (:19)."""
import pycorgi  # noqa: F401
from . import emf, pic, tools  # noqa: F401
