"""jets_b200 -- B200-native operator-application path of Jets.jl behind the Jets API.

The directory is named ``jets.jl_b200`` (the framework's name); import it as ``jets_b200``
through the loader at the repository root.  Everything numerical runs in ``libjets_b200.so``
(hand-written sm_100a CUDA behind the C ABI in ``include/jets_b200.h``); importing this package
fails loudly when the library has not been built.
"""
from ._lib import JetsError, init, lib, LIB_PATH, SIGNATURES  # noqa: F401
from .core import *  # noqa: F401,F403
from .core import (DeviceArray, JetSpace, JetBSpace, JopNl, JopLn, JopAdjoint, Jop)  # noqa: F401
from . import solvers  # noqa: F401
from . import dist  # noqa: F401
from . import pipeline  # noqa: F401
