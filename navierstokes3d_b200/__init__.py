"""navierstokes3d_b200 -- B200-native (sm_100a) hot path of mattbuergler/NavierStokes3D.

The product is ``csrc/libns3d.so`` (hand-written CUDA kernels behind the C ABI of
``include/ns3d.h``); this package is the host side above it: the ctypes binding
(``native``), the scripts' parameter derivation (``params``) and drivers with the
reference's run-script surface (``driver.run_navierstokes3D``, ``driver.runme``).
"""
from . import native, params  # noqa: F401
from .driver import Simulation, StreamedSteps, run_navierstokes3D, runme  # noqa: F401
from .native import FAST, FASTEST, PARITY, VARIANT_G, VARIANT_M, Context, NS3DError  # noqa: F401
from .params import Physics, Setup, SlabGrid, setup_gpu, setup_multi_gpu  # noqa: F401

__version__ = "0.1.0"
