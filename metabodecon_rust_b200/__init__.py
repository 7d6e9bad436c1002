"""metabodecon_rust_b200 -- the deconvolution hot path of metabodecon on NVIDIA B200 (sm_100a).

Same class surface as the reference's Python package (metabodecon-python/metabodecon/__init__.py):
`Deconvoluter`, `Deconvolution`, `Lorentzian`, `Spectrum`, `exceptions`.  All compute goes through
libmdb200.so (hand-written CUDA kernels behind the C ABI of include/mdb200.h).
"""
from . import exceptions
from .deconvoluter import Deconvoluter
from .deconvolution import Deconvolution
from .lorentzian import Lorentzian
from .spectrum import Spectrum

__version__ = "0.1.0"

__all__ = ["__version__", "Deconvoluter", "Deconvolution", "Lorentzian", "Spectrum", "exceptions"]
