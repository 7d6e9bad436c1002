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


def set_devices(n: int) -> None:
    """GPUs used by one `deconvolute_spectra` call: 1 = current device (default), 0 = all visible,
    n = CUDA devices 0..n-1 (contiguous shards of the batch, one pipeline per GPU, no exchange)."""
    from . import _lib
    from .exceptions import raise_for_status
    raise_for_status(_lib.load().mdb_set_device_count(int(n)), _lib.last_error())


__all__ = ["__version__", "set_devices", "Deconvoluter", "Deconvolution", "Lorentzian", "Spectrum", "exceptions"]
