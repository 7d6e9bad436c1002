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


def set_superposition_mode(mode: str) -> None:
    """Arithmetic of the MSE superposition and of `Lorentzian.superposition_vec`: "exact" replays the
    reference's operators bit for bit; "fast" (default) halves the FP64 instructions per evaluation
    and agrees to about 1e-15 relative (1e-13 for the MSE).  Peak sets and Lorentzian parameters are identical in both."""
    from . import _lib
    from .exceptions import raise_for_status
    modes = {"exact": 0, "fast": 1}
    if mode not in modes:
        raise ValueError("mode must be 'exact' or 'fast'")
    raise_for_status(_lib.load().mdb_set_superposition_mode(modes[mode]), _lib.last_error())


def superposition_mode() -> str:
    from . import _lib
    return "fast" if _lib.load().mdb_superposition_mode() == 1 else "exact"


__all__ = ["__version__", "set_devices", "set_superposition_mode", "superposition_mode", "Deconvoluter", "Deconvolution", "Lorentzian", "Spectrum", "exceptions"]
