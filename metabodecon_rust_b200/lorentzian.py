"""`Lorentzian` with the reference's Python surface (metabodecon-python/src/bindings/lorentzian.rs).

Stored in the transformed form (sfhw, hw2, maxp) of deconvolution/lorentzian.rs:138-145.
The vector forms run on the GPU through mdb_superposition_vec; scalar accessors are plain f64
arithmetic in the reference's operation order.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from .exceptions import raise_for_status


def _as_params(lorentzians) -> np.ndarray:
    """list[Lorentzian] | (P,3) array -> contiguous (P,3) f64 array of (sfhw, hw2, maxp)."""
    if isinstance(lorentzians, np.ndarray):
        arr = np.ascontiguousarray(lorentzians, dtype=np.float64)
        if arr.ndim != 2 or arr.shape[1] != 3:
            raise ValueError("expected an array of shape (P, 3) holding (sfhw, hw2, maxp)")
        return arr
    out = np.empty((len(lorentzians), 3), dtype=np.float64)
    for i, l in enumerate(lorentzians):
        out[i, 0], out[i, 1], out[i, 2] = l.sfhw, l.hw2, l._maxp
    return out


def superposition_vec_array(x, params: np.ndarray, mode=None) -> np.ndarray:
    """GPU superposition of the (P,3) parameter array at x (lorentzian.rs:631-635).  mode: None = the
    process default (`set_superposition_mode`), "exact" / "fast" = stated by the caller."""
    lib = _lib.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    if x.ndim != 1:
        raise ValueError("x must be one-dimensional")
    params = np.ascontiguousarray(params, dtype=np.float64)
    out = np.empty_like(x)
    if mode is None:
        st = lib.mdb_superposition_vec(x.ctypes.data, x.size, params.ctypes.data, params.shape[0],
                                       out.ctypes.data, _lib.MDB_MEM_HOST)
    else:
        code = {"exact": _lib.MDB_SUPERPOSITION_EXACT, "fast": _lib.MDB_SUPERPOSITION_FAST}[mode]
        st = lib.mdb_superposition_vec_mode(x.ctypes.data, x.size, params.ctypes.data, params.shape[0],
                                            out.ctypes.data, _lib.MDB_MEM_HOST, code)
    raise_for_status(st, _lib.last_error())
    return out


class Lorentzian:
    __slots__ = ("sfhw", "hw2", "_maxp")

    def __init__(self, sf: float, hw: float, maxp: float) -> None:
        # bindings/lorentzian.rs:26-30: (sf * hw, hw^2, maxp)
        self.sfhw = float(sf) * float(hw)
        self.hw2 = float(hw) * float(hw)
        self._maxp = float(maxp)

    @staticmethod
    def from_transformed(sfhw: float, hw2: float, maxp: float) -> "Lorentzian":
        obj = Lorentzian.__new__(Lorentzian)
        obj.sfhw, obj.hw2, obj._maxp = float(sfhw), float(hw2), float(maxp)
        return obj

    # lorentzian.rs:406-430, 477-504
    @property
    def sf(self) -> float:
        return self.sfhw / self.hw

    @sf.setter
    def sf(self, sf: float) -> None:
        self.sfhw = float(sf) * self.hw

    @property
    def hw(self) -> float:
        return math.sqrt(self.hw2)

    @hw.setter
    def hw(self, hw: float) -> None:
        self.sfhw = self.sf * float(hw)
        self.hw2 = float(hw) * float(hw)

    @property
    def maxp(self) -> float:
        return self._maxp

    @maxp.setter
    def maxp(self, maxp: float) -> None:
        self._maxp = float(maxp)

    def evaluate(self, x: float) -> float:  # lorentzian.rs:546-548
        d = float(x) - self._maxp
        return self.sfhw / (self.hw2 + d * d)

    def evaluate_vec(self, x) -> np.ndarray:  # lorentzian.rs:563-565 (0.0 + t == t)
        # always the exact arithmetic: agrees bit for bit with the scalar `evaluate` above
        return superposition_vec_array(x, _as_params([self]), mode="exact")

    def integral(self) -> float:  # lorentzian.rs:580-582
        return math.pi * self.sf

    @staticmethod
    def superposition(x: float, lorentzians) -> float:  # lorentzian.rs:606-611
        # a single point: always the exact arithmetic (the reference's bit pattern)
        return float(superposition_vec_array(np.array([x], dtype=np.float64), _as_params(lorentzians), mode="exact")[0])

    @staticmethod
    def superposition_vec(x, lorentzians) -> np.ndarray:  # lorentzian.rs:631-635
        return superposition_vec_array(x, _as_params(lorentzians))

    @staticmethod
    def par_superposition_vec(x, lorentzians) -> np.ndarray:  # lorentzian.rs:656-663 (same result)
        return superposition_vec_array(x, _as_params(lorentzians))

    def __repr__(self) -> str:
        return f"Lorentzian(sfhw={self.sfhw!r}, hw2={self.hw2!r}, maxp={self._maxp!r})"
