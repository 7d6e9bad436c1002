"""`Spectrum` with the reference's Python surface (metabodecon-python/src/bindings/spectrum.rs),
reduced to what the deconvolution path consumes plus the two file readers of SURVEY.md §8f.

Validation is Spectrum::new (spectrum/spectrum.rs:179-200) via mdb_spectrum_validate.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .exceptions import raise_for_status


class Spectrum:
    def __init__(self, chemical_shifts, intensities, signal_boundaries) -> None:
        lib = _lib.load()
        x = np.ascontiguousarray(chemical_shifts, dtype=np.float64)
        y = np.ascontiguousarray(intensities, dtype=np.float64)
        if x.ndim != 1 or y.ndim != 1:
            raise ValueError("chemical_shifts and intensities must be one-dimensional")
        sb = (C.c_double * 2)(float(signal_boundaries[0]), float(signal_boundaries[1]))
        ordered = (C.c_double * 2)()
        st = lib.mdb_spectrum_validate(x.ctypes.data, x.size, y.ctypes.data, y.size, sb, ordered)
        raise_for_status(st, _lib.last_error())
        self._x = x
        self._y = y
        self._sb = (ordered[0], ordered[1])
        self.nucleus = "1H"
        self.frequency = 1.0
        self.reference_compound = {"chemical_shift": float(x[0]), "index": 0}

    @property
    def chemical_shifts(self) -> np.ndarray:
        return self._x

    @property
    def intensities(self) -> np.ndarray:
        return self._y

    @property
    def signal_boundaries(self):
        return self._sb

    @signal_boundaries.setter
    def signal_boundaries(self, value) -> None:
        lib = _lib.load()
        sb = (C.c_double * 2)(float(value[0]), float(value[1]))
        ordered = (C.c_double * 2)()
        st = lib.mdb_spectrum_validate(self._x.ctypes.data, self._x.size, self._y.ctypes.data,
                                       self._y.size, sb, ordered)
        raise_for_status(st, _lib.last_error())
        self._sb = (ordered[0], ordered[1])

    def __len__(self) -> int:
        return self._x.size

    # ---- readers (host-side I/O; not part of the accelerated path)
    @staticmethod
    def read_bruker(path: str, experiment: int, processing: int, signal_boundaries) -> "Spectrum":
        from .readers import read_bruker_arrays
        x, y, meta = read_bruker_arrays(path, experiment, processing)
        sp = Spectrum(x, y, signal_boundaries)
        sp.nucleus = meta["nucleus"]
        sp.frequency = meta["frequency"]
        return sp

    @staticmethod
    def read_bruker_set(path: str, experiment: int, processing: int, signal_boundaries):
        entries = sorted(e for e in os.listdir(path) if os.path.isdir(os.path.join(path, e)))
        return [Spectrum.read_bruker(os.path.join(path, e), experiment, processing, signal_boundaries)
                for e in entries]

    @staticmethod
    def read_jcampdx(path: str, signal_boundaries) -> "Spectrum":
        from .readers import read_jcampdx_arrays
        x, y, meta = read_jcampdx_arrays(path)
        sp = Spectrum(x, y, signal_boundaries)
        sp.nucleus = meta.get("nucleus", "1H")
        sp.frequency = meta.get("frequency", 1.0)
        return sp

    @staticmethod
    def read_jcampdx_set(path: str, signal_boundaries):
        entries = sorted(e for e in os.listdir(path) if e.endswith(".dx"))
        return [Spectrum.read_jcampdx(os.path.join(path, e), signal_boundaries) for e in entries]
