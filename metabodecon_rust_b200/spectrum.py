"""`Spectrum` with the reference's Python surface (metabodecon-python/src/bindings/spectrum.rs),
reduced to what the deconvolution path consumes plus the two file readers of SURVEY.md §8f.

Validation is Spectrum::new (spectrum/spectrum.rs:179-200) via mdb_spectrum_validate.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import msgpack
import numpy as np

from . import _lib
from .exceptions import SerializationError, raise_for_status

_NUCLEI = {"1H": "1H", "PROTON": "1H", "HYDROGEN1": "1H", "11B": "11B", "BORON11": "11B", "13C": "13C", "CARBON13": "13C",
           "15N": "15N", "NITROGEN15": "15N", "19F": "19F", "FLUORINE19": "19F", "29SI": "29Si", "SILICON29": "29Si",
           "31P": "31P", "PHOSPHORUS31": "31P"}  # spectrum/meta/nucleus.rs:22-45, 55-68


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

    # ---- serde-compatible storage (spectrum/serialized_spectrum.rs:17-31, bindings/spectrum.rs:194-235).
    # JSON: camelCase keys; the axis is stored as (first, last) + size and rebuilt as first + i * step on
    # reading, exactly as TryFrom<SerializedSpectrum> does (:52-58).  MessagePack: rmp_serde's compact form,
    # structs as arrays in field order: [spectrumBoundaries, signalBoundaries, size, nucleus, frequency,
    # referenceCompound, intensities] with referenceCompound = [chemicalShift, index, (name), (method)]
    # (derived from the serde attributes, not cross-checked against rmp_serde: no Rust toolchain here).
    def _serialized(self) -> dict:
        ref = dict(self.reference_compound)
        rc = {"chemicalShift": float(ref.get("chemical_shift", self._x[0])), "index": int(ref.get("index", 0))}
        if ref.get("name") is not None:
            rc["name"] = str(ref["name"])
        if ref.get("method") is not None:
            rc["method"] = str(ref["method"])
        nucleus = _NUCLEI.get(str(self.nucleus).upper(), str(self.nucleus))
        return {"spectrumBoundaries": [float(self._x[0]), float(self._x[-1])],
                "signalBoundaries": [float(self._sb[0]), float(self._sb[1])], "size": int(self._x.size),
                "nucleus": nucleus, "frequency": float(self.frequency), "referenceCompound": rc,
                "intensities": [float(v) for v in self._y]}

    @staticmethod
    def _from_serialized(obj: dict) -> "Spectrum":
        size = int(obj["size"])
        start, end = (float(v) for v in obj["spectrumBoundaries"])
        step = (end - start) / (float(size) - 1.0)
        x = start + np.arange(size, dtype=np.float64) * step
        sp = Spectrum(x, np.asarray(obj["intensities"], dtype=np.float64), tuple(obj["signalBoundaries"]))
        sp.nucleus = _NUCLEI.get(str(obj["nucleus"]).upper(), str(obj["nucleus"]))
        sp.frequency = float(obj["frequency"])
        rc = obj["referenceCompound"]
        sp.reference_compound = {"chemical_shift": float(rc["chemicalShift"]), "index": int(rc["index"])}
        for key in ("name", "method"):
            if rc.get(key) is not None:
                sp.reference_compound[key] = rc[key]
        return sp

    def write_json(self, path: str) -> None:
        with open(path, "w") as fh:
            fh.write(json.dumps(self._serialized(), indent=2))

    @staticmethod
    def read_json(path: str) -> "Spectrum":
        with open(path, "r") as fh:
            text = fh.read()
        try:
            return Spectrum._from_serialized(json.loads(text))
        except (KeyError, TypeError, ValueError, IndexError) as err:
            raise SerializationError(str(err)) from err

    def write_bin(self, path: str) -> None:
        o = self._serialized()
        rc = o["referenceCompound"]
        ref = [rc["chemicalShift"], rc["index"]] + [rc[k] for k in ("name", "method") if k in rc]
        blob = msgpack.packb([o["spectrumBoundaries"], o["signalBoundaries"], o["size"], o["nucleus"], o["frequency"], ref,
                              o["intensities"]])
        with open(path, "wb") as fh:
            fh.write(blob)

    @staticmethod
    def read_bin(path: str) -> "Spectrum":
        with open(path, "rb") as fh:
            blob = fh.read()
        try:
            obj = msgpack.unpackb(blob, raw=False)
            if not isinstance(obj, dict):
                sb, sig, size, nucleus, freq, ref, ys = obj
                if not isinstance(ref, dict):
                    rc = {"chemicalShift": ref[0], "index": ref[1]}
                    if len(ref) > 2:   # optional fields are skipped when None: a lone third entry is read as the name
                        rc["name"] = ref[2]
                    if len(ref) > 3:
                        rc["method"] = ref[3]
                    ref = rc
                obj = {"spectrumBoundaries": sb, "signalBoundaries": sig, "size": size, "nucleus": nucleus,
                       "frequency": freq, "referenceCompound": ref, "intensities": ys}
            return Spectrum._from_serialized(obj)
        except (KeyError, TypeError, ValueError, IndexError, msgpack.exceptions.UnpackException,
                msgpack.exceptions.ExtraData) as err:
            raise SerializationError(str(err)) from err
