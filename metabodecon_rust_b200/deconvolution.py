"""`Deconvolution` result container (deconvolution/deconvolution.rs:45-56; Python surface
metabodecon-python/src/bindings/deconvolution.rs)."""
from __future__ import annotations

import json
import math

import numpy as np

from .exceptions import SerializationError
from .lorentzian import Lorentzian, superposition_vec_array


class Deconvolution:
    def __init__(self, params: np.ndarray, mse: float, smoothing_settings: dict, selection_settings: dict,
                 fitting_settings: dict, peaks: np.ndarray | None = None) -> None:
        self._params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 3)
        self._mse = float(mse)
        self.smoothing_settings = dict(smoothing_settings)
        self.selection_settings = dict(selection_settings)
        self.fitting_settings = dict(fitting_settings)
        self._peaks = peaks  # (P, 3) selected (left, centre, right); diagnostics, not in the reference

    @property
    def lorentzians(self) -> list:
        return [Lorentzian.from_transformed(a, h, m) for a, h, m in self._params]

    @property
    def parameters(self) -> np.ndarray:
        """(K, 3) array of (sfhw, hw2, maxp) -- zero-copy view of the result."""
        return self._params

    @property
    def peaks(self):
        return self._peaks

    @property
    def mse(self) -> float:
        return self._mse

    def superposition(self, x: float) -> float:
        return float(superposition_vec_array(np.array([x], dtype=np.float64), self._params)[0])

    def superposition_vec(self, x) -> np.ndarray:
        return superposition_vec_array(x, self._params)

    def par_superposition_vec(self, x) -> np.ndarray:
        return superposition_vec_array(x, self._params)

    # ---- serde-compatible JSON (serialized_deconvolution.rs:18-31, serialized_lorentzian.rs:16-43):
    # camelCase keys, Lorentzians stored untransformed as (sf, hw, maxp).
    def _to_serialized(self) -> dict:
        lor = []
        for a, h, m in self._params:
            hw = math.sqrt(h)
            lor.append({"sf": a / hw, "hw": hw, "maxp": m})
        return {"smoothingSettings": self.smoothing_settings, "selectionSettings": self.selection_settings,
                "fittingSettings": self.fitting_settings, "mse": self._mse, "lorentzians": lor}

    def write_json(self, path: str) -> None:
        try:
            text = json.dumps(self._to_serialized(), indent=2)
        except (TypeError, ValueError) as err:
            raise SerializationError(str(err)) from err
        with open(path, "w") as fh:
            fh.write(text)

    @staticmethod
    def read_json(path: str) -> "Deconvolution":
        with open(path, "r") as fh:
            text = fh.read()
        try:
            obj = json.loads(text)
            params = np.array([[l["sf"] * l["hw"], l["hw"] * l["hw"], l["maxp"]] for l in obj["lorentzians"]],
                              dtype=np.float64).reshape(-1, 3)
            return Deconvolution(params, obj["mse"], obj["smoothingSettings"], obj["selectionSettings"],
                                 obj["fittingSettings"])
        except (KeyError, TypeError, ValueError) as err:
            raise SerializationError(str(err)) from err
