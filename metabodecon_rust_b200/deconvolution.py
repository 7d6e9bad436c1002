"""`Deconvolution` result container (deconvolution/deconvolution.rs:45-56; Python surface
metabodecon-python/src/bindings/deconvolution.rs)."""
from __future__ import annotations

import json
import math

import msgpack

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

    # ---- MessagePack (bindings/deconvolution.rs:97-117: rmp_serde::to_vec / from_slice).
    # rmp_serde's default ("compact") form writes every struct as an ARRAY of its fields in
    # declaration order, and serde's internally tagged enums (`tag = "method"`) put the tag first:
    #   Deconvolution        -> [smoothing, selection, fitting, mse, lorentzians]
    #   MovingAverage{..}    -> ["MovingAverage", iterations, windowSize];  Identity -> ["Identity"]
    #   NoiseScoreFilter{..} -> ["NoiseScoreFilter", ["MinimumSum"], threshold];  DetectorOnly -> ["DetectorOnly"]
    #   Analytical{..}       -> ["Analytical", iterations]
    #   Lorentzian           -> [sf, hw, maxp]            (f64 as float64, usize as the smallest uint)
    # Derived from serde's derive rules and the struct definitions (serialized_deconvolution.rs:18-31,
    # serialized_lorentzian.rs:16-24, smoother.rs:21-56, selector.rs:26-57, scorer.rs:15-29,
    # fitter.rs:28-57); not cross-checked against rmp_serde itself (no Rust toolchain here).
    # read_bin also accepts the map form that rmp_serde::to_vec_named produces.
    @staticmethod
    def _settings_to_array(d: dict) -> list:
        m = d["method"]
        if m == "MovingAverage":
            return [m, int(d["iterations"]), int(d["windowSize"])]
        if m == "NoiseScoreFilter":
            return [m, [d["scoringMethod"]["method"]], float(d["threshold"])]
        if m == "Analytical":
            return [m, int(d["iterations"])]
        return [m]

    @staticmethod
    def _settings_from_packed(v, kind: str) -> dict:
        if isinstance(v, dict):
            return v
        m = v[0]
        if m == "MovingAverage":
            return {"method": m, "iterations": int(v[1]), "windowSize": int(v[2])}
        if m == "NoiseScoreFilter":
            sc = v[1]
            return {"method": m, "scoringMethod": sc if isinstance(sc, dict) else {"method": sc[0]}, "threshold": float(v[2])}
        if m == "Analytical":
            return {"method": m, "iterations": int(v[1])}
        if m in ("Identity", "DetectorOnly", "MinimumSum"):
            return {"method": m}
        raise ValueError(f"unknown {kind} settings variant {m!r}")

    def write_bin(self, path: str) -> None:
        ser = self._to_serialized()
        try:
            blob = msgpack.packb([
                self._settings_to_array(ser["smoothingSettings"]), self._settings_to_array(ser["selectionSettings"]),
                self._settings_to_array(ser["fittingSettings"]), float(ser["mse"]),
                [[float(l["sf"]), float(l["hw"]), float(l["maxp"])] for l in ser["lorentzians"]]])
        except (KeyError, TypeError, ValueError) as err:
            raise SerializationError(str(err)) from err
        with open(path, "wb") as fh:
            fh.write(blob)

    @staticmethod
    def read_bin(path: str) -> "Deconvolution":
        with open(path, "rb") as fh:
            blob = fh.read()
        try:
            obj = msgpack.unpackb(blob, raw=False)
            if isinstance(obj, dict):
                sm, se, fi, mse, lor = (obj["smoothingSettings"], obj["selectionSettings"], obj["fittingSettings"],
                                        obj["mse"], obj["lorentzians"])
            else:
                sm, se, fi, mse, lor = obj
            rows = [((l["sf"], l["hw"], l["maxp"]) if isinstance(l, dict) else tuple(l)) for l in lor]
            params = np.array([[sf * hw, hw * hw, maxp] for sf, hw, maxp in rows], dtype=np.float64).reshape(-1, 3)
            return Deconvolution(params, float(mse), Deconvolution._settings_from_packed(sm, "smoothing"),
                                 Deconvolution._settings_from_packed(se, "selection"),
                                 Deconvolution._settings_from_packed(fi, "fitting"))
        except (KeyError, TypeError, ValueError, IndexError, msgpack.exceptions.UnpackException,
                msgpack.exceptions.ExtraData) as err:
            raise SerializationError(str(err)) from err
