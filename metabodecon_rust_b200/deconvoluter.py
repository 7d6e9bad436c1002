"""`Deconvoluter` with the reference's Python surface
(metabodecon-python/src/bindings/deconvoluter.rs:9-170) on top of the C ABI.

All four deconvolution entry points run the same GPU pipeline: the reference's serial and rayon
variants produce identical results (per-point summation order is unchanged), so there is nothing
to distinguish.  Batch semantics follow deconvoluter.rs:655-658: the first failing spectrum in
index order raises, and no partial result is returned.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .deconvolution import Deconvolution
from .exceptions import raise_for_status
from .spectrum import Spectrum


class Deconvoluter:
    def __init__(self) -> None:
        lib = _lib.load()
        handle = C.c_void_p()
        raise_for_status(lib.mdb_deconvoluter_default(C.byref(handle)), _lib.last_error())
        self._h = handle

    def __del__(self) -> None:
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.load().mdb_deconvoluter_free(h)
            except Exception:
                pass
            self._h = None

    # ---- settings (bindings/deconvoluter.rs:23-90)
    def set_identity_smoother(self) -> None:
        s = _lib.SmoothingSettings(_lib.MDB_SMOOTHING_IDENTITY, 0, 0)
        raise_for_status(_lib.load().mdb_deconvoluter_set_smoothing_settings(self._h, C.byref(s)), _lib.last_error())

    def set_moving_average_smoother(self, iterations: int, window_size: int) -> None:
        if iterations < 0 or window_size < 0:
            raise OverflowError("can't convert negative int to unsigned")
        s = _lib.SmoothingSettings(_lib.MDB_SMOOTHING_MOVING_AVERAGE, iterations, window_size)
        raise_for_status(_lib.load().mdb_deconvoluter_set_smoothing_settings(self._h, C.byref(s)), _lib.last_error())

    def set_detector_only(self) -> None:
        s = _lib.SelectionSettings(_lib.MDB_SELECTION_DETECTOR_ONLY, 0, 0.0)
        raise_for_status(_lib.load().mdb_deconvoluter_set_selection_settings(self._h, C.byref(s)), _lib.last_error())

    def set_noise_score_selector(self, threshold: float) -> None:
        s = _lib.SelectionSettings(_lib.MDB_SELECTION_NOISE_SCORE_FILTER, _lib.MDB_SCORING_MINIMUM_SUM, float(threshold))
        raise_for_status(_lib.load().mdb_deconvoluter_set_selection_settings(self._h, C.byref(s)), _lib.last_error())

    def set_analytical_fitter(self, iterations: int) -> None:
        if iterations < 0:
            raise OverflowError("can't convert negative int to unsigned")
        s = _lib.FittingSettings(_lib.MDB_FITTING_ANALYTICAL, iterations)
        raise_for_status(_lib.load().mdb_deconvoluter_set_fitting_settings(self._h, C.byref(s)), _lib.last_error())

    def add_ignore_region(self, boundaries) -> None:
        raise_for_status(_lib.load().mdb_deconvoluter_add_ignore_region(self._h, float(boundaries[0]), float(boundaries[1])),
                         _lib.last_error())

    def clear_ignore_regions(self) -> None:
        _lib.load().mdb_deconvoluter_clear_ignore_regions(self._h)

    # ---- no counterpart in the reference: arithmetic of the MSE superposition of THIS deconvoluter
    def set_superposition_mode(self, mode: str) -> None:
        """Pin the arithmetic of `Deconvolution.mse` for this deconvoluter: "exact" (the reference's bit
        patterns) or "fast" (few-ulp terms, MSE within ~1e-13 relative).  A deconvoluter that was never
        pinned follows the process default (`metabodecon_rust_b200.set_superposition_mode`)."""
        modes = {"exact": _lib.MDB_SUPERPOSITION_EXACT, "fast": _lib.MDB_SUPERPOSITION_FAST}
        if mode not in modes:
            raise ValueError("mode must be 'exact' or 'fast'")
        raise_for_status(_lib.load().mdb_deconvoluter_set_superposition_mode(self._h, modes[mode]), _lib.last_error())

    def superposition_mode(self) -> str:
        return "fast" if _lib.load().mdb_deconvoluter_superposition_mode(self._h) == _lib.MDB_SUPERPOSITION_FAST else "exact"

    def set_threads(self, threads: int) -> None:
        # bindings/deconvoluter.rs:92-106 validates the count; the GPU path has no pool to size.
        if threads <= 1:
            raise ValueError("number of threads must be greater than 1")

    def clear_threads(self) -> None:
        pass

    # ---- settings read-back (deconvoluter.rs:229-298), serde-shaped dicts
    def smoothing_settings(self) -> dict:
        s = _lib.SmoothingSettings()
        _lib.load().mdb_deconvoluter_smoothing_settings(self._h, C.byref(s))
        if s.kind == _lib.MDB_SMOOTHING_IDENTITY:
            return {"method": "Identity"}
        return {"method": "MovingAverage", "iterations": int(s.iterations), "windowSize": int(s.window_size)}

    def selection_settings(self) -> dict:
        s = _lib.SelectionSettings()
        _lib.load().mdb_deconvoluter_selection_settings(self._h, C.byref(s))
        if s.kind == _lib.MDB_SELECTION_DETECTOR_ONLY:
            return {"method": "DetectorOnly"}
        return {"method": "NoiseScoreFilter", "scoringMethod": {"method": "MinimumSum"}, "threshold": float(s.threshold)}

    def fitting_settings(self) -> dict:
        s = _lib.FittingSettings()
        _lib.load().mdb_deconvoluter_fitting_settings(self._h, C.byref(s))
        return {"method": "Analytical", "iterations": int(s.iterations)}

    def ignore_regions(self):
        lib = _lib.load()
        n = lib.mdb_deconvoluter_ignore_regions(self._h, None, 0)
        if n < 0:
            return None
        buf = (C.c_double * (2 * n))()
        lib.mdb_deconvoluter_ignore_regions(self._h, buf, n)
        return [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]

    # ---- deconvolution (bindings/deconvoluter.rs:112-165)
    def _run(self, spectra, memory: int = _lib.MDB_MEM_HOST):
        lib = _lib.load()
        n = len(spectra)
        views = (_lib.SpectrumView * max(n, 1))()
        for i, sp in enumerate(spectra):
            if not isinstance(sp, Spectrum):
                raise TypeError("expected a Spectrum")
            views[i].chemical_shifts = sp.chemical_shifts.ctypes.data
            views[i].intensities = sp.intensities.ctypes.data
            views[i].len = sp.chemical_shifts.size
            views[i].signal_boundaries[0], views[i].signal_boundaries[1] = sp.signal_boundaries
        batch = C.c_void_p()
        st = lib.mdb_deconvolute_spectra(self._h, views, n, memory, C.byref(batch))
        try:
            raise_for_status(st, _lib.last_error())
            sm, se, fi = self.smoothing_settings(), self.selection_settings(), self.fitting_settings()
            if n == 1:  # the single-spectrum calls: no offset tables
                params = np.empty((lib.mdb_batch_n_lorentzians(batch, 0), 3), dtype=np.float64)
                peaks = np.empty((lib.mdb_batch_n_peaks(batch, 0), 3), dtype=np.int32)
                mse1 = C.c_double()
                raise_for_status(lib.mdb_batch_export(batch, None, None, None, C.byref(mse1), params.ctypes.data,
                                                      peaks.ctypes.data), _lib.last_error())
                return [Deconvolution(params, mse1.value, sm, se, fi, peaks)]
            # one bulk export for the whole batch; every Deconvolution holds views into the flat arrays
            tot_l, tot_p = C.c_size_t(), C.c_size_t()
            lib.mdb_batch_totals(batch, C.byref(tot_l), C.byref(tot_p))
            n_lor = np.empty(n, dtype=np.uint64)
            n_pk = np.empty(n, dtype=np.uint64)
            mse = np.empty(n, dtype=np.float64)
            params = np.empty((tot_l.value, 3), dtype=np.float64)
            peaks = np.empty((tot_p.value, 3), dtype=np.int32)
            raise_for_status(lib.mdb_batch_export(batch, None, n_lor.ctypes.data, n_pk.ctypes.data, mse.ctypes.data,
                                                  params.ctypes.data, peaks.ctypes.data), _lib.last_error())
            ol = np.concatenate(([0], np.cumsum(n_lor))).astype(np.int64)
            op = np.concatenate(([0], np.cumsum(n_pk))).astype(np.int64)
            out = [Deconvolution(params[ol[i]:ol[i + 1]], float(mse[i]), sm, se, fi, peaks[op[i]:op[i + 1]])
                   for i in range(n)]
            return out
        finally:
            if batch:
                lib.mdb_batch_free(batch)

    def optimize_settings(self, reference: Spectrum) -> float:
        """`Deconvoluter::optimize_settings` (deconvoluter.rs:761-825, bindings/deconvoluter.rs:167-172):
        810 deconvolutions of `reference` as one GPU batch; keeps the settings with the lowest MSE."""
        if not isinstance(reference, Spectrum):
            raise TypeError("expected a Spectrum")
        view = _lib.SpectrumView()
        view.chemical_shifts = reference.chemical_shifts.ctypes.data
        view.intensities = reference.intensities.ctypes.data
        view.len = reference.chemical_shifts.size
        view.signal_boundaries[0], view.signal_boundaries[1] = reference.signal_boundaries
        mse = C.c_double()
        st = _lib.load().mdb_deconvoluter_optimize_settings(self._h, C.byref(view), _lib.MDB_MEM_HOST, C.byref(mse))
        raise_for_status(st, _lib.last_error())
        return float(mse.value)

    def deconvolute_spectrum(self, spectrum: Spectrum) -> Deconvolution:
        return self._run([spectrum])[0]

    def par_deconvolute_spectrum(self, spectrum: Spectrum) -> Deconvolution:
        return self._run([spectrum])[0]

    def deconvolute_spectra(self, spectra) -> list:
        return self._run(list(spectra))

    def par_deconvolute_spectra(self, spectra) -> list:
        return self._run(list(spectra))
