"""Harness-grade spectrum readers.

These feed the hot path with the same arrays the reference's readers would produce; they are
host-side I/O, not part of the accelerated path.

* Bruker TopSpin: reference `metabodecon/src/spectrum/formats/bruker.rs:260-287` (x-axis
  synthesis), `:447-490` (`1r` decode), metadata keys `:164-166` (acqus) and `:192-196` (procs).
* JCAMP-DX: reference `metabodecon/src/spectrum/formats/jcampdx.rs` (see `read_jcampdx`).
"""
from __future__ import annotations

import os
import re

import numpy as np

_ACQUS = {
    "width": (re.compile(r"(?m)^##\$SW=\s*(\d+(\.\d+)?)"), float),
    "frequency": (re.compile(r"(?m)^##\$SFO1=\s*(\d+(\.\d+)?)"), float),
    "nucleus": (re.compile(r"(?m)^##\$NUC1=\s*<(\w+)"), str),
}
_PROCS = {
    "maximum": (re.compile(r"(?m)^##\$OFFSET=\s*(\d+(\.\d+)?)"), float),
    "exponent": (re.compile(r"(?m)^##\$NC_proc=\s*(-?\d+)"), int),
    "endian": (re.compile(r"(?m)^##\$BYTORDP=\s*(\d)"), int),
    "data_type": (re.compile(r"(?m)^##\$DTYPP=\s*(\d)"), int),
    "data_size": (re.compile(r"(?m)^##\$SI=\s*(\d+)"), int),
}


class MissingMetadataError(ValueError):
    pass


def _extract(table, text, path):
    out = {}
    for key, (rx, conv) in table.items():
        m = rx.search(text)
        if m is None:
            raise MissingMetadataError(f"missing metadata key '{key}' in {path}")
        out[key] = conv(m.group(1))
    return out


def read_bruker_arrays(path: str, experiment: int, processing: int):
    """Returns (chemical_shifts, intensities, meta) for `path/experiment/pdata/processing`."""
    acqus_path = os.path.join(path, str(experiment), "acqus")
    procs_path = os.path.join(path, str(experiment), "pdata", str(processing), "procs")
    one_r_path = os.path.join(path, str(experiment), "pdata", str(processing), "1r")
    with open(acqus_path, "r", errors="replace") as fh:
        acqus = _extract(_ACQUS, fh.read(), acqus_path)
    with open(procs_path, "r", errors="replace") as fh:
        procs = _extract(_PROCS, fh.read(), procs_path)
    size = procs["data_size"]
    # bruker.rs:278-280   x_i = maximum - i * width / (size - 1)   (product first, then divide)
    i = np.arange(size, dtype=np.float64)
    chemical_shifts = procs["maximum"] - i * acqus["width"] / (float(size) - 1.0)
    order = "<" if procs["endian"] == 0 else ">"
    if procs["data_type"] == 0:  # int32, scaled by 2^NC_proc (bruker.rs:459-475)
        raw = np.fromfile(one_r_path, dtype=np.dtype(order + "i4"), count=size)
        if raw.size != size:
            raise ValueError(f"{one_r_path}: expected {size} values, found {raw.size}")
        intensities = raw.astype(np.float64) * (2.0 ** procs["exponent"])
    else:  # f64 (bruker.rs:476-487)
        raw = np.fromfile(one_r_path, dtype=np.dtype(order + "f8"), count=size)
        if raw.size != size:
            raise ValueError(f"{one_r_path}: expected {size} values, found {raw.size}")
        intensities = raw.astype(np.float64)
    meta = {"nucleus": acqus["nucleus"], "frequency": acqus["frequency"]}
    return chemical_shifts, intensities, meta
