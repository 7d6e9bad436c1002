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


from .exceptions import MalformedData, MalformedMetadata, MissingData, MissingMetadata  # noqa: E402

MissingMetadataError = MissingMetadata  # the reference's spectrum::error::Kind::MissingMetadata


def _extract(table, text, path):
    out = {}
    for key, (rx, conv) in table.items():
        m = rx.search(text)
        if m is None:
            raise MissingMetadataError(f"missing metadata key '{key}' in {path}")
        try:
            out[key] = conv(m.group(1))
        except ValueError as err:
            raise MalformedMetadata(f"malformed metadata key '{key}' in {path}: {err}") from err
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
            raise MissingData(f"{one_r_path}: expected {size} values, found {raw.size}")
        intensities = raw.astype(np.float64) * (2.0 ** procs["exponent"])
    else:  # f64 (bruker.rs:476-487)
        raw = np.fromfile(one_r_path, dtype=np.dtype(order + "f8"), count=size)
        if raw.size != size:
            raise MissingData(f"{one_r_path}: expected {size} values, found {raw.size}")
        intensities = raw.astype(np.float64)
    meta = {"nucleus": acqus["nucleus"], "frequency": acqus["frequency"]}
    return chemical_shifts, intensities, meta


def read_bruker(path: str, experiment: int, processing: int, signal_boundaries):
    """`Bruker::read_spectrum` (bruker.rs:260-287) -> Spectrum."""
    from .spectrum import Spectrum
    return Spectrum.read_bruker(path, experiment, processing, signal_boundaries)


# ---------------------------------------------------------------------------------------------
# JCAMP-DX (XYDATA, `(X++(Y..Y))`, AFFN or ASDF = PAC / SQZ / DIF / DUP), jcampdx.rs:555-590.
# The reference decodes ASDF by regex rewriting (jcampdx.rs:925-969); this is a token decoder with
# the same outcome: the trailing DIF ordinate of a line and the Y-check that opens the next line
# are the same sample, which is kept once.
# ---------------------------------------------------------------------------------------------
_DX_HEADER = {
    "version": (re.compile(r"(?m)^(##JCAMP(\s*|_|-)DX=\s*)(?P<v>\d+(\.\d+)?)"), float),
    "type": (re.compile(r"(?m)^(##DATA(\s|_)TYPE=\s*)(?P<v>\w+\s\w+)"), str),
    "format": (re.compile(r"(?m)^(##DATA(\s|_)CLASS=\s*)(?P<v>\w+(\s\w+)?)"), str),
    "frequency": (re.compile(r"(?m)^(##\.OBSERVE(\s|_)FREQUENCY=\s*)(?P<v>\d+(\.\d+)?)"), float),
    "nucleus": (re.compile(r"(?m)^(##\.OBSERVE(\s|_)NUCLEUS=\s*)(?P<v>\^\w+)"), str),
}
_DX_REF_INDEX = re.compile(r"(?m)^(##\.SHIFT(\s|_)REFERENCE=[^,]*,[^,]*,\s*)(?P<v>\d+)")
_DX_REF_SHIFT = re.compile(r"(?m)^(##\.SHIFT(\s|_)REFERENCE=[^,]*,[^,]*,[^,]*,\s*)(?P<v>\d+(\.\d+)?)")
_DX_SOLVENT_SHIFT = re.compile(r"(?m)^(##\.SOLVENT(\s|_)REFERENCE=\s*)(?P<v>\d+(\.\d+))")
_DX_XY = {
    "xunits": (re.compile(r"(?m)^(##XUNITS=\s*)(?P<v>\w+)"), str),
    "factor": (re.compile(r"(?m)^(##YFACTOR=\s*)(?P<v>\d+(\.\d+)?)"), float),
    "first": (re.compile(r"(?m)^(##FIRSTX=\s*)(?P<v>\d+(\.\d+)?)"), float),
    "last": (re.compile(r"(?m)^(##LASTX=\s*)(?P<v>\d+(\.\d+)?)"), float),
    "data_size": (re.compile(r"(?m)^(##NPOINTS=\s*)(?P<v>\d+(\.\d+)?)"), float),
}
_DX_DATA = re.compile(r"(?m)^(##XYDATA=\s*\(X\+\+\([RY]\.\.[RY]\)\)(.*)?)(?P<v>[^#$]*)")

_SQZ = {c: i for i, c in enumerate("@ABCDEFGHI")}
_SQZ.update({c: -(i + 1) for i, c in enumerate("abcdefghi")})
_DIF = {c: i for i, c in enumerate("%JKLMNOPQR")}
_DIF.update({c: -(i + 1) for i, c in enumerate("jklmnopqr")})
_DUP = {c: i + 1 for i, c in enumerate("STUVWXYZs")}


def _dx_capture(table, text, path):
    out = {}
    for key, (rx, conv) in table.items():
        m = rx.search(text)
        if m is None:
            raise MissingMetadataError(f"missing metadata key '{key}' in {path}")
        out[key] = conv(m.group("v"))
    return out


def _decode_xydata(data: str, path: str) -> np.ndarray:
    """Ordinates of an `(X++(Y..Y))` table in AFFN or ASDF form (abscissa column dropped)."""
    if not re.search(r"[@%A-Za-z+-]", data):  # plain AFFN (jcampdx.rs:892-918): skip the abscissa
        out = []
        for line in data.splitlines():
            for tok in line.split()[1:]:
                try:
                    out.append(float(tok))
                except ValueError as err:
                    raise MalformedData(f"{path}: malformed value '{tok}'") from err
        return np.asarray(out, dtype=np.float64)
    values = []
    prev_line_ended_in_dif = False
    token = re.compile(r"[@A-Ia-i%J-Rj-rS-Zs+-]\d*|\d+(?:\.\d+)?")
    for line in data.splitlines():
        fields = token.findall(line)
        if not fields:
            continue
        line_vals = []
        last_kind = "abs"  # what produced the last ordinate: an absolute value or a DIF
        last_dif = 0
        for tok in fields[1:]:  # fields[0] is the abscissa of the X++ column
            head = tok[0]
            if head in _SQZ:
                d = _SQZ[head]
                mag = int(str(abs(d)) + tok[1:])
                line_vals.append(-mag if d < 0 else mag)
                last_kind = "abs"
            elif head in _DIF:
                d = _DIF[head]
                mag = int(str(abs(d)) + tok[1:])
                last_dif = -mag if d < 0 else mag
                if not line_vals:
                    raise MalformedData(f"{path}: DIF token without a preceding ordinate")
                line_vals.append(line_vals[-1] + last_dif)
                last_kind = "dif"
            elif head in _DUP:
                count = int(str(_DUP[head]) + tok[1:])
                if not line_vals:
                    raise MalformedData(f"{path}: DUP token without a preceding ordinate")
                for _ in range(count - 1):
                    line_vals.append(line_vals[-1] + last_dif if last_kind == "dif" else line_vals[-1])
            else:  # PAC (+/- prefixed) or a bare integer
                try:
                    line_vals.append(int(tok))
                except ValueError as err:
                    raise MalformedData(f"{path}: malformed value '{tok}'") from err
                last_kind = "abs"
        if prev_line_ended_in_dif and line_vals:
            line_vals = line_vals[1:]  # Y-check: repeats the last ordinate of the previous line
        values.extend(line_vals)
        prev_line_ended_in_dif = last_kind == "dif"
    return np.asarray(values, dtype=np.float64)


def read_jcampdx_arrays(path: str):
    """Returns (chemical_shifts, intensities, meta) of a JCAMP-DX 5/6 XYDATA NMR spectrum
    (jcampdx.rs:555-590: header :666-710, block :726-760, axis :565-576)."""
    with open(path, "r", errors="replace") as fh:
        dx = fh.read()
    header = _dx_capture(_DX_HEADER, dx, path)
    if int(header["version"]) not in (5, 6) or header["type"].upper() != "NMR SPECTRUM":
        raise ValueError(f"{path}: unsupported JCAMP-DX file")
    if header["format"].upper() != "XYDATA":
        raise ValueError(f"{path}: only XYDATA blocks are supported by this harness reader")
    block = _dx_capture(_DX_XY, dx, path)
    m = _DX_DATA.search(dx)
    if m is None or not m.group("v").strip():
        raise MissingData(f"{path}: missing data table")
    xunits = block["xunits"].upper()
    if xunits not in ("HZ", "PPM"):
        raise ValueError(f"{path}: unsupported x units {xunits}")
    conversion = 1.0 / header["frequency"] if xunits == "HZ" else 1.0
    size = int(block["data_size"])
    step = (block["last"] - block["first"]) * conversion / (float(size) - 1.0)
    mi, ms = _DX_REF_INDEX.search(dx), _DX_REF_SHIFT.search(dx)
    if mi is not None and ms is not None:      # .SHIFT REFERENCE: (shift, index - 1)
        offset = float(ms.group("v")) - float(int(mi.group("v")) - 1) * step
    else:
        msol = _DX_SOLVENT_SHIFT.search(dx)
        offset = float(msol.group("v")) if msol is not None else block["first"] * conversion
    i = np.arange(size, dtype=np.float64)
    chemical_shifts = offset + i * step
    intensities = _decode_xydata(m.group("v").strip(), path) * block["factor"]
    if intensities.size != size:
        raise MalformedData(f"{path}: expected {size} ordinates, decoded {intensities.size}")
    meta = {"nucleus": header["nucleus"].lstrip("^"), "frequency": header["frequency"]}
    return chemical_shifts, intensities, meta
