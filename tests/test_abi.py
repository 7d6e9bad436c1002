"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol the header
declares, and the host logic that needs no GPU (settings validation, ignore-region merge,
Spectrum validation, loud failure without a device) behaves like the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from metabodecon_rust_b200 import Deconvoluter, Spectrum, _lib, exceptions

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "mdb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mdb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = header_functions()
    assert len(declared) >= 35
    bound = {name for name, _, _ in _lib.SIGNATURES}
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mdb200.h but not exported"
        assert name in bound, f"{name} declared in include/mdb200.h but not bound in _lib.SIGNATURES"
    assert bound <= set(declared)
    assert lib.mdb_abi_version() == _lib.ABI_VERSION == 2


def test_rust_sys_crate_declares_every_header_symbol():
    """rust/metabodecon-sys cannot be compiled here (no cargo/rustc); at least keep its extern
    block in step with the header: same symbol set, nothing missing, nothing invented."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "include", "mdb200.h")) as fh:
        header = fh.read()
    with open(os.path.join(root, "rust", "metabodecon-sys", "src", "lib.rs")) as fh:
        rust = fh.read()
    declared = set(re.findall(r"\b(mdb_[a-z0-9_]+)\s*\(", header))
    bound = set(re.findall(r"pub fn (mdb_[a-z0-9_]+)\s*\(", rust))
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))
    assert declared == {name for name, _, _ in _lib.SIGNATURES}


def test_struct_layouts_match_the_header():
    assert C.sizeof(_lib.Lorentzian3) == 24        # lorentzian.rs:138-145: three f64
    assert C.sizeof(_lib.SmoothingSettings) == 24
    assert C.sizeof(_lib.SelectionSettings) == 16
    assert C.sizeof(_lib.FittingSettings) == 16
    assert C.sizeof(_lib.SpectrumView) == 40


def test_default_settings():  # smoother.rs:58-65, selector.rs:59-66, fitter.rs:59-63
    d = Deconvoluter()
    assert d.smoothing_settings() == {"method": "MovingAverage", "iterations": 3, "windowSize": 3}
    assert d.selection_settings() == {"method": "NoiseScoreFilter", "scoringMethod": {"method": "MinimumSum"}, "threshold": 5.0}
    assert d.fitting_settings() == {"method": "Analytical", "iterations": 10}
    assert d.ignore_regions() is None


def test_settings_validation():  # deconvoluter.rs:919-1010
    d = Deconvoluter()
    for it, w in [(0, 3), (3, 0), (3, 1), (0, 0)]:
        with pytest.raises(exceptions.InvalidSmoothingSettings):
            d.set_moving_average_smoother(it, w)
    assert d.smoothing_settings()["iterations"] == 3  # unchanged after a rejected update
    for thr in (0.0, -1.0, float("inf"), float("nan")):
        with pytest.raises(exceptions.InvalidSelectionSettings):
            d.set_noise_score_selector(thr)
    with pytest.raises(exceptions.InvalidFittingSettings):
        d.set_analytical_fitter(0)
    d.set_moving_average_smoother(2, 2)
    d.set_identity_smoother()
    assert d.smoothing_settings() == {"method": "Identity"}
    d.set_detector_only()
    assert d.selection_settings() == {"method": "DetectorOnly"}
    d.set_noise_score_selector(6.4)
    d.set_analytical_fitter(15)
    assert d.fitting_settings()["iterations"] == 15
    with pytest.raises(ValueError):
        d.set_threads(1)
    d.set_threads(8)
    d.clear_threads()


def test_ignore_region_merge():  # deconvoluter.rs:438-472 and its tests :1012-1115
    d = Deconvoluter()
    for bad in [(1.0, 1.0), (float("nan"), 2.0), (1.0, float("inf")), (1.0, 1.0 + 1e-14)]:
        with pytest.raises(exceptions.InvalidIgnoreRegion):
            d.add_ignore_region(bad)
    assert d.ignore_regions() is None
    d.add_ignore_region((4.9, 4.7))
    assert d.ignore_regions() == [(4.7, 4.9)]
    d.add_ignore_region((1.0, 2.0))
    d.add_ignore_region((7.0, 8.0))
    assert d.ignore_regions() == [(1.0, 2.0), (4.7, 4.9), (7.0, 8.0)]
    d.add_ignore_region((4.8, 5.5))          # overlap -> merged
    assert d.ignore_regions() == [(1.0, 2.0), (4.7, 5.5), (7.0, 8.0)]
    d.add_ignore_region((2.0, 3.0))          # touching -> merged
    assert d.ignore_regions() == [(1.0, 3.0), (4.7, 5.5), (7.0, 8.0)]
    d.add_ignore_region((0.0, 10.0))         # swallows everything
    assert d.ignore_regions() == [(0.0, 10.0)]
    d.clear_ignore_regions()
    assert d.ignore_regions() is None


def test_spectrum_validation():  # spectrum.rs:179-200, 756-905
    x = np.linspace(0.0, 10.0, 101)
    y = np.ones(101)
    assert Spectrum(x, y, (9.0, 1.0)).signal_boundaries == (1.0, 9.0)           # re-ordered, increasing axis
    assert Spectrum(x[::-1], y, (1.0, 9.0)).signal_boundaries == (9.0, 1.0)     # decreasing axis
    with pytest.raises(exceptions.EmptyData):
        Spectrum(np.zeros(0), np.zeros(0), (1.0, 2.0))
    with pytest.raises(exceptions.DataLengthMismatch):
        Spectrum(x, y[:-1], (1.0, 9.0))
    bad = x.copy()
    bad[50] += 1e-3
    with pytest.raises(exceptions.NonUniformSpacing):
        Spectrum(bad, y, (1.0, 9.0))
    with pytest.raises(exceptions.NonUniformSpacing):
        Spectrum(np.zeros(101), y, (1.0, 9.0))
    ybad = y.copy()
    ybad[3] = np.nan
    with pytest.raises(exceptions.InvalidIntensities):
        Spectrum(x, ybad, (1.0, 9.0))
    for sb in [(1.0, 1.0), (-1.0, 9.0), (1.0, 11.0), (float("nan"), 2.0)]:
        with pytest.raises(exceptions.InvalidSignalBoundaries):
            Spectrum(x, y, sb)


def test_bruker_reader_matches_reference_checks(golden_dir):
    # macros/check_spectrum.rs:27-50 (check_blood_spectrum!): 2^17 points, range, nucleus, frequency
    sp = Spectrum.read_bruker(os.path.join(golden_dir, "bruker", "blood_01"), 10, 10, (-2.2, 11.8))
    assert len(sp) == 131072 and sp.nucleus == "1H"
    assert abs(sp.chemical_shifts[0] - 14.81146) < 1e-3 and abs(sp.chemical_shifts[-1] - (-5.2121)) < 1e-3
    assert abs(sp.frequency - 600.252821089118) < 1e-9
    assert sp.signal_boundaries == (11.8, -2.2)
    sim = Spectrum.read_bruker(os.path.join(golden_dir, "bruker", "sim_01"), 10, 10, (3.35, 3.55))
    assert len(sim) == 2048


def test_jcampdx_reader_matches_bruker_reader(golden_dir):
    """tests/parsing.rs:7-63 of the reference: the Bruker and JCAMP-DX forms of blood_01 agree
    (intensities exactly -- both are the same int32 samples; chemical shifts to 1e-3)."""
    from metabodecon_rust_b200.readers import read_bruker_arrays, read_jcampdx_arrays
    xb, yb, _ = read_bruker_arrays(os.path.join(golden_dir, "bruker", "blood_01"), 10, 10)
    xj, yj, meta = read_jcampdx_arrays(os.path.join(golden_dir, "jcampdx", "blood_01.dx"))
    assert xj.size == yj.size == 131072
    assert np.array_equal(yb, yj)
    assert np.max(np.abs(xb - xj)) < 1e-3
    assert xj[0] == 14.81146 and meta["nucleus"] == "1H" and abs(meta["frequency"] - 600.252821089118) < 1e-9


def test_jcampdx_asdf_forms_decode_to_the_same_ordinates(tmp_path):
    """The same ten ordinates written as AFFN, PAC, SQZ, DIF (with Y-checks) and DIFDUP
    (JCAMP-DX 5.01 section 5.9 example style) must decode identically."""
    from metabodecon_rust_b200.readers import read_jcampdx_arrays
    want = [1000.0, 1001.0, 1003.0, 1003.0, 1003.0, 1003.0, 998.0, 990.0, 990.0, 1200.0]
    header = ("##TITLE=t\n##JCAMP-DX=5.01\n##DATA TYPE=NMR SPECTRUM\n##DATA CLASS=XYDATA\n"
              "##.OBSERVE FREQUENCY=600.0\n##.OBSERVE NUCLEUS=^1H\n##XUNITS=PPM\n##YUNITS=ARBITRARY UNITS\n"
              "##XFACTOR=1\n##YFACTOR=1\n##FIRSTX=9\n##LASTX=0\n##NPOINTS=10\n##XYDATA=(X++(Y..Y))\n")
    tables = {
        "affn": "9 1000 1001 1003 1003 1003\n4 1003 998 990 990 1200\n",
        "pac": "9+1000+1001+1003+1003+1003\n4+1003+998+990+990+1200\n",
        "sqz": "9A000A001A003A003A003\n4A003I98I90I90A200\n",
        "dif": "9A000JK%%\n5A003%n\n3I98q%\n1I90K10\n0A200\n",
        "difdup": "9A000JK%T\n5A003%n\n3I98q%\n1I90K10\n0A200\n",
    }
    for name, table in tables.items():
        path = tmp_path / f"{name}.dx"
        path.write_text(header + table + "##END=\n")
        x, y, _ = read_jcampdx_arrays(str(path))
        assert y.tolist() == want, name
        assert x[0] == 9.0 and x[-1] == 0.0


@pytest.mark.skipif(_lib.load().mdb_device_count() > 0, reason="only meaningful without a GPU")
def test_compute_fails_loudly_without_a_device():
    x = np.linspace(10.0, 0.0, 100)
    sp = Spectrum(x, np.ones(100), (9.0, 1.0))
    with pytest.raises(exceptions.CudaError):
        Deconvoluter().deconvolute_spectrum(sp)
    from metabodecon_rust_b200 import Lorentzian
    with pytest.raises(exceptions.CudaError):
        Lorentzian.superposition_vec(x, [Lorentzian(1.0, 1.0, 0.0)])


def test_superposition_mode_switch_is_host_logic():
    # include/mdb200.h: process-wide, needs no device; FAST is the default, unknown values are refused
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import metabodecon_rust_b200 as m; from metabodecon_rust_b200 import _lib; "
            "print(m.superposition_mode()); m.set_superposition_mode('exact'); print(m.superposition_mode()); "
            "print(_lib.load().mdb_set_superposition_mode(9) != 0); print(m.superposition_mode())") % ROOT
    for env_value, first in ((None, "fast"), ("exact", "exact"), ("fast", "fast")):
        env = dict(os.environ)
        env.pop("MDB_SUPERPOSITION", None)
        if env_value:
            env["MDB_SUPERPOSITION"] = env_value
        out = subprocess.run([sys.executable, "-c", code], env=env, check=True, capture_output=True, text=True).stdout.split()
        assert out == [first, "exact", "True", "exact"], out


def test_superposition_mode_is_a_property_of_the_deconvoluter():
    """Host logic of VERDICT r1 item 7: the mode lives in the handle (pinned) or follows the process
    default (never pinned); clones inherit; a malformed MDB_SUPERPOSITION is reported, not ignored."""
    import subprocess
    import sys
    lib = _lib.load()
    a, b = Deconvoluter(), Deconvoluter()
    before = lib.mdb_superposition_mode()
    try:
        assert lib.mdb_set_superposition_mode(_lib.MDB_SUPERPOSITION_FAST) == 0
        assert a.superposition_mode() == "fast" and b.superposition_mode() == "fast"
        a.set_superposition_mode("exact")
        assert a.superposition_mode() == "exact" and b.superposition_mode() == "fast"
        assert lib.mdb_set_superposition_mode(_lib.MDB_SUPERPOSITION_EXACT) == 0
        assert b.superposition_mode() == "exact"  # never pinned: follows the default
        b.set_superposition_mode("fast")
        assert lib.mdb_set_superposition_mode(_lib.MDB_SUPERPOSITION_FAST) == 0
        assert a.superposition_mode() == "exact" and b.superposition_mode() == "fast"
        clone = C.c_void_p()
        assert lib.mdb_deconvoluter_clone(a._h, C.byref(clone)) == 0
        assert lib.mdb_deconvoluter_superposition_mode(clone) == _lib.MDB_SUPERPOSITION_EXACT
        lib.mdb_deconvoluter_free(clone)
        with pytest.raises(ValueError):
            a.set_superposition_mode("sloppy")
        assert lib.mdb_deconvoluter_set_superposition_mode(a._h, 9) != 0
        assert lib.mdb_deconvoluter_fit_arithmetic(a._h) == _lib.MDB_FIT_EXACT  # the product default
        assert lib.mdb_deconvoluter_set_fit_arithmetic(a._h, 9) != 0
    finally:
        lib.mdb_set_superposition_mode(before if before >= 0 else _lib.MDB_SUPERPOSITION_FAST)
    code = ("import sys; sys.path.insert(0, %r); from metabodecon_rust_b200 import _lib; lib = _lib.load(); "
            "print(lib.mdb_superposition_mode()); print(lib.mdb_superposition_vec(None, 0, None, 0, None, 0)); "
            "print(_lib.last_error())") % ROOT
    env = dict(os.environ, MDB_SUPERPOSITION="sloppy")
    out = subprocess.run([sys.executable, "-c", code], env=env, check=True, capture_output=True, text=True).stdout.splitlines()
    assert int(out[0]) < 0 and int(out[1]) == _lib.MDB_ERR_INVALID_ARGUMENT and "sloppy" in out[2], out


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "metabodecon_rust_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "mdb_oracle" not in text, f


def test_deconvolution_json_and_messagepack_layouts(tmp_path):
    """Serde-shaped output: camelCase JSON (serialized_deconvolution.rs:18-31) and
    the compact MessagePack form of rmp_serde::to_vec (bindings/deconvolution.rs:97-117) -- structs as
    arrays in field order, internally tagged enums with the tag first, Lorentzians as (sf, hw, maxp)."""
    import json
    import struct
    from metabodecon_rust_b200 import Deconvolution
    from metabodecon_rust_b200.exceptions import SerializationError

    def f64(v):
        return b"\xcb" + struct.pack(">d", v)

    params = np.array([[12.5 * 0.25, 0.25 * 0.25, 5.0]])     # Lorentzian::new(sfhw, hw2, maxp) of (sf 12.5, hw 0.25, maxp 5)
    dec = Deconvolution(params, 0.5, {"method": "MovingAverage", "iterations": 3, "windowSize": 3},
                        {"method": "NoiseScoreFilter", "scoringMethod": {"method": "MinimumSum"}, "threshold": 5.0},
                        {"method": "Analytical", "iterations": 10})
    dec.write_bin(str(tmp_path / "d.bin"))
    want = (b"\x95" + b"\x93\xadMovingAverage\x03\x03" + b"\x93\xb0NoiseScoreFilter\x91\xaaMinimumSum" + f64(5.0)
            + b"\x92\xaaAnalytical\x0a" + f64(0.5) + b"\x91\x93" + f64(12.5) + f64(0.25) + f64(5.0))
    assert (tmp_path / "d.bin").read_bytes() == want
    back = Deconvolution.read_bin(str(tmp_path / "d.bin"))
    assert back.mse == 0.5 and np.array_equal(back.parameters, params)
    assert back.smoothing_settings == dec.smoothing_settings and back.selection_settings == dec.selection_settings
    assert back.fitting_settings == dec.fitting_settings
    # the named (map) form of rmp_serde::to_vec_named reads too
    import msgpack
    named = {"smoothingSettings": {"method": "Identity"}, "selectionSettings": {"method": "DetectorOnly"},
             "fittingSettings": {"method": "Analytical", "iterations": 2}, "mse": 1.25,
             "lorentzians": [{"sf": 2.0, "hw": 0.5, "maxp": -1.0}]}
    (tmp_path / "n.bin").write_bytes(msgpack.packb(named))
    got = Deconvolution.read_bin(str(tmp_path / "n.bin"))
    assert got.smoothing_settings == {"method": "Identity"} and np.array_equal(got.parameters, [[1.0, 0.25, -1.0]])
    # JSON: same content, camelCase keys, untransformed Lorentzians
    dec.write_json(str(tmp_path / "d.json"))
    js = json.loads((tmp_path / "d.json").read_text())
    assert set(js) == {"smoothingSettings", "selectionSettings", "fittingSettings", "mse", "lorentzians"}
    assert js["lorentzians"] == [{"sf": 12.5, "hw": 0.25, "maxp": 5.0}]
    assert np.array_equal(Deconvolution.read_json(str(tmp_path / "d.json")).parameters, params)
    (tmp_path / "bad.bin").write_bytes(b"\x93\x01\x02")
    with pytest.raises(SerializationError):
        Deconvolution.read_bin(str(tmp_path / "bad.bin"))


def test_spectrum_json_and_messagepack_round_trip(tmp_path):
    """Spectrum storage in the reference's serde shape (spectrum/serialized_spectrum.rs:17-62): the axis
    travels as (first, last, size) and is rebuilt as first + i * step."""
    import json
    n = 64
    x = 10.0 - np.arange(n) * (10.0 / (n - 1))
    y = np.sin(np.arange(n) * 0.3) * 100.0
    sp = Spectrum(x, y, (1.0, 9.0))
    sp.nucleus, sp.frequency = "13C", 150.9
    sp.reference_compound = {"chemical_shift": 10.0, "index": 0, "name": "TMS"}
    sp.write_json(str(tmp_path / "s.json"))
    js = json.loads((tmp_path / "s.json").read_text())
    assert set(js) == {"spectrumBoundaries", "signalBoundaries", "size", "nucleus", "frequency", "referenceCompound", "intensities"}
    assert js["size"] == n and js["nucleus"] == "13C" and js["referenceCompound"] == {"chemicalShift": 10.0, "index": 0, "name": "TMS"}
    sp.write_bin(str(tmp_path / "s.bin"))
    for back in (Spectrum.read_json(str(tmp_path / "s.json")), Spectrum.read_bin(str(tmp_path / "s.bin"))):
        assert np.array_equal(back.intensities, y) and back.signal_boundaries == sp.signal_boundaries
        want_x = x[0] + np.arange(n) * ((x[-1] - x[0]) / (n - 1.0))       # serialized_spectrum.rs:54-57
        assert np.array_equal(back.chemical_shifts, want_x)
        assert back.nucleus == "13C" and back.frequency == 150.9 and back.reference_compound["name"] == "TMS"
    (tmp_path / "bad.json").write_text('{"size": 3}')
    with pytest.raises(exceptions.SerializationError):
        Spectrum.read_json(str(tmp_path / "bad.json"))


def test_integration_doc_names_every_header_symbol_and_nothing_else():
    """INTEGRATION.md is the map from the reference's interfaces to the C ABI: every function the header
    declares appears there, and every mdb_* function name used there exists in the header."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "include", "mdb200.h")) as fh:
        header = fh.read()
    with open(os.path.join(root, "INTEGRATION.md")) as fh:
        doc = fh.read()
    declared = set(re.findall(r"\b(mdb_[a-z0-9_]+)\s*\(", header))
    for name in sorted(declared):
        assert re.search(r"\b%s\b" % name, doc), f"{name} is declared in include/mdb200.h but INTEGRATION.md does not mention it"
    for name in sorted(set(re.findall(r"\bmdb_[a-z0-9_]*[a-z0-9]\b", doc))):
        assert re.search(r"\b%s\b" % name, header), f"INTEGRATION.md mentions {name}, which include/mdb200.h does not know"
