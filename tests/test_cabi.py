"""CPU: the C-ABI shared library loads and exports every symbol include/wgk.h declares; the
field table is consistent with the oracle's; without a GPU the product fails loudly."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def libwgk():
    import watergap2_b200 as wg
    if not os.path.exists(wg.LIB_PATH):
        wg.build()
    return wg.lib()


def test_every_declared_symbol_is_exported(libwgk):
    hdr = open(os.path.join(ROOT, "include", "wgk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(wgk_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    for n in names:
        assert hasattr(libwgk, n), f"{n} declared in include/wgk.h but not exported by libwgk.so"


def test_field_table_matches_oracle_names(libwgk, oracle_lib):
    import watergap2_b200 as wg
    from oracle import wg_init
    o = oracle_lib.Oracle(4)
    for name in wg_init.STATIC_FIELDS + wg_init.STATE_FIELDS + wg_init.FLUX_FIELDS:
        assert libwgk.wgk_field_id(name.encode()) >= 0, name
        assert o.has(name), name
    assert libwgk.wgk_field_id(b"no_such_field") < 0
    # dtype / size agreement for a few representative fields
    nm, dt, cnt = ctypes.c_char_p(), ctypes.c_char_p(), ctypes.c_int64()
    for name, want_dt, want_cnt in [("snow_bands", b"f64", 101 * 7), ("elevation", b"i16", 101 * 7), ("rgmax", b"i16", 7),
                                    ("params", b"f64", 26 * 7), ("lct_ddf", b"f64", 18), ("lai_days", b"i32", 7)]:
        f = libwgk.wgk_field_id(name.encode())
        libwgk.wgk_field_info(f, ctypes.byref(nm), ctypes.byref(dt), ctypes.byref(cnt), 7)
        assert dt.value == want_dt and cnt.value == want_cnt, name


def test_no_cpu_fallback():
    """a Model cannot be created without a CUDA device: the product never computes on the CPU"""
    import torch
    import watergap2_b200 as wg
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(wg.WgkError):
        wg.Model(16)


def test_product_does_not_import_oracle():
    """nothing under watergap2_b200/ (python or C++/CUDA) may reference oracle/"""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "watergap2_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|oracle/|wg_oracle|libwgoracle", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_header_is_plain_c(tmp_path):
    """include/wgk.h is the drop-in boundary: it must compile as C99 (no C++ or torch types) for cgo / JNI / ctypes / Fortran bindings"""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "wgk.h"\nint main(void) { return wgk_field_id("snow") == 12345; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])


def test_cell_class_key_with_coldness_bins():
    """the sort key of Model.set_topology(cell_class=...): low 4 bits = water-body class, high 4 bits = the opt-in coldness bin
    (coldest first) from the 0.5 degree row and the mean elevation; off by default"""
    import numpy as np
    import watergap2_b200 as wg
    f = {"loc_lake": np.array([0., 1., 0., 0.]), "loc_wetland": np.array([0., 2., 0., 0.]), "lake_area": np.array([0., 0., 3., 0.]),
         "reservoir_area": np.zeros(4), "glo_wetland": np.zeros(4), "arid": np.array([0, 0, 0, 1]),
         "row": np.array([20, 100, 180, 340], np.int16), "elevation": np.array([[4000] + [0] * 100, [100] + [0] * 100, [0] * 101, [10] * 101], np.int16)}
    assert wg.cell_classes(f, cold_bins=0).tolist() == [0, 3, 4, 8]
    k = wg.cell_classes(f, cold_bins=16)
    assert (k & 15).tolist() == [0, 3, 4, 8]
    lat = 90.25 - 0.5 * f["row"].astype(float)
    t = 27.0 - 0.55 * np.abs(lat) - 0.0045 * f["elevation"][:, 0]
    assert (k >> 4).tolist() == np.clip(np.floor((t + 30.0) / 60.0 * 16), 0, 15).astype(int).tolist()
    assert (k >> 4)[0] < (k >> 4)[2]  # the high mountain cell near the pole sorts before the tropical lowland cell
    assert wg.coldness_bin(np.array([180]), np.array([0]), 4).tolist() == [3]
