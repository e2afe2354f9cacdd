"""ctypes front-end of the CPU oracle (oracle/wg_oracle.c) and reader of the reference dump
container written by oracle/ref_harness.cpp.  TEST INFRASTRUCTURE: importable only from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libwgoracle.so")

DT = {"f64": np.float64, "f32": np.float32, "i32": np.int32, "i16": np.int16, "i8": np.int8, "u16": np.uint16}


def build(force=False):
    src = os.path.join(HERE, "wg_oracle.c")
    hdr = os.path.join(HERE, "wg_oracle.h")
    if (not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-std=gnu11",
                           src, "-o", LIB, "-lm"])
    return LIB


_lib = None


def _bind(_lib):
    if True:
        _lib.wgo_create.restype = ctypes.c_void_p
        _lib.wgo_create.argtypes = [ctypes.c_int]
        _lib.wgo_destroy.argtypes = [ctypes.c_void_p]
        _lib.wgo_field.restype = ctypes.c_void_p
        _lib.wgo_field.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p),
                                   ctypes.POINTER(ctypes.c_int64)]
        _lib.wgo_set_restart.argtypes = [ctypes.c_void_p, ctypes.c_int]
        _lib.wgo_set_subtract_use.argtypes = [ctypes.c_void_p, ctypes.c_int]
        for f in ("wgo_vertical_day", "wgo_routing_day"):
            getattr(_lib, f).argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _lib.wgo_update_land_area_frac.argtypes = [ctypes.c_void_p]
        _lib.wgo_step_days.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4
        _lib.wgo_total_storage_km3.restype = ctypes.c_double
        _lib.wgo_total_storage_km3.argtypes = [ctypes.c_void_p]
        _lib.wgo_rout_prepare.argtypes = [ctypes.c_void_p]
        _lib.wgo_rout_prepare.restype = ctypes.c_int
    return _lib


def lib():
    global _lib
    if _lib is None:
        _lib = _bind(ctypes.CDLL(build()))
    return _lib


class Oracle:
    """Owns a wgo_ctx; fields are exposed as numpy views of the C arrays (zero copy)."""

    def __init__(self, ncell, lib_path=None):
        """lib_path: alternative build of wg_oracle.c (e.g. with FMA contraction) for sensitivity studies"""
        self.ncell = ncell
        self._L = lib() if lib_path is None else _bind(ctypes.CDLL(lib_path))
        self._c = self._L.wgo_create(ncell)
        self._views = {}

    def close(self):
        if self._c:
            self._L.wgo_destroy(self._c)
            self._c = None
            self._views = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def field(self, name):
        if name in self._views:
            return self._views[name]
        dt = ctypes.c_char_p()
        cnt = ctypes.c_int64()
        p = self._L.wgo_field(self._c, name.encode(), ctypes.byref(dt), ctypes.byref(cnt))
        if not p:
            raise KeyError(name)
        npdt = DT[dt.value.decode()]
        buf = (ctypes.c_char * (cnt.value * np.dtype(npdt).itemsize)).from_address(p)
        v = np.frombuffer(buf, dtype=npdt)
        self._views[name] = v
        return v

    def has(self, name):
        try:
            self.field(name)
            return True
        except KeyError:
            return False

    def set(self, name, arr):
        v = self.field(name)
        a = np.asarray(arr).ravel()
        assert a.size == v.size, (name, a.size, v.size)
        v[:] = a.astype(v.dtype, copy=False)

    def load_records(self, recs, day=0, skip=()):
        """Copy every record of `day` whose name is a field of the oracle."""
        n = 0
        for (name, d), a in recs.items():
            if d == day and name not in skip and self.has(name):
                self.set(name, a)
                n += 1
        return n

    def set_forcing_month(self, f):
        self.set("prec31", f["P"])
        self.set("temp31", f["T"])
        self.set("sw31", f["SW"])
        self.set("lw31", f["LW"])

    def vertical_day(self, day, month, dom):
        self._L.wgo_vertical_day(self._c, day, month, dom)

    def routing_day(self, day, month, dom):
        self._L.wgo_routing_day(self._c, day, month, dom)

    def update_land_area_frac(self):
        self._L.wgo_update_land_area_frac(self._c)

    def step_day(self, day, month, dom):
        self.vertical_day(day, month, dom)
        self.routing_day(day, month, dom)
        self.update_land_area_frac()

    def total_storage_km3(self):
        return self._L.wgo_total_storage_km3(self._c)


def read_dump(path, days=None, names=None):
    """Parse a WGD1 container -> {(name, day): ndarray}."""
    out = {}
    with open(path, "rb") as fh:
        while True:
            hdr = fh.read(32 + 4 + 8 + 8)
            if len(hdr) < 52:
                break
            name = hdr[:32].split(b"\0", 1)[0].decode()
            day = int(np.frombuffer(hdr[32:36], "<i4")[0])
            dt = hdr[36:44].split(b"\0", 1)[0].decode()
            cnt = int(np.frombuffer(hdr[44:52], "<i8")[0])
            nbytes = cnt * np.dtype(DT[dt]).itemsize
            if (days is not None and day not in days) or (names is not None and name not in names):
                fh.seek(nbytes, 1)
                continue
            out[(name, day)] = np.frombuffer(fh.read(nbytes), DT[dt]).copy()
    return out


NDAYS = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]


def calendar(simday, start_month=1):
    """simulated day 1.. -> (day-of-year 1..365, month 0..11, day_in_month 1..31) for a run
    starting on the first day of `start_month` (365-day years, integrateWGHM.cpp:100)."""
    doy0 = sum(NDAYS[: start_month - 1])
    d = (doy0 + simday - 1) % 365
    m = 0
    while d >= NDAYS[m]:
        d -= NDAYS[m]
        m += 1
    return (doy0 + simday - 1) % 365 + 1, m, d + 1


class Topology(ctypes.Structure):
    _fields_ = [("ncell", ctypes.c_int), ("ncol", ctypes.c_int), ("nrow", ctypes.c_int),
                ("flowdir", ctypes.c_void_p), ("row", ctypes.c_void_p), ("col", ctypes.c_void_p),
                ("gcrc", ctypes.c_void_p), ("ldd_2", ctypes.c_void_p), ("ldd", ctypes.c_void_p),
                ("inflow9", ctypes.c_void_p), ("flow_acc", ctypes.c_void_p), ("basins", ctypes.c_void_p),
                ("basins2", ctypes.c_void_p), ("cells_to_outlet", ctypes.c_void_p),
                ("outflow_cell", ctypes.c_void_p), ("rout_order", ctypes.c_void_p),
                ("neighbour8", ctypes.c_void_p), ("nlevels", ctypes.c_int32), ("nbasins", ctypes.c_int32),
                ("nbasins2", ctypes.c_int32)]


def rout_prepare(flowdir, row, col, gcrc_colmajor, ncol=720, nrow=360):
    """Restated prepare_routing_files topology. gcrc_colmajor: int32 [ncol][nrow]."""
    ng = len(flowdir)
    a = {"flowdir": np.ascontiguousarray(flowdir, np.int16), "row": np.ascontiguousarray(row, np.int16),
         "col": np.ascontiguousarray(col, np.int16), "gcrc": np.ascontiguousarray(gcrc_colmajor, np.int32),
         "ldd_2": np.zeros(ng, np.int8), "ldd": np.zeros(ng, np.int8), "inflow9": np.zeros((ng, 9), np.int32),
         "flow_acc": np.zeros(ng, np.int16), "basins": np.zeros(ng, np.uint16), "basins2": np.zeros(ng, np.uint16),
         "cells_to_outlet": np.zeros(ng, np.uint16), "outflow_cell": np.zeros(ng, np.int32),
         "rout_order": np.zeros(ng, np.int32), "neighbour8": np.zeros((ng, 8), np.int32)}
    t = Topology()
    t.ncell, t.ncol, t.nrow = ng, ncol, nrow
    for k, v in a.items():
        setattr(t, k, v.ctypes.data)
    rc = lib().wgo_rout_prepare(ctypes.byref(t))
    if rc != 0:
        raise RuntimeError("routing order is not finished (cycle in the flow directions)")
    a["nlevels"], a["nbasins"], a["nbasins2"] = t.nlevels, t.nbasins, t.nbasins2
    return a


def river_geometry(altitude, meandering, outflow_cell, ldd, row, col, nrow=360):
    L = lib()
    cd = np.zeros((9, nrow), np.float32)
    L.wgo_cell_distances.argtypes = [ctypes.c_int, ctypes.c_void_p]
    L.wgo_cell_distances(nrow, cd.ctypes.data)
    ng = len(altitude)
    slope = np.zeros(ng, np.float32)
    length = np.zeros(ng, np.float32)
    L.wgo_river_slope_length.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 9
    args = [np.ascontiguousarray(altitude, np.float32), np.ascontiguousarray(meandering, np.float32),
            np.ascontiguousarray(outflow_cell, np.int32), np.ascontiguousarray(ldd, np.int8),
            np.ascontiguousarray(row, np.int16), np.ascontiguousarray(col, np.int16)]
    L.wgo_river_slope_length(ng, nrow, cd.ctypes.data, *[x.ctypes.data for x in args], slope.ctypes.data,
                             length.ctypes.data)
    return cd, slope, length


def reservoir_prepare(resarea, mean_outflow, mean_outflow12, outflow_cell):
    L = lib()
    ng = len(resarea)
    alloc = np.zeros((ng, 5), np.float32)
    sm = np.zeros(ng, np.int8)
    L.wgo_reservoir_prepare.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 6
    args = [np.ascontiguousarray(resarea, np.float32), np.ascontiguousarray(mean_outflow, np.float32),
            np.ascontiguousarray(mean_outflow12, np.float32), np.ascontiguousarray(outflow_cell, np.int32)]
    L.wgo_reservoir_prepare(ng, *[x.ctypes.data for x in args], alloc.ctypes.data, sm.ctypes.data)
    return alloc, sm
