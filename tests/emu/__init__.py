"""ctypes wrapper of the host emulation of the product's kernel source (TEST INFRASTRUCTURE)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libwgkemu.so")
PMAP = {0: "gamma_hbv", 1: "cfa", 2: "cfs", 4: "p_rivrgh", 7: "p_swoutf", 8: "p_evaredex", 9: "p_netrad", 10: "p_ptc_hum",
        11: "p_ptc_ari", 12: "p_pet_mxdy", 13: "p_mcwh", 15: "p_snowfz", 16: "p_snowmt", 17: "p_degday", 18: "p_gradnt",
        21: "p_pcrit", 22: "p_gwoutf", 25: "p_prec"}


def build():
    srcs = [os.path.join(HERE, "wgk_emu.cpp"), os.path.join(HERE, "cuda_shim.h"),
            os.path.join(HERE, "..", "..", "watergap2_b200", "csrc", "wgk_kernels.cuh"),
            os.path.join(HERE, "..", "..", "watergap2_b200", "csrc", "wgk_fields.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", LIB, srcs[0]])
    return LIB


class Emu:
    def __init__(self, ncell, rout_order, downstream, form="bands"):
        L = ctypes.CDLL(build())
        vp, ci = ctypes.c_void_p, ctypes.c_int
        L.emu_create.restype = vp
        L.emu_create.argtypes = [ci, vp, vp]
        L.emu_set.argtypes = [vp, ctypes.c_char_p, vp]
        L.emu_get.argtypes = [vp, ctypes.c_char_p, vp]
        L.emu_set_forcing.argtypes = [vp] * 5
        L.emu_day.argtypes = [vp] + [ci] * 5
        self.L, self.ncell = L, ncell
        ro = np.ascontiguousarray(rout_order, np.int32)
        dn = np.ascontiguousarray(downstream, np.int32)
        self.e = L.emu_create(ncell, ro.ctypes.data, dn.ctypes.data)
        L.emu_set_form.argtypes = [vp, ci]
        L.emu_set_form(self.e, {"bands": 0, "cells": 1, "bands2": 2, "fused": 3}[form])

    def set(self, name, arr, dtype):
        a = np.ascontiguousarray(np.asarray(arr).astype(dtype))
        assert self.L.emu_set(self.e, name.encode(), a.ctypes.data) == 0, name

    def get(self, name, like):
        out = np.empty_like(like)
        assert self.L.emu_get(self.e, name.encode(), out.ctypes.data) == 0, name
        return out

    def load(self, fields, oracle):
        for k, v in fields.items():
            if k.startswith("_"):
                continue
            if k == "params":
                for kk, nm in PMAP.items():
                    self.set(nm, v[kk], np.float64)
            elif oracle.has(k) and k not in ("routing_cell", "downstream_cell", "glo_lake", "glo_res", "rout_order"):
                self.set(k, v, oracle.field(k).dtype)

    def set_forcing(self, f):
        arrs = [np.ascontiguousarray(f[k], np.float32) for k in ("P", "T", "SW", "LW")]
        self.L.emu_set_forcing(self.e, *[x.ctypes.data for x in arrs])

    def day(self, day, month, dom, slot, tail_level0=-1):
        self.L.emu_day(self.e, day, month, dom, slot, tail_level0)
