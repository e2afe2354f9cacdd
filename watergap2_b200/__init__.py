"""watergap2_b200 — B200-native WaterGAP2 daily hot path.

Python here is plumbing only: a ctypes binding of the C ABI in ``include/wgk.h`` (the same
entry points the C++ look-alike classes in ``csrc/host`` call), used by the tests, the
benchmark and the multi-GPU launcher.  The compute path is ``libwgk.so`` (hand-written
sm_100a kernels, ``csrc/wgk_kernels.cuh``); if the library is missing or no GPU is present
every compute entry point raises — there is no CPU or PyTorch fallback.
"""
import ctypes
import os

import numpy as np

__all__ = ["Model", "WgkError", "lib", "LIB_PATH", "FIELD_DTYPES", "build", "cell_classes"]

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WGK_LIB") or os.path.join(HERE, "libwgk.so")  # WGK_LIB: development override (kernel variants)
NBAND, NLCT, NPARAM = 101, 18, 26
FIELD_DTYPES = {"f64": np.float64, "f32": np.float32, "i32": np.int32, "i16": np.int16, "i8": np.int8}


class WgkError(RuntimeError):
    pass


def cell_classes(fields, cold_bins=None):
    """Sort key per cell for Model.set_topology(cell_class=...): cells of equal key are stored next to each other inside a routing
    level (results do not depend on it).  Low 4 bits: which water-body code a cell needs (bit 0 local lake, bit 1 local wetland,
    bit 2 global lake / reservoir / global wetland, bit 3 arid).  High 4 bits (opt-in, WGK_COLD_BINS=1..16): a coldness bin from
    latitude and mean elevation (coldest first), so that the cells of a warp / CTA tend to be either all in the 100-band snow loop
    or all out of it.  Measured on B200 without gain (one member, ms per simulated year: off 18.4, bins first 19.5, water-body
    class first and bins inside it 18.6): warps of mixed water-body classes cost more than warps of mixed snow regimes."""
    f = fields
    z = lambda k: np.asarray(f[k]).ravel() > 0
    key = (z("loc_lake") * 1 + z("loc_wetland") * 2 + (z("lake_area") | z("reservoir_area") | z("glo_wetland")) * 4
           + (np.asarray(f["arid"]).ravel() == 1) * 8).astype(np.int32)
    if cold_bins is None:
        cold_bins = int(os.environ.get("WGK_COLD_BINS", "0"))
    if cold_bins > 0 and "row" in f and "elevation" in f:
        key += 16 * coldness_bin(np.asarray(f["row"]).ravel(), np.asarray(f["elevation"]).reshape(key.size, -1)[:, 0], min(cold_bins, 16))
    return key.astype(np.uint8)


def coldness_bin(row, elevation_m, nbins=16):
    """0 (coldest) .. nbins-1 from the 0.5 degree row and the mean elevation: a climatological mean temperature of
    27 - 0.55 |lat| - 0.0045 m (any monotone proxy serves - it only orders cells); the host layer uses the same (wg_model.cpp)"""
    lat = 90.25 - 0.5 * row.astype(np.float64)
    t = 27.0 - 0.55 * np.abs(lat) - 0.0045 * elevation_m.astype(np.float64)
    return np.clip(np.floor((t + 30.0) / 60.0 * nbins), 0, nbins - 1).astype(np.int32)


class _Options(ctypes.Structure):
    _fields_ = [("restart", ctypes.c_int), ("tail_threshold", ctypes.c_int), ("use_graph", ctypes.c_int), ("subtract_use", ctypes.c_int)]


_lib = None


def build(verbose=False):
    """Compile csrc/wgk_api.cu into watergap2_b200/libwgk.so for sm_100a (in-tree)."""
    import subprocess

    src = os.path.join(HERE, "csrc", "wgk_api.cu")
    deps = [src, os.path.join(HERE, "csrc", "wgk_kernels.cuh"), os.path.join(HERE, "csrc", "wgk_fields.h"),
            os.path.join(HERE, "..", "include", "wgk.h")]
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "--cudart", "static", "-o", LIB_PATH, src]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return LIB_PATH


def lib():
    """Load libwgk.so (raises if it was not built: the CUDA library IS the product)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WgkError(f"{LIB_PATH} not found - run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cp = ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p
    L.wgk_create.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci, ctypes.POINTER(_Options)]
    L.wgk_destroy.argtypes = [vp]
    L.wgk_destroy.restype = None
    L.wgk_last_error.argtypes = [vp]
    L.wgk_last_error.restype = cp
    L.wgk_synchronize.argtypes = [vp]
    L.wgk_get_stream.argtypes = [vp]
    L.wgk_get_stream.restype = vp
    L.wgk_set_stream.argtypes = [vp, vp]
    L.wgk_set_topology.argtypes = [vp, vp, vp]
    L.wgk_set_cell_classes.argtypes = [vp, vp]
    L.wgk_set_forcing_unf.argtypes = [vp, ci, ci, ci, vp, vp, vp, vp, ci]
    L.wgk_month_begin.argtypes = [vp]
    L.wgk_get_day_state.argtypes = [vp, ci, vp]
    L.wgk_state_vector.argtypes = [vp, ci, ci, vp, ci, vp, vp]
    L.wgk_enkf_update.argtypes = [vp, ci, vp, ci, vp, vp, vp]
    L.wgk_num_levels.argtypes = [vp]
    L.wgk_get_levels.argtypes = [vp, vp]
    L.wgk_field_id.argtypes = [cp]
    L.wgk_field_info.argtypes = [ci, ctypes.POINTER(cp), ctypes.POINTER(cp), ctypes.POINTER(ctypes.c_int64), ci]
    L.wgk_set_field.argtypes = [vp, ci, ci, vp, ctypes.c_size_t]
    L.wgk_get_field.argtypes = [vp, ci, ci, vp, ctypes.c_size_t]
    L.wgk_set_member_pset.argtypes = [vp, ci, ci]
    L.wgk_device_ptr.argtypes = [vp, ci, ci]
    L.wgk_device_ptr.restype = vp
    L.wgk_cell_stride.argtypes = [vp]
    L.wgk_cell_stride.restype = ctypes.c_int64
    L.wgk_member_stride.argtypes = [vp]
    L.wgk_member_stride.restype = ctypes.c_int64
    L.wgk_layout.argtypes = [vp]
    L.wgk_schedule.argtypes = [vp]
    L.wgk_get_device_order.argtypes = [vp, vp]
    L.wgk_forcing_reserve.argtypes = [vp, ci, ci]
    L.wgk_set_forcing.argtypes = [vp, ci, ci, ci, vp, vp, vp, vp, ci]
    L.wgk_vertical_day.argtypes = [vp, ci, ci, ci, ci]
    L.wgk_routing_day.argtypes = [vp, ci, ci, ci]
    L.wgk_update_land_area_frac.argtypes = [vp]
    L.wgk_step_days.argtypes = [vp, ci, ci, ci, ci, ci]
    L.wgk_total_storage_km3.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_double)]
    L.wgk_record_cells.argtypes = [vp, vp, ci, ci]
    L.wgk_get_record.argtypes = [vp, ci, vp, ci]
    L.wgk_record_rewind.argtypes = [vp]
    L.wgk_profile_day.argtypes = [vp, ci, ci, ci, ci, ctypes.POINTER(ctypes.c_float)]
    L.wgk_ensemble_moments.argtypes = [vp, ci, vp, ci, ctypes.POINTER(vp), ctypes.POINTER(vp)]
    L.wgk_moments_finish.argtypes = [vp, ci, vp, vp]
    L.wgk_copy_index.argtypes = [vp, ci, ci, ci]
    L.wgk_fill_field.argtypes = [vp, ci, ci, ctypes.c_double]
    L.wgk_profile_schedule.argtypes = [vp, ci, ci, ci, ci, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ci)]
    L.wgk_stamps.argtypes = [vp, ci, vp]
    L.wgk_fp64_peak.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    L.wgk_kernel_launches.argtypes = [vp]
    L.wgk_kernel_launches.restype = ctypes.c_int64
    _lib = L
    return L


HOST_LIB_PATH = os.path.join(HERE, "libwghost.so")
_host = None


def host_lib():
    """libwghost.so: the C++ drop-in layer (class look-alikes, routing-file builder, checkpoint codecs, integrate_wghm driver)"""
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise WgkError(f"{HOST_LIB_PATH} not found - run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib()  # libwgk.so first (libwghost links it)
        H = ctypes.CDLL(HOST_LIB_PATH)
        H.wg_host_create_context.restype = ctypes.c_void_p
        H.wg_host_create_context.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
        H.wg_host_integrate.restype = ctypes.c_long
        H.wg_host_integrate.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_char_p, ctypes.c_size_t]
        _host = H
    return _host


class Model:
    """One wgk context: `nmember` members of a `ncell` grid on one GPU."""

    @classmethod
    def from_config(cls, config_file, ncell, device=0):
        """a ready-to-step single-member model built by the product's HOST LAYER (libwghost.so) from a reference-format
        configuration: routing files (rout_prepare), init sequence of integrate_wghm_, statics / parameters / start state on the
        device, 31 forcing slots reserved (wg_host_create_context)"""
        err = ctypes.create_string_buffer(1024)
        ptr = host_lib().wg_host_create_context(os.fsencode(config_file), ncell, device, err, 1024)
        if not ptr:
            raise WgkError(f"wg_host_create_context failed: {err.value.decode()}")
        m = cls.__new__(cls)
        m.ncell, m.nmember, m.npset, m.device = ncell, 1, 1, device
        m.subtract_use = 0
        m._L = lib()
        m._c = ctypes.c_void_p(ptr)
        m._ids = {}
        return m

    def __init__(self, ncell, nmember=1, npset=1, device=0, restart=0, tail_threshold=0, use_graph=1, subtract_use=0):
        self.ncell, self.nmember, self.npset, self.device = ncell, nmember, npset, device
        self.subtract_use = subtract_use
        self._L = lib()
        self._c = ctypes.c_void_p()
        opt = _Options(restart, tail_threshold, use_graph, subtract_use)
        rc = self._L.wgk_create(ctypes.byref(self._c), device, ncell, nmember, npset, ctypes.byref(opt))
        if rc != 0:
            msg = self._L.wgk_last_error(self._c).decode() if self._c else "no CUDA device (wgk has no CPU fallback)"
            if self._c:
                self._L.wgk_destroy(self._c)
            self._c = None
            raise WgkError(f"wgk_create failed ({rc}): {msg}")
        self._ids = {}

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise WgkError(f"wgk error {rc}: {self._L.wgk_last_error(self._c).decode()}")

    def close(self):
        if getattr(self, "_c", None):
            self._L.wgk_destroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def field_id(self, name):
        if name not in self._ids:
            f = self._L.wgk_field_id(name.encode())
            if f < 0:
                raise KeyError(name)
            self._ids[name] = f
        return self._ids[name]

    def has_field(self, name):
        try:
            self.field_id(name)
            return True
        except KeyError:
            return False

    def field_info(self, name):
        f = self.field_id(name)
        nm, dt, cnt = ctypes.c_char_p(), ctypes.c_char_p(), ctypes.c_int64()
        scope = self._L.wgk_field_info(f, ctypes.byref(nm), ctypes.byref(dt), ctypes.byref(cnt), self.ncell)
        return FIELD_DTYPES[dt.value.decode()], cnt.value, scope

    # -- topology / fields --------------------------------------------------------------------
    def set_topology(self, rout_order, downstream_cell, cell_class=None):
        ro = np.ascontiguousarray(rout_order, np.int32)
        dc = np.ascontiguousarray(downstream_cell, np.int32)
        assert ro.size == self.ncell and dc.size == self.ncell
        if cell_class is not None:
            cc = np.ascontiguousarray(cell_class, np.uint8)
            assert cc.size == self.ncell
            self._ck(self._L.wgk_set_cell_classes(self._c, cc.ctypes.data))
        self._ck(self._L.wgk_set_topology(self._c, ro.ctypes.data, dc.ctypes.data))

    @property
    def nlevels(self):
        return self._L.wgk_num_levels(self._c)

    def levels(self):
        out = np.zeros(self.ncell, np.int32)
        self._ck(self._L.wgk_get_levels(self._c, out.ctypes.data))
        return out

    def device_order(self):
        out = np.zeros(self.ncell, np.int32)
        self._ck(self._L.wgk_get_device_order(self._c, out.ctypes.data))
        return out

    def set(self, name, arr, index=0):
        dt, cnt, _ = self.field_info(name)
        a = np.ascontiguousarray(np.asarray(arr).astype(dt, copy=False)).ravel()
        if a.size != cnt:
            raise WgkError(f"field {name}: expected {cnt} elements, got {a.size}")
        self._ck(self._L.wgk_set_field(self._c, self.field_id(name), index, a.ctypes.data, a.nbytes))

    def get(self, name, index=0):
        dt, cnt, _ = self.field_info(name)
        out = np.empty(cnt, dt)
        self._ck(self._L.wgk_get_field(self._c, self.field_id(name), index, out.ctypes.data, out.nbytes))
        return out

    def set_member_pset(self, member, pset):
        self._ck(self._L.wgk_set_member_pset(self._c, member, pset))

    def load(self, fields, member=None, pset=None, only=None):
        """Set every entry of dict `fields` that names a wgk field (others are ignored).
        Static fields go to index 0, pset fields to `pset` (default all), member fields to
        `member` (default all)."""
        n = 0
        for name, arr in fields.items():
            if name.startswith("_") or (only is not None and name not in only) or not self.has_field(name):
                continue
            if name.startswith("wu_") and not getattr(self, "subtract_use", 0):
                continue  # the water-use arrays exist only with subtract_use > 0
            _, _, scope = self.field_info(name)
            if scope == 1:
                idx = range(self.npset) if pset is None else [pset]
            elif scope == 2:
                idx = range(self.nmember) if member is None else [member]
            else:
                idx = [0]
            for i in idx:
                self.set(name, arr, i)
            n += 1
        return n

    # -- forcing --------------------------------------------------------------------------------
    def forcing_reserve(self, nslots, per_member=False):
        self._ck(self._L.wgk_forcing_reserve(self._c, nslots, int(per_member)))

    def set_forcing(self, slot0, ndays, prec, temp, sw, lw, member=-1):
        arrs = [np.ascontiguousarray(x, np.float32) for x in (prec, temp, sw, lw)]
        stride = arrs[0].size // self.ncell
        for x in arrs:
            assert x.size == self.ncell * stride
        self._ck(self._L.wgk_set_forcing(self._c, slot0, ndays, member, *[x.ctypes.data for x in arrs], stride))
        self._keep = arrs  # the copies are asynchronous: keep the host buffers alive

    def set_forcing_unf(self, slot0, ndays, prec, temp, sw, lw, member=-1, stride=31):
        """the four buffers are the raw bytes of the reference's big-endian [ncell][stride] float32 UNF0 files"""
        arrs = [np.frombuffer(x, np.uint8) if not isinstance(x, np.ndarray) else x.view(np.uint8).ravel() for x in (prec, temp, sw, lw)]
        for x in arrs:
            assert x.size == self.ncell * stride * 4
        self._ck(self._L.wgk_set_forcing_unf(self._c, slot0, ndays, member, *[x.ctypes.data for x in arrs], stride))
        self._keep = arrs

    # -- hot path ---------------------------------------------------------------------------------
    def vertical_day(self, day, month, dom, slot):
        self._ck(self._L.wgk_vertical_day(self._c, day, month, dom, slot))

    def routing_day(self, day, month, dom):
        self._ck(self._L.wgk_routing_day(self._c, day, month, dom))

    def update_land_area_frac(self):
        self._ck(self._L.wgk_update_land_area_frac(self._c))

    def step_days(self, day, month, dom, slot0, ndays):
        self._ck(self._L.wgk_step_days(self._c, day, month, dom, slot0, ndays))

    def synchronize(self):
        self._ck(self._L.wgk_synchronize(self._c))

    @property
    def stream(self):
        return self._L.wgk_get_stream(self._c)

    def set_stream(self, cuda_stream):
        self._ck(self._L.wgk_set_stream(self._c, ctypes.c_void_p(cuda_stream)))

    # -- EnKF state bridge ------------------------------------------------------------------------
    def month_begin(self):
        self._ck(self._L.wgk_month_begin(self._c))

    def day_state(self, member=0):
        """[7, ncell]: the day's WghmStateFile entry of the routing compartments (mm over the continental area), reference order"""
        out = np.empty((7, self.ncell), np.float64)
        self._ck(self._L.wgk_get_day_state(self._c, member, out.ctypes.data))
        return out

    def state_vector(self, cells, kind="month", member=0, mean_field=None):
        """[ncells, 10] state vector of extract_sub_ (mm over the continental area) for `cells` (0-based)"""
        c = np.ascontiguousarray(cells, np.int32)
        out = np.empty((c.size, 10), np.float64)
        mf = None if mean_field is None else np.ascontiguousarray(mean_field, np.float64)
        self._ck(self._L.wgk_state_vector(self._c, member, {"month": 0, "lastday": 1}[kind], c.ctypes.data, c.size,
                                          None if mf is None else mf.ctypes.data, out.ctypes.data))
        return out

    def enkf_update(self, cells, field, prediction, mean_field, member=0):
        c = np.ascontiguousarray(cells, np.int32)
        a = [np.ascontiguousarray(x, np.float64) for x in (field, prediction, mean_field)]
        assert all(x.size == c.size * 10 for x in a)
        self._ck(self._L.wgk_enkf_update(self._c, member, c.ctypes.data, c.size, *[x.ctypes.data for x in a]))

    # -- ensemble statistics / device-side set-up ---------------------------------------------------
    def ensemble_moments(self, kind="month", cells=None):
        """-> (device pointer of sum, device pointer of sumsq, ncells): [ncells, 10] f64 each, contiguous
        (sumsq directly behind sum), over the members of this context"""
        c = None if cells is None else np.ascontiguousarray(cells, np.int32)
        ps, pq = ctypes.c_void_p(), ctypes.c_void_p()
        self._ck(self._L.wgk_ensemble_moments(self._c, {"month": 0, "lastday": 1}[kind], None if c is None else c.ctypes.data,
                                              0 if c is None else c.size, ctypes.byref(ps), ctypes.byref(pq)))
        return ps.value, pq.value, (self.ncell if c is None else c.size)

    def moments_finish(self, nmember_total, ncells):
        mean = np.empty((ncells, 10), np.float64)
        var = np.empty((ncells, 10), np.float64)
        self._ck(self._L.wgk_moments_finish(self._c, nmember_total, mean.ctypes.data, var.ctypes.data))
        return mean, var

    def copy_member(self, src, dst):
        self._ck(self._L.wgk_copy_index(self._c, 2, src, dst))

    def copy_pset(self, src, dst):
        self._ck(self._L.wgk_copy_index(self._c, 1, src, dst))

    def fill(self, name, value, index=0):
        self._ck(self._L.wgk_fill_field(self._c, self.field_id(name), index, float(value)))

    # -- diagnostics ------------------------------------------------------------------------------
    def total_storage_km3(self, member=0):
        out = ctypes.c_double()
        self._ck(self._L.wgk_total_storage_km3(self._c, member, ctypes.byref(out)))
        return out.value

    def record_cells(self, cells, max_days):
        c = np.ascontiguousarray(cells, np.int32)
        self._ck(self._L.wgk_record_cells(self._c, c.ctypes.data, c.size, max_days))
        self._nrec = c.size

    def record_rewind(self):
        self._ck(self._L.wgk_record_rewind(self._c))

    def get_record(self, ndays, member=0):
        out = np.empty((ndays, self._nrec), np.float64)
        self._ck(self._L.wgk_get_record(self._c, member, out.ctypes.data, ndays))
        return out

    def profile_day(self, day, month, dom, slot):
        """-> dict of phase times in ms for one simulated day run with plain launches"""
        ms = (ctypes.c_float * 6)()
        self._ck(self._L.wgk_profile_day(self._c, day, month, dom, slot, ms))
        return dict(zip(("vertical", "route_local", "route_levels", "route_tail", "route_post", "day"), [float(x) for x in ms]))

    def profile_schedule(self, day, month, dom, slot):
        """-> {class: (ms, launches)} of one simulated day of the schedule step_days uses, every launch timed on its own"""
        ms, n = (ctypes.c_float * 4)(), (ctypes.c_int * 4)()
        self._ck(self._L.wgk_profile_schedule(self._c, day, month, dom, slot, ms, n))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(("vertical", "river_level", "tail", "other"))}

    def stamps(self, enable=True, read=False):
        out = np.zeros((2, 2, 512), np.uint64) if read else None
        self._ck(self._L.wgk_stamps(self._c, int(enable), None if out is None else out.ctypes.data))
        return out

    def fp64_peak_tflops(self):
        v = ctypes.c_double()
        self._ck(self._L.wgk_fp64_peak(self._c, ctypes.byref(v)))
        return v.value

    @property
    def kernel_launches(self):
        return self._L.wgk_kernel_launches(self._c)

    def device_ptr(self, name, member=0):
        return self._L.wgk_device_ptr(self._c, self.field_id(name), member)

    @property
    def cell_stride(self):
        return self._L.wgk_cell_stride(self._c)

    @property
    def member_stride(self):
        return self._L.wgk_member_stride(self._c)

    @property
    def schedule(self):
        """bit flags of wgk_schedule: 1 whole-day kernels (else the (day, level) wavefront), 2 fused level tasks, 4 cell owner"""
        return self._L.wgk_schedule(self._c)

    @property
    def layout(self):
        """'cells' ([member][cell], a warp = 32 cells of one member) or 'members' ([cell][member], a warp = 32 members of one cell)"""
        return "members" if self._L.wgk_layout(self._c) == 1 else "cells"
