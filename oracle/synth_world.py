#!/usr/bin/env python3
"""Synthetic world generator for the WaterGAP2 hot path (TEST INFRASTRUCTURE, not product).

Writes every input the reference needs under the canonical option vector of SURVEY.md
§8(d) / Appendix D, in the reference's own on-disk formats:

* UNF grids, big-endian, suffix = element type (UNF0 f32, UNF1 i8, UNF2 i16, UNF4 i32),
  multi-channel files stored cell-major ``[cell][K]``
  (reference: grid_io_adapters.h:56-127, grid.h:471-474);
* text tables ``LCT_22.DAT`` (daily.cpp:1925-1959), ``LAI_22.DAT`` (lai.cpp:110-148),
  ``OPTIONS.DAT`` (option.cpp:173-620, 36 "Value:" lines), ``OUTPUT_OPTIONS.DAT``
  (option.cpp:933-958, 142 lines), ``ROUTING.DAT`` (routing.cpp:207-239),
  ``STATIONS.DAT`` (calib_basins.cpp:89-94), ``config.txt`` (configFile.cpp:20-231);
* the parameter JSON (calib_param.cpp:140-171, 26 arrays of ng values).

The land mask lives on the 720x360 half-degree raster the reference hard-codes (geo.h:16),
``ng`` cells numbered row-major from the north-west.  Everything is a deterministic
function of (ng, seed).  The flow-direction map is hydrologically conditioned (priority
flood from the coast), so drainage basins are large and flow paths long, as in DDM30.

Usage:  python oracle/synth_world.py OUTDIR [--ng 67420] [--seed 20240607]
                                     [--years 1901 1901] [--months 1 12]
"""
import argparse
import heapq
import json
import os

import numpy as np

NCOL, NROW = 720, 360
EARTH_R = 6371.211

# D8 neighbour offsets (dcol, drow) and the Arc flow-direction code that points that way
# (rout_prepare.cpp:79-113: 1=E 2=SE 4=S 8=SW 16=W 32=NW 64=N 128=NE; row grows southward).
D8 = [(1, 0, 1), (1, 1, 2), (0, 1, 4), (-1, 1, 8), (-1, 0, 16), (-1, -1, 32), (0, -1, 64), (1, -1, 128)]


def be(a, dt):
    return np.ascontiguousarray(a).astype(dt).astype(np.dtype(dt).newbyteorder(">"))


def write_unf(path, a, dt):
    be(a, dt).tofile(path)


def smooth_noise(rng, shape, scale):
    """Periodic-in-x smooth gaussian field via FFT low-pass; unit variance."""
    nr, nc = shape
    f = rng.standard_normal(shape)
    F = np.fft.rfft2(f)
    ky = np.fft.fftfreq(nr)[:, None]
    kx = np.fft.rfftfreq(nc)[None, :]
    k2 = (ky * nr / NROW) ** 2 + (kx * nc / NCOL) ** 2
    F *= np.exp(-k2 * (scale ** 2) * 20.0)
    g = np.fft.irfft2(F, s=shape)
    g -= g.mean()
    g /= g.std() + 1e-30
    return g


class World:
    pass


def build_world(ng=67420, seed=20240607):
    rng = np.random.default_rng(seed)
    w = World()
    w.ng = ng
    w.seed = seed

    # ---------------- land mask --------------------------------------------------
    field = smooth_noise(rng, (NROW, NCOL), 22.0) + 0.45 * smooth_noise(rng, (NROW, NCOL), 6.0)
    lat = 90.0 - (np.arange(NROW) + 0.5) * 0.5
    rows_ok = (np.arange(NROW) >= 20) & (np.arange(NROW) < 340)
    score = np.where(rows_ok[:, None], field, -np.inf)
    if ng < 20000:
        # mini worlds: keep one compact region so that basins are still non-trivial
        cy, cx = 110, 400
        yy, xx = np.mgrid[0:NROW, 0:NCOL]
        score = score - 0.004 * ((yy - cy) ** 2 + (xx - cx) ** 2) / max(1.0, ng / 4000.0)
    flat = score.ravel()
    order = np.argsort(-flat, kind="stable")
    land = np.zeros(NROW * NCOL, bool)
    land[order[:ng]] = True
    land = land.reshape(NROW, NCOL)
    assert land.sum() == ng
    rr, cc = np.nonzero(land)  # row-major scan: north->south, west->east
    w.row = (rr + 1).astype(np.int16)  # 1-based
    w.col = (cc + 1).astype(np.int16)
    gcrc = np.zeros((NROW, NCOL), np.int32)
    gcrc[rr, cc] = np.arange(1, ng + 1)
    w.gcrc = gcrc  # [row][col]; file order is [col][row]
    w.lat = lat[rr]

    # ---------------- cell area per row (km2) -------------------------------------
    phi1 = np.deg2rad(90.0 - np.arange(NROW) * 0.5)
    phi2 = np.deg2rad(90.0 - (np.arange(NROW) + 1) * 0.5)
    w.area_row = (EARTH_R ** 2 * np.deg2rad(0.5) * np.abs(np.sin(phi1) - np.sin(phi2))).astype(np.float32)
    area = w.area_row[rr].astype(np.float64)

    # ---------------- coast / continental fraction ---------------------------------
    coast = np.zeros(ng, bool)
    for dc, dr, _ in D8:
        r2 = rr + dr
        c2 = (cc + dc) % NCOL
        ok = (r2 >= 0) & (r2 < NROW)
        nb_land = np.zeros(ng, bool)
        nb_land[ok] = land[r2[ok], c2[ok]]
        coast |= ~nb_land
    w.coast = coast
    contfreq = np.full(ng, 100.0, np.float32)
    contfreq[coast] = rng.uniform(20.0, 100.0, coast.sum()).astype(np.float32)
    w.contfreq = contfreq

    # ---------------- DEM and hydrologically conditioned D8 ------------------------
    dem = 600.0 + 500.0 * smooth_noise(rng, (NROW, NCOL), 9.0) + 180.0 * smooth_noise(rng, (NROW, NCOL), 2.5)
    # distance-to-coast tilt so that continents have interior highlands
    dist = np.where(land, 1e9, 0.0)
    for _ in range(60):
        m = dist.copy()
        for dc, dr, _c in D8:
            sh = np.roll(np.roll(dist, -dr, axis=0), -dc, axis=1) + (1.0 if dc == 0 or dr == 0 else 1.4142)
            m = np.minimum(m, sh)
        if np.array_equal(m, dist):
            break
        dist = m
    dem = np.maximum(dem + 28.0 * np.minimum(dist, 60.0), 1.0)
    elev = dem[rr, cc]

    # endorheic seeds: a few deep interior local minima (inland sinks, Arc code -1)
    n_sinks = max(1, ng // 1500)
    interior = np.nonzero((dist[rr, cc] > 6))[0]
    sink_cells = []
    if interior.size:
        cand = interior[np.argsort(elev[interior] - 14.0 * dist[rr, cc][interior])]
        taken = np.zeros(ng, bool)
        for n in cand:
            if len(sink_cells) >= n_sinks:
                break
            # keep sinks at least ~12 cells apart
            if sink_cells:
                d2 = (rr[sink_cells] - rr[n]) ** 2 + np.minimum(abs(cc[sink_cells] - cc[n]), NCOL - abs(cc[sink_cells] - cc[n])) ** 2
                if d2.min() < 144:
                    continue
            sink_cells.append(int(n))
            taken[n] = True
    sink_cells = np.array(sink_cells, int)

    flowdir = np.zeros(ng, np.int16)  # 0 = outlet to the ocean
    down = np.full(ng, -1, np.int64)
    done = np.zeros(ng, bool)
    heap = []
    jitter = rng.uniform(0, 1e-3, ng)
    for n in np.nonzero(coast)[0]:
        heapq.heappush(heap, (float(elev[n] + jitter[n]), int(n)))
        done[n] = True
        flowdir[n] = 0
    for n in sink_cells:
        if not done[n]:
            heapq.heappush(heap, (float(elev[n] * 0.5), int(n)))
            done[n] = True
            flowdir[n] = -1
    if not heap:  # no coast (cannot happen with an ocean around) - fall back to lowest cell
        n = int(np.argmin(elev))
        heapq.heappush(heap, (float(elev[n]), n))
        done[n] = True
    idx_of = gcrc - 1
    flood_order = []
    while heap:
        h, n = heapq.heappop(heap)
        flood_order.append(n)
        r0, c0 = int(rr[n]), int(cc[n])
        for dc, dr, _code in D8:
            r2 = r0 + dr
            if r2 < 0 or r2 >= NROW:
                continue
            c2 = (c0 + dc) % NCOL
            m = idx_of[r2, c2]
            if m < 0 or done[m]:
                continue
            done[m] = True
            down[m] = n
            # m drains to n: direction from m to n is (-dc, -dr)
            for dc2, dr2, code2 in D8:
                if dc2 == -dc and dr2 == -dr:
                    flowdir[m] = code2
                    break
            heapq.heappush(heap, (max(h, float(elev[m] + jitter[m])) + 1e-6, int(m)))
    assert done.all()
    w.flowdir = flowdir
    w.down = down
    w.altitude = elev.astype(np.float32)
    # flow accumulation (cells), for bankfull flow / reservoir siting
    acc = np.ones(ng, np.int64)
    for n in reversed(flood_order):
        if down[n] >= 0:
            acc[down[n]] += acc[n]
    w.acc = acc

    # ---------------- static per-cell grids -----------------------------------------
    w.meander = rng.uniform(1.0, 1.6, ng).astype(np.float32)
    sigma = rng.uniform(5.0, 600.0, ng)
    q = (np.arange(1, 101) - 0.5) / 100.0
    # inverse normal CDF (Acklam-free: use erfinv via numpy polynomial approx)
    z = np.sqrt(2.0) * _erfinv(2.0 * q - 1.0)
    bands = np.maximum(elev[:, None] + sigma[:, None] * z[None, :], 0.0)
    er = np.zeros((ng, 101), np.int16)
    er[:, 0] = np.clip(np.rint(elev), 0, 8000).astype(np.int16)
    er[:, 1:] = np.clip(np.rint(np.sort(bands, axis=1)), 0, 8800).astype(np.int16)
    w.elev_range = er

    abslat = np.abs(w.lat)
    lc = np.empty(ng, np.int8)
    u = rng.uniform(0, 1, ng)
    lc[:] = 9
    lc[abslat < 12] = np.where(u[abslat < 12] < 0.7, 1, 11)
    m = (abslat >= 12) & (abslat < 30)
    lc[m] = np.choose((u[m] * 4).astype(int), [7, 10, 16, 11])
    m = (abslat >= 30) & (abslat < 50)
    lc[m] = np.choose((u[m] * 5).astype(int), [4, 5, 10, 13, 14])
    m = (abslat >= 50) & (abslat < 66)
    lc[m] = np.choose((u[m] * 4).astype(int), [3, 4, 6, 17])
    m = (abslat >= 66) & (abslat < 75)
    lc[m] = np.where(u[m] < 0.8, 17, 8)
    lc[abslat >= 75] = 15
    lc[rng.uniform(0, 1, ng) < 0.02] = 2
    w.landcover = lc

    bu = np.zeros(ng, np.float32)
    mb = rng.uniform(0, 1, ng) < 0.05
    bu[mb] = rng.uniform(0, 0.3, mb.sum()).astype(np.float32)
    w.builtup = bu

    arid_score = smooth_noise(rng, (NROW, NCOL), 12.0)[rr, cc] + 1.2 * np.exp(-((abslat - 25.0) / 12.0) ** 2) - 0.5
    w.arid = (arid_score > np.quantile(arid_score, 0.65)).astype(np.int16)

    tawc = rng.uniform(40.0, 250.0, ng).astype(np.float32)
    tawc[lc == 15] = -9999.0
    w.tawc = tawc
    w.slope_class = rng.integers(10, 70, ng).astype(np.int8)
    tex = rng.integers(10, 31, ng).astype(np.int8)
    ut = rng.uniform(0, 1, ng)
    tex[ut < 0.02] = 1
    tex[(ut >= 0.02) & (ut < 0.04)] = 2
    tex[(ut >= 0.04) & (ut < 0.05)] = -1
    w.texture = tex
    pg = np.zeros(ng, np.int8)
    mp = abslat > 55
    pg[mp] = np.clip((abslat[mp] - 55.0) * 4.0 + rng.uniform(-10, 10, mp.sum()), 0, 100).astype(np.int8)
    w.permaglac = pg
    w.aq_factor = rng.integers(0, 101, ng).astype(np.int8)
    w.gw_factor_corr = np.full(ng, -99.0, np.float32)
    w.roughness = rng.uniform(0.03, 0.07, ng).astype(np.float32)
    w.bankfull = (5.0 * acc.astype(np.float64) ** 0.8).astype(np.float32)

    # ---------------- surface water bodies --------------------------------------------
    def sparse_pct(p_nonzero, hi):
        a = np.zeros(ng, np.float64)
        m_ = rng.uniform(0, 1, ng) < p_nonzero
        a[m_] = hi * rng.beta(1.2, 4.0, m_.sum())
        return a

    loclak = sparse_pct(0.15, 30.0)
    locwet = sparse_pct(0.15, 30.0)
    glowet = sparse_pct(0.08, 30.0)
    glolak = np.zeros(ng)
    lakarea = np.zeros(ng)
    resarea = np.zeros(ng)
    glores = np.zeros(ng)  # % of cell (G_RES_<year>)
    reglake = np.zeros(ng)
    reg_status = np.zeros(ng, np.int8)
    locres = np.zeros(ng)

    n_lake = max(2, int(round(600 * ng / 67420)))
    n_res = max(2, int(round(900 * ng / 67420)))
    big = np.argsort(-acc + rng.uniform(0, 0.5, ng))
    pool = big[: max(4 * (n_lake + n_res), 16)]
    pool = pool[rng.permutation(pool.size)]
    lake_cells = pool[:n_lake]
    res_cells = pool[n_lake:n_lake + n_res]
    up_of = {}
    for n in range(ng):
        d = down[n]
        if d >= 0 and (d not in up_of or acc[n] > acc[up_of[d]]):
            up_of[d] = n
    for n in lake_cells:
        f = rng.uniform(5.0, 40.0)
        glolak[n] = f
        a = f / 100.0 * area[n]
        if rng.uniform() < 0.3 and n in up_of:
            m_ = up_of[n]
            f2 = rng.uniform(5.0, 30.0)
            glolak[m_] = max(glolak[m_], f2)
            a += f2 / 100.0 * area[m_]
        lakarea[n] = a
    for n in res_cells:
        f = rng.uniform(2.0, 30.0)
        glores[n] = f
        resarea[n] = f / 100.0 * area[n]
    # a few regulated lakes (status 1) among the reservoir cells
    for n in res_cells[: max(1, n_res // 20)]:
        reg_status[n] = 1
        reglake[n] = glores[n]
    lr = rng.uniform(0, 1, ng) < 0.02
    locres[lr] = rng.uniform(0.1, 3.0, lr.sum())

    # keep total water <= contfreq - 1 (so that land area fraction stays positive) ...
    tot = loclak + locres + locwet + glowet + glolak + glores
    lim = np.maximum(contfreq.astype(np.float64) - 1.0, 0.0)
    over = tot > lim
    sc = np.where(over, lim / np.maximum(tot, 1e-30), 1.0)
    for arr in (loclak, locres, locwet, glowet):
        arr *= sc
    # scaling must not touch the lake/reservoir outflow-cell bookkeeping: only shrink the
    # non-outflow fractions, and if still too large, shrink the local ones to zero
    tot = loclak + locres + locwet + glowet + glolak + glores
    still = tot > lim
    for arr in (loclak, locres, locwet, glowet):
        arr[still] = 0.0
    # ... except for a handful of deliberately land-free cells (landAreaFrac == 0 branch,
    # daily.cpp:829/940/1080, routing.cpp:1885)
    n_zero = max(1, ng // 2500)
    cand = np.nonzero((contfreq == 100.0) & (glolak == 0) & (glores == 0) & (lakarea == 0) & (resarea == 0))[0]
    for n in cand[rng.permutation(cand.size)[:n_zero]]:
        loclak[n], locres[n], locwet[n], glowet[n] = 25.0, 0.0, 15.0, 60.0
    w.loclak = loclak.astype(np.float32)
    w.locwet = locwet.astype(np.float32)
    w.glowet = glowet.astype(np.float32)
    w.glolak = glolak.astype(np.float32)
    w.lakarea = lakarea.astype(np.float32)
    w.resarea = resarea.astype(np.float32)
    w.glores = glores.astype(np.float32)
    w.reglake = reglake.astype(np.float32)
    w.reg_status = reg_status
    w.locres = locres.astype(np.float32)
    w.res_type = np.where(resarea > 0, rng.integers(1, 3, ng), 0).astype(np.int8)
    w.res_start_year = np.where(resarea > 0, rng.integers(1900, 1991, ng), 0).astype(np.int32)
    # mean outflow [km3/month] on every cell (seasonal, so the operational-year start month
    # of rout_prepare.cpp:971-996 is defined everywhere)
    mean_out = 0.04 * acc.astype(np.float64) * rng.uniform(0.6, 1.4, ng)
    w.mean_outflow = mean_out.astype(np.float32)
    phase = np.where(w.lat >= 0, 4.0, 10.0) + rng.integers(-1, 2, ng)
    mm = np.arange(12)[None, :]
    w.mean_outflow12 = (mean_out[:, None] * (1.0 + 0.6 * np.sin(2 * np.pi * (mm - phase[:, None]) / 12.0))).astype(np.float32)
    annual = mean_out * 12.0
    w.stor_cap = np.where(resarea > 0, rng.uniform(0.1, 1.5, ng) * annual, 0.0).astype(np.float32)
    return w


def _erfinv(x):
    # Giles' single-precision-accurate approximation, good enough for elevation quantiles
    x = np.clip(x, -0.999999, 0.999999)
    w_ = -np.log((1.0 - x) * (1.0 + x))
    small = w_ < 5.0
    ws = w_ - 2.5
    p1 = 2.81022636e-08
    for c in (3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503,
              -0.00417768164, 0.246640727, 1.50140941):
        p1 = c + p1 * ws
    wl = np.sqrt(np.maximum(w_, 5.0)) - 3.0
    p2 = -0.000200214257
    for c in (0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613,
              0.00943887047, 1.00167406, 2.83297682):
        p2 = c + p2 * wl
    return np.where(small, p1, p2) * x


# ------------------------------------------------------------------------------------
# forcing: deterministic per (seed, year, month); float32 [ng][31] as in climate.cpp:93-138
# ------------------------------------------------------------------------------------
NDAYS = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
FIRST = [0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334]


def forcing_month(w, year, month):
    """month 1..12 -> dict of float32 arrays [ng][31] (P mm/d, T degC, SW, LW W/m2)."""
    rng = np.random.default_rng([w.seed, year, month, 7])
    ng = w.ng
    lat = w.lat
    doy = FIRST[month - 1] + np.arange(31)
    season = np.cos(2 * np.pi * (doy[None, :] - 200) / 365.0) * np.sign(lat)[:, None]
    tmean = 27.0 - 0.55 * np.abs(lat) - 0.0045 * w.altitude.astype(np.float64)
    amp = 2.0 + 0.28 * np.abs(lat)
    noise = rng.standard_normal((ng, 31))
    ar = np.empty_like(noise)
    ar[:, 0] = noise[:, 0]
    for d in range(1, 31):
        ar[:, d] = 0.7 * ar[:, d - 1] + 0.714 * noise[:, d]
    T = tmean[:, None] + amp[:, None] * season + 3.0 * ar
    wet = rng.uniform(0, 1, (ng, 31)) < 0.3
    theta = np.where(w.arid == 1, 3.0, 11.0)[:, None] * (1.0 + 0.4 * season)
    P = np.where(wet, rng.gamma(0.7, 1.0, (ng, 31)) * theta, 0.0)
    decl = 0.409 * np.sin(2 * np.pi * (doy - 81) / 365.0)
    phi = np.deg2rad(lat)[:, None]
    x = np.clip(-np.tan(phi) * np.tan(decl)[None, :], -1, 1)
    ws = np.arccos(x)
    ra = 1367.0 / np.pi * (ws * np.sin(phi) * np.sin(decl)[None, :] + np.cos(phi) * np.cos(decl)[None, :] * np.sin(ws))
    SW = np.maximum(ra, 0.0) * 0.75 * rng.uniform(0.3, 1.0, (ng, 31))
    LW = np.clip(0.82 * 5.67e-8 * (T + 273.15) ** 4 + rng.uniform(-15, 15, (ng, 31)), 150.0, 450.0)
    return {"P": P.astype(np.float32), "T": T.astype(np.float32),
            "SW": SW.astype(np.float32), "LW": LW.astype(np.float32)}


# ------------------------------------------------------------------------------------
# text tables
# ------------------------------------------------------------------------------------
LCT = [  # idx rootDepth albedo albedoSnow ddf emissivity   (daily.cpp:1946-1957)
    (1, 2.0, 0.11, 0.20, 1.5, 0.9950), (2, 4.0, 0.07, 0.20, 3.0, 0.9950), (3, 2.0, 0.13, 0.30, 1.5, 0.9900),
    (4, 2.0, 0.13, 0.20, 3.0, 0.9900), (5, 2.0, 0.12, 0.25, 2.0, 0.9920), (6, 1.0, 0.13, 0.30, 3.0, 0.9830),
    (7, 0.5, 0.13, 0.40, 4.0, 0.9540), (8, 1.5, 0.13, 0.40, 4.0, 0.9830), (9, 1.5, 0.20, 0.45, 4.0, 0.9930),
    (10, 1.0, 0.25, 0.50, 4.0, 0.9930), (11, 1.0, 0.20, 0.50, 4.0, 0.9830), (12, 1.0, 0.18, 0.50, 4.0, 0.9830),
    (13, 1.0, 0.23, 0.60, 4.0, 0.9810), (14, 1.0, 0.19, 0.60, 4.0, 0.9830), (15, 1.0, 0.60, 0.70, 6.0, 0.9999),
    (16, 0.1, 0.30, 0.60, 6.0, 0.9410), (17, 1.0, 0.15, 0.60, 5.0, 0.9920), (18, 2.0, 0.15, 0.40, 3.0, 0.9950)]
LAI = [  # idx LAI fracDeciduous evergreenReduction initialDays kc_min kc_max (lai.cpp:133-146)
    (1, 4.02, 0.00, 1.0, 1, 1.00, 1.00), (2, 4.78, 0.00, 0.8, 1, 0.95, 1.05), (3, 4.63, 1.00, 0.8, 10, 0.55, 1.05),
    (4, 3.49, 0.00, 0.8, 10, 0.90, 1.00), (5, 4.90, 0.25, 0.8, 10, 0.60, 1.05), (6, 2.18, 0.50, 0.8, 10, 0.40, 0.95),
    (7, 1.71, 0.50, 0.8, 10, 0.40, 0.90), (8, 2.52, 1.00, 0.8, 10, 0.30, 0.90), (9, 2.08, 0.50, 0.5, 10, 0.40, 0.95),
    (10, 1.71, 0.50, 0.8, 10, 0.30, 1.00), (11, 1.54, 0.50, 0.8, 10, 0.30, 1.00), (12, 6.34, 1.00, 0.8, 10, 0.50, 1.10),
    (13, 3.62, 1.00, 0.8, 10, 0.35, 1.10), (14, 3.62, 0.50, 0.8, 10, 0.35, 1.05), (15, 0.00, 0.00, 0.8, 10, 0.50, 0.50),
    (16, 1.31, 1.00, 0.8, 10, 0.25, 0.40), (17, 1.88, 1.00, 0.8, 10, 0.30, 0.90), (18, 2.30, 0.50, 0.8, 10, 0.20, 0.80)]

# canonical OPTIONS.DAT vector; index 2 = grid_store
OPTIONS = [2, 1, 6, 0, 0, 0, 1, 1, 0, 0, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 1, 0, 1, 1, 1, 0, 0, 0,
           2000, 1900, 2010, 0, 1971, 2000, 0, 0]
OPTION_NAMES = ["fileEndianType", "basin", "grid_store", "grid_store_TypeForStorages", "day_store", "time_series",
                "cloud", "intercept", "calc_albedo", "petOpt", "use_kc", "landCoverOpt", "rout_prepare",
                "timeStepCheckFlag", "riverveloOpt", "subtract_use", "use_alloc", "delayedUseSatisfaction", "clclOpt",
                "permaOpt", "resOpt", "statcorrOpt", "aridareaOpt", "fractionalRoutingOpt", "riverEvapoOpt",
                "aggrNUsGloLakResOpt", "climate_spatial_resolution", "resYearOpt", "resYearReference",
                "resYearFirstToUse", "resYearLastToUse", "antNatOpt", "resNUsMeanYearFirst", "resNUsMeanYearLast",
                "calc_wtemp", "glacierOpt"]

PARAM_NAMES = ["gammaHBV_runoff_coeff", "CFA_cellCorrFactor", "CFS_statCorrFactor", "root_depth_multiplier",
               "river_roughness_coeff_mult", "lake_depth", "wetland_depth", "surfacewater_outflow_coefficient",
               "evapo_red_fact_exp_mult", "net_radiation_mult", "PT_coeff_humid", "PT_coeff_arid", "max_daily_PET",
               "mcwh", "LAI_mult", "snow_freeze_temp", "snow_melt_temp", "degree_day_factor_mult",
               "temperature_gradient", "gw_factor_mult", "rg_max_mult", "pcrit_aridgw", "groundwater_outflow_coeff",
               "net_abstraction_surfacewater_mult", "net_abstraction_groundwater_mult", "precip_mult"]
PARAM_DEFAULT = [2.0, 1.0, 1.0, 1.0, 1.0, 5.0, 2.0, 0.01, 1.0, 1.0, 1.26, 1.74, 15.0, 0.3, 1.0, 0.0, 0.0, 1.0,
                 0.006, 1.0, 1.0, 12.5, 0.01, 1.0, 1.0, 1.0]


def default_params(w, variant=0):
    """[26][ng] float64. variant 0 = per-cell perturbed defaults (so that every per-cell lookup
    matters); variant i>0 = calibration set i of SURVEY 8d (config 3)."""
    rng = np.random.default_rng([w.seed, 991, variant])
    ng = w.ng
    p = np.tile(np.array(PARAM_DEFAULT)[:, None], (1, ng))
    if variant == 0:
        p[0] = np.round(rng.uniform(0.3, 4.5, ng), 3)        # gamma
        p[1] = np.round(rng.uniform(0.7, 1.3, ng), 3)        # CFA
        p[7] = np.round(10 ** rng.uniform(-2.3, -1.3, ng), 5)  # sw outflow
        p[22] = np.round(10 ** rng.uniform(-2.3, -1.3, ng), 5)  # gw outflow
        p[15] = np.round(rng.uniform(-1, 1, ng), 2)
        p[16] = np.round(rng.uniform(-1, 1, ng), 2)
        p[25] = np.round(rng.uniform(0.9, 1.1, ng), 3)
    else:
        p[0] = rng.uniform(0.1, 5.0)
        p[1] = rng.uniform(0.5, 1.5)
        p[7] = 10 ** rng.uniform(-3, -1)
        p[22] = 10 ** rng.uniform(-3, -1)
        p[15] = rng.uniform(-1, 1)
        p[16] = rng.uniform(-1, 1)
        for k in (3, 4, 8, 9, 14, 17, 19, 20, 25):
            p[k] = rng.uniform(0.8, 1.2)
    return p


def reservoir_years(w, years):
    """resYearOpt 1 variant of the world's reservoirs (routing.cpp:1052-1412): a third operates from the start - every second of
    them with 60 % of its land-cover fraction in the first year and the full fraction from the second (the "existing reservoir
    grew" branch) -, a third comes on line in the second year, a third in the third.
    -> (start_year [ng] int32, {year: G_RES_<year> fraction [ng]})"""
    idx = np.flatnonzero(w.resarea > 0)
    start = w.res_start_year.copy()
    frac = {}
    for y in range(years[0], years[1] + 1):
        frac[y] = np.zeros(w.ng)
    for i, n in enumerate(idx):
        k = i % 3
        start[n] = years[0] + k if k else min(int(start[n]), years[0] - 1)
        for y in frac:
            if y >= start[n]:
                frac[y][n] = w.glores[n] * (0.6 if (k == 0 and i % 2 and y == years[0]) else 1.0)
    return start.astype(np.int32), frac


def write_world(w, out, years=(1901, 1901), months=(1, 12), grid_store=6, daily_discharge=True, params=None, water_use=False, time_series=0,
                res_year_opt=0):
    inp = os.path.join(out, "input")
    clim = os.path.join(out, "climate")
    rout = os.path.join(out, "routing")
    outd = os.path.join(out, "output")
    for d in (inp, clim, rout, outd, os.path.join(inp, "G_RES")):
        os.makedirs(d, exist_ok=True)
    ng = w.ng
    write_unf(f"{inp}/GCRC.UNF4", w.gcrc.T, "i4")  # file index = col*360+row
    write_unf(f"{inp}/GR.UNF2", w.row, "i2")
    write_unf(f"{inp}/GC.UNF2", w.col, "i2")
    write_unf(f"{inp}/GAREA.UNF0", w.area_row, "f4")
    write_unf(f"{inp}/GCONTFREQ.UNF0", w.contfreq, "f4")
    write_unf(f"{inp}/G_FLOWDIR.UNF2", w.flowdir, "i2")
    write_unf(f"{inp}/GALTMOD.UNF0", w.altitude, "f4")
    write_unf(f"{inp}/G_MEANDERING_RATIO.UNF0", w.meander, "f4")
    write_unf(f"{inp}/G_ELEV_RANGE.101.UNF2", w.elev_range, "i2")
    write_unf(f"{inp}/G_LANDCOVER.UNF1", w.landcover, "i1")
    write_unf(f"{inp}/GBUILTUP.UNF0", w.builtup, "f4")
    write_unf(f"{inp}/G_ARID_HUMID.UNF2", w.arid, "i2")
    write_unf(f"{inp}/G_TAWC.UNF0", w.tawc, "f4")
    write_unf(f"{inp}/G_SLOPE_CLASS.UNF1", w.slope_class, "i1")
    write_unf(f"{inp}/G_TEXTURE.UNF1", w.texture, "i1")
    write_unf(f"{inp}/G_PERMAGLAC.UNF1", w.permaglac, "i1")
    write_unf(f"{inp}/G_AQ_FACTOR.UNF1", w.aq_factor, "i1")
    write_unf(f"{inp}/G_GW_FACTOR_CORR.UNF0", w.gw_factor_corr, "f4")
    write_unf(f"{inp}/G_GLOLAK.UNF0", w.glolak, "f4")
    write_unf(f"{inp}/G_LOCLAK.UNF0", w.loclak, "f4")
    write_unf(f"{inp}/G_GLOWET.UNF0", w.glowet, "f4")
    write_unf(f"{inp}/G_LOCWET.UNF0", w.locwet, "f4")
    write_unf(f"{inp}/G_REGLAKE.UNF0", w.reglake, "f4")
    write_unf(f"{inp}/G_LOCRES.UNF0", w.locres, "f4")
    write_unf(f"{inp}/G_LAKAREA.UNF0", w.lakarea, "f4")
    write_unf(f"{inp}/G_RESAREA.UNF0", w.resarea, "f4")
    write_unf(f"{inp}/G_REG_LAKE.UNF1", w.reg_status, "i1")
    write_unf(f"{inp}/G_RES_TYPE.UNF1", w.res_type, "i1")
    res_start, res_frac = (w.res_start_year, {}) if not res_year_opt else reservoir_years(w, years)
    write_unf(f"{inp}/G_START_YEAR.UNF4", res_start, "i4")
    for y, fr in res_frac.items():
        write_unf(f"{inp}/G_RES/G_RES_{y}.UNF0", fr, "f4")
    write_unf(f"{inp}/G_STORAGE_CAPACITY.UNF0", w.stor_cap, "f4")
    write_unf(f"{inp}/G_MEAN_OUTFLOW.UNF0", w.mean_outflow, "f4")
    write_unf(f"{inp}/G_MEAN_OUTFLOW.12.UNF0", w.mean_outflow12, "f4")
    write_unf(f"{inp}/G_NUs_1971_2000.UNF0", np.zeros(ng), "f4")
    write_unf(f"{inp}/G_RES/G_RES_2000.UNF0", w.glores, "f4")
    write_unf(f"{inp}/G_RES/G_RES_FRAC.UNF0", w.glores, "f4")
    write_unf(f"{inp}/G_OUTFLOW_CELL_ASSIGNMENT.UNF4", np.arange(1, ng + 1), "i4")
    write_unf(f"{inp}/GLWDunits.UNF4", np.arange(1, ng + 1), "i4")
    write_unf(f"{inp}/G_ROUGHNESS.UNF0", w.roughness, "f4")
    write_unf(f"{inp}/G_BANKFULL.UNF0", w.bankfull, "f4")
    if time_series == 1:
        # yearly [cell][365] files of climateYear.cpp:38-58 (time_series 1): the same daily values as the monthly .31 files
        for y in range(years[0], years[1] + 1):
            yr = {k: np.zeros((ng, 365), np.float32) for k in ("P", "T", "SW", "LW")}
            d0 = 0
            for m in range(1, 13):
                nd = (31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31)[m - 1]
                if months[0] <= m <= months[1] or years[0] != years[1]:
                    f = forcing_month(w, y, m)
                    for k in yr:
                        yr[k][:, d0:d0 + nd] = f[k][:, :nd]
                d0 += nd
            write_unf(f"{clim}/G_GPCC_H08day_V20110128_{y}.365.UNF0", yr["P"], "f4")
            write_unf(f"{clim}/G_TEMP_H08_int_{y}.365.UNF0", yr["T"], "f4")
            write_unf(f"{clim}/G_SSRD_H08_int_{y}.365.UNF0", yr["SW"], "f4")
            write_unf(f"{clim}/G_SLRD_H08_int_{y}.365.UNF0", yr["LW"], "f4")
    for y in range(years[0], years[1] + 1):
        for m in (range(1, 13) if years[0] != years[1] else range(months[0], months[1] + 1)):  # several years: every month
            if time_series == 1:
                break
            f = forcing_month(w, y, m)
            write_unf(f"{clim}/GPREC_{y}_{m}.31.UNF0", f["P"], "f4")
            write_unf(f"{clim}/GTEMP_{y}_{m}.31.UNF0", f["T"], "f4")
            write_unf(f"{clim}/GSHORTWAVE_{y}_{m}.31.UNF0", f["SW"], "f4")
            write_unf(f"{clim}/GLONGWAVE_DOWN_{y}_{m}.31.UNF0", f["LW"], "f4")
    with open(f"{inp}/LCT_22.DAT", "w") as fh:
        fh.write("# idx rootingDepth albedo albedoSnow ddf emissivity\n")
        for r in LCT:
            fh.write("%d %.2f %.2f %.2f %.1f %.4f\n" % r)
    with open(f"{inp}/LAI_22.DAT", "w") as fh:
        fh.write("# idx LAI fracDeciduous evergreenLAIreduction initialDays kc_min kc_max\n")
        for r in LAI:
            fh.write("%d %.2f %.2f %.1f %d %.2f %.2f\n" % r)
    opts = list(OPTIONS)
    opts[2] = grid_store
    opts[OPTION_NAMES.index("time_series")] = time_series
    if res_year_opt:  # the reference year has to lie inside [first, last] (option.cpp:639)
        opts[OPTION_NAMES.index("resYearOpt")] = 1
        opts[OPTION_NAMES.index("resYearReference")] = years[0]
        opts[OPTION_NAMES.index("resYearFirstToUse")] = years[0]
        opts[OPTION_NAMES.index("resYearLastToUse")] = years[1]
    if water_use:
        # SURVEY 8f-4 (next row): net abstractions from surface water / groundwater, m3 per month (routing.cpp:884-977),
        # subtract_use = 2 with the other use options at their canonical 0
        opts[OPTION_NAMES.index("subtract_use")] = 2  # 2 (and 3) read the net abstractions, integrateWGHM.cpp:645
        rng = np.random.default_rng([w.seed, 4711])
        has = rng.random(ng) < 0.35
        for y in range(years[0], years[1] + 1):
            season = 1.0 + 0.5 * np.sin(np.arange(12) / 12.0 * 2 * np.pi)
            nus = np.where(has[:, None], rng.gamma(0.6, 4.0e6, (ng, 1)) * season[None, :], 0.0)
            nus[rng.random(ng) < 0.05] *= -0.3          # return flows exceed the withdrawal in a few cells
            nug = np.where(has[:, None], rng.gamma(0.6, 2.0e6, (ng, 1)) * season[None, :], 0.0) * rng.choice([1.0, -0.2], (ng, 1), p=[0.9, 0.1])
            write_unf(f"{inp}/G_NETUSE_SW_m3_{y}.12.UNF0", nus, "f4")
            write_unf(f"{inp}/G_NETUSE_GW_m3_{y}.12.UNF0", nug, "f4")
            write_unf(f"{inp}/G_IRRIG_WITHDRAWAL_USE_SW_m3_{y}.12.UNF0", np.abs(nus) * 1.6, "f4")
            write_unf(f"{inp}/G_IRRIG_CONS_USE_SW_m3_{y}.12.UNF0", np.abs(nus) * 0.8, "f4")
        write_unf(f"{inp}/G_FRACTRETURNGW_IRRIG.UNF0", rng.uniform(0.1, 0.6, ng), "f4")
    with open(f"{out}/OPTIONS.DAT", "w") as fh:
        for name, v in zip(OPTION_NAMES, opts):
            fh.write(f"# {name}\nValue: {v}\n")
    with open(f"{out}/OUTPUT_OPTIONS.DAT", "w") as fh:
        for i in range(142):
            on = 1 if (daily_discharge and i == 92) else 0
            fh.write(f"{on} option{i}\n")
    with open(f"{out}/ROUTING.DAT", "w") as fh:
        fh.write("# river velocity [km/d], time steps per day\n86.4 1\n")
    with open(f"{out}/STATIONS.DAT", "w") as fh:
        fh.write("")
    p = default_params(w, 0) if params is None else params
    js = {"ng_param": ng, "gcrc_cellnumber": list(range(1, ng + 1)), "arc_id": list(range(1, ng + 1))}
    for k, name in enumerate(PARAM_NAMES):
        js[name] = [float(v) for v in p[k]]
    with open(f"{out}/parameters.json", "w") as fh:
        json.dump(js, fh)
    p.astype("<f8").tofile(f"{out}/parameters.f64")  # [26][ng], for the non-JSON consumers
    with open(f"{out}/config.txt", "w") as fh:
        fh.write(f"""# synthetic world ng={ng} seed={w.seed}
param_json {out}/parameters.json
output_state_lastday {outd}/wghm_state_lastday.txt
output_snowInElevation_lastday {outd}/snow_lastday.txt
additionalOutIn_lastday {outd}/additional_lastday.txt
start_month {months[0]}
start_year {years[0]}
end_month {months[1]}
end_year {years[1]}
time_step 1
num_init_years 0
runtime_options {out}/OPTIONS.DAT
output_options {out}/OUTPUT_OPTIONS.DAT
routing {out}/ROUTING.DAT
stations {out}/STATIONS.DAT
input_dir {inp}
output_dir {outd}
climate_dir {clim}
routing_dir {rout}
water_use_dir {inp}
end_of_head
""")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--ng", type=int, default=67420)
    ap.add_argument("--seed", type=int, default=20240607)
    ap.add_argument("--years", type=int, nargs=2, default=[1901, 1901])
    ap.add_argument("--months", type=int, nargs=2, default=[1, 12])
    ap.add_argument("--grid-store", type=int, default=6)
    a = ap.parse_args()
    w = build_world(a.ng, a.seed)
    write_world(w, os.path.abspath(a.out), tuple(a.years), tuple(a.months), a.grid_store)
    print(f"world ng={w.ng} written to {a.out}: {int((w.flowdir == 0).sum())} ocean outlets, "
          f"{int((w.flowdir == -1).sum())} inland sinks, max flow acc {int(w.acc.max())}")


if __name__ == "__main__":
    main()
