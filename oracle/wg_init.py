"""numpy restatement of the reference's INIT-TIME derivations (TEST INFRASTRUCTURE).

Turns a synthetic world (oracle/synth_world.py) plus a [26][ng] parameter block into the
derived static grids and the cold-start state that the daily hot path consumes, following

  geo.cpp:7-45            continental fraction / contcell
  s_max.cpp:40-75         G_Smax  = TAWC * (M_ROOT_D * rootingDepth_lct)        (float grid)
  gw_frac.cpp:36-275      G_gwFactor, G_Rgmax (short, x100)                     (float / short)
  lai.cpp:40-108          G_LAImax, lai_factor_a/b                              (float)
  routing.cpp:131-538     routing.init (lake/wetland/reservoir grids, river geometry)
  routing.cpp:745-765     initFractionStatus
  routing.cpp:789-847     setStoragesToZero
  routing.cpp:979-1291    annualInit (reservoirs in operation, initial land area fraction)
  routing.cpp:5647-5720   setLakeWetlToMaximum
  integrateWGHM.cpp:318-365  G_toBeCalculated, gamma / CFA / CFS from the parameter file

All float32 roundings of the reference (Grid<float> members, float locals) are reproduced
with numpy float32 arithmetic.  tests/test_oracle_vs_ref.py checks every array produced here
bit-for-bit against the in-memory values dumped from the compiled reference (record day 0).
"""
import math

import numpy as np

from . import synth_world as sw
from . import wgo

f32 = np.float32
f64 = np.float64


def derive(w, params=None, resYearReference=2000):
    """-> dict name -> array, names as in oracle/wg_oracle.h (static fields + initial state)."""
    ng = w.ng
    p = sw.default_params(w, 0) if params is None else np.asarray(params, f64)
    out = {}
    topo = wgo.rout_prepare(w.flowdir, w.row, w.col, w.gcrc.T)
    cd, slope_f, length_f = wgo.river_geometry(w.altitude, w.meander, topo["outflow_cell"], topo["ldd"], w.row, w.col)
    alloc, start_month = wgo.reservoir_prepare(w.resarea, w.mean_outflow, w.mean_outflow12, topo["outflow_cell"])
    out["_topology"] = topo
    out["wu_alloc_coeff"] = np.asarray(alloc, f64).reshape(ng, 5)  # G_ALLOC_COEFF.5.UNF0 read into a Grid<double> (routing.cpp:361)

    # ---- geo.cpp -------------------------------------------------------------------------
    area = w.area_row.astype(f64)[w.row.astype(int) - 1]
    contfreq = w.contfreq.astype(f64)
    fwater_const = w.loclak.astype(f64) + w.glolak.astype(f64) + w.locres.astype(f64) + w.glores.astype(f64)
    contfreq = np.where(fwater_const > contfreq, fwater_const, contfreq)
    out["area"] = area
    out["contfreq"] = contfreq
    out["contcell"] = (contfreq >= 0.000001).astype(np.int16)
    out["row"] = w.row.astype(np.int16)
    out["toBeCalculated"] = np.ones(ng, np.int16)
    out["landcover"] = w.landcover.astype(np.int8)
    out["builtup"] = w.builtup.astype(f32)
    out["arid"] = w.arid.astype(np.int16)
    out["ldd"] = topo["ldd_2"].astype(np.int8)  # routing reads G_LDD_2.UNF1 (integrateWGHM.cpp:250)
    out["elevation"] = w.elev_range.astype(np.int16)
    out["texture"] = w.texture.astype(np.int8)
    lc = w.landcover.astype(int) - 1

    # ---- tables ----------------------------------------------------------------------------
    lct = np.array(sw.LCT, f64)
    lai = np.array(sw.LAI, f64)
    rooting_f = lct[:, 1].astype(f32)
    out["lct_albedo"] = lct[:, 2].copy()
    out["lct_albedo_snow"] = lct[:, 3].copy()
    out["lct_ddf"] = lct[:, 4].copy()
    out["lct_emissivity"] = lct[:, 5].copy()
    lai_f = lai[:, 1].astype(f32)
    dec_f = lai[:, 2].astype(f32)
    evr_f = lai[:, 3].astype(f32)
    out["lai_factor_a"] = (0.1 * dec_f.astype(f64)).astype(f32)                 # lai.cpp:93
    out["lai_factor_b"] = ((f32(1) - dec_f) * evr_f).astype(f32)               # lai.cpp:94 (int 1 - float, float math)
    out["lai_initial_days"] = lai[:, 4].astype(np.int16)
    out["lai_kc_min"] = lai[:, 5].copy()
    out["lai_kc_max"] = lai[:, 6].copy()

    # ---- parameters ------------------------------------------------------------------------
    out["params"] = p.copy()
    out["gamma_hbv"] = p[0].copy()
    out["cfa"] = p[1].copy()
    out["cfs"] = p[2].copy()
    # lai.cpp:86  G_LAImax(float) = M_LAI * LeafAreaIndex[lct]
    out["laimax"] = (p[14] * lai_f[lc].astype(f64)).astype(f32)
    # s_max.cpp:58-64
    rootdepth = (p[3] * rooting_f[lc].astype(f64)).astype(f32)
    tawc = w.tawc.astype(f32)
    out["smax"] = np.where(tawc < 0, f32(-9999), tawc * rootdepth).astype(f32)

    # ---- gw_frac.cpp -------------------------------------------------------------------------
    sc = w.slope_class.astype(np.int64).copy()
    sc[sc == 0] = 10
    tab_c = np.array([10, 20, 30, 40, 50, 60, 70])
    tab_f = np.array([1.00, 0.95, 0.90, 0.75, 0.60, 0.30, 0.15], f32)
    slope_factor = np.zeros(ng, f32)
    for n in range(ng):
        s = sc[n]
        if s == 70:
            slope_factor[n] = tab_f[6]
            continue
        for i in range(6):
            if tab_c[i] == s:
                slope_factor[n] = tab_f[i]
                break
            if tab_c[i] < s < tab_c[i + 1]:
                # float + ((float - float) / int) * int  : float arithmetic
                slope_factor[n] = f32(tab_f[i] + f32(f32(f32(tab_f[i + 1] - tab_f[i]) / f32(tab_c[i + 1] - tab_c[i])) * f32(s - tab_c[i])))
                break
    tex = w.texture.astype(np.int64)
    ttab = np.array([10, 20, 30])
    rg_tab = np.array([7., 4.5, 2.5], f32)  # WFD daily values (time_series == 0), gw_frac.cpp:101-107
    tf_tab = np.array([1, 0.95, 0.70], f32)
    texture_factor = np.zeros(ng, f32)
    rgmax_f = np.zeros(ng, f32)
    M_RG = p[20]
    for n in range(ng):
        t = tex[n]
        if t <= 0 or t == 2:
            texture_factor[n] = f32(0.95)
            rgmax_f[n] = 3
            continue
        if t == 1:
            continue
        for i in range(3):
            if ttab[i] == t:
                texture_factor[n] = tf_tab[i]
                rgmax_f[n] = f32(M_RG[n] * f64(rg_tab[i]))
                break
            if i < 2 and ttab[i] < t < ttab[i + 1]:
                texture_factor[n] = f32(tf_tab[i] + f32(f32(f32(tf_tab[i + 1] - tf_tab[i]) / f32(ttab[i + 1] - ttab[i])) * f32(t - ttab[i])))
                # double: (M*R[i]) + ((M * (R[i+1]-R[i]) / int) * int)
                d = f64(rg_tab[i + 1] - rg_tab[i])  # float subtraction, then promoted
                rgmax_f[n] = f32((M_RG[n] * f64(rg_tab[i])) + ((M_RG[n] * d / f64(ttab[i + 1] - ttab[i])) * f64(t - ttab[i])))
                break
    gwf = np.zeros(ng, f32)
    for n in range(ng):
        if texture_factor[n] < 0:
            gwf[n] = -99
            continue
        slopeF = slope_factor[n]
        texF = texture_factor[n]
        aqF = f32(f64(np.int16(w.aq_factor[n])) / 100.0)
        perma = f32(1. - (f64(f32(w.permaglac[n])) / 100.))  # (float)x / 100. is a double division
        slopeF = min(max(slopeF, f32(0)), f32(1))
        aqF = min(max(aqF, f32(0)), f32(1))
        texF = min(max(texF, f32(0)), f32(1))
        perma = min(max(perma, f32(0)), f32(1))
        g = f32(p[19][n] * f64(slopeF) * f64(texF) * f64(aqF) * f64(perma))
        if g > 1.:
            g = f32(0.95)
        gwf[n] = g
    corr = w.gw_factor_corr.astype(f32)
    m = corr > 0
    if m.any():
        g2 = (p[19] * corr.astype(f64)).astype(f32)
        g2[g2 > 1.] = f32(0.95)
        gwf[m] = g2[m]
    out["gwfactor"] = gwf
    # round_to_short(G_Rgmax_f * 100) (gw_frac.cpp:229-234)
    out["rgmax"] = np.where(rgmax_f >= 0, np.floor(rgmax_f.astype(f32) * f32(100) + f32(0.5)), -9999).astype(np.int16)

    # ---- routing.init ---------------------------------------------------------------------
    clip0 = lambda a: np.where(a < 0, 0.0, a.astype(f64))
    glo_lake = clip0(w.glolak)
    loc_lake = clip0(w.loclak)
    glo_wet = clip0(w.glowet)
    loc_wet = clip0(w.locwet)
    lake_area = clip0(w.lakarea)
    res_area_full = clip0(w.resarea)
    loc_lake = loc_lake + w.locres.astype(f64)  # G_loc_lake += G_loc_res (routing.cpp:343)
    mean_NUs = np.zeros(ng)
    down = topo["outflow_cell"]
    mean_demand = mean_NUs.copy()  # all zero without water use; alloc loop adds zeros
    stor_cap_full = np.where(w.stor_cap.astype(f64) < 0, 0.0, w.stor_cap.astype(f64))
    mean_outflow = w.mean_outflow.astype(f64) * 12. * 1000000000. / 31536000.
    mean_demand = mean_demand / 31536000.
    river_length = length_f.astype(f64) * (contfreq / 100.)
    bankfull = np.maximum(w.bankfull.astype(f64), 0.05)
    # libm pow via math.pow (numpy's vectorised power differs from glibc in the last bit)
    width_bf = 2.71 * np.array([math.pow(x, 0.557) for x in bankfull])
    depth_bf = 0.349 * np.array([math.pow(x, 0.341) for x in bankfull])
    bottom = width_bf - 2.0 * 2.0 * depth_bf
    stor_max = river_length * 0.5 * depth_bf / 1000. * (bottom / 1000. + width_bf / 1000.)
    out["loc_lake"] = loc_lake
    out["loc_wetland"] = loc_wet
    out["glo_wetland"] = glo_wet
    out["glo_lake"] = glo_lake
    out["lake_area"] = lake_area
    out["mean_outflow"] = mean_outflow
    out["mean_demand"] = mean_demand
    out["res_type"] = w.res_type.astype(np.int8)
    out["start_month"] = start_month.astype(np.int8)
    out["river_length"] = river_length
    out["river_slope"] = slope_f.astype(f64)
    out["roughness"] = w.roughness.astype(f64)
    out["river_bottom_width"] = bottom
    out["river_width_bf"] = width_bf
    out["river_storage_max"] = stor_max
    out["lake_depth_active"] = p[5] * 0.001
    out["wetl_depth_active"] = p[6] * 0.001
    out["downstream_cell"] = down.astype(np.int32)
    routing_cell = np.zeros(ng, np.int32)
    routing_cell[topo["rout_order"] - 1] = np.arange(1, ng + 1)
    out["routing_cell"] = routing_cell
    out["rout_order"] = topo["rout_order"].astype(np.int32)

    # ---- annualInit (first year; resYearOpt 0 -> G_RES_<resYearReference>) ------------------
    glo_res = w.glores.astype(f64)
    reservoir_area = np.zeros(ng)
    stor_cap = np.zeros(ng)
    reg = w.reg_status.astype(int) == 1
    reservoir_area[reg] = res_area_full[reg]
    stor_cap[reg] = stor_cap_full[reg]
    inop = resYearReference >= w.res_start_year.astype(int)
    reservoir_area[inop] = res_area_full[inop]
    stor_cap[inop] = stor_cap_full[inop]
    # reservoirs of unknown type or without mean outflow become global lakes (routing.cpp:1107-1146)
    bad = (reservoir_area > 0) & inop & ((w.res_type.astype(int) == 0) | (mean_outflow <= 0))
    lake_area = lake_area.copy()
    lake_area[bad] += res_area_full[bad]
    reservoir_area[bad] = 0.
    out["lake_area"] = lake_area
    out["reservoir_area"] = reservoir_area
    out["stor_cap"] = stor_cap
    out["glo_res"] = glo_res
    laf = contfreq - (glo_lake + glo_wet + loc_lake + loc_wet + glo_res)
    laf = np.where(laf < 0, 0.0, laf)

    # ---- initFractionStatus / storages ------------------------------------------------------
    fLocLake = loc_lake / 100.
    fLocWet = loc_wet / 100.
    fGloWet = glo_wet / 100.
    out["f_glo_lake"] = glo_lake / 100.
    out["fswb_init"] = fLocLake + fLocWet + fGloWet
    st = {}
    z = np.zeros(ng)
    for k in ("canopy", "soil", "snow", "lai_precsum", "gw", "river_stor", "land_area_frac_prev", "land_area_frac_next",
              "red_river", "red_res"):
        st[k] = z.copy()
    st["snow_bands"] = np.zeros((ng, 101))
    st["lai_days"] = np.zeros(ng, np.int32)
    st["lai_status"] = np.zeros(ng, np.int32)
    st["status_laf_next"] = np.zeros(ng, np.int16)
    st["fswb_laf"] = out["fswb_init"].copy()
    st["fswb_laf_next"] = out["fswb_init"].copy()
    st["land_area_frac"] = laf
    st["k_release"] = np.full(ng, 0.1)
    # setLakeWetlToMaximum (routing.cpp:5647-5720)
    st["loc_lake_stor"] = (loc_lake / 100.) * area * out["lake_depth_active"]
    st["loc_wetl_stor"] = (loc_wet / 100.) * area * out["wetl_depth_active"]
    st["glo_lake_stor"] = w.lakarea.astype(f64).clip(min=0) * out["lake_depth_active"]
    st["glo_wetl_stor"] = (glo_wet / 100.) * area * out["wetl_depth_active"]
    res_stor = np.zeros(ng)
    full = (resYearReference >= w.res_start_year.astype(int)) & (stor_cap_full > -99)
    res_stor[full] = stor_cap_full[full]
    late = reg & (resYearReference < w.res_start_year.astype(int))
    res_stor[late] += res_area_full[late] * out["lake_depth_active"][late]
    st["res_stor"] = res_stor
    # NB: the second if/else of setLakeWetlToMaximum (resYearOpt == 1 ... else factor = 0)
    # overrides the first one: with resYearOpt == 0 the reservoir reduction factor starts at 0
    st["red_res"] = np.zeros(ng)
    st["red_loc_lake"] = np.ones(ng)
    st["red_loc_wetl"] = np.ones(ng)
    st["red_glo_lake"] = np.ones(ng)
    st["red_glo_wetl"] = np.ones(ng)
    st["red_river"] = np.full(ng, 0.5)
    st["river_area_frac_next"] = st["red_river"] * river_length * width_bf / 1000. / area
    out.update(st)
    return out


STATIC_FIELDS = ["area", "contfreq", "contcell", "row", "toBeCalculated", "landcover", "builtup", "arid", "ldd",
                 "elevation", "texture", "smax", "gwfactor", "rgmax", "laimax", "gamma_hbv", "cfa", "cfs", "params",
                 "lai_factor_a", "lai_factor_b", "lai_initial_days", "lai_kc_min", "lai_kc_max", "lct_albedo",
                 "lct_albedo_snow", "lct_ddf", "lct_emissivity", "loc_lake", "loc_wetland", "glo_wetland", "lake_area",
                 "reservoir_area", "stor_cap", "mean_outflow", "mean_demand", "res_type", "start_month", "river_length",
                 "river_slope", "roughness", "river_bottom_width", "river_width_bf", "river_storage_max",
                 "lake_depth_active", "wetl_depth_active", "fswb_init", "f_glo_lake"]
STATE_FIELDS = ["canopy", "soil", "snow", "snow_bands", "lai_days", "lai_status", "lai_precsum", "gw", "loc_lake_stor",
                "loc_wetl_stor", "glo_lake_stor", "glo_wetl_stor", "res_stor", "river_stor", "red_loc_lake",
                "red_loc_wetl", "red_glo_lake", "red_glo_wetl", "red_res", "red_river", "k_release", "land_area_frac",
                "land_area_frac_prev", "land_area_frac_next", "fswb_laf", "fswb_laf_next", "river_area_frac_next",
                "status_laf_next"]
FLUX_FIELDS = ["lake_balance", "openwater_prec", "openwater_pet", "surface_runoff", "gw_recharge", "storage_transfer",
               "land_aet", "land_aet_uncorr", "discharge", "river_evapo", "gwr_swb", "cell_runoff"]
