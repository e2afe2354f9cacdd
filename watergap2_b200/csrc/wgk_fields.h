// wgk_fields.h — the single table of device arrays of the wgk context.
//
// X(name, ctype, dtype, scope, bands)
//   scope  CELL   : static, shared by all members and parameter sets           [cell]
//          PSET   : derived from the per-cell calibration parameters            [pset][cell]
//          MEMBER : state / daily fluxes of one ensemble member                 [member][cell]
//          TABLE  : land-cover-class table, 18 entries
//   bands  1, or 101 for the elevation-band arrays, which are stored BAND-MAJOR on the device
//          ([band][cell], so that the 100-band snow loop of one warp reads 32 consecutive
//          doubles per band) while the host/reference layout is [cell][101]
//          (daily.h G_Elevation / G_SnowInElevation, VariableChannelGrid grid.h:471-474).
// Device cell order = routing order (rank from G_ROUT_ORDER), rows padded to a multiple of 32.
// Names equal the record names of the test fixtures dumped from the compiled reference.
#pragma once

#define WGK_FIELDS(X) \
    /* ---- shared statics: geometry and masks (geo.h) ---- */ \
    X(area, double, "f64", CELL, 1) \
    X(contfreq, double, "f64", CELL, 1) \
    X(contcell, int16_t, "i16", CELL, 1) \
    X(row, int16_t, "i16", CELL, 1) \
    X(toBeCalculated, int16_t, "i16", CELL, 1) \
    /* ---- shared statics of the vertical balance ---- */ \
    X(landcover, int8_t, "i8", CELL, 1) \
    X(builtup, float, "f32", CELL, 1) \
    X(arid, int16_t, "i16", CELL, 1) \
    X(ldd, int8_t, "i8", CELL, 1) \
    X(texture, int8_t, "i8", CELL, 1) \
    X(elevation, int16_t, "i16", CELL, 101) \
    /* ---- shared statics of the routing cascade (routing.h) ---- */ \
    X(loc_lake, double, "f64", CELL, 1) \
    X(loc_wetland, double, "f64", CELL, 1) \
    X(glo_wetland, double, "f64", CELL, 1) \
    X(lake_area, double, "f64", CELL, 1) \
    X(reservoir_area, double, "f64", CELL, 1) \
    X(stor_cap, double, "f64", CELL, 1) \
    X(mean_outflow, double, "f64", CELL, 1) \
    X(mean_demand, double, "f64", CELL, 1) \
    X(res_type, int8_t, "i8", CELL, 1) \
    X(start_month, int8_t, "i8", CELL, 1) \
    X(river_length, double, "f64", CELL, 1) \
    X(river_slope, double, "f64", CELL, 1) \
    X(roughness, double, "f64", CELL, 1) \
    X(river_bottom_width, double, "f64", CELL, 1) \
    X(river_width_bf, double, "f64", CELL, 1) \
    X(river_storage_max, double, "f64", CELL, 1) \
    X(fswb_init, double, "f64", CELL, 1) \
    X(f_glo_lake, double, "f64", CELL, 1) \
    /* ---- per parameter set ---- */ \
    X(smax, float, "f32", PSET, 1) \
    X(gwfactor, float, "f32", PSET, 1) \
    X(rgmax, int16_t, "i16", PSET, 1) \
    X(laimax, float, "f32", PSET, 1) \
    X(gamma_hbv, double, "f64", PSET, 1) \
    X(cfa, double, "f64", PSET, 1) \
    X(cfs, double, "f64", PSET, 1) \
    X(lake_depth_active, double, "f64", PSET, 1) \
    X(wetl_depth_active, double, "f64", PSET, 1) \
    X(p_prec, double, "f64", PSET, 1) \
    X(p_ptc_hum, double, "f64", PSET, 1) \
    X(p_ptc_ari, double, "f64", PSET, 1) \
    X(p_pet_mxdy, double, "f64", PSET, 1) \
    X(p_netrad, double, "f64", PSET, 1) \
    X(p_snowfz, double, "f64", PSET, 1) \
    X(p_snowmt, double, "f64", PSET, 1) \
    X(p_gradnt, double, "f64", PSET, 1) \
    X(p_degday, double, "f64", PSET, 1) \
    X(p_mcwh, double, "f64", PSET, 1) \
    X(p_pcrit, double, "f64", PSET, 1) \
    X(p_gwoutf, double, "f64", PSET, 1) \
    X(p_evaredex, double, "f64", PSET, 1) \
    X(p_swoutf, double, "f64", PSET, 1) \
    X(p_rivrgh, double, "f64", PSET, 1) \
    /* ---- member state ---- */ \
    X(canopy, double, "f64", MEMBER, 1) \
    X(soil, double, "f64", MEMBER, 1) \
    X(snow, double, "f64", MEMBER, 1) \
    X(snow_bands, double, "f64", MEMBER, 101) \
    X(lai_days, int32_t, "i32", MEMBER, 1) \
    X(lai_status, int32_t, "i32", MEMBER, 1) \
    X(lai_precsum, double, "f64", MEMBER, 1) \
    X(gw, double, "f64", MEMBER, 1) \
    X(loc_lake_stor, double, "f64", MEMBER, 1) \
    X(loc_wetl_stor, double, "f64", MEMBER, 1) \
    X(glo_lake_stor, double, "f64", MEMBER, 1) \
    X(glo_wetl_stor, double, "f64", MEMBER, 1) \
    X(res_stor, double, "f64", MEMBER, 1) \
    X(river_stor, double, "f64", MEMBER, 1) \
    X(red_loc_lake, double, "f64", MEMBER, 1) \
    X(red_loc_wetl, double, "f64", MEMBER, 1) \
    X(red_glo_lake, double, "f64", MEMBER, 1) \
    X(red_glo_wetl, double, "f64", MEMBER, 1) \
    X(red_res, double, "f64", MEMBER, 1) \
    X(red_river, double, "f64", MEMBER, 1) \
    X(k_release, double, "f64", MEMBER, 1) \
    X(land_area_frac, double, "f64", MEMBER, 1) \
    X(land_area_frac_prev, double, "f64", MEMBER, 1) \
    X(land_area_frac_next, double, "f64", MEMBER, 1) \
    X(fswb_laf, double, "f64", MEMBER, 1) \
    X(fswb_laf_next, double, "f64", MEMBER, 1) \
    X(river_area_frac_next, double, "f64", MEMBER, 1) \
    X(river_area_frac_change, double, "f64", MEMBER, 1) \
    X(status_laf_next, int16_t, "i16", MEMBER, 1) \
    /* ---- member daily fluxes (vertical -> routing hand-off and outputs) ---- */ \
    X(lake_balance, double, "f64", MEMBER, 1) \
    X(openwater_prec, double, "f64", MEMBER, 1) \
    X(openwater_pet, double, "f64", MEMBER, 1) \
    X(surface_runoff, double, "f64", MEMBER, 1) \
    X(gw_recharge, double, "f64", MEMBER, 1) \
    X(storage_transfer, double, "f64", MEMBER, 1) \
    X(land_aet, double, "f64", MEMBER, 1) \
    X(land_aet_uncorr, double, "f64", MEMBER, 1) \
    X(discharge, double, "f64", MEMBER, 1) \
    X(river_evapo, double, "f64", MEMBER, 1) \
    X(gwr_swb, double, "f64", MEMBER, 1) \
    X(cell_runoff, double, "f64", MEMBER, 1) \
    /* ---- routing scratch between the cell-parallel pre-pass and the level sweep ---- */ \
    X(t_inflow_local, double, "f64", MEMBER, 1) \
    X(t_runoff_to_river, double, "f64", MEMBER, 1) \
    X(t_gw_to_river, double, "f64", MEMBER, 1) \
    X(t_gwr_loclak, double, "f64", MEMBER, 1) \
    X(t_gwr_locwet, double, "f64", MEMBER, 1) \
    X(t_river_precip, double, "f64", MEMBER, 1) \
    X(t_river_evapo, double, "f64", MEMBER, 1) \
    /* ---- derived once from the statics (k_derive_static) ---- */ \
    X(s_c1, double, "f64", PSET, 1) \
    X(s_slope_pow, double, "f64", PSET, 1) \
    X(s_ekg, double, "f64", PSET, 1) \
    X(s_invkg, double, "f64", PSET, 1) \
    X(s_eks, double, "f64", PSET, 1) \
    X(s_invks, double, "f64", PSET, 1) \
    X(s_flags, int8_t, "i8", CELL, 1) \
    X(s_elev32, int32_t, "i32", CELL, 101) \
    X(s_delev, int16_t, "i16", CELL, 101) \
    X(s_de_min, int16_t, "i16", CELL, 1) \
    X(s_de_max, int16_t, "i16", CELL, 1) \
    /* ---- monthly sums of the seven routing compartments in mm over the continental area (the daily values the \
       reference writes to WghmStateFile, routing.cpp:5002-5020), band-major [7][cell]; only while enabled ---- */ \
    X(mon_acc, double, "f64", MEMBER, 7) \
    /* ---- surface-water-body fractions as formed BEFORE the river-area correction (G_fLocLake / G_fLocWet / G_fGloWet at \
       routing.cpp:5044-5070: columns 45, 47, 46 of the additionalOutIn checkpoint); written only while the monthly accumulation is on ---- */ \
    X(f_loc_lake, double, "f64", MEMBER, 1) \
    X(f_loc_wet, double, "f64", MEMBER, 1) \
    X(f_glo_wet, double, "f64", MEMBER, 1) \
    /* ---- water use (allocated only with wgk_options.subtract_use > 0).  Inputs of the current month in km3 per \
       day: net abstraction from surface water / groundwater times the cell's multiplier (dailyNUInit routing.cpp:884-977, \
       calcNextDay_M :7432-7440), irrigation withdrawal / consumptive use from surface water (:3907-3908); statics: fraction of \
       the irrigation return flow that reaches groundwater, allocation coefficients of irrigation reservoirs [cell][5] \
       (G_ALLOC_COEFF.5.UNF0); state per member: the bookkeeping of updateNetAbstractionGW (:5503-5572) and of :3889-3908 ---- */ \
    X(wu_nus_month, double, "f64", PSET, 1) \
    X(wu_nug_month, double, "f64", PSET, 1) \
    X(wu_wusi_month, double, "f64", CELL, 1) \
    X(wu_cusi_month, double, "f64", CELL, 1) \
    X(wu_frgi, double, "f64", CELL, 1) \
    X(wu_alloc_coeff, double, "f64", CELL, 5) \
    X(wu_total_unsatisfied, double, "f64", MEMBER, 1) \
    X(wu_daily_remaining, double, "f64", MEMBER, 1) \
    X(wu_uns_irr, double, "f64", MEMBER, 1) \
    X(wu_uns_oth, double, "f64", MEMBER, 1) \
    X(wu_red_rf, double, "f64", MEMBER, 1) \
    X(wu_wusi, double, "f64", MEMBER, 1) \
    X(wu_cusi, double, "f64", MEMBER, 1) \
    X(wu_actual_use, double, "f64", MEMBER, 1) \
    X(wu_daily_nug, double, "f64", MEMBER, 1) \
    /* ---- derived from the member state (k_derive_member), maintained by the vertical kernel ---- */ \
    X(s_snowfree, int8_t, "i8", MEMBER, 1) \
    /* ---- land cover tables (LCT_22.DAT / LAI_22.DAT; daily.h:204-208, lai.h) ---- */ \
    X(lai_factor_a, float, "f32", TABLE, 1) \
    X(lai_factor_b, float, "f32", TABLE, 1) \
    X(lai_initial_days, int16_t, "i16", TABLE, 1) \
    X(lai_kc_min, double, "f64", TABLE, 1) \
    X(lai_kc_max, double, "f64", TABLE, 1) \
    X(lct_albedo, double, "f64", TABLE, 1) \
    X(lct_albedo_snow, double, "f64", TABLE, 1) \
    X(lct_ddf, double, "f64", TABLE, 1) \
    X(lct_emissivity, double, "f64", TABLE, 1)

enum wgk_scope { WGK_SCOPE_CELL = 0, WGK_SCOPE_PSET = 1, WGK_SCOPE_MEMBER = 2, WGK_SCOPE_TABLE = 3 };

enum wgk_field_enum {
#define X(name, ctype, dt, scope, bands) WGK_F_##name,
    WGK_FIELDS(X)
#undef X
    WGK_F_COUNT,
    WGK_F_params = 1000 /* pseudo field: f64 [26][ncell] per pset, scattered into the p_* arrays */
};

// Device-side view: one typed pointer per field (base of [index 0]).
struct WgkArrays {
#define X(name, ctype, dt, scope, bands) ctype *name;
    WGK_FIELDS(X)
#undef X
};
