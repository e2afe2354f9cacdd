/* wg_oracle.h — CPU restatement of the WaterGAP2 daily hot path (TEST INFRASTRUCTURE).
 *
 * This is the oracle of the repository: a plain-C, scalar, reference-ordered restatement of
 *   - rout_prepare.cpp (flow topology, routing order),
 *   - dailyWaterBalanceClass::calcNewDay (daily.cpp:94-1264) + lai.cpp:152-306,
 *   - routingClass::routing (routing.cpp:1629-5244) + updateLandAreaFrac (:5343-5352)
 * under the canonical option vector of SURVEY.md 8(d).  It is pinned BIT-EXACTLY against the
 * reference's own sources compiled in oracle/_ref (see tests/test_oracle_vs_ref.py and the
 * golden vectors in tests/golden/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may link or call it; the product (watergap2_b200/) never does.
 */
#ifndef WG_ORACLE_H
#define WG_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WGO_NLCT 18
#define WGO_NPARAM 26
#define WGO_NBAND 101

/* parameter indices = eCalibParam order (calib_param.h:72-101) */
enum {
    WGO_P_GAMRUN_C, WGO_P_CFA, WGO_P_CFS, WGO_M_ROOT_D, WGO_M_RIVRGH_C, WGO_P_LAK_D, WGO_P_WET_D, WGO_P_SWOUTF_C,
    WGO_M_EVAREDEX, WGO_M_NETRAD, WGO_P_PTC_HUM, WGO_P_PTC_ARI, WGO_P_PET_MXDY, WGO_P_MCWH, WGO_M_LAI,
    WGO_P_T_SNOWFZ, WGO_P_T_SNOWMT, WGO_M_DEGDAY_F, WGO_P_T_GRADNT, WGO_M_GW_F, WGO_M_RG_MAX, WGO_P_PCRITGWA,
    WGO_P_GWOUTF_C, WGO_M_NETABSSW, WGO_M_NETABSGW, WGO_M_PREC
};

/* Field table: X(name, ctype, dtype-string, elements per cell (0 = fixed NLCT table)) .
 * Names equal the record names of the reference dump (oracle/ref_harness.cpp). */
#define WGO_FIELDS(X) \
    /* geometry / masks */ \
    X(area, double, "f64", 1) X(contfreq, double, "f64", 1) X(contcell, int16_t, "i16", 1) \
    X(row, int16_t, "i16", 1) X(toBeCalculated, int16_t, "i16", 1) \
    /* vertical statics */ \
    X(landcover, int8_t, "i8", 1) X(builtup, float, "f32", 1) X(arid, int16_t, "i16", 1) X(ldd, int8_t, "i8", 1) \
    X(elevation, int16_t, "i16", WGO_NBAND) X(smax, float, "f32", 1) X(gwfactor, float, "f32", 1) \
    X(rgmax, int16_t, "i16", 1) X(texture, int8_t, "i8", 1) X(laimax, float, "f32", 1) \
    X(gamma_hbv, double, "f64", 1) X(cfa, double, "f64", 1) X(cfs, double, "f64", 1) \
    X(params, double, "f64", WGO_NPARAM) \
    X(lai_factor_a, float, "f32", 0) X(lai_factor_b, float, "f32", 0) X(lai_initial_days, int16_t, "i16", 0) \
    X(lai_kc_min, double, "f64", 0) X(lai_kc_max, double, "f64", 0) X(lct_albedo, double, "f64", 0) \
    X(lct_albedo_snow, double, "f64", 0) X(lct_ddf, double, "f64", 0) X(lct_emissivity, double, "f64", 0) \
    /* routing statics */ \
    X(loc_lake, double, "f64", 1) X(loc_wetland, double, "f64", 1) X(glo_wetland, double, "f64", 1) \
    X(lake_area, double, "f64", 1) X(reservoir_area, double, "f64", 1) X(stor_cap, double, "f64", 1) \
    X(mean_outflow, double, "f64", 1) X(mean_demand, double, "f64", 1) X(res_type, int8_t, "i8", 1) \
    X(start_month, int8_t, "i8", 1) X(river_length, double, "f64", 1) X(river_slope, double, "f64", 1) \
    X(roughness, double, "f64", 1) X(river_bottom_width, double, "f64", 1) X(river_width_bf, double, "f64", 1) \
    X(river_storage_max, double, "f64", 1) X(lake_depth_active, double, "f64", 1) \
    X(wetl_depth_active, double, "f64", 1) X(downstream_cell, int32_t, "i32", 1) X(routing_cell, int32_t, "i32", 1) \
    X(fswb_init, double, "f64", 1) X(f_glo_lake, double, "f64", 1) \
    /* forcing of the current month, reference layout [cell][31] (climate.h:14-21) */ \
    X(prec31, float, "f32", 31) X(temp31, float, "f32", 31) X(sw31, float, "f32", 31) X(lw31, float, "f32", 31) \
    /* state */ \
    X(canopy, double, "f64", 1) X(soil, double, "f64", 1) X(snow, double, "f64", 1) \
    X(snow_bands, double, "f64", WGO_NBAND) X(lai_days, int32_t, "i32", 1) X(lai_status, int32_t, "i32", 1) \
    X(lai_precsum, double, "f64", 1) X(gw, double, "f64", 1) X(loc_lake_stor, double, "f64", 1) \
    X(loc_wetl_stor, double, "f64", 1) X(glo_lake_stor, double, "f64", 1) X(glo_wetl_stor, double, "f64", 1) \
    X(res_stor, double, "f64", 1) X(river_stor, double, "f64", 1) X(red_loc_lake, double, "f64", 1) \
    X(red_loc_wetl, double, "f64", 1) X(red_glo_lake, double, "f64", 1) X(red_glo_wetl, double, "f64", 1) \
    X(red_res, double, "f64", 1) X(red_river, double, "f64", 1) X(k_release, double, "f64", 1) \
    X(land_area_frac, double, "f64", 1) X(land_area_frac_prev, double, "f64", 1) \
    X(land_area_frac_next, double, "f64", 1) X(fswb_laf, double, "f64", 1) X(fswb_laf_next, double, "f64", 1) \
    X(river_area_frac_next, double, "f64", 1) X(status_laf_next, int16_t, "i16", 1) \
    /* daily fluxes / diagnostics (outputs) */ \
    X(lake_balance, double, "f64", 1) X(openwater_prec, double, "f64", 1) X(openwater_pet, double, "f64", 1) \
    X(surface_runoff, double, "f64", 1) X(gw_recharge, double, "f64", 1) X(storage_transfer, double, "f64", 1) \
    X(land_aet, double, "f64", 1) X(land_aet_uncorr, double, "f64", 1) X(discharge, double, "f64", 1) \
    X(river_evapo, double, "f64", 1) X(gwr_swb, double, "f64", 1) X(cell_runoff, double, "f64", 1) \
    X(river_inflow, double, "f64", 1) X(river_area_frac, double, "f64", 1) \
    X(river_area_frac_change, double, "f64", 1) X(thresh_elev, int32_t, "i32", 1) \
    X(wghm_routing_mm, double, "f64", 7) \
    /* ---- water use (subtract_use 2, use_alloc 0, delayedUseSatisfaction 0, aggrNUsGloLakResOpt 0) ---- \
       inputs of the current MONTH in km3/day (dailyNUInit routing.cpp:884-977 + calcNextDay_M :7432-7440; the irrigation \
       withdrawal / consumptive use from surface water :3907-3908), statics, and the per-cell state */ \
    X(wu_nus_month, double, "f64", 1) X(wu_nug_month, double, "f64", 1) X(wu_wusi_month, double, "f64", 1) \
    X(wu_cusi_month, double, "f64", 1) X(wu_frgi, double, "f64", 1) X(wu_alloc_coeff, double, "f64", 5) \
    X(wu_daily_nus, double, "f64", 1) X(wu_daily_nug, double, "f64", 1) X(wu_total_unsatisfied, double, "f64", 1) \
    X(wu_daily_remaining, double, "f64", 1) X(wu_uns_irr, double, "f64", 1) X(wu_uns_oth, double, "f64", 1) \
    X(wu_red_rf, double, "f64", 1) X(wu_wusi, double, "f64", 1) X(wu_cusi, double, "f64", 1) X(wu_actual_use, double, "f64", 1)

typedef struct wgo_ctx wgo_ctx;

wgo_ctx *wgo_create(int ncell);
void wgo_destroy(wgo_ctx *c);
int wgo_ncell(const wgo_ctx *c);
/* by-name access to the arrays owned by the context; returns NULL for unknown names */
void *wgo_field(wgo_ctx *c, const char *name, const char **dtype, int64_t *count);
/* additionalOutIn.additionalfilestatus (daily.cpp:165): 1 = restart from a checkpoint */
void wgo_set_restart(wgo_ctx *c, int restart);
/* options.subtract_use: 0 no water use (canonical), 2 net abstractions from surface water and groundwater */
void wgo_set_subtract_use(wgo_ctx *c, int subtract_use);

/* one simulated day; day 1..365, month 0..11, day_in_month 1..31 (integrateWGHM.cpp:755-798) */
void wgo_vertical_day(wgo_ctx *c, int day, int month, int day_in_month);
void wgo_routing_day(wgo_ctx *c, int day, int month, int day_in_month);
void wgo_update_land_area_frac(wgo_ctx *c);
/* convenience: vertical + routing + updateLandAreaFrac for `ndays` days starting at day_in_month */
void wgo_step_days(wgo_ctx *c, int day, int month, int day_in_month, int ndays);
/* sum over cells of all ten storage compartments in km3 (mass-balance diagnostic) */
double wgo_total_storage_km3(const wgo_ctx *c);

/* ---- flow topology: rout_prepare.cpp ---------------------------------------------------- */
typedef struct {
    int ncell;
    int ncol, nrow;              /* raster (720 x 360 at 0.5 deg) */
    /* inputs */
    const int16_t *flowdir;      /* Arc codes, G_FLOWDIR.UNF2 */
    const int16_t *row, *col;    /* 1-based, GR/GC.UNF2 */
    const int32_t *gcrc;         /* [ncol][nrow] 1-based cell number or 0, GCRC.UNF4 */
    /* outputs (caller-allocated) */
    int8_t *ldd_2;               /* G_LDD_2.UNF1 (written before loop breaking, rout_prepare.cpp:150) */
    int8_t *ldd;                 /* LDD after loop breaking */
    int32_t *inflow9;            /* [ncell][9] G_INFLC.9.UNF4 */
    int16_t *flow_acc;           /* G_FLOW_ACC.UNF2 */
    uint16_t *basins;            /* G_BASINS.UNF2 */
    uint16_t *basins2;           /* G_BASINS_2.UNF2 */
    uint16_t *cells_to_outlet;   /* G_CELLS_TO_OUTLET.UNF2 */
    int32_t *outflow_cell;       /* G_OUTFLC.UNF4 */
    int32_t *rout_order;         /* G_ROUT_ORDER.UNF4 */
    int32_t *neighbour8;         /* [ncell][8] G_NEIGHBOUR_CELLS.8.UNF4 (may be NULL) */
    int32_t nlevels;             /* number of Kahn sweeps (= dependency depth) */
    int32_t nbasins, nbasins2;
} wgo_topology;
int wgo_rout_prepare(wgo_topology *t);
/* float32 geometry of rout_prepare.cpp:700-832; cell_distance is [9][nrow] (layout of GCELLDIST.9.UNF0) */
void wgo_cell_distances(int nrow, float *cell_distance);
void wgo_river_slope_length(int ncell, int nrow, const float *cell_distance, const float *altitude,
                            const float *meandering, const int32_t *outflow_cell, const int8_t *ldd,
                            const int16_t *row, const int16_t *col, float *slope, float *length);
/* reservoir_prepare (rout_prepare.cpp:888-1031): allocation coefficients [ncell][5] and start month */
void wgo_reservoir_prepare(int ncell, const float *resarea_f32, const float *mean_outflow_f32,
                           const float *mean_outflow12_f32, const int32_t *outflow_cell, float *alloc_coeff5,
                           int8_t *start_month);

#ifdef __cplusplus
}
#endif
#endif
