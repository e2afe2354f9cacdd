// wg_model.h — host-side look-alikes of the reference's hot-path classes, backed by the C ABI.
//
//   dailyWaterBalanceClass   daily.h:17-228     calcNewDay (per cell, :24), init, annualInit,
//                                               setStoragesToZero, setStorages, public grids (:45-122)
//   routingClass             routing.h:22-595   routing (:44), updateLandAreaFrac (:46), init,
//                                               annualInit, setStoragesToZero, setLakeWetlToMaximum,
//                                               initFractionStatus, initLake/WetlDepthActive,
//                                               getLandAreaFrac (:246), public grids (:109-259)
//   calibParamClass          calib_param.h:72-222  getValue(eCalibParam, n) over a flat [26][ncell]
//                                               block instead of a JSON DOM
//   optionClass / ConfigFile option.h, configFile.h (the values the hot path reads)
//
// Method names, argument order and meaning follow the reference so that a caller written for
// the reference (integrateWGHM.cpp:127-917) compiles against these classes with the cell count
// turned into a run-time value.  The per-cell calcNewDay() is kept as a shim: the first call
// of a day launches the whole-grid kernel, the others are no-ops; calcNewDayAll() is the
// batched form.  Host grids are mirrors of device arrays, synchronised lazily (pull()/push()).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../../include/wgk.h"
#include "wg_grid.h"
#include "wg_rout_prepare.h"
#include "wg_state_files.h"

namespace wg {

enum eCalibParam {  // calib_param.h:72-101
    P_GAMRUN_C, P_CFA, P_CFS, M_ROOT_D, M_RIVRGH_C, P_LAK_D, P_WET_D, P_SWOUTF_C, M_EVAREDEX, M_NETRAD, P_PTC_HUM,
    P_PTC_ARI, P_PET_MXDY, P_MCWH, M_LAI, P_T_SNOWFZ, P_T_SNOWMT, M_DEGDAY_F, P_T_GRADNT, M_GW_F, M_RG_MAX, P_PCRITGWA,
    P_GWOUTF_C, M_NETABSSW, M_NETABSGW, M_PREC
};

extern const char *const kParamNames[26];  // calib_param.cpp:140-171, eCalibParam order

class calibParamClass {
  public:
    void readJson(const std::string &file, int ncell);  // calib_param.cpp:185-284 (key names :140-171)
    inline double getValue(eCalibParam p, int n) const { return v_[(size_t)p * ncell_ + n]; }
    const double *block() const { return v_.data(); }  // [26][ncell]
    int ncell() const { return ncell_; }

  private:
    int ncell_ = 0;
    std::vector<double> v_;
};

struct ConfigFile {  // configFile.cpp:20-231
    explicit ConfigFile(const std::string &file);
    ConfigFile(const std::string &file, int year, int month, const std::string &progName);  // configFile.cpp: "OL" keeps the file's dates
    std::string startvaluefile, parameterfile, snowInElevationfile, additionalfile, outputmeanfile, outputlastdayfile,
        outputsnowlastdayfile, outputadditionalfile, runtimeoptionsfile, outputoptionsfile, routingoptionsfile, stationsfile,
        inputDir, outputDir, climateDir, routingDir, waterUseDir, calibrationfile, outputparameter;
    int startMonth = 0, startYear = 0, endMonth = 0, endYear = 0, timeStep = 0, numInitYears = 0;
};

struct optionClass {  // option.cpp:173-620, OPTIONS.DAT order
    void init(const ConfigFile &cfg);
    int v[36] = {0};
    int &fileEndianType = v[0], &basin = v[1], &grid_store = v[2], &time_series = v[5], &cloud = v[6], &intercept = v[7],
        &calc_albedo = v[8], &petOpt = v[9], &use_kc = v[10], &rout_prepare = v[12], &riverveloOpt = v[14],
        &subtract_use = v[15], &use_alloc = v[16], &delayedUseSatisfaction = v[17], &clclOpt = v[18], &permaOpt = v[19], &resOpt = v[20], &statcorrOpt = v[21],
        &aridareaOpt = v[22], &fractionalRoutingOpt = v[23], &riverEvapoOpt = v[24], &aggrNUsGloLakResOpt = v[25], &resYearOpt = v[27],
        &resYearReference = v[28], &resYearFirstToUse = v[29], &resYearLastToUse = v[30], &antNatOpt = v[31], &calc_wtemp = v[34], &glacierOpt = v[35];
    std::string input_dir, output_dir, climate_dir, routing_dir, water_use_dir;
    int start_year = 0, end_year = 0;
    void require_canonical() const;  // throws for option values outside the implemented hot path
};

struct geoClass {  // geo.h / geo.cpp:7-45
    void init(const std::string &input_dir, int ncell, int resOpt);
    Grid<int16_t> G_row, G_col, G_contcell;
    Grid<double> G_contfreq;
    std::vector<double> area;  // per row
    double areaOfCellByArrayPos(int n) const { return area[G_row[n] - 1]; }
};

class Engine;  // owns the wgk context and the host<->device mirrors

class dailyWaterBalanceClass {
  public:
    explicit dailyWaterBalanceClass(Engine &e) : eng(e) {}
    void init(const std::string &input_dir, short nBasins);   // daily.cpp:1544-1682 (LCT_22.DAT, parameters)
    void annualInit() {}                                      // only output bookkeeping in the reference (:1684-1740)
    void setStoragesToZero();                                 // daily.cpp:1882-1895
    void setStorages(WghmStateFile &, SnowInElevationFile &, AdditionalOutputInputFile &);  // :1896-1924
    // per-cell signature of daily.h:24; whole-grid launch on the first continental cell of a day
    void calcNewDay(short day, short month, short day_in_month, short last_day_in_month, short year, int n, WghmStateFile &,
                    AdditionalOutputInputFile &, SnowInElevationFile &, short readinstatus, calibParamClass &);
    void calcNewDayAll(short day, short month, short day_in_month);
    Grid<double> G_gammaHBV, G_cellCorrFact, G_snow, G_soilWaterContent, G_canopyWaterContent, G_lakeBalance, G_openWaterPET,
        G_openWaterPrec, G_dailyLocalSurfaceRunoff, G_dailyGwRecharge, G_dailyStorageTransfer;
    Grid<int16_t, 101> G_Elevation;
    Grid<double, 101> G_SnowInElevation;
    float rootingDepth_lct[18];
    double albedo_lct[18], albedoSnow_lct[18], ddf_lct[18], emissivity_lct[18];
    void pull();  // device -> host mirrors of the state and flux grids

  private:
    Engine &eng;
    int last_day_launched = -1;
};

class routingClass {
  public:
    explicit routingClass(Engine &e) : eng(e) {}
    const double minStorVol = 1.e-15;
    void init(short nBasins, const ConfigFile &cfg, WghmStateFile &, AdditionalOutputInputFile &);  // routing.cpp:131-742 (K_release :388-392)
    void annualInit(short year, int start_month, AdditionalOutputInputFile &);  // :979-1495
    void pushYearly();  // statics / storages annualInit changed (reservoirs coming on line, resYearOpt 1) -> device
    void initLakeDepthActive(const calibParamClass &);                        // :5613-5619
    void initWetlDepthActive(const calibParamClass &);                        // :5621-5628
    void setStoragesToZero();                                                 // :789-847
    void setStorages(WghmStateFile &, AdditionalOutputInputFile &);           // :851-882
    void initFractionStatus();                                                // :745-765
    void initFractionStatusAdditionalOI(AdditionalOutputInputFile &);         // :767-787 (restart: fractions from the checkpoint)
    void update_landarea_red_fac_PDAF(calibParamClass &, AdditionalOutputInputFile &);  // :5354-5475 (restart: first day after a checkpoint)
    void setLakeWetlToMaximum(short start_year);                              // :5647-5720
    void routing(short year, short day, short month, short day_in_month, short last_day_in_month, WghmStateFile &,
                 AdditionalOutputInputFile &, short readinstatus, calibParamClass &);  // :1629
    void updateLandAreaFrac(AdditionalOutputInputFile &);                     // :5343-5352
    double getLandAreaFrac(int n) {                                           // routing.h:246-251
        return (0 == statusStarted_landAreaFracNextTimestep[n]) ? G_landAreaFrac[n] : G_landAreaFracNextTimestep[n];
    }
    void pull();
    void updateGloResPrevYear_pct() { G_glores_prevyear = G_glo_res; }
    // water use (subtract_use 2): the year's net abstractions (routing.cpp:884-977), the month's daily values for the
    // device (calcNextDay_M :7432-7440 is the same value on every day of a month), the year-end bookkeeping (:5246-5306)
    void dailyNUInit(const std::string &input_directory, short new_year, calibParamClass &);
    void pushWaterUseMonth(short month);
    void pullWaterUse();
    void pushWaterUseState();
    void annualWaterUsePostProcessing(short year, AdditionalOutputInputFile &);
    Grid<double, 12> G_dailyNUs, G_dailyNUg, G_monthlyWUIrrigFromSwb, G_monthlyCUIrrigFromSwb;
    Grid<double, 5> G_alloc_coeff;
    Grid<double> G_fractreturngw_irrig, G_totalUnsatisfiedUse, G_UnsatisfiedUsePrevYear, G_unsatisfiedNAsFromIrrig,
        G_unsatisfiedNAsFromIrrigPrevYear, G_unsatisfiedNAsFromOtherSectors, G_unsatisfiedNAsFromOtherSectorsPrevYear, G_reducedReturnFlow,
        G_reducedReturnFlowPrevYear, G_dailyRemainingUse, G_withdrawalIrrigFromSwb, G_consumptiveUseIrrigFromSwb, G_actualUse;
    Grid<double> G_statCorrFact, G_landAreaFrac, G_landAreaFracNextTimestep, G_landAreaFracPrevTimestep, G_locLakeStorage,
        G_locWetlStorage, G_gloLakeStorage, G_gloWetlStorage, G_gloResStorage, G_riverStorage, G_groundwaterStorage,
        G_locLakeAreaReductionFactor, G_locWetlAreaReductionFactor, G_gloLakeEvapoReductionFactor,
        G_gloWetlAreaReductionFactor, G_gloResEvapoReductionFactor, G_riverAreaReductionFactor, G_glo_lake, G_loc_lake,
        G_loc_res, G_glo_wetland, G_loc_wetland, G_reg_lake, G_glo_res, G_glores_prevyear, G_lake_area, G_reservoir_area,
        G_reservoir_area_full, G_stor_cap, G_stor_cap_full, G_mean_outflow, G_mean_demand, G_riverLength, G_RiverSlope,
        G_Roughness, G_bankfull_flow, G_RiverWidth_bf, G_RiverDepth_bf, G_riverBottomWidth, G_riverStorageMax,
        G_lakeDepthActive, G_wetlDepthActive, G_fswbInit, G_fswbLandAreaFrac, G_fswbLandAreaFracNextTimestep, G_fGloLake, G_fLocLake,
        G_fLocWet, G_fGloWet,
        G_riverAreaFracNextTimestep_Frac, K_release, G_riverDischarge;
    Grid<int16_t> statusStarted_landAreaFracNextTimestep;
    Grid<int8_t> G_res_type, G_start_month, G_reg_lake_status, G_LDD;
    Grid<int32_t> G_res_start_year, G_downstreamCell, G_routOrder, G_outflow_cell_assignment;
    bool yearlyChanged = false;  // the last annualInit changed reservoir statics / storages the device holds (resYearOpt 1)
    short statusStarted_updateGloResPrevYear = 0;
    std::vector<short> statusStarted_landfreq;

  private:
    Engine &eng;
    std::vector<double> day_state;  // [7][ncell] of the last routed day
};

// Everything the reference keeps in process globals (globals.cpp:7-33), per model instance.
class Engine {
  public:
    Engine(int ncell, int device = 0, int restart = 0, int subtract_use = 0);  // device < 0: host-side initialisation only (no context; tests of the init logic)
    ~Engine();
    int ncell;
    wgk_ctx *ctx = nullptr;
    optionClass options;
    geoClass geo;
    calibParamClass calParam;
    dailyWaterBalanceClass dailyWaterBalance;
    routingClass routing;
    FlowTopology topo;
    Grid<int16_t> G_aindex, G_toBeCalculated;
    Grid<int8_t> G_landCover, G_texture;
    Grid<float> G_built_up, G_Smax, G_gwFactor, G_LAImax;
    Grid<int16_t> G_Rgmax;
    float lai_factor_a[18], lai_factor_b[18];
    int16_t lai_initialDays[18];
    double kc_min[18], kc_max[18];
    Grid<int32_t> lai_days, lai_status;
    Grid<double> lai_precsum;
    // init-time derivations
    void land_init();                  // land.cpp:21-33
    void lai_init(AdditionalOutputInputFile &);  // lai.cpp:40-148 (restart: growing-season state from the checkpoint, :53-65)
    void createMaxSoilWaterCapacityGrid();  // s_max.cpp:40-75
    void createGroundwaterGrids();     // gw_frac.cpp:36-275
    // device synchronisation
    void push_static();                // topology + statics + parameter-derived arrays
    void push_state();                 // host state grids -> device
    void set_forcing_month(int month1, int year);  // climate.cpp:93-138 (.31 files)
    void set_forcing_year(int year);               // climateYear.cpp:38-58 (.365 files, time_series 1)
    void check(int rc, const char *what);
    template <class T, int C> void set(const char *name, const Grid<T, C> &g, int index = 0);
    template <class T, int C> void get(const char *name, Grid<T, C> &g, int index = 0);
};

// integrate_wghm_-shaped driver (integrateWGHM.cpp:38-1168, open-loop "OL" mode, canonical
// options): config.txt in, txt state files out.  Returns simulated days.
long integrate_wghm(const std::string &config_file, int ncell, int device, double *seconds_day_loop);
void init_dump(const std::string &config_file, int ncell, const std::string &dump_file);

// the three state objects of a run (what initialize_wghm_ allocates and integrate_wghm_ works on)
struct ModelState {
    explicit ModelState(int ncell) : wghmState(ncell, 1), additionalOutIn(ncell), snow_in_elevation(ncell) {}
    WghmStateFile wghmState;
    AdditionalOutputInputFile additionalOutIn;
    SnowInElevationFile snow_in_elevation;
};
struct ModelStateRef {
    WghmStateFile &wghmState;
    AdditionalOutputInputFile &additionalOutIn;
    SnowInElevationFile &snow_in_elevation;
    ModelStateRef(WghmStateFile &a, AdditionalOutputInputFile &b, SnowInElevationFile &c) : wghmState(a), additionalOutIn(b), snow_in_elevation(c) {}
    ModelStateRef(ModelState &s) : wghmState(s.wghmState), additionalOutIn(s.additionalOutIn), snow_in_elevation(s.snow_in_elevation) {}
};
// initialize_wghm (initializeWGHM.cpp:32-72): start values and parameters from the files named in the config
void load_start_state(const ConfigFile &cfg, int ncell, ModelState &S, calibParamClass &calParam);
// everything integrate_wghm_ does before its year loop (integrateWGHM.cpp:127-476), on the host grids of E: cold start or restart
// from checkpoint objects (PDAF monthly cycle); E.calParam must be set
void initialize_model(Engine &E, const ConfigFile &cfg, ModelStateRef S);
// initialisation + the year / month / day loop; returns the simulated days
long run_model(Engine &E, const ConfigFile &cfg, ModelStateRef S, double *seconds_day_loop);

}  // namespace wg

// B1, the reference's own FFI (initializeWGHM.h:11-15, integrateWGHM.h:9-13): same names, argument order, by-reference pointers and
// ownership, so that the Fortran (PDAF) side links unchanged
extern "C" {
void initialize_wghm_(const char *s, wg::WghmStateFile *&initstate, wg::calibParamClass *&initcal, wg::AdditionalOutputInputFile *&initaddio,
                      wg::SnowInElevationFile *&initsnow, long *year, long *month, const char *s2, const char *s3, wg::WghmStateFile *&wghmMean);
void integrate_wghm_(const char *s, wg::ConfigFile *&configFile, wg::WghmStateFile *&wghmState, wg::calibParamClass *&calParam,
                     wg::AdditionalOutputInputFile *&additionalOutIn, wg::SnowInElevationFile *&snow_in_elevation, long *step, long *total_steps,
                     long *year, long *month, const char *s2);
}
void initialize_wghm(const char *s, wg::WghmStateFile *&initstate, wg::calibParamClass *&initcal, wg::AdditionalOutputInputFile *&initaddio,
                     wg::SnowInElevationFile *&initsnow, long *year, long *month, const char *s2, const char *s3, wg::WghmStateFile *&wghmMean);
void integrate_wghm(const char *s, wg::ConfigFile *&configFile, wg::WghmStateFile *&wghmState, wg::calibParamClass *&calParam,
                    wg::AdditionalOutputInputFile *&additionalOutIn, wg::SnowInElevationFile *&snow_in_elevation, long *step, long *total_steps,
                    long *year, long *month, const char *s2);

extern "C" {
// C entry used by tests / other languages: runs the drop-in driver on a reference-format config
long wg_host_integrate(const char *config_file, int ncell, int device, double *seconds_day_loop, char *err, size_t errlen);
int wg_host_state_roundtrip(const char *kind, const char *in, const char *out, int ncell, char *err, size_t errlen);
// flow topology from the reference-format input directory into the routing directory
int wg_host_prepare_routing_files(const char *input_dir, const char *routing_dir, int resOpt, int ncell, int *nlevels, char *err, size_t errlen);
// host-side initialisation only (no GPU needed): the start state integrate_wghm would push to the device, written as a record dump
// (name[32], day i32 = 0, dtype[8], count i64, data) with the record names of the test fixtures
// ready-to-step context built by the host layer from a reference-format configuration (the caller owns it: wgk_destroy)
void *wg_host_create_context(const char *config_file, int ncell, int device, char *err, size_t errlen);
int wg_host_init_dump(const char *config_file, int ncell, const char *dump_file, char *err, size_t errlen);
}
