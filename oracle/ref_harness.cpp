// ref_harness.cpp — TEST INFRASTRUCTURE (oracle), not product code.
//
// Drives the UNMODIFIED reference sources (compiled where they lie under
// /root/reference/source by oracle/build_ref.sh) in two ways:
//
//   ref_harness driver <config.txt>
//       calls the reference's own entry points initialize_wghm / integrate_wghm
//       (initializeWGHM.h:14, integrateWGHM.h:12) exactly like watergap.cpp:122-163 does.
//       The run writes the reference's txt state files (16 significant digits).
//
//   ref_harness replay <config.txt> <dump.wgd> [--days A-B] [--every K] [--snow-days A-B]
//                      [--final-state PREFIX] [--time-only] [--day-times FILE] [--deep-snow]
//       replays the orchestration of integrate_wghm_ (integrateWGHM.cpp:127-309 init
//       sequence, :546-922 year/month/day loops) calling the reference's own
//       dailyWaterBalanceClass::calcNewDay / routingClass::routing /
//       routingClass::updateLandAreaFrac objects, and dumps IN-MEMORY doubles
//       (private members reached with `#define private public`, this TU only)
//       after selected days.  Also times the three phases of the day loop
//       (the CPU baseline of BASELINE.md).
//
// The equality "replay final state == driver final state" is checked by
// tests (tests/test_ref_harness.py) so that the replay is pinned to the
// reference's own driver.
//
// Dump container ("WGD1"): a sequence of records
//   replay ... --enkf DIR
//       after the last simulated month: the reference's own PDAF exchange functions on the month just run -
//       extract_sub_ (extractsub.cpp:17) packs the monthly-mean state vector of the cells listed in DIR/ids.txt
//       minus the mean field DIR/meanfield.bin; the "analysis" is that vector plus DIR/perturb.bin;
//       enkf_wghmstate_ (enKF2wghmState.cpp:17) applies it; then the restore of the next cycle
//       (routingClass::setStorages, dailyWaterBalanceClass::setStorages, integrateWGHM.cpp:303/407).
//       Dumped (day tag 9000): enkf_extract, enkf_field, enkf_prediction, enkf_lastday [n][10], enkf_snow_elev
//       [n][101] and the restored in-memory state.
//   ref_harness calib <dir>
//       the calibration loop of integrate_wghm_ (integrateWGHM.cpp:1091-1116) around the reference's own
//       calibGammaClass (calibration.cpp): init / setRunoff / setUpstInflow / setWaterUse / findNewGamma /
//       writeCorrFactors / writeCalibStatus, with the model run replaced by a closed-form response of the annual
//       station discharge to gamma read from <dir>/CALIB_IN.txt.  Runs inside <dir> (the class writes CALIBRATION.OUT,
//       CALIBRATION.LOG, CALIBSTATUS.OUT, STAT_CORR_FACTOR.OUT and G_CORR_FACTOR.UNF0 into the working directory) and prints one
//       "CALIB ..." line per findNewGamma call.
//   ref_harness wu_unit <in.f64> <out.f64>
//       calls the reference's routingClass::updateNetAbstractionGW (routing.cpp:5503-5572) cell by cell on seeded member
//       arrays: in = [8][ng] doubles (dailyRemainingUse, withdrawalIrrigFromSwb, consumptiveUseIrrigFromSwb, fractreturngw_irrig,
//       unsatisfiedNAsFromIrrig, unsatisfiedNAsFromOtherSectors, reducedReturnFlow, dailydailyNUg), out = [6][ng] (returned NAg,
//       then the five arrays the function updates: dailyRemainingUse, unsatisfiedNAsFromIrrig, unsatisfiedNAsFromOtherSectors,
//       reducedReturnFlow, dailydailyNUg).  Groundwork for SURVEY 8f-4.
//   char name[32]; int32 day (0 = static/initial, k = after k-th simulated day);
//   char dtype[8] ("f64","f32","i32","i16","i8"); int64 count; raw little-endian data.

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <array>
#include <map>
#include <cmath>
#ifdef _OPENMP
#include <omp.h>
#endif

#define private public
#define protected public
#include "def.h"
#include "common.h"
#include "globals.h"
#include "initializeWGHM.h"
#include "integrateWGHM.h"
#include "calib_param.h"
#include "extractsub.h"
#include "enKF2wghmState.h"
#include "calibration.h"
#include <unistd.h>
#undef private
#undef protected

static FILE *g_dump = nullptr;

static void put(const char *name, int day, const char *dtype, int64_t count, const void *data, size_t elsize) {
    if (!g_dump) return;
    char nm[32];
    memset(nm, 0, sizeof nm);
    strncpy(nm, name, 31);
    char dt[8];
    memset(dt, 0, sizeof dt);
    strncpy(dt, dtype, 7);
    int32_t d = day;
    fwrite(nm, 1, 32, g_dump);
    fwrite(&d, 4, 1, g_dump);
    fwrite(dt, 1, 8, g_dump);
    fwrite(&count, 8, 1, g_dump);
    fwrite(data, elsize, (size_t)count, g_dump);
}
template <class G> static void put_f64(const char *name, int day, G &g, int64_t count = ng) {
    put(name, day, "f64", count, g.getDataPointer(), 8);
}
template <class G> static void put_f32(const char *name, int day, G &g, int64_t count = ng) {
    put(name, day, "f32", count, g.getDataPointer(), 4);
}
template <class G> static void put_i32(const char *name, int day, G &g, int64_t count = ng) {
    put(name, day, "i32", count, g.getDataPointer(), 4);
}
template <class G> static void put_i16(const char *name, int day, G &g, int64_t count = ng) {
    put(name, day, "i16", count, g.getDataPointer(), 2);
}
template <class G> static void put_i8(const char *name, int day, G &g, int64_t count = ng) {
    put(name, day, "i8", count, g.getDataPointer(), 1);
}

struct Range { int a = 0, b = -1; bool has(int d) const { return d >= a && d <= b; } };
static Range parse_range(const char *s) {
    Range r;
    if (sscanf(s, "%d-%d", &r.a, &r.b) != 2) { r.a = r.b = atoi(s); }
    return r;
}

static void dump_static(calibParamClass &cal) {
    // geometry / masks
    std::vector<double> area(ng), p(ng);
    for (int n = 0; n < ng; n++) area[n] = geo.areaOfCellByArrayPos(n);
    put("area", 0, "f64", ng, area.data(), 8);
    put_f64("contfreq", 0, geo.G_contfreq);
    put_i16("contcell", 0, geo.G_contcell);
    put_i16("row", 0, geo.G_row);
    put_i16("toBeCalculated", 0, G_toBeCalculated);
    put_i8("landcover", 0, land.G_landCover);
    put_f32("builtup", 0, land.G_built_up);
    put_i16("arid", 0, G_aindex);
    put_i8("ldd", 0, G_LDD);
    put_i16("elevation", 0, dailyWaterBalance.G_Elevation, (int64_t)ng * 101);
    put_f32("smax", 0, maxSoilWaterCap.G_Smax);
    put_f32("gwfactor", 0, GW.G_gwFactor);
    put_i16("rgmax", 0, GW.G_Rgmax);
    put_i8("texture", 0, GW.G_texture);
    put_f32("laimax", 0, lai.G_LAImax);
    put("lai_factor_a", 0, "f32", nlct, lai.lai_factor_a, 4);
    put("lai_factor_b", 0, "f32", nlct, lai.lai_factor_b, 4);
    put("lai_initial_days", 0, "i16", nlct, lai.initialDays, 2);
    put("lai_kc_min", 0, "f64", nlct, lai.kc_min, 8);
    put("lai_kc_max", 0, "f64", nlct, lai.kc_max, 8);
    put("lct_albedo", 0, "f64", nlct, dailyWaterBalance.albedo_lct, sizeof(dailyWaterBalance.albedo_lct[0]));
    put("lct_albedo_snow", 0, "f64", nlct, dailyWaterBalance.albedoSnow_lct, sizeof(dailyWaterBalance.albedoSnow_lct[0]));
    put("lct_ddf", 0, "f64", nlct, dailyWaterBalance.ddf_lct, sizeof(dailyWaterBalance.ddf_lct[0]));
    put("lct_emissivity", 0, "f64", nlct, dailyWaterBalance.emissivity_lct, sizeof(dailyWaterBalance.emissivity_lct[0]));
    put("lct_rooting_depth", 0, "f32", nlct, dailyWaterBalance.rootingDepth_lct, sizeof(dailyWaterBalance.rootingDepth_lct[0]));
    put_f64("gamma_hbv", 0, dailyWaterBalance.G_gammaHBV);
    put_f64("cfa", 0, dailyWaterBalance.G_cellCorrFact);
    put_f64("cfs", 0, routing.G_statCorrFact);
    // the 26 per-cell calibration parameters, [26][ng]
    std::vector<double> par((size_t)26 * ng);
    for (int k = 0; k < 26; k++)
        for (int n = 0; n < ng; n++) par[(size_t)k * ng + n] = cal.getValue((eCalibParam)k, n);
    put("params", 0, "f64", (int64_t)26 * ng, par.data(), 8);
    // routing statics
    put_f64("loc_lake", 0, routing.G_loc_lake);
    put_f64("loc_wetland", 0, routing.G_loc_wetland);
    put_f64("glo_wetland", 0, routing.G_glo_wetland);
    put_f64("glo_lake", 0, routing.G_glo_lake);
    put_f64("glo_res", 0, routing.G_glo_res);
    put_f64("lake_area", 0, routing.G_lake_area);
    put_f64("reservoir_area", 0, routing.G_reservoir_area);
    put_f64("stor_cap", 0, routing.G_stor_cap);
    put_f64("mean_outflow", 0, routing.G_mean_outflow);
    put_f64("mean_demand", 0, routing.G_mean_demand);
    put_i8("res_type", 0, routing.G_res_type);
    put_i8("start_month", 0, routing.G_start_month);
    put_f64("river_length", 0, routing.G_riverLength);
    put_f64("river_slope", 0, routing.G_RiverSlope);
    put_f64("roughness", 0, routing.G_Roughness);
    put_f64("river_bottom_width", 0, routing.G_riverBottomWidth);
    put_f64("river_width_bf", 0, routing.G_RiverWidth_bf);
    put_f64("river_storage_max", 0, routing.G_riverStorageMax);
    put_f64("lake_depth_active", 0, routing.G_lakeDepthActive);
    put_f64("wetl_depth_active", 0, routing.G_wetlDepthActive);
    put_i32("downstream_cell", 0, routing.G_downstreamCell);
    put_i32("routing_cell", 0, routing.G_routingCell);
    put_f64("fswb_init", 0, routing.G_fswbInit);
    put_f64("f_glo_lake", 0, routing.G_fGloLake);
}

static void dump_state(int day, bool with_snow) {
    put_f64("canopy", day, dailyWaterBalance.G_canopyWaterContent);
    put_f64("soil", day, dailyWaterBalance.G_soilWaterContent);
    put_f64("snow", day, dailyWaterBalance.G_snow);
    if (with_snow) put_f64("snow_bands", day, dailyWaterBalance.G_SnowInElevation, (int64_t)ng * 101);
    put_i32("lai_days", day, lai.G_days_since_start);
    put_i32("lai_status", day, lai.G_GrowingStatus);
    put_f64("lai_precsum", day, lai.G_PrecSum);
    put_f64("gw", day, routing.G_groundwaterStorage);
    put_f64("loc_lake_stor", day, routing.G_locLakeStorage);
    put_f64("loc_wetl_stor", day, routing.G_locWetlStorage);
    put_f64("glo_lake_stor", day, routing.G_gloLakeStorage);
    put_f64("glo_wetl_stor", day, routing.G_gloWetlStorage);
    put_f64("res_stor", day, routing.G_gloResStorage);
    put_f64("river_stor", day, routing.G_riverStorage);
    put_f64("red_loc_lake", day, routing.G_locLakeAreaReductionFactor);
    put_f64("red_loc_wetl", day, routing.G_locWetlAreaReductionFactor);
    put_f64("red_glo_lake", day, routing.G_gloLakeEvapoReductionFactor);
    put_f64("red_glo_wetl", day, routing.G_gloWetlAreaReductionFactor);
    put_f64("red_res", day, routing.G_gloResEvapoReductionFactor);
    put_f64("red_river", day, routing.G_riverAreaReductionFactor);
    put_f64("k_release", day, routing.K_release);
    put_f64("land_area_frac", day, routing.G_landAreaFrac);
    put_f64("land_area_frac_prev", day, routing.G_landAreaFracPrevTimestep);
    put_f64("land_area_frac_next", day, routing.G_landAreaFracNextTimestep);
    put_f64("fswb_laf", day, routing.G_fswbLandAreaFrac);
    put_f64("fswb_laf_next", day, routing.G_fswbLandAreaFracNextTimestep);
    put_f64("river_area_frac_next", day, routing.G_riverAreaFracNextTimestep_Frac);
    put_i16("status_laf_next", day, routing.statusStarted_landAreaFracNextTimestep);
    if (options.subtract_use > 0) {  // water-use state
        put_f64("wu_total_unsatisfied", day, routing.G_totalUnsatisfiedUse);
        put_f64("wu_daily_remaining", day, routing.G_dailyRemainingUse);
        put_f64("wu_daily_nus", day, routing.G_dailydailyNUs);
        put_f64("wu_daily_nug", day, routing.G_dailydailyNUg);
        put_f64("wu_actual_use", day, routing.G_actualUse);
        put_f64("wu_uns_irr", day, routing.G_unsatisfiedNAsFromIrrig);
        put_f64("wu_uns_oth", day, routing.G_unsatisfiedNAsFromOtherSectors);
        put_f64("wu_red_rf", day, routing.G_reducedReturnFlow);
        put_f64("wu_wusi", day, routing.G_withdrawalIrrigFromSwb);
        put_f64("wu_cusi", day, routing.G_consumptiveUseIrrigFromSwb);
    }
}

static void dump_fluxes(int day, int doy) {
    put_f64("lake_balance", day, dailyWaterBalance.G_lakeBalance);
    put_f64("openwater_prec", day, dailyWaterBalance.G_openWaterPrec);
    put_f64("openwater_pet", day, dailyWaterBalance.G_openWaterPET);
    put_f64("surface_runoff", day, dailyWaterBalance.G_dailyLocalSurfaceRunoff);
    put_f64("gw_recharge", day, dailyWaterBalance.G_dailyGwRecharge);
    put_f64("storage_transfer", day, dailyWaterBalance.G_dailyStorageTransfer);
    put_f64("land_aet", day, dailyWaterBalance.dailyCellLandAET);
    put_f64("land_aet_uncorr", day, dailyWaterBalance.dailyCellLandAET_uncorr);
    // river discharge of the day (transportedVolume, km3/d); kept by the reference in
    // G_daily365RiverAvail when grid_store in {5,6} and output option 92 is on
    // (routing.cpp:4219-4221); cells with LDD<0 keep 0 there.
    if (routing.G_daily365RiverAvail.wasInitialized) {
        std::vector<double> q(ng);
        for (int n = 0; n < ng; n++) q[n] = routing.G_daily365RiverAvail(n, doy - 1);
        put("discharge", day, "f64", ng, q.data(), 8);
    }
    put_f64("river_evapo", day, routing.G_dailyRiverEvapo);
    put_f64("gwr_swb", day, routing.G_gwr_swb);
}

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}


static std::vector<double> read_f64(const std::string &path) {
    std::vector<double> v;
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { perror(path.c_str()); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f) / 8;
    fseek(f, 0, SEEK_SET);
    v.resize(n);
    if (fread(v.data(), 8, n, f) != (size_t)n) { perror("read"); exit(2); }
    fclose(f);
    return v;
}

// The state exchange of one assimilation cycle through the reference's own functions (see the header comment).
static void run_enkf(const std::string &dir, ConfigFile *&configFile, WghmStateFile *&wghmState, calibParamClass *&calParam,
                     AdditionalOutputInputFile *&additionalOutIn, SnowInElevationFile *&snow, int year_, int month_) {
    const std::string ids_file = dir + "/ids.txt";
    std::vector<int> ids;
    {
        std::ifstream in(ids_file);
        std::string line;
        std::getline(in, line);
        while (std::getline(in, line)) {
            if (line.empty()) continue;
            std::stringstream ss(line);
            int id; double lo, la;
            ss >> id >> lo >> la;
            ids.push_back(id);
        }
    }
    const long n = (long)ids.size();
    std::vector<double> meanf = read_f64(dir + "/meanfield.bin"), pert = read_f64(dir + "/perturb.bin");
    if ((long)meanf.size() != n * 10 || (long)pert.size() != n * 10) { fprintf(stderr, "enkf: input sizes\n"); exit(2); }
    WghmStateFile *wghmMean = new WghmStateFile(ng, 1);
    for (long i = 0; i < n; i++) {
        Cell &c = wghmMean->cell(ids[i] - 1);
        c.canopy(0) = meanf[i * 10 + 0]; c.snow(0) = meanf[i * 10 + 1]; c.soil(0) = meanf[i * 10 + 2];
        c.locallake(0) = meanf[i * 10 + 3]; c.localwetland(0) = meanf[i * 10 + 4]; c.globallake(0) = meanf[i * 10 + 5];
        c.globalwetland(0) = meanf[i * 10 + 6]; c.reservoir(0) = meanf[i * 10 + 7]; c.river(0) = meanf[i * 10 + 8];
        c.groundwater(0) = meanf[i * 10 + 9];
    }
    long oy = 0, total_nr_calPar = 26, calpar_size = 0, ny = n * 10, step = 0, total_steps = 100, year = year_, month = month_;
    // optional parameter half (extractsub.cpp:81-340, enKF2wghmState.cpp:127-431): calibration units of the region's cells
    //   calpar_index.bin int32 [nunit][26] (1: parameter is in the state vector), groupmatrixindex.bin int32 [nunit][n]
    //   (0 or the 1-based cell number), calpar_range.bin f64 [2][26], calpar_perturb.bin f64 [number of ones]
    std::vector<int> calpar_index, gmi;
    std::vector<double> calpar_range, calpar_pert;
    {
        FILE *f = fopen((dir + "/calpar_index.bin").c_str(), "rb");
        if (f) {
            fseek(f, 0, SEEK_END);
            const long cnt = ftell(f) / 4;
            fseek(f, 0, SEEK_SET);
            calpar_index.resize(cnt);
            if (fread(calpar_index.data(), 4, cnt, f) != (size_t)cnt) exit(2);
            fclose(f);
            oy = cnt / 26;
            gmi.resize(oy * n);
            f = fopen((dir + "/groupmatrixindex.bin").c_str(), "rb");
            if (!f || fread(gmi.data(), 4, gmi.size(), f) != gmi.size()) { fprintf(stderr, "enkf: groupmatrixindex.bin\n"); exit(2); }
            fclose(f);
            calpar_range = read_f64(dir + "/calpar_range.bin");
            calpar_pert = read_f64(dir + "/calpar_perturb.bin");
            for (int v : calpar_index) calpar_size += v == 1;
            if ((long)calpar_pert.size() != calpar_size || calpar_range.size() != 52) { fprintf(stderr, "enkf: parameter inputs\n"); exit(2); }
        }
    }
    double *output = nullptr;
    extract_sub_(ids_file.c_str(), wghmState, output, &oy, calParam, &total_nr_calPar, &calpar_size,
                 calpar_index.empty() ? nullptr : calpar_index.data(), "", wghmMean, gmi.empty() ? nullptr : gmi.data());
    put("enkf_extract", 9000, "f64", n * 10, output, 8);
    std::vector<double> prediction(output, output + n * 10 + calpar_size), field(n * 10 + calpar_size);
    for (long k = 0; k < n * 10; k++) field[k] = prediction[k] + pert[k];
    for (long k = 0; k < calpar_size; k++) field[n * 10 + k] = prediction[n * 10 + k] + calpar_pert[k];
    delete[] output;
    ny = n * 10 + calpar_size;
    put("enkf_field", 9000, "f64", n * 10, field.data(), 8);
    put("enkf_prediction", 9000, "f64", n * 10, prediction.data(), 8);
    if (calpar_size > 0) {
        put("enkf_par_extract", 9000, "f64", calpar_size, prediction.data() + n * 10, 8);
        put("enkf_par_field", 9000, "f64", calpar_size, field.data() + n * 10, 8);
        configFile->outputparameter = dir + "/parameters_out.json";
        if (configFile->calibrationfile.empty()) configFile->calibrationfile = configFile->parameterfile;  // the run's own JSON
    }
    // monthly mean of every cell as the reference forms it (Cell::mean, wghmStateFile.cpp:711)
    {
        std::vector<double> mm((size_t)ng * 10);
        for (int c = 0; c < ng; c++) {
            Cell m = wghmState->cell(c).mean();
            double v[10] = {m.canopy(0), m.snow(0), m.soil(0), m.locallake(0), m.localwetland(0), m.globallake(0),
                            m.globalwetland(0), m.reservoir(0), m.river(0), m.groundwater(0)};
            for (int k = 0; k < 10; k++) mm[(size_t)c * 10 + k] = v[k];
        }
        put("enkf_month_mean", 9000, "f64", (int64_t)ng * 10, mm.data(), 8);
    }
    configFile->outputmeanfile = dir + "/mean_001.txt";
    configFile->outputlastdayfile = dir + "/lastday_001.txt";
    configFile->outputsnowlastdayfile = dir + "/snowlastday_001.txt";
    configFile->outputadditionalfile = "";
    double *factor = nullptr;
    WghmStateFile *wghmStateMean = nullptr;
    const std::string s2 = dir + "/";
    enkf_wghmstate_(ids_file.c_str(), field.data(), prediction.data(), configFile, wghmState, additionalOutIn, snow, &step, &total_steps,
                    &year, &month, &ny, factor, wghmStateMean, s2.c_str(), calParam, &calpar_size, (dir + "/calpar").c_str(),
                    (dir + "/arcid_gcrc.txt").c_str(), "", wghmMean, calpar_range.empty() ? nullptr : calpar_range.data(), &oy,
                    &total_nr_calPar, calpar_index.empty() ? nullptr : calpar_index.data(), gmi.empty() ? nullptr : gmi.data());
    std::vector<double> last(n * 10), sie(n * 101);
    for (long i = 0; i < n; i++) {
        Cell &c = wghmState->cell(ids[i] - 1);
        double v[10] = {c.canopy(0), c.snow(0), c.soil(0), c.locallake(0), c.localwetland(0), c.globallake(0),
                        c.globalwetland(0), c.reservoir(0), c.river(0), c.groundwater(0)};
        for (int k = 0; k < 10; k++) last[i * 10 + k] = v[k];
        for (int e = 0; e < 101; e++) sie[i * 101 + e] = snow->snowInElevation(ids[i] - 1, e);
    }
    put("enkf_lastday", 9000, "f64", n * 10, last.data(), 8);
    put("enkf_snow_elev", 9000, "f64", n * 101, sie.data(), 8);
    // restore of the next cycle (integrateWGHM.cpp:303, 407)
    routing.setStorages(*wghmState, *additionalOutIn);
    dailyWaterBalance.setStorages(*wghmState, *snow, *additionalOutIn);
    dump_state(9000, true);
    delete wghmMean;
}


static int run_calib(const char *dir) {
    if (chdir(dir) != 0) { perror(dir); return 2; }
    FILE *f = fopen("CALIB_IN.txt", "r");
    if (!f) { perror("CALIB_IN.txt"); return 2; }
    int y0, y1, station;
    double gamma0, s0, s1;
    if (fscanf(f, "%d %d %lf %lf %lf %d", &y0, &y1, &gamma0, &s0, &s1, &station) != 6) return 2;
    const int ny = y1 - y0 + 1;
    std::vector<double> base(ny), inflow(ny), use(ny);
    for (int i = 0; i < ny; i++)
        if (fscanf(f, "%lf %lf %lf", &base[i], &inflow[i], &use[i]) != 3) return 2;
    fclose(f);
    options.evalStartYear = (short)y0;
    options.end_year = (short)y1;
    options.output_dir = ".";
    {   // basin mask for createCorrectionGrid, native int16 [ng]
        FILE *b = fopen("SBASIN.bin", "rb");
        if (b) {
            std::vector<short> sb(ng);
            if (fread(sb.data(), 2, ng, b) == (size_t)ng)
                for (int n = 0; n < ng; n++) G_sbasin[n] = sb[n];
            fclose(b);
        }
    }
    for (int n = 0; n < ng; n++) dailyWaterBalance.G_cellCorrFact[n] = 1.0;
    calibGammaClass cg;
    cg.calibStationNumber = (short)station;
    cg.prepareFiles();
    cg.init();
    float gamma = (float)gamma0;
    bool testRun = false;
    for (int it = 0; it < 60; it++) {
        for (int i = 0; i < ny; i++) {
            // the "model run": annual discharge falls with gamma
            cg.setRunoff(y0 + i, (float)(base[i] * (s0 + s1 / (1.0 + (double)gamma))));
            cg.setWaterUse(y0 + i, (float)use[i]);
            cg.setUpstInflow(y0 + i, (float)inflow[i]);
        }
        if (testRun) {
            cg.writeCorrFactors(gamma, cg.cellCorrFactorInd);
            cg.writeCalibStatus(cg.getCalibStatus());
            printf("CALIB_END %.9g %d %d %.9g\n", (double)gamma, cg.cellCorrFactorInd, cg.getCalibStatus(), (double)cg.getCellCorrFactor());
            break;
        }
        const float gamma_old = gamma;
        gamma = cg.findNewGamma(gamma);
        printf("CALIB %d %.9g %.9g %d %d %.9g %d\n", (int)cg.getCallCounter(), (double)gamma_old, (double)gamma, cg.gammaCond,
               cg.getCalibStatus(), (double)cg.getCellCorrFactor(), cg.cellCorrFactorInd);
        if (gamma < 0) {
            testRun = true;
            gamma = gamma_old;
        }
    }
    return 0;
}


static int run_wu_unit(const char *in, const char *out) {
    std::vector<double> v = read_f64(in);
    if ((long)v.size() != 8L * ng) { fprintf(stderr, "wu_unit: expected 8 x %d doubles\n", (int)ng); return 2; }
    for (int n = 0; n < ng; n++) {
        routing.G_dailyRemainingUse[n] = v[0 * (size_t)ng + n];
        routing.G_withdrawalIrrigFromSwb[n] = v[1 * (size_t)ng + n];
        routing.G_consumptiveUseIrrigFromSwb[n] = v[2 * (size_t)ng + n];
        routing.G_fractreturngw_irrig[n] = v[3 * (size_t)ng + n];
        routing.G_unsatisfiedNAsFromIrrig[n] = v[4 * (size_t)ng + n];
        routing.G_unsatisfiedNAsFromOtherSectors[n] = v[5 * (size_t)ng + n];
        routing.G_reducedReturnFlow[n] = v[6 * (size_t)ng + n];
        routing.G_dailydailyNUg[n] = v[7 * (size_t)ng + n];
    }
    std::vector<double> o(6 * (size_t)ng);
    for (int n = 0; n < ng; n++) {
        o[0 * (size_t)ng + n] = routing.updateNetAbstractionGW(n);
        o[1 * (size_t)ng + n] = routing.G_dailyRemainingUse[n];
        o[2 * (size_t)ng + n] = routing.G_unsatisfiedNAsFromIrrig[n];
        o[3 * (size_t)ng + n] = routing.G_unsatisfiedNAsFromOtherSectors[n];
        o[4 * (size_t)ng + n] = routing.G_reducedReturnFlow[n];
        o[5 * (size_t)ng + n] = routing.G_dailydailyNUg[n];
    }
    FILE *f = fopen(out, "wb");
    if (!f) { perror(out); return 2; }
    fwrite(o.data(), 8, o.size(), f);
    fclose(f);
    return 0;
}

static int run_driver(const char *cfg) {
    std::string progName = "OL", path_mean;
    WghmStateFile *wghmState, *wghmMean;
    calibParamClass *calParam;
    AdditionalOutputInputFile *additionalOutIn;
    SnowInElevationFile *snow;
    ConfigFile *configFile;
    long year = 0, month = 0, total_steps = 0, step = 0;
    double t0 = now();
    initialize_wghm(cfg, wghmState, calParam, additionalOutIn, snow, &year, &month, progName.c_str(), path_mean.c_str(), wghmMean);
    integrate_wghm(cfg, configFile, wghmState, calParam, additionalOutIn, snow, &step, &total_steps, &year, &month, progName.c_str());
    fprintf(stderr, "REF_DRIVER_SECONDS %.6f\n", now() - t0);
    return 0;
}

static int run_replay(int argc, char **argv) {
    const char *cfg = argv[2];
    const char *dumpfile = argv[3];
    Range days, snowdays;
    int every = 0;
    bool time_only = false, deep_snow = false;
    std::string final_prefix, day_times_file, enkf_dir;
    std::vector<double> day_times;
    for (int i = 4; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--days" && i + 1 < argc) days = parse_range(argv[++i]);
        else if (a == "--snow-days" && i + 1 < argc) snowdays = parse_range(argv[++i]);
        else if (a == "--every" && i + 1 < argc) every = atoi(argv[++i]);
        else if (a == "--final-state" && i + 1 < argc) final_prefix = argv[++i];
        else if (a == "--time-only") time_only = true;
        else if (a == "--deep-snow") deep_snow = true;
        else if (a == "--day-times" && i + 1 < argc) day_times_file = argv[++i];
        else if (a == "--enkf" && i + 1 < argc) enkf_dir = argv[++i];
    }
    if (!time_only && strcmp(dumpfile, "-") != 0) {
        g_dump = fopen(dumpfile, "wb");
        if (!g_dump) { perror(dumpfile); return 2; }
    }

    std::string progName = "OL", path_mean;
    WghmStateFile *wghmState, *wghmMean;
    calibParamClass *calParam;
    AdditionalOutputInputFile *additionalOutIn;
    SnowInElevationFile *snow_in_elevation;
    long lyear = 0, lmonth = 0;
    initialize_wghm(cfg, wghmState, calParam, additionalOutIn, snow_in_elevation, &lyear, &lmonth, progName.c_str(), path_mean.c_str(), wghmMean);
    ConfigFile *configFile = new ConfigFile(cfg, (int)lyear, (int)lmonth, progName);

    const short number_of_days_in_month[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    const short last_day_in_month[12] = {30, 58, 89, 119, 150, 180, 211, 242, 272, 303, 333, 364};
    short readinstatus = 1;
#ifdef _OPENMP
    omp_set_num_threads(8);  // integrateWGHM.cpp:106-111 (the reference aborts with any other team size)
#endif
    // ---- init sequence, integrateWGHM.cpp:127-286 -------------------------------------
    options.init(*configFile);
    options.createModelSettingsFile(options.output_dir);
    options.closeModelSettingsFile();
    if (1 == options.rout_prepare)
        prepare_routing_files(options.input_dir, options.routing_dir, 1, options.resOpt);
    geo.init(options.input_dir, options.fileEndianType);
    climate.init();
    short nSpecBasins = cbasin.prepare(options.input_dir, options.output_dir, options.routing_dir);
    dailyWaterBalance.init(options.input_dir, nSpecBasins);
    G_sbasin.read(options.output_dir + "/G_CALIB_BASIN.UNF2");
    directUpstSt.findDirectStations(options.routing_dir);
    allUpstSt.findAllStations(options.routing_dir);
    geo.calibrun = 0;
    land.init(options.input_dir);
    G_aindex.read(options.input_dir + "/G_ARID_HUMID.UNF2");
    G_LDD.read(options.routing_dir + "/G_LDD_2.UNF1");
    dailyWaterBalance.G_Elevation.read(options.input_dir + "/G_ELEV_RANGE.101.UNF2");
    lai.init(options.input_dir, options.output_dir, land.G_landCover, *additionalOutIn, *calParam);
    routing.init(nSpecBasins, *configFile, *wghmState, *additionalOutIn);
    routing.initLakeDepthActive(*calParam);
    routing.initWetlDepthActive(*calParam);
    // ---- body of the (single-pass) calibration do-while, integrateWGHM.cpp:291-476 ------
    if (additionalOutIn->additionalfilestatus == 0) routing.initFractionStatus();
    if (additionalOutIn->additionalfilestatus == 1) routing.initFractionStatusAdditionalOI(*additionalOutIn);
    if (configFile->additionalfile.empty()) routing.setStoragesToZero();
    else routing.setStorages(*wghmState, *additionalOutIn);
    if (configFile->startvaluefile.empty()) routing.setLakeWetlToMaximum(options.start_year);
    if ((readinstatus == 1) && (additionalOutIn->additionalfilestatus == 1)) {  // integrateWGHM.cpp:311-316 (first_day_after_PDAF): restart from checkpoints
        routing.annualInit(options.start_year, configFile->startMonth, *additionalOutIn);
        routing.update_landarea_red_fac_PDAF(*calParam, *additionalOutIn);
    }
    for (int n = 0; n < ng; ++n) G_toBeCalculated[n] = 1;  // options.basin == 1
    for (int n = 0; n < ng; n++) {
        dailyWaterBalance.G_gammaHBV[n] = calParam->getValue(P_GAMRUN_C, n);
        dailyWaterBalance.G_cellCorrFact[n] = calParam->getValue(P_CFA, n);
        routing.G_statCorrFact[n] = calParam->getValue(P_CFS, n);
    }
    if (configFile->additionalfile.empty()) dailyWaterBalance.setStoragesToZero();
    else dailyWaterBalance.setStorages(*wghmState, *snow_in_elevation, *additionalOutIn);
    maxSoilWaterCap.createMaxSoilWaterCapacityGrid(options.input_dir, options.output_dir, land.G_landCover, *calParam);
    GW.createGrids(options.input_dir, options.output_dir, *calParam);
    climate.read_climate_longtermAvg();

    double t_vert = 0, t_rout = 0, t_laf = 0, t_io = 0;
    long ndays_done = 0;
    int simday = 0;
    bool static_done = false;
    short start_year = options.start_year, end_year = options.end_year;
    short day = 0, month = 0, actual_year = 0;
    for (actual_year = start_year; actual_year <= end_year; actual_year += options.time_step) {
        dailyWaterBalance.annualInit();
        routing.annualInit(actual_year, configFile->startMonth, *additionalOutIn);
        // net abstractions of the year (integrateWGHM.cpp:645-647; SURVEY 8f-4, not on the product path yet)
        if (((2 == options.subtract_use) || (3 == options.subtract_use)) && (0 == options.time_series))
            routing.dailyNUInit(options.water_use_dir, actual_year, *calParam);
        day = 0;
        if (actual_year == configFile->startYear)
            for (int m = 1; m < configFile->startMonth; m++) day += number_of_days_in_month[m - 1];
        if ((5 == options.grid_store) || (6 == options.grid_store)) {
            dailyWaterBalance.daily365outInit();
            routing.daily365outInit();
        }
        short start_month = configFile->startMonth, end_month = configFile->endMonth;
        if (start_year == end_year) { /* as configured */ }
        else if (actual_year == start_year) end_month = 12;
        else if (actual_year == end_year) start_month = 1;
        else { start_month = 1; end_month = 12; }
        if (!static_done) {
            if (deep_snow) {
                // fixture for the 1000 mm snow cap (daily.cpp:958-976), which a run from empty storages does
                // not reach within months: preload the band snow of every third cell through the public
                // grid (the reference's code is untouched, only its start state differs)
                for (int n = 0; n < ng; n += 3) {
                    double sum = 0.;
                    for (short e = 1; e <= 100; e++) {
                        double v = ((n / 3) % 2) ? 900. + 3. * e : (double)((n * 131 + e * 37) % 1400) + 0.25 * e;
                        dailyWaterBalance.G_SnowInElevation(n, e) = v;
                        sum += v;
                    }
                    dailyWaterBalance.G_snow[n] = sum / 100.;
                }
            }
            dump_static(*calParam);
            dump_state(0, true);
            static_done = true;
        }
        for (month = start_month - 1; month < end_month; month++) {
            double t0 = now();
            climate.read_climate_data_daily(month + 1, actual_year);
            wghmState->resetCells(number_of_days_in_month[month]);
            t_io += now() - t0;
            for (short day_in_month = 1; day_in_month <= number_of_days_in_month[month]; day_in_month++) {
                day++;
                simday++;
                double ta = now();
#pragma omp parallel for
                for (int n = 0; n < ng; ++n) {
                    if (geo.G_contcell[n])
                        dailyWaterBalance.calcNewDay(day, month, day_in_month, last_day_in_month[month], actual_year, n,
                                                     *wghmState, *additionalOutIn, *snow_in_elevation, readinstatus, *calParam);
                }
                double tb = now();
                if ((2 == options.subtract_use) || (3 == options.subtract_use)) routing.calcNextDay_M(month);  // integrateWGHM.cpp:794-796
                routing.routing(actual_year, day, month, day_in_month, last_day_in_month[month], *wghmState,
                                *additionalOutIn, readinstatus, *calParam);
                double tc = now();
                routing.updateLandAreaFrac(*additionalOutIn);
                double td = now();
                t_vert += tb - ta; t_rout += tc - tb; t_laf += td - tc;
                day_times.push_back(td - ta);
                ndays_done++;
                if (readinstatus == 1) readinstatus = 0;
                bool want = days.has(simday) || (every > 0 && simday % every == 0);
                if (g_dump && want) {
                    dump_state(simday, snowdays.has(simday));
                    dump_fluxes(simday, day);
                    if (month >= 0) {
                        // wghmState of the day (7 routing compartments, mm over continental area; routing.cpp:5002-5020)
                        std::vector<double> v((size_t)7 * ng);
                        for (int n = 0; n < ng; n++) {
                            Cell &c = wghmState->cell(n);
                            v[0 * (size_t)ng + n] = c.locallake(day_in_month - 1);
                            v[1 * (size_t)ng + n] = c.localwetland(day_in_month - 1);
                            v[2 * (size_t)ng + n] = c.globallake(day_in_month - 1);
                            v[3 * (size_t)ng + n] = c.globalwetland(day_in_month - 1);
                            v[4 * (size_t)ng + n] = c.reservoir(day_in_month - 1);
                            v[5 * (size_t)ng + n] = c.river(day_in_month - 1);
                            v[6 * (size_t)ng + n] = c.groundwater(day_in_month - 1);
                        }
                        put("wghm_routing_mm", simday, "f64", (int64_t)7 * ng, v.data(), 8);
                    }
                }
            }
            // month-end rescale into wghmState / snow_in_elevation, integrateWGHM.cpp:830-847
            for (int n = 0; n < ng; ++n) {
                double landAreaFrac = routing.getLandAreaFrac(n);
                for (short elev = 0; elev <= 100; elev++) {
                    if (landAreaFrac == 0.) snow_in_elevation->snowInElevation(n, elev) = 0.;
                    else snow_in_elevation->snowInElevation(n, elev) =
                             dailyWaterBalance.G_SnowInElevation(n, elev) * landAreaFrac / geo.G_contfreq[n];
                }
                for (short dim = 1; dim <= number_of_days_in_month[month]; dim++) {
                    wghmState->cell(n).canopy(dim - 1) = dailyWaterBalance.G_canopyWaterContent[n] * landAreaFrac / geo.G_contfreq[n];
                    wghmState->cell(n).snow(dim - 1) = dailyWaterBalance.G_snow[n] * landAreaFrac / geo.G_contfreq[n];
                    wghmState->cell(n).soil(dim - 1) = dailyWaterBalance.G_soilWaterContent[n] * landAreaFrac / geo.G_contfreq[n];
                }
            }
            if (!enkf_dir.empty() && actual_year == end_year && month + 1 == end_month)
                run_enkf(enkf_dir, configFile, wghmState, calParam, additionalOutIn, snow_in_elevation, actual_year, month + 1);
            if (!final_prefix.empty() && actual_year == end_year && month + 1 == end_month) {
                wghmState->saveDay(final_prefix + "_state.txt", number_of_days_in_month[month] - 1);
                additionalOutIn->save(final_prefix + "_additional.txt");
                snow_in_elevation->save(final_prefix + "_snow.txt");
            }
        }
        if (month == 12) routing.updateGloResPrevYear_pct();
    }
    if (g_dump) fclose(g_dump);
    if (!day_times_file.empty()) {  // seconds spent in the day-loop body, one line per simulated day
        FILE *f = fopen(day_times_file.c_str(), "w");
        if (f) { for (double t : day_times) fprintf(f, "%.9f\n", t); fclose(f); }
    }
    long ncalc = 0;
    for (int n = 0; n < ng; n++) if (geo.G_contcell[n] && G_toBeCalculated[n] == 1) ncalc++;
    double tt = t_vert + t_rout + t_laf;
    printf("REF_TIMING {\"ng\": %d, \"cells\": %ld, \"days\": %ld, \"threads\": 8, \"t_vertical_s\": %.6f, "
           "\"t_routing_s\": %.6f, \"t_landfrac_s\": %.6f, \"t_forcing_io_s\": %.6f, \"cell_days_per_s\": %.6e}\n",
           (int)ng, ncalc, ndays_done, t_vert, t_rout, t_laf, t_io, tt > 0 ? (double)ncalc * ndays_done / tt : 0.0);
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 3 && std::string(argv[1]) == "driver") return run_driver(argv[2]);
    if (argc >= 4 && std::string(argv[1]) == "replay") return run_replay(argc, argv);
    if (argc >= 3 && std::string(argv[1]) == "calib") return run_calib(argv[2]);
    if (argc >= 4 && std::string(argv[1]) == "wu_unit") return run_wu_unit(argv[2], argv[3]);
    fprintf(stderr, "usage: ref_harness driver <config> | replay <config> <dump|-> [--days A-B] [--every K] "
                    "[--snow-days A-B] [--final-state PREFIX] [--time-only] [--day-times FILE]\n");
    return 1;
}
