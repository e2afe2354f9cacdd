// wg_model.cpp — see wg_model.h.  Host-side drop-in layer over the C ABI (include/wgk.h).
#include "wg_model.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace wg {

// ------------------------------------------------------------------------------------------
// parameters / config / options
// ------------------------------------------------------------------------------------------
const char *const kParamNames[26] = {  // calib_param.cpp:140-171
    "gammaHBV_runoff_coeff", "CFA_cellCorrFactor", "CFS_statCorrFactor", "root_depth_multiplier",
    "river_roughness_coeff_mult", "lake_depth", "wetland_depth", "surfacewater_outflow_coefficient",
    "evapo_red_fact_exp_mult", "net_radiation_mult", "PT_coeff_humid", "PT_coeff_arid", "max_daily_PET", "mcwh",
    "LAI_mult", "snow_freeze_temp", "snow_melt_temp", "degree_day_factor_mult", "temperature_gradient", "gw_factor_mult",
    "rg_max_mult", "pcrit_aridgw", "groundwater_outflow_coeff", "net_abstraction_surfacewater_mult",
    "net_abstraction_groundwater_mult", "precip_mult"};

void calibParamClass::readJson(const std::string &file, int ncell) {
    // The parameter file is one JSON object of "name": number | [numbers]; a full DOM is not
    // needed (and is what makes getValue() slow in the reference, calib_param.h:131-222).
    std::ifstream f(file, std::ios::binary);
    if (!f) throw std::runtime_error("Unable to open file '" + file + "'");
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string txt = ss.str();
    ncell_ = ncell > 0 ? ncell : 0;
    v_.assign((size_t)26 * ncell_, 0.0);
    auto find_key = [&](const std::string &key) -> size_t {
        const std::string q = "\"" + key + "\"";
        size_t p = txt.find(q);
        if (p == std::string::npos) return p;
        p = txt.find(':', p + q.size());
        return p == std::string::npos ? p : p + 1;
    };
    size_t p = find_key("ng_param");
    if (p == std::string::npos) throw std::runtime_error("ERROR readJson(): calibParamJson.is_object() == false");
    const int ng_file = (int)strtod(txt.c_str() + p, nullptr);
    if (ncell <= 0) {  // the B1 entry points take the cell count from the parameter file (the reference compiles it in, def.h:10)
        ncell = ng_file;
        ncell_ = ncell;
        v_.assign((size_t)26 * ncell, 0.0);
    }
    if (ng_file != ncell)
        throw std::runtime_error("ERROR readJson(): Predefined number of grid cells " + std::to_string(ncell) +
                                 " does not match number found in JSON object header " + std::to_string(ng_file));
    for (int k = 0; k < 26; k++) {
        p = find_key(kParamNames[k]);
        if (p == std::string::npos) continue;  // missing parameters read as 0, like the reference's DOM lookup
        p = txt.find('[', p);
        const char *s = txt.c_str() + p + 1;
        for (int n = 0; n < ncell; n++) {
            char *e;
            v_[(size_t)k * ncell + n] = strtod(s, &e);
            s = e;
            while (*s == ',' || *s == ' ' || *s == '\n') s++;
        }
    }
}

ConfigFile::ConfigFile(const std::string &file, int year, int month, const std::string &progName) : ConfigFile(file) {
    if (progName != "OL") {  // configFile.cpp:76-113: a PDAF cycle runs exactly the month it is called for
        startMonth = endMonth = month;
        startYear = endYear = year;
    }
}

ConfigFile::ConfigFile(const std::string &file) {
    std::ifstream f(file);
    if (!f) throw std::runtime_error("error by opening file " + file);
    std::string line;
    while (std::getline(f, line)) {
        std::stringstream ss(line);
        std::string tag;
        if (!(ss >> tag) || tag[0] == '#') continue;
        if (tag == "end_of_head") break;
        std::map<std::string, std::string *> str = {
            {"wghm_state", &startvaluefile}, {"param_json", &parameterfile}, {"calibration_parameters", &calibrationfile},
            {"output_calibration_parameters", &outputparameter}, {"snowInElevation_startvalues", &snowInElevationfile},
            {"additionalOutIn_startvalues", &additionalfile}, {"output_state_mean", &outputmeanfile},
            {"output_state_lastday", &outputlastdayfile}, {"output_snowInElevation_lastday", &outputsnowlastdayfile},
            {"additionalOutIn_lastday", &outputadditionalfile}, {"runtime_options", &runtimeoptionsfile},
            {"output_options", &outputoptionsfile}, {"routing", &routingoptionsfile}, {"stations", &stationsfile},
            {"input_dir", &inputDir}, {"output_dir", &outputDir}, {"climate_dir", &climateDir}, {"routing_dir", &routingDir},
            {"water_use_dir", &waterUseDir}};
        std::map<std::string, int *> num = {{"start_month", &startMonth}, {"start_year", &startYear}, {"end_month", &endMonth},
                                            {"end_year", &endYear}, {"time_step", &timeStep}, {"num_init_years", &numInitYears}};
        if (str.count(tag)) ss >> *str[tag];
        else if (num.count(tag)) ss >> *num[tag];
    }
}

void optionClass::init(const ConfigFile &cfg) {
    // defaults of option.cpp:120-165
    const int def[36] = {2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 2000, 1900, 2010, 0, 1971, 2000, 0, 0};
    std::memcpy(v, def, sizeof v);
    std::ifstream f(cfg.runtimeoptionsfile);
    std::string line;
    int i = 0;
    while (f && std::getline(f, line))
        if (line.compare(0, 6, "Value:") == 0 && i < 36) v[i++] = atoi(line.c_str() + 7);
    input_dir = cfg.inputDir; output_dir = cfg.outputDir; climate_dir = cfg.climateDir; routing_dir = cfg.routingDir;
    water_use_dir = cfg.waterUseDir;
    start_year = cfg.startYear; end_year = cfg.endYear;
}

void optionClass::require_canonical() const {
    struct { int idx, want; const char *name; } req[] = {
        {6, 1, "cloud"}, {7, 1, "intercept"}, {8, 0, "calc_albedo"}, {9, 0, "petOpt"}, {10, 1, "use_kc"},
        {14, 1, "riverveloOpt"}, {18, 0, "clclOpt"}, {20, 1, "resOpt"}, {21, 0, "statcorrOpt"},
        {22, 1, "aridareaOpt"}, {23, 1, "fractionalRoutingOpt"}, {24, 1, "riverEvapoOpt"},
        {31, 0, "antNatOpt"}, {34, 0, "calc_wtemp"}, {35, 0, "glacierOpt"}, {1, 1, "basin"}};
    if (v[5] != 0 && v[5] != 1)  // monthly .31 files (climate.cpp:93-138) or yearly .365 files (climateYear.cpp:38-79)
        throw std::runtime_error("option time_series = " + std::to_string(v[5]) + " is outside the implemented hot path (0: .31 files, 1: .365 files)");
    if (v[27] != 0 && v[27] != 1)
        throw std::runtime_error("option resYearOpt = " + std::to_string(v[27]) + " must be 0 (reference year) or 1 (reservoirs start operating in their year)");
    if (v[15] != 0 && v[15] != 2)
        throw std::runtime_error("option subtract_use = " + std::to_string(v[15]) + " is outside the implemented hot path (0: no water use, 2: net abstractions)");
    if (v[15] == 2 && (v[16] != 0 || v[17] != 0 || v[25] != 0))
        throw std::runtime_error("water use is implemented with use_alloc 0, delayedUseSatisfaction 0, aggrNUsGloLakResOpt 0");
    for (auto &r : req)
        if (v[r.idx] != r.want)
            throw std::runtime_error(std::string("option ") + r.name + " = " + std::to_string(v[r.idx]) +
                                     " is outside the implemented hot path (canonical value " + std::to_string(r.want) + ")");
}

void geoClass::init(const std::string &in, int ncell, int resOpt) {  // geo.cpp:7-45
    std::vector<float> a(360);
    read_unf_raw(in + "/GAREA.UNF0", a.data(), a.size());
    area.assign(a.begin(), a.end());
    G_row.initialize(ncell); G_col.initialize(ncell); G_contcell.initialize(ncell); G_contfreq.initialize(ncell);
    G_row.read(in + "/GR.UNF2");
    G_col.read(in + "/GC.UNF2");
    G_contfreq.read(in + "/GCONTFREQ.UNF0");
    Grid<double> fw(ncell), tmp(ncell);
    for (const char *nm : {"/G_LOCLAK.UNF0", "/G_GLOLAK.UNF0"}) { tmp.read(in + nm); for (int n = 0; n < ncell; n++) fw[n] += tmp[n]; }
    if (resOpt == 1)
        for (const char *nm : {"/G_LOCRES.UNF0", "/G_RES/G_RES_FRAC.UNF0"}) { tmp.read(in + nm); for (int n = 0; n < ncell; n++) fw[n] += tmp[n]; }
    for (int n = 0; n < ncell; n++) {
        if (fw[n] > G_contfreq[n]) G_contfreq[n] = fw[n];
        G_contcell[n] = (G_contfreq[n] < 0.000001) ? 0 : 1;
    }
}

// ------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------
Engine::Engine(int nc, int device, int restart, int subtract_use) : ncell(nc), dailyWaterBalance(*this), routing(*this) {
    if (device < 0) return;  // host-side initialisation only: no context, nothing can be stepped
    wgk_options opt{restart, 0, 1, subtract_use};  // restart = additionalOutIn.additionalfilestatus (daily.cpp:165)
    check(wgk_create(&ctx, device, ncell, 1, 1, &opt), "wgk_create");
}
Engine::~Engine() { if (ctx) wgk_destroy(ctx); }

void Engine::check(int rc, const char *what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + (ctx ? wgk_last_error(ctx) : "no CUDA device (wgk has no CPU fallback)"));
}
template <class T, int C> void Engine::set(const char *name, const Grid<T, C> &g, int index) {
    check(wgk_set_field(ctx, wgk_field_id(name), index, g.data(), g.size() * sizeof(T)), name);
}
template <class T, int C> void Engine::get(const char *name, Grid<T, C> &g, int index) {
    if (!g.initialized()) g.initialize(ncell);
    check(wgk_get_field(ctx, wgk_field_id(name), index, g.data(), g.size() * sizeof(T)), name);
}

void Engine::land_init() {
    G_built_up.initialize(ncell); G_landCover.initialize(ncell);
    G_built_up.read(options.input_dir + "/GBUILTUP.UNF0");
    G_landCover.read(options.input_dir + "/G_LANDCOVER.UNF1");
}

static void skip_comments(std::ifstream &f) {
    std::string line;
    while (f && f.peek() == '#') std::getline(f, line);
}

void dailyWaterBalanceClass::init(const std::string &input_dir, short) {  // readLCTdata, daily.cpp:1925-1959
    std::ifstream f(input_dir + "/LCT_22.DAT");
    if (!f) throw std::runtime_error("Can not open file " + input_dir + "/LCT_22.DAT for reading.");
    skip_comments(f);
    for (int i = 0; i < 18; i++) {
        int idx;
        f >> idx >> rootingDepth_lct[i] >> albedo_lct[i] >> albedoSnow_lct[i] >> ddf_lct[i] >> emissivity_lct[i];
        if (idx - 1 != i) throw std::runtime_error("Problem with LCT_22.DAT");
    }
    const int ng = eng.ncell;
    for (auto *g : {&G_gammaHBV, &G_cellCorrFact, &G_snow, &G_soilWaterContent, &G_canopyWaterContent, &G_lakeBalance, &G_openWaterPET,
                    &G_openWaterPrec, &G_dailyLocalSurfaceRunoff, &G_dailyGwRecharge, &G_dailyStorageTransfer})
        g->initialize(ng);
    G_Elevation.initialize(ng);
    G_SnowInElevation.initialize(ng);
}

void Engine::lai_init(AdditionalOutputInputFile &additionalOutIn) {  // lai.cpp:40-148
    std::ifstream f(options.input_dir + "/LAI_22.DAT");
    if (!f) throw std::runtime_error("Can not open file " + options.input_dir + "/LAI_22.DAT for reading.");
    skip_comments(f);
    float LeafAreaIndex[18], decid[18], evergreen[18];
    for (int i = 0; i < 18; i++) {
        int idx;
        f >> idx >> LeafAreaIndex[i] >> decid[i] >> evergreen[i] >> lai_initialDays[i] >> kc_min[i] >> kc_max[i];
        if (idx - 1 != i) throw std::runtime_error("Problem with LAI_22.DAT");
    }
    G_LAImax.initialize(ncell);
    for (int n = 0; n < ncell; n++) G_LAImax[n] = calParam.getValue(M_LAI, n) * LeafAreaIndex[G_landCover[n] - 1];
    for (int i = 0; i < 18; i++) {
        lai_factor_a[i] = 0.1 * decid[i];
        lai_factor_b[i] = (1 - decid[i]) * evergreen[i];
    }
    lai_days.initialize(ncell); lai_status.initialize(ncell); lai_precsum.initialize(ncell);
    if (additionalOutIn.additionalfilestatus != 0)  // :53-65: growing-season state of the checkpoint (a cold start leaves the zeros)
        for (int n = 0; n < ncell; n++) {
            lai_days[n] = (int32_t)additionalOutIn.additionalOutputInput(n, 0);
            lai_status[n] = (int32_t)additionalOutIn.additionalOutputInput(n, 1);
            lai_precsum[n] = additionalOutIn.additionalOutputInput(n, 2);
        }
}

void Engine::createMaxSoilWaterCapacityGrid() {  // s_max.cpp:40-75
    Grid<float> G_TAWC(ncell);
    G_TAWC.read(options.input_dir + "/G_TAWC.UNF0");
    G_Smax.initialize(ncell);
    for (int n = 0; n < ncell; n++) {
        const float rootDepth = calParam.getValue(M_ROOT_D, n) * (double)dailyWaterBalance.rootingDepth_lct[G_landCover[n] - 1];
        if (G_TAWC[n] < 0) G_Smax[n] = -9999;
        else G_Smax[n] = G_TAWC[n] * rootDepth;
    }
}

void Engine::createGroundwaterGrids() {  // gw_frac.cpp:36-275
    const std::string in = options.input_dir;
    Grid<int8_t> slope_class(ncell), permaglac(ncell), aquifer(ncell);
    slope_class.read(in + "/G_SLOPE_CLASS.UNF1");
    G_texture.initialize(ncell);
    G_texture.read(in + "/G_TEXTURE.UNF1");
    permaglac.read(in + "/G_PERMAGLAC.UNF1");
    aquifer.read(in + "/G_AQ_FACTOR.UNF1");
    Grid<float> slope_factor(ncell), texture_factor(ncell), Rgmax_f(ncell), corr(ncell);
    const short slope_class_table[7] = {10, 20, 30, 40, 50, 60, 70};
    const float slope_factor_table[7] = {1.00, 0.95, 0.90, 0.75, 0.60, 0.30, 0.15};
    for (int n = 0; n < ncell; n++) {
        if (slope_class[n] == 0) slope_class[n] = 10;
        if (slope_class_table[6] == slope_class[n]) { slope_factor[n] = slope_factor_table[6]; continue; }
        for (short i = 0; i <= 5; i++) {
            if (slope_class_table[i] == slope_class[n]) { slope_factor[n] = slope_factor_table[i]; break; }
            if (slope_class[n] > slope_class_table[i] && slope_class[n] < slope_class_table[i + 1]) {
                slope_factor[n] = slope_factor_table[i] + (((slope_factor_table[i + 1] - slope_factor_table[i]) / (slope_class_table[i + 1] - slope_class_table[i])) * (slope_class[n] - slope_class_table[i]));
                break;
            }
        }
    }
    const short texture_table[3] = {10, 20, 30};
    float Rgmax_table[3] = {5., 3., 1.5};
    const float Rgmax_tableWFD[3] = {7., 4.5, 2.5};
    if (0 == options.time_series) for (int i = 0; i < 3; i++) Rgmax_table[i] = Rgmax_tableWFD[i];
    const float texture_factor_table[3] = {1, 0.95, 0.70};
    for (int n = 0; n < ncell; n++) {
        const double M_RG_MAX = calParam.getValue(wg::M_RG_MAX, n);
        if (G_texture[n] <= 0 || 2 == G_texture[n]) { texture_factor[n] = 0.95; Rgmax_f[n] = 3; continue; }
        if (1 == G_texture[n]) { texture_factor[n] = 0; Rgmax_f[n] = 0; continue; }
        for (short i = 0; i <= 2; i++) {
            if (texture_table[i] == G_texture[n]) { texture_factor[n] = texture_factor_table[i]; Rgmax_f[n] = M_RG_MAX * Rgmax_table[i]; break; }
            if (i < 2 && G_texture[n] > texture_table[i] && G_texture[n] < texture_table[i + 1]) {
                texture_factor[n] = texture_factor_table[i] + (((texture_factor_table[i + 1] - texture_factor_table[i]) / (texture_table[i + 1] - texture_table[i])) * (G_texture[n] - texture_table[i]));
                Rgmax_f[n] = (M_RG_MAX * Rgmax_table[i]) + ((M_RG_MAX * (Rgmax_table[i + 1] - Rgmax_table[i]) / (texture_table[i + 1] - texture_table[i])) * (G_texture[n] - texture_table[i]));
                break;
            }
        }
    }
    G_gwFactor.initialize(ncell);
    for (int n = 0; n < ncell; n++) {
        if (texture_factor[n] < 0) { G_gwFactor[n] = -99; continue; }
        float slopeFactor = slope_factor[n], textureFactor = texture_factor[n];
        float aquiferFactor = (short)aquifer[n] / 100.0;
        float permaFactor = 1. - (((float)permaglac[n] / 100.));
        auto clamp = [](float x) { return x < 0 ? 0.f : (x > 1 ? 1.f : x); };
        slopeFactor = clamp(slopeFactor); aquiferFactor = clamp(aquiferFactor); textureFactor = clamp(textureFactor); permaFactor = clamp(permaFactor);
        G_gwFactor[n] = calParam.getValue(M_GW_F, n) * slopeFactor * textureFactor * aquiferFactor * permaFactor;
        if (G_gwFactor[n] > 1.) G_gwFactor[n] = 0.95;
    }
    G_Rgmax.initialize(ncell);
    for (int n = 0; n < ncell; n++) G_Rgmax[n] = (Rgmax_f[n] >= 0) ? (short)floor(Rgmax_f[n] * 100 + 0.5) : -9999;
    corr.read(in + "/G_GW_FACTOR_CORR.UNF0");
    for (int n = 0; n < ncell; n++)
        if (corr[n] > 0.) {
            G_gwFactor[n] = calParam.getValue(M_GW_F, n) * corr[n];
            if (G_gwFactor[n] > 1.) G_gwFactor[n] = 0.95;
        }
}

// ------------------------------------------------------------------------------------------
// routing class
// ------------------------------------------------------------------------------------------
void routingClass::init(short, const ConfigFile &, WghmStateFile &, AdditionalOutputInputFile &additionalOutIn) {  // routing.cpp:131-742 (canonical: antNatOpt 0, resOpt 1)
    const int ng = eng.ncell;
    const std::string in = eng.options.input_dir, rd = eng.options.routing_dir;
    for (auto *g : {&G_statCorrFact, &G_landAreaFrac, &G_landAreaFracNextTimestep, &G_landAreaFracPrevTimestep, &G_locLakeStorage,
                    &G_locWetlStorage, &G_gloLakeStorage, &G_gloWetlStorage, &G_gloResStorage, &G_riverStorage, &G_groundwaterStorage,
                    &G_locLakeAreaReductionFactor, &G_locWetlAreaReductionFactor, &G_gloLakeEvapoReductionFactor,
                    &G_gloWetlAreaReductionFactor, &G_gloResEvapoReductionFactor, &G_riverAreaReductionFactor, &G_glo_lake, &G_loc_lake,
                    &G_loc_res, &G_glo_wetland, &G_loc_wetland, &G_reg_lake, &G_glo_res, &G_glores_prevyear, &G_lake_area,
                    &G_reservoir_area, &G_reservoir_area_full, &G_stor_cap, &G_stor_cap_full, &G_mean_outflow, &G_mean_demand,
                    &G_riverLength, &G_RiverSlope, &G_Roughness, &G_bankfull_flow, &G_RiverWidth_bf, &G_RiverDepth_bf,
                    &G_riverBottomWidth, &G_riverStorageMax, &G_lakeDepthActive, &G_wetlDepthActive, &G_fswbInit, &G_fswbLandAreaFrac,
                    &G_fswbLandAreaFracNextTimestep, &G_fGloLake, &G_fLocLake, &G_fLocWet, &G_fGloWet, &G_riverAreaFracNextTimestep_Frac, &K_release,
                    &G_riverDischarge})
        g->initialize(ng);
    statusStarted_landAreaFracNextTimestep.initialize(ng);
    statusStarted_landfreq.assign(ng, 0);
    G_res_type.initialize(ng); G_start_month.initialize(ng); G_reg_lake_status.initialize(ng); G_LDD.initialize(ng);
    G_res_start_year.initialize(ng); G_downstreamCell.initialize(ng); G_routOrder.initialize(ng);
    G_outflow_cell_assignment.initialize(ng);
    G_outflow_cell_assignment.read(in + "/G_OUTFLOW_CELL_ASSIGNMENT.UNF4");  // :290
    G_glo_lake.read(in + "/G_GLOLAK.UNF0"); G_loc_lake.read(in + "/G_LOCLAK.UNF0"); G_glo_wetland.read(in + "/G_GLOWET.UNF0");
    G_loc_wetland.read(in + "/G_LOCWET.UNF0"); G_lake_area.read(in + "/G_LAKAREA.UNF0"); G_reservoir_area_full.read(in + "/G_RESAREA.UNF0");
    G_reg_lake.read(in + "/G_REGLAKE.UNF0"); G_reg_lake_status.read(in + "/G_REG_LAKE.UNF1");
    for (int n = 0; n < ng; n++)
        for (auto *g : {&G_glo_lake, &G_glo_wetland, &G_loc_lake, &G_loc_wetland, &G_lake_area, &G_reservoir_area_full, &G_reg_lake})
            if ((*g)[n] < 0.) (*g)[n] = 0.;
    G_downstreamCell.read(rd + "/G_OUTFLC.UNF4");
    G_loc_res.read(in + "/G_LOCRES.UNF0");
    for (int n = 0; n < ng; n++) G_loc_lake[n] += G_loc_res[n];
    G_res_type.read(in + "/G_RES_TYPE.UNF1"); G_res_start_year.read(in + "/G_START_YEAR.UNF4");
    G_start_month.read(rd + "/G_START_MONTH.UNF1"); G_mean_outflow.read(in + "/G_MEAN_OUTFLOW.UNF0");
    G_stor_cap_full.read(in + "/G_STORAGE_CAPACITY.UNF0");
    Grid<double> mean_NUs(ng);
    mean_NUs.read(in + "/G_NUs_1971_2000.UNF0");
    Grid<double, 5> &alloc = G_alloc_coeff;
    alloc.initialize(ng);
    alloc.read(rd + "/G_ALLOC_COEFF.5.UNF0");
    for (auto *g : {&G_fractreturngw_irrig, &G_totalUnsatisfiedUse, &G_UnsatisfiedUsePrevYear, &G_unsatisfiedNAsFromIrrig, &G_unsatisfiedNAsFromIrrigPrevYear,
                    &G_unsatisfiedNAsFromOtherSectors, &G_unsatisfiedNAsFromOtherSectorsPrevYear, &G_reducedReturnFlow, &G_reducedReturnFlowPrevYear,
                    &G_dailyRemainingUse, &G_withdrawalIrrigFromSwb, &G_consumptiveUseIrrigFromSwb, &G_actualUse})
        g->initialize(ng);
    if (eng.options.subtract_use > 0) G_fractreturngw_irrig.read(in + "/G_FRACTRETURNGW_IRRIG.UNF0");  // :541-543
    for (int n = 0; n < ng; n++) {  // :361-377 (incl. the reference's un-decremented index into G_reservoir_area_full)
        G_mean_demand[n] = mean_NUs[n];
        short i = 0;
        int d = G_downstreamCell[n];
        while (i < 5 && d > 0 && d < ng && G_reservoir_area_full[d] <= 0) {
            G_mean_demand[n] += mean_NUs[d - 1] * alloc(n, i++);
            d = G_downstreamCell[d - 1];
        }
    }
    for (int n = 0; n < ng; n++) {
        if (G_stor_cap_full[n] < 0.) G_stor_cap_full[n] = 0.;
        K_release[n] = (additionalOutIn.additionalfilestatus == 0) ? 0.1 : additionalOutIn.additionalOutputInput(n, 5);  // :388-392
        G_mean_outflow[n] = G_mean_outflow[n] * 12. * 1000000000. / 31536000.;
        G_mean_demand[n] = G_mean_demand[n] / 31536000.;
    }
    G_riverLength.read(rd + "/G_RIVER_LENGTH.UNF0");
    for (int n = 0; n < ng; n++) G_riverLength[n] *= eng.geo.G_contfreq[n] / 100.;
    G_RiverSlope.read(rd + "/G_RIVERSLOPE.UNF0"); G_Roughness.read(in + "/G_ROUGHNESS.UNF0"); G_bankfull_flow.read(in + "/G_BANKFULL.UNF0");
    for (int n = 0; n < ng; n++) {  // :492-507
        if (G_bankfull_flow[n] < 0.05) G_bankfull_flow[n] = 0.05;
        G_RiverWidth_bf[n] = 2.71 * pow(G_bankfull_flow[n], 0.557);
        G_RiverDepth_bf[n] = 0.349 * pow(G_bankfull_flow[n], 0.341);
        G_riverBottomWidth[n] = G_RiverWidth_bf[n] - 2.0 * 2.0 * G_RiverDepth_bf[n];
        G_riverStorageMax[n] = G_riverLength[n] * 0.5 * G_RiverDepth_bf[n] / 1000. * (G_riverBottomWidth[n] / 1000. + G_RiverWidth_bf[n] / 1000.);
    }
    G_routOrder.read(rd + "/G_ROUT_ORDER.UNF4");
    G_LDD.read(rd + "/G_LDD_2.UNF1");
}

void routingClass::initLakeDepthActive(const calibParamClass &cp) { for (int n = 0; n < eng.ncell; n++) G_lakeDepthActive[n] = cp.getValue(P_LAK_D, n) * 0.001; }
void routingClass::initWetlDepthActive(const calibParamClass &cp) { for (int n = 0; n < eng.ncell; n++) G_wetlDepthActive[n] = cp.getValue(P_WET_D, n) * 0.001; }

void routingClass::initFractionStatus() {  // :745-765
    for (int n = 0; n < eng.ncell; n++) {
        statusStarted_landfreq[n] = 0;
        statusStarted_landAreaFracNextTimestep[n] = 0;
        G_fLocLake[n] = G_loc_lake[n] / 100.;
        G_fLocWet[n] = G_loc_wetland[n] / 100.;
        G_fGloLake[n] = G_glo_lake[n] / 100.;
        G_fGloWet[n] = G_glo_wetland[n] / 100.;
        G_fswbInit[n] = G_fLocLake[n] + G_fLocWet[n] + G_fGloWet[n];
        G_fswbLandAreaFrac[n] = G_fswbInit[n];
        G_fswbLandAreaFracNextTimestep[n] = G_fswbLandAreaFrac[n];
    }
    statusStarted_updateGloResPrevYear = 0;
}

void routingClass::initFractionStatusAdditionalOI(AdditionalOutputInputFile &additionalOutIn) {  // :767-787
    for (int n = 0; n < eng.ncell; n++) {
        statusStarted_landfreq[n] = 0;
        statusStarted_landAreaFracNextTimestep[n] = 0;
        G_fLocLake[n] = additionalOutIn.additionalOutputInput(n, 45);
        G_fLocWet[n] = additionalOutIn.additionalOutputInput(n, 47);
        G_fGloLake[n] = G_glo_lake[n] / 100.;
        G_fGloWet[n] = additionalOutIn.additionalOutputInput(n, 46);
        G_fswbInit[n] = additionalOutIn.additionalOutputInput(n, 14);
        G_fswbLandAreaFrac[n] = G_fswbInit[n];
        G_fswbLandAreaFracNextTimestep[n] = G_fswbLandAreaFrac[n];
    }
    statusStarted_updateGloResPrevYear = 0;
}

// First day after a checkpoint (integrateWGHM.cpp:311-316): the reduction factors are re-derived from the restored storages, the
// surface-water-body fractions from them, and the land area fraction of the checkpoint is corrected by the change of those
// fractions against the checkpoint's G_fswbLandAreaFracNextTimestep (column 33).
void routingClass::update_landarea_red_fac_PDAF(calibParamClass &calParam, AdditionalOutputInputFile &additionalOutIn) {
    const double evapoReductionExp = 3.32193, evapoReductionExpReservoir = 2.81383;
    auto clamp01 = [](double x) { return x < 0. ? 0. : (x > 1. ? 1. : x); };
    for (int n = 0; n < eng.ncell; n++) {  // (the reference walks the cells in routing order; the statements are per cell)
        const double ex = calParam.getValue(M_EVAREDEX, n) * evapoReductionExp;
        const double cellArea = eng.geo.areaOfCellByArrayPos(n);
        if (G_loc_lake[n] > 0.) {
            const double maxStorage = ((G_loc_lake[n]) / 100.) * cellArea * G_lakeDepthActive[n];
            G_locLakeAreaReductionFactor[n] = clamp01(1. - pow(fabs(G_locLakeStorage[n] - maxStorage) / (2. * maxStorage), ex));
        }
        if (G_loc_wetland[n] > 0.) {
            const double maxStorage = ((G_loc_wetland[n]) / 100.) * cellArea * G_wetlDepthActive[n];
            G_locWetlAreaReductionFactor[n] = clamp01(1. - pow(fabs(G_locWetlStorage[n] - maxStorage) / (maxStorage), ex));
        }
        if (G_glo_wetland[n] > 0.) {
            const double maxStorage = ((G_glo_wetland[n]) / 100.) * cellArea * G_wetlDepthActive[n];
            G_gloWetlAreaReductionFactor[n] = clamp01(1. - pow(fabs(G_gloWetlStorage[n] - maxStorage) / maxStorage, ex));
        }
        if (G_lake_area[n] > 0.) {
            const double maxStorage = ((double)G_lake_area[n]) * G_lakeDepthActive[n];
            G_gloLakeEvapoReductionFactor[n] = clamp01(1. - pow(fabs(G_gloLakeStorage[n] - maxStorage) / (2. * maxStorage), ex));
        }
        if (G_reservoir_area[n] > 0.) {
            const double maxStorage = G_stor_cap[n];
            G_gloResEvapoReductionFactor[n] = clamp01(1. - pow(fabs(G_gloResStorage[n] - maxStorage) / maxStorage, evapoReductionExpReservoir));
        }
        G_fLocLake[n] = ((G_loc_lake[n] > 0.) && (G_locLakeAreaReductionFactor[n] > 0.)) ? (G_locLakeAreaReductionFactor[n] * G_loc_lake[n] / 100.) : 0.;
        G_fLocWet[n] = ((G_loc_wetland[n] > 0.) && (G_locWetlAreaReductionFactor[n] > 0.)) ? (G_locWetlAreaReductionFactor[n] * G_loc_wetland[n] / 100.) : 0.;
        G_fGloWet[n] = ((G_glo_wetland[n] > 0.) && (G_gloWetlAreaReductionFactor[n] > 0.)) ? (G_gloWetlAreaReductionFactor[n] * G_glo_wetland[n] / 100.) : 0.;
        G_fswbLandAreaFracNextTimestep[n] = G_fLocLake[n] + G_fLocWet[n] + G_fGloWet[n];
        G_fswbLandAreaFrac[n] = additionalOutIn.additionalOutputInput(n, 33);
        const double changePct = G_fswbLandAreaFracNextTimestep[n] * 100. - G_fswbLandAreaFrac[n] * 100.;
        G_landAreaFrac[n] = additionalOutIn.additionalOutputInput(n, 6);
        G_landAreaFrac[n] = G_landAreaFrac[n] - (changePct);
        if (G_landAreaFrac[n] < 0.) G_landAreaFrac[n] = 0.;
        G_landAreaFracPrevTimestep[n] = additionalOutIn.additionalOutputInput(n, 7);
        statusStarted_landfreq[n] = 1;
    }
}

void routingClass::setStoragesToZero() {  // :789-847 (riverveloOpt 1: rivers start empty)
    for (auto *g : {&G_locLakeStorage, &G_locWetlStorage, &G_gloLakeStorage, &G_gloWetlStorage, &G_gloResStorage, &G_riverStorage, &G_groundwaterStorage})
        g->fill(0.);
}

void routingClass::setStorages(WghmStateFile &st, AdditionalOutputInputFile &add) {  // :851-882
    for (int n = 0; n < eng.ncell; n++) {
        G_totalUnsatisfiedUse[n] = add.additionalOutputInput(n, 3);
        G_UnsatisfiedUsePrevYear[n] = add.additionalOutputInput(n, 4);
        G_reducedReturnFlow[n] = add.additionalOutputInput(n, 38);
        G_reducedReturnFlowPrevYear[n] = add.additionalOutputInput(n, 50);
        G_unsatisfiedNAsFromIrrig[n] = add.additionalOutputInput(n, 39);
        G_unsatisfiedNAsFromIrrigPrevYear[n] = add.additionalOutputInput(n, 51);
        G_unsatisfiedNAsFromOtherSectors[n] = add.additionalOutputInput(n, 49);
        G_unsatisfiedNAsFromOtherSectorsPrevYear[n] = add.additionalOutputInput(n, 52);
        // what routing() reloads on the first day after a checkpoint (:1706-1743), as far as use_alloc 0 has it
        G_dailyRemainingUse[n] = add.additionalOutputInput(n, 35);
        G_withdrawalIrrigFromSwb[n] = add.additionalOutputInput(n, 25);
        G_consumptiveUseIrrigFromSwb[n] = add.additionalOutputInput(n, 26);
        const double f = ((eng.geo.areaOfCellByArrayPos(n) * (eng.geo.G_contfreq[n] / 100.)) / 1000000.);
        Cell &c = st.cell(n);
        G_locLakeStorage[n] = c.locallake(0) * f; G_locWetlStorage[n] = c.localwetland(0) * f; G_gloLakeStorage[n] = c.globallake(0) * f;
        G_gloWetlStorage[n] = c.globalwetland(0) * f; G_riverStorage[n] = c.river(0) * f; G_groundwaterStorage[n] = c.groundwater(0) * f;
        G_gloResStorage[n] = c.reservoir(0) * f;
    }
}

void routingClass::setLakeWetlToMaximum(short start_year) {  // :5647-5720
    const int ref = eng.options.resYearReference;
    const bool byYear = eng.options.resYearOpt == 1;
    for (int n = 0; n < eng.ncell; n++) {
        const double A = eng.geo.areaOfCellByArrayPos(n);
        G_locLakeStorage[n] = (G_loc_lake[n] / 100.) * A * G_lakeDepthActive[n];
        G_locWetlStorage[n] = (G_loc_wetland[n] / 100.) * A * G_wetlDepthActive[n];
        G_gloLakeStorage[n] = G_lake_area[n] * G_lakeDepthActive[n];
        G_gloWetlStorage[n] = (G_glo_wetland[n] / 100.) * A * G_wetlDepthActive[n];
        if (!byYear && ref >= G_res_start_year[n] && G_stor_cap_full[n] > -99) G_gloResStorage[n] = G_stor_cap_full[n];
        // the second if/else of the reference (:5678-5683) decides the factor: 0 with resYearOpt 0, 1 for the reservoirs that
        // operate at the start of a resYearOpt 1 run
        if (byYear && start_year >= G_res_start_year[n] && G_stor_cap_full[n] > -99) {
            G_gloResEvapoReductionFactor[n] = 1.;
            G_gloResStorage[n] = G_stor_cap_full[n];
        } else
            G_gloResEvapoReductionFactor[n] = 0.;
        if (G_reg_lake_status[n] == 1 && (byYear ? start_year : ref) < G_res_start_year[n])
            G_gloResStorage[n] += G_reservoir_area_full[n] * G_lakeDepthActive[n];
        G_locLakeAreaReductionFactor[n] = 1.; G_locWetlAreaReductionFactor[n] = 1.; G_gloLakeEvapoReductionFactor[n] = 1.; G_gloWetlAreaReductionFactor[n] = 1.;
        G_riverAreaReductionFactor[n] = 0.5;
        G_riverAreaFracNextTimestep_Frac[n] = G_riverAreaReductionFactor[n] * G_riverLength[n] * G_RiverWidth_bf[n] / 1000. / A;
    }
}

// routing.cpp:979-1415 without the output arrays: the year's reservoir statics (resYearOpt 0: those of the reference year in every
// year; 1: reservoirs and their land-cover fraction G_RES_<year> come on line in their start year), the first year's land area
// fraction, and - parts 3/4 - the land area and the water stored on it that a new reservoir takes.  Works on the host grids
// (synchronised with the device at every month end, run_model); yearlyChanged tells the caller what to send back.
void routingClass::annualInit(short year, int start_month, AdditionalOutputInputFile &additionalOutIn) {
    const int ng = eng.ncell;
    const optionClass &o = eng.options;
    if (additionalOutIn.additionalfilestatus == 0 || start_month == 1 || year > o.start_year)  // :986-1008
        for (int n = 0; n < ng; n++) {
            G_UnsatisfiedUsePrevYear[n] = G_totalUnsatisfiedUse[n];
            G_unsatisfiedNAsFromIrrigPrevYear[n] = G_unsatisfiedNAsFromIrrig[n];
            G_unsatisfiedNAsFromOtherSectorsPrevYear[n] = G_unsatisfiedNAsFromOtherSectors[n];
            G_reducedReturnFlowPrevYear[n] = G_reducedReturnFlow[n];
        }
    const Grid<double> area_before = G_reservoir_area, cap_before = G_stor_cap, lake_before = G_lake_area;
    for (int n = 0; n < ng; n++) { G_reservoir_area[n] = 0.; G_stor_cap[n] = 0.; }
    for (int n = 0; n < ng; n++)
        if (G_reg_lake_status[n] == 1) { G_reservoir_area[n] = G_reservoir_area_full[n]; G_stor_cap[n] = G_stor_cap_full[n]; }
    int resYear = o.resYearReference;  // :1052-1078
    if (o.resYearOpt == 1) resYear = year > o.resYearLastToUse ? o.resYearLastToUse : year < o.resYearFirstToUse ? o.resYearFirstToUse : year;
    G_glo_res.read(o.input_dir + "/G_RES/G_RES_" + std::to_string(resYear) + ".UNF0");
    if (0 == statusStarted_updateGloResPrevYear) { updateGloResPrevYear_pct(); statusStarted_updateGloResPrevYear = 1; }
    for (int n = 0; n < ng; n++)
        if (resYear >= G_res_start_year[n]) { G_reservoir_area[n] = G_reservoir_area_full[n]; G_stor_cap[n] = G_stor_cap_full[n]; }
    for (int n = 0; n < ng; n++)
        if (G_reservoir_area[n] > 0. && resYear >= G_res_start_year[n] && ((G_res_type[n] + 0 == 0) || G_mean_outflow[n] <= 0.)) {
            G_lake_area[n] += G_reservoir_area_full[n];  // treated as a global lake (:1107-1146)
            G_reservoir_area[n] = 0.;
            G_reservoir_area_full[n] = 0.;
        }
    yearlyChanged = false;
    for (int n = 0; n < ng; n++)
        if (G_reservoir_area[n] != area_before[n] || G_stor_cap[n] != cap_before[n] || G_lake_area[n] != lake_before[n]) yearlyChanged = true;
    // part 3 (:1180-1291), first year only (G_fwaterfreq / G_landfreq / G_landWaterExclGloLakAreaFrac feed output grids only)
    for (int n = 0; n < ng; n++) {
        if (0 == statusStarted_landfreq[n]) {
            G_landAreaFrac[n] = (eng.geo.G_contfreq[n] - (G_glo_lake[n] + G_glo_wetland[n] + G_loc_lake[n] + G_loc_wetland[n] + G_glo_res[n]));
            G_landAreaFracPrevTimestep[n] = 0.;
            if (G_landAreaFrac[n] < 0.) G_landAreaFrac[n] = 0.;
            statusStarted_landfreq[n] = 1;
        }
    }
    // part 4 (:1296-1415): a reservoir fraction that grew against the previous year takes land area and the water stored on it.
    // With resYearOpt 0 the fraction is the reference year's in every year and nothing changes; the previous year's fraction of
    // a checkpoint comes back from column 44.
    auto &d = eng.dailyWaterBalance;
    for (int n = 0; n < ng; n++) {
        if (additionalOutIn.additionalfilestatus == 1) G_glores_prevyear[n] = additionalOutIn.additionalOutputInput(n, 44);
        if (!(start_month == 1 || year > o.start_year)) continue;
        const double glores_change = G_glo_res[n] - G_glores_prevyear[n];
        if (!(glores_change > 0.)) continue;
        yearlyChanged = true;
        if (G_glores_prevyear[n] > 0.) G_landAreaFrac[n] = G_landAreaFrac[n] - glores_change;  // an existing reservoir grew
        else G_landAreaFrac[n] = G_landAreaFrac[n] - G_glo_res[n];                              // a new one
        if (G_landAreaFrac[n] < 0.) G_landAreaFrac[n] = 0.;
        // net change against the day before yesterday; not negative = fractional routing only: no change at all
        const double laf_change = G_landAreaFrac[n] - G_landAreaFracPrevTimestep[n];
        if (laf_change >= 0.) {
            G_landAreaFrac[n] = G_landAreaFracPrevTimestep[n];
        } else {
            // the water on the lost land (mm over the cell) goes to the reservoir of the outflow cell, km3
            const double A = eng.geo.areaOfCellByArrayPos(n);
            const double canopy_km3 = d.G_canopyWaterContent[n] * A / 1000000.0 * ((-laf_change) / 100.0);
            const double soil_km3 = d.G_soilWaterContent[n] * A / 1000000.0 * ((-laf_change) / 100.0);
            const double snow_km3 = d.G_snow[n] * A / 1000000.0 * ((-laf_change) / 100.0);
            G_gloResStorage[G_outflow_cell_assignment[n] - 1] += canopy_km3 + soil_km3 + snow_km3;
            G_landAreaFracNextTimestep[n] = G_landAreaFrac[n];  // no second rescaling in calcNewDay
        }
    }
}

// what annualInit may have changed, back to the device (the statics re-derive the kernels' per-cell constants on the next call)
void routingClass::pushYearly() {
    Engine &e = eng;
    e.set("lake_area", G_lake_area); e.set("reservoir_area", G_reservoir_area); e.set("stor_cap", G_stor_cap);
    e.set("land_area_frac", G_landAreaFrac); e.set("land_area_frac_next", G_landAreaFracNextTimestep); e.set("res_stor", G_gloResStorage);
}

// ---- water use (subtract_use 2) --------------------------------------------------------------
void routingClass::dailyNUInit(const std::string &dir, short new_year, calibParamClass &calParam) {  // :884-977 (aggrNUsGloLakResOpt 0)
    const int ng = eng.ncell;
    const short numberOfDaysInMonth[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    for (auto *g : {&G_dailyNUs, &G_dailyNUg, &G_monthlyWUIrrigFromSwb, &G_monthlyCUIrrigFromSwb}) g->initialize(ng);
    const std::string y = std::to_string(new_year);
    G_dailyNUs.read(dir + "/G_NETUSE_SW_m3_" + y + ".12.UNF0");
    G_dailyNUg.read(dir + "/G_NETUSE_GW_m3_" + y + ".12.UNF0");
    G_monthlyWUIrrigFromSwb.read(dir + "/G_IRRIG_WITHDRAWAL_USE_SW_m3_" + y + ".12.UNF0");
    G_monthlyCUIrrigFromSwb.read(dir + "/G_IRRIG_CONS_USE_SW_m3_" + y + ".12.UNF0");
    for (int n = 0; n < ng; n++)
        for (short month = 0; month < 12; month++) {
            G_dailyNUs(n, month) = calParam.getValue(M_NETABSSW, n) * G_dailyNUs(n, month);
            G_dailyNUg(n, month) = calParam.getValue(M_NETABSGW, n) * G_dailyNUg(n, month);
        }
    for (short month = 0; month <= 11; month++)
        for (int n = 0; n <= ng - 1; n++) {
            G_dailyNUs(n, month) = G_dailyNUs(n, month) / (1000000000. * (double)numberOfDaysInMonth[month]);
            G_dailyNUg(n, month) = G_dailyNUg(n, month) / (1000000000. * (double)numberOfDaysInMonth[month]);
        }
}

void routingClass::pushWaterUseMonth(short month) {  // what calcNextDay_M (:7432-7440) and :3907-3908 read on every day of the month
    const int ng = eng.ncell;
    const short numberOfDaysInMonth[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    Grid<double> nus(ng), nug(ng), wusi(ng), cusi(ng);
    for (int n = 0; n < ng; n++) {
        nus[n] = G_dailyNUs(n, month);
        nug[n] = G_dailyNUg(n, month);
        wusi[n] = G_monthlyWUIrrigFromSwb(n, month) / (1000000000. * (double)numberOfDaysInMonth[month]);
        cusi[n] = G_monthlyCUIrrigFromSwb(n, month) / (1000000000. * (double)numberOfDaysInMonth[month]);
    }
    eng.set("wu_nus_month", nus); eng.set("wu_nug_month", nug); eng.set("wu_wusi_month", wusi); eng.set("wu_cusi_month", cusi);
}

void routingClass::pullWaterUse() {
    Engine &e = eng;
    e.get("wu_total_unsatisfied", G_totalUnsatisfiedUse); e.get("wu_uns_irr", G_unsatisfiedNAsFromIrrig);
    e.get("wu_uns_oth", G_unsatisfiedNAsFromOtherSectors); e.get("wu_red_rf", G_reducedReturnFlow);
    e.get("wu_daily_remaining", G_dailyRemainingUse); e.get("wu_wusi", G_withdrawalIrrigFromSwb);
    e.get("wu_cusi", G_consumptiveUseIrrigFromSwb); e.get("wu_actual_use", G_actualUse);
}

void routingClass::pushWaterUseState() {
    Engine &e = eng;
    e.set("wu_total_unsatisfied", G_totalUnsatisfiedUse); e.set("wu_uns_irr", G_unsatisfiedNAsFromIrrig);
    e.set("wu_uns_oth", G_unsatisfiedNAsFromOtherSectors); e.set("wu_red_rf", G_reducedReturnFlow);
    e.set("wu_daily_remaining", G_dailyRemainingUse); e.set("wu_wusi", G_withdrawalIrrigFromSwb);
    e.set("wu_cusi", G_consumptiveUseIrrigFromSwb); e.set("wu_actual_use", G_actualUse);
}

void routingClass::annualWaterUsePostProcessing(short, AdditionalOutputInputFile &add) {  // :5246-5306 (host grids synchronised by the caller)
    for (int n = 0; n <= eng.ncell - 1; n++) {
        if (G_totalUnsatisfiedUse[n] < G_UnsatisfiedUsePrevYear[n]) {
            G_UnsatisfiedUsePrevYear[n] = G_totalUnsatisfiedUse[n];
            G_unsatisfiedNAsFromIrrigPrevYear[n] = G_unsatisfiedNAsFromIrrig[n];
            G_unsatisfiedNAsFromOtherSectorsPrevYear[n] = G_unsatisfiedNAsFromOtherSectors[n];
            G_reducedReturnFlowPrevYear[n] = G_reducedReturnFlow[n];
            G_unsatisfiedNAsFromIrrig[n] = 0.;
            G_unsatisfiedNAsFromOtherSectors[n] = 0.;
            G_reducedReturnFlow[n] = 0.;
            G_totalUnsatisfiedUse[n] = 0.;
        } else {
            G_totalUnsatisfiedUse[n] -= G_UnsatisfiedUsePrevYear[n];
            G_unsatisfiedNAsFromIrrig[n] -= G_unsatisfiedNAsFromIrrigPrevYear[n];
            G_unsatisfiedNAsFromOtherSectors[n] -= G_unsatisfiedNAsFromOtherSectorsPrevYear[n];
            G_reducedReturnFlow[n] -= G_reducedReturnFlowPrevYear[n];
            if ((G_unsatisfiedNAsFromIrrig[n] + G_unsatisfiedNAsFromOtherSectors[n]) < 0) {
                G_unsatisfiedNAsFromIrrig[n] = 0;
                G_unsatisfiedNAsFromOtherSectors[n] = 0;
                G_reducedReturnFlow[n] = 0;
            } else if (G_unsatisfiedNAsFromOtherSectors[n] < 0) {
                G_unsatisfiedNAsFromIrrig[n] += G_unsatisfiedNAsFromOtherSectors[n];
                G_unsatisfiedNAsFromOtherSectors[n] = 0;
            } else if (G_unsatisfiedNAsFromIrrig[n] < 0) {
                G_unsatisfiedNAsFromOtherSectors[n] += G_unsatisfiedNAsFromIrrig[n];
                G_unsatisfiedNAsFromIrrig[n] = 0;
                G_reducedReturnFlow[n] = 0;
            }
            if (G_reducedReturnFlow[n] > 0) {
                G_reducedReturnFlow[n] = 0;
                G_unsatisfiedNAsFromIrrig[n] = 0;
            }
        }
        add.additionalOutputInput(n, 3) = G_totalUnsatisfiedUse[n];
        add.additionalOutputInput(n, 4) = G_UnsatisfiedUsePrevYear[n];
        add.additionalOutputInput(n, 50) = G_reducedReturnFlowPrevYear[n];
        add.additionalOutputInput(n, 51) = G_unsatisfiedNAsFromIrrigPrevYear[n];
        add.additionalOutputInput(n, 52) = G_unsatisfiedNAsFromOtherSectorsPrevYear[n];
        add.additionalOutputInput(n, 38) = G_reducedReturnFlow[n];
        add.additionalOutputInput(n, 39) = G_unsatisfiedNAsFromIrrig[n];
        add.additionalOutputInput(n, 49) = G_unsatisfiedNAsFromOtherSectors[n];
    }
}

void routingClass::routing(short, short day, short month, short dom, short, WghmStateFile &st, AdditionalOutputInputFile &, short, calibParamClass &) {
    eng.check(wgk_routing_day(eng.ctx, day, month, dom), "wgk_routing_day");
    // wghmState of the day (routing.cpp:5002-5020), packed on the device: one copy of [7][ncell] instead of 22 field downloads;
    // the public grids are synchronised lazily (pull()) by whoever reads them
    day_state.resize((size_t)7 * eng.ncell);
    eng.check(wgk_get_day_state(eng.ctx, 0, day_state.data()), "wgk_get_day_state");
    const size_t ng = (size_t)eng.ncell;
    for (size_t n = 0; n < ng; n++) {
        Cell &c = st.cell(n);
        c.locallake(dom - 1) = day_state[0 * ng + n]; c.localwetland(dom - 1) = day_state[1 * ng + n];
        c.globallake(dom - 1) = day_state[2 * ng + n]; c.globalwetland(dom - 1) = day_state[3 * ng + n];
        c.reservoir(dom - 1) = day_state[4 * ng + n]; c.river(dom - 1) = day_state[5 * ng + n];
        c.groundwater(dom - 1) = day_state[6 * ng + n];
    }
}

void routingClass::updateLandAreaFrac(AdditionalOutputInputFile &) {
    // fused into the routing post-pass on the device; columns 6 / 7 of the checkpoint (:5347-5350) are filled from the
    // synchronised grids when a checkpoint is due (fill_additional)
    eng.check(wgk_update_land_area_frac(eng.ctx), "wgk_update_land_area_frac");
}

void routingClass::pull() {
    Engine &e = eng;
    e.get("loc_lake_stor", G_locLakeStorage); e.get("loc_wetl_stor", G_locWetlStorage); e.get("glo_lake_stor", G_gloLakeStorage);
    e.get("glo_wetl_stor", G_gloWetlStorage); e.get("res_stor", G_gloResStorage); e.get("river_stor", G_riverStorage);
    e.get("gw", G_groundwaterStorage); e.get("land_area_frac", G_landAreaFrac); e.get("land_area_frac_prev", G_landAreaFracPrevTimestep);
    e.get("land_area_frac_next", G_landAreaFracNextTimestep); e.get("status_laf_next", statusStarted_landAreaFracNextTimestep);
    e.get("red_loc_lake", G_locLakeAreaReductionFactor); e.get("red_loc_wetl", G_locWetlAreaReductionFactor);
    e.get("red_glo_lake", G_gloLakeEvapoReductionFactor); e.get("red_glo_wetl", G_gloWetlAreaReductionFactor);
    e.get("red_res", G_gloResEvapoReductionFactor); e.get("red_river", G_riverAreaReductionFactor); e.get("k_release", K_release);
    e.get("fswb_laf", G_fswbLandAreaFrac); e.get("fswb_laf_next", G_fswbLandAreaFracNextTimestep);
    e.get("river_area_frac_next", G_riverAreaFracNextTimestep_Frac); e.get("discharge", G_riverDischarge);
}

// ------------------------------------------------------------------------------------------
// daily class
// ------------------------------------------------------------------------------------------
void dailyWaterBalanceClass::setStoragesToZero() {
    G_soilWaterContent.fill(0.); G_canopyWaterContent.fill(0.); G_snow.fill(0.); G_dailyStorageTransfer.fill(0.); G_SnowInElevation.fill(0.);
}

void dailyWaterBalanceClass::setStorages(WghmStateFile &st, SnowInElevationFile &snow, AdditionalOutputInputFile &) {  // :1896-1924
    for (int n = 0; n < eng.ncell; n++) {
        const double laf = eng.routing.getLandAreaFrac(n);
        if (laf <= 0.) {
            G_canopyWaterContent[n] = G_snow[n] = G_soilWaterContent[n] = 0.;
            for (int e = 0; e <= 100; e++) G_SnowInElevation(n, e) = 0.;
        } else {
            const double cf = eng.geo.G_contfreq[n];
            G_canopyWaterContent[n] = st.cell(n).canopy(0) * cf / laf;
            G_snow[n] = st.cell(n).snow(0) * cf / laf;
            G_soilWaterContent[n] = st.cell(n).soil(0) * cf / laf;
            for (int e = 0; e <= 100; e++) G_SnowInElevation(n, e) = snow.snowInElevation(n, e) * cf / laf;
        }
    }
}

void dailyWaterBalanceClass::calcNewDayAll(short day, short month, short dom) {
    // forcing slot of the day: the month's .31 grids lie in slots 0..30, the year's .365 grids in slots 0..364
    const int slot = (eng.options.time_series == 1) ? day - 1 : dom - 1;
    eng.check(wgk_vertical_day(eng.ctx, day, month, dom, slot), "wgk_vertical_day");
}

void dailyWaterBalanceClass::calcNewDay(short day, short month, short dom, short, short year, int, WghmStateFile &, AdditionalOutputInputFile &,
                                        SnowInElevationFile &, short, calibParamClass &) {
    const int key = (int)year * 1000 + day;  // one whole-grid launch per simulated day
    if (key == last_day_launched) return;
    last_day_launched = key;
    calcNewDayAll(day, month, dom);
}

void dailyWaterBalanceClass::pull() {
    Engine &e = eng;
    e.get("canopy", G_canopyWaterContent); e.get("soil", G_soilWaterContent); e.get("snow", G_snow); e.get("snow_bands", G_SnowInElevation);
    e.get("lake_balance", G_lakeBalance); e.get("openwater_pet", G_openWaterPET); e.get("openwater_prec", G_openWaterPrec);
    e.get("surface_runoff", G_dailyLocalSurfaceRunoff); e.get("gw_recharge", G_dailyGwRecharge); e.get("storage_transfer", G_dailyStorageTransfer);
}

// ------------------------------------------------------------------------------------------
// device synchronisation
// ------------------------------------------------------------------------------------------
void Engine::push_static() {
    // class key per cell (which water-body code the cell needs): lets the library store cells of equal class
    // next to each other inside a routing level; results do not depend on it (include/wgk.h)
    // High 4 bits (opt-in, WGK_COLD_BINS=1..16; measured without gain, see watergap2_b200.cell_classes): a coldness bin from
    // latitude and mean elevation (coldest first; 27 - 0.55 |lat| - 0.0045 m as a climatological mean temperature - it only
    // orders cells), so that the cells of a warp / CTA tend to be either all in the 100-band snow loop or all out of it
    std::vector<uint8_t> cls(ncell);
    const char *cb = getenv("WGK_COLD_BINS");
    const int nbins = cb ? std::min(16, atoi(cb)) : 0;
    for (int n = 0; n < ncell; n++) {
        int key = (routing.G_loc_lake[n] > 0.) * 1 + (routing.G_loc_wetland[n] > 0.) * 2
                  + ((routing.G_lake_area[n] > 0.) || (routing.G_reservoir_area_full[n] > 0.) || (routing.G_glo_wetland[n] > 0.)) * 4
                  + (G_aindex[n] == 1) * 8;
        if (nbins > 0) {
            const double lat = 90.25 - 0.5 * (double)geo.G_row[n];
            const double t = 27.0 - 0.55 * std::fabs(lat) - 0.0045 * (double)dailyWaterBalance.G_Elevation(n, 0);
            const int bin = (int)std::floor((t + 30.0) / 60.0 * nbins);
            key += 16 * std::max(0, std::min(nbins - 1, bin));
        }
        cls[n] = (uint8_t)key;
    }
    check(wgk_set_cell_classes(ctx, cls.data()), "wgk_set_cell_classes");
    check(wgk_set_topology(ctx, routing.G_routOrder.data(), routing.G_downstreamCell.data()), "wgk_set_topology");
    Grid<double> area(ncell);
    for (int n = 0; n < ncell; n++) area[n] = geo.areaOfCellByArrayPos(n);
    set("area", area); set("contfreq", geo.G_contfreq); set("contcell", geo.G_contcell); set("row", geo.G_row);
    set("toBeCalculated", G_toBeCalculated); set("landcover", G_landCover); set("builtup", G_built_up); set("arid", G_aindex);
    set("ldd", routing.G_LDD); set("texture", G_texture); set("elevation", dailyWaterBalance.G_Elevation);
    set("loc_lake", routing.G_loc_lake); set("loc_wetland", routing.G_loc_wetland); set("glo_wetland", routing.G_glo_wetland);
    set("lake_area", routing.G_lake_area); set("reservoir_area", routing.G_reservoir_area); set("stor_cap", routing.G_stor_cap);
    set("mean_outflow", routing.G_mean_outflow); set("mean_demand", routing.G_mean_demand); set("res_type", routing.G_res_type);
    set("start_month", routing.G_start_month); set("river_length", routing.G_riverLength); set("river_slope", routing.G_RiverSlope);
    set("roughness", routing.G_Roughness); set("river_bottom_width", routing.G_riverBottomWidth); set("river_width_bf", routing.G_RiverWidth_bf);
    set("river_storage_max", routing.G_riverStorageMax); set("fswb_init", routing.G_fswbInit); set("f_glo_lake", routing.G_fGloLake);
    check(wgk_set_field(ctx, wgk_field_id("params"), 0, calParam.block(), (size_t)26 * ncell * sizeof(double)), "params");
    set("gamma_hbv", dailyWaterBalance.G_gammaHBV); set("cfa", dailyWaterBalance.G_cellCorrFact); set("cfs", routing.G_statCorrFact);
    set("smax", G_Smax); set("gwfactor", G_gwFactor); set("rgmax", G_Rgmax); set("laimax", G_LAImax);
    set("lake_depth_active", routing.G_lakeDepthActive); set("wetl_depth_active", routing.G_wetlDepthActive);
    auto table = [&](const char *name, const void *p, size_t bytes) { check(wgk_set_field(ctx, wgk_field_id(name), 0, p, bytes), name); };
    table("lai_factor_a", lai_factor_a, sizeof lai_factor_a); table("lai_factor_b", lai_factor_b, sizeof lai_factor_b);
    table("lai_initial_days", lai_initialDays, sizeof lai_initialDays); table("lai_kc_min", kc_min, sizeof kc_min); table("lai_kc_max", kc_max, sizeof kc_max);
    table("lct_albedo", dailyWaterBalance.albedo_lct, sizeof dailyWaterBalance.albedo_lct);
    table("lct_albedo_snow", dailyWaterBalance.albedoSnow_lct, sizeof dailyWaterBalance.albedoSnow_lct);
    table("lct_ddf", dailyWaterBalance.ddf_lct, sizeof dailyWaterBalance.ddf_lct);
    table("lct_emissivity", dailyWaterBalance.emissivity_lct, sizeof dailyWaterBalance.emissivity_lct);
    if (options.subtract_use > 0) {
        set("wu_frgi", routing.G_fractreturngw_irrig);
        set("wu_alloc_coeff", routing.G_alloc_coeff);
    }
}

void Engine::push_state() {
    auto &d = dailyWaterBalance;
    auto &r = routing;
    set("canopy", d.G_canopyWaterContent); set("soil", d.G_soilWaterContent); set("snow", d.G_snow); set("snow_bands", d.G_SnowInElevation);
    set("storage_transfer", d.G_dailyStorageTransfer); set("lai_days", lai_days); set("lai_status", lai_status); set("lai_precsum", lai_precsum);
    set("gw", r.G_groundwaterStorage); set("loc_lake_stor", r.G_locLakeStorage); set("loc_wetl_stor", r.G_locWetlStorage);
    set("glo_lake_stor", r.G_gloLakeStorage); set("glo_wetl_stor", r.G_gloWetlStorage); set("res_stor", r.G_gloResStorage);
    set("river_stor", r.G_riverStorage); set("red_loc_lake", r.G_locLakeAreaReductionFactor); set("red_loc_wetl", r.G_locWetlAreaReductionFactor);
    set("red_glo_lake", r.G_gloLakeEvapoReductionFactor); set("red_glo_wetl", r.G_gloWetlAreaReductionFactor);
    set("red_res", r.G_gloResEvapoReductionFactor); set("red_river", r.G_riverAreaReductionFactor); set("k_release", r.K_release);
    set("land_area_frac", r.G_landAreaFrac); set("land_area_frac_prev", r.G_landAreaFracPrevTimestep);
    set("land_area_frac_next", r.G_landAreaFracNextTimestep); set("fswb_laf", r.G_fswbLandAreaFrac);
    set("fswb_laf_next", r.G_fswbLandAreaFracNextTimestep); set("river_area_frac_next", r.G_riverAreaFracNextTimestep_Frac);
    set("status_laf_next", r.statusStarted_landAreaFracNextTimestep);
}

void Engine::set_forcing_month(int month1, int year) {  // climate.cpp:93-123 (.31 files, cloud == 1)
    const std::string c = options.climate_dir, sfx = "_" + std::to_string(year) + "_" + std::to_string(month1) + ".31.UNF0";
    Grid<float, 31> P(ncell), T(ncell), SW(ncell), LW(ncell);
    T.read(c + "/GTEMP" + sfx); P.read(c + "/GPREC" + sfx); SW.read(c + "/GSHORTWAVE" + sfx); LW.read(c + "/GLONGWAVE_DOWN" + sfx);
    check(wgk_set_forcing(ctx, 0, 31, -1, P.data(), T.data(), SW.data(), LW.data(), 31), "wgk_set_forcing");
    check(wgk_synchronize(ctx), "sync");  // the host grids go out of scope
}

// the year's forcing from the [cell][365] files of climateYear.cpp:38-58 (time_series 1): the BYTES of the four big-endian files
// go to the device as they are, byte order and layout are converted there (wgk_set_forcing_unf)
void Engine::set_forcing_year(int year) {
    const std::string c = options.climate_dir, y = std::to_string(year);
    const std::string files[4] = {c + "/G_GPCC_H08day_V20110128_" + y + ".365.UNF0", c + "/G_TEMP_H08_int_" + y + ".365.UNF0",
                                  c + "/G_SSRD_H08_int_" + y + ".365.UNF0", c + "/G_SLRD_H08_int_" + y + ".365.UNF0"};
    std::vector<char> raw[4];
    const size_t bytes = (size_t)ncell * 365 * sizeof(float);
    for (int k = 0; k < 4; k++) {
        std::ifstream f(files[k], std::ios::binary);
        if (!f) throw std::runtime_error(files[k] + " not found.");
        raw[k].resize(bytes);
        f.read(raw[k].data(), (std::streamsize)bytes);
        if ((size_t)f.gcount() != bytes) throw std::runtime_error(files[k] + ": short read");
    }
    check(wgk_set_forcing_unf(ctx, 0, 365, -1, raw[0].data(), raw[1].data(), raw[2].data(), raw[3].data(), 365), "wgk_set_forcing_unf");
    check(wgk_synchronize(ctx), "sync");  // the host buffers go out of scope
}

// ------------------------------------------------------------------------------------------
// integrate_wghm-shaped driver
// ------------------------------------------------------------------------------------------
// initialize_wghm (initializeWGHM.cpp:32-72) + everything integrate_wghm_ does before its year loop (integrateWGHM.cpp:127-476),
// on the host grids: a cold start, or a restart from the three checkpoint files of the config (PDAF monthly cycle)
void load_start_state(const ConfigFile &cfg, int ncell, ModelState &S, calibParamClass &calParam) {  // initializeWGHM.cpp:32-72
    if (!cfg.startvaluefile.empty()) S.wghmState.load(cfg.startvaluefile);
    calParam.readJson(cfg.parameterfile, ncell);
    if (!cfg.additionalfile.empty()) S.additionalOutIn.load(cfg.additionalfile);  // sets additionalfilestatus = 1
    if (!cfg.snowInElevationfile.empty()) S.snow_in_elevation.load(cfg.snowInElevationfile);
}

void initialize_model(Engine &E, const ConfigFile &cfg, ModelStateRef S) {
    const int ncell = E.ncell;
    E.options.init(cfg);
    E.options.require_canonical();
    AdditionalOutputInputFile &additionalOutIn = S.additionalOutIn;
    // init sequence, integrateWGHM.cpp:127-286
    if (1 == E.options.rout_prepare) E.topo = prepare_routing_files(E.options.input_dir, E.options.routing_dir, 1, E.options.resOpt, ncell);
    E.geo.init(E.options.input_dir, ncell, E.options.resOpt);
    E.dailyWaterBalance.init(E.options.input_dir, 0);
    E.land_init();
    E.G_aindex.initialize(ncell);
    E.G_aindex.read(E.options.input_dir + "/G_ARID_HUMID.UNF2");
    E.dailyWaterBalance.G_Elevation.read(E.options.input_dir + "/G_ELEV_RANGE.101.UNF2");
    E.lai_init(additionalOutIn);
    E.routing.init(0, cfg, S.wghmState, additionalOutIn);
    E.routing.initLakeDepthActive(E.calParam);
    E.routing.initWetlDepthActive(E.calParam);
    // body of the single-pass calibration loop, :291-476
    if (additionalOutIn.additionalfilestatus == 0) E.routing.initFractionStatus();
    if (additionalOutIn.additionalfilestatus == 1) E.routing.initFractionStatusAdditionalOI(additionalOutIn);
    if (cfg.additionalfile.empty()) E.routing.setStoragesToZero();
    else E.routing.setStorages(S.wghmState, additionalOutIn);
    if (cfg.startvaluefile.empty()) E.routing.setLakeWetlToMaximum(E.options.start_year);
    if (additionalOutIn.additionalfilestatus == 1) {  // :311-316, first day after a checkpoint
        E.routing.annualInit(E.options.start_year, cfg.startMonth, additionalOutIn);
        E.routing.update_landarea_red_fac_PDAF(E.calParam, additionalOutIn);
    }
    E.G_toBeCalculated.initialize(ncell);
    E.G_toBeCalculated.fill(1);
    for (int n = 0; n < ncell; n++) {
        E.dailyWaterBalance.G_gammaHBV[n] = E.calParam.getValue(P_GAMRUN_C, n);
        E.dailyWaterBalance.G_cellCorrFact[n] = E.calParam.getValue(P_CFA, n);
        E.routing.G_statCorrFact[n] = E.calParam.getValue(P_CFS, n);
    }
    if (cfg.additionalfile.empty()) E.dailyWaterBalance.setStoragesToZero();
    else E.dailyWaterBalance.setStorages(S.wghmState, S.snow_in_elevation, additionalOutIn);
    E.createMaxSoilWaterCapacityGrid();
    E.createGroundwaterGrids();
}

// the columns of the additionalOutIn checkpoint that the hot path owns, as the reference leaves them after the last day of a
// month (lai.cpp:168-173, daily.cpp:1258-1262, routing.cpp:3080, 5050-5071, 5165-5170, 5207-5241, 5347-5350); the water-use
// columns keep what was loaded (zeros on a cold start)
static void fill_additional(Engine &E, AdditionalOutputInputFile &add, short month) {
    const int ncell = E.ncell;
    routingClass &r = E.routing;
    dailyWaterBalanceClass &d = E.dailyWaterBalance;
    Grid<int32_t> ld, ls;
    Grid<double> lp, fll, flw, fgw;
    E.get("lai_days", ld); E.get("lai_status", ls); E.get("lai_precsum", lp);
    E.get("f_loc_lake", fll); E.get("f_loc_wet", flw); E.get("f_glo_wet", fgw);
    for (int n = 0; n < ncell; n++) {
        auto col = [&](int j) -> double & { return add.additionalOutputInput(n, j); };
        if (E.geo.G_contcell[n]) {  // the per-cell writes of lai / calcNewDay only happen for computed cells
            col(0) = ld[n]; col(1) = ls[n]; col(2) = lp[n];
            col(20) = d.G_soilWaterContent[n]; col(22) = d.G_canopyWaterContent[n]; col(24) = d.G_snow[n];
        }
        if (r.G_reservoir_area[n] > 0.) col(5) = r.K_release[n];
        col(6) = r.G_landAreaFrac[n]; col(7) = r.G_landAreaFracPrevTimestep[n];
        col(8) = r.G_locWetlAreaReductionFactor[n]; col(9) = r.G_gloLakeEvapoReductionFactor[n]; col(10) = r.G_groundwaterStorage[n];
        col(11) = r.G_locLakeAreaReductionFactor[n]; col(12) = r.G_gloWetlAreaReductionFactor[n]; col(13) = r.G_gloResEvapoReductionFactor[n];
        col(14) = r.G_fswbInit[n]; col(15) = r.G_gloWetlStorage[n]; col(17) = r.G_locWetlStorage[n]; col(18) = r.G_locLakeStorage[n];
        col(19) = r.G_riverStorage[n]; col(21) = r.G_gloLakeStorage[n]; col(23) = r.G_gloResStorage[n];
        col(33) = r.G_fswbLandAreaFracNextTimestep[n]; col(36) = r.G_fGloLake[n];
        col(44) = (month == 11) ? r.G_glo_res[n] : r.G_glores_prevyear[n];
        col(45) = fll[n]; col(47) = flw[n]; col(46) = fgw[n];
        if (E.options.subtract_use > 0) {  // :5025-5029, 5231-5240 (use_alloc 0: the allocation columns stay 0)
            col(3) = r.G_totalUnsatisfiedUse[n]; col(4) = r.G_UnsatisfiedUsePrevYear[n]; col(50) = r.G_reducedReturnFlowPrevYear[n];
            col(51) = r.G_unsatisfiedNAsFromIrrigPrevYear[n]; col(52) = r.G_unsatisfiedNAsFromOtherSectorsPrevYear[n];
            col(35) = r.G_dailyRemainingUse[n]; col(38) = r.G_reducedReturnFlow[n]; col(39) = r.G_unsatisfiedNAsFromIrrig[n];
            col(49) = r.G_unsatisfiedNAsFromOtherSectors[n]; col(25) = r.G_withdrawalIrrigFromSwb[n]; col(26) = r.G_consumptiveUseIrrigFromSwb[n];
        }
    }
}

long integrate_wghm(const std::string &config_file, int ncell, int device, double *seconds_day_loop) {
    ConfigFile cfg(config_file);
    ModelState S(ncell);
    calibParamClass cal;
    load_start_state(cfg, ncell, S, cal);
    optionClass opt;
    opt.init(cfg);
    Engine E(ncell, device, S.additionalOutIn.additionalfilestatus, opt.subtract_use);
    E.calParam = cal;
    return run_model(E, cfg, S, seconds_day_loop);
}

// initialisation + the year / month / day loop of integrate_wghm_ (integrateWGHM.cpp:127-917) on the state objects of S
long run_model(Engine &E, const ConfigFile &cfg, ModelStateRef S, double *seconds_day_loop) {
    const int ncell = E.ncell;
    initialize_model(E, cfg, S);
    WghmStateFile &wghmState = S.wghmState;
    AdditionalOutputInputFile &additionalOutIn = S.additionalOutIn;
    SnowInElevationFile &snow_in_elevation = S.snow_in_elevation;
    const short number_of_days_in_month[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    const short last_day_in_month[12] = {30, 58, 89, 119, 150, 180, 211, 242, 272, 303, 333, 364};
    short readinstatus = 1;
    const bool yearly_forcing = (E.options.time_series == 1);
    E.check(wgk_forcing_reserve(E.ctx, yearly_forcing ? 365 : 31, 0), "wgk_forcing_reserve");

    long ndays = 0;
    double secs = 0.;
    bool pushed = false;
    for (short year = E.options.start_year; year <= E.options.end_year; year++) {
        E.dailyWaterBalance.annualInit();
        E.routing.annualInit(year, cfg.startMonth, additionalOutIn);
        if (!pushed) {
            E.push_static();
            E.push_state();
            if (E.options.subtract_use > 0) E.routing.pushWaterUseState();
            pushed = true;
        } else if (E.routing.yearlyChanged) {  // resYearOpt 1: reservoirs that start operating this year
            E.routing.pushYearly();
        }
        if (E.options.subtract_use > 0) E.routing.dailyNUInit(E.options.water_use_dir, year, E.calParam);  // integrateWGHM.cpp:645-647
        if (yearly_forcing) E.set_forcing_year(year);  // integrateWGHM.cpp:565-568
        short day = 0;
        if (year == cfg.startYear) for (int m = 1; m < cfg.startMonth; m++) day += number_of_days_in_month[m - 1];
        short start_month = cfg.startMonth, end_month = cfg.endMonth;
        if (E.options.start_year != E.options.end_year) {
            if (year == E.options.start_year) end_month = 12;
            else if (year == E.options.end_year) start_month = 1;
            else { start_month = 1; end_month = 12; }
        }
        for (short month = start_month - 1; month < end_month; month++) {
            if (!yearly_forcing) E.set_forcing_month(month + 1, year);
            if (E.options.subtract_use > 0) E.routing.pushWaterUseMonth(month);  // calcNextDay_M, integrateWGHM.cpp:794-796
            wghmState.resetCells(number_of_days_in_month[month]);
            E.check(wgk_month_begin(E.ctx), "wgk_month_begin");  // the post-pass keeps what the month's checkpoint needs
            const auto t0 = std::chrono::steady_clock::now();
            for (short dom = 1; dom <= number_of_days_in_month[month]; dom++) {
                day++;
                ndays++;
                // the reference's per-cell loop (integrateWGHM.cpp:770-783) through the shim
                for (int n = 0; n < ncell; n++)
                    if (E.geo.G_contcell[n])
                        E.dailyWaterBalance.calcNewDay(day, month, dom, last_day_in_month[month], year, n, wghmState, additionalOutIn,
                                                       snow_in_elevation, readinstatus, E.calParam);
                E.routing.routing(year, day, month, dom, last_day_in_month[month], wghmState, additionalOutIn, readinstatus, E.calParam);
                E.routing.updateLandAreaFrac(additionalOutIn);
                if (readinstatus == 1) readinstatus = 0;
            }
            secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            // month-end rescale, integrateWGHM.cpp:830-847
            E.dailyWaterBalance.pull();
            E.routing.pull();
            for (int n = 0; n < ncell; n++) {
                const double laf = E.routing.getLandAreaFrac(n), cf = E.geo.G_contfreq[n];
                for (short e = 0; e <= 100; e++)
                    snow_in_elevation.snowInElevation(n, e) = (laf == 0.) ? 0. : E.dailyWaterBalance.G_SnowInElevation(n, e) * laf / cf;
                for (short dom = 1; dom <= number_of_days_in_month[month]; dom++) {
                    wghmState.cell(n).canopy(dom - 1) = E.dailyWaterBalance.G_canopyWaterContent[n] * laf / cf;
                    wghmState.cell(n).snow(dom - 1) = E.dailyWaterBalance.G_snow[n] * laf / cf;
                    wghmState.cell(n).soil(dom - 1) = E.dailyWaterBalance.G_soilWaterContent[n] * laf / cf;
                }
            }
            if (E.options.subtract_use > 0) E.routing.pullWaterUse();
            fill_additional(E, additionalOutIn, month);
            if (E.options.subtract_use > 0 && month == 11) {  // integrateWGHM.cpp:944: the year's use bookkeeping
                E.routing.annualWaterUsePostProcessing(year, additionalOutIn);
                E.routing.pushWaterUseState();
            }
            if (year == E.options.end_year && month + 1 == end_month) {  // :853-901
                if (!cfg.outputmeanfile.empty()) wghmState.saveMean(cfg.outputmeanfile);
                if (!cfg.outputlastdayfile.empty()) wghmState.saveDay(cfg.outputlastdayfile, number_of_days_in_month[month] - 1);
                if (!cfg.outputadditionalfile.empty()) additionalOutIn.save(cfg.outputadditionalfile);
                if (!cfg.outputsnowlastdayfile.empty()) snow_in_elevation.save(cfg.outputsnowlastdayfile);
            }
        }
        if (end_month == 12) E.routing.updateGloResPrevYear_pct();  // integrateWGHM.cpp:921-922
    }
    if (seconds_day_loop) *seconds_day_loop = secs;
    return ndays;
}

// host grids after initialize_model (+ the first annualInit of the year loop) as a record dump; no context is created
static void dump_put(FILE *f, const char *name, const char *dtype, int64_t count, const void *data, size_t elsize) {
    char nm[32] = {0}, dt[8] = {0};
    strncpy(nm, name, 31);
    strncpy(dt, dtype, 7);
    const int32_t d = 0;
    fwrite(nm, 1, 32, f); fwrite(&d, 4, 1, f); fwrite(dt, 1, 8, f); fwrite(&count, 8, 1, f); fwrite(data, elsize, (size_t)count, f);
}
void init_dump(const std::string &config_file, int ncell, const std::string &dump_file) {
    ConfigFile cfg(config_file);
    Engine E(ncell, -1);
    ModelState S(ncell);
    load_start_state(cfg, ncell, S, E.calParam);
    initialize_model(E, cfg, S);
    E.dailyWaterBalance.annualInit();
    E.routing.annualInit(E.options.start_year, cfg.startMonth, S.additionalOutIn);
    FILE *f = fopen(dump_file.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + dump_file);
    auto f64 = [&](const char *name, Grid<double> &g) { dump_put(f, name, "f64", ncell, g.data(), 8); };
    routingClass &r = E.routing;
    dailyWaterBalanceClass &d = E.dailyWaterBalance;
    f64("canopy", d.G_canopyWaterContent); f64("soil", d.G_soilWaterContent); f64("snow", d.G_snow);
    dump_put(f, "snow_bands", "f64", (int64_t)ncell * 101, d.G_SnowInElevation.data(), 8);
    dump_put(f, "lai_days", "i32", ncell, E.lai_days.data(), 4); dump_put(f, "lai_status", "i32", ncell, E.lai_status.data(), 4);
    f64("lai_precsum", E.lai_precsum);
    f64("gw", r.G_groundwaterStorage); f64("loc_lake_stor", r.G_locLakeStorage); f64("loc_wetl_stor", r.G_locWetlStorage);
    f64("glo_lake_stor", r.G_gloLakeStorage); f64("glo_wetl_stor", r.G_gloWetlStorage); f64("res_stor", r.G_gloResStorage);
    f64("river_stor", r.G_riverStorage); f64("red_loc_lake", r.G_locLakeAreaReductionFactor); f64("red_loc_wetl", r.G_locWetlAreaReductionFactor);
    f64("red_glo_lake", r.G_gloLakeEvapoReductionFactor); f64("red_glo_wetl", r.G_gloWetlAreaReductionFactor); f64("red_res", r.G_gloResEvapoReductionFactor);
    f64("red_river", r.G_riverAreaReductionFactor); f64("k_release", r.K_release); f64("land_area_frac", r.G_landAreaFrac);
    f64("land_area_frac_prev", r.G_landAreaFracPrevTimestep); f64("land_area_frac_next", r.G_landAreaFracNextTimestep);
    f64("fswb_laf", r.G_fswbLandAreaFrac); f64("fswb_laf_next", r.G_fswbLandAreaFracNextTimestep); f64("fswb_init", r.G_fswbInit);
    f64("river_area_frac_next", r.G_riverAreaFracNextTimestep_Frac); f64("f_glo_lake", r.G_fGloLake);
    dump_put(f, "status_laf_next", "i16", ncell, r.statusStarted_landAreaFracNextTimestep.data(), 2);
    f64("reservoir_area", r.G_reservoir_area); f64("stor_cap", r.G_stor_cap); f64("lake_area", r.G_lake_area);
    dump_put(f, "smax", "f32", ncell, E.G_Smax.data(), 4); dump_put(f, "gwfactor", "f32", ncell, E.G_gwFactor.data(), 4);
    fclose(f);
}

}  // namespace wg

// ------------------------------------------------------------------------------------------
// B1: the Fortran / C entry points of the reference with their signatures and ownership rules (initializeWGHM.h:11-15,
// integrateWGHM.h:9-13).  The callee news the objects into the caller's pointer references; integrate_wghm_ deletes them at the
// configured end date in "OL" mode (integrateWGHM.cpp:1139-1160); errors leave as C++ exceptions, as in the reference.
// The cell count is a run-time value here: it is the ng_param header of the parameter file.  The GPU is $WGK_DEVICE (default 0).
// ------------------------------------------------------------------------------------------
extern "C" {
void initialize_wghm_(const char *s, wg::WghmStateFile *&initstate, wg::calibParamClass *&initcal, wg::AdditionalOutputInputFile *&initaddio,
                      wg::SnowInElevationFile *&initsnow, long *year, long *month, const char *s2, const char *s3, wg::WghmStateFile *&wghmMean) {
    using namespace wg;
    const std::string progName(s2 ? s2 : "OL"), pathMean(s3 ? s3 : "");
    ConfigFile cfg(s, (int)*year, (int)*month, progName);
    initcal = new calibParamClass;
    if (!cfg.parameterfile.empty()) initcal->readJson(cfg.parameterfile, 0);
    const int ncell = initcal->ncell();
    if (ncell <= 0) throw std::runtime_error("initialize_wghm_: the parameter file (param_json) must give the number of cells (ng_param)");
    initstate = new WghmStateFile(ncell, 1);
    if (!cfg.startvaluefile.empty()) initstate->load(cfg.startvaluefile);
    wghmMean = new WghmStateFile(ncell, 1);
    if (!pathMean.empty()) wghmMean->load(pathMean);
    initaddio = new AdditionalOutputInputFile(ncell);
    if (!cfg.additionalfile.empty()) initaddio->load(cfg.additionalfile);  // additionalfilestatus = 1
    initsnow = new SnowInElevationFile(ncell);
    if (!cfg.snowInElevationfile.empty()) initsnow->load(cfg.snowInElevationfile);
}

void integrate_wghm_(const char *s, wg::ConfigFile *&configFile, wg::WghmStateFile *&wghmState, wg::calibParamClass *&calParam,
                     wg::AdditionalOutputInputFile *&additionalOutIn, wg::SnowInElevationFile *&snow_in_elevation, long *step, long *total_steps,
                     long *year, long *month, const char *s2) {
    using namespace wg;
    (void)step; (void)total_steps;
    const std::string progName(s2 ? s2 : "OL");
    configFile = new ConfigFile(s, (int)*year, (int)*month, progName);
    const int ncell = calParam->ncell();
    const char *dev = getenv("WGK_DEVICE");
    {
        optionClass opt;
        opt.init(*configFile);
        Engine E(ncell, dev ? atoi(dev) : 0, additionalOutIn->additionalfilestatus, opt.subtract_use);
        E.calParam = *calParam;
        ModelStateRef S{*wghmState, *additionalOutIn, *snow_in_elevation};
        run_model(E, *configFile, S, nullptr);
    }
    if (progName == "OL") {  // the run is over: the reference frees what initialize_wghm_ allocated
        delete wghmState; wghmState = nullptr;
        delete calParam; calParam = nullptr;
        delete additionalOutIn; additionalOutIn = nullptr;
        delete snow_in_elevation; snow_in_elevation = nullptr;
        delete configFile; configFile = nullptr;
    }
}
// C++ aliases of the reference (initializeWGHM.h:19, integrateWGHM.h:17)
}
void initialize_wghm(const char *s, wg::WghmStateFile *&a, wg::calibParamClass *&b, wg::AdditionalOutputInputFile *&c, wg::SnowInElevationFile *&d,
                     long *year, long *month, const char *s2, const char *s3, wg::WghmStateFile *&m) {
    initialize_wghm_(s, a, b, c, d, year, month, s2, s3, m);
}
void integrate_wghm(const char *s, wg::ConfigFile *&cf, wg::WghmStateFile *&a, wg::calibParamClass *&b, wg::AdditionalOutputInputFile *&c,
                    wg::SnowInElevationFile *&d, long *step, long *total_steps, long *year, long *month, const char *s2) {
    integrate_wghm_(s, cf, a, b, c, d, step, total_steps, year, month, s2);
}

extern "C" {
long wg_host_integrate(const char *config_file, int ncell, int device, double *seconds_day_loop, char *err, size_t errlen) {
    try {
        return wg::integrate_wghm(config_file, ncell, device, seconds_day_loop);
    } catch (std::exception &e) {
        if (err && errlen) snprintf(err, errlen, "%s", e.what());
        return -1;
    }
}
// load + re-save of the three checkpoint files (format round trip, used by the tests)
int wg_host_state_roundtrip(const char *kind, const char *in, const char *out, int ncell, char *err, size_t errlen) {
    try {
        const std::string k(kind);
        if (k == "state") { wg::WghmStateFile f(ncell, 1); f.load(in); f.saveDay(out, 0); }
        else if (k == "snow") { wg::SnowInElevationFile f(ncell); f.load(in); f.save(out); }
        else if (k == "additional") { wg::AdditionalOutputInputFile f(ncell); f.load(in); f.save(out); }
        else throw std::runtime_error("unknown kind " + k);
        return 0;
    } catch (std::exception &e) {
        if (err && errlen) snprintf(err, errlen, "%s", e.what());
        return -1;
    }
}
// a ready-to-step context from a reference-format configuration: initialize_wghm + the init sequence of integrate_wghm_ + the
// first annualInit on the host, statics / parameters / start state pushed to the device, 31 forcing slots reserved.  The caller
// owns the returned wgk_ctx (wgk_destroy); the benchmark and smoke() build their model this way, through the product's own
// host layer (topology builder included) instead of test infrastructure.
void *wg_host_create_context(const char *config_file, int ncell, int device, char *err, size_t errlen) {
    try {
        wg::ConfigFile cfg(config_file);
        wg::ModelState S(ncell);
        wg::calibParamClass cal;
        wg::load_start_state(cfg, ncell, S, cal);
        wg::optionClass opt;
        opt.init(cfg);
        wg::Engine E(ncell, device, S.additionalOutIn.additionalfilestatus, opt.subtract_use);
        E.calParam = cal;
        wg::initialize_model(E, cfg, S);
        E.dailyWaterBalance.annualInit();
        E.routing.annualInit(E.options.start_year, cfg.startMonth, S.additionalOutIn);
        E.push_static();
        E.push_state();
        if (E.options.subtract_use > 0) E.routing.pushWaterUseState();
        E.check(wgk_forcing_reserve(E.ctx, 31, 0), "wgk_forcing_reserve");
        wgk_ctx *ctx = E.ctx;
        E.ctx = nullptr;  // ownership moves to the caller
        return ctx;
    } catch (std::exception &e) {
        if (err && errlen) snprintf(err, errlen, "%s", e.what());
        return nullptr;
    }
}
int wg_host_init_dump(const char *config_file, int ncell, const char *dump_file, char *err, size_t errlen) {
    try {
        wg::init_dump(config_file, ncell, dump_file);
        return 0;
    } catch (std::exception &e) {
        if (err && errlen) snprintf(err, errlen, "%s", e.what());
        return -1;
    }
}
int wg_host_prepare_routing_files(const char *input_dir, const char *routing_dir, int resOpt, int ncell, int *nlevels, char *err, size_t errlen) {
    try {
        wg::FlowTopology t = wg::prepare_routing_files(input_dir, routing_dir, 1, (short)resOpt, ncell);
        if (nlevels) *nlevels = t.nlevels;
        return 0;
    } catch (std::exception &e) {
        if (err && errlen) snprintf(err, errlen, "%s", e.what());
        return -1;
    }
}
}
