/* wg_oracle.c — CPU restatement of the WaterGAP2 daily hot path (TEST INFRASTRUCTURE; see
 * wg_oracle.h).  Scalar, reference-ordered, array-of-cells in the reference's own cell
 * numbering.  Every block cites the reference lines it restates.  Expression shapes
 * (association, operand order, float-vs-double typing) follow the reference so that a build
 * with `gcc -O2 -ffp-contract=off` is bit-identical to oracle/_ref on the same inputs.
 *
 * Canonical options: time_series 0, cloud 1, intercept 1, calc_albedo 0,
 * petOpt 0, use_kc 1, clclOpt 0, riverveloOpt 1, subtract_use 0, resOpt 1, statcorrOpt 0,
 * aridareaOpt 1, fractionalRoutingOpt 1, riverEvapoOpt 1, antNatOpt 0, glacierOpt 0,
 * calc_wtemp 0, timeStepsPerDay 1.
 */
#include "wg_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct wgo_ctx {
    int ncell;
    int restart; /* additionalOutIn.additionalfilestatus */
    int subtract_use; /* options.subtract_use */
#define X(name, ctype, dt, per) ctype *name;
    WGO_FIELDS(X)
#undef X
};

static const double MIN_STOR_VOL = 1.e-15; /* routing.h:24 */

wgo_ctx *wgo_create(int ncell) {
    wgo_ctx *c = (wgo_ctx *)calloc(1, sizeof(wgo_ctx));
    c->ncell = ncell;
#define X(name, ctype, dt, per) \
    c->name = (ctype *)calloc((size_t)((per) == 0 ? WGO_NLCT : (size_t)(per) * (size_t)ncell), sizeof(ctype));
    WGO_FIELDS(X)
#undef X
    return c;
}

void wgo_destroy(wgo_ctx *c) {
    if (!c) return;
#define X(name, ctype, dt, per) free(c->name);
    WGO_FIELDS(X)
#undef X
    free(c);
}

int wgo_ncell(const wgo_ctx *c) { return c->ncell; }
void wgo_set_restart(wgo_ctx *c, int restart) { c->restart = restart; }
void wgo_set_subtract_use(wgo_ctx *c, int subtract_use) { c->subtract_use = subtract_use; }

void *wgo_field(wgo_ctx *c, const char *name, const char **dtype, int64_t *count) {
#define X(fname, ctype, dt, per) \
    if (strcmp(name, #fname) == 0) { \
        if (dtype) *dtype = dt; \
        if (count) *count = (per) == 0 ? WGO_NLCT : (int64_t)(per) * c->ncell; \
        return c->fname; \
    }
    WGO_FIELDS(X)
#undef X
    return NULL;
}

#define PAR(c, k, n) ((c)->params[(size_t)(k) * (size_t)(c)->ncell + (size_t)(n)])

/* ------------------------------------------------------------------------------------------
 * LAI growing-season state machine, lai.cpp:152-306
 * ---------------------------------------------------------------------------------------- */
static double lai_growing(int *days, short initialDays, int *status, int lct, int arid, double LAImin, double LAImax,
                          double *precsum, double prec) { /* lai.cpp:179-241 */
    if (*status == 0) {
        if (*days >= initialDays) {
            (*days)++;
            (*precsum) += prec;
            if (*precsum > 40.) {
                if (*days >= initialDays + 30) {
                    *days = initialDays + 30;
                    *status = 1;
                }
                return (LAImin + (LAImax - LAImin) * (*days - initialDays) / 30.);
            } else {
                *days = initialDays;
                return LAImin;
            }
        } else {
            (*days)++;
            (*precsum) += prec;
            return LAImin;
        }
    } else {
        if (*days <= 30) {
            (*days)--;
            if (lct <= 2) *status = 0;
            if (*days <= 0) {
                *days = 0;
                *status = 0;
                *precsum = 0.;
            }
            return (LAImax - (LAImax - LAImin) * (30 - *days) / 30.);
        } else {
            if (arid && (prec < 0.5)) (*days)--;
            else *days = 30 + initialDays;
            return LAImax;
        }
    }
}

static double lai_nogrowing(int *days, short initialDays, int *status, double LAImin, double LAImax, double *precsum,
                            double prec) { /* lai.cpp:243-293 */
    if (*status == 0) {
        if (*days > initialDays) {
            (*days)++;
            (*precsum) += prec;
            if (*precsum > 40.) {
                if (*days >= initialDays + 30) {
                    *days = initialDays + 30;
                    *status = 1;
                }
                return (LAImin + (LAImax - LAImin) * (*days - initialDays) / 30.);
            } else {
                *days = initialDays;
                return LAImin;
            }
        } else {
            (*precsum) += prec;
            return LAImin;
        }
    } else {
        if (*days <= 30) {
            (*days)--;
            if (*days <= 0) {
                *days = 0;
                *status = 0;
                *precsum = 0.;
            }
            return (LAImax - (LAImax - LAImin) * (30 - *days) / 30.);
        } else {
            (*days)--;
            return LAImax;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * vertical water balance of one cell and one day: dailyWaterBalanceClass::calcNewDay,
 * daily.cpp:94-1264 (output bookkeeping :1281-1539 is out of scope)
 * ---------------------------------------------------------------------------------------- */
static void vertical_cell(wgo_ctx *c, int n, int day_in_month) {
    /* daily.cpp:159-169 (routing.h:246-251 getLandAreaFrac) */
    double landAreaFrac = (0 == c->status_laf_next[n]) ? c->land_area_frac[n] : c->land_area_frac_next[n];
    double lafPrev;
    if (1 == c->status_laf_next[n]) lafPrev = c->land_area_frac_prev[n];
    else if (c->restart == 1) lafPrev = c->land_area_frac_prev[n];
    else lafPrev = landAreaFrac;

    double soil_saturation = 0., soil_water_overflow = 0., neg_land_aet = 0., neg_runoff = 0.;
    if (1 != c->toBeCalculated[n]) return; /* daily.cpp:177 */

    c->land_aet[n] = 0.;
    c->land_aet_uncorr[n] = 0.;
    const int lc = c->landcover[n] - 1;

    /* forcing, daily.cpp:194-205 (float grids promoted to double) */
    double dailyPrec = c->prec31[(size_t)n * 31 + (day_in_month - 1)];
    double dailyTempC = c->temp31[(size_t)n * 31 + (day_in_month - 1)];
    double dailyShortWave = c->sw31[(size_t)n * 31 + (day_in_month - 1)];
    double dailyLongWave = c->lw31[(size_t)n * 31 + (day_in_month - 1)];

    dailyPrec = PAR(c, WGO_M_PREC, n) * dailyPrec; /* :248 */
    double temp2 = dailyTempC + 237.3;
    double e_s = 0.6108 * exp(17.27 * dailyTempC / temp2); /* :250-251 */

    /* arid / humid, daily.cpp:331-348 */
    double alpha, maxDailyPET;
    int arid_gw = 0;
    if (c->arid[n] == 1) {
        alpha = PAR(c, WGO_P_PTC_ARI, n);
        maxDailyPET = PAR(c, WGO_P_PET_MXDY, n);
        arid_gw = 1;
    } else if (c->arid[n] == 0) {
        alpha = PAR(c, WGO_P_PTC_HUM, n);
        maxDailyPET = PAR(c, WGO_P_PET_MXDY, n);
    } else {
        fprintf(stderr, "Error: Invalid value for Arid/humid index: %d\n", (int)c->arid[n]);
        exit(1);
    }

    /* LAI and Kc, daily.cpp:355-356; lai.cpp:152-177, 295-306 (LAImin in float arithmetic) */
    float LAImin_f = c->lai_factor_a[lc] + c->lai_factor_b[lc] * c->laimax[n];
    double LAImin = LAImin_f;
    double LAImaxd = c->laimax[n];
    int days = c->lai_days[n], status = c->lai_status[n];
    double precsum = c->lai_precsum[n];
    double dailyLai;
    if (dailyTempC > 8.)
        dailyLai = lai_growing(&days, c->lai_initial_days[lc], &status, lc + 1, arid_gw, LAImin, LAImaxd, &precsum, dailyPrec);
    else
        dailyLai = lai_nogrowing(&days, c->lai_initial_days[lc], &status, LAImin, LAImaxd, &precsum, dailyPrec);
    c->lai_days[n] = days;
    c->lai_status[n] = status;
    c->lai_precsum[n] = precsum;
    double dailyKc;
    if ((c->laimax[n] - LAImin) == 0.) dailyKc = c->lai_kc_min[lc];
    else dailyKc = c->lai_kc_min[lc] + (c->lai_kc_max[lc] - c->lai_kc_min[lc]) * (dailyLai - LAImin) / (c->laimax[n] - LAImin);

    /* albedo, daily.cpp:366-378 (G_snow = value of the previous day) */
    double albedo;
    if (c->snow[n] > 3.) albedo = c->lct_albedo_snow[lc];
    else albedo = 0.23;

    /* latent heat, :381-386 */
    double lat_heat;
    if (dailyTempC > 0) lat_heat = 2.501 - 0.002361 * dailyTempC;
    else lat_heat = 2.835;

    /* radiation, :404-462 (cloud == 1) */
    double conv_Wm2_to_mmd = 0.0864 / lat_heat;
    double solar_rad = conv_Wm2_to_mmd * dailyShortWave;
    double long_wave_rad_in = conv_Wm2_to_mmd * dailyLongWave;
    double emissivity = c->lct_emissivity[lc];
    double temp_K = dailyTempC + 273.2;
    const double stefan_boltz_const = 0.000000004903;
    double long_wave_rad_out = emissivity * stefan_boltz_const * pow(temp_K, 4.) / lat_heat;
    double net_long_wave_rad = long_wave_rad_in - long_wave_rad_out;
    double net_short_wave_rad = solar_rad * (1. - albedo);
    double net_rad = PAR(c, WGO_M_NETRAD, n) * (net_short_wave_rad + net_long_wave_rad);
    const double openWaterAlbedo = 0.08;
    double openWaterNetShortWaveRad = solar_rad * (1. - openWaterAlbedo);
    double openWaterNetRad = openWaterNetShortWaveRad + net_long_wave_rad;

    /* Priestley-Taylor, :476-522 */
    double dailyPET = 0., dailyOpenWaterPET = 0.;
    double atmos_pres = 101.3;
    double inc_svp = 4098. * e_s / (temp2 * temp2);
    double c3 = 0.0016286 * atmos_pres;
    double gamma = c3 / lat_heat;
    if (net_rad <= 0.) dailyPET = 0.;
    else dailyPET = alpha * (inc_svp * net_rad) / (inc_svp + gamma);
    if (openWaterNetRad <= 0.) dailyOpenWaterPET = 0.;
    else dailyOpenWaterPET = alpha * (inc_svp * openWaterNetRad) / (inc_svp + gamma);

    /* crop coefficients, :768-771 */
    const double kc_OpenWater = 1.05;
    if ((c->snow[n] <= 3.)) {
        dailyPET *= dailyKc;
        dailyOpenWaterPET *= kc_OpenWater;
    }
    const double cfa = c->cfa[n];
    c->lake_balance[n] = (dailyPrec - dailyOpenWaterPET) * cfa; /* :773 */
    c->openwater_prec[n] = dailyPrec;
    c->openwater_pet[n] = dailyOpenWaterPET;

    double landStorageChangeSum = 0., initialStorage = 0.;
    double max_canopy_storage = 0., dailyCanopyEvapo = 0., daily_prec_to_soil = 0., dailySoilPET = 0.;
    double canopy_water_content = 0.;
    double dailySnowEvapo = 0., dailyEffPrec = 0.;
    double immediate_runoff = 0., surface_runoff = 0., dailyAET = 0., daily_runoff = 0., total_daily_runoff = 0.;
    double daily_gw_recharge = 0., pot_gw_recharge = 0.;

    const double P_T_SNOWFZ = PAR(c, WGO_P_T_SNOWFZ, n);
    const double P_T_SNOWMT = PAR(c, WGO_P_T_SNOWMT, n);
    const double P_T_GRADNT = PAR(c, WGO_P_T_GRADNT, n);
    const double M_DEGDAY_F = PAR(c, WGO_M_DEGDAY_F, n);
    const double canopyEvapoExp = 0.66666666; /* daily.cpp:1867 */
    const double runoffFracBuiltUp = 0.5;

    /* interception, :825-894 */
    {
        double canopy_deficiency;
        if (landAreaFrac <= 0.) {
            c->storage_transfer[n] = c->canopy[n];
            c->canopy[n] = 0.;
            dailyCanopyEvapo = 0.;
        } else {
            c->canopy[n] *= lafPrev / landAreaFrac;
            if (fabs(c->canopy[n]) <= MIN_STOR_VOL) c->canopy[n] = 0.;
            initialStorage = c->canopy[n];
            if (dailyLai > 0.00001) {
                max_canopy_storage = PAR(c, WGO_P_MCWH, n) * dailyLai;
                canopy_deficiency = max_canopy_storage - c->canopy[n];
                if (dailyPrec < canopy_deficiency) {
                    c->canopy[n] += dailyPrec;
                    daily_prec_to_soil = 0.;
                } else {
                    c->canopy[n] = max_canopy_storage;
                    daily_prec_to_soil = dailyPrec - canopy_deficiency;
                }
                canopy_water_content = c->canopy[n];
                dailyCanopyEvapo = dailyPET * pow((canopy_water_content / max_canopy_storage), canopyEvapoExp);
                if (dailyCanopyEvapo > canopy_water_content) {
                    dailyCanopyEvapo = canopy_water_content;
                    dailySoilPET = dailyPET - canopy_water_content;
                    c->canopy[n] = 0.0;
                } else {
                    c->canopy[n] -= dailyCanopyEvapo;
                    dailySoilPET = dailyPET - dailyCanopyEvapo;
                }
            } else {
                daily_prec_to_soil = dailyPrec;
                dailySoilPET = dailyPET;
                dailyCanopyEvapo = 0.0;
            }
            landStorageChangeSum += c->canopy[n] - initialStorage;
        }
    }
    if (dailySoilPET < 0.) dailySoilPET = 0.0;

    /* snow in 100 elevation bands, :913-1062 */
    double TempElevMax = 0., snowStorageChange = 0.;
    c->thresh_elev[n] = 0;
    c->snow[n] = 0.;
    const int16_t *elev = c->elevation + (size_t)n * WGO_NBAND;
    double *S = c->snow_bands + (size_t)n * WGO_NBAND;
    for (short e = 1; e < 101; e++) {
        double temp_elev, daily_snow_to_soil_elev = 0., dailyEffPrec_elev, snowmelt_elev = 0.;
        double dailyEffPrecBeforeSnowMelt_elev = 0.;
        temp_elev = dailyTempC - ((elev[e] - elev[0]) * P_T_GRADNT);
        if (landAreaFrac <= 0.) {
            c->storage_transfer[n] += S[e] / 100.;
            S[e] = 0.;
            c->snow[n] = 0.;
        } else {
            S[e] = S[e] * lafPrev / landAreaFrac;
            if (fabs(S[e]) <= MIN_STOR_VOL) S[e] = 0.;
            initialStorage = S[e];
            if (S[e] > 1000.) {
                if (c->thresh_elev[n] == 0.) c->thresh_elev[n] = elev[e];
                else if (c->thresh_elev[n] > 0.) temp_elev = dailyTempC - ((c->thresh_elev[n] - elev[0]) * P_T_GRADNT);
            }
            if (temp_elev <= P_T_SNOWFZ) {
                dailyEffPrecBeforeSnowMelt_elev = 0.;
                daily_snow_to_soil_elev = daily_prec_to_soil;
                S[e] += daily_snow_to_soil_elev;
                if (S[e] > dailySoilPET) {
                    S[e] -= dailySoilPET;
                    dailySnowEvapo += dailySoilPET;
                } else {
                    dailySnowEvapo += S[e];
                    S[e] = 0.;
                }
            } else {
                dailyEffPrecBeforeSnowMelt_elev = daily_prec_to_soil;
            }
            if (temp_elev > P_T_SNOWMT) {
                if (S[e] < 0.) {
                    fprintf(stderr, "G_SnowInElevation(n,elev) < 0 \n");
                } else {
                    snowmelt_elev = M_DEGDAY_F * c->lct_ddf[lc] * (temp_elev - P_T_SNOWMT);
                    if (snowmelt_elev > S[e]) {
                        snowmelt_elev = S[e];
                        S[e] = 0.;
                    } else {
                        S[e] -= snowmelt_elev;
                    }
                }
            }
            dailyEffPrec_elev = dailyEffPrecBeforeSnowMelt_elev + snowmelt_elev;
            snowStorageChange += S[e] - initialStorage;
            if (e == 1) TempElevMax = temp_elev;
            c->snow[n] += S[e];
            dailyEffPrec += dailyEffPrec_elev;
        }
    }
    if (landAreaFrac > 0.) {
        c->snow[n] /= 100.;
        dailyEffPrec /= 100.;
        dailySnowEvapo /= 100.;
        snowStorageChange /= 100.;
        landStorageChangeSum += snowStorageChange;
    }

    /* immediate runoff over built-up area, :1068-1071 */
    if (c->builtup[n] > 0.) {
        immediate_runoff = runoffFracBuiltUp * dailyEffPrec * c->builtup[n];
        dailyEffPrec -= immediate_runoff;
    }

    /* soil and AET, :1080-1239 */
    const double Smax = c->smax[n]; /* float grid promoted on use */
    if (landAreaFrac <= 0.) {
        c->storage_transfer[n] += c->soil[n];
        c->storage_transfer[n] *= cfa;
        c->soil[n] = 0.;
        daily_gw_recharge = 0.;
        total_daily_runoff = 0.;
        c->gw_recharge[n] = 0.;
        c->land_aet[n] = 0.;
        c->land_aet_uncorr[n] = 0.;
    } else {
        c->soil[n] *= lafPrev / landAreaFrac;
        initialStorage = c->soil[n];
        soil_water_overflow = 0;
        if (c->soil[n] > Smax) {
            soil_water_overflow = c->soil[n] - Smax;
            c->soil[n] = Smax;
        }
        if (TempElevMax > P_T_SNOWFZ) {
            if (Smax > 0.) {
                soil_saturation = c->soil[n] / Smax;
                daily_runoff = dailyEffPrec * pow(soil_saturation, (double)c->gamma_hbv[n]);
                if (dailySoilPET > (maxDailyPET - dailyCanopyEvapo) * soil_saturation)
                    dailyAET = (maxDailyPET - dailyCanopyEvapo) * soil_saturation;
                else
                    dailyAET = dailySoilPET;
                c->soil[n] += dailyEffPrec - dailyAET - daily_runoff;
                if (fabs(c->soil[n]) <= MIN_STOR_VOL) c->soil[n] = 0.;
                dailyEffPrec = 0.;
                if (c->soil[n] < 0.) {
                    dailyAET += c->soil[n];
                    c->soil[n] = 0.;
                }
                daily_runoff *= cfa;
                immediate_runoff *= cfa;
                /* groundwater recharge, :1165-1183 (GW.getRgmax short, GW.getgwFactor float) */
                const short Rgmax = c->rgmax[n];
                const float gwFactor = c->gwfactor[n];
                if (((arid_gw) && (c->texture[n] < 21)) && (c->ldd[n] >= 0)) {
                    pot_gw_recharge = 0.;
                    if ((Rgmax / 100.) < (gwFactor * daily_runoff)) daily_gw_recharge = Rgmax / 100.;
                    else daily_gw_recharge = gwFactor * daily_runoff;
                    if (dailyPrec <= PAR(c, WGO_P_PCRITGWA, n)) {
                        pot_gw_recharge = daily_gw_recharge;
                        daily_gw_recharge = 0.;
                    }
                } else {
                    pot_gw_recharge = 0.;
                    if ((Rgmax / 100.) < (gwFactor * daily_runoff)) daily_gw_recharge = Rgmax / 100.;
                    else daily_gw_recharge = gwFactor * daily_runoff;
                }
                daily_runoff -= pot_gw_recharge;
                pot_gw_recharge /= cfa;
                c->soil[n] += pot_gw_recharge;
                if (c->soil[n] > Smax) {
                    soil_water_overflow += c->soil[n] - Smax;
                    c->soil[n] = Smax;
                }
                soil_water_overflow *= cfa;
                total_daily_runoff = daily_runoff + immediate_runoff + soil_water_overflow;
            } else {
                total_daily_runoff = 0.;
                daily_gw_recharge = 0.;
            }
        } else {
            soil_water_overflow *= cfa;
            dailyEffPrec *= cfa;
            total_daily_runoff += soil_water_overflow + dailyEffPrec;
            daily_gw_recharge = 0.;
            dailyAET = 0.;
        }
        c->gw_recharge[n] = daily_gw_recharge; /* :1221 */
        landStorageChangeSum += c->soil[n] - initialStorage;
        c->land_aet[n] = landStorageChangeSum * (cfa - 1.0) - dailyPrec * (cfa - 1.0)
                         + (dailyAET + dailyCanopyEvapo + dailySnowEvapo) * cfa;
        if (c->land_aet[n] < 0.) {
            neg_land_aet = c->land_aet[n];
            c->land_aet[n] = 0.;
        }
        c->land_aet_uncorr[n] = (dailyAET + dailyCanopyEvapo + dailySnowEvapo);
    }

    /* surface runoff, :1244-1257 */
    if (neg_land_aet < 0.) total_daily_runoff = total_daily_runoff + neg_land_aet;
    if (total_daily_runoff < 0.) total_daily_runoff = 0.;
    if ((total_daily_runoff - daily_gw_recharge) < 0.) {
        neg_runoff = total_daily_runoff - daily_gw_recharge;
        daily_gw_recharge = total_daily_runoff;
        c->soil[n] += neg_runoff;
        neg_runoff = 0.;
    }
    surface_runoff = total_daily_runoff - daily_gw_recharge;
    c->surface_runoff[n] = surface_runoff;
    (void)neg_runoff;
}

void wgo_vertical_day(wgo_ctx *c, int day, int month, int day_in_month) {
    (void)day;
    (void)month;
    for (int n = 0; n < c->ncell; n++)
        if (c->contcell[n]) vertical_cell(c, n, day_in_month); /* integrateWGHM.cpp:770-783 */
}

/* ------------------------------------------------------------------------------------------
 * routing
 * ---------------------------------------------------------------------------------------- */
/* routingClass::getRiverVelocity, routing.cpp:7274-7307 */
static double river_velocity(double RiverSlope, double bottomWidth, double Roughness, double riverInflow,
                             double M_RIVRGH_C) {
    double incoming_discharge = (riverInflow * 1000. * 1000. * 1000.) / (60. * 60. * 24.);
    double riverDepth = 0.349 * pow(incoming_discharge, 0.341);
    double crossSectionalArea = riverDepth * (2.0 * riverDepth + bottomWidth);
    double wettedPerimeter = bottomWidth + 2.0 * riverDepth * sqrt(5.0);
    double hydraulicRad = crossSectionalArea / wettedPerimeter;
    double v = 1. / (M_RIVRGH_C * Roughness) * pow(hydraulicRad, (2. / 3.)) * pow(RiverSlope, 0.5);
    v = v * 86.4;
    if (v < 0.00001) return 0.00001;
    return v;
}

/* groundwater linear reservoir, identical text at routing.cpp:1938-1958, 1986-2006, 2075-2094,
 * 2131-2150, 3331-3346 */
static double gw_step(double *Sg, double netGWin, double k) {
    double prev = *Sg;
    *Sg = prev * exp(-1. * k) + (1. / k) * netGWin * (1. - exp(-1. * k));
    if (fabs(*Sg) <= MIN_STOR_VOL) *Sg = 0.;
    double q = prev - *Sg + netGWin;
    if (q <= 0.) {
        q = 0.;
        *Sg = prev + netGWin;
        if (fabs(*Sg) <= MIN_STOR_VOL) *Sg = 0.;
    }
    return q;
}

/* routingClass::updateNetAbstractionGW, routing.cpp:5503-5572: the day's net abstraction from groundwater, adapted to the
 * surface-water use that stayed unsatisfied the day before (return flows reduced) or was satisfied late (reintroduced) */
static double update_net_abstraction_gw(wgo_ctx *c, int n) {
    double NAgnew, returnflowChange, eff, frgi, factor, WUsiNew, dailyRemainingUseFromIrrig, reintroducedRatio;
    if (c->wu_daily_remaining[n] > 1.e-12) {
        if (c->wu_wusi[n] > 0.) {
            eff = c->wu_cusi[n] / c->wu_wusi[n];
            frgi = c->wu_frgi[n];
            factor = 1 - (1 - frgi) * (1 - eff);
            WUsiNew = 1 / factor * (c->wu_wusi[n] * factor - c->wu_daily_remaining[n]);
            if (WUsiNew < 0.) {
                WUsiNew = 0.;
                c->wu_uns_irr[n] += c->wu_wusi[n] * factor;
                c->wu_uns_oth[n] += c->wu_daily_remaining[n] - (c->wu_wusi[n] * factor);
            } else {
                c->wu_uns_irr[n] += c->wu_daily_remaining[n];
            }
            returnflowChange = (frgi * (1 - eff) * (WUsiNew - c->wu_wusi[n]));
            c->wu_red_rf[n] += returnflowChange;
            NAgnew = c->wu_daily_nug[n] - returnflowChange;
            c->wu_daily_remaining[n] = 0.;
            return NAgnew;
        } else {
            c->wu_uns_oth[n] += c->wu_daily_remaining[n];
            return c->wu_daily_nug[n];
        }
    } else if (c->wu_daily_remaining[n] < -1.e-12) {
        dailyRemainingUseFromIrrig = c->wu_daily_remaining[n] + c->wu_uns_oth[n];
        if (dailyRemainingUseFromIrrig < 0.) {
            c->wu_uns_oth[n] = 0.;
        } else {
            c->wu_uns_oth[n] += c->wu_daily_remaining[n];
            c->wu_daily_remaining[n] = 0.;
            return c->wu_daily_nug[n];
        }
        if (c->wu_uns_irr[n] == 0.) {
            c->wu_daily_remaining[n] = 0.;
            return c->wu_daily_nug[n];
        }
        reintroducedRatio = (dailyRemainingUseFromIrrig / c->wu_uns_irr[n]);
        if (reintroducedRatio < -1.) {
            reintroducedRatio = -1.;
            dailyRemainingUseFromIrrig = c->wu_uns_irr[n] * -1.;
        }
        returnflowChange = (reintroducedRatio * c->wu_red_rf[n]);
        c->wu_uns_irr[n] += dailyRemainingUseFromIrrig;
        c->wu_red_rf[n] += returnflowChange;
        NAgnew = c->wu_daily_nug[n] - returnflowChange;
        c->wu_daily_remaining[n] = 0.;
        return NAgnew;
    }
    return c->wu_daily_nug[n];
}
/* the adaptation in front of each of the four groundwater balances (routing.cpp:1929-1935, 1991-1996, 2076-2081, 2132-2137, 3325-3330) */
static double net_gw_use(wgo_ctx *c, int n) {
    if (c->subtract_use <= 0) return 0.;
    if (c->wu_daily_remaining[n] != 0.) c->wu_daily_nug[n] = update_net_abstraction_gw(c, n);
    return c->wu_daily_nug[n];
}

void wgo_routing_day(wgo_ctx *c, int day, int month, int day_in_month) {
    (void)day_in_month;
    const int ng = c->ncell;
    const int wu = c->subtract_use > 0;
    static const int numberOfDaysInMonthWU[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    (void)numberOfDaysInMonthWU;
    if (wu) /* routingClass::calcNextDay_M, routing.cpp:7432-7440 (integrateWGHM.cpp:794-796) */
        for (int n = 0; n < ng; n++) {
            c->wu_daily_nus[n] = c->wu_nus_month[n];
            c->wu_daily_nug[n] = c->wu_nug_month[n];
        }
    const short first_day_in_month[12] = {1, 32, 60, 91, 121, 152, 182, 213, 244, 274, 305, 335}; /* routing.cpp:1634 */
    const double lakeOutflowExp = 1.5, wetlOutflowExp = 2.5, evapoReductionExp = 3.32193,
                 evapoReductionExpReservoir = 2.81383; /* :124-129 */
    const double Kswbgw = 10.;                         /* :1664 */

    for (int n = 0; n < ng; n++) { /* :1772-1811 */
        c->river_inflow[n] = 0.;
        c->river_evapo[n] = 0.;
    }

    for (int routingCell = 0; routingCell < ng; routingCell++) { /* :1835 */
        const int n = c->routing_cell[routingCell] - 1;
        const double kG = PAR(c, WGO_P_GWOUTF_C, n);
        const double M_EVAREDEX = PAR(c, WGO_M_EVAREDEX, n);
        if (!c->contcell[n]) continue; /* :1878 */
        double cellArea = c->area[n];
        const double cfa = c->cfa[n];
        const double owPrec = c->openwater_prec[n], owPET = c->openwater_pet[n];
        const int ldd = c->ldd[n];
        const int aridc = (1 == c->arid[n]) && (ldd >= 0); /* aridareaOpt == 1 */
        double dailyLocalSurfaceRunoff;
        double localRunoff = 0., localRunoffIntoRiver = 0., localGWRunoff = 0., localGWRunoffIntoRiver = 0.;
        double fswb_catchment = 0.;
        double gwr_loclak = 0., gwr_glolak = 0., gwr_locwet = 0., gwr_glowet = 0., gwr_res = 0.;
        double groundwater_runoff_km3, netGWin;

        /* :1885-1891 */
        if (c->land_area_frac[n] <= 0.)
            dailyLocalSurfaceRunoff = c->storage_transfer[n] * cellArea / 1000000. * c->land_area_frac_prev[n] / 100.;
        else
            dailyLocalSurfaceRunoff = c->surface_runoff[n] * cellArea / 1000000. * c->land_area_frac[n] / 100.;
        /* :1898-1908 */
        if (ldd >= 0) {
            fswb_catchment = c->fswb_init[n] * 20.;
            if (fswb_catchment > 1.) fswb_catchment = 1.;
            localRunoffIntoRiver = (1. - fswb_catchment) * dailyLocalSurfaceRunoff;
        }
        /* arid, not an inland sink: GW is done after the surface water bodies (:1910-1915) */
        if ((1 == c->arid[n]) && (ldd >= 0)) {
            localGWRunoff = 0.;
            localRunoff = fswb_catchment * dailyLocalSurfaceRunoff;
        }
        /* humid, not an inland sink (:1979-2033) */
        if ((0 == c->arid[n]) && (ldd >= 0)) {
            netGWin = (double)c->gw_recharge[n] * cellArea * ((double)c->land_area_frac[n] / 100.) / 1000000.;
            if (wu) netGWin -= net_gw_use(c, n);
            groundwater_runoff_km3 = gw_step(&c->gw[n], netGWin, kG);
            localGWRunoffIntoRiver = (1. - fswb_catchment) * groundwater_runoff_km3;
            localGWRunoff = fswb_catchment * groundwater_runoff_km3;
            localRunoff = (fswb_catchment * dailyLocalSurfaceRunoff) + localGWRunoff;
        }
        /* inland sinks (:2123-2176) */
        if (ldd < 0) {
            netGWin = c->gw_recharge[n] * cellArea * (c->land_area_frac[n] / 100.) / 1000000.;
            if (wu) netGWin -= net_gw_use(c, n);
            groundwater_runoff_km3 = gw_step(&c->gw[n], netGWin, kG);
            if (c->land_area_frac[n] == 0.)
                dailyLocalSurfaceRunoff = c->storage_transfer[n] * cellArea / 1000000. * c->land_area_frac_prev[n] / 100.;
            else
                dailyLocalSurfaceRunoff = c->surface_runoff[n] * cellArea / 1000000. * c->land_area_frac[n] / 100.;
            localGWRunoff = groundwater_runoff_km3;
            localRunoff = dailyLocalSurfaceRunoff + localGWRunoff;
        }

        /* :2193-2296 with use_alloc 0, delayedUseSatisfaction 0, aggrNUsGloLakResOpt 0: the day's desired use is the net
         * abstraction from surface water (negative = return flow) */
        double remainingUse = 0., dailyActualUse = 0., remainingUseGloLake = 0., remainingUseRes = 0.;
        if (wu && 1 == c->toBeCalculated[n]) remainingUse = c->wu_daily_nus[n];

        double transportedVolume = 0., inflowUpstream = 0.;
        if (0 != c->toBeCalculated[n]) { /* :2308 */
            double inflow = localRunoff, outflow = 0., maxStorage, totalInflow, PETgwr, PETgwrMax, locLakeMaxStorage = 0.;
            const double kS = PAR(c, WGO_P_SWOUTF_C, n);
            const double contf = c->contfreq[n];

            /* local lake, :2318-2490 */
            if (c->loc_lake[n] > 0.) {
                const double prev = c->loc_lake_stor[n];
                maxStorage = ((c->loc_lake[n]) / 100.) * cellArea * c->lake_depth_active[n];
                locLakeMaxStorage = maxStorage; /* G_locLakeMaxStorage, :2321 */
                const double r = c->red_loc_lake[n];
                double evapo = ((1.0 - cfa) * owPrec * r) + (cfa * (owPET * r));
                if (evapo < 0.) evapo = 0.;
                totalInflow = inflow + (owPrec * r) * (cellArea / 1000000.) * (c->loc_lake[n] / 100.);
                if (aridc) gwr_loclak = Kswbgw * r * c->loc_lake[n] / 100. / (contf / 100.);
                else gwr_loclak = 0.;
                PETgwr = evapo * (cellArea / 1000000.) * (c->loc_lake[n] / 100.)
                         + gwr_loclak * cellArea * (contf / 100.) / 1000000.;
                PETgwrMax = prev + maxStorage + totalInflow;
                if (PETgwrMax < 0.) PETgwrMax = 0.;
                if (PETgwr > PETgwrMax) {
                    c->loc_lake_stor[n] = (-1.) * maxStorage;
                    gwr_loclak *= PETgwrMax / PETgwr;
                    evapo *= PETgwrMax / PETgwr;
                } else {
                    c->loc_lake_stor[n] = prev + totalInflow - PETgwr;
                }
                if (prev > 0.) {
                    outflow = kS * prev * pow((prev / maxStorage), lakeOutflowExp);
                    if (c->loc_lake_stor[n] <= 0.) outflow = 0;
                    else if (outflow > c->loc_lake_stor[n]) outflow = c->loc_lake_stor[n];
                } else
                    outflow = 0.;
                c->loc_lake_stor[n] -= outflow;
                if (fabs(c->loc_lake_stor[n]) <= MIN_STOR_VOL) c->loc_lake_stor[n] = 0.;
                if (c->loc_lake_stor[n] > maxStorage) {
                    outflow += (c->loc_lake_stor[n] - maxStorage);
                    c->loc_lake_stor[n] = maxStorage;
                }
                inflow = outflow;
                c->red_loc_lake[n] = 1. - pow(fabs(c->loc_lake_stor[n] - maxStorage) / (2. * maxStorage),
                                              (M_EVAREDEX * evapoReductionExp));
                if (c->red_loc_lake[n] < 0.) c->red_loc_lake[n] = 0.;
                if (c->red_loc_lake[n] > 1.) c->red_loc_lake[n] = 1.;
            }

            /* local wetland, :2495-2617 */
            if (c->loc_wetland[n] > 0.) {
                const double prev = c->loc_wetl_stor[n];
                maxStorage = ((c->loc_wetland[n]) / 100.) * cellArea * c->wetl_depth_active[n];
                const double r = c->red_loc_wetl[n];
                double evapo = ((1.0 - cfa) * owPrec * r) + (cfa * (owPET * r));
                if (evapo < 0.) evapo = 0.;
                totalInflow = inflow + (owPrec * r * (cellArea / 1000000.) * (c->loc_wetland[n] / 100.));
                if (aridc) gwr_locwet = Kswbgw * r * c->loc_wetland[n] / 100. / (contf / 100.);
                else gwr_locwet = 0.;
                PETgwr = evapo * (cellArea / 1000000.) * (c->loc_wetland[n] / 100.)
                         + gwr_locwet * cellArea * (contf / 100.) / 1000000.;
                PETgwrMax = prev + totalInflow;
                if (PETgwr > PETgwrMax) {
                    c->loc_wetl_stor[n] = 0.;
                    gwr_locwet *= PETgwrMax / PETgwr;
                    evapo *= PETgwrMax / PETgwr;
                } else {
                    c->loc_wetl_stor[n] = prev + totalInflow - PETgwr;
                }
                if (fabs(c->loc_wetl_stor[n]) <= MIN_STOR_VOL) c->loc_wetl_stor[n] = 0.;
                if (c->loc_wetl_stor[n] > 0.) {
                    outflow = kS * c->loc_wetl_stor[n] * pow((c->loc_wetl_stor[n] / maxStorage), wetlOutflowExp);
                    if (outflow > c->loc_wetl_stor[n]) outflow = c->loc_wetl_stor[n];
                } else
                    outflow = 0.;
                c->loc_wetl_stor[n] -= outflow;
                if (c->loc_wetl_stor[n] > maxStorage) {
                    outflow += (c->loc_wetl_stor[n] - maxStorage);
                    c->loc_wetl_stor[n] = maxStorage;
                }
                inflow = outflow;
                c->red_loc_wetl[n] = 1. - pow(fabs(c->loc_wetl_stor[n] - maxStorage) / (maxStorage),
                                              (M_EVAREDEX * evapoReductionExp));
                if (c->red_loc_wetl[n] < 0.) c->red_loc_wetl[n] = 0.;
                if (c->red_loc_wetl[n] > 1.) c->red_loc_wetl[n] = 1.;
            }

            /* water from upstream cells joins here, :2623-2624 */
            inflow += c->river_inflow[n];
            inflowUpstream = c->river_inflow[n];

            /* global lake, :2630-2804 */
            if (c->lake_area[n] > 0.) {
                const double prev = c->glo_lake_stor[n];
                maxStorage = ((double)c->lake_area[n]) * c->lake_depth_active[n];
                const double r = c->red_glo_lake[n];
                double evapo = ((1.0 - cfa) * owPrec) + (cfa * owPET * r);
                if (evapo < 0.) evapo = 0.;
                totalInflow = inflow + (owPrec * ((double)c->lake_area[n] / 1000000.));
                if (aridc) gwr_glolak = Kswbgw * r * ((double)c->lake_area[n] / (cellArea * (contf / 100.)));
                else gwr_glolak = 0.;
                if (c->reservoir_area[n] > 0.) remainingUseGloLake = 0.5 * remainingUse; /* :2681-2686 */
                else remainingUseGloLake = remainingUse;
                const double remainingUseGloLakeStart = remainingUseGloLake;
                double PETgwrRemUse = evapo * ((double)c->lake_area[n] / 1000000.)
                                      + gwr_glolak * cellArea * (contf / 100.) / 1000000. + remainingUseGloLake;
                double PETgwrRemUseMax = totalInflow + maxStorage + prev;
                if (PETgwrRemUse > PETgwrRemUseMax) {
                    c->glo_lake_stor[n] = (-1.) * maxStorage;
                    outflow = 0.;
                    evapo *= PETgwrRemUseMax / PETgwrRemUse;
                    gwr_glolak *= PETgwrRemUseMax / PETgwrRemUse;
                    if (remainingUseGloLake > 0.) remainingUseGloLake -= remainingUseGloLake * PETgwrRemUseMax / PETgwrRemUse; /* :2710-2716 */
                    else remainingUseGloLake = 0.;
                } else {
                    c->glo_lake_stor[n] = prev * exp(-1. * kS) + (1. / kS) * (totalInflow - PETgwrRemUse) * (1. - exp(-1. * kS));
                    outflow = totalInflow + prev - c->glo_lake_stor[n] - PETgwrRemUse;
                    if (c->glo_lake_stor[n] > maxStorage) {
                        outflow += (c->glo_lake_stor[n] - maxStorage);
                        c->glo_lake_stor[n] = maxStorage;
                    }
                    if (outflow < 0.) {
                        outflow = 0.;
                        c->glo_lake_stor[n] = prev + totalInflow - PETgwrRemUse;
                    }
                    remainingUseGloLake = 0.; /* :2752 */
                }
                if (fabs(c->glo_lake_stor[n]) <= MIN_STOR_VOL) c->glo_lake_stor[n] = 0.;
                inflow = outflow;
                dailyActualUse = remainingUseGloLakeStart - remainingUseGloLake; /* :2788 */
                c->red_glo_lake[n] = 1. - pow(fabs(c->glo_lake_stor[n] - maxStorage) / (2. * maxStorage),
                                              (M_EVAREDEX * evapoReductionExp));
                if (c->red_glo_lake[n] < 0.) c->red_glo_lake[n] = 0.;
                if (c->red_glo_lake[n] > 1.) c->red_glo_lake[n] = 1.;
            }

            /* reservoir (Hanasaki), :2807-3082 */
            if (c->reservoir_area[n] > 0.) {
                double c_ratio, prov_rel = 0., release, dailyUse, monthlyUse;
                c_ratio = c->stor_cap[n] / (c->mean_outflow[n] * 31536000. / 1000000000.);
                maxStorage = c->stor_cap[n];
                const double prev = c->res_stor[n];
                const double r = c->red_res[n];
                double evapo = ((1.0 - cfa) * owPrec) + (cfa * (owPET * r));
                if (evapo < 0.) evapo = 0.;
                totalInflow = inflow + (owPrec * ((double)c->reservoir_area[n] / 1000000.));
                if (aridc) gwr_res = Kswbgw * r * ((double)c->reservoir_area[n] / (cellArea * (contf / 100.)));
                else gwr_res = 0.;
                if (c->lake_area[n] > 0.) remainingUseRes = 0.5 * remainingUse + remainingUseGloLake; /* :2869-2875 */
                else remainingUseRes = remainingUse;
                const double remainingUseResStart = remainingUseRes;
                PETgwr = evapo * ((double)c->reservoir_area[n] / 1000000.) + gwr_res * cellArea * (contf / 100.) / 1000000.;
                PETgwrMax = prev + totalInflow;
                if (PETgwr > PETgwrMax) {
                    c->res_stor[n] = prev + totalInflow - PETgwrMax;
                    gwr_res *= PETgwrMax / PETgwr;
                    evapo *= PETgwrMax / PETgwr;
                } else {
                    c->res_stor[n] = prev + totalInflow - PETgwr;
                }
                if (wu) { /* :2922-2946 */
                    if (remainingUseRes < 0.) {
                        c->res_stor[n] -= remainingUseRes;
                        remainingUseRes = 0.;
                    } else {
                        if (c->res_stor[n] > (0.1 * c->stor_cap[n])) {
                            if (remainingUseRes < (c->res_stor[n] - (c->stor_cap[n] * 0.1))) {
                                c->res_stor[n] -= remainingUseRes;
                                remainingUseRes = 0.;
                            } else {
                                remainingUseRes -= (c->res_stor[n] - (c->stor_cap[n] * 0.1));
                                c->res_stor[n] = c->stor_cap[n] * 0.1;
                            }
                        }
                    }
                }
                if (fabs(c->res_stor[n]) <= MIN_STOR_VOL) c->res_stor[n] = 0.;
                if (month == c->start_month[n] - 1) { /* :2945-2956 */
                    if (day == first_day_in_month[month]) {
                        if (c->res_stor[n] < (c->stor_cap[n] * 0.1)) c->k_release[n] = 0.1;
                        else c->k_release[n] = c->res_stor[n] / (maxStorage * 0.85);
                    }
                }
                if ((c->res_type[n] + 0) == 1) { /* irrigation; without water use dailyUse == 0 (:2960-2977) */
                    dailyUse = 0.0;
                    static const int numberOfDaysInMonth[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
                    if (wu) { /* :2961-2973: own use plus the share of up to 5 downstream cells without a reservoir (note the
                                 reference's un-decremented index into G_reservoir_area) */
                        dailyUse = c->wu_daily_nus[n];
                        short i = 0;
                        int downstreamCell = c->downstream_cell[n];
                        while (i < 5 && downstreamCell > 0 && downstreamCell < ng && c->reservoir_area[downstreamCell] <= 0) {
                            dailyUse += c->wu_daily_nus[downstreamCell - 1] * c->wu_alloc_coeff[(size_t)n * 5 + i++];
                            downstreamCell = c->downstream_cell[downstreamCell - 1];
                        }
                    }
                    monthlyUse = dailyUse * numberOfDaysInMonth[month];
                    monthlyUse = monthlyUse * 1000000000. / (numberOfDaysInMonth[month] * 86400.);
                    if (c->mean_demand[n] >= 0.5 * c->mean_outflow[n])
                        prov_rel = c->mean_outflow[n] / 2. * (1. + monthlyUse / c->mean_demand[n]);
                    else
                        prov_rel = c->mean_outflow[n] + monthlyUse - c->mean_demand[n];
                } else if ((c->res_type[n] + 0) == 2) {
                    prov_rel = c->mean_outflow[n];
                } else
                    fprintf(stderr, "unknown reservoir type in gcrcNumber %d\n", n + 1);
                if (c_ratio >= 0.5) release = c->k_release[n] * prov_rel;
                else
                    release = ((4. * c_ratio * c_ratio) * c->k_release[n] * prov_rel)
                              + ((1.0 - ((4. * c_ratio * c_ratio))) * inflow * 1000000000. / (24. * 3600.));
                if (c->res_stor[n] >= (c->stor_cap[n] * 0.1)) outflow = release * (24. * 3600.) / 1000000000.;
                else outflow = 0.1 * release * (24. * 3600.) / 1000000000.;
                if (outflow < 0.) outflow = 0.;
                c->res_stor[n] -= outflow;
                if (c->res_stor[n] > maxStorage) {
                    outflow += (c->res_stor[n] - maxStorage);
                    c->res_stor[n] = maxStorage;
                }
                if (c->res_stor[n] < 0.) {
                    outflow += c->res_stor[n];
                    c->res_stor[n] = 0.;
                }
                inflow = outflow;
                dailyActualUse += remainingUseResStart - remainingUseRes; /* :3068 */
                c->red_res[n] = 1. - pow(fabs(c->res_stor[n] - maxStorage) / maxStorage, evapoReductionExpReservoir);
                if (c->red_res[n] < 0.) c->red_res[n] = 0.;
                if (c->red_res[n] > 1.) c->red_res[n] = 1.;
            }

            /* :3166-3172 (aggrNUsGloLakResOpt 0) */
            if (wu && (c->reservoir_area[n] > 0. || c->lake_area[n] > 0.)) {
                if (c->reservoir_area[n] > 0.) remainingUse = remainingUseRes;
                else if (c->lake_area[n] > 0.) remainingUse = remainingUseGloLake;
            }

            /* global wetland, :3178-3297 */
            if (c->glo_wetland[n] > 0) {
                const double prev = c->glo_wetl_stor[n];
                maxStorage = ((c->glo_wetland[n]) / 100.) * cellArea * c->wetl_depth_active[n];
                const double r = c->red_glo_wetl[n];
                double evapo = ((1.0 - cfa) * (owPrec * r)) + (cfa * (owPET * r));
                if (evapo < 0.) evapo = 0.;
                totalInflow = inflow + (owPrec * r * (cellArea / 1000000.) * (c->glo_wetland[n] / 100.));
                if (aridc) gwr_glowet = Kswbgw * r * c->glo_wetland[n] / 100. / (contf / 100.);
                else gwr_glowet = 0.;
                PETgwr = evapo * (cellArea / 1000000.) * ((c->glo_wetland[n]) / 100.)
                         + gwr_glowet * cellArea * (contf / 100.) / 1000000.;
                PETgwrMax = totalInflow + prev;
                if (PETgwr > PETgwrMax) {
                    c->glo_wetl_stor[n] = 0.;
                    outflow = 0.;
                    gwr_glowet *= PETgwrMax / PETgwr;
                    evapo *= PETgwrMax / PETgwr;
                } else {
                    c->glo_wetl_stor[n] = prev * exp(-1. * kS) + (1. / kS) * (totalInflow - PETgwr) * (1. - exp(-1. * kS));
                    outflow = totalInflow + prev - c->glo_wetl_stor[n] - PETgwr;
                }
                if (c->glo_wetl_stor[n] > maxStorage) {
                    outflow += (c->glo_wetl_stor[n] - maxStorage);
                    c->glo_wetl_stor[n] = maxStorage;
                }
                if (fabs(c->glo_wetl_stor[n]) <= MIN_STOR_VOL) c->glo_wetl_stor[n] = 0.;
                inflow = outflow;
                c->red_glo_wetl[n] = 1. - pow(fabs(c->glo_wetl_stor[n] - maxStorage) / maxStorage,
                                              (M_EVAREDEX * evapoReductionExp));
                if (c->red_glo_wetl[n] < 0.) c->red_glo_wetl[n] = 0.;
                if (c->red_glo_wetl[n] > 1.) c->red_glo_wetl[n] = 1.;
            }

            /* (semi-)arid groundwater incl. recharge below surface water bodies, :3305-3386 */
            if (aridc) {
                c->gwr_swb[n] = gwr_loclak + gwr_glolak + gwr_locwet + gwr_glowet + gwr_res;
                netGWin = c->gwr_swb[n] * cellArea * (contf / 100.) / 1000000.
                          + (double)c->gw_recharge[n] * cellArea * ((double)c->land_area_frac[n] / 100.) / 1000000.;
                if (wu) netGWin -= net_gw_use(c, n);
                groundwater_runoff_km3 = gw_step(&c->gw[n], netGWin, kG);
                localGWRunoffIntoRiver = groundwater_runoff_km3;
            }

            /* river, :3388-3586 */
            c->river_inflow[n] = inflow;
            if (ldd >= 0) {
                c->river_inflow[n] += localRunoffIntoRiver;
                c->river_inflow[n] += localGWRunoffIntoRiver;
            }
            double riverVelocity = river_velocity(c->river_slope[n], c->river_bottom_width[n], c->roughness[n],
                                                  c->river_inflow[n], PAR(c, WGO_M_RIVRGH_C, n));
            double K = riverVelocity / c->river_length[n];
            const double prevR = c->river_stor[n];
            c->river_area_frac[n] = c->river_area_frac_next[n];
            c->river_evapo[n] = ((1.0 - cfa) * (owPrec) + (cfa * owPET)) * c->river_area_frac[n] / 100. * cellArea / 1000000.;
            double riverPrecip = owPrec * c->river_area_frac[n] / 100. * cellArea / 1000000.;
            c->river_inflow[n] += riverPrecip;
            double RiverEvapoRemUse = remainingUse + c->river_evapo[n];
            const double remainingUseRiverStart = remainingUse;
            double RivEvapoRemUseMax = c->river_inflow[n] + (K * prevR * exp(-1. * K)) / (1. - exp(-1. * K));
            if (RiverEvapoRemUse > RivEvapoRemUseMax) {
                c->river_stor[n] = 0.;
                transportedVolume = c->river_inflow[n] + prevR - RivEvapoRemUseMax;
                if (transportedVolume < 0.) transportedVolume = 0.;
                if (remainingUse > 0.) remainingUse -= remainingUse * RivEvapoRemUseMax / RiverEvapoRemUse; /* :3489-3492 */
                else remainingUse = 0.;
                c->river_evapo[n] *= RivEvapoRemUseMax / RiverEvapoRemUse;
            } else {
                c->river_stor[n] = prevR * exp(-1. * K) + (1. / K) * (c->river_inflow[n] - RiverEvapoRemUse) * (1. - exp(-1. * K));
                if (fabs(c->river_stor[n]) <= MIN_STOR_VOL) c->river_stor[n] = 0.;
                transportedVolume = c->river_inflow[n] + prevR - c->river_stor[n] - RiverEvapoRemUse;
                if (transportedVolume < 0.) transportedVolume = 0.;
                remainingUse = 0.; /* :3512 */
            }
            /* river width / area fraction of the next day, :3546-3586 */
            {
                double crossSectionalArea = c->river_stor[n] / c->river_length[n];
                double bw = c->river_bottom_width[n];
                double riverDepth = -bw / (4. * 1000.) + sqrt(bw / 1000. * bw / (16. * 1000.) + 0.5 * crossSectionalArea);
                double width = bw / 1000. + 4. * riverDepth;
                if (width > c->river_width_bf[n] / 1000.) width = c->river_width_bf[n] / 1000.;
                c->red_river[n] = 1. - pow(fabs(c->river_stor[n] - c->river_storage_max[n]) / c->river_storage_max[n],
                                           (M_EVAREDEX * evapoReductionExp));
                if (c->red_river[n] < 0.) c->red_river[n] = 0.;
                if (c->red_river[n] > 1.) c->red_river[n] = 1.;
                c->river_area_frac_next[n] = c->red_river[n] * c->river_length[n] * width * 100. / cellArea;
                c->river_area_frac_change[n] = c->river_area_frac_next[n] - c->river_area_frac[n];
            }
            if (wu) {
                /* what the river could not supply is taken from the local lake, :3590-3624 */
                if ((c->loc_lake[n] > 0.) && (remainingUse > 0.)) {
                    if (c->loc_lake_stor[n] > (-1.) * locLakeMaxStorage) {
                        if (remainingUse < (locLakeMaxStorage + c->loc_lake_stor[n])) {
                            c->loc_lake_stor[n] -= remainingUse;
                            remainingUse = 0.;
                        } else {
                            remainingUse -= (locLakeMaxStorage + c->loc_lake_stor[n]);
                            c->loc_lake_stor[n] = (-1.) * locLakeMaxStorage;
                        }
                    }
                    c->red_loc_lake[n] = 1. - pow(fabs(c->loc_lake_stor[n] - locLakeMaxStorage) / (2. * locLakeMaxStorage),
                                                  (M_EVAREDEX * evapoReductionExp));
                    if (c->red_loc_lake[n] < 0.) c->red_loc_lake[n] = 0.;
                    if (c->red_loc_lake[n] > 1.) c->red_loc_lake[n] = 1.;
                }
                dailyActualUse += remainingUseRiverStart - remainingUse; /* :3625 */
                c->wu_actual_use[n] += dailyActualUse;                   /* :3736 */
                /* use_alloc 0, delayedUseSatisfaction 0, :3889-3900 */
                c->wu_total_unsatisfied[n] += remainingUse;
                c->wu_daily_remaining[n] = remainingUse;
                c->wu_wusi[n] = c->wu_wusi_month[n]; /* :3907-3908 */
                c->wu_cusi[n] = c->wu_cusi_month[n];
            }
            /* :3923-3928 */
            if (ldd < 0) c->cell_runoff[n] = 0. - inflowUpstream;
            else c->cell_runoff[n] = (transportedVolume - inflowUpstream);
            /* push to the downstream cell, :3955-3958 */
            if ((c->downstream_cell[n] - 1) >= 0) {
                if (0 != c->toBeCalculated[(c->downstream_cell[n] - 1)])
                    c->river_inflow[(c->downstream_cell[n] - 1)] += transportedVolume;
            }
            /* reference keeps discharge of inland sinks out of G_daily365RiverAvail (:4219-4221) */
            c->discharge[n] = (ldd >= 0) ? transportedVolume : 0.;
        }
        c->river_inflow[n] = 0.; /* :4517 */
    }

    /* wghmState of the day, :5002-5020 */
    for (int n = 0; n < ng; n++) {
        double cellArea = c->area[n];
        double conv = ((cellArea * (c->contfreq[n] / 100.)) / 1000000.);
        double *w = c->wghm_routing_mm;
        w[0 * (size_t)ng + n] = c->loc_lake_stor[n] / conv;
        w[1 * (size_t)ng + n] = c->loc_wetl_stor[n] / conv;
        w[2 * (size_t)ng + n] = c->glo_lake_stor[n] / conv;
        w[3 * (size_t)ng + n] = c->glo_wetl_stor[n] / conv;
        w[4 * (size_t)ng + n] = c->res_stor[n] / conv;
        w[5 * (size_t)ng + n] = c->river_stor[n] / conv;
        w[6 * (size_t)ng + n] = c->gw[n] / conv;
    }

    /* surface-water-body fractions and next-day land area fraction, :5034-5188 */
    for (int n = 0; n < ng; n++) {
        double fLocLake, fLocWet, fGloWet;
        if ((c->loc_lake[n] > 0.) && (c->red_loc_lake[n] > 0.)) fLocLake = (c->red_loc_lake[n] * c->loc_lake[n] / 100.);
        else fLocLake = 0.;
        if ((c->loc_wetland[n] > 0.) && (c->red_loc_wetl[n] > 0.)) fLocWet = (c->red_loc_wetl[n] * c->loc_wetland[n] / 100.);
        else fLocWet = 0.;
        if ((c->glo_wetland[n] > 0.) && (c->red_glo_wetl[n] > 0.)) fGloWet = (c->red_glo_wetl[n] * c->glo_wetland[n] / 100.);
        else fGloWet = 0.;
        c->fswb_laf[n] = c->fswb_laf_next[n];
        c->fswb_laf_next[n] = fLocLake + fLocWet + fGloWet;
        double changePct = c->fswb_laf_next[n] * 100. - c->fswb_laf[n] * 100.0;
        /* riverEvapoOpt == 1 */
        double maxRiverAreaFrac = c->contfreq[n] / 100. - c->f_glo_lake[n];
        if (((c->lake_area[n] > 0.) || (c->reservoir_area[n] > 0.)) && (c->f_glo_lake[n] == 1.)) {
            c->river_area_frac_next[n] = 0.;
            c->river_area_frac_change[n] = 0.;
            c->red_river[n] = 0.;
            c->river_evapo[n] = 0.;
        } else {
            if (c->river_area_frac_next[n] <= maxRiverAreaFrac) {
                if (c->fswb_laf_next[n] > (maxRiverAreaFrac - c->river_area_frac_next[n])) {
                    double fswbFracCorr = (maxRiverAreaFrac - c->river_area_frac_next[n]) / c->fswb_laf_next[n];
                    c->red_loc_lake[n] *= fswbFracCorr;
                    c->red_loc_wetl[n] *= fswbFracCorr;
                    c->red_glo_wetl[n] *= fswbFracCorr;
                    if ((fLocLake > 0.) && (c->red_loc_lake[n] > 0.)) fLocLake = (c->red_loc_lake[n] * c->loc_lake[n] / 100.);
                    else { c->red_loc_lake[n] = 0.; fLocLake = 0.; }
                    if ((fLocWet > 0.) && (c->red_loc_wetl[n] > 0.)) fLocWet = (c->red_loc_wetl[n] * c->loc_wetland[n] / 100.);
                    else { c->red_loc_wetl[n] = 0.; fLocWet = 0.; }
                    if ((fGloWet > 0.) && (c->red_glo_wetl[n] > 0.)) fGloWet = (c->red_glo_wetl[n] * c->glo_wetland[n] / 100.);
                    else { c->red_glo_wetl[n] = 0.; fGloWet = 0.; }
                }
            } else {
                double riverAreaFracDeficit = c->river_area_frac_next[n] - maxRiverAreaFrac;
                c->river_area_frac_change[n] -= riverAreaFracDeficit;
                c->red_river[n] *= maxRiverAreaFrac / c->river_area_frac_next[n];
                c->river_area_frac_next[n] = maxRiverAreaFrac;
                fLocLake = 0.;
                fLocWet = 0.;
                fGloWet = 0.;
            }
        }
        c->fswb_laf_next[n] = fLocLake + fLocWet + fGloWet;
        changePct = c->fswb_laf_next[n] * 100. - c->fswb_laf[n] * 100.;
        c->status_laf_next[n] = 1;
        c->land_area_frac_next[n] = c->land_area_frac[n] - (changePct + (c->river_area_frac_change[n] * 100.));
        if (c->land_area_frac_next[n] < 0.) c->land_area_frac_next[n] = 0.;
    }
}

/* routingClass::updateLandAreaFrac, routing.cpp:5343-5352 */
void wgo_update_land_area_frac(wgo_ctx *c) {
    for (int n = 0; n < c->ncell; n++) {
        c->land_area_frac_prev[n] = c->land_area_frac[n];
        c->land_area_frac[n] = c->land_area_frac_next[n];
    }
}

void wgo_step_days(wgo_ctx *c, int day, int month, int day_in_month, int ndays) {
    for (int i = 0; i < ndays; i++) {
        wgo_vertical_day(c, day + i, month, day_in_month + i);
        wgo_routing_day(c, day + i, month, day_in_month + i);
        wgo_update_land_area_frac(c);
    }
}

double wgo_total_storage_km3(const wgo_ctx *c) {
    double s = 0.;
    for (int n = 0; n < c->ncell; n++) {
        double laf = (0 == c->status_laf_next[n]) ? c->land_area_frac[n] : c->land_area_frac_next[n];
        double land_km3 = (c->canopy[n] + c->snow[n] + c->soil[n]) * c->area[n] / 1000000. * laf / 100.;
        s += land_km3 + c->gw[n] + c->loc_lake_stor[n] + c->loc_wetl_stor[n] + c->glo_lake_stor[n] + c->glo_wetl_stor[n]
             + c->res_stor[n] + c->river_stor[n];
    }
    return s;
}

/* ------------------------------------------------------------------------------------------
 * flow topology: rout_prepare.cpp
 * ---------------------------------------------------------------------------------------- */
/* padded index raster GCRC_2 (rout_prepare.cpp:134-147): (ncol+2) x (nrow+2), wraps in longitude */
static int32_t *make_gcrc2(const wgo_topology *t) {
    const int nc = t->ncol, nr = t->nrow;
    int32_t *g2 = (int32_t *)calloc((size_t)(nc + 2) * (nr + 2), sizeof(int32_t));
#define G2(j, i) g2[(size_t)(j) * (nr + 2) + (i)]
#define G1(j, i) t->gcrc[(size_t)(j) * nr + (i)]
    for (int i = 0; i < nr; i++) {
        G2(0, i + 1) = G1(nc - 1, i);
        G2(nc + 1, i + 1) = G1(0, i);
        for (int j = 0; j < nc; j++) G2(j + 1, i + 1) = G1(j, i);
    }
    return g2;
}

/* inflow_cells, rout_prepare.cpp:338-439 */
static void inflow_cells(const wgo_topology *t, const int32_t *g2, const int8_t *ldd, int32_t *inf) {
    const int nr = t->nrow;
    memset(inf, 0, sizeof(int32_t) * 9 * (size_t)t->ncell);
    static const int dcol[8] = {+1, 0, -1, +1, -1, +1, 0, -1};
    static const int drow[8] = {-1, -1, -1, 0, 0, +1, +1, +1};
    static const int code[8] = {1, 2, 3, 4, 6, 7, 8, 9};
    for (int n = 0; n < t->ncell; n++) {
        int row = t->row[n], col = t->col[n];
        for (int k = 0; k < 8; k++) {
            int32_t m = G2(col + dcol[k], row + drow[k]);
            if (m != 0 && ldd[m - 1] == code[k]) inf[(size_t)n * 9 + (9 - code[k])] = m;
        }
    }
}

/* flow_accumulation, rout_prepare.cpp:442-494 (int16 arithmetic as in the reference) */
static void flow_accumulation(const wgo_topology *t, const int32_t *inf, int16_t *acc) {
    const int ng = t->ncell;
    for (int n = 0; n < ng; n++) acc[n] = 1;
    for (int n = 0; n < ng; n++)
        for (int i = 0; i <= 8; i++)
            if (inf[(size_t)n * 9 + i] != 0) acc[n] = 0;
    int cells_left = 99999, cells_left_prev;
    do {
        cells_left_prev = cells_left;
        cells_left = 0;
        for (int n = 0; n < ng; n++)
            if (acc[n] == 0) {
                int later = 0;
                for (int i = 0; i <= 8; i++) {
                    int32_t m = inf[(size_t)n * 9 + i];
                    if (m != 0 && acc[m - 1] == 0) later = 1;
                }
                if (!later) {
                    int16_t fa = 1;
                    for (int i = 0; i <= 8; i++) {
                        int32_t m = inf[(size_t)n * 9 + i];
                        if (m != 0) fa = (int16_t)(fa + acc[m - 1]);
                    }
                    acc[n] = fa;
                } else
                    cells_left++;
            }
    } while ((cells_left > 0) && (cells_left_prev - cells_left != 0));
}

int wgo_rout_prepare(wgo_topology *t) {
    const int ng = t->ncell;
    const int nr = t->nrow;
    int8_t *ldd = t->ldd;
    /* Arc flow direction -> LDD, rout_prepare.cpp:78-114 */
    for (int n = 0; n < ng; n++) {
        switch (t->flowdir[n]) {
            case -1: ldd[n] = -1; break;
            case 8: ldd[n] = 1; break;
            case 4: ldd[n] = 2; break;
            case 2: ldd[n] = 3; break;
            case 16: ldd[n] = 4; break;
            case 0: ldd[n] = 5; break;
            case 1: ldd[n] = 6; break;
            case 32: ldd[n] = 7; break;
            case 64: ldd[n] = 8; break;
            case 128: ldd[n] = 9; break;
            default: ldd[n] = 5; break;
        }
    }
    memcpy(t->ldd_2, ldd, (size_t)ng); /* G_LDD_2 is written before the loop check (:150) */
    int32_t *g2 = make_gcrc2(t);
    inflow_cells(t, g2, ldd, t->inflow9);
    flow_accumulation(t, t->inflow9, t->flow_acc);
    /* cells in loops keep accumulation 0 -> made sinks, :164-177 */
    int correction = 0;
    for (int n = 0; n < ng; n++)
        if (0 == t->flow_acc[n]) {
            ldd[n] = 5;
            t->flow_acc[n] = 1;
            correction = 1;
        }
    /* (G_FLOW_ACC.UNF2 is written here, :171; keep a copy semantics: the caller reads flow_acc
     *  after this function, which the reference recomputes below when corrected) */
    int16_t *acc_file = (int16_t *)malloc(sizeof(int16_t) * (size_t)ng);
    memcpy(acc_file, t->flow_acc, sizeof(int16_t) * (size_t)ng);
    if (correction) {
        inflow_cells(t, g2, ldd, t->inflow9);
        flow_accumulation(t, t->inflow9, t->flow_acc);
    }
    memcpy(t->flow_acc, acc_file, sizeof(int16_t) * (size_t)ng); /* file content = first pass */
    free(acc_file);

    /* derive_basins, :497-583 */
    int8_t *done = (int8_t *)malloc((size_t)ng);
    int i = 0, j = 0;
    for (int n = 0; n < ng; n++) { done[n] = -99; t->basins[n] = 0; t->cells_to_outlet[n] = 0; }
    for (int n = 0; n < ng; n++) {
        if ((0 == ldd[n]) || (-99 == ldd[n])) { done[n] = 2; t->basins[n] = 0; j++; ldd[n] = 5; }
        if ((5 == ldd[n]) || (-1 == ldd[n])) { done[n] = 1; i++; t->basins[n] = (uint16_t)i; }
    }
    t->nbasins = i;
    int k;
    i = 1;
    do {
        k = 0;
        for (int n = 0; n < ng; n++) {
            if (1 == done[n]) {
                t->cells_to_outlet[n] = (uint16_t)(i - 1);
                k++;
                done[n] = 2;
                for (int q = 0; q <= 8; q++) {
                    int32_t m = t->inflow9[(size_t)n * 9 + q];
                    if (m != 0) { t->basins[m - 1] = t->basins[n]; done[m - 1] = -1; }
                }
            }
        }
        for (int n = 0; n < ng; n++)
            if (-1 == done[n]) done[n] = 1;
        i++;
    } while (k != 0);
    free(done);

    /* reindex_waterbasins, :585-617 (basins with one cell -> 0, others renumbered 1..) */
    {
        int nb = t->nbasins;
        int32_t *cnt = (int32_t *)calloc((size_t)nb + 2, sizeof(int32_t));
        int32_t *newid = (int32_t *)calloc((size_t)nb + 2, sizeof(int32_t));
        for (int n = 0; n < ng; n++) cnt[t->basins[n]]++;
        int jj = 0;
        for (int b = 1; b <= nb; b++) {
            if (1 == cnt[b]) newid[b] = 0;
            else newid[b] = ++jj;
        }
        for (int n = 0; n < ng; n++) t->basins2[n] = (uint16_t)newid[t->basins[n]];
        t->nbasins2 = jj;
        free(cnt);
        free(newid);
    }

    /* outflow_cell, :619-670 */
    {
        static const int dcol[10] = {0, -1, 0, +1, -1, 0, +1, -1, 0, +1};
        static const int drow[10] = {0, +1, +1, +1, 0, 0, 0, -1, -1, -1};
        for (int n = 0; n < ng; n++) {
            int row = t->row[n], col = t->col[n];
            int d = ldd[n];
            if (d == 5 || d == -1) t->outflow_cell[n] = 0;
            else if (d >= 1 && d <= 9) t->outflow_cell[n] = G2(col + dcol[d], row + drow[d]);
            else t->outflow_cell[n] = 0;
        }
    }
    /* neighbouring_cells, :672-698 */
    if (t->neighbour8) {
        static const int dcol[8] = {+1, +1, 0, -1, -1, -1, 0, +1};
        static const int drow[8] = {0, -1, -1, -1, 0, +1, +1, +1};
        for (int n = 0; n < ng; n++)
            for (int q = 0; q < 8; q++) t->neighbour8[(size_t)n * 8 + q] = G2(t->col[n] + dcol[q], t->row[n] + drow[q]);
    }
    free(g2);

    /* rout_order, :834-886: Kahn sweeps in ascending cell number */
    {
        int32_t *routing = (int32_t *)calloc((size_t)ng, sizeof(int32_t));
        int8_t *list = (int8_t *)calloc((size_t)ng, 1);
        for (int n = 0; n < ng; n++) t->rout_order[n] = -99;
        for (int n = 0; n < ng; n++)
            for (int q = 0; q < 9; q++)
                if (t->inflow9[(size_t)n * 9 + q] > 0) routing[n]++;
        int routOrder = 1, counterStep, levels = 0;
        do {
            counterStep = 0;
            for (int n = 0; n < ng; n++) {
                if (routing[n] == 0) { t->rout_order[n] = routOrder++; counterStep++; list[n] = 1; }
                else list[n] = 0;
            }
            for (int n = 0; n < ng; n++)
                if (list[n]) {
                    routing[n]--;
                    if (t->outflow_cell[n] > 0) routing[t->outflow_cell[n] - 1]--;
                }
            if (counterStep > 0) levels++;
        } while (counterStep > 0);
        t->nlevels = levels;
        int bad = 0;
        for (int n = 0; n < ng; n++)
            if (routing[n] >= 0 || t->rout_order[n] == -99) bad = 1;
        free(routing);
        free(list);
        if (bad) return -1;
    }
    return 0;
}
#undef G1
#undef G2

/* calculate_distances, rout_prepare.cpp:700-759 (float arithmetic with a float pi) */
void wgo_cell_distances(int nrow, float *cd) {
#define CD(ch, i) cd[(size_t)(ch) * nrow + (i)] /* file layout [9][nrow], GCELLDIST.9.UNF0 */
    const float pi = 3.141592653589793;
    const float earth_radius = 6371.211;
    float vert_dist, horiz_dist, l;
    vert_dist = 2 * pi * earth_radius * 0.5 / 360.0;
    for (int i = 0; i < nrow; i++) {
        l = 0.25 + i * 0.5;
        horiz_dist = 2 * pi * earth_radius * sin(l * pi / 180.0) * 0.5 / 360.0;
        CD(1, i) = vert_dist;
        CD(3, i) = horiz_dist;
        CD(4, i) = 0;
        CD(5, i) = horiz_dist;
        CD(7, i) = vert_dist;
    }
    for (int i = 0; i < nrow - 1; i++) CD(0, i) = sqrt(CD(3, i) * CD(3, i + 1) + CD(1, i) * CD(1, i));
    CD(0, nrow - 1) = -99;
    for (int i = 1; i < nrow; i++) CD(6, i) = sqrt(CD(3, i) * CD(3, i - 1) + CD(1, i) * CD(1, i));
    CD(6, 0) = -99;
    for (int i = 0; i < nrow; i++) {
        CD(2, i) = CD(0, i);
        CD(8, i) = CD(6, i);
    }
}

/* calculate_river_slope (:761-794) and calculate_river_length (:796-832) */
void wgo_river_slope_length(int ncell, int nrow, const float *cd, const float *altitude, const float *meandering,
                            const int32_t *outflow_cell, const int8_t *ldd, const int16_t *row, const int16_t *col,
                            float *slope, float *length) {
    for (int n = 0; n < ncell; n++) {
        if ((outflow_cell[n] != 0) && ((ldd[n] != 5) && (ldd[n] != -1))) {
            slope[n] = (altitude[n] - altitude[outflow_cell[n] - 1])
                       / (1000.0 * CD(ldd[n] - 1, row[n] - 1) * ((meandering[n] + meandering[outflow_cell[n] - 1]) / 2));
        } else
            slope[n] = 0;
    }
    float minslope = 0.00001;
    for (int n = 0; n < ncell; n++)
        if (slope[n] < minslope) slope[n] = minslope;
    for (int n = 0; n < ncell; n++) {
        int n_outflow = outflow_cell[n] - 1;
        if (n_outflow == -1) length[n] = 55.;
        else if (row[n] == row[n_outflow]) length[n] = CD(3, row[n] - 1);
        else {
            if (row[n] > row[n_outflow]) {
                if (col[n] == col[n_outflow]) length[n] = CD(7, row[n] - 1);
                else length[n] = CD(8, row[n] - 1);
            } else {
                if (col[n] == col[n_outflow]) length[n] = CD(1, row[n] - 1);
                else length[n] = CD(2, row[n] - 1);
            }
        }
        if (n_outflow == -1) length[n] *= meandering[n];
        else length[n] *= ((meandering[n] + meandering[n_outflow]) / 2);
    }
}
#undef CD

/* reservoir_prepare, rout_prepare.cpp:888-1031.  G_RESAREA.UNF0 (float) is read into a
 * Grid<int> there: only the sign of the bit pattern is used, which for finite floats equals
 * the sign of the value. */
void wgo_reservoir_prepare(int ncell, const float *resarea_f32, const float *mean_outflow_f32,
                           const float *mean_outflow12_f32, const int32_t *outflow_cell, float *alloc_coeff5,
                           int8_t *start_month) {
    double *alloc = (double *)malloc(sizeof(double) * 5 * (size_t)ncell);
    double *inflowRes = (double *)calloc((size_t)ncell, sizeof(double));
    int32_t *ra = (int32_t *)malloc(sizeof(int32_t) * (size_t)ncell);
    memcpy(ra, resarea_f32, sizeof(int32_t) * (size_t)ncell);
    for (int n = 0; n < ncell; n++)
        for (int i = 0; i < 5; i++) alloc[(size_t)n * 5 + i] = 1.0;
    for (int n = 0; n < ncell; n++) {
        if (ra[n] > 0) {
            short i = 0;
            int d = outflow_cell[n];
            while (i < 5 && d > 0 && ra[d - 1] <= 0) {
                inflowRes[d - 1] += (double)mean_outflow_f32[n];
                d = outflow_cell[d - 1];
                i++;
            }
        }
    }
    for (int n = 0; n < ncell; n++) {
        if (ra[n] > 0) {
            short i = 0;
            int d = outflow_cell[n];
            while (i < 5 && d > 0 && ra[d - 1] <= 0) {
                alloc[(size_t)n * 5 + i] = (double)mean_outflow_f32[n] / inflowRes[d - 1];
                d = outflow_cell[d - 1];
                i++;
            }
        }
    }
    for (size_t q = 0; q < (size_t)ncell * 5; q++) alloc_coeff5[q] = (float)alloc[q];
    /* first month of the operational year = first month after the longest "dry" run
     * (monthly outflow below the annual mean), :971-996 */
    for (int n = 0; n < ncell; n++) {
        int start = 0, dry_length = 0, start_ = 0, counter_length = 0, last_dry = 1;
        int month = 0;
        double mo = (double)mean_outflow_f32[n];
        while (month < 12 && (double)mean_outflow12_f32[(size_t)n * 12 + month] < mo) month++;
        int beg_month = month;
        for (int m = 0; m < 12; m++) {
            month = m + beg_month;
            if (month > 11) month -= 12;
            if ((double)mean_outflow12_f32[(size_t)n * 12 + month] < mo) {
                counter_length++;
                if (!last_dry) { start_ = month; last_dry = 1; }
            } else if (last_dry) {
                last_dry = 0;
                if (counter_length > dry_length) { start = start_; dry_length = counter_length; counter_length = 0; }
            }
        }
        if (counter_length > dry_length) { start = start_; dry_length = counter_length; }
        start_month[n] = (int8_t)(start + 1);
    }
    free(alloc);
    free(inflowRes);
    free(ra);
}
