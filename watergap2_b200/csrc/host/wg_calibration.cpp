// wg_calibration.cpp - see wg_calibration.h.  Written from the description of the procedure (decision table below), single
// precision throughout; reference lines are cited per block.
#include "wg_calibration.h"

#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>

calibGammaClass::calibGammaClass() : gammaUpperLimit(5.f), gammaLowerLimit(0.1f) {}  // calibration.cpp:45-49

std::string calibGammaClass::path(const char *name) const { return dir.empty() ? std::string(name) : dir + "/" + name; }

void calibGammaClass::configure(short evalStartYear, short end_year, short station, const std::string &directory) {
    evalStart = evalStartYear;
    endYear = end_year;
    calibStationNumber = station;
    dir = directory;
}

void calibGammaClass::prepareFiles() {  // :51-84 (the time stamp comment line is not written)
    for (const char *n : {"CALIBRATION.OUT", "CALIBRATION.LOG", "CALIBSTATUS.OUT"}) std::ofstream(path(n), std::ios::trunc);
}

void calibGammaClass::init() {  // :239-249
    const int n = endYear - evalStart + 1;
    measuredRunoff.assign(n, -99.f);
    simulatedRunoff.assign(n, -99.f);
    simulatedInflow.assign(n, -99.f);
    simulatedWaterUse.assign(n, -99.f);
    readObservedData();
}

void calibGammaClass::setObserved(int year, float q) {  // m3/s -> km3/year, :233-235 (float * int ... / double, stored as float)
    measuredRunoff.at(year - evalStart) = (float)((double)(q * 365 * 24 * 60 * 60) / 1000000000.0);
}

void calibGammaClass::readObservedData() {  // :212-237
    std::ifstream in(path("RIVER.DAT"));
    int year;
    float q;
    while (in >> year >> q)
        if (year >= evalStart && year <= endYear) setObserved(year, q);
}

void calibGammaClass::setRunoff(int year, float v) { simulatedRunoff.at(year - evalStart) = v; }
void calibGammaClass::setUpstInflow(int year, float v) { simulatedInflow.at(year - evalStart) = v; }
void calibGammaClass::setWaterUse(int year, float v) { simulatedWaterUse.at(year - evalStart) = v; }

// Decision table of the search (:315-526).  sod = sum over the observed years of (simulated - observed).
//   |sod / (n * mean observed)| < 1 %                    -> done (gamma = -99, gammaCond 1, status 1)
//   first call:  sod > 0 (too much runoff, gamma too small) -> upper limit (or done when already there); else lower limit
//   later calls: bisect between the bracketing values; while one side of the bracket is unknown, double / halve;
//                at a limit: stay there once (second call), then give up (gammaCond 2) and let the 10 % rule / CFA decide
float calibGammaClass::findNewGamma(float gamma_old) {
    callCounter++;
    CallCounterNo = callCounter;
    const float UP = gammaUpperLimit, LO = gammaLowerLimit;
    float sod = 0, msum = 0, isum = 0, usum = 0, rsum = 0;
    int n = 0;
    for (size_t i = 0; i < measuredRunoff.size(); i++)  // :300-313
        if (measuredRunoff[i] > -1) {
            n++;
            sod += simulatedRunoff[i] - measuredRunoff[i];
            msum += measuredRunoff[i];
            rsum += simulatedRunoff[i];
            isum += simulatedInflow[i];
            usum += simulatedWaterUse[i];
        }
    const float avg = msum / n;
    float gamma = NAN;  // the reference leaves it unassigned on the paths that cannot occur
    auto half = [&]() { return (float)((gamma_low + gamma_high) / 2.0); };
    if (std::fabs(sod / (n * avg)) < 0.01) {
        gamma = -99;
        gammaCond = 1;
        calibStatus = 1;
    } else if (callCounter == 1) {
        if (sod > 0) {
            if (gamma_old >= UP) gamma = -99;
            else { gamma_low = gamma_old; gamma = UP; }
        } else {
            if (gamma_old <= LO) gamma = -99;
            else { gamma_high = gamma_old; gamma = LO; }
        }
    } else if (sod > 0) {
        if (gamma_old == UP) {
            if (callCounter == 2) { gamma = gamma_old; gamma_high = UP; }
            else { gamma = -99; gammaCond = 2; }
        } else if (gamma_old < UP) {
            gamma_low = gamma_old;
            gamma = gamma_high < 0 ? (float)(gamma_old * 2.0) : half();
        }
    } else {
        if (gamma_old == LO) {
            if (callCounter == 2) { gamma = gamma_old; gamma_low = LO; }
            else { gamma = -99; gammaCond = 2; }
        } else if (gamma_old > LO) {
            const bool open = gamma_low < 0;
            gamma_high = gamma_old;
            gamma = open ? (float)(gamma_old / 2.0) : half();
        }
    }
    float nse = -99;  // Nash-Sutcliffe, only with more than one year (:529-546)
    if (endYear - evalStart > 0) {
        float s1 = 0, s2 = 0;
        for (size_t i = 0; i < measuredRunoff.size(); i++)
            if (measuredRunoff[i] > -1) {
                s1 += (measuredRunoff[i] - avg) * (measuredRunoff[i] - avg);
                s2 += (simulatedRunoff[i] - measuredRunoff[i]) * (simulatedRunoff[i] - measuredRunoff[i]);
            }
        nse = (s1 - s2) / s1;
    }
    // gamma stuck at a limit: accept within 10 % of the observation, otherwise a cell correction factor (:551-592)
    float avgAdapt = avg;
    const float simAvg = rsum / n;
    if (gamma_old == UP && gammaCond == 2) {
        avgAdapt = (float)(avg * 1.1);
        if (simAvg < avgAdapt) calibStatus = 2;
        else cellCorrFactorInd = 99;
    }
    if (gamma_old == LO && gammaCond == 2) {
        avgAdapt = (float)(avg * 0.9);
        if (simAvg > avgAdapt) calibStatus = 2;
        else cellCorrFactorInd = 99;
    }
    std::string generated = "?";  // uninitialised in the reference until CFA is computed
    if (cellCorrFactorInd == 99) {
        cellCorrFactor = (avgAdapt + (usum - isum) / n) / ((rsum + usum - isum) / n);
        const float g = (rsum / n) + (usum / n) - (isum / n);
        std::ostringstream o;
        o << g;
        generated = o.str();
    }
    {   // CALIBRATION.OUT (:594-607)
        std::ostringstream o;
        o << gamma_old << '\t' << nse << '\t' << sod << '\t' << gamma_low << '\t' << gamma_high << '\t' << avgAdapt << '\t' << n << '\t'
          << (avgAdapt / simAvg) << '\t' << isum / n << '\t' << usum / n << '\t' << 0 << '\t' << simAvg << '\t' << generated << '\t'
          << cellCorrFactor << '\t';
        lastLine = o.str();
        std::ofstream(path("CALIBRATION.OUT"), std::ios::app) << lastLine << std::endl;
        std::ofstream(path("CALIBRATION.LOG"), std::ios::app) << "call " << callCounter << ": gamma " << gamma_old << " -> " << gamma
                                                             << ", 1% criterion " << std::fabs(sod / (n * avg)) << std::endl;
    }
    if (gamma > 0 && gamma < 0.99 * LO && callCounter > 2) gamma = -99;                                   // :609-613
    if (gamma > 0 && std::fabs((sumOfDifferences_old / sod) - 1) < 0.0001) gamma = -99;                 // :614-619
    sumOfDifferences_old = sod;
    gridDue = false;
    if (gamma < 0 && std::fabs(cellCorrFactor - 1) > 0.01) {                                            // :623-627
        gridDue = true;
        gridSim = simAvg;
        gridMeas = avgAdapt;
        calibStatus = 3;
    }
    gammaOfPreviousRun = gamma_old;
    return gamma;
}

void calibGammaClass::createCorrectionGrid(const std::vector<std::vector<float>> &annual, const short *sbasin, int ng,
                                           float simulatedDischarge, float measuredDischarge, double *cellCorrFact) const {  // :730-801
    std::vector<float> mean(ng, 0.f);
    for (const auto &y : annual)
        for (int c = 0; c < ng; c++) mean[c] += y[c];
    const short years = (short)annual.size();
    for (int c = 0; c < ng; c++) mean[c] /= years;
    float total = 0;
    for (int c = 0; c < ng; c++)
        if (sbasin[c] == calibStationNumber) total += std::fabs(mean[c]);
    for (int c = 0; c < ng; c++) {
        if (sbasin[c] != calibStationNumber) continue;
        const int sg = mean[c] > 0 ? 1 : (mean[c] == 0 ? 0 : -1);
        double f = 1 - ((sg * (simulatedDischarge - measuredDischarge)) / total);
        if (f > 1.5) f = 1.5;  // the correction factor is limited to 0.5 .. 1.5
        if (f < 0.5) f = 0.5;
        cellCorrFact[c] = f;
    }
}

void calibGammaClass::writeCorrFactors(float gamma, int cellCorrFactInd) {  // :660-728
    float msum = 0, rsum = 0;
    int n = 0;
    for (size_t i = 0; i < measuredRunoff.size(); i++)
        if (measuredRunoff[i] > -1) {
            n++;
            msum += measuredRunoff[i];
            rsum += simulatedRunoff[i];
        }
    if (gamma == gammaUpperLimit && cellCorrFactInd == 99) msum = (float)(msum * 1.1);
    if (gamma == gammaLowerLimit && cellCorrFactInd == 99) msum = (float)(msum * 0.9);
    float cfs = cellCorrFactInd == 99 ? msum / rsum : 1.0f;
    if ((cfs > 1.0 && cfs < 1.01) || (cfs < 1.0 && cfs > 0.99)) cfs = 1.0f;  // within 1 %: no station correction
    if (cfs > 1.0 || cfs < 1.0) calibStatus = 4;
    stationCorrFactor = cfs;
    std::ostringstream o;
    o << msum / n << '\t' << rsum / n << '\t' << gammaOfPreviousRun << '\t' << cellCorrFactor << '\t' << cfs;
    lastCorrLine = o.str();
    std::ofstream f(path("STAT_CORR_FACTOR.OUT"));
    f << "# Measured runoff\tSimulated runoff\tGamma\tCell correction factor\tStation correction factor" << std::endl << lastCorrLine << std::endl;
}

void calibGammaClass::writeCalibStatus(int status) { std::ofstream(path("CALIBSTATUS.OUT"), std::ios::app) << status << std::endl; }  // :653-659

// ---- C entry points for tests and other languages ---------------------------------------------------------------
extern "C" {
void *wg_calib_create(short evalStartYear, short endYear, short station, const char *directory) {
    auto *c = new calibGammaClass();
    c->configure(evalStartYear, endYear, station, directory ? directory : "");
    c->prepareFiles();
    c->init();
    return c;
}
void wg_calib_destroy(void *h) { delete (calibGammaClass *)h; }
void wg_calib_set_observed(void *h, int year, float m3s) { ((calibGammaClass *)h)->setObserved(year, m3s); }
void wg_calib_set_year(void *h, int year, float runoff, float waterUse, float upstInflow) {
    auto *c = (calibGammaClass *)h;
    c->setRunoff(year, runoff);
    c->setWaterUse(year, waterUse);
    c->setUpstInflow(year, upstInflow);
}
// out[6] = {gamma, callCounter, gammaCond, calibStatus, cellCorrFactor, cellCorrFactorInd}; line = the CALIBRATION.OUT data line
void wg_calib_find_new_gamma(void *h, float gamma_old, double out[6], char *line, size_t linelen) {
    auto *c = (calibGammaClass *)h;
    const float g = c->findNewGamma(gamma_old);
    out[0] = g; out[1] = c->getCallCounter(); out[2] = c->gammaCond; out[3] = c->getCalibStatus();
    out[4] = c->getCellCorrFactor(); out[5] = c->cellCorrFactorInd;
    if (line && linelen) snprintf(line, linelen, "%s", c->lastResultLine().c_str());
}
// -> calibStatus after the test run; line = the STAT_CORR_FACTOR.OUT data line
int wg_calib_finish(void *h, float gamma, char *line, size_t linelen) {
    auto *c = (calibGammaClass *)h;
    c->writeCorrFactors(gamma, c->cellCorrFactorInd);
    c->writeCalibStatus(c->getCalibStatus());
    if (line && linelen) snprintf(line, linelen, "%s", c->lastCorrFactorLine().c_str());
    return c->getCalibStatus();
}
// 1 when findNewGamma asked for the correction grid; applies it to cellCorrFact [ng]
int wg_calib_correction_grid(void *h, const float *annual, int nyears, const short *sbasin, int ng, double *cellCorrFact) {
    auto *c = (calibGammaClass *)h;
    if (!c->correctionGridDue()) return 0;
    std::vector<std::vector<float>> a(nyears);
    for (int y = 0; y < nyears; y++) a[y].assign(annual + (size_t)y * ng, annual + (size_t)(y + 1) * ng);
    c->createCorrectionGrid(a, sbasin, ng, c->gridSimulated(), c->gridMeasured(), cellCorrFact);
    return 1;
}
}
