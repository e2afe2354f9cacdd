// wg_calibration.h - host-side look-alike of the reference's calibGammaClass (calibration.h / calibration.cpp:45-801):
// the search for the runoff coefficient gamma of one calibration basin, driven by integrate_wghm_
// (integrateWGHM.cpp:213-264, 967-972, 1091-1116).  Same method names and argument meaning; the model runs between two
// findNewGamma() calls are the caller's (GPU runs through the C ABI; watergap2_b200/calibration.py::calibrate_gamma is that loop).
//
// Everything is single precision in the reference's operand order.  Files: CALIBRATION.OUT and STAT_CORR_FACTOR.OUT
// get the reference's data lines, CALIBSTATUS.OUT its status; the prose of CALIBRATION.LOG is reduced to one line per call.
// Pinned against the compiled reference on seven scenarios (tests/golden/ref_calibration.json, tests/test_host_library.py);
// the same logic exists as Python host code in watergap2_b200/calibration.py (GammaCalibration).
#pragma once
#include <string>
#include <vector>

class calibGammaClass {
public:
    calibGammaClass();
    int cellCorrFactorInd = 1;
    int gammaCond = 0;
    int calibStatus = 0;  // 1 gamma found, 2 within 10 % at a limit, 3 CFA computed, 4 CFS needed
    short CallCounterNo = 0;

    // evaluation period and station (the reference reads them from the global options / C_STATION.DAT)
    void configure(short evalStartYear, short endYear, short calibStationNumber, const std::string &directory);
    void prepareFiles();                                   // truncates CALIBRATION.OUT / .LOG, CALIBSTATUS.OUT
    void init();                                           // resets the yearly arrays and reads RIVER.DAT if present
    void readObservedData();                               // RIVER.DAT: year, m3/s -> km3/year
    void setObserved(int year, float m3_per_s);            // the same for one year, without the file
    void setRunoff(int year, float value);
    void setUpstInflow(int year, float value);
    void setWaterUse(int year, float value);
    float findNewGamma(float gamma_old);                   // -99: the search has ended
    void writeCorrFactors(float gamma, int cellCorrFactInd);
    void writeCalibStatus(int status);
    float getCellCorrFactor() const { return cellCorrFactor; }
    short getCallCounter() const { return CallCounterNo; }
    int getCalibStatus() const { return calibStatus; }
    float getStationCorrFactor() const { return stationCorrFactor; }
    // createCorrectionGrid: annual G_POT_CELL_RUNOFF grids [nyears][ng] (float), basin ids [ng]; updates cellCorrFact [ng]
    void createCorrectionGrid(const std::vector<std::vector<float>> &annualPotCellRunoff, const short *sbasin, int ng,
                              float simulatedDischarge, float measuredDischarge, double *cellCorrFact) const;
    bool correctionGridDue() const { return gridDue; }
    float gridSimulated() const { return gridSim; }
    float gridMeasured() const { return gridMeas; }
    const std::string &lastResultLine() const { return lastLine; }
    const std::string &lastCorrFactorLine() const { return lastCorrLine; }

private:
    std::string dir;
    short evalStart = 0, endYear = 0, calibStationNumber = 0;
    std::vector<float> measuredRunoff, simulatedRunoff, simulatedInflow, simulatedWaterUse;
    float gammaUpperLimit, gammaLowerLimit;
    float gammaOfPreviousRun = 0.f, cellCorrFactor = 1.f, stationCorrFactor = 1.f;
    // function-local statics of the reference's findNewGamma
    short callCounter = 0;
    float gamma_low = -99.f, gamma_high = -99.f, sumOfDifferences_old = 0.f;
    bool gridDue = false;
    float gridSim = 0.f, gridMeas = 0.f;
    std::string lastLine, lastCorrLine;
    std::string path(const char *name) const;
};
