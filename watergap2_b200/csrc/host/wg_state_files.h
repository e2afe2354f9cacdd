// wg_state_files.h — checkpoint containers and their text codecs (drop-in file formats).
//   Cell / WghmStateFile          wghmStateFile.h:7-73, txt format wghmStateFile.cpp:160-229, 540-599
//   SnowInElevationFile           snowInElevationFile.h:9-28, txt format snowInElevationFile.cpp:29-131, 205-248
//   AdditionalOutputInputFile     additionalOutputInputFile.h:14-32, txt format .cpp:131-234, 237-406
// Numbers are written with precision(16) scientific (17 significant digits: lossless doubles).
// The NetCDF variants of the reference (.nc) are out of scope (netcdf-c is not available).
#pragma once
#include <array>
#include <fstream>
#include <iomanip>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace wg {

class Cell {
  public:
    Cell() = default;
    Cell(size_t id, size_t l) : id_(id) { resize(l); }
    void resize(size_t l) { for (auto &v : c_) v.resize(l); }
    size_t size() const { return c_[0].size(); }
    int id() const { return (int)id_; }
    double &canopy(int i) { return c_[0].at(i); }
    double &snow(int i) { return c_[1].at(i); }
    double &soil(int i) { return c_[2].at(i); }
    double &locallake(int i) { return c_[3].at(i); }
    double &localwetland(int i) { return c_[4].at(i); }
    double &globallake(int i) { return c_[5].at(i); }
    double &globalwetland(int i) { return c_[6].at(i); }
    double &reservoir(int i) { return c_[7].at(i); }
    double &river(int i) { return c_[8].at(i); }
    double &groundwater(int i) { return c_[9].at(i); }
    double &compartment(int k, int i) { return c_[k].at(i); }
    double tws(int i) {  // wghmStateFile.cpp:730-735 (same summation order)
        return canopy(i) + snow(i) + soil(i) + locallake(i) + localwetland(i) + globallake(i) + globalwetland(i) + reservoir(i) + river(i) + groundwater(i);
    }
    Cell mean() const {  // wghmStateFile.cpp:711-727
        Cell m(id_, 1);
        for (int k = 0; k < 10; k++) m.c_[k][0] = std::accumulate(c_[k].begin(), c_[k].end(), 0.0) / (double)c_[k].size();
        return m;
    }

  private:
    size_t id_ = 0;
    std::array<std::vector<double>, 10> c_;
};

class WghmStateFile {
  public:
    explicit WghmStateFile(size_t ncell = 0, size_t len = 0) {
        cells_.resize(ncell);
        for (size_t i = 0; i < ncell; i++) cells_[i] = Cell(i + 1, len);
    }
    size_t size() const { return cells_.size(); }
    Cell &cell(int i) { return cells_.at(i); }
    void resetCells(size_t l = 0) { for (auto &c : cells_) { c.resize(0); c.resize(l); } }
    void load(const std::string &fn) {
        if (ext(fn) != "txt") return;  // like the reference: unknown extensions are silently ignored
        std::ifstream f(fn);
        if (!f) throw std::runtime_error("In WghmStateFile::load(): error by opening file " + fn);
        std::string line;
        while (std::getline(f, line)) {
            if (line.empty()) break;
            std::stringstream ss(line);
            char c;
            ss >> c;
            ss.putback(c);
            if (isalpha((unsigned char)c)) continue;
            int id;
            double tws;
            ss >> id >> tws;
            Cell cell(id, 1);
            for (int k = 0; k < 10; k++) ss >> cell.compartment(k, 0);
            if ((size_t)id > cells_.size()) cells_.resize(id);
            cells_.at(id - 1) = cell;
        }
    }
    void saveDay(const std::string &fn, int day) { write(fn, day, false); }
    void saveMean(const std::string &fn) { write(fn, 0, true); }

  private:
    static std::string ext(const std::string &fn) { return fn.substr(fn.find_last_of(".") + 1); }
    void write(const std::string &fn, int day, bool mean) {
        if (ext(fn) != "txt") throw std::runtime_error("WghmStateFile: only .txt is supported: " + fn);
        std::ofstream s(fn, std::ios::binary);
        if (!s) throw std::runtime_error("In WghmStateFile::saveDay(): error by opening file " + fn);
        const unsigned width = 33;
        s << "WGHM water storage states" << std::endl;
        s << std::setw(6) << std::setfill(' ') << "ID";
        for (const char *h : {"TWS", "CANOPY", "SNOW", "SOIL", "LOCALLAKE", "LOCALWETLAND", "GLOBALLAKE", "GLOBALWETLAND", "RESERVOIR", "RIVER", "GROUNDWATER"})
            s << std::setw(width) << std::setfill(' ') << h;
        s << std::endl;
        for (auto &c0 : cells_) {
            Cell c = mean ? c0.mean() : c0;
            const int d = mean ? 0 : day;
            s.precision(16);
            s << std::setw(6) << std::setfill(' ') << c.id();
            s << std::setw(width) << std::scientific << std::setfill(' ') << c.tws(d);
            for (int k = 0; k < 10; k++) s << std::setw(width) << std::scientific << std::setfill(' ') << c.compartment(k, d);
            s << std::endl;
        }
    }
    std::vector<Cell> cells_;
};

class SnowInElevationFile {
  public:
    explicit SnowInElevationFile(size_t ncell = 0) : ncell_(ncell), v_(ncell * 101, 0.0) {}
    double &snowInElevation(int cell, int elev) { return v_[(size_t)cell * 101 + elev]; }
    double *data() { return v_.data(); }
    void save(const std::string &fn) {
        std::ofstream s(fn, std::ios::binary);
        if (!s) throw std::runtime_error("In SnowInElevationFile::save(): error by opening file " + fn);
        s << "Snow in elevation for land cells" << std::endl;
        s.precision(16);
        for (size_t i = 0; i < ncell_; i++) {
            s << i + 1 << "\t";
            for (int j = 0; j < 100; j++) s << std::scientific << v_[i * 101 + j] << "\t";
            s << std::scientific << v_[i * 101 + 100];
            if (i + 1 < ncell_) s << std::endl;
        }
    }
    void load(const std::string &fn) {
        std::ifstream f(fn);
        if (!f) throw std::runtime_error("In SnowInElevationFile::load(): error by opening file " + fn);
        std::string line;
        std::getline(f, line);  // header
        for (size_t i = 0; i < ncell_; i++) {
            size_t id;
            f >> id;
            for (int j = 0; j < 101; j++) f >> v_[(id - 1) * 101 + j];
        }
    }

  private:
    size_t ncell_;
    std::vector<double> v_;
};

class AdditionalOutputInputFile {
  public:
    explicit AdditionalOutputInputFile(size_t ncell = 0) : ncell_(ncell), v_(ncell * 53, 0.0) {
        for (size_t i = 0; i < ncell; i++) {  // start values of additionalOutputInputFile.cpp:72-84: K_release 0.1, reduction factors 1
            v_[i * 53 + 5] = 0.1;
            for (int j : {8, 9, 11, 12, 13}) v_[i * 53 + j] = 1.;
        }
    }
    int additionalfilestatus = 0;
    double &additionalOutputInput(int i, int j) { return v_[(size_t)i * 53 + j]; }
    void save(const std::string &fn) {
        std::ofstream s(fn, std::ios::binary);
        if (!s) throw std::runtime_error("In AdditionalOutputInputFile::save(): error by opening file " + fn);
        const unsigned width = 33;
        // title and column names are part of the file format (additionalOutputInputFile.cpp:327-383), kept byte for
        // byte - including the leading blanks and the spelling - so that a file written here equals the reference's
        static const char *const kColumns[53] = {
            "G_days_since_start", "G_GrowingStatus", "G_PrecSum", "G_totalUnsatisfiedUse", "G_UnsatisfiedUsePrevYear", "K_release",
            "G_landAreaFrac", "G_landAreaFracPrevTimestep", "G_locWetlAreaReductionFactor", "G_gloLakeEvapoReductionFactor",
            "G_groundwaterStorage", "G_locLakeAreaReductionFactor", "G_gloWetlAreaReductionFactor", "G_gloResEvapoReductionFactor",
            "G_fswbInit[n]", "G_gloWetlStorage", "G_fswbLandAreaFrac", "G_locWetlStorage[n]", "G_locLakeStorage[n]",
            "G_riverStorage[n]", "G_soilwatercontent[n]", "G_gloLakeStorage[n]", "G_canopywatercontent[n]", "G_gloResStorage[n]",
            "G_snow[n]", "G_withdrawalIrrigFromSwb[n]", "G_consumptiveUseIrrigFromSwb[n]", "G_unsatAllocUSE", "G_AllocUSE",
            "G_secondcell", "G_daily_UnsatAllocUseNextDay[n]", " G_daily_allocatedUseNextDay[n]", " G_AllocUSETOneigborcell[n]",
            " G_fswbLandAreaFracNextTimestep[n]", " G_PrevUnsatAllocUse[n]", " G_dailyRemainingUse", " G_dailyAllocatedUse[n]",
            " G_PrevTotalUnstatisfieduse[n]", " G_reducedReturnFlow[n]", " G_unsatisfiedNAsFromIrrig[n]",
            " G_dailySatisAllocatedUseInSecondCell", " G_dailyAllocatedUse", " G_unsatUseRiparian[n]", " G_ActualUse[secondCell]",
            " G_glores_prevyear[n]", " G_fLocLake[n]", " G_fGloWet[n]", " G_fLocWet[n]", " G_unsatUseRiparian[i]",
            " G_unsatisfiedNAsFromOtherSectors[n]", " G_reducedReturnFlowPrevYear[n]", " G_unsatisfiedNAsFromIrrigPrevYear[n]",
            " G_unsatisfiedNAsFromOtherSectorsPrevYear[n]"};
        s << "additional out- and input" << std::endl;
        s << std::setw(6) << std::setfill(' ') << "ID";
        for (int j = 0; j < 53; j++) s << std::setw(width) << std::setfill(' ') << kColumns[j];
        s << std::endl;
        s.precision(16);
        for (size_t i = 0; i < ncell_; i++) {
            s << std::setw(6) << std::setfill(' ') << i + 1;
            for (int j = 0; j < 53; j++) s << std::setw(width) << std::setfill(' ') << std::scientific << v_[i * 53 + j];
            if (i + 1 < ncell_) s << std::endl;
        }
    }
    void load(const std::string &fn) {
        std::ifstream f(fn);
        if (!f) throw std::runtime_error("In AdditionalOutputInputFile::load(): error by opening file " + fn);
        std::string line;
        std::getline(f, line);
        std::getline(f, line);
        for (size_t i = 0; i < ncell_; i++) {
            size_t id;
            f >> id;
            for (int j = 0; j < 53; j++) f >> v_[(id - 1) * 53 + j];
        }
        additionalfilestatus = 1;
    }

  private:
    size_t ncell_;
    std::vector<double> v_;
};

}  // namespace wg
