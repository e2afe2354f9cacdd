// wg_pdaf_bridge.h - host-side look-alikes of the state exchange of the reference's PDAF coupling, over the host
// WghmStateFile / SnowInElevationFile classes (wg_state_files.h):
//   extract_sub      extractsub.cpp:17-79    monthly-mean state vector of the region's cells minus the temporal mean field
//   enkf_wghmstate   enKF2wghmState.cpp:17-121, 440-529 (state part): last day += analysis - prediction with the storage
//                    limits, snow in elevation rescaled to the analysed snow, monthly mean after the assimilation
//   extract_sub_parameters  extractsub.cpp:81-340   sub-basin means of the calibrated parameters, appended to the state vector
//   enkf_parameters         enKF2wghmState.cpp:127-399  analysed parameters clamped to their range, sub-basin means otherwise
//   parameterJsonFile       parameterJsonFile.cpp:65-163, 207-525, 563-596  the per-cell parameter file of the next cycle
// The device-side equivalents (no WghmStateFile round trip) are wgk_state_vector / wgk_enkf_update (include/wgk.h).
// Equal to the compiled reference's functions on tests/golden/ref_ng1000_enkf.npz and ref_ng1000_enkf_par.npz
// (tests/test_host_library.py).
#pragma once
#include <string>
#include <vector>

#include "wg_model.h"
#include "wg_state_files.h"

namespace wg {
// ids: 1-based cell numbers of the region (the reference reads them from the GIM header file); output [ids.size() * 10]
void extract_sub(const std::vector<int> &ids, WghmStateFile &wghmState, WghmStateFile &wghmMean, double *output);
// wghmState holds the days of the month on entry and only the updated last day on return; snow [cell][101];
// wghmStateMean (out, one day): the monthly mean after the assimilation
void enkf_wghmstate(const std::vector<int> &ids, const double *field, const double *prediction, WghmStateFile &wghmState,
                    SnowInElevationFile &snow_in_elevation, WghmStateFile &wghmMean, WghmStateFile &wghmStateMean);

// ---- parameter half.  nunit calibration units ("sub-basins"); calPar_index [nunit][26] (1: the parameter of that unit is part of
// the assimilated vector); groupmatrixindex [nunit][ids.size()]: 0, or the 1-based cell number of region cell j when it belongs to
// unit i (the Fortran side hands both over column-major, which is this row-major shape).
// -> out: one value per calPar_index == 1 in (unit, parameter) order, the mean of the parameter over the unit's cells
void extract_sub_parameters(const std::vector<int> &ids, const calibParamClass &calParam, int nunit, const int *calPar_index,
                            const int *groupmatrixindex, std::vector<double> &out);
// field_par: the analysed values in the order of extract_sub_parameters; calpar_range [2][26] (lower row, upper row);
// -> mat [26][nunit]
void enkf_parameters(const std::vector<int> &ids, const double *field_par, const calibParamClass &calParam, int nunit,
                     const int *calPar_index, const int *groupmatrixindex, const double *calpar_range, std::vector<double> &mat);

class parameterJsonFile {
  public:
    explicit parameterJsonFile(const calibParamClass &calParam);  // parameterJsonFile.cpp:65-125 (the caller has read the JSON)
    // every cell of a unit gets the unit's column of mat (:127-163)
    void parameterJsonFile_cda(const std::vector<double> &mat, const int *groupmatrixindex, int nr_cda_unit, int ids);
    // descriptors, the two ordinators (cell number, ArcID from filenameInputArcID: header line, then "arcid gcrc" pairs) and the 26
    // parameter arrays, numbers in the stream's default format (6 significant digits) like the reference (:207-525)
    void save(const std::string &filenameOutput, const std::string &filenameInputArcID) const;
    // one line per parameter: name and the units' values, width 33, scientific (:563-596)
    static void save_cda_txt(const std::string &filename, int nr_subbasins, const std::vector<double> &mat);
    const std::vector<double> &values() const { return v_; }  // [26][ncell]

  private:
    int ncell_;
    std::vector<double> v_;
};
}  // namespace wg
