// wg_pdaf_bridge.h - host-side look-alikes of the state exchange of the reference's PDAF coupling, over the host
// WghmStateFile / SnowInElevationFile classes (wg_state_files.h):
//   extract_sub      extractsub.cpp:17-79    monthly-mean state vector of the region's cells minus the temporal mean field
//   enkf_wghmstate   enKF2wghmState.cpp:17-121, 440-529 (state part; the calibration-parameter update of :127-431 needs the
//                    parameter JSON writer and is not reproduced): last day += analysis - prediction with the storage limits,
//                    snow in elevation rescaled to the analysed snow, monthly mean after the assimilation
// The device-side equivalents (no WghmStateFile round trip) are wgk_state_vector / wgk_enkf_update (include/wgk.h).
// Equal to the compiled reference's functions on tests/golden/ref_ng1000_enkf.npz (tests/test_host_library.py).
#pragma once
#include <vector>

#include "wg_state_files.h"

namespace wg {
// ids: 1-based cell numbers of the region (the reference reads them from the GIM header file); output [ids.size() * 10]
void extract_sub(const std::vector<int> &ids, WghmStateFile &wghmState, WghmStateFile &wghmMean, double *output);
// wghmState holds the days of the month on entry and only the updated last day on return; snow [cell][101];
// wghmStateMean (out, one day): the monthly mean after the assimilation
void enkf_wghmstate(const std::vector<int> &ids, const double *field, const double *prediction, WghmStateFile &wghmState,
                    SnowInElevationFile &snow_in_elevation, WghmStateFile &wghmMean, WghmStateFile &wghmStateMean);
}  // namespace wg
