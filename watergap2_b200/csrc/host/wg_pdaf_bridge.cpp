// wg_pdaf_bridge.cpp - see wg_pdaf_bridge.h
#include "wg_pdaf_bridge.h"

#include <cstddef>

namespace wg {

void extract_sub(const std::vector<int> &ids, WghmStateFile &wghmState, WghmStateFile &wghmMean, double *output) {
    size_t j = 0;
    for (int id : ids) {  // extractsub.cpp:65-79: Cell::mean of every compartment minus the temporal mean field
        const Cell m = wghmState.cell(id - 1).mean();
        Cell mm = m;
        for (int k = 0; k < 10; k++) output[j++] = mm.compartment(k, 0) - wghmMean.cell(id - 1).compartment(k, 0);
    }
}

namespace {
// compartments that may not become negative (enKF2wghmState.cpp:89-121: all but local lake, global lake, groundwater)
constexpr bool kNonNegative[10] = {true, true, true, false, true, false, true, true, true, false};
inline void limit(Cell &c) {
    for (int k = 0; k < 10; k++)
        if (kNonNegative[k] && c.compartment(k, 0) < 0.) c.compartment(k, 0) = 0.;
    if (c.snow(0) > 1000.) c.snow(0) = 1000.;
}
}  // namespace

void enkf_wghmstate(const std::vector<int> &ids, const double *field, const double *prediction, WghmStateFile &wghmState,
                    SnowInElevationFile &snow, WghmStateFile &wghmMean, WghmStateFile &wghmStateMean) {
    const size_t ncell = wghmState.size();
    // monthly mean before the assimilation, last day of the month (:20-24)
    wghmStateMean = WghmStateFile(ncell, 1);
    for (size_t c = 0; c < ncell; c++) {
        Cell &src = wghmState.cell((int)c);
        wghmStateMean.cell((int)c) = src.mean();
        Cell last((size_t)src.id(), 1);
        for (int k = 0; k < 10; k++) last.compartment(k, 0) = src.compartment(k, (int)src.size() - 1);
        src = last;
    }
    // last day += analysis - prediction, with the storage limits (:89-121)
    for (size_t i = 0; i < ids.size(); i++) {
        Cell &c = wghmState.cell(ids[i] - 1);
        for (int k = 0; k < 10; k++) c.compartment(k, 0) += field[i * 10 + k] - prediction[i * 10 + k];
        limit(c);
    }
    // snow in elevation follows the analysed monthly snow (:440-471)
    for (size_t i = 0; i < ids.size(); i++) {
        const int n = ids[i] - 1;
        const double before = wghmStateMean.cell(n).snow(0);
        const double after = field[i * 10 + 1] + wghmMean.cell(n).snow(0);
        const double factor = before == 0 ? 0. : after / before;
        for (int e = 1; e < 101; e++) {
            double &s = snow.snowInElevation(n, e);
            s = before == 0 ? after / 100 : s * factor;
            if (s < 0.) s = 0.;
            if (s > 1000.) s = 1000.;
        }
    }
    // monthly mean after the assimilation = analysis + temporal mean field, same limits (:485-529)
    for (size_t i = 0; i < ids.size(); i++) {
        Cell &c = wghmStateMean.cell(ids[i] - 1);
        for (int k = 0; k < 10; k++) c.compartment(k, 0) = field[i * 10 + k] + wghmMean.cell(ids[i] - 1).compartment(k, 0);
        limit(c);
    }
}

}  // namespace wg

// C entry for tests / other languages: one assimilation cycle on plain arrays.  n cells (the region = all of them),
// daily [n][10][ndays], snow_elev [n][101] (in/out), mean_field / perturb [n][10]; the "analysis" is the extracted vector
// plus perturb.  Outputs: extract, field, lastday, mean_after [n][10].
extern "C" int wg_host_pdaf_cycle(int n, int ndays, const double *daily, double *snow_elev, const double *mean_field,
                                  const double *perturb, double *extract, double *field, double *lastday, double *mean_after) {
    using namespace wg;
    WghmStateFile state(n, ndays), mean(n, 1), after;
    SnowInElevationFile snow(n);
    std::vector<int> ids(n);
    for (int c = 0; c < n; c++) {
        ids[c] = c + 1;
        for (int k = 0; k < 10; k++) {
            for (int d = 0; d < ndays; d++) state.cell(c).compartment(k, d) = daily[((size_t)c * 10 + k) * ndays + d];
            mean.cell(c).compartment(k, 0) = mean_field[(size_t)c * 10 + k];
        }
        for (int e = 0; e < 101; e++) snow.snowInElevation(c, e) = snow_elev[(size_t)c * 101 + e];
    }
    extract_sub(ids, state, mean, extract);
    for (size_t k = 0; k < (size_t)n * 10; k++) field[k] = extract[k] + perturb[k];
    enkf_wghmstate(ids, field, extract, state, snow, mean, after);
    for (int c = 0; c < n; c++) {
        for (int k = 0; k < 10; k++) {
            lastday[(size_t)c * 10 + k] = state.cell(c).compartment(k, 0);
            mean_after[(size_t)c * 10 + k] = after.cell(c).compartment(k, 0);
        }
        for (int e = 0; e < 101; e++) snow_elev[(size_t)c * 101 + e] = snow.snowInElevation(c, e);
    }
    return 0;
}
