// wg_pdaf_bridge.cpp - see wg_pdaf_bridge.h
#include "wg_pdaf_bridge.h"

#include <cstddef>
#include <cstdio>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>

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

// ------------------------------------------------------------------------------------------
// parameter half
// ------------------------------------------------------------------------------------------
namespace {
// mean of parameter j over the cells of unit i, summed in region order (extractsub.cpp:121-338 / enKF2wghmState.cpp:183-397)
double unit_mean(const std::vector<int> &ids, const calibParamClass &cal, const int *gmi, int i, int j) {
    const size_t n = ids.size();
    double summe = 0;
    int count = 0;
    for (size_t c = 0; c < n; c++)
        if (gmi[c + (size_t)i * n] != 0) {
            summe += cal.getValue((eCalibParam)j, ids[c] - 1);
            count++;
        }
    return summe / count;  // a unit without cells gives 0/0 like the reference
}
}  // namespace

void extract_sub_parameters(const std::vector<int> &ids, const calibParamClass &cal, int nunit, const int *calPar_index,
                            const int *gmi, std::vector<double> &out) {
    out.clear();
    for (int i = 0; i < nunit; i++)
        for (int j = 0; j < 26; j++)
            if (calPar_index[j + i * 26] == 1) out.push_back(unit_mean(ids, cal, gmi, i, j));
}

void enkf_parameters(const std::vector<int> &ids, const double *field_par, const calibParamClass &cal, int nunit,
                     const int *calPar_index, const int *gmi, const double *range, std::vector<double> &mat) {
    mat.assign((size_t)26 * nunit, 0.);
    int index = 0;
    for (int i = 0; i < nunit; i++)
        for (int j = 0; j < 26; j++) {
            double &m = mat[(size_t)j * nunit + i];
            if (calPar_index[j + i * 26] == 1) {  // :170-180
                m = field_par[index++];
                if (m < range[j]) m = range[j];
                if (m > range[j + 26]) m = range[j + 26];
            } else if (calPar_index[j + i * 26] == 0) {
                m = unit_mean(ids, cal, gmi, i, j);
            }
        }
}

parameterJsonFile::parameterJsonFile(const calibParamClass &cal) : ncell_(cal.ncell()), v_(cal.block(), cal.block() + (size_t)26 * cal.ncell()) {}

void parameterJsonFile::parameterJsonFile_cda(const std::vector<double> &mat, const int *gmi, int nunit, int ids) {
    for (int i = 0; i < nunit; i++)
        for (int n = 0; n < ids; n++) {
            const int cell = gmi[n + (size_t)i * ids];
            if (cell <= 0) continue;
            if (cell > ncell_) throw std::out_of_range("parameterJsonFile_cda: cell number beyond the grid");  // vector::at
            for (int j = 0; j < 26; j++) v_[(size_t)j * ncell_ + cell - 1] = mat[(size_t)j * nunit + i];
        }
}

void parameterJsonFile::save_cda_txt(const std::string &filename, int nunit, const std::vector<double> &mat) {
    std::ofstream stream(filename.c_str(), std::ios::binary);
    if (!stream.good()) throw std::runtime_error("error by opening file");
    for (int j = 0; j < 26; j++) {
        stream << std::setw(33) << std::setfill(' ') << kParamNames[j];
        for (int i = 0; i < nunit; i++) stream << std::setw(33) << std::setfill(' ') << std::scientific << mat[(size_t)j * nunit + i];
        stream << std::endl;
    }
}

void parameterJsonFile::save(const std::string &out, const std::string &arcid_file) const {
    std::vector<long> arcid(ncell_, 0), gcrc(ncell_, 0);
    {
        std::ifstream in(arcid_file, std::ios::in | std::ios::binary);
        if (in.is_open()) {
            std::string header;
            std::getline(in, header, '\n');
            long a, g;
            for (int n = 0; n < ncell_ && (in >> a >> g); n++) {
                arcid[n] = a;
                gcrc[n] = g;
            }
        }
    }
    char now[64];
    {
        const time_t t = time(nullptr);
        strftime(now, sizeof now, "%Y-%m-%dT%H:%M:%S", localtime(&t));
    }
    std::ofstream f(out, std::ios::out | std::ios::binary);
    if (!f.is_open()) throw std::runtime_error("Unable to open file '" + out + "'");
    f << "{" << std::endl;
    f << "\"file_encoding\": \"UTF-8 without BOM\"," << std::endl;
    f << "\"n_descriptors\": 13," << std::endl;
    f << "\"n_ordinators\": 2," << std::endl;
    f << "\"n_parameters\": 26," << std::endl;
    f << "\"ng_param\": " << ncell_ << "," << std::endl;
    f << "\"watergap_landmask\": \"WLM\"," << std::endl;
    f << "\"watergap_version\": \"WaterGAP2.2b\"," << std::endl;
    f << "\"reference_year\": 1901," << std::endl;
    f << "\"reference_month\": 1," << std::endl;
    f << "\"creation_datetime\": \"" << now << "\"," << std::endl;
    f << "\"creation_institution\": \"BonnUniversity-IGG\"," << std::endl;
    f << "\"creation_staff\": \"APMG\"," << std::endl;
    f << "\"comments\": \"Cells filled with read-in data: (1) G_GAMMA_HBV.UNF0, G_CORR_FACTOR.UNF0, G_STAT_CORR.UNF0 (2) global "
         "parameters\"," << std::endl;
    for (int i = 0; i < 2; i++) {
        f << "\"" << (i == 0 ? "gcrc_cellnumber" : "arc_id") << "\": [";
        for (int n = 0; n < ncell_; n++) {
            if (gcrc[n] == n + 1) f << (i == 0 ? (long)n + 1 : arcid[n]);
            else f << "inconsistent_gcrc_in_at_n" << n;
            if (n < ncell_ - 1) f << ",";
        }
        f << "]," << std::endl;
    }
    for (int j = 0; j < 26; j++) {
        f << "\"" << kParamNames[j] << "\": [";
        for (int n = 0; n < ncell_; n++) {
            f << v_[(size_t)j * ncell_ + n];
            if (n < ncell_ - 1) f << ",";
        }
        f << "]";
        if (j < 25) f << "," << std::endl;
    }
    f << std::endl << "}";
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

// C entry for tests / other languages: the parameter half of one assimilation cycle.  parameter_json: the cycle's per-cell
// parameter file; ids [nids] 1-based cells of the region; calPar_index [nunit][26]; groupmatrixindex [nunit][nids];
// calpar_range [2][26]; perturb [number of ones in calPar_index]: analysis = extracted + perturb.  Writes cda_txt (time-evolution
// line file) and json_out (parameter file of the next cycle, ArcIDs from arcid_file); extract / field_par [number of ones],
// mat [26][nunit].  Returns the number of assimilated parameters, -1 on error (message on stderr).
extern "C" int wg_host_pdaf_parameters(const char *parameter_json, int nids, const int *ids_in, int nunit, const int *calPar_index,
                                       const int *groupmatrixindex, const double *calpar_range, const double *perturb,
                                       const char *cda_txt, const char *json_out, const char *arcid_file, double *extract,
                                       double *field_par, double *mat_out) {
    using namespace wg;
    try {
        calibParamClass cal;
        cal.readJson(parameter_json, 0);
        std::vector<int> ids(ids_in, ids_in + nids);
        std::vector<double> ex, mat;
        extract_sub_parameters(ids, cal, nunit, calPar_index, groupmatrixindex, ex);
        std::vector<double> field(ex.size());
        for (size_t k = 0; k < ex.size(); k++) {
            extract[k] = ex[k];
            field_par[k] = field[k] = ex[k] + perturb[k];
        }
        enkf_parameters(ids, field.data(), cal, nunit, calPar_index, groupmatrixindex, calpar_range, mat);
        for (size_t k = 0; k < mat.size(); k++) mat_out[k] = mat[k];
        parameterJsonFile::save_cda_txt(cda_txt, nunit, mat);
        parameterJsonFile pj(cal);
        pj.parameterJsonFile_cda(mat, groupmatrixindex, nunit, nids);
        pj.save(json_out, arcid_file);
        return (int)ex.size();
    } catch (const std::exception &e) {
        fprintf(stderr, "wg_host_pdaf_parameters: %s\n", e.what());
        return -1;
    }
}

// ------------------------------------------------------------------------------------------
// The symbols the PDAF Fortran side binds (extractsub.h / enKF2wghmState.h), argument for argument; the objects are the ones
// initialize_wghm_ / integrate_wghm_ (wg_model.cpp) hand out.  Files written: the monthly mean before the assimilation
// (output_state_mean + date), the updated last day, snow in elevation and additional file (the start values of the next
// cycle), the mean after the assimilation (<s2>states_mean_update_<member><date>.txt), and with calpar_size > 0 the
// time-evolution file <calparsample><date>.txt and the parameter JSON output_calibration_parameters.
// ------------------------------------------------------------------------------------------
namespace {
std::vector<int> read_region_ids(const char *file) {  // extractsub.cpp:32-56: header line, then "ID lon lat"
    std::ifstream stream(file, std::ios::binary);
    if (!stream.good()) throw std::runtime_error("error by opening file");
    std::vector<int> ids;
    std::string line;
    std::getline(stream, line);
    while (std::getline(stream, line)) {
        if (line.empty()) continue;
        std::stringstream ss(line);
        int id;
        double lambda, phi;
        ss >> id >> lambda >> phi;
        ids.push_back(id);
    }
    return ids;
}
}  // namespace

extern "C" {
void extract_sub_(const char *s, wg::WghmStateFile *&wghmState, double *&output, long *oy, wg::calibParamClass *&calParam,
                  long *total_nr_calPar, long *calpar_size, int *calPar_index, const char *calPar_filename, wg::WghmStateFile *&wghmMean,
                  int *groupmatrixindex) {
    using namespace wg;
    (void)calPar_filename;
    try {
    const std::vector<int> ids = read_region_ids(s);
    const long nr_par = *calpar_size;
    output = new double[ids.size() * 10 + nr_par];
    extract_sub(ids, *wghmState, *wghmMean, output);
    if (nr_par > 0) {
        if (*total_nr_calPar != 26) throw std::runtime_error("extract_sub_: 26 calibration parameters expected");
        std::vector<double> par;
        extract_sub_parameters(ids, *calParam, (int)*oy, calPar_index, groupmatrixindex, par);
        if ((long)par.size() != nr_par) throw std::runtime_error("extract_sub_: calpar_size does not match calPar_index");
        for (long k = 0; k < nr_par; k++) output[ids.size() * 10 + k] = par[k];
    }
    } catch (std::exception &e) {  // the reference rethrows with this prefix (extractsub.cpp:341-344)
        fprintf(stderr, "extract_sub_: %s\n", e.what());
        throw std::runtime_error(std::string("In watergap::main():\n") + e.what());
    }
}

void enkf_wghmstate_(const char *s, double *field, double *prediction, wg::ConfigFile *&configFile, wg::WghmStateFile *&wghmState,
                     wg::AdditionalOutputInputFile *&additionalOutIn, wg::SnowInElevationFile *&snow_in_elevation, long *step,
                     long *total_steps, long *year, long *month, long *ny, double *&output, wg::WghmStateFile *&wghmStateMean,
                     const char *s2, wg::calibParamClass *&calParam, long *calpar_size, const char *calparsample,
                     const char *path_calpar_IDs, const char *calPar_filename, wg::WghmStateFile *&wghmMean, double *calpar_range,
                     long *oy, long *total_nr_calPar, int *calPar_index, int *groupmatrixindex) {
    using namespace wg;
    (void)calPar_filename;
    try {
    const long nr_par = *calpar_size;
    char date[32];
    snprintf(date, sizeof date, "_%ld-%02ld", *year, *month);
    if (!configFile->outputmeanfile.empty()) {  // monthly mean before the assimilation (:38-49)
        std::string fn = configFile->outputmeanfile;
        const std::string::size_type pos = fn.rfind('.');
        if (pos == std::string::npos) fn += date;
        else fn.insert(pos, date);
        wghmState->saveMean(fn);
    }
    const std::vector<int> ids = read_region_ids(s);
    if ((long)ids.size() * 10 + nr_par != *ny) throw std::runtime_error("enkf_wghmstate_: the state vector does not match the region");
    wghmStateMean = new WghmStateFile;
    enkf_wghmstate(ids, field, prediction, *wghmState, *snow_in_elevation, *wghmMean, *wghmStateMean);
    if (!configFile->outputlastdayfile.empty()) wghmState->saveDay(configFile->outputlastdayfile, 0);
    if (nr_par > 0) {  // :127-431
        if (*total_nr_calPar != 26) throw std::runtime_error("enkf_wghmstate_: 26 calibration parameters expected");
        calibParamClass fileCal;  // the reference re-reads the cycle's parameter file for the JSON it rewrites
        fileCal.readJson(configFile->calibrationfile, calParam->ncell());
        std::vector<double> mat;
        enkf_parameters(ids, field + ids.size() * 10, *calParam, (int)*oy, calPar_index, groupmatrixindex, calpar_range, mat);
        parameterJsonFile::save_cda_txt(std::string(calparsample) + date + ".txt", (int)*oy, mat);
        parameterJsonFile paramJson(fileCal);
        paramJson.parameterJsonFile_cda(mat, groupmatrixindex, (int)*oy, (int)ids.size());
        paramJson.save(configFile->outputparameter, path_calpar_IDs);
    }
    if (!configFile->outputsnowlastdayfile.empty()) snow_in_elevation->save(configFile->outputsnowlastdayfile);
    {   // mean after the assimilation: <s2>states_mean_update_<3 characters before the extension of output_state_mean><date>.txt
        const std::string &str2 = configFile->outputmeanfile;
        std::string str = std::string(s2) + "states_mean_update_";
        if (str2.size() >= 7 && str2.substr(str2.find_last_of(".") + 1) == "txt") {
            str.append(str2.end() - 7, str2.end() - 4);
            str += date;
            str += ".txt";
        }
        if (str.size() > 4 && str.compare(str.size() - 4, 4, ".txt") == 0) wghmStateMean->saveMean(str);
    }
    if (!configFile->outputadditionalfile.empty()) additionalOutIn->save(configFile->outputadditionalfile);
    output = nullptr;  // the reference frees its snow factors before returning
    delete wghmStateMean;
    wghmStateMean = nullptr;
    if (*step == *total_steps - 1) {
        delete wghmState; wghmState = nullptr;
        delete calParam; calParam = nullptr;
        delete additionalOutIn; additionalOutIn = nullptr;
        delete snow_in_elevation; snow_in_elevation = nullptr;
        delete configFile; configFile = nullptr;
    }
    } catch (std::exception &e) {  // enKF2wghmState.cpp:586-589
        fprintf(stderr, "enkf_wghmstate_: %s\n", e.what());
        throw std::runtime_error(std::string("In watergap::main():\n") + e.what());
    }
}
}  // extern "C"
// C++ aliases of the reference (extractsub.cpp:349, enKF2wghmState.cpp:597)
void extract_sub(const char *s, wg::WghmStateFile *&wghmState, double *&output, long *oy, wg::calibParamClass *&calParam,
                 long *total_nr_calPar, long *calpar_size, int *calPar_index, const char *calPar_filename, wg::WghmStateFile *&wghmMean,
                 int *groupmatrixindex) {
    extract_sub_(s, wghmState, output, oy, calParam, total_nr_calPar, calpar_size, calPar_index, calPar_filename, wghmMean, groupmatrixindex);
}
void enkf_wghmstate(const char *s, double *field, double *prediction, wg::ConfigFile *&configFile, wg::WghmStateFile *&wghmState,
                    wg::AdditionalOutputInputFile *&additionalOutIn, wg::SnowInElevationFile *&snow_in_elevation, long *step,
                    long *total_steps, long *year, long *month, long *ny, double *&output, wg::WghmStateFile *&wghmStateMean,
                    const char *s2, wg::calibParamClass *&calParam, long *calpar_size, const char *calparsample,
                    const char *path_calpar_IDs, const char *calPar_filename, wg::WghmStateFile *&wghmMean, double *calpar_range,
                    long *oy, long *total_nr_calPar, int *calPar_index, int *groupmatrixindex) {
    enkf_wghmstate_(s, field, prediction, configFile, wghmState, additionalOutIn, snow_in_elevation, step, total_steps, year, month, ny,
                    output, wghmStateMean, s2, calParam, calpar_size, calparsample, path_calpar_IDs, calPar_filename, wghmMean,
                    calpar_range, oy, total_nr_calPar, calPar_index, groupmatrixindex);
}
