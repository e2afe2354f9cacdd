// wg_rout_prepare.cpp — see wg_rout_prepare.h.  Vector based, every pass O(ncell) (the
// reference's reindex_waterbasins is O(nbasins x ncell), rout_prepare.cpp:585-617); integer
// outputs are bit-identical to the reference's files (tests/test_host_library.py).
#include "wg_rout_prepare.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>

#include "wg_grid.h"

namespace wg {
namespace {

struct Raster {  // index raster with a one-cell frame, wrapped in longitude (rout_prepare.cpp:134-147)
    int ncol, nrow;
    std::vector<int32_t> v;
    Raster(const std::vector<int32_t> &gcrc, int nc, int nr) : ncol(nc), nrow(nr), v((size_t)(nc + 2) * (nr + 2), 0) {
        for (int i = 0; i < nr; i++) {
            at(0, i + 1) = gcrc[(size_t)(nc - 1) * nr + i];
            at(nc + 1, i + 1) = gcrc[i];
            for (int j = 0; j < nc; j++) at(j + 1, i + 1) = gcrc[(size_t)j * nr + i];
        }
    }
    int32_t &at(int col, int row) { return v[(size_t)col * (nrow + 2) + row]; }
    int32_t at(int col, int row) const { return v[(size_t)col * (nrow + 2) + row]; }
};

// LDD code -> (dcol, drow) of the cell it points at; 1=SW 2=S 3=SE 4=W 6=E 7=NW 8=N 9=NE
const int kDc[10] = {0, -1, 0, +1, -1, 0, +1, -1, 0, +1};
const int kDr[10] = {0, +1, +1, +1, 0, 0, 0, -1, -1, -1};

void inflow_cells(const Raster &g, const std::vector<int16_t> &row, const std::vector<int16_t> &col,
                  const std::vector<int8_t> &ldd, std::vector<int32_t> &inf) {
    const int ng = (int)ldd.size();
    std::fill(inf.begin(), inf.end(), 0);
    for (int n = 0; n < ng; n++)
        for (int code = 1; code <= 9; code++) {
            if (code == 5) continue;
            // the neighbour whose LDD `code` points at n sits opposite to the direction of `code`
            const int32_t m = g.at(col[n] - kDc[code], row[n] - kDr[code]);
            if (m != 0 && ldd[m - 1] == code) inf[(size_t)n * 9 + (9 - code)] = m;  // rout_prepare.cpp:360-410
        }
}

void flow_accumulation(const std::vector<int32_t> &inf, std::vector<int16_t> &acc) {  // rout_prepare.cpp:442-494, int16 as there
    const int ng = (int)acc.size();
    for (int n = 0; n < ng; n++) {
        acc[n] = 1;
        for (int i = 0; i < 9; i++)
            if (inf[(size_t)n * 9 + i] != 0) acc[n] = 0;
    }
    int left = 99999, prev;
    do {
        prev = left;
        left = 0;
        for (int n = 0; n < ng; n++) {
            if (acc[n] != 0) continue;
            bool later = false;
            for (int i = 0; i < 9; i++) {
                const int32_t m = inf[(size_t)n * 9 + i];
                if (m != 0 && acc[m - 1] == 0) later = true;
            }
            if (later) { left++; continue; }
            int16_t fa = 1;
            for (int i = 0; i < 9; i++) {
                const int32_t m = inf[(size_t)n * 9 + i];
                if (m != 0) fa = (int16_t)(fa + acc[m - 1]);
            }
            acc[n] = fa;
        }
    } while (left > 0 && prev - left != 0);
}

}  // namespace

FlowTopology build_flow_topology(const std::vector<int16_t> &flowdir, const std::vector<int16_t> &row,
                                 const std::vector<int16_t> &col, const std::vector<int32_t> &gcrc, int ncol, int nrow) {
    FlowTopology t;
    const int ng = (int)flowdir.size();
    t.ncell = ng; t.ncol = ncol; t.nrow = nrow;
    t.ldd.resize(ng);
    for (int n = 0; n < ng; n++) {  // Arc codes -> LDD, rout_prepare.cpp:78-114
        int8_t d;
        switch (flowdir[n]) {
            case -1: d = -1; break; case 8: d = 1; break; case 4: d = 2; break; case 2: d = 3; break;
            case 16: d = 4; break; case 0: d = 5; break; case 1: d = 6; break; case 32: d = 7; break;
            case 64: d = 8; break; case 128: d = 9; break; default: d = 5; break;
        }
        t.ldd[n] = d;
    }
    t.ldd_2 = t.ldd;
    const Raster g(gcrc, ncol, nrow);
    t.inflow9.assign((size_t)ng * 9, 0);
    t.flow_acc.assign(ng, 0);
    inflow_cells(g, row, col, t.ldd, t.inflow9);
    flow_accumulation(t.inflow9, t.flow_acc);
    bool corrected = false;  // cells caught in loops never get a value: turn them into sinks (:164-177)
    for (int n = 0; n < ng; n++)
        if (t.flow_acc[n] == 0) { t.ldd[n] = 5; t.flow_acc[n] = 1; corrected = true; }
    if (corrected) inflow_cells(g, row, col, t.ldd, t.inflow9);  // G_FLOW_ACC.UNF2 keeps the first pass (:171)

    // basins: breadth-first from every outlet / inland sink, sweep by sweep as derive_basins (:497-583)
    t.basins.assign(ng, 0); t.cells_to_outlet.assign(ng, 0);
    std::vector<int32_t> frontier, next;
    int nb = 0;
    for (int n = 0; n < ng; n++) {
        if (t.ldd[n] == 0 || t.ldd[n] == -99) t.ldd[n] = 5;
        if (t.ldd[n] == 5 || t.ldd[n] == -1) { t.basins[n] = (uint16_t)++nb; frontier.push_back(n); }
    }
    t.nbasins = nb;
    for (int dist = 0; !frontier.empty(); dist++) {
        next.clear();
        for (int32_t n : frontier) {
            t.cells_to_outlet[n] = (uint16_t)dist;
            for (int j = 0; j < 9; j++) {
                const int32_t m = t.inflow9[(size_t)n * 9 + j];
                if (m != 0) { t.basins[m - 1] = t.basins[n]; next.push_back(m - 1); }
            }
        }
        std::sort(next.begin(), next.end());  // the reference sweeps cells in ascending order
        frontier.swap(next);
    }
    // basins with a single cell -> 0, the others renumbered in order of appearance (:585-617)
    std::vector<int32_t> cnt(nb + 1, 0), newid(nb + 1, 0);
    for (int n = 0; n < ng; n++) cnt[t.basins[n]]++;
    int j = 0;
    for (int b = 1; b <= nb; b++) newid[b] = (cnt[b] == 1) ? 0 : ++j;
    t.nbasins2 = j;
    t.basins2.resize(ng);
    for (int n = 0; n < ng; n++) t.basins2[n] = (uint16_t)newid[t.basins[n]];

    t.outflow_cell.assign(ng, 0);  // :619-670
    for (int n = 0; n < ng; n++) {
        const int d = t.ldd[n];
        if (d >= 1 && d <= 9 && d != 5) t.outflow_cell[n] = g.at(col[n] + kDc[d], row[n] + kDr[d]);
    }
    t.neighbour8.assign((size_t)ng * 8, 0);  // :672-698, E NE N NW W SW S SE
    const int nc8[8] = {+1, +1, 0, -1, -1, -1, 0, +1}, nr8[8] = {0, -1, -1, -1, 0, +1, +1, +1};
    for (int n = 0; n < ng; n++)
        for (int q = 0; q < 8; q++) t.neighbour8[(size_t)n * 8 + q] = g.at(col[n] + nc8[q], row[n] + nr8[q]);

    // routing order (:834-886): Kahn's algorithm by sweeps; a sweep numbers, in ascending cell
    // number, every cell whose upstream cells are all numbered.  Implemented level by level:
    // level(n) = 1 + max level of its upstream cells, then a stable sort by (level, n).
    std::vector<int32_t> remaining(ng, 0), level(ng, 0), order;
    for (int n = 0; n < ng; n++)
        for (int q = 0; q < 9; q++)
            if (t.inflow9[(size_t)n * 9 + q] > 0) remaining[n]++;
    frontier.clear();
    for (int n = 0; n < ng; n++)
        if (remaining[n] == 0) frontier.push_back(n);
    t.rout_order.assign(ng, -99);
    int rank = 1, nlev = 0;
    while (!frontier.empty()) {
        nlev++;
        next.clear();
        for (int32_t n : frontier) t.rout_order[n] = rank++;
        for (int32_t n : frontier) {
            const int32_t d = t.outflow_cell[n];
            if (d > 0 && --remaining[d - 1] == 0) next.push_back(d - 1);
        }
        std::sort(next.begin(), next.end());
        frontier.swap(next);
    }
    t.nlevels = nlev;
    for (int n = 0; n < ng; n++)
        if (t.rout_order[n] == -99) throw std::runtime_error("Routing order is not finished. Something is wrong in G_upstreamCells or/and G_outflow_cell");
    return t;
}

void river_geometry(FlowTopology &t, const std::vector<int16_t> &row, const std::vector<int16_t> &col,
                    const std::vector<float> &altitude, const std::vector<float> &meandering) {
    const int nrow = t.nrow, ng = t.ncell;
    auto &cd = t.cell_distance;
    cd.assign((size_t)9 * nrow, 0.f);
    auto CD = [&](int ch, int i) -> float & { return cd[(size_t)ch * nrow + i]; };
    // calculate_distances (:700-759): float arithmetic with a float pi, as in the reference
    const float pi = 3.141592653589793;
    const float earth_radius = 6371.211;
    const float vert_dist = 2 * pi * earth_radius * 0.5 / 360.0;
    for (int i = 0; i < nrow; i++) {
        const float l = 0.25 + i * 0.5;
        const float horiz = 2 * pi * earth_radius * sin(l * pi / 180.0) * 0.5 / 360.0;
        CD(1, i) = vert_dist; CD(3, i) = horiz; CD(4, i) = 0; CD(5, i) = horiz; CD(7, i) = vert_dist;
    }
    for (int i = 0; i < nrow - 1; i++) CD(0, i) = sqrt(CD(3, i) * CD(3, i + 1) + CD(1, i) * CD(1, i));
    CD(0, nrow - 1) = -99;
    for (int i = 1; i < nrow; i++) CD(6, i) = sqrt(CD(3, i) * CD(3, i - 1) + CD(1, i) * CD(1, i));
    CD(6, 0) = -99;
    for (int i = 0; i < nrow; i++) { CD(2, i) = CD(0, i); CD(8, i) = CD(6, i); }
    t.river_slope.assign(ng, 0.f);
    t.river_length.assign(ng, 0.f);
    for (int n = 0; n < ng; n++) {  // :761-794
        const int32_t o = t.outflow_cell[n];
        if (o != 0 && t.ldd[n] != 5 && t.ldd[n] != -1)
            t.river_slope[n] = (altitude[n] - altitude[o - 1]) / (1000.0 * CD(t.ldd[n] - 1, row[n] - 1) * ((meandering[n] + meandering[o - 1]) / 2));
        else
            t.river_slope[n] = 0;
    }
    const float minslope = 0.00001;
    for (int n = 0; n < ng; n++)
        if (t.river_slope[n] < minslope) t.river_slope[n] = minslope;
    for (int n = 0; n < ng; n++) {  // :796-832
        const int o = t.outflow_cell[n] - 1;
        float len;
        if (o == -1) len = 55.;
        else if (row[n] == row[o]) len = CD(3, row[n] - 1);
        else if (row[n] > row[o]) len = (col[n] == col[o]) ? CD(7, row[n] - 1) : CD(8, row[n] - 1);
        else len = (col[n] == col[o]) ? CD(1, row[n] - 1) : CD(2, row[n] - 1);
        if (o == -1) len *= meandering[n];
        else len *= ((meandering[n] + meandering[o]) / 2);
        t.river_length[n] = len;
    }
}

void reservoir_prepare(FlowTopology &t, const std::vector<float> &resarea, const std::vector<float> &mean_outflow,
                       const std::vector<float> &mean_outflow12) {  // rout_prepare.cpp:888-1031
    const int ng = t.ncell;
    // the reference reads the float file into a Grid<int> and only tests > 0 / <= 0 on the bit pattern
    std::vector<int32_t> ra(ng);
    std::memcpy(ra.data(), resarea.data(), sizeof(int32_t) * ng);
    std::vector<double> alloc((size_t)ng * 5, 1.0), inflowRes(ng, 0.0);
    for (int pass = 0; pass < 2; pass++)
        for (int n = 0; n < ng; n++) {
            if (ra[n] <= 0) continue;
            int d = t.outflow_cell[n];
            for (int i = 0; i < 5 && d > 0 && ra[d - 1] <= 0; i++, d = t.outflow_cell[d - 1]) {
                if (pass == 0) inflowRes[d - 1] += (double)mean_outflow[n];
                else alloc[(size_t)n * 5 + i] = (double)mean_outflow[n] / inflowRes[d - 1];
            }
        }
    t.alloc_coeff.resize((size_t)ng * 5);
    for (size_t q = 0; q < alloc.size(); q++) t.alloc_coeff[q] = (float)alloc[q];
    t.start_month.assign(ng, 1);
    for (int n = 0; n < ng; n++) {  // first month after the longest run of below-average months
        const double mo = (double)mean_outflow[n];
        auto dry = [&](int m) { return (double)mean_outflow12[(size_t)n * 12 + m] < mo; };
        int start = 0, dry_length = 0, start_ = 0, counter = 0, month = 0;
        bool last_dry = true;
        while (month < 12 && dry(month)) month++;
        const int beg = month;
        for (int m = 0; m < 12; m++) {
            month = m + beg;
            if (month > 11) month -= 12;
            if (dry(month)) {
                counter++;
                if (!last_dry) { start_ = month; last_dry = true; }
            } else if (last_dry) {
                last_dry = false;
                if (counter > dry_length) { start = start_; dry_length = counter; counter = 0; }
            }
        }
        if (counter > dry_length) start = start_;
        t.start_month[n] = (int8_t)(start + 1);
    }
}

FlowTopology prepare_routing_files(const std::string &in, const std::string &out, short arcNumbersForFlowDir, short resOpt, int ng) {
    if (arcNumbersForFlowDir != 1) throw std::runtime_error("only Arc flow direction codes (G_FLOWDIR.UNF2) are supported");
    Grid<int16_t> flowdir(ng), row(ng), col(ng);
    flowdir.read(in + "/G_FLOWDIR.UNF2");
    row.read(in + "/GR.UNF2");
    col.read(in + "/GC.UNF2");
    const int ncol = 720, nrow = 360;
    std::vector<int32_t> gcrc((size_t)ncol * nrow);
    read_unf_raw(in + "/GCRC.UNF4", gcrc.data(), gcrc.size());
    std::vector<int16_t> fd(flowdir.data(), flowdir.data() + ng), r(row.data(), row.data() + ng), c(col.data(), col.data() + ng);
    FlowTopology t = build_flow_topology(fd, r, c, gcrc, ncol, nrow);
    Grid<float> alt(ng), mea(ng);
    alt.read(in + "/GALTMOD.UNF0");
    mea.read(in + "/G_MEANDERING_RATIO.UNF0");
    river_geometry(t, r, c, std::vector<float>(alt.data(), alt.data() + ng), std::vector<float>(mea.data(), mea.data() + ng));
    write_unf_raw(out + "/G_LDD_2.UNF1", t.ldd_2.data(), t.ldd_2.size());
    write_unf_raw(out + "/G_INFLC.9.UNF4", t.inflow9.data(), t.inflow9.size());
    write_unf_raw(out + "/G_FLOW_ACC.UNF2", t.flow_acc.data(), t.flow_acc.size());
    write_unf_raw(out + "/G_CELLS_TO_OUTLET.UNF2", t.cells_to_outlet.data(), t.cells_to_outlet.size());
    write_unf_raw(out + "/G_BASINS.UNF2", t.basins.data(), t.basins.size());
    write_unf_raw(out + "/G_BASINS_2.UNF2", t.basins2.data(), t.basins2.size());
    write_unf_raw(out + "/G_OUTFLC.UNF4", t.outflow_cell.data(), t.outflow_cell.size());
    write_unf_raw(out + "/G_NEIGHBOUR_CELLS.8.UNF4", t.neighbour8.data(), t.neighbour8.size());
    write_unf_raw(out + "/GCELLDIST.9.UNF0", t.cell_distance.data(), t.cell_distance.size());
    write_unf_raw(out + "/G_RIVERSLOPE.UNF0", t.river_slope.data(), t.river_slope.size());
    write_unf_raw(out + "/G_RIVER_LENGTH.UNF0", t.river_length.data(), t.river_length.size());
    write_unf_raw(out + "/G_ROUT_ORDER.UNF4", t.rout_order.data(), t.rout_order.size());
    if (resOpt == 1) {
        Grid<float> ra(ng), mo(ng);
        Grid<float, 12> mo12(ng);
        ra.read(in + "/G_RESAREA.UNF0");
        mo.read(in + "/G_MEAN_OUTFLOW.UNF0");
        mo12.read(in + "/G_MEAN_OUTFLOW.12.UNF0");
        reservoir_prepare(t, std::vector<float>(ra.data(), ra.data() + ng), std::vector<float>(mo.data(), mo.data() + ng),
                          std::vector<float>(mo12.data(), mo12.data() + (size_t)ng * 12));
        write_unf_raw(out + "/G_ALLOC_COEFF.5.UNF0", t.alloc_coeff.data(), t.alloc_coeff.size());
        write_unf_raw(out + "/G_START_MONTH.UNF1", t.start_month.data(), t.start_month.size());
    }
    return t;
}

}  // namespace wg
