// wg_rout_prepare.h — flow topology from D8 flow directions (product side).
// Replaces prepare_routing_files() (rout_prepare.h:4, rout_prepare.cpp:61-1031): same inputs,
// same output files, byte for byte; additionally returns the topology in memory.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace wg {

struct FlowTopology {
    int ncell = 0, ncol = 720, nrow = 360;
    std::vector<int8_t> ldd_2;            // G_LDD_2.UNF1 (before loop breaking)
    std::vector<int8_t> ldd;              // after loop breaking / nodata -> sink
    std::vector<int32_t> inflow9;         // [ncell][9] G_INFLC.9.UNF4
    std::vector<int16_t> flow_acc;        // G_FLOW_ACC.UNF2
    std::vector<uint16_t> basins, basins2, cells_to_outlet;
    std::vector<int32_t> outflow_cell;    // G_OUTFLC.UNF4
    std::vector<int32_t> neighbour8;      // [ncell][8]
    std::vector<int32_t> rout_order;      // G_ROUT_ORDER.UNF4 (1-based rank)
    std::vector<float> cell_distance;     // [9][nrow] GCELLDIST.9.UNF0
    std::vector<float> river_slope, river_length;
    std::vector<float> alloc_coeff;       // [ncell][5]
    std::vector<int8_t> start_month;
    int nlevels = 0, nbasins = 0, nbasins2 = 0;
};

// in-memory builder: flowdir = Arc codes (G_FLOWDIR.UNF2), row/col 1-based, gcrc [ncol][nrow]
FlowTopology build_flow_topology(const std::vector<int16_t> &flowdir, const std::vector<int16_t> &row,
                                 const std::vector<int16_t> &col, const std::vector<int32_t> &gcrc, int ncol, int nrow);
void river_geometry(FlowTopology &t, const std::vector<int16_t> &row, const std::vector<int16_t> &col,
                    const std::vector<float> &altitude, const std::vector<float> &meandering);
void reservoir_prepare(FlowTopology &t, const std::vector<float> &resarea, const std::vector<float> &mean_outflow,
                       const std::vector<float> &mean_outflow12);

// drop-in for the reference's prepare_routing_files: reads input_dir, writes routing_dir
FlowTopology prepare_routing_files(const std::string &input_dir, const std::string &routing_dir, short arcNumbersForFlowDir,
                                   short resOpt, int ncell);

}  // namespace wg
