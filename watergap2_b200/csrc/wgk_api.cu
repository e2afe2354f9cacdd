// wgk_api.cu — C ABI (include/wgk.h) over the sm_100a kernels of wgk_kernels.cuh.
//
// Owns device memory, the routing-order permutation, the level index and the per-day CUDA
// graph.  There is no CPU fallback anywhere in this file: every entry point either runs the
// CUDA path or returns an error code.
#include "../../include/wgk.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cuda_runtime.h>

namespace wgk { constexpr int NBAND = 101; }
#define WGK_NBAND_K wgk::NBAND
#include "wgk_kernels.cuh"  // namespace wgk: kernels of the cell-minor layout
#undef WGK_MM
#define WGK_MM 1
#include "wgk_kernels.cuh"  // namespace wgk_mm: the same kernels compiled for the member-minor layout (lane = member)
#undef WGK_WU
#define WGK_WU 1
#include "wgk_kernels.cuh"  // namespace wgk_mm_wu: member-minor, the kernels that contain water use compiled with it
#undef WGK_MM
#define WGK_MM 0
#include "wgk_kernels.cuh"  // namespace wgk_wu: cell-minor with water use
// a layout-dependent kernel of the context's layout
#define WGK_K(c_, name_) ((c_)->mm ? wgk_mm::name_ : wgk::name_)
// a kernel that contains water-use code: by layout and by wgk_options.subtract_use
#define WGK_KW(c_, name_) ((c_)->opt.subtract_use > 0 ? ((c_)->mm ? wgk_mm_wu::name_ : wgk_wu::name_) : ((c_)->mm ? wgk_mm::name_ : wgk::name_))

namespace {

struct FieldInfo {
    const char *name;
    const char *dtype;
    int scope;
    int bands;
    int elsize;
    size_t offset;  // of the pointer inside WgkArrays
};

const FieldInfo kFields[] = {
#define X(name, ctype, dt, scope, bands) {#name, dt, WGK_SCOPE_##scope, bands, (int)sizeof(ctype), offsetof(WgkArrays, name)},
    WGK_FIELDS(X)
#undef X
};

// eCalibParam -> device parameter array (calib_param.h:72-101)
struct ParamMap { int k; int field; };
const ParamMap kParamMap[] = {
    {0, WGK_F_gamma_hbv}, {1, WGK_F_cfa}, {2, WGK_F_cfs}, {4, WGK_F_p_rivrgh}, {7, WGK_F_p_swoutf},
    {8, WGK_F_p_evaredex}, {9, WGK_F_p_netrad}, {10, WGK_F_p_ptc_hum}, {11, WGK_F_p_ptc_ari},
    {12, WGK_F_p_pet_mxdy}, {13, WGK_F_p_mcwh}, {15, WGK_F_p_snowfz}, {16, WGK_F_p_snowmt},
    {17, WGK_F_p_degday}, {18, WGK_F_p_gradnt}, {21, WGK_F_p_pcrit}, {22, WGK_F_p_gwoutf}, {25, WGK_F_p_prec},
};

}  // namespace

struct wgk_ctx {
    int device = 0;
    int ncell = 0, stride = 0, nmember = 0, npset = 0;
    // layout of the member / parameter-set arrays (WgkParams): cell-minor [member][cell] or member-minor [cell][member]
    bool mm = false;
    int mpad = 0, ppad = 0;  // rows of the member / parameter-set arrays (padded to 32 when member-minor)
    void *d_stage = nullptr;  // device staging row for strided uploads / downloads (member-minor)
    size_t d_stage_bytes = 0;
    wgk_options opt{};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;

    WgkArrays arrays{};
    std::vector<void *> allocs;

    // topology
    bool have_topology = false;
    int nlevels = 0;
    int tail_level0 = 0;
    // "rank" below = device position: the routing rank, or - with cell classes - the routing rank re-sorted by
    // class inside each dependency level
    std::vector<int32_t> rank_of_cell, cell_of_rank, level_off, level_of_rank;
    std::vector<uint8_t> cell_class;
    int32_t *d_rank_of_cell = nullptr;
    std::vector<int32_t> down_pos;       // downstream device position of every device position, or -1
    int32_t *d_wu_res_idx = nullptr;     // water use: [5][stride] downstream cells an irrigation reservoir serves
    int32_t *d_cell_of_rank = nullptr, *d_up_off = nullptr, *d_up_idx = nullptr, *d_down = nullptr, *d_level_off = nullptr;
    int32_t *d_member_pset = nullptr;
    std::vector<int32_t> member_pset;
    int32_t *d_cal = nullptr;

    // forcing
    float4 *d_forcing = nullptr;
    int forcing_nslots = 0, forcing_per_member = 0;
    float *d_fstage = nullptr;  // 4 staging grids [ncell][stride]
    // forcing uploads run on their own stream, so that the host->device copy and the pack of the next
    // period overlap the stepping of the current one; ordering against the steps is by events:
    // a step waits for the uploads issued before it, an upload waits for the steps that still read its slots
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_forcing = nullptr;
    bool forcing_pending = false;
    struct SlotUse { int lo, n; cudaEvent_t ev; };
    std::vector<SlotUse> slot_uses;
    size_t fstage_elems = 0;

    // record
    double *d_record = nullptr;
    int32_t *d_record_cells = nullptr;
    int nrec = 0, record_max_days = 0;
    int record_rows = 0;  // rows written since the record was (re)started

    unsigned long long *d_stamps = nullptr;  // wgk_stamps: globaltimer stamps of the level-0 tasks

    // ensemble moments (sum, sum of squares of the state vector over this context's members), [ncells][10] each
    double *d_mom_sum = nullptr, *d_mom_sumsq = nullptr;
    int32_t *d_mom_pos = nullptr;
    size_t mom_cap = 0;
    int mom_ncells = 0;

    // staging
    void *h_stage = nullptr;
    size_t h_stage_bytes = 0;
    double *d_partial = nullptr;

    // per-call calendar table and the discharge buffers of the days in flight
    int32_t *d_cal_days = nullptr;
    double *d_qbuf = nullptr;
    std::vector<int> chunk_lo;  // tail chunks: levels [chunk_lo[c], chunk_lo[c+1])
    int wave_level0 = 0;             // the wavefront's own split into wide levels [0, wave_level0) and tail chunks (fused level
    std::vector<int> wave_chunk_lo;  // tasks treat every level as wide; otherwise = tail_level0 / chunk_lo)
    int wave_tail_threshold = -1;    // < 0: the context's tail threshold

    // CUDA graphs of the (day, level) wavefront, one per call length
    std::map<int, cudaGraphExec_t> graphs;
    std::map<int, int> graph_nodes;
    int64_t launches = 0;
    bool derived_dirty = true;  // s_c1 / s_slope_pow / s_flags need (re)computation
    bool member_dirty = true;   // s_snowfree needs (re)computation (band state uploaded or exposed)
    bool month_acc = false;     // EnKF bridge: accumulate the daily WghmStateFile entries of the month
    int month_days = 0;
    bool whole_day = false;     // many members: whole-grid kernels day after day instead of the (day, level) wavefront
    int level_edges = 0;        // fused level tasks: 0 programmatic edge from the upstream level, 1 full edge
    int level_tasks = 0;        // wavefront: V(d, l) and R(d, l) of a wide level as one task (k_level_day): 0 no, 1 every wide level
                                // (programmatic edges between the levels), 2 level 0 only (headwater cells: no upstream level)
    int form = 0;               // vertical kernel form: 0 thread per cell, 1 band-parallel 5 threads/cell, 2 band-parallel 2 threads/cell
    int32_t *d_gidx = nullptr;  // [ncell] index into the global-water-body scratch or -1
    double *d_gbody = nullptr;  // [nmember][ngbody][GB_N]
    int ngbody = 0;

    // cell-owner schedule (k_days_owner): one launch per call, every thread owns one cell for all its days
    int owner_mode = 0;         // 1 = WGK_DAY_SCHEDULE=owner (opt-in)
    int owner_fits = -1;        // cached co-residency check (-1 unknown)
    int owner_nwarps = 0;
    int32_t *d_own_warp_begin = nullptr, *d_own_warp_end = nullptr, *d_own_cell_warp = nullptr, *d_own_abort = nullptr;
    uint32_t *d_own_progress = nullptr;
    unsigned long long *d_own_ring = nullptr;  // [QBUF_K][nmember][stride][2] tagged discharge entries
    uint32_t owner_base = 0;                   // days stepped by k_days_owner so far (tag base)
    int32_t *d_rec_head = nullptr, *d_rec_next = nullptr;
    long long *d_own_timing = nullptr;  // WGK_OWNER_TIMING=1: per-warp cycle counts of the last call (development aid)
};

namespace {

int fail(wgk_ctx *c, int code, const char *fmt, ...) {
    if (c) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        c->err = buf;
    }
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(c, WGK_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

void **field_slot(wgk_ctx *c, int f) { return (void **)((char *)&c->arrays + kFields[f].offset); }

size_t field_rows(const wgk_ctx *c, int f) {
    switch (kFields[f].scope) {
        case WGK_SCOPE_CELL: return 1;
        case WGK_SCOPE_PSET: return (size_t)c->ppad;
        case WGK_SCOPE_MEMBER: return (size_t)c->mpad;
        default: return 1;
    }
}
// valid indices of a field (members / parameter sets; the padded rows of the member-minor layout are not addressable)
size_t field_index_count(const wgk_ctx *c, int f) {
    switch (kFields[f].scope) {
        case WGK_SCOPE_PSET: return (size_t)c->npset;
        case WGK_SCOPE_MEMBER: return (size_t)c->nmember;
        default: return 1;
    }
}
size_t field_row_elems(const wgk_ctx *c, int f) {
    if (kFields[f].scope == WGK_SCOPE_TABLE) return WGK_NLCT;
    return (size_t)c->stride * kFields[f].bands;
}

int ensure_stage(wgk_ctx *c, size_t bytes) {
    if (c->h_stage_bytes >= bytes) return 0;
    if (c->h_stage) cudaFreeHost(c->h_stage);
    c->h_stage = nullptr;
    c->h_stage_bytes = 0;
    CU(cudaMallocHost(&c->h_stage, bytes));
    c->h_stage_bytes = bytes;
    return 0;
}

// distance (in elements) between two consecutive cells of one index of a member / parameter-set field: 1 in the cell-minor
// layout, the padded row count in the member-minor layout ([band][cell][index])
int index_pad(const wgk_ctx *c, int f) {
    if (!c->mm) return 1;
    if (kFields[f].scope == WGK_SCOPE_MEMBER) return c->mpad;
    if (kFields[f].scope == WGK_SCOPE_PSET) return c->ppad;
    return 1;
}
int ensure_dstage(wgk_ctx *c, size_t bytes) {
    if (c->d_stage_bytes >= bytes) return 0;
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_stage);
    c->d_stage = nullptr;
    c->d_stage_bytes = 0;
    CU(cudaMalloc(&c->d_stage, bytes));
    c->d_stage_bytes = bytes;
    return 0;
}
// dst[k * dst_stride] = src[k * src_stride] for k < n, elements of `es` bytes (1, 2, 4 or 8), on the context's stream
void launch_strided(wgk_ctx *c, char *dst, size_t dst_stride, const char *src, size_t src_stride, size_t n, int es) {
    const unsigned grid = (unsigned)((n + 255) / 256);
    switch (es) {
        case 1: wgk::k_strided_copy<uint8_t><<<grid, 256, 0, c->stream>>>((uint8_t *)dst, dst_stride, (const uint8_t *)src, src_stride, n); break;
        case 2: wgk::k_strided_copy<uint16_t><<<grid, 256, 0, c->stream>>>((uint16_t *)dst, dst_stride, (const uint16_t *)src, src_stride, n); break;
        case 4: wgk::k_strided_copy<uint32_t><<<grid, 256, 0, c->stream>>>((uint32_t *)dst, dst_stride, (const uint32_t *)src, src_stride, n); break;
        default: wgk::k_strided_copy<unsigned long long><<<grid, 256, 0, c->stream>>>((unsigned long long *)dst, dst_stride, (const unsigned long long *)src, src_stride, n); break;
    }
    c->launches++;
}

WgkParams make_params(const wgk_ctx *c) {
    WgkParams p{};
    p.a = c->arrays;
    p.member_pset = c->d_member_pset;
    p.forcing = c->d_forcing;
    p.up_off = c->d_up_off;
    p.up_idx = c->d_up_idx;
    p.down = c->d_down;
    p.level_off = c->d_level_off;
    p.cal = c->d_cal;
    p.cal_days = c->d_cal_days;
    p.qbuf = c->d_qbuf;
    p.gidx = c->d_gidx;
    p.gbody = c->d_gbody;
    p.ngbody = c->ngbody;
    p.record = c->d_record;
    p.record_cells = c->d_record_cells;
    p.nrec = c->nrec;
    p.record_max_days = c->record_max_days;
    p.ncell = c->ncell;
    p.stride = c->stride;
    p.nmember = c->nmember;
    p.npset = c->npset;
    p.forcing_nslots = c->forcing_nslots;
    p.forcing_per_member = c->forcing_per_member;
    p.restart = c->opt.restart;
    p.month_acc = c->month_acc ? 1 : 0;
    p.subtract_use = c->opt.subtract_use;
    p.wu_res_idx = c->d_wu_res_idx;
    p.nlevels = c->nlevels;
    p.mm = c->mm ? 1 : 0;
    p.mpad = c->mpad;
    p.ppad = c->ppad;
    p.stamps = c->d_stamps;
    {
        const char *e = getenv("WGK_STAMP_LEVEL");
        p.stamp_level = e ? atoi(e) : 0;
    }
    return p;
}

void drop_graph(wgk_ctx *c) {
    for (auto &kv : c->graphs) cudaGraphExecDestroy(kv.second);
    c->graphs.clear();
    c->graph_nodes.clear();
}

constexpr int MAX_CALL_DAYS = 366;
int levels_per_chunk() {  // tuning knob (WGK_LEVELS_PER_CHUNK); measured optimum on B200: 1
    const char *e = getenv("WGK_LEVELS_PER_CHUNK");
    const int v = e ? atoi(e) : 1;
    return v > 0 ? v : 1;
}

void *cells_pre_fn(const wgk_ctx *c) {
    if (c->opt.subtract_use > 0 && c->form)
        return c->form == 1 ? (void *)wgk_wu::k_cells_pre<wgk_wu::VCfgSmall> : (void *)wgk_wu::k_cells_pre<wgk_wu::VCfgMid>;
    return c->form == 1 ? (void *)wgk::k_cells_pre<wgk::VCfgSmall> : c->form == 2 ? (void *)wgk::k_cells_pre<wgk::VCfgMid> : (void *)WGK_KW(c, k_cells_pre_tpc);
}
void *vertical_fn(const wgk_ctx *c) {
    return c->form == 1 ? (void *)wgk::k_vertical<wgk::VCfgSmall> : c->form == 2 ? (void *)wgk::k_vertical<wgk::VCfgMid> : (void *)WGK_K(c, k_vertical_tpc);
}
// grid of a cell-parallel kernel of `block` threads over `ncells` device positions (wgk::map_thread)
dim3 cell_grid(const wgk_ctx *c, int ncells, int block) {
#if WGK_MM_UNIFORM_CELL
    if (c->mm) return dim3(ncells, (c->mpad + std::min(block, c->mpad) - 1) / std::min(block, c->mpad));
#else
    if (c->mm) return dim3((unsigned)(((long long)ncells * c->mpad + block - 1) / block), 1);
#endif
    return dim3((ncells + block - 1) / block, c->nmember);
}
// threads per CTA of a cell-parallel kernel whose cell-minor form uses `block`
dim3 cell_block(const wgk_ctx *c, int block) {
#if WGK_MM_UNIFORM_CELL
    if (c->mm) return dim3(std::min(block, c->mpad));
#endif
    (void)c;
    return dim3(block);
}
dim3 cells_pre_block(const wgk_ctx *c) { return c->form == 1 ? dim3(wgk::VCfgSmall::THREADS) : c->form == 2 ? dim3(wgk::VCfgMid::THREADS) : cell_block(c, wgk::VBLOCK); }
dim3 cells_pre_grid(const wgk_ctx *c, int begin, int end) {
    return c->form ? dim3(wgk::v_num_tiles(begin, end), c->nmember) : cell_grid(c, end - begin, wgk::VBLOCK);
}
cudaError_t launch_cells_pre(wgk_ctx *c, const WgkParams &p, int d, int begin, int end) {
    void *args[] = {(void *)&p, &d, &begin, &end};
    return cudaLaunchKernel(cells_pre_fn(c), cells_pre_grid(c, begin, end), cells_pre_block(c), args, 0, c->stream);
}
cudaError_t launch_vertical(wgk_ctx *c, const WgkParams &p, int d) {
    void *args[] = {(void *)&p, &d};
    return cudaLaunchKernel(vertical_fn(c), cells_pre_grid(c, 0, c->ncell), cells_pre_block(c), args, 0, c->stream);
}

// plain launches of one simulated day (day offset `d` of the current call) on c->stream, phase
// by phase over the whole grid; used by the three-call class-shim path and by wgk_profile_day
int enqueue_vertical(wgk_ctx *c, const WgkParams &p, int d) {
    launch_vertical(c, p, d);
    return 1;
}
int enqueue_routing(wgk_ctx *c, const WgkParams &p, int d) {
    int n = 0;
    const dim3 block = cell_block(c, 128), grid = cell_grid(c, c->ncell, 128);
    WGK_KW(c, k_route_local)<<<grid, block, 0, c->stream>>>(p, d);
    n++;
    for (int l = 0; l < c->tail_level0; l++) {
        const int cnt = c->level_off[l + 1] - c->level_off[l];
        const dim3 g = cell_grid(c, cnt, 128);
        WGK_KW(c, k_route_level)<<<g, block, 0, c->stream>>>(p, d, l);
        n++;
    }
    if (c->tail_level0 < c->nlevels) {
        (c->opt.subtract_use > 0 ? wgk_wu::k_route_tail : wgk::k_route_tail)<<<c->nmember, 256, 0, c->stream>>>(p, d, c->tail_level0, c->nlevels);
        n++;
    }
    WGK_K(c, k_route_post)<<<grid, block, 0, c->stream>>>(p);
    n++;
    return n;
}

// the kernels of `ndays` days as (day, level) tasks in dependency order; serial order on one
// stream satisfies every dependency (used when use_graph == 0)
int enqueue_wavefront_serial(wgk_ctx *c, const WgkParams &p, int ndays) {
    int n = 0;
    const dim3 block = cell_block(c, 128);
    for (int d = 0; d < ndays; d++) {
        for (int l = 0; l < c->wave_level0; l++) {
            const int begin = c->level_off[l], end = c->level_off[l + 1];
            const dim3 g = cell_grid(c, end - begin, 128);
            if (c->level_tasks == 1 || (c->level_tasks == 2 && l == 0)) {
                WGK_KW(c, k_level_day)<<<cell_grid(c, end - begin, wgk::VBLOCK), cell_block(c, wgk::VBLOCK), 0, c->stream>>>(p, d, l);
                n += 1;
                continue;
            }
            launch_cells_pre(c, p, d, begin, end);
            WGK_KW(c, k_river_level)<<<g, block, 0, c->stream>>>(p, d, l);
            n += 2;
        }
        for (size_t k = 0; k + 1 < c->wave_chunk_lo.size(); k++) {
            const int lo = c->wave_chunk_lo[k], hi = c->wave_chunk_lo[k + 1];
            const int begin = c->level_off[lo], end = c->level_off[hi];
            launch_cells_pre(c, p, d, begin, end);
            (c->opt.subtract_use > 0 ? wgk_wu::k_tail_chunk : wgk::k_tail_chunk)<<<c->nmember, 256, 0, c->stream>>>(p, d, lo, hi);
            n += 2;
        }
        if (c->d_record) {
            WGK_K(c, k_end_of_day)<<<1, 256, 0, c->stream>>>(p, d);
            n++;
        }
    }
    return n;
}

// whole-grid kernels day after day (many members), in stream order: the vertical balance, the local routing, the wide
// levels one launch each, the narrow tail in one persistent CTA per member, the post-pass.  These are the UNFUSED kernels of
// the three-call path: with every kernel filling the GPU, the fused forms of the wavefront (vertical + local routing in
// k_cells_pre*, river + post in k_river_level) lose more to their register count and code size than they save in re-loads
// (measured on B200, 0.5 degree grid, us per member-day, fused / unfused: 64 members 44.5 / 39.1, 128 members 41.2 / 35.8).
int enqueue_whole_days(wgk_ctx *c, const WgkParams &p, int ndays) {
    int n = 0;
    for (int d = 0; d < ndays; d++) {
        n += enqueue_vertical(c, p, d);
        n += enqueue_routing(c, p, d);
        if (c->d_record) {
            WGK_K(c, k_end_of_day)<<<1, 256, 0, c->stream>>>(p, d);
            n++;
        }
    }
    return n;
}

// The same tasks as a CUDA graph whose edges are exactly the data dependencies:
//   V(d, l)  <- R(d-1, l)             vertical balance + local routing: own state of the previous day
//   R(d, l)  <- V(d, l), R(d, l-1)    river + post: upstream discharge of the same day
//   first R task of day d <- end of day d-QBUF_K   (discharge buffer reuse)
// so that R(d, l), R(d+1, l-1), R(d+2, l-2) ... and all V tasks execute concurrently.
int build_wavefront_graph(wgk_ctx *c, const WgkParams &p, int ndays, cudaGraphExec_t *out, int *nnodes) {
    cudaGraph_t g;
    CU(cudaGraphCreate(&g, 0));
    const int W = c->wave_level0;
    const int C = (int)c->wave_chunk_lo.size() - 1 > 0 ? (int)c->wave_chunk_lo.size() - 1 : 0;
    std::vector<cudaGraphNode_t> prevW(W, nullptr), prevT(C, nullptr), dayEnd(ndays, nullptr);
    int count = 0;
    auto add = [&](void *fn, dim3 grid, dim3 block, void **args, std::vector<cudaGraphNode_t> deps, cudaGraphNode_t *node) -> cudaError_t {
        cudaKernelNodeParams kp{};
        kp.func = fn;
        kp.gridDim = grid;
        kp.blockDim = block;
        kp.sharedMemBytes = 0;
        kp.kernelParams = args;
        kp.extra = nullptr;
        std::vector<cudaGraphNode_t> dd;
        for (auto x : deps) if (x) dd.push_back(x);
        count++;
        return cudaGraphAddKernelNode(node, g, dd.data(), dd.size(), &kp);
    };
    WgkParams pp = p;
    void *pre_fn = cells_pre_fn(c);
    const dim3 pre_block = cells_pre_block(c);
    auto pre_grid = [&](int begin, int end) { return cells_pre_grid(c, begin, end); };
    for (int d = 0; d < ndays; d++) {
        cudaGraphNode_t reuse = (d >= wgk::QBUF_K) ? dayEnd[d - wgk::QBUF_K] : nullptr;
        cudaGraphNode_t last = nullptr;
        bool first_sweep = true;
        for (int l = 0; l < W; l++) {
            int begin = c->level_off[l], end = c->level_off[l + 1];
            int dd = d, ll = l;
            const dim3 grid = cell_grid(c, end - begin, 128);
            if (c->level_tasks == 1 || (c->level_tasks == 2 && l == 0)) {
                // VR(d, l): full edges from the level's own previous day (and the discharge-buffer reuse), a programmatic edge from
                // the upstream level of the same day: the grid starts once the upstream grid is resident and waits inside the
                // kernel, after its vertical part, for the upstream grid to complete
                void *a[] = {&pp, &dd, &ll};
                cudaGraphNode_t node;
                const bool plain = c->level_edges == 1;  // full edge from the upstream level instead of the programmatic one
                CU(add((void *)WGK_KW(c, k_level_day), cell_grid(c, end - begin, wgk::VBLOCK), cell_block(c, wgk::VBLOCK), a,
                       {prevW[l], plain ? last : nullptr, first_sweep ? reuse : nullptr}, &node));
                if (last && !plain) {
                    cudaGraphEdgeData ed{};
                    ed.from_port = cudaGraphKernelNodePortLaunchCompletion;
                    ed.to_port = 0;
                    ed.type = cudaGraphDependencyTypeProgrammatic;
                    CU(cudaGraphAddDependencies_v2(g, &last, &node, &ed, 1));
                }
                first_sweep = false;
                prevW[l] = node;
                last = node;
                continue;
            }
            // V(d, l): vertical balance + local routing, waits only for the cells' own previous day
            void *a1[] = {&pp, &dd, &begin, &end};
            cudaGraphNode_t pre, node;
            CU(add(pre_fn, pre_grid(begin, end), pre_block, a1, {prevW[l]}, &pre));
            // R(d, l): river + post, additionally waits for the upstream level of the same day
            // (fusing R(d, l) with V(d + 1, l) into one task - one kernel boundary per day on the own-cell
            //  recurrence instead of two - was measured SLOWER: 27.4 vs 24.6 ms per simulated year, with 32 or 64 buffers)
            void *a2[] = {&pp, &dd, &ll};
            CU(add((void *)WGK_KW(c, k_river_level), grid, cell_block(c, 128), a2, {pre, last, first_sweep ? reuse : nullptr}, &node));
            first_sweep = false;
            prevW[l] = node;
            last = node;
        }
        for (int k = 0; k < C; k++) {
            int lo = c->wave_chunk_lo[k], hi = c->wave_chunk_lo[k + 1];
            int begin = c->level_off[lo], end = c->level_off[hi], dd = d;
            void *a1[] = {&pp, &dd, &begin, &end};
            cudaGraphNode_t pre, sweep;
            CU(add(pre_fn, pre_grid(begin, end), pre_block, a1, {prevT[k]}, &pre));
            void *a2[] = {&pp, &dd, &lo, &hi};
            CU(add((void *)(c->opt.subtract_use > 0 ? wgk_wu::k_tail_chunk : wgk::k_tail_chunk), dim3(c->nmember), dim3(256), a2, {pre, last, first_sweep ? reuse : nullptr}, &sweep));
            first_sweep = false;
            prevT[k] = sweep;
            last = sweep;
        }
        if (c->d_record) {
            int dd = d;
            void *a3[] = {&pp, &dd};
            cudaGraphNode_t e;
            CU(add((void *)WGK_K(c, k_end_of_day), dim3(1), dim3(256), a3, {last, d > 0 ? dayEnd[d - 1] : nullptr}, &e));  // day ends in order
            last = e;
        }
        dayEnd[d] = last;
    }
    cudaError_t e = cudaGraphInstantiate(out, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(c, WGK_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    *nnodes = count;
    return 0;
}

// cell-owner schedule: usable when every CTA of k_days_owner is resident at the same time (the threads wait for
// each other inside the kernel), which the cooperative launch then guarantees
bool owner_usable(wgk_ctx *c) {
    if (c->owner_mode == 0 || c->owner_nwarps <= 0) return false;
    if (c->owner_fits < 0) {
        int nb = 0, sms = 0, coop = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
        const size_t smem = sizeof(wgk::SnowStage) * (wgk::OWN_BLOCK / wgk::VBLOCK);
        cudaFuncSetAttribute((c->opt.subtract_use > 0 ? (const void *)wgk_wu::k_days_owner : (const void *)wgk::k_days_owner), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (c->opt.subtract_use > 0 ? (const void *)wgk_wu::k_days_owner : (const void *)wgk::k_days_owner), wgk::OWN_BLOCK, smem);
        const long long ctas = (long long)((c->owner_nwarps + wgk::OWN_BLOCK / 32 - 1) / (wgk::OWN_BLOCK / 32)) * c->nmember;
        c->owner_fits = (coop && ctas <= (long long)nb * sms) ? 1 : 0;
    }
    return c->owner_fits == 1;
}
int launch_owner(wgk_ctx *c, const WgkParams &p, int ndays) {
    WgkOwner s{};
    s.warp_begin = c->d_own_warp_begin;
    s.warp_end = c->d_own_warp_end;
    s.cell_warp = c->d_own_cell_warp;
    s.progress = c->d_own_progress;
    s.abort = c->d_own_abort;
    s.rec_head = c->d_rec_head;
    s.rec_next = c->d_rec_next;
    s.nwarps = c->owner_nwarps;
    if (getenv("WGK_OWNER_TIMING")) {
        if (!c->d_own_timing) CU(cudaMalloc(&c->d_own_timing, sizeof(long long) * 4 * (size_t)c->nmember * c->owner_nwarps));
        s.timing = c->d_own_timing;
    }
    if (!c->d_own_ring) {
        const size_t bytes = sizeof(unsigned long long) * 2 * wgk::QBUF_K * (size_t)c->nmember * c->stride;
        CU(cudaMalloc(&c->d_own_ring, bytes));
        CU(cudaMemsetAsync(c->d_own_ring, 0, bytes, c->stream));
    }
    s.ring = c->d_own_ring;
    s.base = c->owner_base;
    c->owner_base += (uint32_t)ndays;
    const dim3 grid((c->owner_nwarps + wgk::OWN_BLOCK / 32 - 1) / (wgk::OWN_BLOCK / 32), c->nmember);
    void *args[] = {(void *)&p, (void *)&s, (void *)&ndays};
    CU(cudaLaunchCooperativeKernel((c->opt.subtract_use > 0 ? (const void *)wgk_wu::k_days_owner : (const void *)wgk::k_days_owner), grid, dim3(wgk::OWN_BLOCK), args,
                                   sizeof(wgk::SnowStage) * (wgk::OWN_BLOCK / wgk::VBLOCK), c->stream));
    c->launches += 1;
    return WGK_OK;
}

// inflow-independent river constants and cell class flags, recomputed after any static or
// parameter upload (never inside a graph capture)
int ensure_derived(wgk_ctx *c) {
    if (c->member_dirty) {
        const dim3 block = cell_block(c, 128), grid = cell_grid(c, c->ncell, 128);
        WGK_K(c, k_derive_member)<<<grid, block, 0, c->stream>>>(make_params(c));
        c->launches++;
        c->member_dirty = false;
    }
    if (!c->derived_dirty) return 0;
    dim3 block(128), grid((c->ncell + 127) / 128, c->npset);
    WGK_K(c, k_derive_static)<<<grid, block, 0, c->stream>>>(make_params(c));
    c->launches++;
    // cells with a global lake / reservoir / global wetland get a slot in the per-day scratch
    std::vector<int8_t> flags(c->stride);
    CU(cudaMemcpyAsync(flags.data(), c->arrays.s_flags, c->stride, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    std::vector<int32_t> gidx(c->ncell, -1);
    int n = 0;
    for (int r = 0; r < c->ncell; r++)
        if (flags[r] & (wgk::FL_LAKE | wgk::FL_RES | wgk::FL_GLOWET)) gidx[r] = n++;
    if (c->d_gidx) cudaFree(c->d_gidx);
    if (c->d_gbody) cudaFree(c->d_gbody);
    c->d_gidx = nullptr;
    c->d_gbody = nullptr;
    CU(cudaMalloc(&c->d_gidx, sizeof(int32_t) * c->ncell));
    CU(cudaMemcpy(c->d_gidx, gidx.data(), sizeof(int32_t) * c->ncell, cudaMemcpyHostToDevice));
    const size_t nb = (size_t)c->mpad * std::max(1, n) * wgk::GB_N;
    CU(cudaMalloc(&c->d_gbody, nb * sizeof(double)));
    CU(cudaMemset(c->d_gbody, 0, nb * sizeof(double)));
    c->ngbody = n;
    if (c->opt.subtract_use > 0) {
        // irrigation reservoirs release for their own use and for up to 5 downstream cells without a reservoir (routing.cpp:2961-2973).
        // The reference tests G_reservoir_area[downstreamCell] with the 1-based cell number as a 0-based index, i.e. the cell
        // numbered downstreamCell + 1; that indexing is reproduced here (an index of ng ends the chain).
        std::vector<double> res((size_t)c->stride);
        CU(cudaMemcpyAsync(res.data(), c->arrays.reservoir_area, sizeof(double) * c->stride, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        std::vector<int32_t> idx((size_t)5 * c->stride, -1);
        const int ng = c->ncell;
        auto refnum_down = [&](int x) { return c->down_pos[x] >= 0 ? c->cell_of_rank[c->down_pos[x]] + 1 : 0; };
        for (int x = 0; x < ng; x++) {
            int i = 0, dc = refnum_down(x);
            while (i < 5 && dc > 0 && dc < ng && res[c->rank_of_cell[dc]] <= 0) {
                const int xd = c->rank_of_cell[dc - 1];
                idx[(size_t)i * c->stride + x] = xd;
                i++;
                dc = refnum_down(xd);
            }
        }
        if (!c->d_wu_res_idx) CU(cudaMalloc(&c->d_wu_res_idx, sizeof(int32_t) * idx.size()));
        CU(cudaMemcpy(c->d_wu_res_idx, idx.data(), sizeof(int32_t) * idx.size(), cudaMemcpyHostToDevice));
    }
    // cells that are inactive (now) never write their discharge entry: no stale value of an earlier configuration may
    // feed a downstream gather or the published discharge field (every active cell rewrites its entry before it is read)
    CU(cudaMemsetAsync(c->d_qbuf, 0, (size_t)wgk::QBUF_K * c->mpad * c->stride * sizeof(double), c->stream));
    c->derived_dirty = false;
    drop_graph(c);
    return 0;
}

int check_ready(wgk_ctx *c) {
    if (!c) return WGK_ERR_ARG;
    if (!c->have_topology) return fail(c, WGK_ERR_STATE, "wgk_set_topology must be called first");
    if (!c->d_forcing) return fail(c, WGK_ERR_STATE, "no forcing on the device (wgk_forcing_reserve / wgk_set_forcing)");
    return 0;
}

}  // namespace

extern "C" {

int wgk_create(wgk_ctx **out, int device, int ncell, int nmember, int npset, const wgk_options *opt) {
    if (!out || ncell <= 0 || nmember <= 0 || npset <= 0) return WGK_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        *out = nullptr;
        return WGK_ERR_CUDA;  // no GPU: the product path fails loudly, there is no CPU fallback
    }
    wgk_ctx *c = new wgk_ctx();
    *out = c;
    c->device = device;
    c->ncell = ncell;
    c->stride = (ncell + 31) / 32 * 32;
    c->nmember = nmember;
    c->npset = npset;
    if (opt) c->opt = *opt;
    else { c->opt.restart = 0; c->opt.tail_threshold = 0; c->opt.use_graph = 1; c->opt.subtract_use = 0; }
    if (c->opt.subtract_use != 0 && c->opt.subtract_use != 2)
        return fail(c, WGK_ERR_ARG, "wgk_options.subtract_use %d: 0 (no water use) or 2 (net abstractions) are implemented", c->opt.subtract_use);
    if ((c->opt.restart != 0 && c->opt.restart != 1) || (c->opt.use_graph != 0 && c->opt.use_graph != 1) || c->opt.tail_threshold < 0)
        return fail(c, WGK_ERR_ARG, "wgk_options: restart %d / use_graph %d must be 0 or 1, tail_threshold %d >= 0", c->opt.restart,
                    c->opt.use_graph, c->opt.tail_threshold);
    {   // Form of the vertical kernel.  The band-parallel tile forms (5 or 2 threads per cell in the band loop) were the default
        // of small problems in round 1, when the thread-per-cell kernels ran as two-kernel tasks; against the fused (day, level)
        // task of round 2 they lose at every size (one member, ms per simulated year, tile form / thread per cell in fused
        // tasks: 3 000 cells 14.1 / 10.4, 10 000 16.4 / 10.9, 20 000 16.7 / 11.1, 30 000 18.5 / 11.9, 67 420 28.8 / 16.9),
        // so the thread-per-cell form is the default everywhere and the tile forms stay selectable (and tested).
        const char *e = getenv("WGK_VERTICAL_FORM");  // "cells" | "bands" | "bands2" (tests exercise all)
        if (e && !strcmp(e, "bands")) c->form = 1;
        else if (e && !strcmp(e, "bands2")) c->form = 2;
        else c->form = 0;
    }
    {   // Schedule of a multi-day call.  Few members: the (day, level) wavefront, which hides the level-to-level
        // latency chain of a day behind the following days.  Many members: every kernel fills the GPU on its own, and
        // whole-grid (unfused) kernels day after day win (measured on B200, 0.5 degree grid, us per member-day,
        // wavefront / whole-day: 16 members 46.7 / 56.4, 32 members 45.4 / 44.4, 64 members 44.3 / 38.1,
        // 128 members 43.6 / 35.1).  A single member on a 2.2 M-cell grid stays with the wavefront (475 vs 590 ms per
        // year): its narrow levels are one CTA per member.
        const char *e = getenv("WGK_DAY_SCHEDULE");  // "owner" | "wavefront" | "wholeday"
        if (e && !strcmp(e, "wavefront")) c->whole_day = false;
        else if (e && !strcmp(e, "wholeday")) c->whole_day = true;
        else c->whole_day = (nmember >= 32 && (long long)nmember * ncell >= 6400000);
        // "owner": one launch per call, a thread owns its cell for all days (k_days_owner); needs every cell-member
        // co-resident (<= 75 776 on B200).  Opt-in: measured slower than the wavefront (93 vs 67 us per day, see the kernel).
        c->owner_mode = (e && !strcmp(e, "owner")) ? 1 : 0;
    }
    {   // Layout.  From 32 members on: member-minor, a warp = 32 members of one cell (lane = member), so that the lanes share the
        // cell's statics, land cover, water-body class and nearly the same weather (cell-minor warps run 20 of 32 lanes per
        // instruction and are issue-bound, profiles/r1_k_vertical_tpc_m16.md; member-minor: 32 of 32, 34 % fewer warp
        // instructions, profiles/r2_k_vertical_tpc_m64_*.md).  It comes with the thread-per-cell kernels and one launch per
        // routing level (no persistent narrow-level CTA).  Results are bit-identical between the layouts
        // (tests/test_gpu_round2.py::test_member_minor_layout_is_bit_identical).
        // Measured on B200, 0.5 degree grid, 10^9 cell-days/s (members: member-minor + wavefront / member-minor + whole-day /
        // cell-minor + whole-day / cell-minor + wavefront): 32: 1.66 / 1.47 / 1.43 / 1.42, 64: 1.96 / 1.89 / 1.69 / 1.46,
        // 128: 2.05 / 2.13 / 1.87 / 1.49, 256: - / 2.29 / 1.93 / - : hence the whole-day schedule from ~6.4 M cell-members.
        const char *e = getenv("WGK_LAYOUT");  // "members" | "cells"
        if (e && !strcmp(e, "members")) c->mm = true;
        else if (e && !strcmp(e, "cells")) c->mm = false;
        else c->mm = c->form == 0 && c->owner_mode == 0 && nmember >= 32;
        if (c->mm) {
            c->form = 0;
            c->owner_mode = 0;
        }
        c->mpad = c->mm ? (nmember + 31) / 32 * 32 : nmember;
        c->ppad = c->mm ? (npset == 1 ? 1 : (npset + 31) / 32 * 32) : npset;
    }
    {   // Tasks of the wavefront.  "split": two kernels per (day, wide level) - vertical balance + local routing, then river +
        // post once the upstream level is through - and a persistent one-CTA sweep per narrow level.  "fused": ONE kernel per
        // (day, level) for every level (k_level_day: V part, griddepcontrol.wait, R part) with a programmatic
        // (launch-completion) edge from the upstream level, so that the grid starts while the upstream grid runs; "fusedfull":
        // the same with a full edge; "fused0": only the headwater level.  Measured on B200 (0.5 degree grid, one member, ms per
        // simulated year; k_level_day compiled for 2 resident CTAs per SM = no register cap, WGK_LEVEL_MINB):
        //   split, narrow levels <= 256 cells as tail chunks   21.7        fused, tail chunks <= 256 cells   24.0
        //   split, every level wide                            20.8        fused, tail chunks <= 8 cells     22.3
        //   fusedfull, every level a fused task                27.0        fused, every level a fused task   19.9
        // With the 128-register cap of the thread-per-cell kernels the fused task loses (20.6, and 23.7 with tail chunks - the
        // round-2 measurement that had kept "split" the default): the V and R parts share one register budget.  The narrow
        // levels must be fused tasks as well: every level has the own-cell recurrence V(d,l) -> R(d,l) -> V(d+1,l), and the
        // slowest one paces the whole wavefront.
        const char *e = getenv("WGK_LEVEL_TASKS");  // "split" | "fused" | "fused0" | "fusedfull"
        // The fused task wins while the run is latency-bound and loses once the kernels fill the GPU (one member, cells on the GPU,
        // ms per simulated month fused / split: 135 k 3.02 / 3.59, 270 k 5.38 / 5.76, 539 k 10.04 / 10.27, 1.08 M 19.13 / 18.92,
        // 2.16 M 37.2 / 35.8): default below 800 k cell-members.
        const bool fused_default = c->form == 0 && !c->mm && !c->whole_day && c->opt.subtract_use == 0 && (long long)nmember * ncell < 800000;
        const int want = (e && (!strcmp(e, "fused") || !strcmp(e, "fusedfull"))) ? 1 : (e && !strcmp(e, "fused0")) ? 2 : (e && !strcmp(e, "split")) ? 0
                         : (fused_default ? 1 : 0);
        c->level_tasks = c->form != 0 ? 0 : want;
        c->level_edges = (e && !strcmp(e, "fusedfull")) ? 1 : 0;
        const char *t = getenv("WGK_WAVE_TAIL_THRESHOLD");  // narrow levels of the wavefront as tail chunks up to this many cells
        c->wave_tail_threshold = t ? atoi(t) : (c->level_tasks == 1 ? 0 : -1);
    }
    if (c->opt.tail_threshold <= 0) {
        const char *e = getenv("WGK_TAIL_THRESHOLD");
        c->opt.tail_threshold = (e && atoi(e) > 0) ? atoi(e) : 256;
    }
    CU(cudaSetDevice(device));
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_forcing, cudaEventDisableTiming));
    // the tile kernels keep up to 36 KB per CTA in shared memory: ask for the largest carve-out
    for (const void *fn : {(const void *)wgk::k_cells_pre<wgk::VCfgSmall>, (const void *)wgk::k_cells_pre<wgk::VCfgMid>,
                           (const void *)wgk::k_vertical<wgk::VCfgSmall>, (const void *)wgk::k_vertical<wgk::VCfgMid>,
                           (const void *)wgk::k_cells_pre_tpc, (const void *)wgk::k_vertical_tpc,
                           (const void *)wgk_mm::k_cells_pre_tpc, (const void *)wgk_mm::k_vertical_tpc,
                           (const void *)wgk_wu::k_cells_pre<wgk_wu::VCfgSmall>, (const void *)wgk_wu::k_cells_pre<wgk_wu::VCfgMid>,
                           (const void *)wgk_wu::k_cells_pre_tpc, (const void *)wgk_mm_wu::k_cells_pre_tpc})
        CU(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    for (int f = 0; f < WGK_F_COUNT; f++) {
        if (c->opt.subtract_use == 0 && strncmp(kFields[f].name, "wu_", 3) == 0) continue;  // the water-use arrays exist only with water use
        const size_t bytes = field_rows(c, f) * field_row_elems(c, f) * kFields[f].elsize;
        void *d = nullptr;
        cudaError_t e = cudaMalloc(&d, bytes);
        if (e != cudaSuccess) return fail(c, WGK_ERR_NOMEM, "cudaMalloc(%zu) for field %s: %s", bytes, kFields[f].name, cudaGetErrorString(e));
        CU(cudaMemsetAsync(d, 0, bytes, c->stream));
        *field_slot(c, f) = d;
        c->allocs.push_back(d);
    }
    c->member_pset.assign(nmember, 0);
    for (int m = 0; m < nmember; m++) c->member_pset[m] = (npset == nmember) ? m : 0;
    CU(cudaMalloc(&c->d_member_pset, sizeof(int32_t) * nmember));
    CU(cudaMemcpyAsync(c->d_member_pset, c->member_pset.data(), sizeof(int32_t) * nmember, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMalloc(&c->d_cal, sizeof(int32_t) * 8));
    CU(cudaMemsetAsync(c->d_cal, 0, sizeof(int32_t) * 8, c->stream));
    CU(cudaMalloc(&c->d_partial, sizeof(double) * 256));
    CU(cudaMalloc(&c->d_cal_days, sizeof(int32_t) * 4 * MAX_CALL_DAYS));
    CU(cudaMemsetAsync(c->d_cal_days, 0, sizeof(int32_t) * 4 * MAX_CALL_DAYS, c->stream));
    const size_t qn = (size_t)wgk::QBUF_K * c->mpad * c->stride;
    if (cudaMalloc(&c->d_qbuf, qn * sizeof(double)) != cudaSuccess) return fail(c, WGK_ERR_NOMEM, "cudaMalloc discharge buffers");
    CU(cudaMemsetAsync(c->d_qbuf, 0, qn * sizeof(double), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return WGK_OK;
}

void wgk_destroy(wgk_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    for (auto &u : c->slot_uses) cudaEventDestroy(u.ev);
    if (c->ev_forcing) cudaEventDestroy(c->ev_forcing);
    drop_graph(c);
    for (void *d : c->allocs) cudaFree(d);
    cudaFree(c->d_rank_of_cell); cudaFree(c->d_wu_res_idx);
    cudaFree(c->d_cell_of_rank); cudaFree(c->d_up_off); cudaFree(c->d_up_idx); cudaFree(c->d_down);
    cudaFree(c->d_level_off); cudaFree(c->d_member_pset); cudaFree(c->d_cal); cudaFree(c->d_forcing);
    cudaFree(c->d_gidx); cudaFree(c->d_gbody); cudaFree(c->d_cal_days); cudaFree(c->d_qbuf);
    cudaFree(c->d_mom_sum); cudaFree(c->d_mom_sumsq); cudaFree(c->d_mom_pos); cudaFree(c->d_stamps); cudaFree(c->d_stage);
    cudaFree(c->d_fstage); cudaFree(c->d_record); cudaFree(c->d_record_cells); cudaFree(c->d_partial);
    cudaFree(c->d_own_warp_begin); cudaFree(c->d_own_warp_end); cudaFree(c->d_own_cell_warp); cudaFree(c->d_own_progress);
    cudaFree(c->d_own_abort); cudaFree(c->d_rec_head); cudaFree(c->d_rec_next); cudaFree(c->d_own_timing); cudaFree(c->d_own_ring);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char *wgk_last_error(const wgk_ctx *c) { return c ? c->err.c_str() : "null context"; }

int wgk_synchronize(wgk_ctx *c) {
    if (!c) return WGK_ERR_ARG;
    CU(cudaStreamSynchronize(c->copy_stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->d_own_abort) {  // a hand-off wait of the cell-owner schedule timed out (never expected; not a hang)
        int32_t ab = 0;
        CU(cudaMemcpy(&ab, c->d_own_abort, sizeof ab, cudaMemcpyDeviceToHost));
        if (ab) {
            cudaMemset(c->d_own_abort, 0, sizeof ab);
            return fail(c, WGK_ERR_CUDA, "cell-owner schedule aborted: a discharge hand-off was not published in time; the state is incomplete");
        }
    }
    return WGK_OK;
}

void *wgk_get_stream(wgk_ctx *c) { return c ? (void *)c->stream : nullptr; }

int wgk_set_stream(wgk_ctx *c, void *s) {
    if (!c) return WGK_ERR_ARG;
    CU(cudaStreamSynchronize(c->stream));
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)s;
    c->own_stream = false;
    drop_graph(c);
    return WGK_OK;
}

// ---------------------------------------------------------------------------------------
// topology
// ---------------------------------------------------------------------------------------
int wgk_set_cell_classes(wgk_ctx *c, const uint8_t *cell_class) {
    if (!c) return WGK_ERR_ARG;
    if (cell_class) c->cell_class.assign(cell_class, cell_class + c->ncell);
    else c->cell_class.clear();
    return WGK_OK;
}

int wgk_set_topology(wgk_ctx *c, const int32_t *rout_order, const int32_t *downstream_cell) {
    if (!c || !rout_order || !downstream_cell) return WGK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    const int ng = c->ncell;
    // reference routing rank (G_ROUT_ORDER) <-> cell
    std::vector<int32_t> cell_of_r(ng, -1), r_of_cell(ng, -1);
    for (int n = 0; n < ng; n++) {
        const int r = rout_order[n] - 1;
        if (r < 0 || r >= ng || cell_of_r[r] != -1) return fail(c, WGK_ERR_TOPOLOGY, "rout_order is not a permutation of 1..ncell (cell %d)", n + 1);
        cell_of_r[r] = n;
        r_of_cell[n] = r;
    }
    std::vector<int32_t> down_r(ng, -1);
    for (int n = 0; n < ng; n++) {
        const int d = downstream_cell[n];
        if (d < 0 || d > ng) return fail(c, WGK_ERR_TOPOLOGY, "downstream cell of cell %d out of range", n + 1);
        if (d > 0) {
            const int rd = r_of_cell[d - 1], rn = r_of_cell[n];
            if (rd <= rn) return fail(c, WGK_ERR_TOPOLOGY, "cell %d is routed before its upstream cell %d", d, n + 1);
            down_r[rn] = rd;
        }
    }
    // dependency level = longest path from a headwater; the rank order of rout_prepare.cpp's
    // Kahn sweeps is level-major, which is verified here
    std::vector<int32_t> level_r(ng, 0);
    for (int r = 0; r < ng; r++)
        if (down_r[r] >= 0) level_r[down_r[r]] = std::max(level_r[down_r[r]], level_r[r] + 1);
    for (int r = 1; r < ng; r++)
        if (level_r[r] < level_r[r - 1])
            return fail(c, WGK_ERR_TOPOLOGY, "routing order is not level-major at rank %d (not produced by rout_order sweeps)", r);
    c->nlevels = level_r[ng - 1] + 1;
    c->level_off.assign(c->nlevels + 1, 0);
    for (int r = 0; r < ng; r++) c->level_off[level_r[r] + 1]++;
    for (int l = 0; l < c->nlevels; l++) c->level_off[l + 1] += c->level_off[l];
    // device position: rank order, stably re-sorted by class inside each level
    std::vector<int32_t> r_of_pos(ng), pos_of_r(ng);
    for (int r = 0; r < ng; r++) r_of_pos[r] = r;
    if (!c->cell_class.empty()) {
        const char *e = getenv("WGK_CLASS_ORDER");  // "asc" | "desc" | "cost"
        // default "cost": inside a level the classes with the longest code path first (global water body, then local lake /
        // wetland), so that the slowest warps of a task are launched first instead of last (level-0 task 42.6 -> 40.3 us, year
        // 17.1 -> 16.8 ms); "asc" = plain ascending key (round 1)
        const int mode = (e && !strcmp(e, "desc")) ? 1 : (e && !strcmp(e, "asc")) ? 0 : (e && !strcmp(e, "asc2")) ? 3 : (e && !strcmp(e, "cost2")) ? 4 : 2;
        auto key = [&](int r) -> int {
            const int k = c->cell_class[cell_of_r[r]];
            if (mode == 0) return k;
            if (mode == 1) return 255 - k;
            const int cost = 3 * ((k >> 2) & 1) + (k & 1) + ((k >> 1) & 1);  // global water body, local lake, local wetland
            if (mode == 3) return (k & 15) * 16 + (k >> 4);                    // water-body class first, the high 4 bits inside it
            if (mode == 4) return (7 - cost) * 256 + (k & 15) * 16 + (k >> 4);
            return (7 - cost) * 256 + k;
        };
        for (int l = 0; l < c->nlevels; l++)
            std::stable_sort(r_of_pos.begin() + c->level_off[l], r_of_pos.begin() + c->level_off[l + 1],
                             [&](int a, int b) { return key(a) < key(b); });
    }
    for (int x = 0; x < ng; x++) pos_of_r[r_of_pos[x]] = x;
    c->rank_of_cell.assign(ng, -1);
    c->cell_of_rank.assign(ng, -1);
    c->level_of_rank.assign(ng, 0);
    std::vector<int32_t> down(ng, -1), nup(ng, 0);
    for (int x = 0; x < ng; x++) {
        const int r = r_of_pos[x];
        c->cell_of_rank[x] = cell_of_r[r];
        c->rank_of_cell[cell_of_r[r]] = x;
        c->level_of_rank[x] = level_r[r];
        if (down_r[r] >= 0) {
            down[x] = pos_of_r[down_r[r]];
            nup[down[x]]++;
        }
    }
    // upstream CSR in device positions; the entries of a cell ascend in reference rank = the order in
    // which the reference adds them to G_riverInflow (routing.cpp:3955-3958)
    std::vector<int32_t> up_off(ng + 1, 0);
    for (int x = 0; x < ng; x++) up_off[x + 1] = up_off[x] + nup[x];
    std::vector<int32_t> up_idx(std::max(1, up_off[ng])), fill(ng, 0);
    for (int r = 0; r < ng; r++)
        if (down_r[r] >= 0) {
            const int xd = pos_of_r[down_r[r]];
            up_idx[up_off[xd] + fill[xd]++] = pos_of_r[r];
        }
    // narrow tail: first level from which every level has <= tail_threshold cells
    c->tail_level0 = c->nlevels;
    for (int l = c->nlevels - 1; l >= 0 && !c->mm; l--) {  // (member-minor: every level is a launch of its own, nmember threads per cell)
        if (c->level_off[l + 1] - c->level_off[l] <= c->opt.tail_threshold) c->tail_level0 = l;
        else break;
    }
    c->chunk_lo.clear();
    for (int l = c->tail_level0; l < c->nlevels; l += levels_per_chunk()) c->chunk_lo.push_back(l);
    if (c->tail_level0 < c->nlevels) c->chunk_lo.push_back(c->nlevels);
    c->wave_level0 = c->tail_level0;
    c->wave_chunk_lo = c->chunk_lo;
    if (c->wave_tail_threshold >= 0 && !c->mm) {
        c->wave_level0 = c->nlevels;
        for (int l = c->nlevels - 1; l >= 0; l--) {
            if (c->level_off[l + 1] - c->level_off[l] <= c->wave_tail_threshold) c->wave_level0 = l;
            else break;
        }
        c->wave_chunk_lo.clear();
        for (int l = c->wave_level0; l < c->nlevels; l += levels_per_chunk()) c->wave_chunk_lo.push_back(l);
        if (c->wave_level0 < c->nlevels) c->wave_chunk_lo.push_back(c->nlevels);
    }
    auto upload = [&](int32_t *&dptr, const std::vector<int32_t> &v) -> cudaError_t {
        if (dptr) cudaFree(dptr);
        dptr = nullptr;
        cudaError_t e = cudaMalloc(&dptr, sizeof(int32_t) * std::max<size_t>(1, v.size()));
        if (e != cudaSuccess) return e;
        return cudaMemcpy(dptr, v.data(), sizeof(int32_t) * v.size(), cudaMemcpyHostToDevice);
    };
    CU(upload(c->d_cell_of_rank, c->cell_of_rank));
    CU(upload(c->d_rank_of_cell, c->rank_of_cell));
    CU(upload(c->d_up_off, up_off));
    CU(upload(c->d_up_idx, up_idx));
    CU(upload(c->d_down, down));
    c->down_pos = down;
    CU(upload(c->d_level_off, c->level_off));
    {   // cell-owner schedule: warps of <= 32 consecutive cells that never straddle a dependency level
        std::vector<int32_t> wb, we, cw(std::max(1, ng), 0);
        for (int l = 0; l < c->nlevels; l++)
            for (int b = c->level_off[l]; b < c->level_off[l + 1]; b += 32) {
                const int e = std::min(b + 32, c->level_off[l + 1]);
                for (int x = b; x < e; x++) cw[x] = (int32_t)wb.size();
                wb.push_back(b);
                we.push_back(e);
            }
        c->owner_nwarps = (int)wb.size();
        CU(upload(c->d_own_warp_begin, wb));
        CU(upload(c->d_own_warp_end, we));
        CU(upload(c->d_own_cell_warp, cw));
        if (c->d_own_progress) cudaFree(c->d_own_progress);
        c->d_own_progress = nullptr;
        CU(cudaMalloc(&c->d_own_progress, sizeof(uint32_t) * (size_t)c->nmember * std::max(1, c->owner_nwarps)));  // (cell-minor only)
        CU(cudaMemset(c->d_own_progress, 0, sizeof(uint32_t) * (size_t)c->nmember * std::max(1, c->owner_nwarps)));
        if (c->d_own_ring) CU(cudaMemset(c->d_own_ring, 0, sizeof(unsigned long long) * 2 * wgk::QBUF_K * (size_t)c->nmember * c->stride));
        c->owner_base = 0;
        if (!c->d_own_abort) {
            CU(cudaMalloc(&c->d_own_abort, sizeof(int32_t)));
            CU(cudaMemset(c->d_own_abort, 0, sizeof(int32_t)));
        }
        c->owner_fits = -1;
    }
    c->have_topology = true;
    drop_graph(c);
    return WGK_OK;
}

int wgk_num_levels(const wgk_ctx *c) { return c && c->have_topology ? c->nlevels : WGK_ERR_STATE; }

int wgk_get_levels(const wgk_ctx *c, int32_t *level) {
    if (!c || !level) return WGK_ERR_ARG;
    if (!c->have_topology) return WGK_ERR_STATE;
    for (int n = 0; n < c->ncell; n++) level[n] = c->level_of_rank[c->rank_of_cell[n]];
    return WGK_OK;
}

int wgk_get_device_order(const wgk_ctx *c, int32_t *rank_of_cell) {
    if (!c || !rank_of_cell) return WGK_ERR_ARG;
    if (!c->have_topology) return WGK_ERR_STATE;
    memcpy(rank_of_cell, c->rank_of_cell.data(), sizeof(int32_t) * c->ncell);
    return WGK_OK;
}

int64_t wgk_cell_stride(const wgk_ctx *c) { return c ? (c->mm ? c->mpad : 1) : 0; }
int64_t wgk_member_stride(const wgk_ctx *c) { return c ? (c->mm ? 1 : c->stride) : 0; }
int wgk_layout(const wgk_ctx *c) { return c ? (c->mm ? 1 : 0) : WGK_ERR_ARG; }
int wgk_schedule(const wgk_ctx *c) {
    return c ? ((c->whole_day ? 1 : 0) | (!c->whole_day && c->level_tasks ? 2 : 0) | (c->owner_mode == 1 ? 4 : 0)) : WGK_ERR_ARG;
}

// ---------------------------------------------------------------------------------------
// fields
// ---------------------------------------------------------------------------------------
int wgk_field_id(const char *name) {
    if (!name) return WGK_ERR_ARG;
    if (strcmp(name, "params") == 0) return WGK_F_params;
    for (int f = 0; f < WGK_F_COUNT; f++)
        if (strcmp(name, kFields[f].name) == 0) return f;
    return WGK_ERR_ARG;
}

int wgk_field_info(int field, const char **name, const char **dtype, int64_t *count, int ncell) {
    if (field == WGK_F_params) {
        if (name) *name = "params";
        if (dtype) *dtype = "f64";
        if (count) *count = (int64_t)WGK_NPARAM * ncell;
        return WGK_SCOPE_PSET;
    }
    if (field < 0 || field >= WGK_F_COUNT) return WGK_ERR_ARG;
    if (name) *name = kFields[field].name;
    if (dtype) *dtype = kFields[field].dtype;
    if (count) *count = kFields[field].scope == WGK_SCOPE_TABLE ? WGK_NLCT : (int64_t)ncell * kFields[field].bands;
    return kFields[field].scope;
}

static int set_field_raw(wgk_ctx *c, int f, int index, const void *host, size_t bytes) {
    const FieldInfo &fi = kFields[f];
    if (index < 0 || (size_t)index >= field_index_count(c, f)) return fail(c, WGK_ERR_ARG, "index %d out of range for field %s", index, fi.name);
    if (!*field_slot(c, f)) return fail(c, WGK_ERR_STATE, "field %s exists only with wgk_options.subtract_use > 0", fi.name);
    const size_t row_elems = field_row_elems(c, f);
    char *dst = (char *)*field_slot(c, f) + (size_t)index * row_elems * fi.elsize;
    if (fi.scope == WGK_SCOPE_TABLE) {
        if (bytes != (size_t)WGK_NLCT * fi.elsize) return fail(c, WGK_ERR_ARG, "field %s expects %zu bytes, got %zu", fi.name, (size_t)WGK_NLCT * fi.elsize, bytes);
        CU(cudaMemcpyAsync(dst, host, bytes, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        return WGK_OK;
    }
    const size_t want = (size_t)c->ncell * fi.bands * fi.elsize;
    if (bytes != want) return fail(c, WGK_ERR_ARG, "field %s expects %zu bytes, got %zu", fi.name, want, bytes);
    if (!c->have_topology) return fail(c, WGK_ERR_STATE, "wgk_set_topology must precede wgk_set_field (device order = routing order)");
    if (f == WGK_F_arid) {  // daily.cpp:345-348: any value other than 0/1 aborts the reference
        const int16_t *v = (const int16_t *)host;
        for (int n = 0; n < c->ncell; n++)
            if (v[n] != 0 && v[n] != 1) return fail(c, WGK_ERR_ARG, "Invalid value for Arid/humid index: %d (cell %d)", (int)v[n], n + 1);
    }
    const size_t row_bytes = row_elems * fi.elsize;
    int rc = ensure_stage(c, row_bytes);
    if (rc) return rc;
    memset(c->h_stage, 0, row_bytes);
    const int es = fi.elsize, nb = fi.bands, ng = c->ncell, st = c->stride;
    const char *src = (const char *)host;
    char *tmp = (char *)c->h_stage;
    for (int r = 0; r < ng; r++) {
        const size_t n = (size_t)c->cell_of_rank[r];
        for (int b = 0; b < nb; b++) memcpy(tmp + ((size_t)b * st + r) * es, src + (n * nb + b) * es, es);
    }
    const int pad = index_pad(c, f);
    if (pad > 1) {  // member-minor: the dense row [band][cell] goes to a device staging row and is scattered with stride `pad`
        rc = ensure_dstage(c, row_bytes);
        if (rc) return rc;
        CU(cudaMemcpyAsync(c->d_stage, tmp, row_bytes, cudaMemcpyHostToDevice, c->stream));
        launch_strided(c, (char *)*field_slot(c, f) + (size_t)index * es, (size_t)pad, (const char *)c->d_stage, 1, row_elems, es);
        CU(cudaGetLastError());
    } else {
        CU(cudaMemcpyAsync(dst, tmp, row_bytes, cudaMemcpyHostToDevice, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    if (fi.scope != WGK_SCOPE_MEMBER) c->derived_dirty = true;
    if (f == WGK_F_snow_bands) c->member_dirty = true;
    return WGK_OK;
}

int wgk_set_field(wgk_ctx *c, int field, int index, const void *host, size_t bytes) {
    if (!c || !host) return WGK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    if (field == WGK_F_params) {
        const size_t want = (size_t)WGK_NPARAM * c->ncell * sizeof(double);
        if (bytes != want) return fail(c, WGK_ERR_ARG, "params expects %zu bytes, got %zu", want, bytes);
        const double *p = (const double *)host;
        for (const ParamMap &pm : kParamMap) {
            int rc = set_field_raw(c, pm.field, index, p + (size_t)pm.k * c->ncell, (size_t)c->ncell * sizeof(double));
            if (rc) return rc;
        }
        // active lake / wetland depth in km (routing.cpp:5613-5628, conversion.h)
        std::vector<double> d(c->ncell);
        for (int n = 0; n < c->ncell; n++) d[n] = p[(size_t)5 * c->ncell + n] * 0.001;
        int rc = set_field_raw(c, WGK_F_lake_depth_active, index, d.data(), d.size() * sizeof(double));
        if (rc) return rc;
        for (int n = 0; n < c->ncell; n++) d[n] = p[(size_t)6 * c->ncell + n] * 0.001;
        return set_field_raw(c, WGK_F_wetl_depth_active, index, d.data(), d.size() * sizeof(double));
    }
    if (field < 0 || field >= WGK_F_COUNT) return fail(c, WGK_ERR_ARG, "unknown field %d", field);
    return set_field_raw(c, field, index, host, bytes);
}

int wgk_get_field(wgk_ctx *c, int f, int index, void *host, size_t bytes) {
    if (!c || !host) return WGK_ERR_ARG;
    if (f < 0 || f >= WGK_F_COUNT) return fail(c, WGK_ERR_ARG, "unknown field %d", f);
    CU(cudaSetDevice(c->device));
    const FieldInfo &fi = kFields[f];
    if (index < 0 || (size_t)index >= field_index_count(c, f)) return fail(c, WGK_ERR_ARG, "index %d out of range for field %s", index, fi.name);
    if (!*field_slot(c, f)) return fail(c, WGK_ERR_STATE, "field %s exists only with wgk_options.subtract_use > 0", fi.name);
    const size_t row_elems = field_row_elems(c, f);
    const char *src = (const char *)*field_slot(c, f) + (size_t)index * row_elems * fi.elsize;
    if (fi.scope == WGK_SCOPE_TABLE) {
        if (bytes != (size_t)WGK_NLCT * fi.elsize) return fail(c, WGK_ERR_ARG, "field %s: wrong size", fi.name);
        CU(cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        return WGK_OK;
    }
    const size_t want = (size_t)c->ncell * fi.bands * fi.elsize;
    if (bytes != want) return fail(c, WGK_ERR_ARG, "field %s expects %zu bytes, got %zu", fi.name, want, bytes);
    if (!c->have_topology) return fail(c, WGK_ERR_STATE, "no topology");
    const size_t row_bytes = row_elems * fi.elsize;
    int rc = ensure_stage(c, row_bytes);
    if (rc) return rc;
    const int pad = index_pad(c, f);
    if (pad > 1) {  // member-minor: gather the strided elements of this index into a dense device row first
        rc = ensure_dstage(c, row_bytes);
        if (rc) return rc;
        launch_strided(c, (char *)c->d_stage, 1, (const char *)*field_slot(c, f) + (size_t)index * fi.elsize, (size_t)pad, row_elems, fi.elsize);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(c->h_stage, c->d_stage, row_bytes, cudaMemcpyDeviceToHost, c->stream));
    } else {
        CU(cudaMemcpyAsync(c->h_stage, src, row_bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    const int es = fi.elsize, nb = fi.bands, ng = c->ncell, st = c->stride;
    const char *tmp = (const char *)c->h_stage;
    char *dst = (char *)host;
    for (int r = 0; r < ng; r++) {
        const size_t n = (size_t)c->cell_of_rank[r];
        for (int b = 0; b < nb; b++) memcpy(dst + (n * nb + b) * es, tmp + ((size_t)b * st + r) * es, es);
    }
    return WGK_OK;
}

int wgk_set_member_pset(wgk_ctx *c, int member, int pset) {
    if (!c || member < 0 || member >= c->nmember || pset < 0 || pset >= c->npset) return WGK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    c->member_pset[member] = pset;
    CU(cudaMemcpyAsync(c->d_member_pset + member, &c->member_pset[member], sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return WGK_OK;
}

void *wgk_device_ptr(wgk_ctx *c, int f, int member) {
    if (!c || f < 0 || f >= WGK_F_COUNT) return nullptr;
    if (member < 0 || (size_t)member >= field_index_count(c, f) || !*field_slot(c, f)) return nullptr;
    if (f == WGK_F_snow_bands) c->member_dirty = true;  // the caller may write the bands on the device
    return (char *)*field_slot(c, f) + (size_t)member * (index_pad(c, f) > 1 ? 1 : field_row_elems(c, f)) * kFields[f].elsize;
}

// ---------------------------------------------------------------------------------------
// forcing
// ---------------------------------------------------------------------------------------
int wgk_forcing_reserve(wgk_ctx *c, int nslots, int per_member) {
    if (!c || nslots <= 0) return WGK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->copy_stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->d_forcing) cudaFree(c->d_forcing);
    c->d_forcing = nullptr;
    const size_t n = (size_t)nslots * (per_member ? c->mpad : 1) * c->stride;
    cudaError_t e = cudaMalloc(&c->d_forcing, n * sizeof(float4));
    if (e != cudaSuccess) return fail(c, WGK_ERR_NOMEM, "cudaMalloc forcing (%zu bytes): %s", n * sizeof(float4), cudaGetErrorString(e));
    CU(cudaMemsetAsync(c->d_forcing, 0, n * sizeof(float4), c->stream));
    c->forcing_nslots = nslots;
    c->forcing_per_member = per_member ? 1 : 0;
    drop_graph(c);
    return WGK_OK;
}

static int set_forcing_impl(wgk_ctx *c, int slot0, int ndays, int member, const float *prec, const float *temp,
                            const float *sw, const float *lw, int stride, bool big_endian) {
    if (!c || !prec || !temp || !sw || !lw || ndays <= 0 || stride < ndays) return WGK_ERR_ARG;
    if (!c->have_topology) return fail(c, WGK_ERR_STATE, "wgk_set_topology must precede wgk_set_forcing");
    if (!c->d_forcing) {
        int rc = wgk_forcing_reserve(c, std::max(31, slot0 + ndays), member >= 0);
        if (rc) return rc;
    }
    if (slot0 < 0 || slot0 + ndays > c->forcing_nslots) return fail(c, WGK_ERR_ARG, "forcing slots %d..%d exceed the %d reserved", slot0, slot0 + ndays - 1, c->forcing_nslots);
    if ((member >= 0) != (c->forcing_per_member != 0))
        return fail(c, WGK_ERR_ARG, "forcing was reserved %s", c->forcing_per_member ? "per member" : "shared");
    if (member >= c->nmember) return WGK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    const size_t elems = (size_t)c->ncell * stride;
    if (c->fstage_elems < elems) {
        if (c->d_fstage) cudaFree(c->d_fstage);
        c->d_fstage = nullptr;
        CU(cudaMalloc(&c->d_fstage, 4 * elems * sizeof(float)));
        c->fstage_elems = elems;
    }
    // wait for the steps that still read these slots (and drop the records of finished steps)
    for (size_t k = 0; k < c->slot_uses.size();) {
        wgk_ctx::SlotUse &u = c->slot_uses[k];
        if (cudaEventQuery(u.ev) == cudaSuccess) {
            cudaEventDestroy(u.ev);
            c->slot_uses.erase(c->slot_uses.begin() + k);
            continue;
        }
        if (u.lo < slot0 + ndays && slot0 < u.lo + u.n) CU(cudaStreamWaitEvent(c->copy_stream, u.ev, 0));
        k++;
    }
    float *dP = c->d_fstage, *dT = dP + elems, *dS = dT + elems, *dL = dS + elems;
    cudaStream_t cs = c->copy_stream;
    CU(cudaMemcpyAsync(dP, prec, elems * sizeof(float), cudaMemcpyHostToDevice, cs));
    CU(cudaMemcpyAsync(dT, temp, elems * sizeof(float), cudaMemcpyHostToDevice, cs));
    CU(cudaMemcpyAsync(dS, sw, elems * sizeof(float), cudaMemcpyHostToDevice, cs));
    CU(cudaMemcpyAsync(dL, lw, elems * sizeof(float), cudaMemcpyHostToDevice, cs));
    // element (slot, member, cell) of the device forcing: WgkParams::fi
    const int F = c->forcing_per_member ? c->mpad : 1;
    const size_t pitch = (size_t)F * c->stride;
    const bool strided = c->mm && c->forcing_per_member;
    const size_t cell_stride = strided ? (size_t)c->mpad : 1;
    float4 *dst = c->d_forcing + (size_t)slot0 * pitch + (size_t)(c->forcing_per_member ? member : 0) * (strided ? 1 : c->stride);
    dim3 block(128), grid((c->ncell + 127) / 128, std::min(ndays, 31));
    if (big_endian) wgk::k_forcing_pack<true><<<grid, block, 0, cs>>>(dst, dP, dT, dS, dL, c->d_cell_of_rank, c->ncell, ndays, stride, pitch, cell_stride);
    else wgk::k_forcing_pack<false><<<grid, block, 0, cs>>>(dst, dP, dT, dS, dL, c->d_cell_of_rank, c->ncell, ndays, stride, pitch, cell_stride);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev_forcing, cs));
    c->forcing_pending = true;
    return WGK_OK;
}

int wgk_set_forcing(wgk_ctx *c, int slot0, int ndays, int member, const float *prec, const float *temp,
                    const float *sw, const float *lw, int stride) {
    return set_forcing_impl(c, slot0, ndays, member, prec, temp, sw, lw, stride, false);
}

int wgk_set_forcing_unf(wgk_ctx *c, int slot0, int ndays, int member, const void *prec, const void *temp,
                        const void *sw, const void *lw, int stride) {
    return set_forcing_impl(c, slot0, ndays, member, (const float *)prec, (const float *)temp, (const float *)sw, (const float *)lw, stride, true);
}

// ---------------------------------------------------------------------------------------
// hot path
// ---------------------------------------------------------------------------------------
// a step is ordered after every forcing upload issued before it ...
static int forcing_before_step(wgk_ctx *c) {
    if (c->forcing_pending) {
        CU(cudaStreamWaitEvent(c->stream, c->ev_forcing, 0));
        c->forcing_pending = false;
    }
    return WGK_OK;
}
// ... and remembers which slots it reads until it has finished
static int forcing_after_step(wgk_ctx *c, int slot0, int ndays) {
    for (size_t k = 0; k < c->slot_uses.size();) {  // forget the steps that have finished
        if (cudaEventQuery(c->slot_uses[k].ev) == cudaSuccess) {
            cudaEventDestroy(c->slot_uses[k].ev);
            c->slot_uses.erase(c->slot_uses.begin() + k);
        } else {
            k++;
        }
    }
    wgk_ctx::SlotUse u;
    u.lo = slot0;
    u.n = ndays;
    if (slot0 + ndays > c->forcing_nslots) { u.lo = 0; u.n = c->forcing_nslots; }  // the slots wrap around
    CU(cudaEventCreateWithFlags(&u.ev, cudaEventDisableTiming));
    CU(cudaEventRecord(u.ev, c->stream));
    c->slot_uses.push_back(u);
    return WGK_OK;
}

static int fill_calendar(wgk_ctx *c, int day, int month, int dom, int slot, int ndays, bool record = false) {
    if (day < 1 || day > 365 || month < 0 || month > 11 || dom < 1 || dom > 31) return fail(c, WGK_ERR_ARG, "bad date day=%d month=%d day_in_month=%d", day, month, dom);
    if (slot < 0 || slot >= c->forcing_nslots) return fail(c, WGK_ERR_ARG, "forcing slot %d not reserved", slot);
    if (ndays < 1 || ndays > MAX_CALL_DAYS) return fail(c, WGK_ERR_ARG, "1..%d days per call", MAX_CALL_DAYS);
    // station record: the rows of this call follow those of the calls before it; a call that does not fit into the
    // remaining rows restarts the record at row 0 (wgk.h)
    int rec_base = 0;
    if (c->d_record && record) {
        if (c->record_rows + ndays > c->record_max_days) c->record_rows = 0;
        rec_base = c->record_rows;
        c->record_rows = std::min(c->record_max_days, c->record_rows + ndays);
    }
    wgk::k_fill_calendar<<<1, 1, 0, c->stream>>>(c->d_cal_days, c->d_cal, day, month, dom, slot, ndays, c->forcing_nslots, rec_base);
    c->launches++;
    return WGK_OK;
}

// make the discharge of day offset `d` of the finished call the value of the "discharge" field
static int publish_discharge(wgk_ctx *c, int d) {
    const size_t n = (size_t)c->mpad * c->stride;
    CU(cudaMemcpyAsync(c->arrays.discharge, c->d_qbuf + (size_t)(d % wgk::QBUF_K) * n, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return WGK_OK;
}

int wgk_vertical_day(wgk_ctx *c, int day, int month, int dom, int slot) {
    int rc = check_ready(c);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    rc = fill_calendar(c, day, month, dom, slot, 1);
    if (rc) return rc;
    rc = ensure_derived(c);
    if (rc) return rc;
    rc = forcing_before_step(c);
    if (rc) return rc;
    c->launches += enqueue_vertical(c, make_params(c), 0);
    CU(cudaGetLastError());
    return forcing_after_step(c, slot, 1);
}

int wgk_routing_day(wgk_ctx *c, int day, int month, int dom) {
    int rc = check_ready(c);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    rc = fill_calendar(c, day, month, dom, 0, 1);  // the routing kernels do not read the forcing slot
    if (rc) return rc;
    rc = ensure_derived(c);
    if (rc) return rc;
    c->launches += enqueue_routing(c, make_params(c), 0);
    CU(cudaGetLastError());
    if (c->month_acc) c->month_days += 1;
    return publish_discharge(c, 0);
}

// development aid (not part of include/wgk.h): per-warp cycle counts of the last cell-owner call, [nwarps][4]
// = {vertical + local routing, hand-off waits, river + release, post-pass}, and the first cell / level width of each warp
extern "C" int wgk_debug_owner_timing(wgk_ctx *c, long long *out, int32_t *warp_begin, int max_warps) {
    if (!c || !c->d_own_timing) return -1;
    const int n = std::min(max_warps, c->owner_nwarps);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
    if (cudaMemcpy(out, c->d_own_timing, sizeof(long long) * 4 * n, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (warp_begin && cudaMemcpy(warp_begin, c->d_own_warp_begin, sizeof(int32_t) * n, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return c->owner_nwarps;
}

int wgk_update_land_area_frac(wgk_ctx *c) {
    // routingClass::updateLandAreaFrac is fused into the routing post-pass (prev <- cur <- next
    // per cell); the entry point exists so that the three-call sequence of
    // integrateWGHM.cpp:779-798 maps one to one.
    return check_ready(c);
}

int wgk_step_days(wgk_ctx *c, int day, int month, int dom, int slot0, int ndays) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (ndays <= 0) return WGK_OK;
    CU(cudaSetDevice(c->device));
    rc = fill_calendar(c, day, month, dom, slot0, ndays, true);
    if (rc) return rc;
    rc = ensure_derived(c);
    if (rc) return rc;
    rc = forcing_before_step(c);
    if (rc) return rc;
    const WgkParams p = make_params(c);
    if (c->owner_mode == 1 && !owner_usable(c))
        return fail(c, WGK_ERR_STATE, "WGK_DAY_SCHEDULE=owner: %d warps x %d members are not co-resident on this GPU", c->owner_nwarps, c->nmember);
    bool owner = false;
    if ((c->opt.use_graph || c->owner_mode == 1) && owner_usable(c)) {
        rc = launch_owner(c, p, ndays);
        if (rc) return rc;
        owner = true;  // the discharge field is written by the kernel itself
    } else if (c->opt.use_graph) {
        auto it = c->graphs.find(ndays);
        if (it == c->graphs.end()) {
            cudaGraphExec_t ex;
            int nn = 0;
            if (c->whole_day) {  // linear chain, recorded by stream capture
                cudaGraph_t g;
                CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                nn = enqueue_whole_days(c, p, ndays);
                CU(cudaStreamEndCapture(c->stream, &g));
                cudaError_t e = cudaGraphInstantiate(&ex, g, 0);
                cudaGraphDestroy(g);
                if (e != cudaSuccess) return fail(c, WGK_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
            } else {
                rc = build_wavefront_graph(c, p, ndays, &ex, &nn);
                if (rc) return rc;
            }
            c->graphs[ndays] = ex;
            c->graph_nodes[ndays] = nn;
            it = c->graphs.find(ndays);
        }
        CU(cudaGraphLaunch(it->second, c->stream));
        c->launches += c->graph_nodes[ndays];
    } else {
        c->launches += c->whole_day ? enqueue_whole_days(c, p, ndays) : enqueue_wavefront_serial(c, p, ndays);
    }
    CU(cudaGetLastError());
    if (!owner) {
        rc = publish_discharge(c, ndays - 1);
        if (rc) return rc;
    }
    if (c->month_acc) c->month_days += ndays;
    return forcing_after_step(c, slot0, ndays);
}

// ---------------------------------------------------------------------------------------
// EnKF state bridge
// ---------------------------------------------------------------------------------------
int wgk_month_begin(wgk_ctx *c) {
    if (!c) return WGK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(c->arrays.mon_acc, 0, (size_t)c->mpad * 7 * c->stride * sizeof(double), c->stream));
    if (!c->month_acc) drop_graph(c);  // the flag is part of the kernel parameters baked into the graphs
    c->month_acc = true;
    c->month_days = 0;
    return WGK_OK;
}

namespace {
// device copies of the region's cell list (as device positions) and of up to three [ncells][10] host arrays
struct BridgeBuffers {
    int32_t *pos = nullptr;
    double *d[3] = {nullptr, nullptr, nullptr};
    ~BridgeBuffers() {
        cudaFree(pos);
        for (double *x : d) cudaFree(x);
    }
};
int bridge_upload(wgk_ctx *c, BridgeBuffers &b, const int32_t *cells, int ncells, const double *h0, const double *h1, const double *h2) {
    if (!c->have_topology) return fail(c, WGK_ERR_STATE, "no topology");
    std::vector<int32_t> pos(ncells);
    for (int k = 0; k < ncells; k++) {
        if (cells[k] < 0 || cells[k] >= c->ncell) return fail(c, WGK_ERR_ARG, "cell %d out of range", cells[k]);
        pos[k] = c->rank_of_cell[cells[k]];
    }
    CU(cudaMalloc(&b.pos, sizeof(int32_t) * std::max(1, ncells)));
    CU(cudaMemcpyAsync(b.pos, pos.data(), sizeof(int32_t) * ncells, cudaMemcpyHostToDevice, c->stream));
    const double *h[3] = {h0, h1, h2};
    for (int k = 0; k < 3; k++) {
        CU(cudaMalloc(&b.d[k], sizeof(double) * 10 * std::max(1, ncells)));
        if (h[k]) CU(cudaMemcpyAsync(b.d[k], h[k], sizeof(double) * 10 * ncells, cudaMemcpyHostToDevice, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));  // pos / h* are host temporaries of the caller
    return WGK_OK;
}
}  // namespace

int wgk_state_vector(wgk_ctx *c, int member, int kind, const int32_t *cells, int ncells, const double *mean_field, double *out) {
    if (!c || !cells || !out || ncells < 0 || member < 0 || member >= c->nmember || (kind != 0 && kind != 1)) return WGK_ERR_ARG;
    if (kind == 0 && (!c->month_acc || c->month_days <= 0)) return fail(c, WGK_ERR_STATE, "wgk_month_begin and at least one stepped day must precede a monthly state vector");
    if (ncells == 0) return WGK_OK;
    CU(cudaSetDevice(c->device));
    BridgeBuffers b;
    int rc = bridge_upload(c, b, cells, ncells, mean_field, nullptr, nullptr);
    if (rc) return rc;
    WGK_K(c, k_state_vector)<<<(ncells + 127) / 128, 128, 0, c->stream>>>(make_params(c), member, kind, c->month_days, b.pos, ncells,
                                                                     mean_field ? b.d[0] : nullptr, b.d[1]);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, b.d[1], sizeof(double) * 10 * ncells, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return WGK_OK;
}

int wgk_enkf_update(wgk_ctx *c, int member, const int32_t *cells, int ncells, const double *field, const double *prediction,
                    const double *mean_field) {
    if (!c || !cells || !field || !prediction || !mean_field || ncells < 0 || member < 0 || member >= c->nmember) return WGK_ERR_ARG;
    if (!c->month_acc || c->month_days <= 0) return fail(c, WGK_ERR_STATE, "wgk_month_begin and the month's days must precede wgk_enkf_update");
    if (ncells == 0) return WGK_OK;
    CU(cudaSetDevice(c->device));
    BridgeBuffers b;
    int rc = bridge_upload(c, b, cells, ncells, field, prediction, mean_field);
    if (rc) return rc;
    WGK_K(c, k_enkf_update)<<<(ncells + 127) / 128, 128, 0, c->stream>>>(make_params(c), member, c->month_days, b.pos, ncells, b.d[0], b.d[1], b.d[2]);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    c->member_dirty = true;  // the band state changed: s_snowfree is recomputed before the next step
    return WGK_OK;
}

int wgk_get_day_state(wgk_ctx *c, int member, double *out) {
    if (!c || !out || member < 0 || member >= c->nmember) return WGK_ERR_ARG;
    if (!c->have_topology) return fail(c, WGK_ERR_STATE, "no topology");
    CU(cudaSetDevice(c->device));
    const size_t bytes = sizeof(double) * 7 * (size_t)c->ncell;
    int rc = ensure_dstage(c, bytes);
    if (rc) return rc;
    WGK_K(c, k_pack_day_state)<<<(c->ncell + 127) / 128, 128, 0, c->stream>>>(make_params(c), member, c->d_rank_of_cell, (double *)c->d_stage);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, c->d_stage, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return WGK_OK;
}

// ---------------------------------------------------------------------------------------
// ensemble statistics (the only exchange between ranks, once per assimilation cycle)
// ---------------------------------------------------------------------------------------
int wgk_ensemble_moments(wgk_ctx *c, int kind, const int32_t *cells, int ncells, void **d_sum, void **d_sumsq) {
    if (!c || (kind != 0 && kind != 1) || ncells < 0 || !d_sum || !d_sumsq) return WGK_ERR_ARG;
    if (!c->have_topology) return fail(c, WGK_ERR_STATE, "no topology");
    if (kind == 0 && (!c->month_acc || c->month_days <= 0)) return fail(c, WGK_ERR_STATE, "wgk_month_begin and at least one stepped day must precede monthly moments");
    CU(cudaSetDevice(c->device));
    const int n = cells ? ncells : c->ncell;
    if (n == 0) return fail(c, WGK_ERR_ARG, "empty cell list");
    if (c->mom_cap < (size_t)n) {
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_mom_sum); cudaFree(c->d_mom_sumsq); cudaFree(c->d_mom_pos);
        c->d_mom_sum = c->d_mom_sumsq = nullptr;
        c->d_mom_pos = nullptr;
        c->mom_cap = 0;
        // one allocation for both, so that a single all-reduce of 2 * n * 10 doubles covers sum and sumsq
        CU(cudaMalloc(&c->d_mom_sum, sizeof(double) * 20 * (size_t)n));
        c->d_mom_sumsq = nullptr;
        CU(cudaMalloc(&c->d_mom_pos, sizeof(int32_t) * (size_t)n));
        c->mom_cap = (size_t)n;
    }
    double *sum = c->d_mom_sum, *sumsq = c->d_mom_sum + (size_t)10 * n;
    if (cells) {
        std::vector<int32_t> pos(n);
        for (int k = 0; k < n; k++) {
            if (cells[k] < 0 || cells[k] >= c->ncell) return fail(c, WGK_ERR_ARG, "cell %d out of range", cells[k]);
            pos[k] = c->rank_of_cell[cells[k]];
        }
        CU(cudaMemcpyAsync(c->d_mom_pos, pos.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));  // pos is a host temporary
    }
    WGK_K(c, k_ensemble_moments)<<<(n + 127) / 128, 128, 0, c->stream>>>(make_params(c), kind, c->month_days, cells ? c->d_mom_pos : nullptr,
                                                                    c->d_cell_of_rank, n, sum, sumsq);
    c->launches++;
    CU(cudaGetLastError());
    c->mom_ncells = n;
    *d_sum = sum;
    *d_sumsq = sumsq;
    return WGK_OK;
}

int wgk_moments_finish(wgk_ctx *c, int nmember_total, double *mean, double *var) {
    if (!c || nmember_total <= 0) return WGK_ERR_ARG;
    if (!c->d_mom_sum || c->mom_ncells <= 0) return fail(c, WGK_ERR_STATE, "wgk_ensemble_moments must be called first");
    CU(cudaSetDevice(c->device));
    const size_t n = (size_t)10 * c->mom_ncells;
    double *sum = c->d_mom_sum, *sumsq = c->d_mom_sum + n;
    wgk::k_moments_finish<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(sum, sumsq, n, (double)nmember_total);
    c->launches++;
    CU(cudaGetLastError());
    if (mean) CU(cudaMemcpyAsync(mean, sum, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (var) CU(cudaMemcpyAsync(var, sumsq, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return WGK_OK;
}

// ---------------------------------------------------------------------------------------
// device-side replication and uniform parameters (set-up of large ensembles / parameter sweeps)
// ---------------------------------------------------------------------------------------
int wgk_copy_index(wgk_ctx *c, int scope, int src, int dst) {
    if (!c || (scope != WGK_SCOPE_PSET && scope != WGK_SCOPE_MEMBER)) return WGK_ERR_ARG;
    const int rows = scope == WGK_SCOPE_PSET ? c->npset : c->nmember;
    if (src < 0 || src >= rows || dst < 0 || dst >= rows) return fail(c, WGK_ERR_ARG, "index out of range (%d -> %d of %d)", src, dst, rows);
    if (src == dst) return WGK_OK;
    CU(cudaSetDevice(c->device));
    for (int f = 0; f < WGK_F_COUNT; f++) {
        if (kFields[f].scope != scope || !*field_slot(c, f)) continue;
        const size_t bytes = field_row_elems(c, f) * kFields[f].elsize;
        char *base = (char *)*field_slot(c, f);
        const int pad = index_pad(c, f);
        if (pad > 1) {
            const int es = kFields[f].elsize;
            launch_strided(c, base + (size_t)dst * es, (size_t)pad, base + (size_t)src * es, (size_t)pad, field_row_elems(c, f), es);
        } else {
            CU(cudaMemcpyAsync(base + (size_t)dst * bytes, base + (size_t)src * bytes, bytes, cudaMemcpyDeviceToDevice, c->stream));
        }
    }
    if (scope == WGK_SCOPE_PSET) c->derived_dirty = true;
    // (s_snowfree is a member field and is copied with the bands it describes)
    return WGK_OK;
}

int wgk_fill_field(wgk_ctx *c, int field, int index, double value) {
    if (!c) return WGK_ERR_ARG;
    if (field < 0 || field >= WGK_F_COUNT) return fail(c, WGK_ERR_ARG, "unknown field %d", field);
    const FieldInfo &fi = kFields[field];
    if (strcmp(fi.dtype, "f64") != 0 || fi.bands != 1 || fi.scope == WGK_SCOPE_TABLE)
        return fail(c, WGK_ERR_ARG, "wgk_fill_field: %s is not a per-cell f64 field", fi.name);
    if (index < 0 || (size_t)index >= field_index_count(c, field)) return fail(c, WGK_ERR_ARG, "index %d out of range for field %s", index, fi.name);
    CU(cudaSetDevice(c->device));
    const int pad = index_pad(c, field);
    double *dst = (double *)*field_slot(c, field) + (size_t)index * (pad > 1 ? 1 : c->stride);
    wgk::k_fill_f64<<<(c->ncell + 255) / 256, 256, 0, c->stream>>>(dst, c->ncell, (size_t)pad, value);
    c->launches++;
    CU(cudaGetLastError());
    if (fi.scope != WGK_SCOPE_MEMBER) c->derived_dirty = true;
    return WGK_OK;
}

// ---------------------------------------------------------------------------------------
// diagnostics
// ---------------------------------------------------------------------------------------
int wgk_total_storage_km3(wgk_ctx *c, int member, double *out) {
    if (!c || !out || member < 0 || member >= c->nmember) return WGK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    const int nblk = 128;
    WGK_K(c, k_total_storage)<<<nblk, 256, 0, c->stream>>>(make_params(c), member, c->d_partial);
    c->launches++;
    double part[128];
    CU(cudaMemcpyAsync(part, c->d_partial, sizeof(double) * nblk, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    double s = 0.;
    for (int i = 0; i < nblk; i++) s += part[i];
    *out = s;
    return WGK_OK;
}

int wgk_record_cells(wgk_ctx *c, const int32_t *cells, int ncells, int max_days) {
    if (!c || ncells < 0 || max_days < 0) return WGK_ERR_ARG;
    if (!c->have_topology) return fail(c, WGK_ERR_STATE, "no topology");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (c->d_record) cudaFree(c->d_record);
    if (c->d_record_cells) cudaFree(c->d_record_cells);
    if (c->d_rec_head) cudaFree(c->d_rec_head);
    if (c->d_rec_next) cudaFree(c->d_rec_next);
    c->d_record = nullptr;
    c->d_record_cells = nullptr;
    c->d_rec_head = nullptr;
    c->d_rec_next = nullptr;
    c->nrec = 0;
    c->record_max_days = 0;
    if (ncells > 0 && max_days > 0) {
        std::vector<int32_t> ranks(ncells);
        for (int k = 0; k < ncells; k++) {
            if (cells[k] < 0 || cells[k] >= c->ncell) return fail(c, WGK_ERR_ARG, "record cell %d out of range", cells[k]);
            ranks[k] = c->rank_of_cell[cells[k]];
        }
        CU(cudaMalloc(&c->d_record_cells, sizeof(int32_t) * ncells));
        CU(cudaMemcpy(c->d_record_cells, ranks.data(), sizeof(int32_t) * ncells, cudaMemcpyHostToDevice));
        // the same list per cell, for the cell-owner schedule (a cell may be recorded more than once)
        std::vector<int32_t> head(c->ncell, -1), next(ncells, -1);
        for (int k = ncells - 1; k >= 0; k--) {
            next[k] = head[ranks[k]];
            head[ranks[k]] = k;
        }
        CU(cudaMalloc(&c->d_rec_head, sizeof(int32_t) * c->ncell));
        CU(cudaMalloc(&c->d_rec_next, sizeof(int32_t) * ncells));
        CU(cudaMemcpy(c->d_rec_head, head.data(), sizeof(int32_t) * c->ncell, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_rec_next, next.data(), sizeof(int32_t) * ncells, cudaMemcpyHostToDevice));
        const size_t n = (size_t)max_days * c->nmember * ncells;
        CU(cudaMalloc(&c->d_record, n * sizeof(double)));
        CU(cudaMemset(c->d_record, 0, n * sizeof(double)));
        c->nrec = ncells;
        c->record_max_days = max_days;
    }
    c->record_rows = 0;
    drop_graph(c);
    return WGK_OK;
}

int wgk_record_rewind(wgk_ctx *c) {
    if (!c) return WGK_ERR_ARG;
    c->record_rows = 0;
    return WGK_OK;
}

int wgk_get_record(wgk_ctx *c, int member, double *out, int ndays) {
    if (!c || !out || member < 0 || member >= c->nmember || ndays < 0) return WGK_ERR_ARG;
    if (!c->d_record) return fail(c, WGK_ERR_STATE, "wgk_record_cells was not called");
    if (ndays > c->record_rows) return fail(c, WGK_ERR_ARG, "the station record holds %d days (of at most %d), %d requested", c->record_rows, c->record_max_days, ndays);
    CU(cudaSetDevice(c->device));
    const size_t pitch = (size_t)c->nmember * c->nrec;
    CU(cudaMemcpy2DAsync(out, (size_t)c->nrec * sizeof(double), c->d_record + (size_t)member * c->nrec, pitch * sizeof(double),
                         (size_t)c->nrec * sizeof(double), ndays, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return WGK_OK;
}

int wgk_profile_day(wgk_ctx *c, int day, int month, int dom, int slot, float ms[6]) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (!ms) return WGK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    rc = fill_calendar(c, day, month, dom, slot, 1);
    if (rc) return rc;
    rc = ensure_derived(c);
    if (rc) return rc;
    rc = forcing_before_step(c);
    if (rc) return rc;
    cudaEvent_t ev[6];
    for (auto &e : ev) CU(cudaEventCreate(&e));
    const WgkParams p = make_params(c);
    const dim3 block = cell_block(c, 128), grid = cell_grid(c, c->ncell, 128);
    CU(cudaEventRecord(ev[0], c->stream));
    launch_vertical(c, p, 0);
    CU(cudaEventRecord(ev[1], c->stream));
    WGK_KW(c, k_route_local)<<<grid, block, 0, c->stream>>>(p, 0);
    CU(cudaEventRecord(ev[2], c->stream));
    int n = 3;
    for (int l = 0; l < c->tail_level0; l++) {
        const int cnt = c->level_off[l + 1] - c->level_off[l];
        const dim3 g = cell_grid(c, cnt, 128);
        WGK_KW(c, k_route_level)<<<g, block, 0, c->stream>>>(p, 0, l);
        n++;
    }
    CU(cudaEventRecord(ev[3], c->stream));
    if (c->tail_level0 < c->nlevels) {
        (c->opt.subtract_use > 0 ? wgk_wu::k_route_tail : wgk::k_route_tail)<<<c->nmember, 256, 0, c->stream>>>(p, 0, c->tail_level0, c->nlevels);
        n++;
    }
    CU(cudaEventRecord(ev[4], c->stream));
    WGK_K(c, k_route_post)<<<grid, block, 0, c->stream>>>(p);
    CU(cudaEventRecord(ev[5], c->stream));
    c->launches += n;
    CU(cudaEventSynchronize(ev[5]));
    for (int k = 0; k < 5; k++) CU(cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]));
    CU(cudaEventElapsedTime(&ms[5], ev[0], ev[5]));
    for (auto &e : ev) cudaEventDestroy(e);
    return publish_discharge(c, 0);
}

// ---------------------------------------------------------------------------------------
// measurement aids of bench.py (include/wgk.h "diagnostics")
// ---------------------------------------------------------------------------------------
int wgk_stamps(wgk_ctx *c, int enable, unsigned long long *out) {
    if (!c) return WGK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    const size_t n = (size_t)2 * 2 * wgk::STAMP_DAYS;
    if (out) {
        if (!c->d_stamps) return fail(c, WGK_ERR_STATE, "stamps are not enabled");
        CU(cudaMemcpy(out, c->d_stamps, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
    if (enable && !c->d_stamps) {
        CU(cudaMalloc(&c->d_stamps, n * sizeof(unsigned long long)));
        drop_graph(c);  // the pointer is part of the kernel parameters baked into the graphs
    } else if (!enable && c->d_stamps) {
        cudaFree(c->d_stamps);
        c->d_stamps = nullptr;
        drop_graph(c);
    }
    if (c->d_stamps) {  // reset: start stamps to the largest value, end stamps to 0
        std::vector<unsigned long long> z(n, 0ull);
        for (int k = 0; k < 2; k++)
            for (int d = 0; d < wgk::STAMP_DAYS; d++) z[(size_t)(k * 2) * wgk::STAMP_DAYS + d] = ~0ull;
        CU(cudaMemcpy(c->d_stamps, z.data(), n * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    }
    return WGK_OK;
}

// one simulated day of the schedule wgk_step_days uses for this context, as plain launches with a CUDA event pair
// around EVERY launch: ms[k] / launches[k] per kernel class, k = 0 vertical (+ local routing) kernels, 1 river level
// kernels, 2 narrow-level tail kernels, 3 the rest (local routing / post-pass of the whole-day schedule, station record)
int wgk_profile_schedule(wgk_ctx *c, int day, int month, int dom, int slot, float ms[4], int launches[4]) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (!ms || !launches) return WGK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    rc = fill_calendar(c, day, month, dom, slot, 1);
    if (rc) return rc;
    rc = ensure_derived(c);
    if (rc) return rc;
    rc = forcing_before_step(c);
    if (rc) return rc;
    const WgkParams p = make_params(c);
    struct Rec { cudaEvent_t a, b; int cls; };
    std::vector<Rec> recs;
    auto timed = [&](int cls, auto &&launch) {
        Rec r{};
        r.cls = cls;
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        cudaEventRecord(r.a, c->stream);
        launch();
        cudaEventRecord(r.b, c->stream);
        recs.push_back(r);
    };
    const dim3 block = cell_block(c, 128);
    if (c->whole_day) {
        const dim3 grid = cell_grid(c, c->ncell, 128);
        timed(0, [&] { launch_vertical(c, p, 0); });
        timed(3, [&] { WGK_KW(c, k_route_local)<<<grid, block, 0, c->stream>>>(p, 0); });
        for (int l = 0; l < c->tail_level0; l++) {
            const dim3 g = cell_grid(c, c->level_off[l + 1] - c->level_off[l], 128);
            timed(1, [&] { WGK_KW(c, k_route_level)<<<g, block, 0, c->stream>>>(p, 0, l); });
        }
        if (c->tail_level0 < c->nlevels) timed(2, [&] { (c->opt.subtract_use > 0 ? wgk_wu::k_route_tail : wgk::k_route_tail)<<<c->nmember, 256, 0, c->stream>>>(p, 0, c->tail_level0, c->nlevels); });
        timed(3, [&] { WGK_K(c, k_route_post)<<<grid, block, 0, c->stream>>>(p); });
    } else {
        for (int l = 0; l < c->wave_level0; l++) {
            const int begin = c->level_off[l], end = c->level_off[l + 1];
            const dim3 g = cell_grid(c, end - begin, 128);
            if (c->level_tasks == 1 || (c->level_tasks == 2 && l == 0)) {
                timed(0, [&] { WGK_KW(c, k_level_day)<<<cell_grid(c, end - begin, wgk::VBLOCK), cell_block(c, wgk::VBLOCK), 0, c->stream>>>(p, 0, l); });
                continue;
            }
            timed(0, [&] { launch_cells_pre(c, p, 0, begin, end); });
            timed(1, [&] { WGK_KW(c, k_river_level)<<<g, block, 0, c->stream>>>(p, 0, l); });
        }
        for (size_t k = 0; k + 1 < c->wave_chunk_lo.size(); k++) {
            const int lo = c->wave_chunk_lo[k], hi = c->wave_chunk_lo[k + 1];
            const int begin = c->level_off[lo], end = c->level_off[hi];
            timed(0, [&] { launch_cells_pre(c, p, 0, begin, end); });
            timed(2, [&] { (c->opt.subtract_use > 0 ? wgk_wu::k_tail_chunk : wgk::k_tail_chunk)<<<c->nmember, 256, 0, c->stream>>>(p, 0, lo, hi); });
        }
    }
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    for (int k = 0; k < 4; k++) { ms[k] = 0.f; launches[k] = 0; }
    for (Rec &r : recs) {
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        ms[r.cls] += t;
        launches[r.cls]++;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    c->launches += (int64_t)recs.size();
    if (c->month_acc) c->month_days += 1;
    return publish_discharge(c, 0);
}

int wgk_fp64_peak(wgk_ctx *c, double *tflops) {
    if (!c || !tflops) return WGK_ERR_ARG;
    CU(cudaSetDevice(c->device));
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    const int iters = 1 << 16, blocks = sms * 8, threads = 256;
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    double best = 0.;
    for (int rep = 0; rep < 4; rep++) {  // first repetition warms up
        CU(cudaEventRecord(a, c->stream));
        wgk::k_fp64_peak<<<blocks, threads, 0, c->stream>>>(c->d_partial, iters, 0.999999, 1e-6);
        CU(cudaEventRecord(b, c->stream));
        CU(cudaEventSynchronize(b));
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, a, b));
        const double tf = 2. * 8. * iters * (double)blocks * threads / (t * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    c->launches += 4;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = best;
    return WGK_OK;
}

#ifdef WGK_PHASE_TIMING
// development builds only (-DWGK_PHASE_TIMING): read and reset the in-situ warp durations of the thread-per-cell task kernels
int wgk_debug_insitu(wgk_ctx *c, unsigned long long out[8]) {
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpyFromSymbol(out, wgk::g_insitu, sizeof(unsigned long long) * 8));
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CU(cudaMemcpyToSymbol(wgk::g_insitu, z, sizeof z));
    return WGK_OK;
}
// development builds only: cycles per phase of the thread-per-cell vertical step, out[2][2][8] (see g_vphase); read and reset
int wgk_debug_vphases(wgk_ctx *c, unsigned long long *out) {
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpyFromSymbol(out, wgk::g_vphase, sizeof(unsigned long long) * 32));
    unsigned long long z[32] = {0};
    CU(cudaMemcpyToSymbol(wgk::g_vphase, z, sizeof z));
    return WGK_OK;
}
// development builds only: cycles of every level-0 vertical warp on four days, out[4][1024]
int wgk_debug_warpdur(wgk_ctx *c, unsigned int *out) {
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpyFromSymbol(out, wgk::g_warpdur, sizeof(unsigned int) * 4 * 1024));
    return WGK_OK;
}
// development builds only: globaltimer stamps of the level-0 tasks of the last call, out[2][2][512]; reset = start stamps to ~0ull
int wgk_debug_stamps(wgk_ctx *c, unsigned long long *out, int reset) {
    CU(cudaStreamSynchronize(c->stream));
    if (out) CU(cudaMemcpyFromSymbol(out, wgk::g_stamp, sizeof(unsigned long long) * 2 * 2 * 512));
    if (reset) {
        std::vector<unsigned long long> z(2 * 2 * 512, 0ull);
        for (int k = 0; k < 2; k++)
            for (int d = 0; d < 512; d++) z[(k * 2 + 0) * 512 + d] = ~0ull;
        CU(cudaMemcpyToSymbol(wgk::g_stamp, z.data(), sizeof(unsigned long long) * z.size()));
    }
    return WGK_OK;
}
// development builds only (-DWGK_PHASE_TIMING): read and reset the per-phase cycle counters of the tile kernels
int wgk_debug_phases(wgk_ctx *c, unsigned long long out[8]) {
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpyFromSymbol(out, wgk::g_phase, sizeof(unsigned long long) * 8));
    unsigned long long z[8] = {0};
    CU(cudaMemcpyToSymbol(wgk::g_phase, z, sizeof z));
    return WGK_OK;
}
#endif

int64_t wgk_kernel_launches(const wgk_ctx *c) { return c ? c->launches : 0; }

}  // extern "C"
