/* wgk.h — C ABI of the B200-native WaterGAP2 daily hot path ("wgk" = WaterGAP kernels).
 *
 * Drop-in boundary: the reference's hot path is entered through two C++
 * class interfaces called from the day loop of integrate_wghm_ (integrateWGHM.cpp:755-798):
 *
 *   dailyWaterBalanceClass::calcNewDay(day, month, day_in_month, last_day_in_month, year, n, ...)
 *                                                            daily.h:24,  daily.cpp:94-1264
 *   routingClass::routing(year, day, month, day_in_month, last_day_in_month, ...)
 *                                                            routing.h:44, routing.cpp:1629-5244
 *   routingClass::updateLandAreaFrac(...)                    routing.h:46, routing.cpp:5343-5352
 *
 * plus the flow order produced by prepare_routing_files (rout_prepare.h:4, rout_prepare.cpp:61)
 * and read back in routingClass::init (routing.cpp:526-538).  The entry points below are what a
 * binding of those interfaces needs: plain pointers and sizes, int status codes, no C++ or
 * torch types.  The host-side look-alike classes in watergap2_b200/csrc/host/ and the ctypes
 * binding in watergap2_b200/__init__.py both sit on top of exactly this header; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *  - every host array is in the REFERENCE's cell order (array position n = cell number - 1)
 *    and in the reference's layout ([cell] or [cell][K]); the library permutes into its own
 *    device layout (routing order, structure of arrays, band-major snow);
 *  - the caller owns host buffers, the library owns device memory;
 *  - all functions return 0 on success or a negative wgk_status; they never throw or exit;
 *    wgk_last_error() gives the message of the last failure on that context;
 *  - one context lives on one GPU; contexts are independent (thread-compatible);
 *  - work is stream-ordered on the context's stream; wgk_synchronize() waits for it;
 *  - `member` = one ensemble member / calibration run sharing the grid and topology
 *    (enKF2wghmState.cpp / calibration.cpp run them as separate processes);
 *    `pset` = one set of per-cell calibration parameters (calib_param.h:72-101); each member
 *    points at one pset (wgk_set_member_pset), so an EnKF ensemble shares one pset and a
 *    calibration sweep has one pset per member.
 */
#ifndef WGK_H
#define WGK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WGK_VERSION 100
#define WGK_NBAND 101   /* G_Elevation / G_SnowInElevation channels: [0] mean, [1..100] bands (daily.h) */
#define WGK_NLCT 18     /* land cover classes, def.h:20 */
#define WGK_NPARAM 26   /* eCalibParam, calib_param.h:72-101 */

typedef struct wgk_ctx wgk_ctx;

typedef enum {
    WGK_OK = 0,
    WGK_ERR_ARG = -1,        /* bad argument (unknown field, wrong size, index out of range) */
    WGK_ERR_STATE = -2,      /* call order violated (e.g. step before topology) */
    WGK_ERR_CUDA = -3,       /* CUDA runtime error, see wgk_last_error */
    WGK_ERR_NOMEM = -4,
    WGK_ERR_TOPOLOGY = -5    /* routing order inconsistent with the downstream map */
} wgk_status;

/* run-time options; defaults = canonical option vector (OPTIONS.DAT order of
 * option.cpp:173-620).  Only the values listed are implemented; others return WGK_ERR_ARG. */
typedef struct {
    int restart;            /* additionalOutIn.additionalfilestatus (daily.cpp:165): 0 cold start, 1 from checkpoint */
    int tail_threshold;     /* routing levels with <= this many cells are run by one persistent CTA per member (0 = auto) */
    int use_graph;          /* 1: one CUDA graph of (day, level) tasks per wgk_step_days call (default), 0: plain launches in dependency order */
    int subtract_use;       /* options.subtract_use (option.cpp): 0 no water use (canonical), 2 net abstractions from surface water and
                               groundwater satisfied inside routing() with use_alloc 0, delayedUseSatisfaction 0, aggrNUsGloLakResOpt 0
                               (routing.cpp:1929-1935, 2193-2296, 2678-2788, 2866-2977, 3466-3512, 3590-3625, 3889-3908, 5503-5572).
                               The "wu_*" fields exist only then: the month's inputs in km3 per day are set by the caller at every
                               month start (wu_nus_month, wu_nug_month per parameter set; wu_wusi_month, wu_cusi_month), wu_frgi and
                               wu_alloc_coeff [ncell][5] once. */
} wgk_options;

/* ---- life cycle ---------------------------------------------------------------------- */
int wgk_create(wgk_ctx **ctx, int device, int ncell, int nmember, int npset, const wgk_options *opt);
void wgk_destroy(wgk_ctx *ctx);
const char *wgk_last_error(const wgk_ctx *ctx);
int wgk_synchronize(wgk_ctx *ctx);
/* the cudaStream_t all work of this context is ordered on (as void* to keep the header C-only);
 * wgk_set_stream adopts a caller stream (e.g. torch.cuda.current_stream().cuda_stream) */
void *wgk_get_stream(wgk_ctx *ctx);
int wgk_set_stream(wgk_ctx *ctx, void *cuda_stream);

/* ---- topology: replaces routingClass::init's reading of G_ROUT_ORDER / G_OUTFLC
 *      (routing.cpp:326, 526-538) ------------------------------------------------------ */
/* rout_order[n]: 1-based routing rank of cell n (G_ROUT_ORDER.UNF4, rout_prepare.cpp:834-886);
 * downstream_cell[n]: 1-based number of the cell n drains to, 0 = none (G_OUTFLC.UNF4). */
int wgk_set_topology(wgk_ctx *ctx, const int32_t *rout_order, const int32_t *downstream_cell);
/* Optional, BEFORE wgk_set_topology: a small class key per cell: bit 0 local lake, bit 1 local wetland,
 * bit 2 global lake/reservoir/wetland, bit 3 arid (bits 4-7: free for the caller, a secondary key).  Inside one
 * dependency level the device order is free (the upstream sums keep the reference's order through the CSR), so
 * cells of equal class are stored next to each other and the warps of the cell-parallel kernels skip the
 * water-body code they do not need; the classes with the longest code path (bit 2, then bits 0 / 1) come first
 * inside a level, so that the slowest warps of a kernel start first.  Results do not depend on it.  NULL resets
 * to plain routing order. */
int wgk_set_cell_classes(wgk_ctx *ctx, const uint8_t *cell_class);
int wgk_num_levels(const wgk_ctx *ctx);
/* level[n] (0-based dependency level of cell n) for ncell cells; for tests and basin sharding */
int wgk_get_levels(const wgk_ctx *ctx, int32_t *level);

/* ---- fields ---------------------------------------------------------------------------- */
/* Fields are addressed by the names used throughout the repository (wgk_fields.h; the same
 * names as the oracle and the reference dump).  `index` is the member for state/flux
 * fields, the pset for parameter-derived fields and ignored (0) for shared statics.
 * `bytes` must equal the field's host size (checked).  Host layouts:
 *   per-cell fields  [ncell];  "elevation" i16 [ncell][101];  "snow_bands" f64 [ncell][101];
 *   "params" f64 [26][ncell] (eCalibParam order);  table fields [18]. */
int wgk_field_id(const char *name);                     /* < 0 if unknown */
int wgk_field_info(int field, const char **name, const char **dtype, int64_t *host_count_per_index, int ncell);
int wgk_set_field(wgk_ctx *ctx, int field, int index, const void *host, size_t bytes);
int wgk_get_field(wgk_ctx *ctx, int field, int index, void *host, size_t bytes);
int wgk_set_member_pset(wgk_ctx *ctx, int member, int pset);
/* raw device pointer of element (member, device position 0) of a per-member field, for zero-copy access (NCCL via
 * torch).  Device layout: positions in routing order; element (member, position x) lies wgk_member_stride() * member +
 * wgk_cell_stride() * x elements behind the field's base.  wgk_layout: 0 = cell-minor [member][cell] (few members),
 * 1 = member-minor [cell][member] with the members padded to a multiple of 32 (many members: a warp works on 32 members
 * of one cell; chosen by wgk_create from the member count, results are bit-identical between the layouts). */
void *wgk_device_ptr(wgk_ctx *ctx, int field, int member);
int64_t wgk_cell_stride(const wgk_ctx *ctx);
int64_t wgk_member_stride(const wgk_ctx *ctx);
int wgk_layout(const wgk_ctx *ctx);
/* schedule of a multi-day call chosen by wgk_create (bit flags): 1 = whole-grid kernels day after day (many members), else
 * the (day, level) wavefront graph; 2 = the wavefront's levels run as ONE task per (day, level) (vertical part and river part in
 * one kernel, programmatic edge from the upstream level; the default below 800 000 cell-members), else as two kernels per wide
 * level and one persistent CTA per narrow level; 4 = cell-owner kernel (opt-in). */
int wgk_schedule(const wgk_ctx *ctx);
/* rank_of_cell[n] = position of reference cell n in the device (routing) order */
int wgk_get_device_order(const wgk_ctx *ctx, int32_t *rank_of_cell);

/* ---- forcing: replaces climateClass::read_climate_data_daily's in-memory grids
 *      G_precipitation_d / G_temperature_d / G_shortwave_d / G_longwave_d
 *      (climate.cpp:93-138, climate.h:14-21; float [ncell][31]) --------------------------- */
/* The context holds `nslots` forcing days on the device ([slot][cell] of float4 P,T,SW,LW).
 * wgk_set_forcing copies `ndays` days starting at channel 0 of the four host grids (reference
 * layout [ncell][stride], stride = 31 for the .31 monthly files) into slots slot0.. ;
 * member < 0 = shared by all members, else per-member forcing (EnKF perturbed forcing).
 * The transposition / permutation runs on the device.  The copy is ASYNCHRONOUS on the library's copy stream
 * (it overlaps the stepping of other slots): the four host buffers must stay valid and unchanged until the next
 * wgk_synchronize() or until a later wgk_step_days call that reads these slots has been synchronised; pinned
 * buffers make the copy truly asynchronous, pageable ones are staged by the CUDA runtime before the call returns. */
int wgk_forcing_reserve(wgk_ctx *ctx, int nslots, int per_member);
int wgk_set_forcing(wgk_ctx *ctx, int slot0, int ndays, int member, const float *prec, const float *temp,
                    const float *shortwave, const float *longwave, int stride);
/* the same with the four buffers holding the BYTES of the reference's big-endian UNF0 files
 * (GPREC_<year>_<month>.31.UNF0 ..., climate.cpp:100-123; e.g. the mmap'ed files): byte order and layout are
 * converted on the device */
int wgk_set_forcing_unf(wgk_ctx *ctx, int slot0, int ndays, int member, const void *prec, const void *temp,
                        const void *shortwave, const void *longwave, int stride);

/* ---- the hot path ---------------------------------------------------------------------- */
/* One call = what the day loop of integrateWGHM.cpp:755-798 does for all cells and members:
 * calcNewDay for every continental cell, routing(), updateLandAreaFrac().
 * day 1..365 (day of the 365-day model year), month 0..11, day_in_month 1..31, slot =
 * forcing slot of that day.  The three-call form exists for the class shims; wgk_step_days
 * runs `ndays` (<= 366) consecutive days (365-day model years, forcing slots slot0, slot0+1, ...
 * modulo the reserved slots) as ONE CUDA graph of (day, routing level) tasks whose edges are the
 * data dependencies, so that the level chain of one day overlaps the following days. */
int wgk_vertical_day(wgk_ctx *ctx, int day, int month, int day_in_month, int slot);
int wgk_routing_day(wgk_ctx *ctx, int day, int month, int day_in_month);
int wgk_update_land_area_frac(wgk_ctx *ctx);
int wgk_step_days(wgk_ctx *ctx, int day, int month, int day_in_month, int slot0, int ndays);

/* ---- diagnostics ----------------------------------------------------------------------- */
/* total water storage of one member in km3 (canopy+snow+soil on the land fraction, plus the
 * seven routing compartments): the daily global mass-balance check of BASELINE.md */
int wgk_total_storage_km3(wgk_ctx *ctx, int member, double *out);
/* per-day record of river discharge at `ncells` chosen cells (station series, routing.cpp:4232-4238), `max_days` rows.
 * wgk_record_cells starts the record at row 0; every wgk_step_days call appends its days after those of the calls
 * before it (a multi-year calibration run is several calls); a call whose days do not fit into the remaining rows
 * restarts the record at row 0 with its first day, and wgk_record_rewind does so explicitly.  wgk_get_record copies
 * the first `ndays` rows to host out[ndays][ncells] and fails (WGK_ERR_ARG) when fewer rows have been recorded. */
int wgk_record_cells(wgk_ctx *ctx, const int32_t *cells, int ncells, int max_days);
int wgk_record_rewind(wgk_ctx *ctx);
int wgk_get_record(wgk_ctx *ctx, int member, double *out, int ndays);
/* ---- EnKF state bridge (what the PDAF coupling of the reference exchanges with the model) ----------
 * wgk_month_begin: start accumulating the daily WghmStateFile entries of the seven routing compartments
 * (routing.cpp:5002-5020) for Cell::mean (wghmStateFile.cpp:711-728); every stepped day counts.
 * wgk_state_vector: extract_sub_ (extractsub.cpp:65-79) for one member: out[ncells][10] = canopy, snow, soil,
 * local lake, local wetland, global lake, global wetland, reservoir, river, groundwater of the given cells
 * (0-based) in mm over the continental area, kind 0 = mean of the month so far, kind 1 = last day, minus
 * mean_field[ncells][10] (NULL: nothing subtracted).
 * wgk_enkf_update: enkf_wghmstate_ (enKF2wghmState.cpp:89-121, 440-471) for one member: last-day state +=
 * field - prediction with the reference's limits (canopy, snow <= 1000, soil, local / global wetland,
 * reservoir, river >= 0; lakes and groundwater may go negative), snow bands rescaled by (field + mean_field) /
 * monthly mean snow (or set to 1/100 of it where that mean is 0) within [0, 1000], and the result restored as
 * the start state of the next cycle (daily.cpp:1896-1924, routing.cpp:851-882).  field, prediction, mean_field:
 * host [ncells][10]. */
int wgk_month_begin(wgk_ctx *ctx);
/* the day's WghmStateFile entry of the seven routing compartments (routing.cpp:5002-5020: local lake, local wetland, global lake,
 * global wetland, reservoir, river, groundwater in mm over the continental area) of all cells in reference order, host
 * out[7][ncell]: packed on the device, ONE device-to-host copy (what the class shim routingClass::routing needs per day) */
int wgk_get_day_state(wgk_ctx *ctx, int member, double *out);
int wgk_state_vector(wgk_ctx *ctx, int member, int kind, const int32_t *cells, int ncells, const double *mean_field, double *out);
int wgk_enkf_update(wgk_ctx *ctx, int member, const int32_t *cells, int ncells, const double *field, const double *prediction,
                    const double *mean_field);

/* ---- ensemble statistics: the one exchange between GPUs -------------------------------------
 * wgk_ensemble_moments: over ALL members of this context, sum and sum of squares of the extract_sub_ state vector
 * (the values of wgk_state_vector with mean_field NULL, kind as there) of `cells` (0-based; NULL = all ncell cells in
 * reference order), each [ncells][10] f64, members added in ascending order.  The results stay on the device in ONE
 * library-owned buffer (*d_sumsq == *d_sum + 10 * ncells, valid until the next call), written stream-ordered on the
 * context's stream, so that the caller can all-reduce 2 * 10 * ncells doubles over NCCL in place (one process per
 * GPU; enKF2wghmState.cpp runs the members as separate processes and PDAF forms the statistics over MPI).
 * wgk_moments_finish: mean = sum / nmember_total and population variance max(0, sumsq / nmember_total - mean^2) of
 * the (all-reduced) buffers, in place on the device, copied to host mean / var ([ncells][10], either may be NULL). */
int wgk_ensemble_moments(wgk_ctx *ctx, int kind, const int32_t *cells, int ncells, void **d_sum, void **d_sumsq);
int wgk_moments_finish(wgk_ctx *ctx, int nmember_total, double *mean, double *var);

/* ---- set-up of large ensembles and parameter sweeps on the device -------------------------------------------
 * wgk_copy_index: copy every field of one scope (1 = parameter set, 2 = member: state and fluxes) from index src
 * to index dst, device to device (an ensemble starts from one state; calibration runs share everything but the
 * calibrated parameters).  wgk_fill_field: one value for all cells of a per-cell f64 field of one index (the
 * reference calibrates one gamma / CFA per basin, calibration.cpp:266-527). */
int wgk_copy_index(wgk_ctx *ctx, int scope, int src, int dst);
int wgk_fill_field(wgk_ctx *ctx, int field, int index, double value);

/* one simulated day with plain launches and CUDA events between the phases, on the context's
 * stream: ms[0] vertical, ms[1] routing pre-pass (cell-parallel), ms[2] wide routing levels
 * (one launch each), ms[3] narrow-level tail (one persistent CTA per member), ms[4] routing
 * post-pass (cell-parallel), ms[5] whole day.
 * Advances the model state by that day. Used by bench.py for the per-kernel roofline. */
int wgk_profile_day(wgk_ctx *ctx, int day, int month, int day_in_month, int slot, float ms[6]);
/* the same for the schedule wgk_step_days really runs for this context ((day, level) wavefront tasks, or the whole-day
 * kernels of many-member runs): one simulated day as plain launches with an event pair around every launch, summed per
 * kernel class: 0 vertical balance (+ local routing) kernels, 1 river-level kernels, 2 narrow-level tail kernels,
 * 3 the rest.  Advances the model state by that day. */
int wgk_profile_schedule(wgk_ctx *ctx, int day, int month, int day_in_month, int slot, float ms[4], int launches[4]);
/* %globaltimer stamps of the level-0 tasks (the dominant kernels) INSIDE the running graph: enable != 0 switches them on
 * (and resets them), 0 off; out (may be NULL) receives the stamps of the calls since the last reset as
 * u64 [2: vertical task, river task][2: first warp start, last warp end][512 day offsets] in ns. */
int wgk_stamps(wgk_ctx *ctx, int enable, unsigned long long *out);
/* measured DFMA throughput of the context's GPU in TFLOP/s (8 independent FMA chains per thread, 8 CTAs of 256 per SM):
 * the denominator of the FP64-pipe fractions bench.py reports */
int wgk_fp64_peak(wgk_ctx *ctx, double *tflops);
/* number of kernels this context has launched (graph nodes counted per replay) */
int64_t wgk_kernel_launches(const wgk_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* WGK_H */
