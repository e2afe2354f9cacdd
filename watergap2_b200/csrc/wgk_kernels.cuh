// wgk_kernels.cuh — sm_100a kernels of the WaterGAP2 daily hot path.
//
// FP64 structure-of-arrays; device cell order = routing order by dependency level (inside a level optionally
// sorted by cell class), so every level of the river network is one contiguous, coalesced index range.
//
// timed path (wgk_step_days: one CUDA graph of (day, level) tasks)
//   k_cells_pre_tpc / k_vertical_tpc   dailyWaterBalanceClass::calcNewDay (daily.cpp:94-1264, lai.cpp:152-306), one
//                     thread per (member, cell), band columns staged by cp.async; k_cells_pre* adds the cell-parallel
//                     part of routingClass::routing that does not depend on upstream cells: runoff split,
//                     groundwater of humid cells and inland sinks, local lake, local wetland (routing.cpp:1878-2617)
//   k_cells_pre<VCfg> / k_vertical<VCfg>   the same, CTA-cooperative and band-parallel (tile of 32 cells x NW warps)
//   k_river_level     one dependency level of the ordered cell loop, reduced to what depends on upstream cells:
//                     inflow gather, global lake / reservoir / global wetland / arid groundwater where present,
//                     river reach (routing.cpp:2623-3545), then the post-pass of the same cells: river width /
//                     area fraction, surface-water-body fractions, next-day land area fraction,
//                     updateLandAreaFrac (routing.cpp:3546-3586, 5034-5188, 5343-5352)
//   k_tail_chunk      the same for narrow levels inside ONE persistent CTA per member, levels separated by
//                     __syncthreads() instead of kernel launches, next level's inputs prefetched before the barrier
//   k_end_of_day      station discharge record
//   k_days_owner      opt-in alternative to the graph: one launch per call, a thread owns its cell for all days
// three-call class-shim path (wgk_vertical_day / wgk_routing_day) and wgk_profile_day
//   k_route_local, k_route_level, k_route_tail, k_route_post   (also the whole-day schedule of many-member runs)
// set-up and exchange
//   k_derive_static, k_derive_member   quantities derived once from statics / parameters / uploaded state
//   k_forcing_pack    [cell][31] float grids -> [slot][cell] float4 in device order
//   k_fill_calendar   calendar of a multi-day call on the device (one graph replays for any start date)
//   k_state_vector, k_enkf_update      EnKF state bridge (extractsub.cpp:65-79, enKF2wghmState.cpp:89-121, 440-471)
//   k_total_storage   global storage in km3 (mass-balance check)
//
// Build with -fmad=false: the reference CPU build has no FMA contraction, and results are
// compared at 1e-10 relative.  Expression shapes follow the reference line by line.
// (no include guard: wgk_api.cu includes this header twice, once per layout - see WGK_MM)

#include <cstdint>
#ifndef WGK_EMU  // tests/emu compiles this header for the host to check the kernel logic
#include <cuda_pipeline.h>
#include <cuda_runtime.h>
#endif
#include "wgk_fields.h"

#ifndef WGK_MM_UNIFORM_CELL
#define WGK_MM_UNIFORM_CELL 1  // member-minor thread mapping: one cell per CTA (see map_thread); measured on B200 against the linear mapping,
                               // with 5 resident CTAs per SM: 1024 sets 2.65 vs 2.55, 256 members 2.29 vs 2.17 x 10^9 cell-days/s
#endif
#ifndef WGK_MM
#define WGK_MM 0  // 0: kernels of the cell-minor layout in namespace wgk, 1: of the member-minor layout in namespace wgk_mm
#endif
#ifndef WGK_PARAMS_DEFINED
#define WGK_PARAMS_DEFINED
struct WgkParams {
    WgkArrays a;
    const int32_t *member_pset;   // [nmember]
    const float4 *forcing;        // [slot][fmember][stride]  (P, T, SW, LW)
    const int32_t *up_off;        // [ncell+1] CSR of upstream cells, device order
    const int32_t *up_idx;        // upstream ranks, ascending (= reference accumulation order)
    const int32_t *down;          // [ncell] downstream rank or -1
    const int32_t *level_off;     // [nlevels+1]
    int32_t *cal;                 // base of the current call {day, month, day_in_month, slot, first row of the station record}
    int32_t *cal_days;            // [max days per call][4] {day of year, month, day in month, forcing slot}
    double *qbuf;                 // [QBUF_K][nmember][stride] river discharge of the days in flight
    const int32_t *gidx;          // [ncell] index into the global-water-body scratch or -1
    double *gbody;                // [nmember][ngbody][GB_N] inflow-independent terms of the day
    int ngbody;
    double *record;               // [max_days][nmember][nrec] discharge record, or null
    const int32_t *record_cells;  // [nrec] device ranks
    int nrec, record_max_days;
    int ncell, stride, nmember, npset;
    int forcing_nslots, forcing_per_member;
    int restart;
    int month_acc;                // accumulate the daily WghmStateFile values of the month (EnKF bridge)
    int subtract_use;             // options.subtract_use: 0 no water use, 2 net abstractions
    const int32_t *wu_res_idx;    // [5][stride] device position of the i-th downstream cell whose use an irrigation reservoir serves, or -1
    int nlevels;
    // Layout of the member- and parameter-set-scoped arrays:
    //   mm == 0  cell-minor   [member][band][cell]: a warp = 32 consecutive cells of one member (few members, latency regime)
    //   mm == 1  member-minor [band][cell][member]: a warp = 32 members of ONE cell (many members, throughput regime): statics
    //            are warp-uniform loads and the 32 lanes share land cover, water-body class and nearly the same weather
    // The kernels are compiled once per layout (namespaces wgk / wgk_mm, WGK_MM below) so that the index arithmetic of either
    // is free of run-time selects; the index helpers mi() ... gs() live in the namespace.
    int mm, mpad, ppad;           // mpad / ppad: rows of the member / parameter-set arrays (members padded to 32 when mm)
    unsigned long long *stamps;   // optional (wgk_stamps): [2: V, R][2: first warp start, last warp end][STAMP_DAYS] %globaltimer ns of the
                                  // level-0 tasks of a call, i.e. their duration INSIDE the running graph; null = off
    int stamp_level;              // level whose fused task (k_level_day) is stamped (0; WGK_STAMP_LEVEL for the timeline tools)
};

// cell-owner schedule (k_days_owner): its hand-off structures
struct WgkOwner {
    const int32_t *warp_begin, *warp_end;  // [nwarps] device-order cell range of a warp (one level, <= 32 cells)
    const int32_t *cell_warp;              // [ncell] warp that owns a cell
    unsigned long long *ring;              // [QBUF_K][nmember][stride][2] tagged discharge entries
    uint32_t *progress;                    // [nmember][nwarps] tag of the last day a warp has consumed
    int32_t *abort;                        // set when a wait timed out
    const int32_t *rec_head, *rec_next;    // station records of a cell: rec_head[cell] -> k, rec_next[k] -> k' (or -1)
    long long *timing;                     // optional [nwarps][4] cycles of lane 0: vertical+local, waits, river+hand-off, post (tools/owner_timing.py)
    uint32_t base;                         // tag of day offset d of this call = base + d + 1 (days stepped so far by this context)
    int nwarps;
};
#endif  // WGK_PARAMS_DEFINED

#ifndef WGK_WU
#define WGK_WU 0  // 1: the kernels that contain water use compiled WITH it, in namespaces wgk_wu / wgk_mm_wu; the canonical
                  // configuration (subtract_use 0) runs kernels without a trace of it (it cost 5 % of the single-member year as a run-time branch)
#endif
#undef WGK_NS
#if WGK_MM && WGK_WU
#define WGK_NS wgk_mm_wu
#elif WGK_MM
#define WGK_NS wgk_mm
#elif WGK_WU
#define WGK_NS wgk_wu
#else
#define WGK_NS wgk
#endif
namespace WGK_NS {

constexpr bool MM = (WGK_MM != 0);
constexpr bool WU = (WGK_WU != 0);
#undef WGK_WU_PARAMS
#undef WGK_WU_ARGS
#if WGK_WU  // the out-of-line global-water-body function carries the day's use only in the water-use build
#define WGK_WU_PARAMS , double &remainingUse, double &dailyActualUse
#define WGK_WU_ARGS , remainingUse, dailyActualUse
#else
#define WGK_WU_PARAMS
#define WGK_WU_ARGS
#endif
// element (member m, device position r) of a member array / of a parameter-set array
__host__ __device__ __forceinline__ size_t mi(const WgkParams &p, const int m, const int r) { return MM ? (size_t)r * p.mpad + m : (size_t)m * p.stride + r; }
__host__ __device__ __forceinline__ size_t qi_of(const WgkParams &p, const int ps, const int r) { return MM ? (size_t)r * p.ppad + ps : (size_t)ps * p.stride + r; }
__device__ __forceinline__ size_t qi(const WgkParams &p, const int m, const int r) { return qi_of(p, p.member_pset[m], r); }
// element (m, band b, r) of a band array with nb bands = bi(p, m, r, nb) + b * band_stride(p)
__host__ __device__ __forceinline__ size_t bi(const WgkParams &p, const int m, const int r, const int nb) { return MM ? (size_t)r * p.mpad + m : (size_t)m * nb * p.stride + r; }
__host__ __device__ __forceinline__ size_t band_stride(const WgkParams &p) { return MM ? (size_t)p.stride * p.mpad : (size_t)p.stride; }
// forcing [slot][fmember][cell] (cell-minor) / [slot][cell][fmember] (member-minor)
__host__ __device__ __forceinline__ size_t fi(const WgkParams &p, const int slot, const int m, const int r) {
    if (!p.forcing_per_member) return (size_t)slot * p.stride + r;
    return MM ? ((size_t)slot * p.stride + r) * p.mpad + m : ((size_t)slot * p.nmember + m) * p.stride + r;
}
// per-day scratch of the cells with a global water body: element k of (member m, slot gi) = gbody[gb(p, m, gi, nk) + k * gbody_stride(p)]
__host__ __device__ __forceinline__ size_t gb(const WgkParams &p, const int m, const int gi, const int nk) { return MM ? (size_t)gi * nk * p.mpad + m : ((size_t)m * p.ngbody + gi) * nk; }
__host__ __device__ __forceinline__ size_t gbody_stride(const WgkParams &p) { return MM ? (size_t)p.mpad : (size_t)1; }


constexpr double MIN_STOR_VOL = 1.e-15;  // routing.h:24

// Division by one of the unit-conversion constants of the reference (/ 100., / 1000000., / 1000., / 30.):
// q0 = x * RN(1/c), then one FMA residual correction (Markstein).  With RN(1/c) correctly rounded and the
// significand of c not all ones the result IS the correctly rounded quotient x / c, i.e. bit-identical to the
// division it replaces, in 3 dependent FP64 instructions instead of ~25 (there are ~80 such divisions on the
// path of one cell-day).  tests/test_kernel_logic_cpu.py::test_const_division_is_exact checks 4e7 operands.
// the two 32-bit words of a double (integer tests on the bits)
__device__ __forceinline__ int hi_word(const double x) {
#ifdef __CUDA_ARCH__
    return __double2hiint(x);
#else
    long long b; memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
__device__ __forceinline__ int lo_word(const double x) {
#ifdef __CUDA_ARCH__
    return __double2loint(x);
#else
    long long b; memcpy(&b, &x, 8); return (int)(b & 0xffffffffll);
#endif
}
// pow() of the hot path.  WGK_POW_OUTLINE: ONE out-of-line copy per kernel variant instead of ~300 inlined instructions at each of
// the ~12 call sites of a cell-day (code size against the instruction cache; the fused task k_level_day is 120 KB of SASS)
#if defined(WGK_POW_OUTLINE) && defined(__CUDA_ARCH__)
__device__ __noinline__ double wg_pow(const double x, const double y) { return pow(x, y); }
#else
__device__ __forceinline__ double wg_pow(const double x, const double y) { return pow(x, y); }
#endif
struct ConstDiv {
    double c, rc;
};
__device__ __forceinline__ double operator/(const double x, const ConstDiv d) {
    const double q = x * d.rc;
    return fma(fma(-d.c, q, x), d.rc, q);
}
constexpr ConstDiv C100{100., 0.01}, C1E6{1000000., 1e-6}, C1000{1000., 0.001}, C30{30., 1. / 30.},
                   C86400{86400., 1. / 86400.};

// thread -> (device position r in [begin, end), member m) for the cell-parallel kernels.  Cell-minor layout: blockIdx.y = member,
// consecutive threads = consecutive cells.  Member-minor layout: consecutive threads = consecutive members of ONE cell (a warp
// never straddles two cells, mpad is a multiple of 32); lanes beyond the last member idle.
__device__ __forceinline__ bool map_thread(const WgkParams &p, const int begin, const int end, int &r, int &m) {
    if (MM) {
#if WGK_MM_UNIFORM_CELL
        // one cell per CTA (blockIdx.x), the CTA's threads = min(128, mpad) consecutive members, blockIdx.y = member block: the cell
        // index is uniform over the CTA, so the compiler keeps the cell's statics in uniform registers
        r = begin + blockIdx.x;
        m = blockIdx.y * blockDim.x + threadIdx.x;
        return m < p.nmember;
#else
        const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        r = begin + (int)(t / p.mpad);
        m = (int)(t % p.mpad);
        return r < end && m < p.nmember;
#endif
    }
    r = begin + blockIdx.x * blockDim.x + threadIdx.x;
    m = blockIdx.y;
    return r < end;
}

// Shared-memory staging of the 100 snow bands of a cell: the band columns of the CTA's 128
// cells are brought in by cp.async (LDGSTS) in chunks of SNOW_CH bands, double buffered, so
// that all loads of the band loop are in flight while the thread evaluates radiation / PET /
// canopy, and the loop itself reads shared memory.  Every thread copies and reads only its own
// column, so no block-wide barrier is involved.
#ifndef WGK_SNOW_CH
#define WGK_SNOW_CH 5  // bands per staged chunk (divides 100); measured with the select-form band loop, ms per simulated year
                      // at 1 / 64 members: 10 bands 23.5 / 869, 5 bands 22.0 / 850, 4 bands 22.1 / 857, 2 bands 23.3 / 934
#endif
#ifndef WGK_VBLOCK
#define WGK_VBLOCK 128  // threads per CTA of the thread-per-cell kernels
#endif
constexpr int SNOW_CH = WGK_SNOW_CH, SNOW_NCH = 100 / WGK_SNOW_CH, VBLOCK = WGK_VBLOCK;
#ifndef WGK_TPC_MINB
#define WGK_TPC_MINB 4  // resident CTAs per SM the thread-per-cell kernels are compiled for (register cap 65536 / (128 * MINB));
                        // 5 and 6 spill and measured 9 % / 14 % slower at one member on B200
#endif
#ifndef WGK_TPC_MINB_MM
#define WGK_TPC_MINB_MM 5  // member-minor instantiation: with the cell index uniform over the CTA the kernel fits 96 registers without spills
#endif
#undef WGK_TPC_MINB_EFF
#if WGK_MM
#define WGK_TPC_MINB_EFF WGK_TPC_MINB_MM
#else
#define WGK_TPC_MINB_EFF WGK_TPC_MINB
#endif
#ifndef WGK_SNOW_NBUF
#define WGK_SNOW_NBUF 2  // staged chunks in flight per thread
#endif
#ifndef WGK_BAND_FORM
#define WGK_BAND_FORM 1  // 0: band by band, 1: the bands of a chunk stage by stage (see vertical_cell)
#endif
constexpr int SNOW_NBUF = WGK_SNOW_NBUF;
static_assert(SNOW_NBUF >= 2 && SNOW_NBUF <= SNOW_NCH && SNOW_NBUF <= 8, "stage depth");
struct SnowStage {
    double s[SNOW_NBUF][SNOW_CH][VBLOCK];
    int32_t e[SNOW_NBUF][SNOW_CH][VBLOCK];
};
// before chunk c is read: all but the youngest min(NBUF - 1, chunks after c) copy groups must have landed
__device__ __forceinline__ void stage_wait(const int remaining) {
    if (remaining >= SNOW_NBUF - 1) { __pipeline_wait_prior(SNOW_NBUF - 1); return; }
#pragma unroll
    for (int n = SNOW_NBUF - 2; n >= 1; n--)
        if (remaining == n) { __pipeline_wait_prior(n); return; }
    __pipeline_wait_prior(0);
}
__device__ __forceinline__ void stage_issue(SnowStage *st, const int buf, const int t, const double *S, const int32_t *E,
                                            const size_t sstride, const size_t estride) {
#pragma unroll
    for (int k = 0; k < SNOW_CH; k++) {
        __pipeline_memcpy_async(&st->s[buf][k][t], S + (size_t)k * sstride, sizeof(double));
        __pipeline_memcpy_async(&st->e[buf][k][t], E + (size_t)k * estride, sizeof(int32_t));
    }
    __pipeline_commit();
}


// ----------------------------------------------------------------------------------------
// LAI growing-season state machine (lai.cpp:179-293)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ double lai_growing(int &days, int initialDays, int &status, int lct, bool arid,
                                              double LAImin, double LAImax, double &precsum, double prec) {
    if (status == 0) {
        if (days >= initialDays) {
            days++;
            precsum += prec;
            if (precsum > 40.) {
                if (days >= initialDays + 30) {
                    days = initialDays + 30;
                    status = 1;
                }
                return (LAImin + (LAImax - LAImin) * (days - initialDays) / C30);
            } else {
                days = initialDays;
                return LAImin;
            }
        } else {
            days++;
            precsum += prec;
            return LAImin;
        }
    } else {
        if (days <= 30) {
            days--;
            if (lct <= 2) status = 0;
            if (days <= 0) {
                days = 0;
                status = 0;
                precsum = 0.;
            }
            return (LAImax - (LAImax - LAImin) * (30 - days) / C30);
        } else {
            if (arid && (prec < 0.5)) days--;
            else days = 30 + initialDays;
            return LAImax;
        }
    }
}

__device__ __forceinline__ double lai_nogrowing(int &days, int initialDays, int &status, double LAImin, double LAImax,
                                                double &precsum, double prec) {
    if (status == 0) {
        if (days > initialDays) {
            days++;
            precsum += prec;
            if (precsum > 40.) {
                if (days >= initialDays + 30) {
                    days = initialDays + 30;
                    status = 1;
                }
                return (LAImin + (LAImax - LAImin) * (days - initialDays) / C30);
            } else {
                days = initialDays;
                return LAImin;
            }
        } else {
            precsum += prec;
            return LAImin;
        }
    } else {
        if (days <= 30) {
            days--;
            if (days <= 0) {
                days = 0;
                status = 0;
                precsum = 0.;
            }
            return (LAImax - (LAImax - LAImin) * (30 - days) / C30);
        } else {
            days--;
            return LAImax;
        }
    }
}

// inputs of the cell-parallel routing pre-pass that are known before today's vertical step (statics,
// parameters, yesterday's state): the band-parallel kernel loads them at its very start, in a warp of
// its own, so that no global-memory latency is left between the vertical step and the local routing
struct LocalIn {
    double ekg, invkg, evaredex, area, cfa, contf, fswb_init, loc_lake, loc_wetland, kS, lake_depth, wetl_depth;
    double laf, laf_prev, gw, loc_lake_stor, red_loc_lake, loc_wetl_stor, red_loc_wetl, raf_next;
    int contcell, flags, ldd, arid;
};
// today's fluxes of the vertical step (daily.h G_openWaterPrec ... G_dailyStorageTransfer)
struct LocalFlux {
    double owPrec, owPET, storage_transfer, surface_runoff, gw_recharge;
};

__device__ __forceinline__ LocalIn local_load(const WgkParams &p, const int r, const int m) {
    const WgkArrays &a = p.a;
    const size_t i = mi(p, m, r);
    const size_t q = qi(p, m, r);
    LocalIn li;
    li.contcell = a.contcell[r];
    li.flags = a.s_flags[r];
    li.ldd = a.ldd[r];
    li.arid = a.arid[r];
    li.ekg = a.s_ekg[q];
    li.invkg = a.s_invkg[q];
    li.evaredex = a.p_evaredex[q];
    li.area = a.area[r];
    li.cfa = a.cfa[q];
    li.contf = a.contfreq[r];
    li.fswb_init = a.fswb_init[r];
    li.loc_lake = a.loc_lake[r];
    li.loc_wetland = a.loc_wetland[r];
    li.kS = a.p_swoutf[q];
    li.lake_depth = a.lake_depth_active[q];
    li.wetl_depth = a.wetl_depth_active[q];
    li.laf = a.land_area_frac[i];
    li.laf_prev = a.land_area_frac_prev[i];
    li.gw = a.gw[i];
    li.loc_lake_stor = a.loc_lake_stor[i];
    li.red_loc_lake = a.red_loc_lake[i];
    li.loc_wetl_stor = a.loc_wetl_stor[i];
    li.red_loc_wetl = a.red_loc_wetl[i];
    li.raf_next = a.river_area_frac_next[i];
    return li;
}

__device__ __forceinline__ LocalFlux local_flux_load(const WgkParams &p, const int r, const int m) {
    const WgkArrays &a = p.a;
    const size_t i = mi(p, m, r);
    LocalFlux fx;
    fx.owPrec = a.openwater_prec[i];
    fx.owPET = a.openwater_pet[i];
    fx.storage_transfer = a.storage_transfer[i];
    fx.surface_runoff = a.surface_runoff[i];
    fx.gw_recharge = a.gw_recharge[i];
    return fx;
}

#ifdef WGK_PHASE_TIMING  // development aid: cycles per phase of vertical_cell / local routing, per warp (tools/vphase_timing.py)
__device__ unsigned long long g_vphase[2][2][8];  // [level 0?][warp runs the band loop?][phase 0..6 cycles summed over warps, 7 = warps]
#define WGK_VT_BEGIN() long long vt_ = clock64(), vt_d_[7] = {0, 0, 0, 0, 0, 0, 0}; int vt_lane_ = 0, vt_cls_ = 0
#define WGK_VT_CLASS(run_) do { const unsigned b_ = __ballot_sync(__activemask(), (run_)); vt_cls_ = b_ != 0; \
    vt_lane_ = (b_ ? __ffs(b_) : __ffs(__activemask())) - 1; } while (0)
#define WGK_VT(k_) do { const long long t_ = clock64(); vt_d_[k_] += t_ - vt_; vt_ = t_; } while (0)
#define WGK_VT_FLUSH(l0_) do { if ((int)(threadIdx.x & 31) == vt_lane_) { \
    for (int k_ = 0; k_ < 7; k_++) atomicAdd(&g_vphase[l0_][vt_cls_][k_], (unsigned long long)vt_d_[k_]); atomicAdd(&g_vphase[l0_][vt_cls_][7], 1ull); } } while (0)
#else
#define WGK_VT_BEGIN() do { } while (0)
#define WGK_VT_CLASS(run_) do { } while (0)
#define WGK_VT(k_) do { } while (0)
#define WGK_VT_FLUSH(l0_) do { } while (0)
#endif
// ----------------------------------------------------------------------------------------
// vertical water balance, one thread per cell (throughput form, used when members x cells fill the GPU)
// ----------------------------------------------------------------------------------------
// -> true when the cell was computed; then *li (if given) holds the inputs of the local routing, loaded together
// with those of the soil part, and *fx today's fluxes (otherwise the caller reads both from global memory)
// EARLY: the inputs of the soil part and of the local routing are loaded before the band loop instead of after it (the compiler
// cannot move a load above the band stores), one memory round trip less on the cell-day chain at the price of ~40 registers
// held across the loop: for the kernel without a register cap (k_level_day)
template <bool EARLY = false>
__device__ __forceinline__ bool vertical_cell(const WgkParams &p, const int r, const int m, const int slot, SnowStage *st,
                                              LocalIn *li = nullptr, LocalFlux *fx = nullptr) {
    const WgkArrays &a = p.a;
    const size_t i = mi(p, m, r);
    const size_t q = qi(p, m, r);
    const int tid = threadIdx.x & (VBLOCK - 1);  // column of the (128-thread) staging block handed in by the kernel
    WGK_VT_BEGIN();
    const size_t bs = band_stride(p);  // distance between two bands of the member's snow column
    double *__restrict__ S = a.snow_bands + bi(p, m, r, WGK_NBAND_K) + bs;
    const int32_t *__restrict__ E = a.s_elev32 + r + p.stride;

    // Every input of the head is loaded here, unconditionally and before the first store: one memory round
    // trip instead of one per use (the compiler may not move a load above a store or an early return).
    const int in_contcell = a.contcell[r], in_tbc = a.toBeCalculated[r], started = a.status_laf_next[i];
    const double in_laf = a.land_area_frac[i], in_laf_next = a.land_area_frac_next[i], in_laf_prev = a.land_area_frac_prev[i];
    const float4 f = p.forcing[fi(p, slot, m, r)];
    const int lc = a.landcover[r] - 1;
    const double P_T_SNOWFZ = a.p_snowfz[q];
    const double P_T_SNOWMT = a.p_snowmt[q];
    const double P_T_GRADNT = a.p_gradnt[q];
    const double in_degday = a.p_degday[q];
    const int in_snowfree = a.s_snowfree[i], in_de_max = a.s_de_max[r], in_de_min = a.s_de_min[r];
    const double in_p_prec = a.p_prec[q], in_ptc_ari = a.p_ptc_ari[q], in_ptc_hum = a.p_ptc_hum[q], in_pet_mxdy = a.p_pet_mxdy[q];
    const int in_arid = a.arid[r];
    const float in_laimax = a.laimax[q];
    const int in_lai_days = a.lai_days[i], in_lai_status = a.lai_status[i];
    const double in_lai_precsum = a.lai_precsum[i], in_snow = a.snow[i], in_netrad = a.p_netrad[q], in_cfa = a.cfa[q];
    const double in_canopy = a.canopy[i], in_mcwh = a.p_mcwh[q];
    const int elev0 = a.s_elev32[r];
    if (!in_contcell) return false;  // integrateWGHM.cpp:772

    // daily.cpp:159-169, routing.h:246-251
    const double landAreaFrac = (0 == started) ? in_laf : in_laf_next;
    double lafPrev;
    if (1 == started) lafPrev = in_laf_prev;
    else if (p.restart == 1) lafPrev = in_laf_prev;
    else lafPrev = landAreaFrac;

    if (1 != in_tbc) return false;  // daily.cpp:177
    WGK_VT(0);

    // second round: the land-cover tables (18 entries each, cache resident)
    const double ddf = in_degday * a.lct_ddf[lc];  // (M_DEGDAY_F * ddf_lct) * (...) keeps the reference association
    const float in_fa = a.lai_factor_a[lc], in_fb = a.lai_factor_b[lc];
    const int in_initial_days = a.lai_initial_days[lc];
    const double in_kc_min = a.lai_kc_min[lc], in_kc_max = a.lai_kc_max[lc], in_albedo_snow = a.lct_albedo_snow[lc], in_emissivity = a.lct_emissivity[lc];
    const bool noland = (landAreaFrac <= 0.);
    // cells without snow whose lowest and highest band are above the freezing threshold take no part
    // in the band loop (see the note on the band-parallel form below); the band columns of the others
    // start streaming now (bands 1..100; band 0 is the unused mean slot)
    bool bare = false;
    if (in_snowfree) {
        const double t_top = (double)f.y - ((double)in_de_max * P_T_GRADNT), t_bot = (double)f.y - ((double)in_de_min * P_T_GRADNT);
        bare = noland || (ddf >= 0. && t_top > P_T_SNOWFZ && t_bot > P_T_SNOWFZ);
    }
    WGK_VT_CLASS(!bare);
    WGK_VT(1);
    if (!bare) {
#pragma unroll
        for (int b = 0; b < SNOW_NBUF; b++) stage_issue(st, b, tid, S + (size_t)b * SNOW_CH * bs, E + (size_t)b * SNOW_CH * p.stride, bs, p.stride);
    }

    double dailyPrec = (double)f.x;
    const double dailyTempC = (double)f.y;
    const double dailyShortWave = (double)f.z;
    const double dailyLongWave = (double)f.w;

    dailyPrec = in_p_prec * dailyPrec;  // :248
    const double temp2 = dailyTempC + 237.3;
    const double e_s = 0.6108 * exp(17.27 * dailyTempC / temp2);

    // arid / humid (:331-348); any other index value is rejected on upload
    const bool arid_gw = (in_arid == 1);
    const double alpha = arid_gw ? in_ptc_ari : in_ptc_hum;
    const double maxDailyPET = in_pet_mxdy;

    // LAI / Kc (:355-356); LAImin in float arithmetic as in lai.cpp:154
    const float laimax_f = in_laimax;
    const float LAImin_f = __fadd_rn(in_fa, __fmul_rn(in_fb, laimax_f));
    const double LAImin = (double)LAImin_f;
    const double LAImaxd = (double)laimax_f;
    int days = in_lai_days, status = in_lai_status;
    double precsum = in_lai_precsum;
    double dailyLai;
    if (dailyTempC > 8.)
        dailyLai = lai_growing(days, in_initial_days, status, lc + 1, arid_gw, LAImin, LAImaxd, precsum, dailyPrec);
    else
        dailyLai = lai_nogrowing(days, in_initial_days, status, LAImin, LAImaxd, precsum, dailyPrec);
    a.lai_days[i] = days;
    a.lai_status[i] = status;
    a.lai_precsum[i] = precsum;
    double dailyKc;
    if ((LAImaxd - LAImin) == 0.) dailyKc = in_kc_min;
    else dailyKc = in_kc_min + (in_kc_max - in_kc_min) * (dailyLai - LAImin) / (LAImaxd - LAImin);

    const double snow_prev = in_snow;
    double albedo;
    if (snow_prev > 3.) albedo = in_albedo_snow;  // :366
    else albedo = 0.23;                                   // use_kc == 1

    double lat_heat;
    if (dailyTempC > 0) lat_heat = 2.501 - 0.002361 * dailyTempC;
    else lat_heat = 2.835;

    const double conv_Wm2_to_mmd = 0.0864 / lat_heat;
    const double solar_rad = conv_Wm2_to_mmd * dailyShortWave;
    const double long_wave_rad_in = conv_Wm2_to_mmd * dailyLongWave;
    const double emissivity = in_emissivity;
    const double temp_K = dailyTempC + 273.2;
    const double stefan_boltz_const = 0.000000004903;
    const double temp_K2 = temp_K * temp_K;  // pow(temp_K, 4.) (:423) as two squarings (<= 1 ulp apart)
    const double long_wave_rad_out = emissivity * stefan_boltz_const * (temp_K2 * temp_K2) / lat_heat;
    const double net_long_wave_rad = long_wave_rad_in - long_wave_rad_out;
    const double net_short_wave_rad = solar_rad * (1. - albedo);
    const double net_rad = in_netrad * (net_short_wave_rad + net_long_wave_rad);
    const double openWaterNetShortWaveRad = solar_rad * (1. - 0.08);
    const double openWaterNetRad = openWaterNetShortWaveRad + net_long_wave_rad;

    double dailyPET, dailyOpenWaterPET;
    const double inc_svp = 4098. * e_s / (temp2 * temp2);
    const double c3 = 0.0016286 * 101.3;
    const double gamma = c3 / lat_heat;
    if (net_rad <= 0.) dailyPET = 0.;
    else dailyPET = alpha * (inc_svp * net_rad) / (inc_svp + gamma);
    if (openWaterNetRad <= 0.) dailyOpenWaterPET = 0.;
    else dailyOpenWaterPET = alpha * (inc_svp * openWaterNetRad) / (inc_svp + gamma);
    if (snow_prev <= 3.) {  // :768-771
        dailyPET *= dailyKc;
        dailyOpenWaterPET *= 1.05;
    }
    const double cfa = in_cfa;
    a.lake_balance[i] = (dailyPrec - dailyOpenWaterPET) * cfa;
    a.openwater_prec[i] = dailyPrec;
    a.openwater_pet[i] = dailyOpenWaterPET;

    double landStorageChangeSum = 0., initialStorage = 0.;
    double dailyCanopyEvapo = 0., daily_prec_to_soil = 0., dailySoilPET = 0.;
    double dailySnowEvapo = 0., dailyEffPrec = 0.;
    double immediate_runoff = 0., dailyAET = 0., daily_runoff = 0., total_daily_runoff = 0.;
    double daily_gw_recharge = 0., pot_gw_recharge = 0.;
    double soil_water_overflow = 0., neg_land_aet = 0.;
    double storage_transfer = 0.;
    double land_aet = 0., land_aet_uncorr = 0.;


    // interception (:825-894)
    double canopy = in_canopy;
    if (noland) {
        storage_transfer = canopy;
        canopy = 0.;
        dailyCanopyEvapo = 0.;
    } else {
        canopy *= lafPrev / landAreaFrac;
        if (fabs(canopy) <= MIN_STOR_VOL) canopy = 0.;
        initialStorage = canopy;
        if (dailyLai > 0.00001) {
            const double max_canopy_storage = in_mcwh * dailyLai;
            const double canopy_deficiency = max_canopy_storage - canopy;
            if (dailyPrec < canopy_deficiency) {
                canopy += dailyPrec;
                daily_prec_to_soil = 0.;
            } else {
                canopy = max_canopy_storage;
                daily_prec_to_soil = dailyPrec - canopy_deficiency;
            }
            const double canopy_water_content = canopy;
            dailyCanopyEvapo = dailyPET * wg_pow((canopy_water_content / max_canopy_storage), 0.66666666);
            if (dailyCanopyEvapo > canopy_water_content) {
                dailyCanopyEvapo = canopy_water_content;
                dailySoilPET = dailyPET - canopy_water_content;
                canopy = 0.0;
            } else {
                canopy -= dailyCanopyEvapo;
                dailySoilPET = dailyPET - dailyCanopyEvapo;
            }
        } else {
            daily_prec_to_soil = dailyPrec;
            dailySoilPET = dailyPET;
            dailyCanopyEvapo = 0.0;
        }
        landStorageChangeSum += canopy - initialStorage;
    }
    a.canopy[i] = canopy;
    if (dailySoilPET < 0.) dailySoilPET = 0.0;

    WGK_VT(2);
    float e_builtup = 0.f, e_smax = 0.f, e_gwfactor = 0.f;
    double e_soil = 0., e_gamma = 0., e_pcrit = 0., e_transfer_old = 0.;
    short e_rgmax = 0;
    int e_texture = 0, e_ldd = 0;
    if (EARLY) {
        e_builtup = a.builtup[r]; e_smax = a.smax[q]; e_gwfactor = a.gwfactor[q];
        e_soil = a.soil[i]; e_gamma = a.gamma_hbv[q]; e_pcrit = a.p_pcrit[q];
        e_rgmax = a.rgmax[q];
        e_texture = a.texture[r]; e_ldd = a.ldd[r];
        e_transfer_old = a.storage_transfer[i];
        if (li) *li = local_load(p, r, m);
    }
    // snow in 100 elevation bands (:913-1062)
    double TempElevMax = 0., snowStorageChange = 0., snow = 0.;
    int nz = 0;
    if (bare) {
        // the additions the band loop performs on all-zero bands
        if (noland) {
            storage_transfer += 0.;
        } else {
            const double x = daily_prec_to_soil + 0.;
#pragma unroll 10
            for (int e = 0; e < 100; e++) dailyEffPrec += x;
            dailyEffPrec /= 100.;
            TempElevMax = dailyTempC - ((a.s_elev32[(size_t)p.stride + r] - elev0) * P_T_GRADNT);
        }
    } else if (noland) {
        __pipeline_wait_prior(0);
#pragma unroll 4
        for (int e = 1; e < 101; e++) {
            storage_transfer += *S / C100;
            *S = 0.;
            S += bs;
        }
        snow = 0.;
    } else {
        // S * lafPrev / landAreaFrac (:947): the quotient is formed with one correctly rounded
        // reciprocal and a fused residual correction (Markstein), 4 instructions instead of a
        // ~25-instruction division per band; the result is the correctly rounded quotient
        const double inv_laf = 1. / landAreaFrac;
        int thresh_elev = 0;
#if WGK_BAND_FORM >= 1
        // Staged form: the bands of a chunk go through every stage together (rescale, 1000 mm rule, temperatures, the cold and
        // the warm branch as selects, the ordered sums), so that the instruction stream the scheduler sees is already
        // interleaved: the ~15 dependent FP64 operations of a band overlap with those of the other bands of the chunk, and
        // the only band-to-band dependencies left are the integer select of the 1000 mm rule and the four running sums
        // (same additions in the same order as the band-by-band form: same bits).
        // Form 2 additionally trims the instruction count of a band (the loop is half of all instructions of a cell-day, and
        // the FP64 instructions take two issue slots each):
        //  * the rare 1000 mm rule behind ONE test per chunk on the largest high word of the rescaled bands (a band can only
        //    exceed 1000 mm if its high word reaches that of 1000.0; the exact comparison is made on the rare path);
        //  * "x = c ? (o ? a - b : 0) : y" as ONE subtraction of a selected subtrahend: below the freezing threshold the new storage
        //    is (s + P) - sublimation with sublimation = o ? PET : s + P, and (s + P) - (s + P) is the +0 the branch assigns; above
        //    the melting threshold it is s - melt with melt = all ? s : m, and s - s = +0; where the branch is not taken the
        //    subtrahend is +0 and x - 0 = x.  Same bits, four selects and no extra subtraction per band less;
        //  * the "any band non-zero" flag from the bits of the result instead of an FP64 compare.
        for (int c = 0; c < SNOW_NCH; c++) {
            const int buf = c % SNOW_NBUF;
            stage_wait(SNOW_NCH - 1 - c);
            double s0[SNOW_CH], sv[SNOW_CH], tv[SNOW_CH], ev[SNOW_CH], mv[SNOW_CH];
            int el[SNOW_CH];
#if WGK_BAND_FORM == 2
            int hmax = 0;
#endif
#pragma unroll
            for (int k = 0; k < SNOW_CH; k++) {
                el[k] = st->e[buf][k][tid];
                const double num = st->s[buf][k][tid] * lafPrev;
                double s = num * inv_laf;
                s = fma(fma(-landAreaFrac, s, num), inv_laf, s);
                s0[k] = (fabs(s) <= MIN_STOR_VOL) ? 0. : s;
#if WGK_BAND_FORM == 2
                { const int hw_ = hi_word(s0[k]); hmax = hw_ > hmax ? hw_ : hmax; }
#endif
            }
#if WGK_BAND_FORM == 2
            if (hmax >= 0x408F4000 || thresh_elev != 0)  // (0x408F4000 = high word of 1000.0) rare: the rule stays off the common path
#endif
#pragma unroll
            for (int k = 0; k < SNOW_CH; k++) {  // the 1000 mm rule (:958-976): integer selects only
                const bool big = s0[k] > 1000.;
                const bool first = big && thresh_elev == 0;
                const int e_eff = (big && !first && thresh_elev > 0) ? thresh_elev : el[k];
                thresh_elev = first ? el[k] : thresh_elev;
                el[k] = e_eff;
            }
#pragma unroll
            for (int k = 0; k < SNOW_CH; k++) tv[k] = dailyTempC - ((el[k] - elev0) * P_T_GRADNT);
            if (c == 0) TempElevMax = tv[0];
#if WGK_BAND_FORM == 2
#pragma unroll
            for (int k = 0; k < SNOW_CH; k++) {  // accumulation and sublimation below the freezing threshold (:982-999)
                const bool cold = tv[k] <= P_T_SNOWFZ;
                const double s_in = s0[k] + daily_prec_to_soil;
                const double sub = (s_in > dailySoilPET) ? dailySoilPET : s_in;
                ev[k] = cold ? sub : 0.;
                sv[k] = (cold ? s_in : s0[k]) - ev[k];
                mv[k] = cold ? 0. : daily_prec_to_soil;
            }
#pragma unroll
            for (int k = 0; k < SNOW_CH; k++) {  // melt above the melting threshold (:1003-1019)
                const double s = sv[k];
                const bool melt = tv[k] > P_T_SNOWMT && !(s < 0.);
                const double m_raw = ddf * (tv[k] - P_T_SNOWMT);
                const double mm = melt ? ((m_raw > s) ? s : m_raw) : 0.;
                mv[k] += mm;
                sv[k] = s - mm;
            }
#pragma unroll
            for (int k = 0; k < SNOW_CH; k++) {
                dailySnowEvapo += ev[k];
                snowStorageChange += sv[k] - s0[k];
                snow += sv[k];
                dailyEffPrec += mv[k];
                nz |= (hi_word(sv[k]) & 0x7fffffff) | lo_word(sv[k]);  // != 0. on the bits (-0. counts as zero)
                S[(size_t)(c * SNOW_CH + k) * bs] = sv[k];
            }
#else
#pragma unroll
            for (int k = 0; k < SNOW_CH; k++) {  // accumulation and sublimation below the freezing threshold (:982-999)
                const bool cold = tv[k] <= P_T_SNOWFZ;
                const double s_in = s0[k] + daily_prec_to_soil;
                const bool over = s_in > dailySoilPET;
                ev[k] = cold ? (over ? dailySoilPET : s_in) : 0.;
                sv[k] = cold ? (over ? s_in - dailySoilPET : 0.) : s0[k];
                mv[k] = cold ? 0. : daily_prec_to_soil;
            }
#pragma unroll
            for (int k = 0; k < SNOW_CH; k++) {  // melt above the melting threshold (:1003-1019)
                const double s = sv[k];
                const bool melt = tv[k] > P_T_SNOWMT && !(s < 0.);
                const double m_raw = ddf * (tv[k] - P_T_SNOWMT);
                const bool all = m_raw > s;
                mv[k] += melt ? (all ? s : m_raw) : 0.;
                sv[k] = melt ? (all ? 0. : s - m_raw) : s;
            }
#pragma unroll
            for (int k = 0; k < SNOW_CH; k++) {
                dailySnowEvapo += ev[k];
                snowStorageChange += sv[k] - s0[k];
                snow += sv[k];
                dailyEffPrec += mv[k];
                nz |= (sv[k] != 0.);
                S[(size_t)(c * SNOW_CH + k) * bs] = sv[k];
            }
#endif
            if (c + SNOW_NBUF < SNOW_NCH)
                stage_issue(st, buf, tid, S + (size_t)(c + SNOW_NBUF) * SNOW_CH * bs, E + (size_t)(c + SNOW_NBUF) * SNOW_CH * p.stride, bs, p.stride);
        }
#else
        for (int c = 0; c < SNOW_NCH; c++) {
            const int buf = c % SNOW_NBUF;
            stage_wait(SNOW_NCH - 1 - c);
#pragma unroll
            for (int k = 0; k < SNOW_CH; k++) {
                const int elev_e = st->e[buf][k][tid];
                const double sraw = st->s[buf][k][tid];
                double temp_elev = dailyTempC - ((elev_e - elev0) * P_T_GRADNT);
                const double num = sraw * lafPrev;
                double s = num * inv_laf;
                s = fma(fma(-landAreaFrac, s, num), inv_laf, s);
                if (fabs(s) <= MIN_STOR_VOL) s = 0.;
                const double s0 = s;
                {   // the 1000 mm rule (:958-976, rare): the first band above 1000 mm fixes the elevation whose temperature all
                    // later bands above 1000 mm use.  Branch-free like the rest of the band, so that the unrolled bands of a
                    // chunk form one basic block (21.8 vs 22.2 ms per simulated year).
                    const bool big = s > 1000.;
                    const bool first = big && thresh_elev == 0;
                    const double t_thresh = dailyTempC - ((thresh_elev - elev0) * P_T_GRADNT);
                    temp_elev = (big && !first && thresh_elev > 0) ? t_thresh : temp_elev;
                    thresh_elev = first ? elev_e : thresh_elev;
                }
                // Accumulation and sublimation below the freezing threshold (:982-999), melt above the melting threshold
                // (:1003-1019), written as selects: without branches the compiler interleaves the unrolled bands of a chunk and
                // the dependent FP64 operations of a band overlap with those of its neighbours (the bands only meet in the
                // ordered sums).  Measured against short if-bodies: 23.3 vs 24.5 ms per simulated year (one member), 869 vs
                // 894 ms (64 members); a two-stage form (all bands of a chunk rescaled first, one test for the rare 1000 mm
                // rule per chunk) was slower again (26.3 ms).  Adding +0. leaves a non-negative sum unchanged.
                const bool cold = temp_elev <= P_T_SNOWFZ;
                const double s_in = s + daily_prec_to_soil;
                const bool over = s_in > dailySoilPET;
                dailySnowEvapo += cold ? (over ? dailySoilPET : s_in) : 0.;
                s = cold ? (over ? s_in - dailySoilPET : 0.) : s;
                double effmelt = cold ? 0. : daily_prec_to_soil;
                const bool melt = temp_elev > P_T_SNOWMT && !(s < 0.);
                const double m_raw = ddf * (temp_elev - P_T_SNOWMT);
                const bool all = m_raw > s;
                effmelt += melt ? (all ? s : m_raw) : 0.;
                s = melt ? (all ? 0. : s - m_raw) : s;
                snowStorageChange += s - s0;
                if (c == 0 && k == 0) TempElevMax = temp_elev;
                snow += s;
                nz |= (s != 0.);
                dailyEffPrec += effmelt;
                S[(size_t)(c * SNOW_CH + k) * bs] = s;
            }
            // refill the buffer just consumed with the chunk after next (same thread: no barrier)
            if (c + SNOW_NBUF < SNOW_NCH)
                stage_issue(st, buf, tid, S + (size_t)(c + SNOW_NBUF) * SNOW_CH * bs, E + (size_t)(c + SNOW_NBUF) * SNOW_CH * p.stride, bs, p.stride);
        }
#endif
        snow /= 100.;
        dailyEffPrec /= 100.;
        dailySnowEvapo /= 100.;
        snowStorageChange /= 100.;
        landStorageChangeSum += snowStorageChange;
    }
    a.snow[i] = snow;
    if (!bare) a.s_snowfree[i] = (int8_t)(nz == 0);
    WGK_VT(3);

    double gw_recharge_out = 0.;
    // inputs of the soil part: one round of loads
    const float builtup = EARLY ? e_builtup : a.builtup[r], in_smax = EARLY ? e_smax : a.smax[q], gwFactor = EARLY ? e_gwfactor : a.gwfactor[q];
    const double in_soil = EARLY ? e_soil : a.soil[i], in_gamma = EARLY ? e_gamma : a.gamma_hbv[q], in_pcrit = EARLY ? e_pcrit : a.p_pcrit[q];
    const short Rgmax = EARLY ? e_rgmax : a.rgmax[q];
    const int in_texture = EARLY ? e_texture : a.texture[r], in_ldd = EARLY ? e_ldd : a.ldd[r];
    const double in_transfer_old = EARLY ? e_transfer_old : a.storage_transfer[i];
    if (li && !EARLY) *li = local_load(p, r, m);
    // immediate runoff (:1068-1071)
    if (builtup > 0.) {
        immediate_runoff = 0.5 * dailyEffPrec * builtup;
        dailyEffPrec -= immediate_runoff;
    }
    WGK_VT(4);

    // soil and AET (:1080-1239)
    const double Smax = (double)in_smax;
    double soil = in_soil;
    if (noland) {
        storage_transfer += soil;
        storage_transfer *= cfa;
        soil = 0.;
        daily_gw_recharge = 0.;
        total_daily_runoff = 0.;
        a.gw_recharge[i] = 0.;
        a.storage_transfer[i] = storage_transfer;
        gw_recharge_out = 0.;
    } else {
        soil *= lafPrev / landAreaFrac;
        initialStorage = soil;
        soil_water_overflow = 0;
        if (soil > Smax) {
            soil_water_overflow = soil - Smax;
            soil = Smax;
        }
        if (TempElevMax > P_T_SNOWFZ) {
            if (Smax > 0.) {
                const double soil_saturation = soil / Smax;
                daily_runoff = dailyEffPrec * wg_pow(soil_saturation, in_gamma);
                if (dailySoilPET > (maxDailyPET - dailyCanopyEvapo) * soil_saturation)
                    dailyAET = (maxDailyPET - dailyCanopyEvapo) * soil_saturation;
                else
                    dailyAET = dailySoilPET;
                soil += dailyEffPrec - dailyAET - daily_runoff;
                if (fabs(soil) <= MIN_STOR_VOL) soil = 0.;
                dailyEffPrec = 0.;
                if (soil < 0.) {
                    dailyAET += soil;
                    soil = 0.;
                }
                daily_runoff *= cfa;
                immediate_runoff *= cfa;
                if ((Rgmax / C100) < (gwFactor * daily_runoff)) daily_gw_recharge = Rgmax / C100;
                else daily_gw_recharge = gwFactor * daily_runoff;
                pot_gw_recharge = 0.;
                if (((arid_gw) && (in_texture < 21)) && (in_ldd >= 0)) {  // :1165-1176
                    if (dailyPrec <= in_pcrit) {
                        pot_gw_recharge = daily_gw_recharge;
                        daily_gw_recharge = 0.;
                    }
                }
                daily_runoff -= pot_gw_recharge;
                pot_gw_recharge /= cfa;
                soil += pot_gw_recharge;
                if (soil > Smax) {
                    soil_water_overflow += soil - Smax;
                    soil = Smax;
                }
                soil_water_overflow *= cfa;
                total_daily_runoff = daily_runoff + immediate_runoff + soil_water_overflow;
            } else {
                total_daily_runoff = 0.;
                daily_gw_recharge = 0.;
            }
        } else {
            soil_water_overflow *= cfa;
            dailyEffPrec *= cfa;
            total_daily_runoff += soil_water_overflow + dailyEffPrec;
            daily_gw_recharge = 0.;
            dailyAET = 0.;
        }
        a.gw_recharge[i] = daily_gw_recharge;  // :1221 (not updated by the fix-up below)
        gw_recharge_out = daily_gw_recharge;
        landStorageChangeSum += soil - initialStorage;
        land_aet = landStorageChangeSum * (cfa - 1.0) - dailyPrec * (cfa - 1.0)
                   + (dailyAET + dailyCanopyEvapo + dailySnowEvapo) * cfa;
        if (land_aet < 0.) {
            neg_land_aet = land_aet;
            land_aet = 0.;
        }
        land_aet_uncorr = (dailyAET + dailyCanopyEvapo + dailySnowEvapo);
    }
    a.land_aet[i] = land_aet;
    a.land_aet_uncorr[i] = land_aet_uncorr;

    // surface runoff (:1244-1257)
    if (neg_land_aet < 0.) total_daily_runoff = total_daily_runoff + neg_land_aet;
    if (total_daily_runoff < 0.) total_daily_runoff = 0.;
    if ((total_daily_runoff - daily_gw_recharge) < 0.) {
        const double neg_runoff = total_daily_runoff - daily_gw_recharge;
        daily_gw_recharge = total_daily_runoff;
        soil += neg_runoff;
    }
    a.soil[i] = soil;
    a.surface_runoff[i] = total_daily_runoff - daily_gw_recharge;
    if (fx) {
        fx->owPrec = dailyPrec;
        fx->owPET = dailyOpenWaterPET;
        fx->storage_transfer = noland ? storage_transfer : in_transfer_old;  // G_dailyStorageTransfer keeps its last value (:1113)
        fx->surface_runoff = total_daily_runoff - daily_gw_recharge;
        fx->gw_recharge = gw_recharge_out;
    }
    WGK_VT(5);
    WGK_VT_FLUSH(r < p.level_off[1]);
    return true;
}

// ----------------------------------------------------------------------------------------
// vertical water balance (daily.cpp:94-1264), CTA-cooperative and band-parallel
//
// One CTA of NW warps works on a TILE of 32 consecutive cells (lane = cell):
//   v_mode   warp 0   per-cell regime (off / no land / bare and warm / needs the band loop) and the
//                     scalars the band threads need that do not depend on the head
//   v_head   warp 0   forcing, LAI, radiation, PET, interception          (daily.cpp:186-894)
//   slabs    all      the 100 snow bands in NSLAB slabs of SLAB bands; inside a slab warp w owns
//                     V_BPW consecutive bands of the 32 cells (coalesced 256 B rows), the loads of
//                     the next slab are in flight while the current one is evaluated     (:913-1062)
//   v_sum    warps 0-3  the four band sums (storage change, effective precipitation, sublimation,
//                     snow) are accumulated IN BAND ORDER from per-band contributions staged in
//                     shared memory: the value is the one the serial loop of the reference produces
//   v_tail   warp 0   immediate runoff, soil, AET, runoff split                        (:1068-1257)
// so the dependent instruction chain of one cell-day is ~1/4 of the one-thread-per-cell form; this
// chain (cell-day -> cell-day) is what bounds a single-member run.
//
// Cells whose 100 bands are all exactly zero (s_snowfree, maintained by v_sum) and whose lowest and
// highest band are both above the freezing threshold take no part in the band loop: every band
// would read 0, add 0 and write 0 (the elevation-band temperature is monotone in the elevation, so
// checking the two extreme bands covers all of them).  The sums of such a cell are formed by the
// same additions the loop would do (100 x "+= precipitation"), so results are unchanged; a tile
// made of such cells only never touches the band arrays.
// ----------------------------------------------------------------------------------------
// NW warps per tile, BPW bands per warp and slab; PRE: inputs of head / tail / local routing preloaded into
// shared memory by the warps that idle while warp 0 determines the regimes
template <int NW_, int BPW_, bool PRE_>
struct VCfg {
    static constexpr int NW = NW_, BPW = BPW_, SLAB = NW_ * BPW_, NSLAB = 100 / (NW_ * BPW_), THREADS = 32 * NW_;
    static constexpr int NQ = (4 + NW_ - 1) / NW_;  // band sums per warp
    static constexpr bool PRE = PRE_;
    static_assert(SLAB * NSLAB == 100, "slabs must tile the 100 bands");
    static_assert(!PRE_ || NW_ >= 2, "the preload needs a warp besides warp 0");
};
using VCfgSmall = VCfg<5, 4, true>;  // 5 threads per cell: shortest chain, for problems far too small to fill the GPU
using VCfgMid = VCfg<2, 5, false>;   // 2 threads per cell: all tiles of a 0.5 degree member resident at once
enum { VM_ACTIVE = 1, VM_NOLAND = 2, VM_BARE = 4 };
enum { Q_CHG = 0, Q_EFF = 1, Q_SUB = 2, Q_SNOW = 3 };

// slots of the preloaded per-cell inputs (VTile::pd / pf / pk)
enum { LI_ekg, LI_invkg, LI_evaredex, LI_area, LI_cfa, LI_contf, LI_fswb_init, LI_loc_lake, LI_loc_wetland, LI_kS, LI_lake_depth,
       LI_wetl_depth, LI_laf, LI_laf_prev, LI_gw, LI_loc_lake_stor, LI_red_loc_lake, LI_loc_wetl_stor, LI_red_loc_wetl,
       LI_raf_next, LI_transfer_old, TI_soil, TI_gamma, TI_pcrit, HI_p_prec, HI_ptc_ari, HI_ptc_hum, HI_lai_precsum, HI_snow, HI_netrad,
       HI_canopy, HI_mcwh, HI_pet_mxdy, VP_ND };
enum { TI_builtup, TI_smax, TI_gwfactor, HI_laimax, VP_NF };
enum { K_contcell, K_flags, K_ldd, K_arid, K_texture, K_rgmax, K_lc, K_lai_days, K_lai_status, VP_NK };

template <class C>
struct VTile {
    // band-loop scalars of the tile's cells
    double T[32], grad[32], lafPrev[32], laf[32], inv_laf[32], fz[32], mt[32], ddf[32], prec[32], pet[32];
    int elev0[32], mode[32];
    int cap_first[32], cap_elev[32], cap_any[C::NSLAB], nband_cells;
    double temp1[32];  // temperature of band 1 (TempElevMax, daily.cpp:1030)
    double contrib[4][C::SLAB][32], acc_init[32], fin[4][32];
    int nz[32];
    // head -> tail
    double h_prec[32], h_cfa[32], h_canopy_evapo[32], h_lsc[32], h_maxpet[32], h_owpet[32];
    // inputs of the head, the tail and the local routing, brought in by warps 1-4 (cp.async) while
    // warp 0 determines the regimes: after the first barrier warp 0 computes from shared memory only
    double pd[C::PRE ? VP_ND : 1][32];
    float pf[C::PRE ? VP_NF : 1][32];
    float4 pforce[C::PRE ? 32 : 1];
    int pk[C::PRE ? VP_NK : 1][32];
};

// registers of one thread that live across the barriers of a tile
template <class C>
struct VThread {
    double pre_s[C::BPW], cur_s[C::BPW], acc[C::NQ];
    int pre_e[C::BPW], cur_e[C::BPW], nz;
};
// input of the scalar phases: from the preloaded shared-memory slot or straight from global memory
#define VIN_D(slot_, expr_) (C::PRE ? sm.pd[C::PRE ? (slot_) : 0][lane] : (double)(expr_))
#define VIN_F(slot_, expr_) (C::PRE ? sm.pf[C::PRE ? (slot_) : 0][lane] : (float)(expr_))
#define VIN_K(slot_, expr_) (C::PRE ? sm.pk[C::PRE ? (slot_) : 0][lane] : (int)(expr_))

// regime of the cell and head-independent scalars (warp 0, lane = cell)
template <class C>
__device__ __forceinline__ void v_mode(const WgkParams &p, VTile<C> &sm, const int r0, const int begin, const int end, const int m,
                                       const int slot, const int lane) {
    const WgkArrays &a = p.a;
    const int r = r0 + lane;
    if (lane == 0) sm.nband_cells = 0;
    if (lane < C::NSLAB) sm.cap_any[lane] = 0;
    sm.cap_first[lane] = 1000;
    sm.cap_elev[lane] = 0;
#ifndef WGK_EMU
    __syncwarp();
#endif
    int mode = 0;
    if (r >= begin && r < end && a.contcell[r]) {  // integrateWGHM.cpp:772
        const size_t i = mi(p, m, r);
        const size_t q = qi(p, m, r);
        // daily.cpp:159-169, routing.h:246-251 (all candidates are loaded at once: one memory round trip)
        const int started = a.status_laf_next[i];
        const double laf_cur = a.land_area_frac[i], laf_next = a.land_area_frac_next[i], laf_prev = a.land_area_frac_prev[i];
        const double landAreaFrac = (0 == started) ? laf_cur : laf_next;
        double lafPrev;
        if (1 == started) lafPrev = laf_prev;
        else if (p.restart == 1) lafPrev = laf_prev;
        else lafPrev = landAreaFrac;
        if (1 == a.toBeCalculated[r]) {  // daily.cpp:177
            mode = VM_ACTIVE;
            const bool noland = (landAreaFrac <= 0.);
            if (noland) mode |= VM_NOLAND;
            const float4 f = p.forcing[fi(p, slot, m, r)];
            const double T = (double)f.y;
            const double grad = a.p_gradnt[q], fz = a.p_snowfz[q];
            const double ddf = a.p_degday[q] * a.lct_ddf[a.landcover[r] - 1];  // (M_DEGDAY_F * ddf_lct) * (...) keeps the reference association
            sm.T[lane] = T;
            sm.grad[lane] = grad;
            sm.fz[lane] = fz;
            sm.mt[lane] = a.p_snowmt[q];
            sm.ddf[lane] = ddf;
            sm.lafPrev[lane] = lafPrev;
            sm.laf[lane] = landAreaFrac;
            // S * lafPrev / landAreaFrac (:947): the quotient is formed with one correctly rounded
            // reciprocal and a fused residual correction (Markstein), 4 instructions instead of a
            // ~25-instruction division per band; the result is the correctly rounded quotient
            sm.inv_laf[lane] = noland ? 0. : 1. / landAreaFrac;
            sm.elev0[lane] = a.elevation[r];
            if (a.s_snowfree[i]) {
                const double t_top = T - ((double)a.s_de_max[r] * grad), t_bot = T - ((double)a.s_de_min[r] * grad);
                if (noland || (ddf >= 0. && t_top > fz && t_bot > fz)) mode |= VM_BARE;
            }
            if (!(mode & VM_BARE)) atomicAdd(&sm.nband_cells, 1);
        }
    }
    sm.mode[lane] = mode;
}

// issue the loads of slab `slab` (bands slab*C::SLAB + w*C::BPW + j + 1) into the prefetch registers
template <class C>
__device__ __forceinline__ void v_prefetch(const WgkParams &p, const VTile<C> &sm, VThread<C> &ts, const int r0, const int m, const int slab,
                                           const int w, const int lane) {
    const int mode = sm.mode[lane];
    const bool on = (mode & VM_ACTIVE) && !(mode & VM_BARE);
    const size_t r = (size_t)(r0 + lane);
    const int e0 = slab * C::SLAB + w * C::BPW + 1;
    const size_t bs = band_stride(p);
    const double *__restrict__ S = p.a.snow_bands + bi(p, m, (int)r, WGK_NBAND_K) + (size_t)e0 * bs;
    const int16_t *__restrict__ E = p.a.s_delev + (size_t)e0 * p.stride + r;
#pragma unroll
    for (int j = 0; j < C::BPW; j++) {
        ts.pre_s[j] = on ? S[(size_t)j * bs] : 0.;
        ts.pre_e[j] = on ? (int)E[(size_t)j * p.stride] : 0;
    }
}

// forcing, LAI, radiation, PET, interception (warp 0, lane = cell)
template <class C>
__device__ __forceinline__ void v_head(const WgkParams &p, VTile<C> &sm, const int r0, const int m, const int slot, const int lane) {
    const WgkArrays &a = p.a;
    const int mode = sm.mode[lane];
    if (!(mode & VM_ACTIVE)) return;
    const int r = r0 + lane;
    const size_t i = mi(p, m, r);
    const size_t q = qi(p, m, r);
    const double landAreaFrac = sm.laf[lane], lafPrev = sm.lafPrev[lane];
    const int lc = VIN_K(K_lc, a.landcover[r]) - 1;
    const float4 f = C::PRE ? sm.pforce[C::PRE ? lane : 0]
                            : p.forcing[fi(p, slot, m, r)];
    double dailyPrec = (double)f.x;
    const double dailyTempC = (double)f.y;
    const double dailyShortWave = (double)f.z;
    const double dailyLongWave = (double)f.w;

    dailyPrec = VIN_D(HI_p_prec, a.p_prec[q]) * dailyPrec;  // :248
    const double temp2 = dailyTempC + 237.3;
    const double e_s = 0.6108 * exp(17.27 * dailyTempC / temp2);

    // arid / humid (:331-348); any other index value is rejected on upload
    const bool arid_gw = (VIN_K(K_arid, a.arid[r]) == 1);
    const double alpha = arid_gw ? VIN_D(HI_ptc_ari, a.p_ptc_ari[q]) : VIN_D(HI_ptc_hum, a.p_ptc_hum[q]);

    // LAI / Kc (:355-356); LAImin in float arithmetic as in lai.cpp:154
    const float laimax_f = VIN_F(HI_laimax, a.laimax[q]);
    const float LAImin_f = __fadd_rn(a.lai_factor_a[lc], __fmul_rn(a.lai_factor_b[lc], laimax_f));
    const double LAImin = (double)LAImin_f;
    const double LAImaxd = (double)laimax_f;
    int days = VIN_K(K_lai_days, a.lai_days[i]), status = VIN_K(K_lai_status, a.lai_status[i]);
    double precsum = VIN_D(HI_lai_precsum, a.lai_precsum[i]);
    double dailyLai;
    if (dailyTempC > 8.)
        dailyLai = lai_growing(days, a.lai_initial_days[lc], status, lc + 1, arid_gw, LAImin, LAImaxd, precsum, dailyPrec);
    else
        dailyLai = lai_nogrowing(days, a.lai_initial_days[lc], status, LAImin, LAImaxd, precsum, dailyPrec);
    a.lai_days[i] = days;
    a.lai_status[i] = status;
    a.lai_precsum[i] = precsum;
    double dailyKc;
    if ((LAImaxd - LAImin) == 0.) dailyKc = a.lai_kc_min[lc];
    else dailyKc = a.lai_kc_min[lc] + (a.lai_kc_max[lc] - a.lai_kc_min[lc]) * (dailyLai - LAImin) / (LAImaxd - LAImin);

    const double snow_prev = VIN_D(HI_snow, a.snow[i]);
    double albedo;
    if (snow_prev > 3.) albedo = a.lct_albedo_snow[lc];  // :366
    else albedo = 0.23;                                   // use_kc == 1

    double lat_heat;
    if (dailyTempC > 0) lat_heat = 2.501 - 0.002361 * dailyTempC;
    else lat_heat = 2.835;

    const double conv_Wm2_to_mmd = 0.0864 / lat_heat;
    const double solar_rad = conv_Wm2_to_mmd * dailyShortWave;
    const double long_wave_rad_in = conv_Wm2_to_mmd * dailyLongWave;
    const double emissivity = a.lct_emissivity[lc];
    const double temp_K = dailyTempC + 273.2;
    const double stefan_boltz_const = 0.000000004903;
    const double temp_K2 = temp_K * temp_K;  // pow(temp_K, 4.) (:423) as two squarings (<= 1 ulp apart)
    const double long_wave_rad_out = emissivity * stefan_boltz_const * (temp_K2 * temp_K2) / lat_heat;
    const double net_long_wave_rad = long_wave_rad_in - long_wave_rad_out;
    const double net_short_wave_rad = solar_rad * (1. - albedo);
    const double net_rad = VIN_D(HI_netrad, a.p_netrad[q]) * (net_short_wave_rad + net_long_wave_rad);
    const double openWaterNetShortWaveRad = solar_rad * (1. - 0.08);
    const double openWaterNetRad = openWaterNetShortWaveRad + net_long_wave_rad;

    double dailyPET, dailyOpenWaterPET;
    const double inc_svp = 4098. * e_s / (temp2 * temp2);
    const double c3 = 0.0016286 * 101.3;
    const double gamma = c3 / lat_heat;
    if (net_rad <= 0.) dailyPET = 0.;
    else dailyPET = alpha * (inc_svp * net_rad) / (inc_svp + gamma);
    if (openWaterNetRad <= 0.) dailyOpenWaterPET = 0.;
    else dailyOpenWaterPET = alpha * (inc_svp * openWaterNetRad) / (inc_svp + gamma);
    if (snow_prev <= 3.) {  // :768-771
        dailyPET *= dailyKc;
        dailyOpenWaterPET *= 1.05;
    }
    const double cfa = VIN_D(LI_cfa, a.cfa[q]);
    a.lake_balance[i] = (dailyPrec - dailyOpenWaterPET) * cfa;
    a.openwater_prec[i] = dailyPrec;
    a.openwater_pet[i] = dailyOpenWaterPET;

    double landStorageChangeSum = 0.;
    double dailyCanopyEvapo = 0., daily_prec_to_soil = 0., dailySoilPET = 0.;
    // interception (:825-894)
    double canopy = VIN_D(HI_canopy, a.canopy[i]);
    double acc_init = 0.;
    if (mode & VM_NOLAND) {
        acc_init = canopy;  // storage_transfer starts with the canopy water (:829)
        canopy = 0.;
    } else {
        canopy *= lafPrev / landAreaFrac;
        if (fabs(canopy) <= MIN_STOR_VOL) canopy = 0.;
        const double initialStorage = canopy;
        if (dailyLai > 0.00001) {
            const double max_canopy_storage = VIN_D(HI_mcwh, a.p_mcwh[q]) * dailyLai;
            const double canopy_deficiency = max_canopy_storage - canopy;
            if (dailyPrec < canopy_deficiency) {
                canopy += dailyPrec;
                daily_prec_to_soil = 0.;
            } else {
                canopy = max_canopy_storage;
                daily_prec_to_soil = dailyPrec - canopy_deficiency;
            }
            const double canopy_water_content = canopy;
            dailyCanopyEvapo = dailyPET * wg_pow((canopy_water_content / max_canopy_storage), 0.66666666);
            if (dailyCanopyEvapo > canopy_water_content) {
                dailyCanopyEvapo = canopy_water_content;
                dailySoilPET = dailyPET - canopy_water_content;
                canopy = 0.0;
            } else {
                canopy -= dailyCanopyEvapo;
                dailySoilPET = dailyPET - dailyCanopyEvapo;
            }
        } else {
            daily_prec_to_soil = dailyPrec;
            dailySoilPET = dailyPET;
            dailyCanopyEvapo = 0.0;
        }
        landStorageChangeSum += canopy - initialStorage;
    }
    a.canopy[i] = canopy;
    if (dailySoilPET < 0.) dailySoilPET = 0.0;
    sm.prec[lane] = daily_prec_to_soil;
    sm.pet[lane] = dailySoilPET;
    sm.acc_init[lane] = acc_init;
    sm.h_prec[lane] = dailyPrec;
    sm.h_cfa[lane] = cfa;
    sm.h_canopy_evapo[lane] = dailyCanopyEvapo;
    sm.h_lsc[lane] = landStorageChangeSum;
    sm.h_maxpet[lane] = VIN_D(HI_pet_mxdy, a.p_pet_mxdy[q]);
    sm.h_owpet[lane] = dailyOpenWaterPET;
}

// take over the prefetched slab, start the next one, rescale to the new land area fraction and
// look for bands above the 1000 mm cap (:947-976)
template <class C>
__device__ __forceinline__ void v_scale(const WgkParams &p, VTile<C> &sm, VThread<C> &ts, const int r0, const int m, const int slab,
                                        const int w, const int lane) {
    const int mode = sm.mode[lane];
#pragma unroll
    for (int j = 0; j < C::BPW; j++) {
        ts.cur_s[j] = ts.pre_s[j];
        ts.cur_e[j] = ts.pre_e[j];
    }
    if (slab + 1 < C::NSLAB) v_prefetch<C>(p, sm, ts, r0, m, slab + 1, w, lane);
    if (!(mode & VM_ACTIVE) || (mode & VM_NOLAND)) return;
    const double lafPrev = sm.lafPrev[lane], laf = sm.laf[lane], inv_laf = sm.inv_laf[lane];
    bool capped = false;
#pragma unroll
    for (int j = 0; j < C::BPW; j++) {
        const double num = ts.cur_s[j] * lafPrev;
        double s = num * inv_laf;
        s = fma(fma(-laf, s, num), inv_laf, s);
        if (fabs(s) <= MIN_STOR_VOL) s = 0.;
        ts.cur_s[j] = s;
        if (s > 1000.) {
            capped = true;
            if (ts.cur_e[j] + sm.elev0[lane] != 0) atomicMin(&sm.cap_first[lane], slab * C::SLAB + w * C::BPW + j + 1);
        }
    }
    if (capped) sm.cap_any[slab] = 1;
}

// only when a band of the tile is above the cap: the first such band of a cell (in band order,
// with a non-zero elevation) fixes the elevation whose temperature all later capped bands use
template <class C>
__device__ __forceinline__ void v_cap_resolve(VTile<C> &sm, const VThread<C> &ts, const int slab, const int w, const int lane) {
#pragma unroll
    for (int j = 0; j < C::BPW; j++)
        if (slab * C::SLAB + w * C::BPW + j + 1 == sm.cap_first[lane]) sm.cap_elev[lane] = ts.cur_e[j] + sm.elev0[lane];
}

// the bands of one slab: accumulation, sublimation, melt (:978-1045); per-band contributions to
// the four ordered sums go to shared memory, the new band storage to global memory
template <class C>
__device__ __forceinline__ void v_band(const WgkParams &p, VTile<C> &sm, const VThread<C> &ts, const int r0, const int m, const int slab,
                                       const int w, const int lane) {
    const int mode = sm.mode[lane];
    const bool store = (mode & VM_ACTIVE) && !(mode & VM_BARE);
    const int k0 = w * C::BPW, e0 = slab * C::SLAB + k0 + 1;
    const size_t bs = band_stride(p);
    double *__restrict__ S = p.a.snow_bands + bi(p, m, r0 + lane, WGK_NBAND_K) + (size_t)e0 * bs;
    if (mode & VM_NOLAND) {  // :916-922
#pragma unroll
        for (int j = 0; j < C::BPW; j++) {
            sm.contrib[Q_CHG][k0 + j][lane] = ts.cur_s[j] / C100;
            sm.contrib[Q_EFF][k0 + j][lane] = 0.;
            sm.contrib[Q_SUB][k0 + j][lane] = 0.;
            sm.contrib[Q_SNOW][k0 + j][lane] = 0.;
            if (store) S[(size_t)j * bs] = 0.;
        }
        return;
    }
    const double T = sm.T[lane], grad = sm.grad[lane], fz = sm.fz[lane], mt = sm.mt[lane], ddf = sm.ddf[lane];
    const double prec = sm.prec[lane], pet = sm.pet[lane];
    const bool cap_slab = sm.cap_any[slab] != 0 || sm.cap_elev[lane] != 0;
#pragma unroll
    for (int j = 0; j < C::BPW; j++) {
        const int de = ts.cur_e[j];
        double temp_elev = T - ((double)de * grad);
        double s = ts.cur_s[j];
        const double s0 = s;
        if (cap_slab && s > 1000.) {  // :958-976
            const int ce = sm.cap_elev[lane];
            if (ce > 0 && e0 + j > sm.cap_first[lane]) temp_elev = T - ((double)(ce - sm.elev0[lane]) * grad);
        }
        // accumulation and sublimation (:982-999), melt (:1003-1019) as selects, so that the unrolled bands interleave
        // (see vertical_cell)
        const bool cold = temp_elev <= fz;
        const double s_in = s + prec;
        const bool over = s_in > pet;
        const double sub = cold ? (over ? pet : s_in) : 0.;
        s = cold ? (over ? s_in - pet : 0.) : s;
        const double eff = cold ? 0. : prec;
        const bool mlt = temp_elev > mt && !(s < 0.);
        const double m_raw = ddf * (temp_elev - mt);
        const bool all = m_raw > s;
        const double melt = mlt ? (all ? s : m_raw) : 0.;
        s = mlt ? (all ? 0. : s - m_raw) : s;
        if (slab == 0 && k0 + j == 0) sm.temp1[lane] = temp_elev;
        sm.contrib[Q_CHG][k0 + j][lane] = s - s0;
        sm.contrib[Q_EFF][k0 + j][lane] = eff + melt;
        sm.contrib[Q_SUB][k0 + j][lane] = sub;
        sm.contrib[Q_SNOW][k0 + j][lane] = s;
        if (store) S[(size_t)j * bs] = s;
    }
}

// ordered accumulation of the slab's contributions: warp w sums the quantities w, w + NW, ... of the 32 cells
template <class C>
__device__ __forceinline__ void v_sum(VTile<C> &sm, VThread<C> &ts, const int slab, const int w, const int lane) {
#pragma unroll
    for (int qi = 0; qi < C::NQ; qi++) {
        const int q = w + qi * C::NW;
        if (q >= 4) break;
        double acc = (slab == 0) ? ((q == Q_CHG) ? sm.acc_init[lane] : 0.) : ts.acc[qi];
        int nz = (slab == 0) ? 0 : ts.nz;
#pragma unroll
        for (int k = 0; k < C::SLAB; k++) {
            const double c = sm.contrib[q][k][lane];
            acc += c;
            if (q == Q_SNOW) nz |= (c != 0.);
        }
        ts.acc[qi] = acc;
        if (q == Q_SNOW) ts.nz = nz;
        if (slab == C::NSLAB - 1) {
            sm.fin[q][lane] = acc;
            if (q == Q_SNOW) sm.nz[lane] = nz;
        }
    }
}

// a tile without any cell in the band loop: the sums of its bare cells by the additions the loop
// would perform on all-zero bands
template <class C>
__device__ __forceinline__ void v_bare_sums(const WgkParams &p, VTile<C> &sm, const int r0, const int lane) {
    const int mode = sm.mode[lane];
    if (!(mode & VM_ACTIVE)) return;
    double eff = 0.;
    if (!(mode & VM_NOLAND)) {
        const double x = sm.prec[lane] + 0.;
#pragma unroll 10
        for (int e = 0; e < 100; e++) eff += x;
        sm.temp1[lane] = sm.T[lane] - ((double)p.a.s_delev[(size_t)p.stride + r0 + lane] * sm.grad[lane]);
    }
    sm.fin[Q_CHG][lane] = sm.acc_init[lane] + 0.;
    sm.fin[Q_EFF][lane] = eff;
    sm.fin[Q_SUB][lane] = 0.;
    sm.fin[Q_SNOW][lane] = 0.;
    sm.nz[lane] = 0;
}

// immediate runoff, soil, AET, runoff split (warp 0, lane = cell)
template <class C>
__device__ __forceinline__ void v_tail(const WgkParams &p, VTile<C> &sm, const int r0, const int m, const int lane, double flux[3]) {
    const WgkArrays &a = p.a;
    const int mode = sm.mode[lane];
    if (!(mode & VM_ACTIVE)) return;
    const int r = r0 + lane;
    const size_t i = mi(p, m, r);
    const size_t q = qi(p, m, r);
    const bool noland = (mode & VM_NOLAND) != 0;
    const double landAreaFrac = sm.laf[lane], lafPrev = sm.lafPrev[lane];
    const double dailyPrec = sm.h_prec[lane], cfa = sm.h_cfa[lane], dailyCanopyEvapo = sm.h_canopy_evapo[lane];
    const double dailySoilPET = sm.pet[lane], maxDailyPET = sm.h_maxpet[lane];
    const double P_T_SNOWFZ = sm.fz[lane];
    double landStorageChangeSum = sm.h_lsc[lane];
    double storage_transfer = 0., snow = 0., dailyEffPrec = 0., dailySnowEvapo = 0., TempElevMax = 0.;
    if (noland) {
        storage_transfer = sm.fin[Q_CHG][lane];
    } else {
        snow = sm.fin[Q_SNOW][lane] / C100;
        dailyEffPrec = sm.fin[Q_EFF][lane] / C100;
        dailySnowEvapo = sm.fin[Q_SUB][lane] / C100;
        landStorageChangeSum += sm.fin[Q_CHG][lane] / C100;
        TempElevMax = sm.temp1[lane];
    }
    a.snow[i] = snow;
    if (!(mode & VM_BARE)) a.s_snowfree[i] = (int8_t)(sm.nz[lane] == 0);

    double immediate_runoff = 0., dailyAET = 0., daily_runoff = 0., total_daily_runoff = 0.;
    double daily_gw_recharge = 0., pot_gw_recharge = 0.;
    double soil_water_overflow = 0., neg_land_aet = 0.;
    double land_aet = 0., land_aet_uncorr = 0.;

    // immediate runoff (:1068-1071)
    const float builtup = VIN_F(TI_builtup, a.builtup[r]);
    if (builtup > 0.) {
        immediate_runoff = 0.5 * dailyEffPrec * builtup;
        dailyEffPrec -= immediate_runoff;
    }

    // soil and AET (:1080-1239)
    const double Smax = (double)VIN_F(TI_smax, a.smax[q]);
    double soil = VIN_D(TI_soil, a.soil[i]);
    if (noland) {
        storage_transfer += soil;
        storage_transfer *= cfa;
        soil = 0.;
        daily_gw_recharge = 0.;
        total_daily_runoff = 0.;
        a.gw_recharge[i] = 0.;
        a.storage_transfer[i] = storage_transfer;
        flux[2] = 0.;
    } else {
        soil *= lafPrev / landAreaFrac;
        const double initialStorage = soil;
        soil_water_overflow = 0;
        if (soil > Smax) {
            soil_water_overflow = soil - Smax;
            soil = Smax;
        }
        if (TempElevMax > P_T_SNOWFZ) {
            if (Smax > 0.) {
                const double soil_saturation = soil / Smax;
                daily_runoff = dailyEffPrec * wg_pow(soil_saturation, VIN_D(TI_gamma, a.gamma_hbv[q]));
                if (dailySoilPET > (maxDailyPET - dailyCanopyEvapo) * soil_saturation)
                    dailyAET = (maxDailyPET - dailyCanopyEvapo) * soil_saturation;
                else
                    dailyAET = dailySoilPET;
                soil += dailyEffPrec - dailyAET - daily_runoff;
                if (fabs(soil) <= MIN_STOR_VOL) soil = 0.;
                dailyEffPrec = 0.;
                if (soil < 0.) {
                    dailyAET += soil;
                    soil = 0.;
                }
                daily_runoff *= cfa;
                immediate_runoff *= cfa;
                const short Rgmax = (short)VIN_K(K_rgmax, a.rgmax[q]);
                const float gwFactor = VIN_F(TI_gwfactor, a.gwfactor[q]);
                if ((Rgmax / C100) < (gwFactor * daily_runoff)) daily_gw_recharge = Rgmax / C100;
                else daily_gw_recharge = gwFactor * daily_runoff;
                pot_gw_recharge = 0.;
                if (((VIN_K(K_arid, a.arid[r]) == 1) && (VIN_K(K_texture, a.texture[r]) < 21)) && (VIN_K(K_ldd, a.ldd[r]) >= 0)) {  // :1165-1176
                    if (dailyPrec <= VIN_D(TI_pcrit, a.p_pcrit[q])) {
                        pot_gw_recharge = daily_gw_recharge;
                        daily_gw_recharge = 0.;
                    }
                }
                daily_runoff -= pot_gw_recharge;
                pot_gw_recharge /= cfa;
                soil += pot_gw_recharge;
                if (soil > Smax) {
                    soil_water_overflow += soil - Smax;
                    soil = Smax;
                }
                soil_water_overflow *= cfa;
                total_daily_runoff = daily_runoff + immediate_runoff + soil_water_overflow;
            } else {
                total_daily_runoff = 0.;
                daily_gw_recharge = 0.;
            }
        } else {
            soil_water_overflow *= cfa;
            dailyEffPrec *= cfa;
            total_daily_runoff += soil_water_overflow + dailyEffPrec;
            daily_gw_recharge = 0.;
            dailyAET = 0.;
        }
        a.gw_recharge[i] = daily_gw_recharge;  // :1221 (not updated by the fix-up below)
        flux[2] = daily_gw_recharge;
        landStorageChangeSum += soil - initialStorage;
        land_aet = landStorageChangeSum * (cfa - 1.0) - dailyPrec * (cfa - 1.0)
                   + (dailyAET + dailyCanopyEvapo + dailySnowEvapo) * cfa;
        if (land_aet < 0.) {
            neg_land_aet = land_aet;
            land_aet = 0.;
        }
        land_aet_uncorr = (dailyAET + dailyCanopyEvapo + dailySnowEvapo);
    }
    a.land_aet[i] = land_aet;
    a.land_aet_uncorr[i] = land_aet_uncorr;

    // surface runoff (:1244-1257)
    if (neg_land_aet < 0.) total_daily_runoff = total_daily_runoff + neg_land_aet;
    if (total_daily_runoff < 0.) total_daily_runoff = 0.;
    if ((total_daily_runoff - daily_gw_recharge) < 0.) {
        const double neg_runoff = total_daily_runoff - daily_gw_recharge;
        daily_gw_recharge = total_daily_runoff;
        soil += neg_runoff;
    }
    a.soil[i] = soil;
    a.surface_runoff[i] = total_daily_runoff - daily_gw_recharge;
    flux[0] = noland ? storage_transfer : VIN_D(LI_transfer_old, a.storage_transfer[i]);  // G_dailyStorageTransfer keeps its last value (:1113)
    flux[1] = total_daily_runoff - daily_gw_recharge;
}

// number of days of river discharge kept in flight (temporal wavefront over the level graph)
#ifndef WGK_QBUF_K
#define WGK_QBUF_K 32
#endif
constexpr int QBUF_K = WGK_QBUF_K;
// run-time stamps of the level-0 tasks (bench.py: duration of the dominant kernel inside the timed graph)
constexpr int STAMP_DAYS = 512;
__device__ __forceinline__ void stamp_task(const WgkParams &p, const int kind, const int which, const int dayofs) {
#ifndef WGK_EMU
    if (p.stamps && (threadIdx.x & 31) == 0 && dayofs < STAMP_DAYS) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        unsigned long long *e = p.stamps + ((size_t)(kind * 2 + which)) * STAMP_DAYS + dayofs;
        if (which) atomicMax(e, t);
        else atomicMin(e, t);
    }
#endif
}
#ifdef WGK_PHASE_TIMING  // development aid: in-situ warp durations of the thread-per-cell task kernels (tools/insitu_timing.py)
__device__ unsigned long long g_insitu[8];  // {V cycles, V warps, R cycles, R warps, V level-0 cycles, V level-0 warps, R level-0 cycles, R level-0 warps}
__device__ unsigned long long g_stamp[2][2][512];  // level-0 tasks: [V, R][first warp start, last warp end][day offset], globaltimer ns
__device__ __forceinline__ unsigned long long insitu_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define WGK_INSITU_STAMP(kind_, which_, day_) do { if ((threadIdx.x & 31) == 0 && (day_) < 512) { \
    if (which_) atomicMax(&g_stamp[kind_][1][day_], insitu_ns()); else atomicMin(&g_stamp[kind_][0][day_], insitu_ns()); } } while (0)
__device__ unsigned int g_warpdur[4][1024];  // cycles of every level-0 vertical warp on day offsets 100, 101, 200, 300 (persistence of slow warps)
#define WGK_INSITU_WARPDUR(day_) do { const int slot_ = (day_) == 100 ? 0 : (day_) == 101 ? 1 : (day_) == 200 ? 2 : (day_) == 300 ? 3 : -1; \
    const int w_ = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; \
    if (slot_ >= 0 && (threadIdx.x & 31) == 0 && w_ < 1024 && blockIdx.y == 0) g_warpdur[slot_][w_] = (unsigned int)(clock64() - insitu_t0_); } while (0)
#define WGK_INSITU_WARPDUR2(day_, which_) do { const int slot_ = (day_) == 100 ? (which_) : (day_) == 200 ? 2 + (which_) : -1; \
    const int w_ = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; \
    if (slot_ >= 0 && (threadIdx.x & 31) == 0 && w_ < 1024 && blockIdx.y == 0) g_warpdur[slot_][w_] = (unsigned int)(clock64() - insitu_t0_); } while (0)
#define WGK_INSITU_BEGIN() const long long insitu_t0_ = clock64()
#define WGK_INSITU_END(k_, l0_) do { if ((threadIdx.x & 31) == 0) { const unsigned long long dt_ = (unsigned long long)(clock64() - insitu_t0_); \
    atomicAdd(&g_insitu[k_], dt_); atomicAdd(&g_insitu[(k_) + 1], 1ull); if (l0_) { atomicAdd(&g_insitu[(k_) + 4], dt_); atomicAdd(&g_insitu[(k_) + 5], 1ull); } } } while (0)
#else
#define WGK_INSITU_BEGIN() do { } while (0)
#define WGK_INSITU_END(k_, l0_) do { } while (0)
#define WGK_INSITU_STAMP(kind_, which_, day_) do { } while (0)
#define WGK_INSITU_WARPDUR(day_) do { } while (0)
#define WGK_INSITU_WARPDUR2(day_, which_) do { } while (0)
#endif

// ----------------------------------------------------------------------------------------
// routing helpers
// ----------------------------------------------------------------------------------------
// ----------------------------------------------------------------------------------------
// water use: subtract_use 2 with use_alloc 0, delayedUseSatisfaction 0, aggrNUsGloLakResOpt 0
// ----------------------------------------------------------------------------------------
// The day's net abstraction from groundwater in front of a groundwater balance (routing.cpp:1929-1935 and its four copies):
// calcNextDay_M resets it to the month's value, updateNetAbstractionGW (:5503-5572) adapts it to the surface-water use that
// stayed unsatisfied the day before (return flows reduced) or was satisfied late (reintroduced).  Returns what is subtracted
// from the net groundwater inflow.
__device__ __noinline__ double wu_net_gw_use(const WgkParams &p, const int r, const size_t i, const size_t q) {
    const WgkArrays &a = p.a;
    const double nug = a.wu_nug_month[q];
    double out = nug;
    double rem = a.wu_daily_remaining[i];
    if (rem != 0.) {
        if (rem > 1.e-12) {
            const double wusi = a.wu_wusi[i];
            if (wusi > 0.) {
                const double eff = a.wu_cusi[i] / wusi;
                const double frgi = a.wu_frgi[r];
                const double factor = 1 - (1 - frgi) * (1 - eff);
                double WUsiNew = 1 / factor * (wusi * factor - rem);
                if (WUsiNew < 0.) {
                    WUsiNew = 0.;
                    a.wu_uns_irr[i] += wusi * factor;
                    a.wu_uns_oth[i] += rem - (wusi * factor);
                } else {
                    a.wu_uns_irr[i] += rem;
                }
                const double returnflowChange = (frgi * (1 - eff) * (WUsiNew - wusi));
                a.wu_red_rf[i] += returnflowChange;
                out = nug - returnflowChange;
                a.wu_daily_remaining[i] = 0.;
            } else {
                a.wu_uns_oth[i] += rem;
            }
        } else if (rem < -1.e-12) {
            const double uns_oth = a.wu_uns_oth[i], uns_irr = a.wu_uns_irr[i];
            double fromIrrig = rem + uns_oth;
            if (fromIrrig < 0.) {
                a.wu_uns_oth[i] = 0.;
                if (uns_irr == 0.) {
                    a.wu_daily_remaining[i] = 0.;
                } else {
                    double ratio = (fromIrrig / uns_irr);
                    if (ratio < -1.) {
                        ratio = -1.;
                        fromIrrig = uns_irr * -1.;
                    }
                    const double red_rf = a.wu_red_rf[i];
                    const double returnflowChange = (ratio * red_rf);
                    a.wu_uns_irr[i] = uns_irr + fromIrrig;
                    a.wu_red_rf[i] = red_rf + returnflowChange;
                    out = nug - returnflowChange;
                    a.wu_daily_remaining[i] = 0.;
                }
            } else {
                a.wu_uns_oth[i] = uns_oth + rem;
                a.wu_daily_remaining[i] = 0.;
            }
        }
    }
    a.wu_daily_nug[i] = out;
    return out;
}

// groundwater linear reservoir (routing.cpp:1938-1958 and four identical copies)
__device__ __forceinline__ double gw_step(double &Sg, double netGWin, double ek, double invk) {
    // ek = exp(-1. * k) and invk = (1. / k) depend on the parameter P_GWOUTF_C only: derived once (k_derive_static)
    const double prev = Sg;
    Sg = prev * ek + invk * netGWin * (1. - ek);
    if (fabs(Sg) <= MIN_STOR_VOL) Sg = 0.;
    double qq = prev - Sg + netGWin;
    if (qq <= 0.) {
        qq = 0.;
        Sg = prev + netGWin;
        if (fabs(Sg) <= MIN_STOR_VOL) Sg = 0.;
    }
    return qq;
}

__device__ __forceinline__ double clamp01(double x) {
    if (x < 0.) x = 0.;
    if (x > 1.) x = 1.;
    return x;
}

// Inflow-independent terms of the global lake / reservoir / global wetland / arid groundwater
// equations of one cell and one day, computed by the cell-parallel pre-pass so that the ordered
// sweep only evaluates what really depends on the upstream inflow.
enum { GB_L_PREC, GB_L_PET, GB_L_MAX, GB_L_GWR, GB_R_PREC, GB_R_PET, GB_R_GWR, GB_R_C, GB_R_PROV, GB_R_CAP,
       GB_W_PREC, GB_W_PET, GB_W_MAX, GB_W_GWR, GB_EKS, GB_INVKS, GB_EKG, GB_INVKG, GB_GWRECH, GB_LOC_GWR_LAK,
       GB_LOC_GWR_WET, GB_N = 24 };

// cell class bits (s_flags, derived once from the statics by k_derive_static)
constexpr int FL_ACTIVE = 1, FL_LDD_OUT = 2, FL_ARIDC = 4, FL_LAKE = 8, FL_RES = 16, FL_GLOWET = 32, FL_TBC1 = 64;  // FL_TBC1: G_toBeCalculated == 1

#if !WGK_WU  // no water use inside: compiled once per layout
// one-time derivation of inflow-independent river constants (routing.cpp:7293-7296):
//   s_c1 = 1 / (M_RIVRGH_C * roughness),  s_slope_pow = pow(slope, 0.5),  s_flags
__global__ void __launch_bounds__(128) k_derive_static(const __grid_constant__ WgkParams p) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int ps = blockIdx.y;
    if (r >= p.ncell) return;
    const WgkArrays &a = p.a;
    const size_t q = qi_of(p, ps, r);
    a.s_c1[q] = 1. / (a.p_rivrgh[q] * a.roughness[r]);
    a.s_ekg[q] = exp(-1. * a.p_gwoutf[q]);  // routing.cpp:1940
    a.s_invkg[q] = (1. / a.p_gwoutf[q]);
    a.s_eks[q] = exp(-1. * a.p_swoutf[q]);  // exp(-k) and 1/k of the global lakes / wetlands (routing.cpp:2741, 3237)
    a.s_invks[q] = (1. / a.p_swoutf[q]);
    a.s_slope_pow[q] = pow(a.river_slope[r], 0.5);
    if (ps == 0) {
        // band elevation minus cell mean elevation (the integer the lapse rate multiplies, daily.cpp:935) and its extremes
        const int e0 = a.elevation[r];
        int lo = 32767, hi = -32768;
        for (int b = 1; b < WGK_NBAND_K; b++) {
            const int de = a.elevation[(size_t)b * p.stride + r] - e0;
            a.s_delev[(size_t)b * p.stride + r] = (int16_t)de;
            lo = de < lo ? de : lo;
            hi = de > hi ? de : hi;
        }
        a.s_delev[r] = 0;
        for (int b = 0; b < WGK_NBAND_K; b++) a.s_elev32[(size_t)b * p.stride + r] = a.elevation[(size_t)b * p.stride + r];
        a.s_de_min[r] = (int16_t)lo;
        a.s_de_max[r] = (int16_t)hi;
        int f = 0;
        const int ldd = a.ldd[r];
        if (a.contcell[r] && (0 != a.toBeCalculated[r])) f |= FL_ACTIVE;
        if (ldd >= 0) f |= FL_LDD_OUT;
        if ((1 == a.arid[r]) && (ldd >= 0)) f |= FL_ARIDC;
#ifndef WGK_EXP_NOGLOB  // (timing experiment only: no global water bodies anywhere)
        if (a.lake_area[r] > 0.) f |= FL_LAKE;
        if (a.reservoir_area[r] > 0.) f |= FL_RES;
        if (a.glo_wetland[r] > 0) f |= FL_GLOWET;
#endif
        if (a.contcell[r] && (1 == a.toBeCalculated[r])) f |= FL_TBC1;
        a.s_flags[r] = (int8_t)f;
    }
}

// s_snowfree: all 100 elevation bands of the cell hold exactly zero snow (recomputed after every
// upload of the band state; afterwards maintained by the vertical kernel)
__global__ void __launch_bounds__(128) k_derive_member(const __grid_constant__ WgkParams p) {
    int r, m;
    if (!map_thread(p, 0, p.ncell, r, m)) return;
    const double *S = p.a.snow_bands + bi(p, m, r, WGK_NBAND_K);
    const size_t bs = band_stride(p);
    int nz = 0;
    for (int b = 1; b < WGK_NBAND_K; b++) nz |= (S[(size_t)b * bs] != 0.);
    p.a.s_snowfree[mi(p, m, r)] = (int8_t)(nz == 0);
}

#endif  // !WGK_WU
// ----------------------------------------------------------------------------------------
// cell-parallel pre-pass of the routing day: everything that does not depend on upstream cells
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void local_compute(const WgkParams &p, const int r, const int m, const LocalIn &li, const LocalFlux &fx,
                                              const int wu_month) {  // wu_month: month 0..11 of the day (only read with water use)
    const WgkArrays &a = p.a;
    const size_t i = mi(p, m, r);
    a.river_evapo[i] = 0.;  // routing.cpp:1781
    if (!li.contcell) return;
    const int flags = li.flags;
    const double M_EVAREDEX = li.evaredex;
    const double cellArea = li.area;
    const double cfa = li.cfa;
    const double owPrec = fx.owPrec, owPET = fx.owPET;
    const int ldd = li.ldd;
    const int arid = li.arid;
    const bool aridc = (flags & FL_ARIDC) != 0;
    const double laf = li.laf;
    const double contf = li.contf;
    double dailyLocalSurfaceRunoff;
    double localRunoff = 0., localRunoffIntoRiver = 0., localGWRunoffIntoRiver = 0., fswb_catchment = 0.;
    double gwr_loclak = 0., gwr_locwet = 0.;

    if (laf <= 0.)  // :1885-1891
        dailyLocalSurfaceRunoff = fx.storage_transfer * cellArea / C1E6 * li.laf_prev / C100;
    else
        dailyLocalSurfaceRunoff = fx.surface_runoff * cellArea / C1E6 * laf / C100;
    if (ldd >= 0) {  // :1898-1908
        fswb_catchment = li.fswb_init * 20.;
        if (fswb_catchment > 1.) fswb_catchment = 1.;
        localRunoffIntoRiver = (1. - fswb_catchment) * dailyLocalSurfaceRunoff;
    }
    if ((1 == arid) && (ldd >= 0)) {  // :1910-1915
        localRunoff = fswb_catchment * dailyLocalSurfaceRunoff;
    }
    if ((0 == arid) && (ldd >= 0)) {  // :1979-2033
        double netGWin = fx.gw_recharge * cellArea * (laf / C100) / C1E6;
        if (WU) netGWin -= wu_net_gw_use(p, r, i, qi(p, m, r));
        double Sg = li.gw;
        const double qg = gw_step(Sg, netGWin, li.ekg, li.invkg);
        a.gw[i] = Sg;
        localGWRunoffIntoRiver = (1. - fswb_catchment) * qg;
        const double localGWRunoff = fswb_catchment * qg;
        localRunoff = (fswb_catchment * dailyLocalSurfaceRunoff) + localGWRunoff;
    }
    if (ldd < 0) {  // :2123-2176
        double netGWin = fx.gw_recharge * cellArea * (laf / C100) / C1E6;
        if (WU) netGWin -= wu_net_gw_use(p, r, i, qi(p, m, r));
        double Sg = li.gw;
        const double qg = gw_step(Sg, netGWin, li.ekg, li.invkg);
        a.gw[i] = Sg;
        if (laf == 0.)
            dailyLocalSurfaceRunoff = fx.storage_transfer * cellArea / C1E6 * li.laf_prev / C100;
        else
            dailyLocalSurfaceRunoff = fx.surface_runoff * cellArea / C1E6 * laf / C100;
        localRunoff = dailyLocalSurfaceRunoff + qg;
    }

    double inflow = localRunoff;
    if (flags & FL_ACTIVE) {
        const double kS = li.kS;
        const double loc_lake = li.loc_lake;
        if (loc_lake > 0.) {  // local lake, :2318-2490
            const double prev = li.loc_lake_stor;
            const double maxStorage = ((loc_lake) / C100) * cellArea * li.lake_depth;
            const double rf = li.red_loc_lake;
            double evapo = ((1.0 - cfa) * owPrec * rf) + (cfa * (owPET * rf));
            if (evapo < 0.) evapo = 0.;
            const double totalInflow = inflow + (owPrec * rf) * (cellArea / C1E6) * (loc_lake / C100);
            if (aridc) gwr_loclak = 10. * rf * loc_lake / C100 / (contf / C100);
            const double PETgwr = evapo * (cellArea / C1E6) * (loc_lake / C100) + gwr_loclak * cellArea * (contf / C100) / C1E6;
            double PETgwrMax = prev + maxStorage + totalInflow;
            if (PETgwrMax < 0.) PETgwrMax = 0.;
            double S;
            if (PETgwr > PETgwrMax) {
                S = (-1.) * maxStorage;
                gwr_loclak *= PETgwrMax / PETgwr;
            } else {
                S = prev + totalInflow - PETgwr;
            }
            double outflow;
            if (prev > 0.) {
                {
                    const double x = (prev / maxStorage);  // pow(x, 1.5) (:2440) as x * sqrt(x), <= 1 ulp apart
                    #ifdef WGK_LIBM_POW
                    outflow = kS * prev * wg_pow(x, 1.5);
#else
                    outflow = kS * prev * (x * sqrt(x));
#endif
                }
                if (S <= 0.) outflow = 0;
                else if (outflow > S) outflow = S;
            } else
                outflow = 0.;
            S -= outflow;
            if (fabs(S) <= MIN_STOR_VOL) S = 0.;
            if (S > maxStorage) {
                outflow += (S - maxStorage);
                S = maxStorage;
            }
            inflow = outflow;
            a.loc_lake_stor[i] = S;
            a.red_loc_lake[i] = clamp01(1. - wg_pow(fabs(S - maxStorage) / (2. * maxStorage), (M_EVAREDEX * 3.32193)));
        }
        const double loc_wetland = li.loc_wetland;
        if (loc_wetland > 0.) {  // local wetland, :2495-2617
            const double prev = li.loc_wetl_stor;
            const double maxStorage = ((loc_wetland) / C100) * cellArea * li.wetl_depth;
            const double rf = li.red_loc_wetl;
            double evapo = ((1.0 - cfa) * owPrec * rf) + (cfa * (owPET * rf));
            if (evapo < 0.) evapo = 0.;
            const double totalInflow = inflow + (owPrec * rf * (cellArea / C1E6) * (loc_wetland / C100));
            if (aridc) gwr_locwet = 10. * rf * loc_wetland / C100 / (contf / C100);
            const double PETgwr = evapo * (cellArea / C1E6) * (loc_wetland / C100) + gwr_locwet * cellArea * (contf / C100) / C1E6;
            const double PETgwrMax = prev + totalInflow;
            double S;
            if (PETgwr > PETgwrMax) {
                S = 0.;
                gwr_locwet *= PETgwrMax / PETgwr;
            } else {
                S = prev + totalInflow - PETgwr;
            }
            if (fabs(S) <= MIN_STOR_VOL) S = 0.;
            double outflow;
            if (S > 0.) {
                {
                    const double x = (S / maxStorage);  // pow(x, 2.5) (:2580) as x * x * sqrt(x)
                    #ifdef WGK_LIBM_POW
                    outflow = kS * S * wg_pow(x, 2.5);
#else
                    outflow = kS * S * ((x * x) * sqrt(x));
#endif
                }
                if (outflow > S) outflow = S;
            } else
                outflow = 0.;
            S -= outflow;
            if (S > maxStorage) {
                outflow += (S - maxStorage);
                S = maxStorage;
            }
            inflow = outflow;
            a.loc_wetl_stor[i] = S;
            a.red_loc_wetl[i] = clamp01(1. - wg_pow(fabs(S - maxStorage) / (maxStorage), (M_EVAREDEX * 3.32193)));
        }
        // arid cells without global lake / reservoir / global wetland: the groundwater step below
        // the surface water bodies (:3305-3386) does not depend on upstream inflow either
        if (aridc && !(flags & (FL_LAKE | FL_RES | FL_GLOWET))) {
            const double gwr_swb = gwr_loclak + 0. + gwr_locwet + 0. + 0.;
            a.gwr_swb[i] = gwr_swb;
            double netGWin = gwr_swb * cellArea * (contf / C100) / C1E6 + fx.gw_recharge * cellArea * (laf / C100) / C1E6;
            if (WU) netGWin -= wu_net_gw_use(p, r, i, qi(p, m, r));
            double Sg = li.gw;
            localGWRunoffIntoRiver = gw_step(Sg, netGWin, li.ekg, li.invkg);
            a.gw[i] = Sg;
        }
        if (flags & (FL_LAKE | FL_RES | FL_GLOWET)) {
            // Cells with a global water body are 8 % of the grid and the slowest warps of every level (they pace the single
            // member: without them the simulated year takes 15.6 instead of 18.4 ms).  Every input of the three blocks below is
            // therefore loaded HERE, in one round and before the first store (the compiler cannot move a load above the stores
            // to g[]), instead of one dependent round of loads per block; exp(-k) and 1/k come from the one-time derivation.
            const size_t q_ = qi(p, m, r);
            const bool hasL = (flags & FL_LAKE) != 0, hasR = (flags & FL_RES) != 0, hasW = (flags & FL_GLOWET) != 0;
            const double in_eks = a.s_eks[q_], in_invks = a.s_invks[q_];
            const double in_lake_area = hasL ? a.lake_area[r] : 0., in_red_glo_lake = hasL ? a.red_glo_lake[i] : 0.;
            const double in_reservoir_area = hasR ? a.reservoir_area[r] : 0., in_stor_cap = hasR ? a.stor_cap[r] : 0.;
            const double in_mean_outflow = hasR ? a.mean_outflow[r] : 0., in_red_res = hasR ? a.red_res[i] : 0.;
            const double in_mean_demand = hasR ? a.mean_demand[r] : 0.;
            const int in_res_type = hasR ? a.res_type[r] : 0;
            const double in_glo_wetland = hasW ? a.glo_wetland[r] : 0., in_red_glo_wetl = hasW ? a.red_glo_wetl[i] : 0.;
            double *g = p.gbody + gb(p, m, p.gidx[r], GB_N);
            const size_t gs = gbody_stride(p);
            g[GB_EKS * gs] = in_eks;
            g[GB_INVKS * gs] = in_invks;
            g[GB_EKG * gs] = li.ekg;
            g[GB_INVKG * gs] = li.invkg;
            g[GB_GWRECH * gs] = fx.gw_recharge * cellArea * (laf / C100) / C1E6;
            g[GB_LOC_GWR_LAK * gs] = gwr_loclak;
            g[GB_LOC_GWR_WET * gs] = gwr_locwet;
            if (flags & FL_LAKE) {  // :2630-2676
                const double lake_area = in_lake_area;
                const double rf = in_red_glo_lake;
                double evapo = ((1.0 - cfa) * owPrec) + (cfa * owPET * rf);
                if (evapo < 0.) evapo = 0.;
                const double gwr = aridc ? 10. * rf * (lake_area / (cellArea * (contf / C100))) : 0.;
                g[GB_L_PREC * gs] = (owPrec * (lake_area / C1E6));
                g[GB_L_GWR * gs] = gwr;
                g[GB_L_PET * gs] = evapo * (lake_area / C1E6) + gwr * cellArea * (contf / C100) / C1E6 + 0.;
                g[GB_L_MAX * gs] = (lake_area)*li.lake_depth;
            }
            if (flags & FL_RES) {  // :2807-2870, 2960-2983
                const double reservoir_area = in_reservoir_area;
                const double stor_cap = in_stor_cap;
                const double mean_outflow = in_mean_outflow;
                const double rf = in_red_res;
                double evapo = ((1.0 - cfa) * owPrec) + (cfa * (owPET * rf));
                if (evapo < 0.) evapo = 0.;
                const double gwr = aridc ? 10. * rf * (reservoir_area / (cellArea * (contf / C100))) : 0.;
                g[GB_R_PREC * gs] = (owPrec * (reservoir_area / C1E6));
                g[GB_R_GWR * gs] = gwr;
                g[GB_R_PET * gs] = evapo * (reservoir_area / C1E6) + gwr * cellArea * (contf / C100) / C1E6;
                g[GB_R_C * gs] = stor_cap / (mean_outflow * 31536000. / 1000000000.);
                g[GB_R_CAP * gs] = stor_cap;
                double prov_rel = 0.;
                const int res_type = in_res_type;
                if (res_type == 1) {  // irrigation reservoir (:2960-2977)
                    double monthlyUse = 0.;
                    if (WU) {
                        // own use plus the share of up to 5 downstream cells that have no reservoir (the chain is resolved on the
                        // host, wgk_api.cu ensure_derived, with the reference's indexing of G_reservoir_area)
                        double dailyUse = a.wu_nus_month[qi(p, m, r)];
                        for (int k = 0; k < 5; k++) {
                            const int x = p.wu_res_idx[(size_t)k * p.stride + r];
                            if (x < 0) break;
                            dailyUse += a.wu_nus_month[qi(p, m, x)] * a.wu_alloc_coeff[(size_t)k * p.stride + r];
                        }
                        const int ndim[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
                        monthlyUse = dailyUse * ndim[wu_month];
                        monthlyUse = monthlyUse * 1000000000. / (ndim[wu_month] * 86400.);
                    }
                    const double mean_demand = in_mean_demand;
                    if (mean_demand >= 0.5 * mean_outflow) prov_rel = mean_outflow / 2. * (1. + monthlyUse / mean_demand);
                    else prov_rel = mean_outflow + monthlyUse - mean_demand;
                } else if (res_type == 2) {
                    prov_rel = mean_outflow;
                }
                g[GB_R_PROV * gs] = prov_rel;
            }
            if (flags & FL_GLOWET) {  // :3178-3200
                const double glo_wetland = in_glo_wetland;
                const double rf = in_red_glo_wetl;
                double evapo = ((1.0 - cfa) * (owPrec * rf)) + (cfa * (owPET * rf));
                if (evapo < 0.) evapo = 0.;
                const double gwr = aridc ? 10. * rf * glo_wetland / C100 / (contf / C100) : 0.;
                g[GB_W_PREC * gs] = (owPrec * rf * (cellArea / C1E6) * (glo_wetland / C100));
                g[GB_W_GWR * gs] = gwr;
                g[GB_W_PET * gs] = evapo * (cellArea / C1E6) * ((glo_wetland) / C100) + gwr * cellArea * (contf / C100) / C1E6;
                g[GB_W_MAX * gs] = ((glo_wetland) / C100) * cellArea * li.wetl_depth;
            }
        }
        // river evaporation and precipitation on yesterday's river area fraction (:3425-3441)
        const double raf = li.raf_next;
        a.t_river_evapo[i] = ((1.0 - cfa) * (owPrec) + (cfa * owPET)) * raf / C100 * cellArea / C1E6;
        a.t_river_precip[i] = owPrec * raf / C100 * cellArea / C1E6;
    }
    a.t_inflow_local[i] = inflow;
    a.t_runoff_to_river[i] = localRunoffIntoRiver;
    a.t_gw_to_river[i] = localGWRunoffIntoRiver;
    a.t_gwr_loclak[i] = gwr_loclak;
    a.t_gwr_locwet[i] = gwr_locwet;
}


__device__ __forceinline__ void route_local_cell(const WgkParams &p, const int r, const int m, const int wu_month) {
    const LocalIn li = local_load(p, r, m);
    local_compute(p, r, m, li, local_flux_load(p, r, m), wu_month);
}

__global__ void __launch_bounds__(128) k_route_local(const __grid_constant__ WgkParams p, const int dayofs) {
    int r, m;
    if (!map_thread(p, 0, p.ncell, r, m)) return;
    route_local_cell(p, r, m, p.cal_days[4 * dayofs + 1]);
}

// ----------------------------------------------------------------------------------------
// the ordered sweep: only what depends on upstream cells stays on the level-to-level chain
// ----------------------------------------------------------------------------------------
// inflow-independent inputs of one cell, loaded before the hand-off of the previous level
struct RiverCtx {
    double inflow_local, runoff_to_river, gw_to_river, c1, slope_pow, bw, river_length, prevR, precip, evapo;
    int up0, up1, flags;
};

__device__ __forceinline__ RiverCtx load_ctx(const WgkParams &p, const int r, const size_t i, const size_t q) {
    const WgkArrays &a = p.a;
    RiverCtx c;
    c.flags = a.s_flags[r];
    c.up0 = p.up_off[r];
    c.up1 = p.up_off[r + 1];
    c.inflow_local = a.t_inflow_local[i];
    c.runoff_to_river = a.t_runoff_to_river[i];
    c.gw_to_river = a.t_gw_to_river[i];
    c.c1 = a.s_c1[q];
    c.slope_pow = a.s_slope_pow[q];
    c.bw = a.river_bottom_width[r];
    c.river_length = a.river_length[r];
    c.prevR = a.river_stor[i];
    c.precip = a.t_river_precip[i];
    c.evapo = a.t_river_evapo[i];
    return c;
}

// global lake, reservoir, global wetland and the arid groundwater below them (routing.cpp:2630-3386)
// for the few cells that have them: only the inflow-dependent arithmetic; exp(), 1/k, evaporation
// and recharge demands come from the pre-pass (GBody), the reduction-factor pow() is done by the
// post-pass.  Returns the inflow handed to the river.
// BATCH (the fused task k_level_day, where these cells are the slowest warps of every level): every input of the blocks in ONE
// round of loads before the first store, and the routine inlined - a block's own loads after the previous block's store cost one
// dependent round trip each, the call a stack frame (17.4 -> 17.1 ms per simulated year).  The throughput kernels keep the
// out-of-line form with the loads inside the blocks: batched, k_route_level needs 103 instead of 80 registers.
template <bool BATCH>
__device__ __forceinline__ double route_global_bodies_impl(const WgkParams &p, const int r, const int m, const size_t i,
                                                   double inflow, const int flags, const int day, const int month,
                                                   double &gwToRiver WGK_WU_PARAMS) {
#if !WGK_WU
    double remainingUse = 0., dailyActualUse = 0.;  // (no water use: the terms below fold away)
#endif
    // remainingUse / dailyActualUse: water use (0 / untouched without it).  The day's surface-water use is taken from the global
    // lake (routing.cpp:2678-2788) and the reservoir (:2866-2946, 3068); what they cannot supply goes on to the river (:3166-3172).
    constexpr bool wu = WU;
    double remainingUseGloLake = 0., remainingUseRes = 0.;
    const WgkArrays &a = p.a;
    const double *g = p.gbody + gb(p, m, p.gidx[r], GB_N);  // (no __restrict__: written earlier by the same thread in k_days_persistent)
    const size_t gs = gbody_stride(p);
    const bool hasL = BATCH && (flags & FL_LAKE) != 0, hasR = BATCH && (flags & FL_RES) != 0, hasW = BATCH && (flags & FL_GLOWET) != 0,
               hasA = BATCH && (flags & FL_ARIDC) != 0;
    const double ek = g[GB_EKS * gs], invk = g[GB_INVKS * gs];
    const double l_prev = hasL ? a.glo_lake_stor[i] : 0., l_max = hasL ? g[GB_L_MAX * gs] : 0., l_pet = hasL ? g[GB_L_PET * gs] : 0.;
    const double l_gwr = hasL ? g[GB_L_GWR * gs] : 0., l_prec = hasL ? g[GB_L_PREC * gs] : 0.;
    const double r_cap = hasR ? g[GB_R_CAP * gs] : 0., r_prev = hasR ? a.res_stor[i] : 0., r_pet = hasR ? g[GB_R_PET * gs] : 0.;
    const double r_gwr = hasR ? g[GB_R_GWR * gs] : 0., r_prec = hasR ? g[GB_R_PREC * gs] : 0., r_krel = hasR ? a.k_release[i] : 0.;
    const double r_c = hasR ? g[GB_R_C * gs] : 0., r_prov = hasR ? g[GB_R_PROV * gs] : 0.;
    const int r_start_month = hasR ? a.start_month[r] : 0;
    const double w_prev = hasW ? a.glo_wetl_stor[i] : 0., w_max = hasW ? g[GB_W_MAX * gs] : 0., w_pet = hasW ? g[GB_W_PET * gs] : 0.;
    const double w_gwr = hasW ? g[GB_W_GWR * gs] : 0., w_prec = hasW ? g[GB_W_PREC * gs] : 0.;
    const double a_gwr_lak = hasA ? g[GB_LOC_GWR_LAK * gs] : 0., a_gwr_wet = hasA ? g[GB_LOC_GWR_WET * gs] : 0., a_gwrech = hasA ? g[GB_GWRECH * gs] : 0.;
    const double a_ekg = hasA ? g[GB_EKG * gs] : 0., a_invkg = hasA ? g[GB_INVKG * gs] : 0.;
    const double a_area = hasA ? a.area[r] : 0., a_contf = hasA ? a.contfreq[r] : 0., a_gw = hasA ? a.gw[i] : 0.;
    double gwr_glolak = 0., gwr_res = 0., gwr_glowet = 0.;
    if (flags & FL_LAKE) {  // :2677-2720
        const double prev = BATCH ? l_prev : a.glo_lake_stor[i];
        remainingUseGloLake = (flags & FL_RES) ? 0.5 * remainingUse : remainingUse;  // :2681-2686 (0 without water use)
        const double remainingUseGloLakeStart = remainingUseGloLake;
        const double maxStorage = BATCH ? l_max : g[GB_L_MAX * gs], PET = (BATCH ? l_pet : g[GB_L_PET * gs]) + remainingUseGloLake;
        gwr_glolak = BATCH ? l_gwr : g[GB_L_GWR * gs];
        const double totalInflow = inflow + (BATCH ? l_prec : g[GB_L_PREC * gs]);
        const double PETmax = totalInflow + maxStorage + prev;
        double S, outflow;
        if (PET > PETmax) {
            S = (-1.) * maxStorage;
            outflow = 0.;
            gwr_glolak *= PETmax / PET;
            if (remainingUseGloLake > 0.) remainingUseGloLake -= remainingUseGloLake * PETmax / PET;  // :2710-2716
            else remainingUseGloLake = 0.;
        } else {
            S = prev * ek + invk * (totalInflow - PET) * (1. - ek);
            outflow = totalInflow + prev - S - PET;
            if (S > maxStorage) {
                outflow += (S - maxStorage);
                S = maxStorage;
            }
            if (outflow < 0.) {
                outflow = 0.;
                S = prev + totalInflow - PET;
            }
            remainingUseGloLake = 0.;  // :2752
        }
        if (fabs(S) <= MIN_STOR_VOL) S = 0.;
        inflow = outflow;
        dailyActualUse = remainingUseGloLakeStart - remainingUseGloLake;  // :2788
        a.glo_lake_stor[i] = S;
    }
    if (flags & FL_RES) {  // :2871-3040
        const double stor_cap = BATCH ? r_cap : g[GB_R_CAP * gs];
        const double maxStorage = stor_cap;
        const double prev = BATCH ? r_prev : a.res_stor[i];
        const double PET = BATCH ? r_pet : g[GB_R_PET * gs];
        gwr_res = BATCH ? r_gwr : g[GB_R_GWR * gs];
        const double totalInflow = inflow + (BATCH ? r_prec : g[GB_R_PREC * gs]);
        const double PETmax = prev + totalInflow;
        remainingUseRes = (flags & FL_LAKE) ? 0.5 * remainingUse + remainingUseGloLake : remainingUse;  // :2869-2875
        const double remainingUseResStart = remainingUseRes;
        double S;
        if (PET > PETmax) {
            S = prev + totalInflow - PETmax;
            gwr_res *= PETmax / PET;
        } else {
            S = prev + totalInflow - PET;
        }
        if (wu) {  // :2922-2946: return flows are added; abstractions only above 10 % of the capacity
            if (remainingUseRes < 0.) {
                S -= remainingUseRes;
                remainingUseRes = 0.;
            } else if (S > (0.1 * stor_cap)) {
                if (remainingUseRes < (S - (stor_cap * 0.1))) {
                    S -= remainingUseRes;
                    remainingUseRes = 0.;
                } else {
                    remainingUseRes -= (S - (stor_cap * 0.1));
                    S = stor_cap * 0.1;
                }
            }
        }
        if (fabs(S) <= MIN_STOR_VOL) S = 0.;
        double Krel = BATCH ? r_krel : a.k_release[i];
        const int fdim[12] = {1, 32, 60, 91, 121, 152, 182, 213, 244, 274, 305, 335};
        if (month == (BATCH ? r_start_month : (int)a.start_month[r]) - 1 && day == fdim[month]) {  // :2945-2956
            if (S < (stor_cap * 0.1)) Krel = 0.1;
            else Krel = S / (maxStorage * 0.85);
            a.k_release[i] = Krel;
        }
        const double c_ratio = BATCH ? r_c : g[GB_R_C * gs], prov_rel = BATCH ? r_prov : g[GB_R_PROV * gs];
        double release;
        if (c_ratio >= 0.5) release = Krel * prov_rel;
        else
            release = ((4. * c_ratio * c_ratio) * Krel * prov_rel)
                      + ((1.0 - ((4. * c_ratio * c_ratio))) * inflow * 1000000000. / (24. * 3600.));
        double outflow;
        if (S >= (stor_cap * 0.1)) outflow = release * (24. * 3600.) / 1000000000.;
        else outflow = 0.1 * release * (24. * 3600.) / 1000000000.;
        if (outflow < 0.) outflow = 0.;
        S -= outflow;
        if (S > maxStorage) {
            outflow += (S - maxStorage);
            S = maxStorage;
        }
        if (S < 0.) {
            outflow += S;
            S = 0.;
        }
        inflow = outflow;
        dailyActualUse += remainingUseResStart - remainingUseRes;  // :3068
        a.res_stor[i] = S;
    }
    if (wu && (flags & (FL_LAKE | FL_RES))) remainingUse = (flags & FL_RES) ? remainingUseRes : remainingUseGloLake;  // :3166-3172
    if (flags & FL_GLOWET) {  // :3201-3260
        const double prev = BATCH ? w_prev : a.glo_wetl_stor[i];
        const double maxStorage = BATCH ? w_max : g[GB_W_MAX * gs], PET = BATCH ? w_pet : g[GB_W_PET * gs];
        gwr_glowet = BATCH ? w_gwr : g[GB_W_GWR * gs];
        const double totalInflow = inflow + (BATCH ? w_prec : g[GB_W_PREC * gs]);
        const double PETmax = totalInflow + prev;
        double S, outflow;
        if (PET > PETmax) {
            S = 0.;
            outflow = 0.;
            gwr_glowet *= PETmax / PET;
        } else {
            S = prev * ek + invk * (totalInflow - PET) * (1. - ek);
            outflow = totalInflow + prev - S - PET;
        }
        if (S > maxStorage) {
            outflow += (S - maxStorage);
            S = maxStorage;
        }
        if (fabs(S) <= MIN_STOR_VOL) S = 0.;
        inflow = outflow;
        a.glo_wetl_stor[i] = S;
    }
    if (flags & FL_ARIDC) {  // :3305-3386
        const double gwr_swb = (BATCH ? a_gwr_lak : g[GB_LOC_GWR_LAK * gs]) + gwr_glolak + (BATCH ? a_gwr_wet : g[GB_LOC_GWR_WET * gs]) + gwr_glowet + gwr_res;
        a.gwr_swb[i] = gwr_swb;
        double netGWin = gwr_swb * (BATCH ? a_area : a.area[r]) * ((BATCH ? a_contf : a.contfreq[r]) / C100) / C1E6 + (BATCH ? a_gwrech : g[GB_GWRECH * gs]);
        if (wu) netGWin -= wu_net_gw_use(p, r, i, qi(p, m, r));  // :3325-3330
        const double prev = BATCH ? a_gw : a.gw[i];
        const double ekg = BATCH ? a_ekg : g[GB_EKG * gs];
        double Sg = prev * ekg + (BATCH ? a_invkg : g[GB_INVKG * gs]) * netGWin * (1. - ekg);
        if (fabs(Sg) <= MIN_STOR_VOL) Sg = 0.;
        double qq = prev - Sg + netGWin;
        if (qq <= 0.) {
            qq = 0.;
            Sg = prev + netGWin;
            if (fabs(Sg) <= MIN_STOR_VOL) Sg = 0.;
        }
        gwToRiver = qq;
        a.gw[i] = Sg;
    }
    (void)m;
    return inflow;
}

// out-of-line form for the throughput kernels (one copy of the code, small register footprint at the call site)
__device__ __noinline__ double route_global_bodies(const WgkParams &p, const int r, const int m, const size_t i,
                                                   double inflow, const int flags, const int day, const int month,
                                                   double &gwToRiver WGK_WU_PARAMS) {
    return route_global_bodies_impl<false>(p, r, m, i, inflow, flags, day, month, gwToRiver WGK_WU_ARGS);
}

// river reach of one cell (routing.cpp:3388-3545), given the inflow-independent context and
// the sum of upstream discharges; writes discharge / storage and returns nothing
template <bool BATCH = false>
__device__ __forceinline__ double route_river(const WgkParams &p, const RiverCtx &c, const int r, const int m, const size_t i,
                                            const size_t q, const double inflowUpstream, const int day, const int month,
                                            double *__restrict__ qday, double *qout = nullptr, double *red_loc_lake_out = nullptr) {
    const WgkArrays &a = p.a;
    double inflow = c.inflow_local + inflowUpstream;  // :2623
    double gwToRiver = c.gw_to_river;
    // water use (:2193-2296 with use_alloc 0, delayedUseSatisfaction 0, aggrNUsGloLakResOpt 0): the day's desired use is the month's
    // net abstraction from surface water (negative = return flow)
    constexpr bool wu = WU;
    double remainingUse = 0., dailyActualUse = 0.;
    if (wu && (c.flags & FL_TBC1)) remainingUse = a.wu_nus_month[q];
    if (c.flags & (FL_LAKE | FL_RES | FL_GLOWET))
        inflow = BATCH ? route_global_bodies_impl<true>(p, r, m, i, inflow, c.flags, day, month, gwToRiver WGK_WU_ARGS)
                       : route_global_bodies(p, r, m, i, inflow, c.flags, day, month, gwToRiver WGK_WU_ARGS);
    double riverInflow = inflow;
    if (c.flags & FL_LDD_OUT) {
        riverInflow += c.runoff_to_river;
        riverInflow += gwToRiver;
    }
    // routingClass::getRiverVelocity (routing.cpp:7274-7307); pow(x, 2/3) is evaluated as
    // cbrt(x*x): same value to ~1 ulp with a much shorter dependent chain
    const double incoming_discharge = (riverInflow * 1000. * 1000. * 1000.) / C86400;  // (60. * 60. * 24.)
    const double riverDepth = 0.349 * wg_pow(incoming_discharge, 0.341);
    const double crossSectionalArea = riverDepth * (2.0 * riverDepth + c.bw);
    const double wettedPerimeter = c.bw + 2.0 * riverDepth * sqrt(5.0);
    const double hydraulicRad = crossSectionalArea / wettedPerimeter;
    double v = c.c1 * cbrt(hydraulicRad * hydraulicRad) * c.slope_pow;
    v = v * 86.4;
    if (v < 0.00001) v = 0.00001;  // an empty river gives exp(-inf) = 0 -> v = 0 -> floor, as pow(0, y) does
    const double K = v / c.river_length;
    const double prevR = c.prevR;
    double riverEvapo = c.evapo;
    riverInflow += c.precip;
    const double RiverEvapoRemUse = remainingUse + riverEvapo;  // (0. + evaporation without water use)
    const double remainingUseRiverStart = remainingUse;
    const double eK = exp(-1. * K);
    const double RivEvapoRemUseMax = riverInflow + (K * prevR * eK) / (1. - eK);
    double Sr, transportedVolume;
    if (RiverEvapoRemUse > RivEvapoRemUseMax) {
        Sr = 0.;
        transportedVolume = riverInflow + prevR - RivEvapoRemUseMax;
        if (transportedVolume < 0.) transportedVolume = 0.;
        if (remainingUse > 0.) remainingUse -= remainingUse * RivEvapoRemUseMax / RiverEvapoRemUse;  // :3489-3492
        else remainingUse = 0.;
        riverEvapo *= RivEvapoRemUseMax / RiverEvapoRemUse;
    } else {
        Sr = prevR * eK + (1. / K) * (riverInflow - RiverEvapoRemUse) * (1. - eK);
        if (fabs(Sr) <= MIN_STOR_VOL) Sr = 0.;
        transportedVolume = riverInflow + prevR - Sr - RiverEvapoRemUse;
        if (transportedVolume < 0.) transportedVolume = 0.;
        remainingUse = 0.;  // :3512
    }
    if (wu) {
        // what the river could not supply is taken from the local lake (:3590-3624), whose reduction factor is formed anew
        const double loc_lake = a.loc_lake[r];
        if ((loc_lake > 0.) && (remainingUse > 0.)) {
            const double maxStorage = ((loc_lake) / C100) * a.area[r] * a.lake_depth_active[q];
            double S = a.loc_lake_stor[i];
            if (S > (-1.) * maxStorage) {
                if (remainingUse < (maxStorage + S)) {
                    S -= remainingUse;
                    remainingUse = 0.;
                } else {
                    remainingUse -= (maxStorage + S);
                    S = (-1.) * maxStorage;
                }
            }
            a.loc_lake_stor[i] = S;
            const double red = clamp01(1. - wg_pow(fabs(S - maxStorage) / (2. * maxStorage), (a.p_evaredex[q] * 3.32193)));
            a.red_loc_lake[i] = red;
            if (red_loc_lake_out) *red_loc_lake_out = red;
        }
        dailyActualUse += remainingUseRiverStart - remainingUse;  // :3625
        a.wu_actual_use[i] += dailyActualUse;                     // :3736
        a.wu_total_unsatisfied[i] += remainingUse;                // :3897-3900 (use_alloc 0, delayedUseSatisfaction 0)
        a.wu_daily_remaining[i] = remainingUse;
        a.wu_wusi[i] = a.wu_wusi_month[r];                        // :3907-3908
        a.wu_cusi[i] = a.wu_cusi_month[r];
    }
    // (inland sinks have no downstream cell; the reference keeps their river outflow out of the
    //  discharge grid, routing.cpp:4219-4221, and books it as evaporation, :3935-3937)
    const bool out = (c.flags & FL_LDD_OUT) != 0;
    qday[i] = out ? transportedVolume : 0.;
    if (qout) *qout = out ? transportedVolume : 0.;
    a.cell_runoff[i] = out ? (transportedVolume - inflowUpstream) : (0. - inflowUpstream);
    a.river_stor[i] = Sr;
    a.river_evapo[i] = riverEvapo;
    return Sr;
}

__device__ __forceinline__ double gather_upstream(const WgkParams &p, const RiverCtx &c, const int m,
                                                  const double *qday) {
    // upstream inflow in routing order (= order of the += at routing.cpp:3957)
    double s = 0.;
    for (int k = c.up0; k < c.up1; k++) s += qday[mi(p, m, p.up_idx[k])];
    return s;
}

__device__ __forceinline__ double *qbuf_of_day(const WgkParams &p, const int dayofs) {
    return p.qbuf + (size_t)(dayofs % QBUF_K) * p.mpad * p.stride;
}

// one dependency level per launch (wide levels), routing sweep only
__global__ void __launch_bounds__(128) k_route_level(const __grid_constant__ WgkParams p, const int dayofs, const int level) {
    int r, m;
    if (!map_thread(p, p.level_off[level], p.level_off[level + 1], r, m)) return;
    const size_t i = mi(p, m, r), q = qi(p, m, r);
    const RiverCtx c = load_ctx(p, r, i, q);
    if (!(c.flags & FL_ACTIVE)) return;
    double *qday = qbuf_of_day(p, dayofs);
    route_river(p, c, r, m, i, q, gather_upstream(p, c, m, qday), p.cal_days[4 * dayofs], p.cal_days[4 * dayofs + 1], qday);
}

// levels [level_lo, level_hi) inside one persistent CTA per member; levels are separated by
// __syncthreads(), which also orders the global-memory hand-off of the discharge values.  The
// context of the next level's cell is loaded BEFORE the barrier so that only the gather of the
// upstream discharges and the river arithmetic remain on the level-to-level critical path.
__device__ __forceinline__ void sweep_levels(const WgkParams &p, const int m, const int dayofs, const int level_lo, const int level_hi) {
    const int day = p.cal_days[4 * dayofs], month = p.cal_days[4 * dayofs + 1];
    double *qday = qbuf_of_day(p, dayofs);
    int begin = p.level_off[level_lo], end = p.level_off[level_lo + 1];
    int r = begin + threadIdx.x;
    RiverCtx c;
    c.flags = 0;
    if (r < end) c = load_ctx(p, r, mi(p, m, r), qi(p, m, r));
    for (int level = level_lo; level < level_hi; level++) {
        if (r < end) {
            if (c.flags & FL_ACTIVE) route_river(p, c, r, m, mi(p, m, r), qi(p, m, r), gather_upstream(p, c, m, qday), day, month, qday);
            // levels wider than the CTA (only possible when the tail threshold is raised)
            for (int r2 = r + blockDim.x; r2 < end; r2 += blockDim.x) {
                const RiverCtx c2 = load_ctx(p, r2, mi(p, m, r2), qi(p, m, r2));
                if (c2.flags & FL_ACTIVE) route_river(p, c2, r2, m, mi(p, m, r2), qi(p, m, r2), gather_upstream(p, c2, m, qday), day, month, qday);
            }
        }
        if (level + 1 < level_hi) {
            begin = end;
            end = p.level_off[level + 2];
            r = begin + threadIdx.x;
            c.flags = 0;
            if (r < end) c = load_ctx(p, r, mi(p, m, r), qi(p, m, r));
        }
        __syncthreads();
    }
}

#if !WGK_MM  // one persistent CTA per member: cell-minor layout only
__global__ void __launch_bounds__(256) k_route_tail(const __grid_constant__ WgkParams p, const int dayofs, const int level_lo, const int level_hi) {
    sweep_levels(p, blockIdx.x, dayofs, level_lo, level_hi);
}

#endif  // !WGK_MM
// ----------------------------------------------------------------------------------------
// cell-parallel post-pass: river width / area fraction of the next day (:3546-3586), surface
// water body fractions and next-day land area fraction (:5034-5188), updateLandAreaFrac
// (:5343-5352)
// ----------------------------------------------------------------------------------------
// inputs of the post-pass that do not depend on today's river step: loaded together with the river
// context, before the hand-off from the upstream level is awaited
struct PostIn {
    double contf, area, red_loc_lake, red_loc_wetl, red_glo_wetl, red_river, raf_next, raf_change, laf, glo_wetland;
    double river_length, bw, wbf, smaxr, evaredex, loc_lake, loc_wetland, fswb_old, f_glo_lake;
    int flags;
};

__device__ __forceinline__ PostIn post_load(const WgkParams &p, const int r, const int m) {
    const WgkArrays &a = p.a;
    const size_t i = mi(p, m, r);
    const size_t q = qi(p, m, r);
    PostIn in;
    in.flags = a.s_flags[r];
    in.contf = a.contfreq[r];
    in.area = a.area[r];
    in.red_loc_lake = a.red_loc_lake[i];
    in.red_loc_wetl = a.red_loc_wetl[i];
    in.red_glo_wetl = a.red_glo_wetl[i];
    in.red_river = a.red_river[i];
    in.raf_next = a.river_area_frac_next[i];
    in.raf_change = a.river_area_frac_change[i];
    in.laf = a.land_area_frac[i];
    in.glo_wetland = a.glo_wetland[r];
    in.river_length = a.river_length[r];
    in.bw = a.river_bottom_width[r];
    in.wbf = a.river_width_bf[r];
    in.smaxr = a.river_storage_max[r];
    in.evaredex = a.p_evaredex[q];
    in.loc_lake = a.loc_lake[r];
    in.loc_wetland = a.loc_wetland[r];
    in.fswb_old = a.fswb_laf_next[i];
    in.f_glo_lake = a.f_glo_lake[r];
    return in;
}

__device__ __forceinline__ void route_post_compute(const WgkParams &p, const int r, const int m, const PostIn &in, const double Sr) {
    const WgkArrays &a = p.a;
    const size_t i = mi(p, m, r);
    const int flags = in.flags;
    const double contf = in.contf;
    const double cellArea = in.area;
    double red_loc_lake = in.red_loc_lake, red_loc_wetl = in.red_loc_wetl, red_glo_wetl = in.red_glo_wetl;
    double red_river = in.red_river;
    double raf_next = in.raf_next;
    double raf_change = in.raf_change;
    const double laf = in.laf;
    const double glo_wetland = in.glo_wetland;
    if (flags & FL_ACTIVE) {
        const double raf = raf_next;  // G_riverAreaFrac[n] = G_riverAreaFracNextTimestep_Frac[n] (:3424)
        const double river_length = in.river_length;
        const double bw = in.bw;
        const double crossSectionalArea = Sr / river_length;
        const double riverDepth = -bw / (4. * 1000.) + sqrt(bw / C1000 * bw / (16. * 1000.) + 0.5 * crossSectionalArea);
        double width = bw / C1000 + 4. * riverDepth;
        const double wbf = in.wbf;
        if (width > wbf / C1000) width = wbf / C1000;
        const double smaxr = in.smaxr;
        red_river = clamp01(1. - wg_pow(fabs(Sr - smaxr) / smaxr, (in.evaredex * 3.32193)));
        raf_next = red_river * river_length * width * 100. / cellArea;
        raf_change = raf_next - raf;
    }
    if ((flags & FL_ACTIVE) && (flags & (FL_LAKE | FL_RES | FL_GLOWET))) {
        // evaporation reduction factors of the global water bodies (:2790-2802, 3068-3078, 3287-3296): the six inputs in one round
        // of loads; in a warp that holds lakes, reservoirs and wetlands the three pow() are evaluated back to back for every lane
        // (an absent body gets harmless arguments and no store), so that their chains overlap instead of following each other
        const double *g = p.gbody + gb(p, m, p.gidx[r], GB_N);  // (no __restrict__: written earlier by the same thread in k_days_persistent)
        const size_t gs = gbody_stride(p);
        const double xexp = (in.evaredex * 3.32193);
        const bool hasL = (flags & FL_LAKE) != 0, hasR = (flags & FL_RES) != 0, hasW = (flags & FL_GLOWET) != 0;
        const double maxL = hasL ? g[GB_L_MAX * gs] : 1., storL = hasL ? a.glo_lake_stor[i] : 0.;
        const double maxR = hasR ? g[GB_R_CAP * gs] : 1., storR = hasR ? a.res_stor[i] : 0.;
        const double maxW = hasW ? g[GB_W_MAX * gs] : 1., storW = hasW ? a.glo_wetl_stor[i] : 0.;
        // how many kinds of water body the lanes of this warp hold: with one kind a plain branch (no wasted pow - a member-minor
        // warp is 32 members of ONE cell), with several all three back to back
#if defined(__CUDA_ARCH__) && !defined(WGK_NO_POW3)
        const unsigned am = __activemask();
        const int kinds = (__ballot_sync(am, hasL) != 0) + (__ballot_sync(am, hasR) != 0) + (__ballot_sync(am, hasW) != 0);
#else
        const int kinds = 1;
#endif
        if (kinds > 1) {
            const double powL = wg_pow(fabs(storL - maxL) / (2. * maxL), xexp);
            const double powR = wg_pow(fabs(storR - maxR) / maxR, 2.81383);
            const double powW = wg_pow(fabs(storW - maxW) / maxW, xexp);
            if (hasL) a.red_glo_lake[i] = clamp01(1. - powL);
            if (hasR) a.red_res[i] = clamp01(1. - powR);
            if (hasW) red_glo_wetl = clamp01(1. - powW);
        } else {
            if (hasL) a.red_glo_lake[i] = clamp01(1. - wg_pow(fabs(storL - maxL) / (2. * maxL), xexp));
            if (hasR) a.red_res[i] = clamp01(1. - wg_pow(fabs(storR - maxR) / maxR, 2.81383));
            if (hasW) red_glo_wetl = clamp01(1. - wg_pow(fabs(storW - maxW) / maxW, xexp));
        }
    }
    const double loc_lake = in.loc_lake, loc_wetland = in.loc_wetland;
    double fLocLake = ((loc_lake > 0.) && (red_loc_lake > 0.)) ? (red_loc_lake * loc_lake / C100) : 0.;
    double fLocWet = ((loc_wetland > 0.) && (red_loc_wetl > 0.)) ? (red_loc_wetl * loc_wetland / C100) : 0.;
    double fGloWet = ((glo_wetland > 0.) && (red_glo_wetl > 0.)) ? (red_glo_wetl * glo_wetland / C100) : 0.;
    if (p.month_acc) {  // the uncorrected fractions, as the checkpoint keeps them (routing.cpp:5050-5071)
        a.f_loc_lake[i] = fLocLake;
        a.f_loc_wet[i] = fLocWet;
        a.f_glo_wet[i] = fGloWet;
    }
    const double fswb_old = in.fswb_old;
    double fswb_next = fLocLake + fLocWet + fGloWet;
    const double fGloLake = in.f_glo_lake;
    const double maxRiverAreaFrac = contf / C100 - fGloLake;
    if ((flags & (FL_LAKE | FL_RES)) && (fGloLake == 1.)) {
        raf_next = 0.;
        raf_change = 0.;
        red_river = 0.;
        a.river_evapo[i] = 0.;
    } else {
        if (raf_next <= maxRiverAreaFrac) {
            if (fswb_next > (maxRiverAreaFrac - raf_next)) {
                const double fswbFracCorr = (maxRiverAreaFrac - raf_next) / fswb_next;
                red_loc_lake *= fswbFracCorr;
                red_loc_wetl *= fswbFracCorr;
                red_glo_wetl *= fswbFracCorr;
                if ((fLocLake > 0.) && (red_loc_lake > 0.)) fLocLake = (red_loc_lake * loc_lake / C100);
                else { red_loc_lake = 0.; fLocLake = 0.; }
                if ((fLocWet > 0.) && (red_loc_wetl > 0.)) fLocWet = (red_loc_wetl * loc_wetland / C100);
                else { red_loc_wetl = 0.; fLocWet = 0.; }
                if ((fGloWet > 0.) && (red_glo_wetl > 0.)) fGloWet = (red_glo_wetl * glo_wetland / C100);
                else { red_glo_wetl = 0.; fGloWet = 0.; }
            }
        } else {
            const double riverAreaFracDeficit = raf_next - maxRiverAreaFrac;
            raf_change -= riverAreaFracDeficit;
            red_river *= maxRiverAreaFrac / raf_next;
            raf_next = maxRiverAreaFrac;
            fLocLake = 0.;
            fLocWet = 0.;
            fGloWet = 0.;
        }
    }
    fswb_next = fLocLake + fLocWet + fGloWet;
    const double changePct = fswb_next * 100. - fswb_old * 100.;
    double laf_next = laf - (changePct + (raf_change * 100.));
    if (laf_next < 0.) laf_next = 0.;
    a.red_loc_lake[i] = red_loc_lake;
    a.red_loc_wetl[i] = red_loc_wetl;
    a.red_glo_wetl[i] = red_glo_wetl;
    a.red_river[i] = red_river;
    a.river_area_frac_next[i] = raf_next;
    a.river_area_frac_change[i] = raf_change;
    a.fswb_laf[i] = fswb_old;
    a.fswb_laf_next[i] = fswb_next;
    if (p.month_acc) {
        // the day's entry of WghmStateFile for the seven routing compartments (routing.cpp:5002-5020), summed in
        // day order as Cell::mean does (wghmStateFile.cpp:711-728)
        const double denom = ((cellArea * (contf / C100)) / C1E6);
        double *__restrict__ acc = a.mon_acc + bi(p, m, r, 7);
        const size_t bs = band_stride(p);
        acc[0 * bs] += a.loc_lake_stor[i] / denom;
        acc[1 * bs] += a.loc_wetl_stor[i] / denom;
        acc[2 * bs] += a.glo_lake_stor[i] / denom;
        acc[3 * bs] += a.glo_wetl_stor[i] / denom;
        acc[4 * bs] += a.res_stor[i] / denom;
        acc[5 * bs] += Sr / denom;
        acc[6 * bs] += a.gw[i] / denom;
    }
    a.status_laf_next[i] = 1;
    // updateLandAreaFrac fused: prev <- cur, cur <- next
    a.land_area_frac_next[i] = laf_next;
    a.land_area_frac_prev[i] = laf;
    a.land_area_frac[i] = laf_next;
}

__device__ __forceinline__ void route_post_cell(const WgkParams &p, const int r, const int m) {
    const PostIn in = post_load(p, r, m);
    route_post_compute(p, r, m, in, p.a.river_stor[mi(p, m, r)]);
}

#if !WGK_WU  // no water use inside
__global__ void __launch_bounds__(128) k_route_post(const __grid_constant__ WgkParams p) {
    int r, m;
    if (!map_thread(p, 0, p.ncell, r, m)) return;
    route_post_cell(p, r, m);
}

#endif  // !WGK_WU
// ----------------------------------------------------------------------------------------
// temporal wavefront: one kernel per (day, dependency level).  A cell of level l on day d needs
// its own state of day d-1 and the discharge of its upstream cells (levels < l) of day d, so
// (d, l), (d+1, l-1), (d+2, l-2) ... are independent and run concurrently as branches of one
// CUDA graph; the level-to-level latency chain of a day is hidden behind the work of the
// following days instead of serialising the run.
// ----------------------------------------------------------------------------------------
// river reach + post-pass of the cells of one wide level (the part of the day that waits for the
// upstream level); the vertical balance and the local routing of the same cells run in a separate,
// earlier task (k_cells_pre) that only waits for the cells' own previous day
__global__ void __launch_bounds__(128) k_river_level(const __grid_constant__ WgkParams p, const int dayofs, const int level) {
    int r, m;
    if (!map_thread(p, p.level_off[level], p.level_off[level + 1], r, m)) return;
    const size_t i = mi(p, m, r), q = qi(p, m, r);
    WGK_INSITU_BEGIN();
    if (level == 0) WGK_INSITU_STAMP(1, 0, dayofs);
    if (level == 0) stamp_task(p, 1, 0, dayofs);
    const RiverCtx c = load_ctx(p, r, i, q);
    PostIn in = post_load(p, r, m);  // same round of loads as the river context
    double Sr = c.prevR;
    if (c.flags & FL_ACTIVE) {
        double *qday = qbuf_of_day(p, dayofs);
        double red_ll = in.red_loc_lake;  // (water use may take from the local lake after the river: its reduction factor changes)
        Sr = route_river(p, c, r, m, i, q, gather_upstream(p, c, m, qday), p.cal_days[4 * dayofs], p.cal_days[4 * dayofs + 1], qday, nullptr,
                         WU ? &red_ll : nullptr);
        if (WU) in.red_loc_lake = red_ll;
    }
    route_post_compute(p, r, m, in, Sr);
    WGK_INSITU_END(2, level == 0);
    if (level == 0) WGK_INSITU_STAMP(1, 1, dayofs);
    if (level == 0) stamp_task(p, 1, 1, dayofs);
}

// vertical balance (+ local routing) of the cells [begin, end), one CTA per tile of 32 cells; tiles
// start at a multiple of 4 cells so that every 256 B band row of a warp is sector aligned
__host__ __device__ __forceinline__ int v_num_tiles(const int begin, const int end) { return (end - (begin & ~3) + 31) / 32; }

// warps 1-4: bring the per-cell inputs of the head, the tail and the local routing into shared memory
// (cp.async for 4/8/16-byte items, plain loads for the 1/2-byte flags) while warp 0 is in v_mode
template <class C>
__device__ __forceinline__ void v_preload(const WgkParams &p, VTile<C> &sm, const int r0, const int begin, const int end, const int m,
                                          const int slot, const int w, const int lane) {
    const WgkArrays &a = p.a;
    const int r = r0 + lane;
    if (!C::PRE || r < begin || r >= end) return;
    const size_t i = mi(p, m, r);
    const size_t q = qi(p, m, r);
    // four groups of copies, dealt round-robin to the warps 1 .. NW-1
#pragma unroll
    for (int task = 1; task <= 4; task++) {
    if (1 + (task - 1) % (C::NW > 1 ? C::NW - 1 : 1) != w) continue;
#define VP_D(slot_, src_) __pipeline_memcpy_async(&sm.pd[slot_][lane], &(src_), sizeof(double))
#define VP_F(slot_, src_) __pipeline_memcpy_async(&sm.pf[slot_][lane], &(src_), sizeof(float))
    if (task == 1) {
        VP_D(LI_ekg, a.s_ekg[q]); VP_D(LI_invkg, a.s_invkg[q]); VP_D(LI_evaredex, a.p_evaredex[q]); VP_D(LI_area, a.area[r]); VP_D(LI_cfa, a.cfa[q]);
        VP_D(LI_contf, a.contfreq[r]); VP_D(LI_fswb_init, a.fswb_init[r]); VP_D(LI_loc_lake, a.loc_lake[r]);
        VP_D(LI_loc_wetland, a.loc_wetland[r]); VP_D(LI_kS, a.p_swoutf[q]); VP_D(LI_lake_depth, a.lake_depth_active[q]);
        VP_D(LI_wetl_depth, a.wetl_depth_active[q]);
    } else if (task == 2) {
        VP_D(LI_laf, a.land_area_frac[i]); VP_D(LI_laf_prev, a.land_area_frac_prev[i]); VP_D(LI_gw, a.gw[i]);
        VP_D(LI_loc_lake_stor, a.loc_lake_stor[i]); VP_D(LI_red_loc_lake, a.red_loc_lake[i]);
        VP_D(LI_loc_wetl_stor, a.loc_wetl_stor[i]); VP_D(LI_red_loc_wetl, a.red_loc_wetl[i]);
        VP_D(LI_raf_next, a.river_area_frac_next[i]); VP_D(LI_transfer_old, a.storage_transfer[i]);
        sm.pk[K_contcell][lane] = a.contcell[r];
        sm.pk[K_flags][lane] = a.s_flags[r];
        sm.pk[K_ldd][lane] = a.ldd[r];
    } else if (task == 3) {
        __pipeline_memcpy_async(&sm.pforce[lane],
                                &p.forcing[fi(p, slot, m, r)],
                                sizeof(float4));
        VP_D(HI_p_prec, a.p_prec[q]); VP_D(HI_ptc_ari, a.p_ptc_ari[q]); VP_D(HI_ptc_hum, a.p_ptc_hum[q]);
        VP_D(HI_lai_precsum, a.lai_precsum[i]); VP_D(HI_snow, a.snow[i]); VP_D(HI_netrad, a.p_netrad[q]);
        VP_D(HI_canopy, a.canopy[i]); VP_D(HI_mcwh, a.p_mcwh[q]); VP_D(HI_pet_mxdy, a.p_pet_mxdy[q]);
        VP_F(HI_laimax, a.laimax[q]);
        __pipeline_memcpy_async(&sm.pk[K_lai_days][lane], &a.lai_days[i], sizeof(int32_t));
        __pipeline_memcpy_async(&sm.pk[K_lai_status][lane], &a.lai_status[i], sizeof(int32_t));
        sm.pk[K_lc][lane] = a.landcover[r];
        sm.pk[K_arid][lane] = a.arid[r];
    } else if (task == 4) {
        VP_D(TI_soil, a.soil[i]); VP_D(TI_gamma, a.gamma_hbv[q]); VP_D(TI_pcrit, a.p_pcrit[q]);
        VP_F(TI_builtup, a.builtup[r]); VP_F(TI_smax, a.smax[q]); VP_F(TI_gwfactor, a.gwfactor[q]);
        sm.pk[K_texture][lane] = a.texture[r];
        sm.pk[K_rgmax][lane] = a.rgmax[q];
    }
#undef VP_D
#undef VP_F
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
}

template <class C>
__device__ __forceinline__ LocalIn local_from_tile(const VTile<C> &sm, const int lane) {
    LocalIn li;
    li.contcell = sm.pk[K_contcell][lane];
    li.flags = sm.pk[K_flags][lane];
    li.ldd = sm.pk[K_ldd][lane];
    li.arid = sm.pk[K_arid][lane];
    li.ekg = sm.pd[LI_ekg][lane];
    li.invkg = sm.pd[LI_invkg][lane];
    li.evaredex = sm.pd[LI_evaredex][lane];
    li.area = sm.pd[LI_area][lane];
    li.cfa = sm.pd[LI_cfa][lane];
    li.contf = sm.pd[LI_contf][lane];
    li.fswb_init = sm.pd[LI_fswb_init][lane];
    li.loc_lake = sm.pd[LI_loc_lake][lane];
    li.loc_wetland = sm.pd[LI_loc_wetland][lane];
    li.kS = sm.pd[LI_kS][lane];
    li.lake_depth = sm.pd[LI_lake_depth][lane];
    li.wetl_depth = sm.pd[LI_wetl_depth][lane];
    li.laf = sm.pd[LI_laf][lane];
    li.laf_prev = sm.pd[LI_laf_prev][lane];
    li.gw = sm.pd[LI_gw][lane];
    li.loc_lake_stor = sm.pd[LI_loc_lake_stor][lane];
    li.red_loc_lake = sm.pd[LI_red_loc_lake][lane];
    li.loc_wetl_stor = sm.pd[LI_loc_wetl_stor][lane];
    li.red_loc_wetl = sm.pd[LI_red_loc_wetl][lane];
    li.raf_next = sm.pd[LI_raf_next][lane];
    return li;
}

// tail of the tile (warp 0): soil / runoff, then the local routing with the fluxes handed over in registers
template <class C, bool LOCAL>
__device__ __forceinline__ void v_finish(const WgkParams &p, VTile<C> &sm, const int r0, const int begin, const int end, const int m,
                                         const int lane, const int wu_month) {
    double flux[3] = {0., 0., 0.};
    v_tail<C>(p, sm, r0, m, lane, flux);
    const int r = r0 + lane;
    if (LOCAL && r >= begin && r < end) {
        LocalFlux fx;
        if (sm.mode[lane] & VM_ACTIVE) {
            fx.owPrec = sm.h_prec[lane];
            fx.owPET = sm.h_owpet[lane];
            fx.storage_transfer = flux[0];
            fx.surface_runoff = flux[1];
            fx.gw_recharge = flux[2];
        } else {
            fx = local_flux_load(p, r, m);  // cells outside the computed region keep their last fluxes
        }
        local_compute(p, r, m, C::PRE ? local_from_tile<C>(sm, lane) : local_load(p, r, m), fx, wu_month);
    }
}

#ifdef WGK_PHASE_TIMING  // development aid: SM-clock cycles per phase of the tile kernels, summed over tiles
__device__ unsigned long long g_phase[8];
#define WGK_TICK(k_) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_phase[k_], (unsigned long long)(t_ - tick_)); tick_ = t_; } } while (0)
#define WGK_TICK0() long long tick_ = clock64()
#else
#define WGK_TICK(k_) do { } while (0)
#define WGK_TICK0() do { } while (0)
#endif

template <class C, bool LOCAL>
__device__ __forceinline__ void vertical_tile(const WgkParams &p, VTile<C> &sm, const int begin, const int end, const int m, const int dayofs) {
    const int slot = p.cal_days[4 * dayofs + 3];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = (begin & ~3) + 32 * blockIdx.x;
    VThread<C> ts;
    WGK_TICK0();
    if (w == 0) v_mode<C>(p, sm, r0, begin, end, m, slot, lane);
    else v_preload<C>(p, sm, r0, begin, end, m, slot, w, lane);
    __syncthreads();
    WGK_TICK(0);
    const bool bands = sm.nband_cells != 0;  // else nothing but bare / inactive cells: the band arrays are not touched
    if (!bands && w != 0) return;
    if (bands) v_prefetch<C>(p, sm, ts, r0, m, 0, w, lane);
    if (w == 0) v_head<C>(p, sm, r0, m, slot, lane);
    WGK_TICK(1);
    if (bands) {
        __syncthreads();
#pragma unroll 1
        for (int slab = 0; slab < C::NSLAB; slab++) {
            v_scale<C>(p, sm, ts, r0, m, slab, w, lane);
            __syncthreads();
            if (sm.cap_any[slab]) {
                v_cap_resolve<C>(sm, ts, slab, w, lane);
                __syncthreads();
            }
            v_band<C>(p, sm, ts, r0, m, slab, w, lane);
            __syncthreads();
            if (w < 4) v_sum<C>(sm, ts, slab, w, lane);
        }
        __syncthreads();
        WGK_TICK(2);
        if (w != 0) return;
    } else {
        v_bare_sums<C>(p, sm, r0, lane);
        WGK_TICK(3);
    }
    v_finish<C, LOCAL>(p, sm, r0, begin, end, m, lane, p.cal_days[4 * dayofs + 1]);
    WGK_TICK(4);
#ifdef WGK_PHASE_TIMING
    if (threadIdx.x == 0) atomicAdd(&g_phase[bands ? 6 : 7], 1ull);
#endif
}

#if !WGK_MM  // the band-parallel tile kernels run on the cell-minor layout only
template <class C>
__global__ void __launch_bounds__(C::THREADS, C::NW == 5 ? 6 : 12) k_cells_pre(const __grid_constant__ WgkParams p, const int dayofs, const int begin, const int end) {
    __shared__ VTile<C> sm;
    vertical_tile<C, true>(p, sm, begin, end, blockIdx.y, dayofs);
}

#if !WGK_WU  // no water use inside
// the vertical balance alone over the whole grid (wgk_vertical_day, wgk_profile_day)
template <class C>
__global__ void __launch_bounds__(C::THREADS, C::NW == 5 ? 6 : 12) k_vertical(const __grid_constant__ WgkParams p, const int dayofs) {
    __shared__ VTile<C> sm;
    vertical_tile<C, false>(p, sm, 0, p.ncell, blockIdx.y, dayofs);
}

#endif  // !WGK_MM
#endif  // !WGK_WU
// thread-per-cell forms of k_vertical and k_cells_pre
#if !WGK_WU  // no water use inside
__global__ void __launch_bounds__(VBLOCK, WGK_TPC_MINB_EFF) k_vertical_tpc(const __grid_constant__ WgkParams p, const int dayofs) {
    __shared__ SnowStage stage;
    int r, m;
    if (!map_thread(p, 0, p.ncell, r, m)) return;
    vertical_cell(p, r, m, p.cal_days[4 * dayofs + 3], &stage);
}


#endif  // !WGK_WU
#ifndef WGK_PRE_MINB_MM
#define WGK_PRE_MINB_MM 4  // vertical + local routing in one kernel spills at 96 registers; measured on B200 (10^9 cell-days/s, wavefront,
                           // 4 / 5 resident CTAs): 32 members 1.78 / 1.67, 64 members 2.00 / 1.97
#endif
#undef WGK_PRE_MINB_EFF
#if WGK_MM
#define WGK_PRE_MINB_EFF WGK_PRE_MINB_MM
#else
#define WGK_PRE_MINB_EFF WGK_TPC_MINB
#endif
__global__ void __launch_bounds__(VBLOCK, WGK_PRE_MINB_EFF) k_cells_pre_tpc(const __grid_constant__ WgkParams p, const int dayofs, const int begin, const int end) {
    __shared__ SnowStage stage;
    int r, m;
    if (!map_thread(p, begin, end, r, m)) return;
    WGK_INSITU_BEGIN();
    if (begin == 0) WGK_INSITU_STAMP(0, 0, dayofs);
    if (begin == 0) stamp_task(p, 0, 0, dayofs);
    LocalIn li;
    LocalFlux fx;
    if (vertical_cell(p, r, m, p.cal_days[4 * dayofs + 3], &stage, &li, &fx)) local_compute(p, r, m, li, fx, p.cal_days[4 * dayofs + 1]);
    else route_local_cell(p, r, m, p.cal_days[4 * dayofs + 1]);
    WGK_INSITU_END(0, begin == 0);
    if (begin == 0) WGK_INSITU_WARPDUR(dayofs);
    if (begin == 0) WGK_INSITU_STAMP(0, 1, dayofs);
    if (begin == 0) stamp_task(p, 0, 1, dayofs);
}

// V(d, l) and R(d, l) of one wide level in ONE task: the vertical balance + local routing of the level's cells, then - after
// the upstream level of the same day has completed - their river reach + post-pass, cell by cell in the same thread.  In the
// wavefront graph the edge from the upstream level's task is a PROGRAMMATIC one (launch-completion port): this grid starts as
// soon as the upstream grid is resident, works through its vertical part next to it, and griddepcontrol.wait holds it until the
// upstream grid has completed and its discharge is visible.  The own-cell recurrence of a level then crosses one kernel
// boundary per day instead of two, and the graph has half the nodes.  (Launched without a programmatic edge - plain stream
// order - the wait returns at once.)
#ifndef WGK_LEVEL_MINB
#define WGK_LEVEL_MINB 2  // the fused task holds the V and the R part in one register budget: no cap (164 registers); with the
                          // 128 of the split kernels it measured 20.6 instead of 19.9 ms per simulated year
#endif
#ifndef WGK_LEVEL_EARLY
#define WGK_LEVEL_EARLY 0
#endif
#undef WGK_LEVEL_MINB_EFF
#if WGK_MM
#define WGK_LEVEL_MINB_EFF WGK_PRE_MINB_MM
#else
#define WGK_LEVEL_MINB_EFF WGK_LEVEL_MINB
#endif
__global__ void __launch_bounds__(VBLOCK, WGK_LEVEL_MINB_EFF) k_level_day(const __grid_constant__ WgkParams p, const int dayofs, const int level) {
    __shared__ SnowStage stage;
    int r, m;
    const bool on = map_thread(p, p.level_off[level], p.level_off[level + 1], r, m);
    WGK_INSITU_BEGIN();
    if (on) {
        if (level == p.stamp_level) stamp_task(p, 0, 0, dayofs);
        LocalIn li;
        LocalFlux fx;
        if (vertical_cell<WGK_LEVEL_EARLY != 0>(p, r, m, p.cal_days[4 * dayofs + 3], &stage, &li, &fx)) local_compute(p, r, m, li, fx, p.cal_days[4 * dayofs + 1]);
        else route_local_cell(p, r, m, p.cal_days[4 * dayofs + 1]);
        if (level == p.stamp_level) stamp_task(p, 0, 1, dayofs);
        if (level == 0) WGK_INSITU_WARPDUR2(dayofs, 0);
    }
    // everything the R part needs that does not come from the upstream level is loaded BEFORE the wait (own-cell values this
    // thread has just written, statics, the indices of the upstream cells): after the wait only the gather of the upstream
    // discharge is left between the hand-off and the arithmetic
    size_t i = 0, q = 0;
    RiverCtx c;
    PostIn in;
    int up[8];
    c.flags = 0;
    if (on) {
        i = mi(p, m, r);
        q = qi(p, m, r);
        c = load_ctx(p, r, i, q);
        in = post_load(p, r, m);
#pragma unroll
        for (int k = 0; k < 8; k++) up[k] = (c.up0 + k < c.up1) ? p.up_idx[c.up0 + k] : -1;
    }
#ifdef __CUDA_ARCH__
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
    if (!on) return;
    if (level == p.stamp_level) stamp_task(p, 1, 0, dayofs);
    double Sr = c.prevR;
    if (c.flags & FL_ACTIVE) {
        double *qday = qbuf_of_day(p, dayofs);
        double red_ll = in.red_loc_lake;
        // upstream inflow in routing order (= order of the += at routing.cpp:3957); a cell has at most 8 neighbours
        double inflow_up = 0.;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (up[k] >= 0) inflow_up += qday[mi(p, m, up[k])];
        for (int k = c.up0 + 8; k < c.up1; k++) inflow_up += qday[mi(p, m, p.up_idx[k])];  // (never on a D8 network)
        Sr = route_river<true>(p, c, r, m, i, q, inflow_up, p.cal_days[4 * dayofs], p.cal_days[4 * dayofs + 1], qday, nullptr,
                               WU ? &red_ll : nullptr);
        if (WU) in.red_loc_lake = red_ll;
    }
    route_post_compute(p, r, m, in, Sr);
    if (level == p.stamp_level) stamp_task(p, 1, 1, dayofs);
    if (level == 0) WGK_INSITU_WARPDUR2(dayofs, 1);
    WGK_INSITU_END(0, level == 0);
}

#if !WGK_MM  // k_tail_chunk and the cell-owner schedule run on the cell-minor layout only
// narrow levels [level_lo, level_hi) of one day in one persistent CTA per member, then the
// post-pass of those cells
__global__ void __launch_bounds__(256) k_tail_chunk(const __grid_constant__ WgkParams p, const int dayofs, const int level_lo, const int level_hi) {
    const int m = blockIdx.x;
    sweep_levels(p, m, dayofs, level_lo, level_hi);
    for (int r = p.level_off[level_lo] + threadIdx.x; r < p.level_off[level_hi]; r += blockDim.x) route_post_cell(p, r, m);
}

#ifndef WGK_EMU
// ----------------------------------------------------------------------------------------
// Cell-owner schedule: ONE launch for all days of a call.  Every thread owns one cell for the whole
// call and walks its days in order - vertical balance, local routing, river reach, post-pass - so the
// cell's own day -> day recurrence needs no kernel boundary at all.  The only exchange between
// threads is the river discharge handed to the downstream cell.  It travels through a ring of
// QBUF_K days of 16-byte entries, each 8-byte half carrying 32 bits of the value and a 32-bit day tag:
// an aligned 8-byte store is single-copy atomic, so a consumer that polls the entry until both tags
// name the day it needs holds a complete value - no fence on either side (a release fence, MEMBAR.GPU,
// measured ~50 us per warp and day here because it waits for the band stores of the whole SM; an
// acquire load empties the SM's L1).  Back pressure for the reuse of a ring slot: a warp stores the
// number of days it has consumed into its progress word, and a cell waits until the warp of its
// downstream cell is less than QBUF_K days behind before it overwrites an entry.
// Warps never straddle a dependency level (WgkOwner::warp_begin/end), so a lane never waits for its
// own warp; every wait is on a strictly earlier (day, level), and all CTAs are co-resident
// (cooperative launch), hence the schedule cannot deadlock.  A wait that exceeds WGK_SPIN_LIMIT cycles
// raises the abort word and ends the kernel (reported by the host) instead of hanging the GPU.
// The arithmetic is that of k_cells_pre_tpc + k_river_level: results are bit-identical to the
// (day, level) wavefront (tests/test_gpu_parity.py::test_cell_owner_schedule_equals_wavefront).
// MEASURED (B200, 0.5 degree grid, one member): 93 us per simulated day against 67 us of the
// wavefront graph, so this schedule is opt-in (WGK_DAY_SCHEDULE=owner) and not the default.  The cell-day is
// ~130 KB of SASS and the instruction cache holds 32 KB: warps that drift apart in the day loop wait for
// instructions (ncu: no_instruction 47 % of the vertical step, the step itself 2x slower); the day barrier that
// keeps a CTA's warps in step (below) repairs that but couples them to the slowest hand-off among them
// (CTAs of 512 / 256 / 128 threads: 112 / 90 / 90 us per day; without any hand-off wait the barrier form runs at 63 us).
// ----------------------------------------------------------------------------------------
constexpr long long WGK_SPIN_LIMIT = 4000000000LL;  // ~2 s at 1.965 GHz

// all polling goes to L2 (relaxed at GPU scope), never to the non-coherent L1
__device__ __forceinline__ uint32_t ld_progress(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_progress(uint32_t *p, const uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void ring_put(unsigned long long *e, const double v, const uint32_t tag) {
    const unsigned long long lo = ((unsigned long long)tag << 32) | (uint32_t)__double2loint(v);
    const unsigned long long hi = ((unsigned long long)tag << 32) | (uint32_t)__double2hiint(v);
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(e), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ bool ring_try(const unsigned long long *e, const uint32_t tag, double &v) {
    unsigned long long lo, hi;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(e) : "memory");
    v = __hiloint2double((int)(uint32_t)hi, (int)(uint32_t)lo);
    return (uint32_t)(lo >> 32) == tag && (uint32_t)(hi >> 32) == tag;
}
// a wait gives up when the run was aborted by any thread or after WGK_SPIN_LIMIT cycles
__device__ __forceinline__ bool spin_check(unsigned &it, const long long t0, int32_t *abort) {
    if ((++it & 63u) != 0u) return true;
    if (*(volatile int32_t *)abort) return false;
    if (clock64() - t0 > WGK_SPIN_LIMIT) {
        atomicExch(abort, 1);
        return false;
    }
    return true;
}
__device__ __forceinline__ bool ring_get(const unsigned long long *e, const uint32_t tag, double &v, int32_t *abort) {
    if (ring_try(e, tag, v)) return true;
    const long long t0 = clock64();
    unsigned it = 0;
    while (!ring_try(e, tag, v))
        if (!spin_check(it, t0, abort)) return false;
    return true;
}
// tags are compared as wrapping 32-bit distances
__device__ __forceinline__ bool wait_progress(const uint32_t *flag, const uint32_t target, int32_t *abort) {
    if ((int32_t)(ld_progress(flag) - target) >= 0) return true;
    const long long t0 = clock64();
    unsigned it = 0;
    while ((int32_t)(ld_progress(flag) - target) < 0)
        if (!spin_check(it, t0, abort)) return false;
    return true;
}

// CTAs of OWN_BLOCK threads with a CTA barrier at the start of every day: the warps of a CTA then walk
// the ~130 KB of code of a cell-day in step and share its instruction fetches (the instruction cache holds 32 KB;
// warps left to drift apart spent half of their vertical step waiting for instructions - ncu no_instruction 47 %).
#ifndef WGK_OWN_BLOCK
#define WGK_OWN_BLOCK 256  // measured: 512 -> 112 us, 256 -> 90 us, 128 -> 90 us per simulated day (wavefront graph: 67 us)
#endif
constexpr int OWN_BLOCK = WGK_OWN_BLOCK;
__global__ void __launch_bounds__(OWN_BLOCK, WGK_TPC_MINB * VBLOCK / OWN_BLOCK) k_days_owner(const __grid_constant__ WgkParams p, const __grid_constant__ WgkOwner s,
                                                            const int ndays) {
    extern __shared__ __align__(16) unsigned char own_smem[];
    SnowStage *stage = reinterpret_cast<SnowStage *>(own_smem) + threadIdx.x / VBLOCK;
    const int w = blockIdx.x * (OWN_BLOCK / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int r = w < s.nwarps ? s.warp_begin[w] + lane : 0;
    const bool valid = w < s.nwarps && r < s.warp_end[w];  // lanes without a cell only take part in the barriers
    const unsigned mask = __ballot_sync(0xffffffffu, valid);  // the lanes of this warp that own a cell
    const int m = blockIdx.y;
    const size_t mb = mi(p, m, 0);  // (the cell-owner schedule runs on the cell-minor layout only)
    const size_t q = qi(p, m, r);
    uint32_t *prog = s.progress + (size_t)m * s.nwarps;
    const int up0 = valid ? p.up_off[r] : 0, up1 = valid ? p.up_off[r + 1] : 0;
    const int dn = valid ? p.down[r] : -1;
    const int dn_w = dn >= 0 ? s.cell_warp[dn] : -1;
    const int rec0 = (valid && p.record && s.rec_head) ? s.rec_head[r] : -1;
    long long tacc[4] = {0, 0, 0, 0}, tk = s.timing ? clock64() : 0;
#define WGK_OWNER_TICK(k_) do { if (s.timing) { const long long t_ = clock64(); tacc[k_] += t_ - tk; tk = t_; } } while (0)
    bool ok = true;
    for (int d = 0; d < ndays; d++) {
        if (__syncthreads_or(!ok)) return;  // day barrier; a timed-out wait ends the CTA (the others see the abort word)
        if (!valid) continue;
        {   // vertical balance + local routing of the day (k_cells_pre_tpc)
            LocalIn li;
            LocalFlux fx;
            if (vertical_cell(p, r, m, p.cal_days[4 * d + 3], stage, &li, &fx)) local_compute(p, r, m, li, fx, p.cal_days[4 * d + 1]);
            else route_local_cell(p, r, m, p.cal_days[4 * d + 1]);
        }
        // river reach + post-pass (k_river_level); everything that does not depend on upstream cells is loaded
        // before the hand-off is awaited
        const RiverCtx c = load_ctx(p, r, mb + r, q);
        PostIn in = post_load(p, r, m);
        const uint32_t tag = s.base + (uint32_t)d + 1u;
        unsigned long long *ring = s.ring + ((size_t)(d % QBUF_K) * p.mpad * p.stride + mb) * 2;
        double Sr = c.prevR, qv = 0.;
        WGK_OWNER_TICK(0);
        if (c.flags & FL_ACTIVE) {
            double inflowUpstream = 0.;  // in routing order (= order of the += at routing.cpp:3957)
            for (int k = up0; k < up1 && ok; k++) {
                double v;
                ok = ring_get(ring + 2 * (size_t)p.up_idx[k], tag, v, s.abort);
                inflowUpstream += v;
            }
            WGK_OWNER_TICK(1);
            double red_ll = in.red_loc_lake;
            if (ok) Sr = route_river(p, c, r, m, mb + r, q, inflowUpstream, p.cal_days[4 * d], p.cal_days[4 * d + 1], p.a.discharge, &qv, WU ? &red_ll : nullptr);
            if (WU) in.red_loc_lake = red_ll;
        } else {
            p.a.discharge[mb + r] = 0.;
        }
        // ring slot reuse: the downstream cell's warp must have consumed day d - QBUF_K
        if (d >= QBUF_K && dn_w >= 0 && ok) ok = wait_progress(prog + dn_w, tag - (uint32_t)QBUF_K, s.abort);
        if (ok) ring_put(ring + 2 * (size_t)r, qv, tag);
        WGK_OWNER_TICK(2);
        if (ok) route_post_compute(p, r, m, in, Sr);
        if (rec0 >= 0 && p.cal[4] + d < p.record_max_days)
            for (int k = rec0; k >= 0; k = s.rec_next[k]) p.record[((size_t)(p.cal[4] + d) * p.nmember + m) * p.nrec + k] = qv;
        // once EVERY lane's gather of day d has returned (its values were used above) the warp has consumed day d;
        // the full mask matters: lanes still polling must not be overtaken by the producer of day d + QBUF_K
        __syncwarp(mask);
        if (lane == 0) st_progress(prog + w, tag);
        WGK_OWNER_TICK(3);
    }
    if (s.timing && valid && lane == 0)
        for (int k = 0; k < 4; k++) s.timing[((size_t)m * s.nwarps + w) * 4 + k] = tacc[k];
#undef WGK_OWNER_TICK
}
#endif  // WGK_EMU

#endif  // !WGK_MM
#if !WGK_WU  // calendar, forcing, diagnostics, state bridge: no water use inside
// ----------------------------------------------------------------------------------------
// calendar, forcing, diagnostics
// ----------------------------------------------------------------------------------------
// calendar of the `ndays` days of one call, starting at (day, month, day_in_month, slot);
// 365-day years (integrateWGHM.cpp:100-102), forcing slots cycle through the reserved ones
__global__ void k_fill_calendar(int32_t *cal_days, int32_t *cal, int day, int month, int dom, int slot, int ndays, int nslots, int rec_base) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    cal[0] = day; cal[1] = month; cal[2] = dom; cal[3] = slot;
    cal[4] = rec_base;  // row of the station record that day offset 0 of this call writes (graphs are replayed unchanged)
    const int nd[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    for (int d = 0; d < ndays; d++) {
        cal_days[4 * d] = day; cal_days[4 * d + 1] = month; cal_days[4 * d + 2] = dom; cal_days[4 * d + 3] = slot;
        dom++;
        day++;
        if (dom > nd[month]) { dom = 1; month++; }
        if (month > 11) { month = 0; day = 1; }
        slot++;
        if (slot >= nslots) slot = 0;
    }
}

// end of a simulated day: record the discharge of the station cells
__global__ void k_end_of_day(const __grid_constant__ WgkParams p, const int dayofs) {
    if (!p.record) return;
    const int row = p.cal[4] + dayofs;  // rows accumulate over the calls since the record was (re)started
    if (row >= p.record_max_days) return;
    const double *qday = qbuf_of_day(p, dayofs);
    const int total = p.nmember * p.nrec;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        const int m = k / p.nrec, cidx = k % p.nrec;
        p.record[(size_t)row * total + k] = qday[mi(p, m, p.record_cells[cidx])];
    }
}

// host grids [cell][stride] (reference order) staged on the device -> [slot][cell] float4 (device order).
// SWAP: the staged words are the bytes of the reference's big-endian .31 / .365 UNF0 files as they lie on
// disk (climate.cpp:100-123, gridio byte order), swapped here instead of on the host.
template <bool SWAP>
__global__ void k_forcing_pack(float4 *__restrict__ dst, const float *__restrict__ P, const float *__restrict__ T,
                               const float *__restrict__ SW, const float *__restrict__ LW,
                               const int32_t *__restrict__ cell_of_rank, int ncell, int ndays, int src_stride, size_t slot_pitch,
                               size_t cell_stride) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ncell) return;
    const int n = cell_of_rank[r];
    auto word = [](const float *a, const size_t k) -> float {
#ifndef WGK_EMU
        if (SWAP) return __uint_as_float(__byte_perm(__float_as_uint(a[k]), 0, 0x0123));
#endif
        return a[k];
    };
    for (int d = blockIdx.y; d < ndays; d += gridDim.y) {
        const size_t s = (size_t)n * src_stride + d;
        dst[(size_t)d * slot_pitch + (size_t)r * cell_stride] = make_float4(word(P, s), word(T, s), word(SW, s), word(LW, s));
    }
}

// total water storage of one member (km3), block-reduced in a fixed order so that the value
// is reproducible run to run
__global__ void __launch_bounds__(256) k_total_storage(const __grid_constant__ WgkParams p, const int m, double *partial) {
    __shared__ double sh[256];
    const WgkArrays &a = p.a;
    double s = 0.;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < p.ncell; r += gridDim.x * blockDim.x) {
        const size_t i = mi(p, m, r);
        const double laf = (0 == a.status_laf_next[i]) ? a.land_area_frac[i] : a.land_area_frac_next[i];
        const double land = (a.canopy[i] + a.snow[i] + a.soil[i]) * a.area[r] / C1E6 * laf / C100;
        s += land + a.gw[i] + a.loc_lake_stor[i] + a.loc_wetl_stor[i] + a.glo_lake_stor[i] + a.glo_wetl_stor[i]
             + a.res_stor[i] + a.river_stor[i];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// ----------------------------------------------------------------------------------------
// EnKF state bridge: the state vector PDAF sees and the analysis increment it hands back
// ----------------------------------------------------------------------------------------
// land area fraction as routingClass::getLandAreaFrac (routing.h:246-251)
__device__ __forceinline__ double laf_of(const WgkArrays &a, const size_t i) {
    return (0 == a.status_laf_next[i]) ? a.land_area_frac[i] : a.land_area_frac_next[i];
}

// the ten compartments of one cell in mm over the continental area, order of extractsub.cpp:65-79:
// canopy, snow, soil, local lake, local wetland, global lake, global wetland, reservoir, river, groundwater.
// kind 1: the last day (integrateWGHM.cpp:838-847, routing.cpp:5002-5020); kind 0: the mean of the month's
// `ndays` entries (Cell::mean): canopy / snow / soil carry the month-end value on every day of the month.
__device__ __forceinline__ void state_of_cell(const WgkParams &p, const int x, const int m, const int kind, const int ndays, double v[10]) {
    const WgkArrays &a = p.a;
    const size_t i = mi(p, m, x);
    const double laf = laf_of(a, i), contf = a.contfreq[x];
    const double land[3] = {a.canopy[i] * laf / contf, a.snow[i] * laf / contf, a.soil[i] * laf / contf};
    const double denom = ((a.area[x] * (contf / C100)) / C1E6);
    const double stor[7] = {a.loc_lake_stor[i], a.loc_wetl_stor[i], a.glo_lake_stor[i], a.glo_wetl_stor[i], a.res_stor[i], a.river_stor[i], a.gw[i]};
    if (kind == 1) {
        for (int k = 0; k < 3; k++) v[k] = land[k];
        for (int k = 0; k < 7; k++) v[3 + k] = stor[k] / denom;
    } else {
        for (int k = 0; k < 3; k++) {
            double s = 0.0;
            for (int d = 0; d < ndays; d++) s += land[k];
            v[k] = s / (double)ndays;
        }
        for (int k = 0; k < 7; k++) v[3 + k] = a.mon_acc[bi(p, m, x, 7) + (size_t)k * band_stride(p)] / (double)ndays;
    }
}

// extract_sub_ (extractsub.cpp:65-79): state vector of the region's cells minus the temporal mean field
__global__ void __launch_bounds__(128) k_state_vector(const __grid_constant__ WgkParams p, const int m, const int kind, const int ndays,
                                                      const int32_t *__restrict__ pos, const int ncells,
                                                      const double *__restrict__ mean_field, double *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncells) return;
    double v[10];
    state_of_cell(p, pos[j], m, kind, ndays, v);
    for (int k = 0; k < 10; k++) out[(size_t)j * 10 + k] = v[k] - (mean_field ? mean_field[(size_t)j * 10 + k] : 0.);
}

// the day's WghmStateFile entry of the seven routing compartments (routing.cpp:5002-5020: km3 -> mm over the continental area) of
// every cell, in REFERENCE cell order, out[k][n], k = local lake, local wetland, global lake, global wetland, reservoir, river,
// groundwater: one packed device-to-host copy per simulated day for the class shim routingClass::routing
__global__ void __launch_bounds__(128) k_pack_day_state(const __grid_constant__ WgkParams p, const int m, const int32_t *__restrict__ rank_of_cell,
                                                        double *__restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.ncell) return;
    const WgkArrays &a = p.a;
    const int x = rank_of_cell[n];
    const size_t i = mi(p, m, x);
    const double denom = ((a.area[x] * (a.contfreq[x] / C100)) / C1E6);
    const double stor[7] = {a.loc_lake_stor[i], a.loc_wetl_stor[i], a.glo_lake_stor[i], a.glo_wetl_stor[i], a.res_stor[i], a.river_stor[i], a.gw[i]};
#pragma unroll
    for (int k = 0; k < 7; k++) out[(size_t)k * p.ncell + n] = stor[k] / denom;
}

// Ensemble moments of the state vector over the members of this context (what an assimilation
// cycle exchanges once per month): sum[j][k] = sum over members of v, sumsq[j][k] = sum of v * v, v = the extract_sub_
// value of compartment k of cell j (state_of_cell, so the values are those of k_state_vector bit for bit).  Every
// member array is read exactly once, coalesced: a thread owns one device position and walks the members in
// ascending order (fixed summation order -> reproducible, and equal to a sequential host sum).  `pos` == nullptr:
// all cells, thread = device position, results stored at the reference cell number (cell_of_rank); else thread j
// gathers device position pos[j].  The two outputs are plain device buffers the host all-reduces over NCCL.
__global__ void __launch_bounds__(128) k_ensemble_moments(const __grid_constant__ WgkParams p, const int kind, const int ndays,
                                                          const int32_t *__restrict__ pos, const int32_t *__restrict__ cell_of_rank,
                                                          const int ncells, double *__restrict__ sum, double *__restrict__ sumsq) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncells) return;
    const int x = pos ? pos[j] : j;
    const size_t o = (size_t)(pos ? j : cell_of_rank[j]) * 10;
    double s[10], ss[10];
#pragma unroll
    for (int k = 0; k < 10; k++) { s[k] = 0.; ss[k] = 0.; }
    for (int m = 0; m < p.nmember; m++) {
        double v[10];
        state_of_cell(p, x, m, kind, ndays, v);
#pragma unroll
        for (int k = 0; k < 10; k++) {
            s[k] += v[k];
            ss[k] += v[k] * v[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 10; k++) {
        sum[o + k] = s[k];
        sumsq[o + k] = ss[k];
    }
}

// mean and (population) variance from the all-reduced sums, in place: sum <- mean, sumsq <- max(0, E[x^2] - mean^2)
__global__ void __launch_bounds__(256) k_moments_finish(double *__restrict__ sum, double *__restrict__ sumsq, const size_t n, const double nmember_total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double mean = sum[i] / nmember_total;
    double var = sumsq[i] / nmember_total - mean * mean;
    if (var < 0.) var = 0.;
    sum[i] = mean;
    sumsq[i] = var;
}

// DFMA throughput of this GPU (north_star: "FP64 pipe utilisation against peak"): every thread runs 8 independent chains
// of dependent fused multiply-adds from registers; flops = 2 * 8 * iters per thread.  `out` keeps the compiler honest.
__global__ void __launch_bounds__(256) k_fp64_peak(double *__restrict__ out, const int iters, const double a, const double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1., x2 = x0 + 2., x3 = x0 + 3., x4 = x0 + 4., x5 = x0 + 5., x6 = x0 + 6., x7 = x0 + 7.;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == 123.456) out[0] = r;  // never true for the arguments used
}

// uniform value for one row of a per-cell f64 array (a calibration parameter of one parameter set, calibration.cpp
// assigns one gamma / CFA per basin)
__global__ void __launch_bounds__(256) k_fill_f64(double *__restrict__ dst, const int n, const size_t stride, const double v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[(size_t)i * stride] = v;
}

// dst[k * dst_stride] = src[k * src_stride]: uploads, downloads and copies of one index of a member-minor array
template <class T>
__global__ void __launch_bounds__(256) k_strided_copy(T *__restrict__ dst, const size_t dst_stride, const T *__restrict__ src, const size_t src_stride,
                                                      const size_t n) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) dst[k * dst_stride] = src[k * src_stride];
}

// enkf_wghmstate_ (enKF2wghmState.cpp:89-121, 440-471) followed by the restore of the next cycle's start
// (daily.cpp:1896-1924, routing.cpp:851-882): the last day's state plus (analysis - prediction) with the
// reference's limits, the snow bands rescaled by assimilated / predicted monthly snow, back to device units
__global__ void __launch_bounds__(128) k_enkf_update(const __grid_constant__ WgkParams p, const int m, const int ndays,
                                                     const int32_t *__restrict__ pos, const int ncells,
                                                     const double *__restrict__ field, const double *__restrict__ prediction,
                                                     const double *__restrict__ mean_field) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncells) return;
    const WgkArrays &a = p.a;
    const int x = pos[j];
    const size_t i = mi(p, m, x);
    double w[10], mon[10];
    state_of_cell(p, x, m, 1, ndays, w);
    state_of_cell(p, x, m, 0, ndays, mon);
    const double *fl = field + (size_t)j * 10, *pr = prediction + (size_t)j * 10, *mf = mean_field + (size_t)j * 10;
    for (int k = 0; k < 10; k++) w[k] += (fl[k] - pr[k]);
    if (w[0] < 0.) w[0] = 0.;
    if (w[1] < 0.) w[1] = 0.;
    if (w[1] > 1000.) w[1] = 1000.;
    if (w[2] < 0.) w[2] = 0.;
    if (w[4] < 0.) w[4] = 0.;
    if (w[6] < 0.) w[6] = 0.;
    if (w[7] < 0.) w[7] = 0.;
    if (w[8] < 0.) w[8] = 0.;
    // snow in elevation bands (:440-471), in the file's units (band * landAreaFrac / contfreq, integrateWGHM.cpp:830-836)
    const double laf = laf_of(a, i), contf = a.contfreq[x];
    const double snow_mean_before = mon[1];
    const double snow_after = fl[1] + mf[1];
    double *__restrict__ S = a.snow_bands + bi(p, m, x, WGK_NBAND_K);
    const size_t bs = band_stride(p);
    for (int e = 1; e < WGK_NBAND_K; e++) {
        double sie = (laf == 0.) ? 0. : S[(size_t)e * bs] * laf / contf;
        if (snow_mean_before == 0) sie = snow_after / 100;
        else sie *= snow_after / snow_mean_before;
        if (sie < 0.) sie = 0.;
        if (sie > 1000.) sie = 1000.;
        S[(size_t)e * bs] = (laf <= 0.) ? 0. : sie * contf / laf;  // daily.cpp:1904-1922
    }
    if (laf <= 0.) {
        a.canopy[i] = 0.;
        a.snow[i] = 0.;
        a.soil[i] = 0.;
    } else {
        a.canopy[i] = w[0] * contf / laf;
        a.snow[i] = w[1] * contf / laf;
        a.soil[i] = w[2] * contf / laf;
    }
    const double denom = ((a.area[x] * (contf / C100)) / C1E6);
    a.loc_lake_stor[i] = w[3] * denom;
    a.loc_wetl_stor[i] = w[4] * denom;
    a.glo_lake_stor[i] = w[5] * denom;
    a.glo_wetl_stor[i] = w[6] * denom;
    a.res_stor[i] = w[7] * denom;
    a.river_stor[i] = w[8] * denom;
    a.gw[i] = w[9] * denom;
}

#endif  // !WGK_WU

}  // namespace wgk / wgk_mm
