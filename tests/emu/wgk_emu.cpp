// wgk_emu.cpp — TEST INFRASTRUCTURE: runs the CUDA kernel source of the product on the host
// (see cuda_shim.h).  Device-order bookkeeping is re-derived here independently of
// wgk_api.cu; only the kernel header is shared with the product.
#include "cuda_shim.h"
#define cuda_runtime_h_skip
namespace wgk { constexpr int NBAND = 101; }
#define WGK_NBAND_K wgk::NBAND
#define WGK_EMU 1
#include "../../watergap2_b200/csrc/wgk_kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct Field { const char *name; const char *dtype; int scope, bands, elsize; size_t offset; };
static const Field kFields[] = {
#define X(name, ctype, dt, scope, bands) {#name, dt, WGK_SCOPE_##scope, bands, (int)sizeof(ctype), offsetof(WgkArrays, name)},
    WGK_FIELDS(X)
#undef X
};

struct Emu {
    int ncell, stride, nlevels;
    WgkParams p{};
    std::vector<int32_t> rank_of_cell, cell_of_rank, up_off, up_idx, down, level_off, member_pset;
    std::vector<float4> forcing;
    std::vector<int32_t> gidx;
    std::vector<double> gbody, qbuf;
    int32_t cal_days[8] = {0};
    int32_t cal[8] = {0};
    bool derived = false;
    int form = 0;  // 0: band-parallel tiles (5 warps), 1: thread per cell, 2: band-parallel tiles (2 warps),
                   // 3: the fused (day, level) task k_level_day (thread per cell; vertical + local routing + river + post)
};

template <class K, class... A> static void launch(K kern, dim3 grid, dim3 block, A... args) {
    gridDim = grid; blockDim = block;
    for (unsigned by = 0; by < grid.y; by++)
        for (unsigned bx = 0; bx < grid.x; bx++)
            for (unsigned tx = 0; tx < block.x; tx++) { blockIdx = dim3(bx, by); threadIdx = dim3(tx); kern(args...); }
}

extern "C" {

Emu *emu_create(int ncell, const int32_t *rout_order, const int32_t *downstream) {
    Emu *e = new Emu();
    e->ncell = ncell;
    e->stride = (ncell + 31) / 32 * 32;
    for (const Field &f : kFields) {
        size_t n = f.scope == WGK_SCOPE_TABLE ? 18 : (size_t)e->stride * f.bands;
        *(void **)((char *)&e->p.a + f.offset) = calloc(n, f.elsize);
    }
    e->rank_of_cell.assign(ncell, 0); e->cell_of_rank.assign(ncell, 0);
    for (int n = 0; n < ncell; n++) { e->rank_of_cell[n] = rout_order[n] - 1; e->cell_of_rank[rout_order[n] - 1] = n; }
    e->down.assign(ncell, -1);
    std::vector<int> nup(ncell, 0), lvl(ncell, 0);
    for (int n = 0; n < ncell; n++) if (downstream[n] > 0) { int rd = e->rank_of_cell[downstream[n] - 1]; e->down[e->rank_of_cell[n]] = rd; nup[rd]++; }
    e->up_off.assign(ncell + 1, 0);
    for (int r = 0; r < ncell; r++) e->up_off[r + 1] = e->up_off[r] + nup[r];
    e->up_idx.assign(std::max(1, e->up_off[ncell]), 0);
    std::vector<int> fill(ncell, 0);
    for (int r = 0; r < ncell; r++) if (e->down[r] >= 0) e->up_idx[e->up_off[e->down[r]] + fill[e->down[r]]++] = r;
    for (int r = 0; r < ncell; r++) if (e->down[r] >= 0) lvl[e->down[r]] = std::max(lvl[e->down[r]], lvl[r] + 1);
    e->nlevels = lvl[ncell - 1] + 1;
    e->level_off.assign(e->nlevels + 1, 0);
    for (int r = 0; r < ncell; r++) e->level_off[lvl[r] + 1]++;
    for (int l = 0; l < e->nlevels; l++) e->level_off[l + 1] += e->level_off[l];
    e->member_pset.assign(1, 0);
    e->forcing.assign((size_t)31 * e->stride, float4{0, 0, 0, 0});
    WgkParams &p = e->p;
    p.member_pset = e->member_pset.data(); p.forcing = e->forcing.data(); p.up_off = e->up_off.data();
    p.up_idx = e->up_idx.data(); p.down = e->down.data(); p.level_off = e->level_off.data(); p.cal = e->cal;
    p.record = nullptr; p.record_cells = nullptr; p.nrec = 0; p.record_max_days = 0;
    p.ncell = ncell; p.stride = e->stride; p.nmember = 1; p.npset = 1; p.forcing_nslots = 31; p.forcing_per_member = 0;
    p.restart = 0; p.nlevels = e->nlevels; p.mm = 0; p.mpad = 1; p.ppad = 1;
    e->qbuf.assign((size_t)wgk::QBUF_K * e->stride, 0.0);
    p.qbuf = e->qbuf.data(); p.cal_days = e->cal_days;
    return e;
}

static const Field *find(const char *name) {
    for (const Field &f : kFields) if (!strcmp(f.name, name)) return &f;
    return nullptr;
}

int emu_set(Emu *e, const char *name, const void *host) {
    const Field *f = find(name);
    if (!f) return -1;
    e->derived = false;
    char *dst = *(char **)((char *)&e->p.a + f->offset);
    if (f->scope == WGK_SCOPE_TABLE) { memcpy(dst, host, (size_t)18 * f->elsize); return 0; }
    for (int r = 0; r < e->ncell; r++)
        for (int b = 0; b < f->bands; b++)
            memcpy(dst + ((size_t)b * e->stride + r) * f->elsize, (const char *)host + ((size_t)e->cell_of_rank[r] * f->bands + b) * f->elsize, f->elsize);
    return 0;
}

int emu_get(Emu *e, const char *name, void *host) {
    const Field *f = find(name);
    if (!f) return -1;
    const char *src = *(char **)((char *)&e->p.a + f->offset);
    if (f->scope == WGK_SCOPE_TABLE) { memcpy(host, src, (size_t)18 * f->elsize); return 0; }
    for (int r = 0; r < e->ncell; r++)
        for (int b = 0; b < f->bands; b++)
            memcpy((char *)host + ((size_t)e->cell_of_rank[r] * f->bands + b) * f->elsize, src + ((size_t)b * e->stride + r) * f->elsize, f->elsize);
    return 0;
}

void emu_set_forcing(Emu *e, const float *P, const float *T, const float *SW, const float *LW) {
    for (int d = 0; d < 31; d++)
        for (int r = 0; r < e->ncell; r++) {
            size_t s = (size_t)e->cell_of_rank[r] * 31 + d;
            e->forcing[(size_t)d * e->stride + r] = float4{P[s], T[s], SW[s], LW[s]};
        }
}

}  // extern "C"

// k_cells_pre on the host: the phases of vertical_tile() in the order of the kernel, each phase run
// for all threads of the CTA before the next one starts (= the barriers of the kernel)
template <class C>
static void emu_cells_pre_tiles(Emu *e, int begin, int end) {
    using namespace wgk;
    WgkParams &p = e->p;
    static VTile<C> sm;
    static VThread<C> ts[C::THREADS];
    const int slot = p.cal_days[3], m = 0;
    for (int tile = 0; tile < v_num_tiles(begin, end); tile++) {
        const int r0 = (begin & ~3) + 32 * tile;
        auto each = [&](auto fn) { for (int t = 0; t < C::THREADS; t++) fn(t >> 5, t & 31, ts[t]); };
        each([&](int w, int lane, VThread<C> &) { if (w == 0) v_mode<C>(p, sm, r0, begin, end, m, slot, lane); else v_preload<C>(p, sm, r0, begin, end, m, slot, w, lane); });
        if (sm.nband_cells == 0) {
            for (int lane = 0; lane < 32; lane++) v_head<C>(p, sm, r0, m, slot, lane);
            for (int lane = 0; lane < 32; lane++) v_bare_sums<C>(p, sm, r0, lane);
        } else {
            each([&](int w, int lane, VThread<C> &t) { v_prefetch<C>(p, sm, t, r0, m, 0, w, lane); });
            for (int lane = 0; lane < 32; lane++) v_head<C>(p, sm, r0, m, slot, lane);
            for (int slab = 0; slab < C::NSLAB; slab++) {
                each([&](int w, int lane, VThread<C> &t) { v_scale<C>(p, sm, t, r0, m, slab, w, lane); });
                if (sm.cap_any[slab]) each([&](int w, int lane, VThread<C> &t) { v_cap_resolve<C>(sm, t, slab, w, lane); });
                each([&](int w, int lane, VThread<C> &t) { v_band<C>(p, sm, t, r0, m, slab, w, lane); });
                each([&](int w, int lane, VThread<C> &t) { if (w < 4) v_sum<C>(sm, t, slab, w, lane); });
            }
        }
        for (int lane = 0; lane < 32; lane++) v_finish<C, true>(p, sm, r0, begin, end, m, lane, p.cal_days[1]);
    }
}

extern "C" {

static void emu_cells_pre(Emu *e, int begin, int end) {
    using namespace wgk;
    if (e->form == 1) {  // thread-per-cell form
        launch(k_cells_pre_tpc, dim3((end - begin + VBLOCK - 1) / VBLOCK, 1), dim3(VBLOCK), e->p, 0, begin, end);
    } else if (e->form == 2) {
        emu_cells_pre_tiles<VCfgMid>(e, begin, end);
    } else {
        emu_cells_pre_tiles<VCfgSmall>(e, begin, end);
    }
}

static void emu_river(Emu *e, int l, int day, int month) {
    WgkParams &p = e->p;
    double *qday = wgk::qbuf_of_day(p, 0);
    for (int r = e->level_off[l]; r < e->level_off[l + 1]; r++) {
        const wgk::RiverCtx c = wgk::load_ctx(p, r, (size_t)r, (size_t)r);
        if (c.flags & wgk::FL_ACTIVE) wgk::route_river(p, c, r, 0, (size_t)r, (size_t)r, wgk::gather_upstream(p, c, 0, qday), day, month, qday);
    }
}

// tail_level0 < 0: every level as its own (V, R) task pair; else the levels >= tail_level0 as one
// k_cells_pre over all their cells followed by the level-by-level river sweep and the post-pass
void emu_day(Emu *e, int day, int month, int dom, int slot, int tail_level0) {
    e->cal_days[0] = day; e->cal_days[1] = month; e->cal_days[2] = dom; e->cal_days[3] = slot;
    WgkParams &p = e->p;
    dim3 block(128), grid((e->ncell + 127) / 128, 1);
    if (!e->derived) {
        launch(wgk::k_derive_static, grid, block, p);
        launch(wgk::k_derive_member, grid, block, p);
        e->derived = true;
    }
    if (e->gidx.empty()) {
        e->gidx.assign(e->ncell, -1);
        int n = 0;
        for (int r = 0; r < e->ncell; r++)
            if (e->p.a.s_flags[r] & (wgk::FL_LAKE | wgk::FL_RES | wgk::FL_GLOWET)) e->gidx[r] = n++;
        e->gbody.assign((size_t)std::max(1, n) * wgk::GB_N, 0.0);
        e->p.gidx = e->gidx.data(); e->p.gbody = e->gbody.data(); e->p.ngbody = n;
    }
    if (e->form == 3) {  // one fused task per level, in level order (what the wavefront graph runs for small problems)
        for (int l = 0; l < e->nlevels; l++) {
            const int n = e->level_off[l + 1] - e->level_off[l];
            launch(wgk::k_level_day, dim3((n + wgk::VBLOCK - 1) / wgk::VBLOCK, 1), dim3(wgk::VBLOCK), p, 0, l);
        }
        memcpy(p.a.discharge, wgk::qbuf_of_day(p, 0), sizeof(double) * e->stride);
        return;
    }
    int t0 = tail_level0 < 0 ? e->nlevels : tail_level0;
    for (int l = 0; l < t0; l++) {
        emu_cells_pre(e, e->level_off[l], e->level_off[l + 1]);
        emu_river(e, l, day, month);
        for (int r = e->level_off[l]; r < e->level_off[l + 1]; r++) wgk::route_post_cell(p, r, 0);
    }
    if (t0 < e->nlevels) {
        int begin = e->level_off[t0], end = e->level_off[e->nlevels];
        emu_cells_pre(e, begin, end);
        for (int l = t0; l < e->nlevels; l++) emu_river(e, l, day, month);
        for (int r = begin; r < end; r++) wgk::route_post_cell(p, r, 0);
    }
    memcpy(p.a.discharge, wgk::qbuf_of_day(p, 0), sizeof(double) * e->stride);
}

// exactness of the constant-division helper of the kernels: number of operands (out of n pseudo-random ones
// spread over 600 binades, plus neighbours of powers of two and of multiples of c) whose quotient differs
// from the IEEE division
long emu_test_constdiv(long n, unsigned long long seed) {
    const wgk::ConstDiv cs[5] = {wgk::C100, wgk::C1E6, wgk::C1000, wgk::C30, wgk::C86400};
    long bad = 0;
    unsigned long long x = seed ? seed : 88172645463325252ull;
    for (long k = 0; k < n; k++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;  // xorshift64
        unsigned long long bits = (x & 0x000fffffffffffffull) | ((unsigned long long)(1023 - 300 + (int)((x >> 52) % 600)) << 52) | (x & 0x8000000000000000ull);
        double v;
        memcpy(&v, &bits, 8);
        for (const wgk::ConstDiv &c : cs) {
            const volatile double ref = v / c.c;
            if (!((v / c) == ref)) bad++;
            // operands whose quotient is at or next to a representable number
            const double w = std::nextafter(ref * c.c, (k & 1) ? 1e308 : -1e308);
            const volatile double ref2 = w / c.c;
            if (!((w / c) == ref2)) bad++;
        }
    }
    return bad;
}

void emu_set_form(Emu *e, int form) { e->form = form; }
int emu_nlevels(Emu *e) { return e->nlevels; }
}
