// Context, memory plan and the C ABI entry points of librgbl_b200.so (see include/rgbl_b200.h).
// There is no CPU fallback anywhere in this file: every compute entry point needs the CUDA device
// the context was created on and reports RGBL_E_CUDA otherwise.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "rgbl_ctx.h"

namespace rgbl {

static thread_local std::string g_create_error;

template <class T>
static cudaError_t dmalloc(T** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(T)); }
template <class T>
static cudaError_t hmalloc(T** p, size_t n) { return cudaMallocHost((void**)p, n * sizeof(T)); }

static void release(Ctx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    void* dev[] = {c->d_levels, c->d_cells, c->d_coefs, c->d_pyr, c->d_blur, c->d_slots, c->d_counts, c->d_cell_off,
                   c->d_level_cnt, c->d_frame_total, c->d_overflow, c->d_dense, c->d_sel, c->d_n_sel, c->d_kps,
                   c->d_kps_un, c->d_desc, c->d_pts, c->d_n_pts, c->d_idx_map, c->d_raw, c->d_processed, c->d_depth,
                   c->d_uright, c->d_scratch, c->d_kps_in, c->d_n_kp_in, c->qt_scr.perm_a, c->qt_scr.perm_b, c->qt_scr.node_a,
                   c->qt_scr.node_b, c->qt_scr.scan, c->qt_scr.quad, c->d_sel_lvl, c->d_n_sel_lvl, c->d_lvl_region};
    for (void* p : dev) if (p) cudaFree(p);
    c->trk.release();
    for (cudaEvent_t& e : c->chain_tev) if (e) cudaEventDestroy(e);
    void* host[] = {c->h_chain_ovf, c->h_scalars, c->h_level_cnt, c->h_frame_total, c->h_overflow, c->h_n_sel, c->h_n_pts, c->h_dense, c->h_sel};
    for (void* p : host) if (p) cudaFreeHost(p);
    for (int i = 0; i < kNumStages; ++i) { if (c->ev_b[i]) cudaEventDestroy(c->ev_b[i]); if (c->ev_e[i]) cudaEventDestroy(c->ev_e[i]); }
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    if (c->ev_pyr) cudaEventDestroy(c->ev_pyr);
    if (c->ev_blur) cudaEventDestroy(c->ev_blur);
    if (c->d_pts_raw) cudaFree(c->d_pts_raw);
    if (c->d_png_raw) cudaFree(c->d_png_raw);
    if (c->d_png_band) cudaFree(c->d_png_band);
    if (c->d_png_status) cudaFree(c->d_png_status);
    if (c->h_png_raw) cudaFreeHost(c->h_png_raw);
    if (c->h_png_status) cudaFreeHost(c->h_png_status);
    if (c->d_strips) cudaFree(c->d_strips);
    if (c->map_arena) cudaFree(c->map_arena);
    if (c->st_trk) { cudaStreamSynchronize(c->st_trk); cudaStreamDestroy(c->st_trk); }
    if (c->ev_snap) cudaEventDestroy(c->ev_snap);
    for (int i = 0; i < 2; ++i) {
        if (c->chain_exec[i]) cudaGraphExecDestroy(c->chain_exec[i]);
        if (c->ev_chain_b[i]) cudaEventDestroy(c->ev_chain_b[i]);
        if (c->ev_chain_e[i]) cudaEventDestroy(c->ev_chain_e[i]);
        if (c->ev_chain_done[i]) cudaEventDestroy(c->ev_chain_done[i]);
    }
    for (Ctx::StageSlot& sl : c->stage) { if (sl.img) cudaFree(sl.img); if (sl.pts) cudaFree(sl.pts); if (sl.n_pts) cudaFree(sl.n_pts); }
    if (c->h_chain_f) cudaFreeHost(c->h_chain_f);
    if (c->h_chain_i) cudaFreeHost(c->h_chain_i);
    if (c->st) cudaStreamDestroy(c->st);
    if (c->st_aux) cudaStreamDestroy(c->st_aux);
    delete c;
}

static int create(const rgbl_config* cfg, Ctx** out) {
    Ctx* c = new Ctx();
    c->cfg = *cfg;
    auto fail = [&](int rc) { g_create_error = c->err; release(c); return rc; };
    if (cfg->width < 1 || cfg->height < 1 || cfg->max_batch < 1 || cfg->max_points < 0) { c->err = "invalid configuration"; return fail(RGBL_E_INVALID); }
    if (cfg->max_points >= (1 << 22) - 1) { c->err = "max_points must be < 4194303"; return fail(RGBL_E_UNSUPPORTED); }
    int rc = compute_orb_tables(cfg->orb, c->tab);
    if (rc) { c->err = "invalid ORB parameters"; return fail(rc); }
    rc = build_geometry(cfg->width, cfg->height, c->tab, c->levels, c->cells, c->coefs, c->frame_bytes, c->err);
    if (rc) return fail(rc);
    c->n_cells = (int)c->cells.size();
    // DistributeOctTree returns at most max(quota + 2, 4 * nIni) keypoints per level (the very
    // first subdivision pass is unguarded, later ones stop within +3 of the budget).
    c->cap_kp = 0;
    for (const LevelGeom& g : c->levels) {
        const int n_ini = (int)std::round(static_cast<float>(g.max_bx - g.min_bx) / (g.max_by - g.min_by));
        c->cap_kp += std::max(g.quota + 3, 4 * n_ini);
    }
    const int B = cfg->max_batch;
    const int per_frame_cand = cfg->max_candidates > 0 ? cfg->max_candidates : std::max(32768, cfg->width * cfg->height / 8);
    c->dense_cap = per_frame_cand * B;

    cudaError_t e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) { c->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e) + " (librgbl_b200 has no CPU fallback)"; return fail(RGBL_E_CUDA); }
    const int nl = c->tab.nlevels;
    const size_t WH = (size_t)cfg->width * cfg->height;
#define CUF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { c->err = std::string(#call) + ": " + cudaGetErrorString(e_); return fail(RGBL_E_CUDA); } } while (0)
    CUF(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    CUF(cudaStreamCreateWithFlags(&c->st_aux, cudaStreamNonBlocking));
    CUF(cudaEventCreate(&c->ev_t0));
    CUF(cudaEventCreate(&c->ev_t1));
    CUF(cudaEventCreateWithFlags(&c->ev_pyr, cudaEventDisableTiming));
    CUF(cudaEventCreateWithFlags(&c->ev_blur, cudaEventDisableTiming));
    for (int i = 0; i < kNumStages; ++i) { CUF(cudaEventCreate(&c->ev_b[i])); CUF(cudaEventCreate(&c->ev_e[i])); }
    CUF(dmalloc(&c->d_levels, nl));
    CUF(dmalloc(&c->d_cells, c->cells.size()));
    CUF(dmalloc(&c->d_coefs, std::max<size_t>(c->coefs.size(), 1)));
    CUF(cudaMemcpy(c->d_levels, c->levels.data(), nl * sizeof(LevelGeom), cudaMemcpyHostToDevice));
    CUF(cudaMemcpy(c->d_cells, c->cells.data(), c->cells.size() * sizeof(CellInfo), cudaMemcpyHostToDevice));
    if (!c->coefs.empty()) CUF(cudaMemcpy(c->d_coefs, c->coefs.data(), c->coefs.size() * sizeof(LinCoef), cudaMemcpyHostToDevice));
    CUF(dmalloc(&c->d_pyr, c->frame_bytes * B));
    CUF(dmalloc(&c->d_blur, c->frame_bytes * B));
    CUF(cudaMemset(c->d_pyr, 0, c->frame_bytes * B));
    CUF(cudaMemset(c->d_blur, 0, c->frame_bytes * B));
    CUF(dmalloc(&c->d_slots, (size_t)B * c->n_cells * kCellCap));
    CUF(dmalloc(&c->d_counts, (size_t)B * c->n_cells));
    CUF(dmalloc(&c->d_cell_off, (size_t)B * c->n_cells));
    CUF(dmalloc(&c->d_level_cnt, (size_t)B * RGBL_MAX_LEVELS));
    CUF(dmalloc(&c->d_frame_total, (size_t)B));
    CUF(dmalloc(&c->d_overflow, 4));
    CUF(cudaMemset(c->d_overflow, 0, 4 * sizeof(int)));
    CUF(dmalloc(&c->d_dense, (size_t)c->dense_cap));
    {   // device quad-tree: per-level survivor regions + scratch mirroring the dense candidate buffer
        std::vector<int> region(nl + 1, 0);
        bool fits = true;
        for (int l = 0; l < nl; ++l) {
            const LevelGeom& g = c->levels[l];
            const int n_ini = (int)std::round(static_cast<float>(g.max_bx - g.min_bx) / (g.max_by - g.min_by));
            region[l + 1] = region[l] + std::max(g.quota + 3, 4 * n_ini);
            if (g.quota + 3 > 1024 || 4 * n_ini > 1024 || n_ini < 1 || n_ini > 64) fits = false;
            c->qt_max_nodes = std::max(c->qt_max_nodes, std::max(g.quota + 3, 4 * n_ini));
        }
        int dev_smem = 0;
        cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
        const char* env = getenv("RGBL_HOST_QUADTREE");
        c->qt_device_ok = fits && dev_smem >= quadtree_smem_bytes();
        c->device_quadtree = c->qt_device_ok && !(env && env[0] == '1');
        CUF(dmalloc(&c->d_lvl_region, nl + 1));
        CUF(cudaMemcpy(c->d_lvl_region, region.data(), (nl + 1) * sizeof(int), cudaMemcpyHostToDevice));
        CUF(dmalloc(&c->d_sel_lvl, (size_t)B * c->cap_kp));
        CUF(dmalloc(&c->d_n_sel_lvl, (size_t)B * RGBL_MAX_LEVELS));
        CUF(dmalloc(&c->qt_scr.perm_a, (size_t)c->dense_cap)); CUF(dmalloc(&c->qt_scr.perm_b, (size_t)c->dense_cap));
        CUF(dmalloc(&c->qt_scr.node_a, (size_t)c->dense_cap)); CUF(dmalloc(&c->qt_scr.node_b, (size_t)c->dense_cap));
        CUF(dmalloc(&c->qt_scr.scan, (size_t)c->dense_cap + (size_t)B * nl + 8));
        CUF(dmalloc(&c->qt_scr.quad, (size_t)c->dense_cap));
    }
    {   // strip FAST, staged describe, dilation with the empty-tile shortcut: measured on B200 in round 2 (profiles/r02_variants.md),
        // bit-exact and faster, hence the defaults; RGBL_<NAME>=0 selects the round-1 kernel for A/B runs
        const char* envd = getenv("RGBL_DESCRIBE_STAGED");
        c->describe_staged = !(envd && envd[0] == '0');
        const char* envl = getenv("RGBL_DILATE_V2");
        c->dilate_v2 = !(envl && envl[0] == '0');
        const char* envt = getenv("RGBL_LEVEL_TMA");
        c->level_tma = !(envt && envt[0] == '0') && make_level_tensor_maps(c->d_pyr, c->frame_bytes, B, c->levels.data(), nl, &c->level_tms) == 0;
        const char* env = getenv("RGBL_FAST_STRIPS");
        if (!(env && env[0] == '0')) {
            build_fast_strips(c->cells, 8, 264, c->strips, c->strip_rows_cap, c->strip_list_cap);
            CUF(dmalloc(&c->d_strips, c->strips.size()));
            CUF(cudaMemcpy(c->d_strips, c->strips.data(), c->strips.size() * sizeof(StripInfo), cudaMemcpyHostToDevice));
            c->fast_strips = true;
        }
    }
    CUF(dmalloc(&c->d_sel, (size_t)B * c->cap_kp));
    CUF(dmalloc(&c->d_n_sel, (size_t)B));
    CUF(dmalloc(&c->d_kps, (size_t)B * c->cap_kp));
    CUF(dmalloc(&c->d_kps_un, (size_t)c->cap_kp));
    CUF(dmalloc(&c->d_kps_in, (size_t)c->cap_kp));
    CUF(dmalloc(&c->d_n_kp_in, 1));
    CUF(dmalloc(&c->d_desc, (size_t)B * c->cap_kp * 32));
    CUF(dmalloc(&c->d_depth, (size_t)B * c->cap_kp));
    CUF(dmalloc(&c->d_uright, (size_t)B * c->cap_kp));
    if (cfg->max_points > 0) {
        CUF(dmalloc(&c->d_pts, (size_t)B * 4 * cfg->max_points));
        CUF(dmalloc(&c->d_n_pts, (size_t)B));
        CUF(dmalloc(&c->d_idx_map, (size_t)B * WH));
        CUF(cudaMemset(c->d_idx_map, 0, (size_t)B * WH * sizeof(uint32_t)));
        CUF(dmalloc(&c->d_raw, (size_t)B * WH));
        CUF(dmalloc(&c->d_processed, (size_t)B * WH));
        CUF(hmalloc(&c->h_n_pts, (size_t)B));
    }
    c->scratch_bytes = (size_t)(cfg->width + 2 * kEdgeThreshold + 64) * (cfg->height + 2 * kEdgeThreshold);
    CUF(dmalloc(&c->d_scratch, c->scratch_bytes));
    CUF(hmalloc(&c->h_scalars, 16));
    CUF(hmalloc(&c->h_chain_ovf, 4));
    for (int i = 0; i < 4; ++i) c->h_chain_ovf[i] = 0;
    c->chain_timing_on = std::getenv("RGBL_CHAIN_TIMING") != nullptr;
    c->chain_graphs_on = !(std::getenv("RGBL_CHAIN_GRAPH") && std::getenv("RGBL_CHAIN_GRAPH")[0] == '0');
    c->chain_pdl_on = !(std::getenv("RGBL_CHAIN_PDL") && std::getenv("RGBL_CHAIN_PDL")[0] == '0');
    CUF(hmalloc(&c->h_level_cnt, (size_t)B * RGBL_MAX_LEVELS));
    CUF(hmalloc(&c->h_frame_total, (size_t)B));
    CUF(hmalloc(&c->h_overflow, 4));
    CUF(hmalloc(&c->h_n_sel, (size_t)B));
    CUF(hmalloc(&c->h_dense, (size_t)c->dense_cap));
    CUF(hmalloc(&c->h_sel, (size_t)B * c->cap_kp));
#undef CUF
    *out = c;
    return RGBL_OK;
}

// ---- profiling helpers -----------------------------------------------------------------------------
void stage_begin(Ctx* c, int stage, cudaStream_t st) {
    if (c->prof_on) { cudaEventRecord(c->ev_b[stage], st); c->st_used[stage] = true; }
}
void stage_end(Ctx* c, int stage, cudaStream_t st, int launches) {
    c->total_launches += launches;
    c->st_pending_launches[stage] += launches;
    if (c->prof_on) cudaEventRecord(c->ev_e[stage], st);
}
// call after both streams are idle
void prof_collect(Ctx* c) {
    for (int i = 0; i < kNumStages; ++i) {
        if (c->prof_on && c->st_used[i]) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, c->ev_b[i], c->ev_e[i]) == cudaSuccess) { c->st_ms[i] += ms; c->st_calls[i] += 1; c->st_launches[i] += c->st_pending_launches[i]; }
        }
        c->st_used[i] = false;
        c->st_pending_launches[i] = 0;
    }
}

// ---- extraction pipeline -----------------------------------------------------------------------
// Stage 0 (copy):   images -> level 0 of each frame slot (upload_images).
// Stage 1 (device): pyramid, FAST + compaction on the main stream; blur on the aux stream.
// Stage 2 (host):   quad-tree per (frame, level) on worker threads -> SelKp lists.
// Stage 3 (device): describe.  Results stay in d_kps / d_desc (copied out by the callers).
static int upload_images(Ctx* c, int n_frames, const uint8_t* const* gray, int stride, cudaStream_t st) {
    const LevelGeom& l0 = c->levels[0];
    for (int f = 0; f < n_frames; ++f)
        CU(cudaMemcpy2DAsync(c->d_pyr + (size_t)f * c->frame_bytes + l0.off, l0.pitch, gray[f], stride, c->cfg.width,
                             c->cfg.height, cudaMemcpyHostToDevice, st));
    return RGBL_OK;
}

static int run_extract(Ctx* c, int n_frames, const std::function<void()>& aux_work = nullptr) {
    const int nl = c->tab.nlevels;
    stage_begin(c, ST_PYRAMID, c->st);
    // fused TMA tile kernel: level l's launch writes blur(l) and level l+1 (the separate blur stage below is then empty)
    const bool fused_levels = c->level_tma && launch_level_tiles(c->st, c->level_tms, c->d_pyr, c->d_blur, c->frame_bytes, c->levels.data(), nl, c->d_coefs, n_frames) == 0;
    if (!fused_levels) launch_pyramid(c->st, c->d_pyr, c->frame_bytes, c->levels.data(), nl, c->d_coefs, n_frames);
    stage_end(c, ST_PYRAMID, c->st, fused_levels ? nl : nl - 1);
    stage_begin(c, ST_FAST, c->st);
    if (c->fast_strips) {
        if (launch_fast_strips(c->st, c->d_pyr, c->frame_bytes, c->d_levels, c->d_cells, c->n_cells, c->d_strips, (int)c->strips.size(),
                               c->strip_rows_cap, c->strip_list_cap, c->cfg.orb.ini_th_fast, c->cfg.orb.min_th_fast, c->d_slots,
                               c->d_counts, c->d_overflow, n_frames) != 0) {
            c->err = "strip FAST kernel needs more shared memory than this device allows"; return RGBL_E_CUDA;
        }
    } else {
        launch_fast(c->st, c->d_pyr, c->frame_bytes, c->d_levels, c->d_cells, c->n_cells, c->cfg.orb.ini_th_fast,
                    c->cfg.orb.min_th_fast, c->d_slots, c->d_counts, c->d_overflow, n_frames);
    }
    stage_end(c, ST_FAST, c->st, 1);
    stage_begin(c, ST_COMPACT, c->st);
    launch_compact(c->st, c->d_levels, nl, c->n_cells, c->d_slots, c->d_counts, c->d_cell_off, c->d_level_cnt,
                   c->d_frame_total, c->d_dense, c->dense_cap, c->d_overflow, n_frames);
    stage_end(c, ST_COMPACT, c->st, 2);
    // The aux stream starts once FAST + compaction are done, i.e. it runs the blur (and, for RGB-L frames, the
    // depth maps queued by the caller via `aux_work`) while the host is busy with the quad-tree.
    CU(cudaEventRecord(c->ev_pyr, c->st));
    CU(cudaStreamWaitEvent(c->st_aux, c->ev_pyr, 0));
    stage_begin(c, ST_BLUR, c->st_aux);
    if (!fused_levels) launch_blur(c->st_aux, c->d_pyr, c->d_blur, c->frame_bytes, c->levels.data(), nl, n_frames);
    stage_end(c, ST_BLUR, c->st_aux, fused_levels ? 0 : nl);
    if (aux_work) aux_work();
    CU(cudaEventRecord(c->ev_blur, c->st_aux));
    if (c->prof_serial) CU(cudaStreamWaitEvent(c->st, c->ev_blur, 0));      // rgbl_profile_enable(ctx, 2): no kernel of this context overlaps another
    if (c->device_quadtree) {
        // Fully on-device keypoint distribution: no host round trip between FAST and describe.
        stage_begin(c, ST_QUADTREE, c->st);
        if (launch_quadtree(c->st, c->d_dense, c->d_level_cnt, c->d_frame_total, c->d_levels, nl, c->qt_scr, c->d_sel_lvl, c->d_n_sel_lvl,
                            c->d_lvl_region, c->cap_kp, c->d_overflow + 1, c->d_sel, c->d_n_sel, n_frames, c->qt_max_nodes) != 0) {
            c->err = "quad-tree kernel needs more shared memory than this device allows"; return RGBL_E_CUDA;
        }
        stage_end(c, ST_QUADTREE, c->st, 2);
        CU(cudaStreamWaitEvent(c->st, c->ev_blur, 0));
        stage_begin(c, ST_DESCRIBE, c->st);
        (c->describe_staged ? launch_describe_staged : launch_describe)(c->st, c->d_pyr, c->d_blur, c->frame_bytes, c->d_levels, c->d_sel,
                                                                        c->d_n_sel, c->cap_kp, c->cap_kp, c->tab.umax, c->d_kps, c->d_desc,
                                                                        n_frames);
        stage_end(c, ST_DESCRIBE, c->st, 1);
        CU(cudaGetLastError());
        c->last_frames = n_frames;
        c->blur_valid = true;
        c->host_counts_valid = false;
        return c->cap_kp;        // upper bound of keypoints per frame (exact counts are in d_n_sel)
    }
    CU(cudaMemcpyAsync(c->h_level_cnt, c->d_level_cnt, (size_t)n_frames * RGBL_MAX_LEVELS * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(c->h_frame_total, c->d_frame_total, (size_t)n_frames * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(c->h_overflow, c->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    if (*c->h_overflow) {
        cudaMemsetAsync(c->d_overflow, 0, sizeof(int), c->st);
        c->err = (*c->h_overflow == 1) ? "FAST cell slot overflow (>256 survivors in one cell)" : "candidate buffer overflow: raise max_candidates";
        return RGBL_E_CAPACITY;
    }
    size_t total = 0;
    std::vector<size_t> fbase(n_frames + 1, 0);
    for (int f = 0; f < n_frames; ++f) { total += c->h_frame_total[f]; fbase[f + 1] = total; }
    if (total > (size_t)c->dense_cap) { c->err = "candidate buffer overflow: raise max_candidates"; return RGBL_E_CAPACITY; }
    if (total) CU(cudaMemcpyAsync(c->h_dense, c->d_dense, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));

    // host quad-tree: tasks = (frame, level)
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::vector<SelKp>> sel_lists((size_t)n_frames * nl);
    std::atomic<int> next{0};
    std::atomic<int> status{0};
    const int n_tasks = n_frames * nl;
    auto worker = [&]() {
        std::vector<int32_t> xys, idx;
        for (;;) {
            const int t = next.fetch_add(1);
            if (t >= n_tasks) break;
            const int f = t / nl, l = t % nl;
            size_t off = fbase[f];
            for (int k = 0; k < l; ++k) off += c->h_level_cnt[f * RGBL_MAX_LEVELS + k];
            const int n = c->h_level_cnt[f * RGBL_MAX_LEVELS + l];
            const LevelGeom& lg = c->levels[l];
            xys.resize((size_t)n * 3);
            for (int k = 0; k < n; ++k) {
                const uint32_t p = c->h_dense[off + k];
                xys[3 * k] = (int)(p & 0xfff); xys[3 * k + 1] = (int)((p >> 12) & 0xfff); xys[3 * k + 2] = (int)(p >> 24);
            }
            idx.resize((size_t)c->cap_kp);
            int m = quadtree_select(xys.data(), n, lg.min_bx, lg.max_bx, lg.min_by, lg.max_by, lg.quota, idx.data(), (int)idx.size());
            if (m < 0) { status.store(m); continue; }
            std::vector<SelKp>& out = sel_lists[t];
            out.resize(m);
            for (int k = 0; k < m; ++k) {
                const int q = idx[k];
                out[k].x = (uint16_t)(xys[3 * q] + lg.min_bx);
                out[k].y = (uint16_t)(xys[3 * q + 1] + lg.min_by);
                out[k].level = (uint8_t)l; out[k].score = (uint8_t)xys[3 * q + 2]; out[k].pad = 0;
            }
        }
    };
    int n_threads = (int)std::thread::hardware_concurrency();
    n_threads = std::max(1, std::min(std::min(n_threads, 32), n_tasks));
    if (n_threads == 1) worker();
    else {
        std::vector<std::thread> pool;
        for (int i = 0; i < n_threads; ++i) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    if (status.load()) { c->err = "quad-tree selection failed"; return status.load(); }
    int max_n = 0;
    for (int f = 0; f < n_frames; ++f) {
        int n = 0;
        for (int l = 0; l < nl; ++l) {
            const std::vector<SelKp>& sl = sel_lists[(size_t)f * nl + l];
            if (n + (int)sl.size() > c->cap_kp) { c->err = "keypoint capacity exceeded"; return RGBL_E_CAPACITY; }
            if (!sl.empty()) std::memcpy(c->h_sel + (size_t)f * c->cap_kp + n, sl.data(), sl.size() * sizeof(SelKp));
            n += (int)sl.size();
        }
        c->h_n_sel[f] = n;
        max_n = std::max(max_n, n);
    }
    c->host_quadtree_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    CU(cudaMemcpyAsync(c->d_sel, c->h_sel, (size_t)n_frames * c->cap_kp * sizeof(SelKp), cudaMemcpyHostToDevice, c->st));
    CU(cudaMemcpyAsync(c->d_n_sel, c->h_n_sel, (size_t)n_frames * sizeof(int), cudaMemcpyHostToDevice, c->st));
    CU(cudaStreamWaitEvent(c->st, c->ev_blur, 0));
    stage_begin(c, ST_DESCRIBE, c->st);
    (c->describe_staged ? launch_describe_staged : launch_describe)(c->st, c->d_pyr, c->d_blur, c->frame_bytes, c->d_levels, c->d_sel,
                                                                    c->d_n_sel, c->cap_kp, max_n, c->tab.umax, c->d_kps, c->d_desc, n_frames);
    stage_end(c, ST_DESCRIBE, c->st, max_n > 0 ? 1 : 0);
    CU(cudaGetLastError());
    c->last_frames = n_frames;
    c->blur_valid = true;
    c->host_counts_valid = true;
    return max_n;
}

// Device-quad-tree mode: bring the per-frame keypoint counts and the status flags to the host (one small sync).
static int fetch_counts(Ctx* c, int n_frames) {
    if (c->host_counts_valid) return RGBL_OK;
    CU(cudaMemcpyAsync(c->h_n_sel, c->d_n_sel, (size_t)n_frames * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(c->h_overflow, c->d_overflow, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    if (c->h_overflow[0] || c->h_overflow[1]) {
        const int a = c->h_overflow[0], b = c->h_overflow[1];
        cudaMemsetAsync(c->d_overflow, 0, 2 * sizeof(int), c->st);
        c->err = b ? "device quad-tree capacity exceeded (set RGBL_HOST_QUADTREE=1)" : (a == 1 ? "FAST cell slot overflow (>256 survivors in one cell)" : "candidate buffer overflow: raise max_candidates");
        return RGBL_E_CAPACITY;
    }
    return RGBL_OK;
}

// Depth maps for n_frames resident point clouds (d_pts / d_n_pts) on stream st.
static void run_depth_maps(Ctx* c, const DepthDev& dd, int n_frames, int max_pts, float* raw, cudaStream_t st) {
    const int W = c->cfg.width, H = c->cfg.height;
    stage_begin(c, ST_DEPTH_PROJECT, st);
    launch_depth_project(st, c->d_pts, 4 * c->cfg.max_points, c->d_n_pts, max_pts, dd, W, H, c->d_idx_map, c->stamp, n_frames);
    stage_end(c, ST_DEPTH_PROJECT, st, max_pts > 0 ? 1 : 0);
    stage_begin(c, ST_DEPTH_DILATE, st);
    const bool need_raw = dd.method == RGBL_DEPTH_AVERAGE_FILTERING || dd.method == RGBL_DEPTH_NEAREST_NEIGHBOR_PIXEL;
    float* raw_out = need_raw ? c->d_raw : raw;
    (c->dilate_v2 ? launch_depth_resolve_dilate_v2 : launch_depth_resolve_dilate)(st, c->d_pts, 4 * c->cfg.max_points, c->d_n_pts, dd, W, H, c->d_idx_map,
                                                                                  c->stamp, raw_out, c->d_processed, n_frames);
    int launches = 1;
    if (dd.method == RGBL_DEPTH_AVERAGE_FILTERING) { launch_depth_average_filter(st, c->d_raw, W, H, dd.avg_kernel, c->d_processed, n_frames); ++launches; }
    stage_end(c, ST_DEPTH_DILATE, st, launches);
}

// GetFeatureDepthFromDepthMap / the per-keypoint part of Upsample_NearestNeighbor_Pixel
static void run_depth_keypoints(Ctx* c, const DepthDev& dd, const rgbl_keypoint* kps, const rgbl_keypoint* kps_un, const int* n_kp, int max_n,
                                int n_frames, cudaStream_t st) {
    const int W = c->cfg.width, H = c->cfg.height;
    stage_begin(c, ST_DEPTH_GATHER, st);
    if (dd.method == RGBL_DEPTH_NEAREST_NEIGHBOR_PIXEL)
        launch_depth_nn_pixel(st, c->d_raw, W, H, kps, kps_un, n_kp, c->cap_kp, max_n, dd.bf, dd.nn_radius, c->d_depth, c->d_uright, n_frames);
    else
        launch_depth_gather(st, c->d_processed, W, H, kps, kps_un, n_kp, c->cap_kp, max_n, dd.method == RGBL_DEPTH_NONE ? -1.f : dd.bf,
                            c->d_depth, c->d_uright, n_frames);
    stage_end(c, ST_DEPTH_GATHER, st, max_n > 0 ? 1 : 0);
}

static int setup_depth(Ctx* c, const float P[12], const rgbl_depth_params* prm, DepthDev& dd) {
    if (!c->d_pts) { c->err = "context was created with max_points == 0"; return RGBL_E_INVALID; }
    std::memcpy(dd.P, P, sizeof(float) * 12);
    dd.min_dist = prm->min_dist; dd.max_dist = prm->max_dist; dd.bf = prm->bf;
    dd.inv_scale_m = prm->max_dist * prm->inv_dilation_scale;
    dd.method = prm->method; dd.avg_kernel = prm->avg_kernel; dd.nn_radius = prm->nn_search_radius;
    switch (prm->method) {
        case RGBL_DEPTH_INVERSE_DILATION:
            if (prm->ku < 1 || prm->kv < 1 || prm->ku > 9 || prm->kv > 9) { c->err = "structuring element must be 1..9"; return RGBL_E_INVALID; }
            dd.ku = prm->ku; dd.kv = prm->kv;
            std::memcpy(dd.mask, prm->mask, 81);
            break;
        case RGBL_DEPTH_AVERAGE_FILTERING:
            if (prm->avg_kernel < 1 || prm->avg_kernel > 9) { c->err = "AverageFiltering kernel size must be 1..9"; return RGBL_E_INVALID; }
            dd.ku = dd.kv = 1; std::memset(dd.mask, 0, 81); dd.mask[0] = 1;          // resolve only: Raw is the filter input
            break;
        case RGBL_DEPTH_NEAREST_NEIGHBOR_PIXEL:
            if (!(prm->nn_search_radius >= 1.f) || prm->nn_search_radius > 30.f) { c->err = "NearestNeighborPixel search distance must be 1..30"; return RGBL_E_INVALID; }
            dd.ku = dd.kv = 1; std::memset(dd.mask, 0, 81); dd.mask[0] = 1;
            break;
        case RGBL_DEPTH_NONE:
            dd.ku = dd.kv = 1; std::memset(dd.mask, 0, 81); dd.mask[0] = 1;
            break;
        default:
            c->err = "LiDAR.Method not implemented (IPBasic has no definition in the reference either, src/DepthModule.cc:62-77)";
            return RGBL_E_UNSUPPORTED;
    }
    if (++c->stamp >= 1023u) {
        // stamp wrap: clear the index map once every 1022 calls (both streams idle first)
        cudaStreamSynchronize(c->st); cudaStreamSynchronize(c->st_aux);
        if (cudaMemset(c->d_idx_map, 0, (size_t)c->cfg.max_batch * c->cfg.width * c->cfg.height * sizeof(uint32_t)) != cudaSuccess) {
            c->err = "cudaMemset(idx_map) failed"; return RGBL_E_CUDA;
        }
        c->stamp = 1;
    }
    return RGBL_OK;
}

}  // namespace rgbl

using namespace rgbl;

extern "C" {

int rgbl_create(const rgbl_config* cfg, rgbl_ctx** out) {
    if (!cfg || !out) return RGBL_E_INVALID;
    *out = nullptr;
    Ctx* c = nullptr;
    int rc = create(cfg, &c);
    if (rc) return rc;
    *out = reinterpret_cast<rgbl_ctx*>(c);
    return RGBL_OK;
}

void rgbl_destroy(rgbl_ctx* ctx) { release(reinterpret_cast<Ctx*>(ctx)); }

int rgbl_keypoint_capacity(const rgbl_ctx* ctx) { return ctx ? reinterpret_cast<const Ctx*>(ctx)->cap_kp : RGBL_E_INVALID; }

const char* rgbl_last_error(const rgbl_ctx* ctx) {
    if (!ctx) return g_create_error.c_str();
    return reinterpret_cast<const Ctx*>(ctx)->err.c_str();
}

int rgbl_orb_extract_batch(rgbl_ctx* ctx, int n_frames, const uint8_t* const* gray, int width, int height, int stride,
                           int lap0, int lap1, rgbl_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_index) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    if (!gray || n_frames < 1 || !kps || !desc || !n_out) { c->err = "null argument"; return RGBL_E_INVALID; }
    if (width <= 0 || height <= 0) { c->err = "empty image"; return RGBL_E_EMPTY; }
    for (int f = 0; f < n_frames; ++f) if (!gray[f]) { c->err = "empty image"; return RGBL_E_EMPTY; }
    if (width != c->cfg.width || height != c->cfg.height || stride < width) { c->err = "image size does not match the context"; return RGBL_E_INVALID; }
    if (n_frames > c->cfg.max_batch) { c->err = "n_frames exceeds max_batch"; return RGBL_E_CAPACITY; }
    CU(cudaSetDevice(c->cfg.device));
    int rc = upload_images(c, n_frames, gray, stride, c->st);
    if (rc) return rc;
    rc = run_extract(c, n_frames);
    if (rc < 0) return rc;
    rc = fetch_counts(c, n_frames); if (rc) return rc;
    for (int f = 0; f < n_frames; ++f) {
        const int n = c->h_n_sel[f];
        n_out[f] = n;
        if (n > cap) { c->err = "output capacity too small"; return RGBL_E_CAPACITY; }
        if (n) {
            CU(cudaMemcpyAsync(kps + (size_t)f * cap, c->d_kps + (size_t)f * c->cap_kp, (size_t)n * sizeof(rgbl_keypoint), cudaMemcpyDeviceToHost, c->st));
            CU(cudaMemcpyAsync(desc + (size_t)f * cap * 32, c->d_desc + (size_t)f * c->cap_kp * 32, (size_t)n * 32, cudaMemcpyDeviceToHost, c->st));
        }
    }
    CU(cudaStreamSynchronize(c->st));
    CU(cudaStreamSynchronize(c->st_aux));
    prof_collect(c);
    // vLappingArea placement (src/ORBextractor.cc:1153-1162): lapping keypoints fill the back.
    for (int f = 0; f < n_frames; ++f) {
        const int n = n_out[f];
        int mono = n;
        if (!(lap0 == 0 && lap1 == 0)) {
            rgbl_keypoint* k = kps + (size_t)f * cap;
            uint8_t* d = desc + (size_t)f * cap * 32;
            std::vector<rgbl_keypoint> k2(n);
            std::vector<uint8_t> d2((size_t)n * 32);
            int mi = 0, si = n - 1;
            for (int i = 0; i < n; ++i) {
                const bool lapping = k[i].x >= (float)lap0 && k[i].x <= (float)lap1;
                const int slot = lapping ? si-- : mi++;
                k2[slot] = k[i];
                std::memcpy(&d2[(size_t)slot * 32], d + (size_t)i * 32, 32);
            }
            std::memcpy(k, k2.data(), (size_t)n * sizeof(rgbl_keypoint));
            std::memcpy(d, d2.data(), (size_t)n * 32);
            mono = mi;
        }
        if (mono_index) mono_index[f] = mono;
    }
    return RGBL_OK;
}

int rgbl_orb_extract(rgbl_ctx* ctx, const uint8_t* gray, int width, int height, int stride, int lap0, int lap1,
                     rgbl_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_index) {
    const uint8_t* g[1] = {gray};
    return rgbl_orb_extract_batch(ctx, 1, g, width, height, stride, lap0, lap1, kps, desc, cap, n_out, mono_index);
}

static int check_frame_level(Ctx* c, int frame, int level) {
    if (frame < 0 || frame >= c->last_frames || level < 0 || level >= c->tab.nlevels) { c->err = "frame/level out of range (no extraction yet?)"; return RGBL_E_INVALID; }
    return RGBL_OK;
}

int rgbl_orb_get_pyramid(rgbl_ctx* ctx, int frame, int level, uint8_t* dst, int dst_stride, int* w_out, int* h_out) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !dst) return RGBL_E_INVALID;
    int rc = check_frame_level(c, frame, level); if (rc) return rc;
    const LevelGeom& lg = c->levels[level];
    const int W = lg.w + 2 * kEdgeThreshold, H = lg.h + 2 * kEdgeThreshold;
    if (dst_stride < W) { c->err = "dst_stride too small"; return RGBL_E_INVALID; }
    CU(cudaSetDevice(c->cfg.device));
    const int pitch = (W + 63) & ~63;
    launch_padded_level(c->st, c->d_pyr, c->frame_bytes, frame, lg, c->d_scratch, pitch);
    CU(cudaMemcpy2DAsync(dst, dst_stride, c->d_scratch, pitch, W, H, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    if (w_out) *w_out = lg.w;
    if (h_out) *h_out = lg.h;
    return RGBL_OK;
}

static int get_plane(Ctx* c, const uint8_t* base, int frame, int level, uint8_t* dst, int dst_stride, int* w_out, int* h_out) {
    int rc = check_frame_level(c, frame, level); if (rc) return rc;
    const LevelGeom& lg = c->levels[level];
    if (dst_stride < lg.w) { c->err = "dst_stride too small"; return RGBL_E_INVALID; }
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaMemcpy2DAsync(dst, dst_stride, base + (size_t)frame * c->frame_bytes + lg.off, lg.pitch, lg.w, lg.h, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    if (w_out) *w_out = lg.w;
    if (h_out) *h_out = lg.h;
    return RGBL_OK;
}

int rgbl_orb_get_level(rgbl_ctx* ctx, int frame, int level, uint8_t* dst, int dst_stride, int* w_out, int* h_out) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !dst) return RGBL_E_INVALID;
    return get_plane(c, c->d_pyr, frame, level, dst, dst_stride, w_out, h_out);
}

int rgbl_orb_get_blurred_level(rgbl_ctx* ctx, int frame, int level, uint8_t* dst, int dst_stride) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !dst) return RGBL_E_INVALID;
    if (!c->blur_valid) { c->err = "no extraction yet"; return RGBL_E_INVALID; }
    CU(cudaStreamSynchronize(c->st_aux));
    return get_plane(c, c->d_blur, frame, level, dst, dst_stride, nullptr, nullptr);
}

int rgbl_orb_get_candidates(rgbl_ctx* ctx, int frame, int level, int32_t* xys, int cap, int* n_out) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !xys || !n_out) return RGBL_E_INVALID;
    int rc = check_frame_level(c, frame, level); if (rc) return rc;
    if (!c->host_counts_valid) {
        CU(cudaSetDevice(c->cfg.device));
        CU(cudaMemcpyAsync(c->h_level_cnt, c->d_level_cnt, (size_t)c->last_frames * RGBL_MAX_LEVELS * sizeof(int), cudaMemcpyDeviceToHost, c->st));
        CU(cudaMemcpyAsync(c->h_frame_total, c->d_frame_total, (size_t)c->last_frames * sizeof(int), cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        size_t tot = 0;
        for (int f = 0; f < c->last_frames; ++f) tot += c->h_frame_total[f];
        if (tot > (size_t)c->dense_cap) { c->err = "candidate buffer overflow"; return RGBL_E_CAPACITY; }
        if (tot) CU(cudaMemcpyAsync(c->h_dense, c->d_dense, tot * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
    }
    size_t off = 0;
    for (int f = 0; f < frame; ++f) off += c->h_frame_total[f];
    for (int l = 0; l < level; ++l) off += c->h_level_cnt[frame * RGBL_MAX_LEVELS + l];
    const int n = c->h_level_cnt[frame * RGBL_MAX_LEVELS + level];
    *n_out = n;
    if (n > cap) { c->err = "capacity too small"; return RGBL_E_CAPACITY; }
    for (int k = 0; k < n; ++k) {
        const uint32_t p = c->h_dense[off + k];
        xys[3 * k] = (int)(p & 0xfff); xys[3 * k + 1] = (int)((p >> 12) & 0xfff); xys[3 * k + 2] = (int)(p >> 24);
    }
    return RGBL_OK;
}

int rgbl_depth_from_pcd(rgbl_ctx* ctx, const float* pts4xn, int n_pts, const float P[12], int width, int height,
                        const rgbl_depth_params* prm, const rgbl_keypoint* kps, const rgbl_keypoint* kps_un, int n_kp,
                        float* depth, float* uright, float* raw_map, float* processed_map) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    if (!pts4xn || !P || !prm || n_pts < 0 || n_kp < 0 || (n_kp > 0 && (!kps || !kps_un || !depth || !uright))) { c->err = "null argument"; return RGBL_E_INVALID; }
    if (width != c->cfg.width || height != c->cfg.height) { c->err = "image size does not match the context"; return RGBL_E_INVALID; }
    if (n_pts > c->cfg.max_points) { c->err = "n_pts exceeds max_points"; return RGBL_E_CAPACITY; }
    if (n_kp > c->cap_kp) { c->err = "n_kp exceeds keypoint capacity"; return RGBL_E_CAPACITY; }
    CU(cudaSetDevice(c->cfg.device));
    DepthDev dd;
    int rc = setup_depth(c, P, prm, dd); if (rc) return rc;
    const size_t WH = (size_t)width * height;
    c->h_n_pts[0] = n_pts;
    int* h_nkp = c->h_overflow;          // 1-int pinned scratch (overflow flag is re-read on every extract)
    *h_nkp = n_kp;
    if (n_pts) CU(cudaMemcpyAsync(c->d_pts, pts4xn, (size_t)4 * n_pts * sizeof(float), cudaMemcpyHostToDevice, c->st));
    CU(cudaMemcpyAsync(c->d_n_pts, c->h_n_pts, sizeof(int), cudaMemcpyHostToDevice, c->st));
    CU(cudaMemcpyAsync(c->d_n_kp_in, h_nkp, sizeof(int), cudaMemcpyHostToDevice, c->st));
    if (n_kp) {
        CU(cudaMemcpyAsync(c->d_kps_in, kps, (size_t)n_kp * sizeof(rgbl_keypoint), cudaMemcpyHostToDevice, c->st));
        CU(cudaMemcpyAsync(c->d_kps_un, kps_un, (size_t)n_kp * sizeof(rgbl_keypoint), cudaMemcpyHostToDevice, c->st));
    }
    run_depth_maps(c, dd, 1, n_pts, c->d_raw, c->st);
    run_depth_keypoints(c, dd, c->d_kps_in, c->d_kps_un, c->d_n_kp_in, n_kp, 1, c->st);
    CU(cudaGetLastError());
    if (n_kp) {
        CU(cudaMemcpyAsync(depth, c->d_depth, (size_t)n_kp * sizeof(float), cudaMemcpyDeviceToHost, c->st));
        CU(cudaMemcpyAsync(uright, c->d_uright, (size_t)n_kp * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    }
    if (raw_map) CU(cudaMemcpyAsync(raw_map, c->d_raw, WH * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    if (processed_map) CU(cudaMemcpyAsync(processed_map, c->d_processed, WH * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    prof_collect(c);
    return RGBL_OK;
}

/* Frame::ComputeStereoFromRGBD (src/Frame.cc:1074-1095), the depth association of System::TrackRGBD: d = imDepth(kp.y, kp.x) at the
 * DISTORTED keypoint (C-style truncation of the float coordinates), mvDepth = d and mvuRight = kpUn.x - mbf / d where d > 0, else -1.
 * The gather runs on the device: the depth image (H x W float32, already scaled by DepthMapFactor, src/Tracking.cc:1568) is uploaded
 * into the context's processed-depth plane and read by the same kernel that serves DepthModule::GetFeatureDepthFromDepthMap.        */
int rgbl_depth_from_map(rgbl_ctx* ctx, const float* depth_map, int width, int height, int stride_floats, float bf, const rgbl_keypoint* kps,
                        const rgbl_keypoint* kps_un, int n_kp, float* depth, float* uright) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    if (!depth_map || n_kp < 0 || (n_kp > 0 && (!kps || !kps_un || !depth || !uright))) { c->err = "null argument"; return RGBL_E_INVALID; }
    if (width != c->cfg.width || height != c->cfg.height || stride_floats < width) { c->err = "image size does not match the context"; return RGBL_E_INVALID; }
    if (!c->d_processed) { c->err = "context was created with max_points == 0 (no depth planes)"; return RGBL_E_INVALID; }
    if (n_kp > c->cap_kp) { c->err = "n_kp exceeds keypoint capacity"; return RGBL_E_CAPACITY; }
    if (n_kp == 0) return RGBL_OK;
    CU(cudaSetDevice(c->cfg.device));
    int* h_nkp = c->h_overflow;
    *h_nkp = n_kp;
    CU(cudaMemcpy2DAsync(c->d_processed, (size_t)width * sizeof(float), depth_map, (size_t)stride_floats * sizeof(float), (size_t)width * sizeof(float), height,
                         cudaMemcpyHostToDevice, c->st));
    CU(cudaMemcpyAsync(c->d_n_kp_in, h_nkp, sizeof(int), cudaMemcpyHostToDevice, c->st));
    CU(cudaMemcpyAsync(c->d_kps_in, kps, (size_t)n_kp * sizeof(rgbl_keypoint), cudaMemcpyHostToDevice, c->st));
    CU(cudaMemcpyAsync(c->d_kps_un, kps_un, (size_t)n_kp * sizeof(rgbl_keypoint), cudaMemcpyHostToDevice, c->st));
    stage_begin(c, ST_DEPTH_GATHER, c->st);
    launch_depth_gather(c->st, c->d_processed, width, height, c->d_kps_in, c->d_kps_un, c->d_n_kp_in, c->cap_kp, n_kp, bf, c->d_depth, c->d_uright, 1);
    stage_end(c, ST_DEPTH_GATHER, c->st, 1);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(depth, c->d_depth, (size_t)n_kp * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(uright, c->d_uright, (size_t)n_kp * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    prof_collect(c);
    return RGBL_OK;
}

static int check_batch_args(Ctx* c, int n_frames, int width, int height, int stride) {
    if (width <= 0 || height <= 0) { c->err = "empty image"; return RGBL_E_EMPTY; }
    if (width != c->cfg.width || height != c->cfg.height || stride < width) { c->err = "image size does not match the context"; return RGBL_E_INVALID; }
    if (n_frames < 1) { c->err = "n_frames < 1"; return RGBL_E_INVALID; }
    if (n_frames > c->cfg.max_batch) { c->err = "n_frames exceeds max_batch"; return RGBL_E_CAPACITY; }
    return RGBL_OK;
}

// H2D of one batch of RGB-L inputs into the context's frame slots (images on the main stream, clouds on aux).
// layout 0: planar 4 x N rows (x, y, z, 1) as LoadPointcloudBinaryMat builds them; layout 1: the raw KITTI .bin records
// (x, y, z, reflectance) x N, de-interleaved on the device (the 4th row becomes 1, Examples/RGB-L/rgbl_kitti.cc:168-177)
static int upload_rgbl(Ctx* c, int n_frames, const uint8_t* const* gray, int stride, const float* const* pts4xn, const int* n_pts, int* max_pts_out,
                       int layout = 0) {
    if (!c->d_pts) { c->err = "context was created with max_points == 0"; return RGBL_E_INVALID; }
    if (layout == 1 && !c->d_pts_raw) {
        if (cudaMalloc((void**)&c->d_pts_raw, (size_t)c->cfg.max_batch * 4 * c->cfg.max_points * sizeof(float)) != cudaSuccess) {
            cudaGetLastError(); c->err = "cudaMalloc failed (raw point records)"; return RGBL_E_CUDA;
        }
    }
    int max_pts = 0;
    for (int f = 0; f < n_frames; ++f) {
        if (gray && !gray[f]) { c->err = "empty image"; return RGBL_E_EMPTY; }
        if (n_pts[f] < 0 || n_pts[f] > c->cfg.max_points || (n_pts[f] && !pts4xn[f])) { c->err = "bad point cloud"; return RGBL_E_CAPACITY; }
        max_pts = std::max(max_pts, n_pts[f]);
    }
    for (int f = 0; f < n_frames; ++f) {
        c->h_n_pts[f] = n_pts[f];
        float* dst = (layout == 1 ? c->d_pts_raw : c->d_pts) + (size_t)f * 4 * c->cfg.max_points;
        if (n_pts[f]) CU(cudaMemcpyAsync(dst, pts4xn[f], (size_t)4 * n_pts[f] * sizeof(float), cudaMemcpyHostToDevice, c->st_aux));
    }
    CU(cudaMemcpyAsync(c->d_n_pts, c->h_n_pts, (size_t)n_frames * sizeof(int), cudaMemcpyHostToDevice, c->st_aux));
    if (layout == 1 && max_pts > 0) launch_deinterleave_xyzr(c->st_aux, c->d_pts_raw, c->d_pts, 4 * c->cfg.max_points, c->d_n_pts, max_pts, n_frames);
    if (gray) { int rc = upload_images(c, n_frames, gray, stride, c->st); if (rc) return rc; }      // else: level 0 was written by decode_png_to_level0
    c->resident_frames = n_frames;
    c->resident_max_pts = max_pts;
    *max_pts_out = max_pts;
    return RGBL_OK;
}

// cv::imread(PNG, IMREAD_UNCHANGED) + cvtColor to gray (Examples/RGB-L/rgbl_kitti.cc:87, src/Tracking.cc:1567-1580) into level 0 of the
// frame slots 0..n_frames-1: host inflate into pinned staging, H2D of the filtered scanlines, reconstruction + gray on the device.
static int decode_png_to_level0(Ctx* c, int n_frames, const uint8_t* const* png, const size_t* png_bytes, int camera_rgb, cudaStream_t st) {
    const int w = c->cfg.width, h = c->cfg.height;
    if (!c->d_png_raw) {
        c->png_raw_stride = (((size_t)w * 4 + 1) * h + 255) & ~(size_t)255;
        const size_t total = c->png_raw_stride * c->cfg.max_batch;
        if (cudaMalloc((void**)&c->d_png_raw, total) != cudaSuccess || cudaMallocHost((void**)&c->h_png_raw, total) != cudaSuccess ||
            cudaMalloc((void**)&c->d_png_band, (size_t)c->cfg.max_batch * w * sizeof(uint32_t)) != cudaSuccess ||
            cudaMalloc((void**)&c->d_png_status, sizeof(int)) != cudaSuccess || cudaMallocHost((void**)&c->h_png_status, sizeof(int)) != cudaSuccess) {
            cudaGetLastError(); c->err = "allocation of the PNG staging buffers failed"; return RGBL_E_CUDA;
        }
        CU(cudaMemset(c->d_png_status, 0, sizeof(int)));
    }
    for (int f = 0; f < n_frames; ++f) if (!png[f] || !png_bytes[f]) { c->err = "empty image"; return RGBL_E_EMPTY; }
    int ch = 0;
    std::string perr;
    const int prc = png_inflate_batch(n_frames, png, png_bytes, w, h, c->h_png_raw, c->png_raw_stride, &ch, perr);
    if (prc) { c->err = perr; return prc == -2 ? RGBL_E_UNSUPPORTED : RGBL_E_INVALID; }
    const size_t used = ((size_t)w * ch + 1) * h;
    for (int f = 0; f < n_frames; ++f)
        CU(cudaMemcpyAsync(c->d_png_raw + (size_t)f * c->png_raw_stride, c->h_png_raw + (size_t)f * c->png_raw_stride, used, cudaMemcpyHostToDevice, st));
    launch_png_unfilter_gray(st, c->d_png_raw, c->png_raw_stride, w, h, ch, camera_rgb, c->d_pyr, c->frame_bytes, c->levels[0], c->d_png_band, c->d_png_status, n_frames);
    CU(cudaMemcpyAsync(c->h_png_status, c->d_png_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));          // the staging buffer is free again, and a bad scanline filter byte is an error of THIS call
    if (*c->h_png_status) {
        CU(cudaMemsetAsync(c->d_png_status, 0, sizeof(int), st));
        c->err = "corrupt PNG: scanline filter type above 4"; return RGBL_E_INVALID;
    }
    return RGBL_OK;
}

// Everything of Frame::Frame (RGB-L) after the inputs are in HBM: depth maps (aux stream) || extraction, then gather.
static int process_rgbl(Ctx* c, int n_frames, int max_pts, const float P[12], const rgbl_depth_params* prm) {
    DepthDev dd;
    int rc = setup_depth(c, P, prm, dd); if (rc) return rc;
    // depth maps run on the aux stream behind the blur, overlapping the host quad-tree; describe waits for both
    int max_n = run_extract(c, n_frames, [&]() { run_depth_maps(c, dd, n_frames, max_pts, nullptr, c->st_aux); });
    if (max_n < 0) return max_n;
    run_depth_keypoints(c, dd, c->d_kps, c->d_kps, c->d_n_sel, max_n, n_frames, c->st);
    CU(cudaGetLastError());
    return max_n;
}

static int download_rgbl(Ctx* c, int n_frames, rgbl_keypoint* kps, uint8_t* desc, float* depth, float* uright, int cap, int* n_out) {
    { int rc = fetch_counts(c, n_frames); if (rc) return rc; }
    for (int f = 0; f < n_frames; ++f) {
        const int n = c->h_n_sel[f];
        n_out[f] = n;
        if (n > cap) { c->err = "output capacity too small"; return RGBL_E_CAPACITY; }
        if (!n) continue;
        CU(cudaMemcpyAsync(kps + (size_t)f * cap, c->d_kps + (size_t)f * c->cap_kp, (size_t)n * sizeof(rgbl_keypoint), cudaMemcpyDeviceToHost, c->st));
        CU(cudaMemcpyAsync(desc + (size_t)f * cap * 32, c->d_desc + (size_t)f * c->cap_kp * 32, (size_t)n * 32, cudaMemcpyDeviceToHost, c->st));
        CU(cudaMemcpyAsync(depth + (size_t)f * cap, c->d_depth + (size_t)f * c->cap_kp, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->st));
        CU(cudaMemcpyAsync(uright + (size_t)f * cap, c->d_uright + (size_t)f * c->cap_kp, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    }
    CU(cudaStreamSynchronize(c->st));
    CU(cudaStreamSynchronize(c->st_aux));
    prof_collect(c);
    return RGBL_OK;
}

int rgbl_frame_rgbl_batch(rgbl_ctx* ctx, int n_frames, const uint8_t* const* gray, int width, int height, int stride,
                          const float* const* pts4xn, const int* n_pts, const float P[12], const rgbl_depth_params* prm,
                          rgbl_keypoint* kps, uint8_t* desc, float* depth, float* uright, int cap, int* n_out) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    if (!gray || !pts4xn || !n_pts || !P || !prm || !kps || !desc || !depth || !uright || !n_out) { c->err = "null argument"; return RGBL_E_INVALID; }
    int rc = check_batch_args(c, n_frames, width, height, stride); if (rc) return rc;
    CU(cudaSetDevice(c->cfg.device));
    int max_pts = 0;
    rc = upload_rgbl(c, n_frames, gray, stride, pts4xn, n_pts, &max_pts); if (rc) return rc;
    rc = process_rgbl(c, n_frames, max_pts, P, prm); if (rc < 0) return rc;
    return download_rgbl(c, n_frames, kps, desc, depth, uright, cap, n_out);
}

/* Resident form used to measure device throughput: inputs are uploaded once, processing can then be
 * repeated without any host->device input traffic; results stay in HBM until rgbl_resident_download. */
int rgbl_resident_upload(rgbl_ctx* ctx, int n_frames, const uint8_t* const* gray, int width, int height, int stride,
                         const float* const* pts4xn, const int* n_pts) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    if (!gray || !pts4xn || !n_pts) { c->err = "null argument"; return RGBL_E_INVALID; }
    int rc = check_batch_args(c, n_frames, width, height, stride); if (rc) return rc;
    CU(cudaSetDevice(c->cfg.device));
    int max_pts = 0;
    rc = upload_rgbl(c, n_frames, gray, stride, pts4xn, n_pts, &max_pts); if (rc) return rc;
    CU(cudaStreamSynchronize(c->st));
    CU(cudaStreamSynchronize(c->st_aux));
    return RGBL_OK;
}

int rgbl_resident_upload_kitti(rgbl_ctx* ctx, int n_frames, const uint8_t* const* gray, int width, int height, int stride,
                               const float* const* xyzr, const int* n_pts) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    if (!gray || !xyzr || !n_pts) { c->err = "null argument"; return RGBL_E_INVALID; }
    int rc = check_batch_args(c, n_frames, width, height, stride); if (rc) return rc;
    CU(cudaSetDevice(c->cfg.device));
    int max_pts = 0;
    rc = upload_rgbl(c, n_frames, gray, stride, xyzr, n_pts, &max_pts, 1); if (rc) return rc;
    CU(cudaStreamSynchronize(c->st));
    CU(cudaStreamSynchronize(c->st_aux));
    return RGBL_OK;
}

int rgbl_resident_upload_kitti_png(rgbl_ctx* ctx, int n_frames, const uint8_t* const* png, const size_t* png_bytes, int camera_rgb,
                                   const float* const* xyzr, const int* n_pts) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    if (!png || !png_bytes || !xyzr || !n_pts) { c->err = "null argument"; return RGBL_E_INVALID; }
    int rc = check_batch_args(c, n_frames, c->cfg.width, c->cfg.height, c->cfg.width); if (rc) return rc;
    CU(cudaSetDevice(c->cfg.device));
    rc = decode_png_to_level0(c, n_frames, png, png_bytes, camera_rgb, c->st); if (rc) return rc;
    int max_pts = 0;
    rc = upload_rgbl(c, n_frames, nullptr, 0, xyzr, n_pts, &max_pts, 1); if (rc) return rc;
    CU(cudaStreamSynchronize(c->st));
    CU(cudaStreamSynchronize(c->st_aux));
    return RGBL_OK;
}

int rgbl_decode_png_gray(rgbl_ctx* ctx, int n_frames, const uint8_t* const* png, const size_t* png_bytes, int camera_rgb, uint8_t* const* gray_out,
                         int stride) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    if (!png || !png_bytes || !gray_out) { c->err = "null argument"; return RGBL_E_INVALID; }
    int rc = check_batch_args(c, n_frames, c->cfg.width, c->cfg.height, stride); if (rc) return rc;
    CU(cudaSetDevice(c->cfg.device));
    rc = decode_png_to_level0(c, n_frames, png, png_bytes, camera_rgb, c->st); if (rc) return rc;
    const LevelGeom& l0 = c->levels[0];
    for (int f = 0; f < n_frames; ++f) {
        if (!gray_out[f]) { c->err = "null output image"; return RGBL_E_INVALID; }
        CU(cudaMemcpy2DAsync(gray_out[f], stride, c->d_pyr + (size_t)f * c->frame_bytes + l0.off, l0.pitch, l0.w, l0.h, cudaMemcpyDeviceToHost, c->st));
    }
    CU(cudaStreamSynchronize(c->st));
    return RGBL_OK;
}

int rgbl_resident_process(rgbl_ctx* ctx, const float P[12], const rgbl_depth_params* prm, int* n_out) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    if (!P || !prm) { c->err = "null argument"; return RGBL_E_INVALID; }
    if (c->resident_frames < 1) { c->err = "nothing uploaded"; return RGBL_E_INVALID; }
    CU(cudaSetDevice(c->cfg.device));
    int rc = process_rgbl(c, c->resident_frames, c->resident_max_pts, P, prm); if (rc < 0) return rc;
    if (n_out) { rc = fetch_counts(c, c->resident_frames); if (rc) return rc; }
    CU(cudaStreamSynchronize(c->st));
    CU(cudaStreamSynchronize(c->st_aux));
    prof_collect(c);
    if (n_out) for (int f = 0; f < c->resident_frames; ++f) n_out[f] = c->h_n_sel[f];
    return RGBL_OK;
}

/* ---- sequence runner: many consecutive batches of ONE sequence per host call ------------------------------------------------------
 * The per-batch calls above return to the caller between batches; with a slow caller (Python, several ranks on one host) the device
 * waits for the host.  rgbl_track_sequence keeps the whole loop native: per batch  inputs (host buffers, or a staged device slot) ->
 * frame construction -> tracking chain (queued two deep on the tracking stream, continue_sequence from the second batch on) -> poses
 * (and, if asked for, the frame-construction outputs) back to the host.  Nothing synchronises with the device until a chain's results
 * are collected, one batch behind.                                                                                                   */
int rgbl_resident_stage(rgbl_ctx* ctx, int slot, int n_frames, const uint8_t* const* gray, int width, int height, int stride,
                        const float* const* pts4xn, const int* n_pts) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    if (!gray || !pts4xn || !n_pts) { c->err = "null argument"; return RGBL_E_INVALID; }
    if (slot < 0 || slot >= Ctx::kMaxStageSlots) { c->err = "stage slot out of range"; return RGBL_E_INVALID; }
    if (!c->d_pts) { c->err = "context was created with max_points == 0"; return RGBL_E_INVALID; }
    int rc = check_batch_args(c, n_frames, width, height, stride); if (rc) return rc;
    CU(cudaSetDevice(c->cfg.device));
    const LevelGeom& l0 = c->levels[0];
    const size_t img_bytes = (size_t)l0.pitch * c->cfg.height, pts_floats = (size_t)4 * c->cfg.max_points;
    Ctx::StageSlot& sl = c->stage[slot];
    if (!sl.img) {
        if (cudaMalloc((void**)&sl.img, (size_t)c->cfg.max_batch * img_bytes) != cudaSuccess ||
            cudaMalloc((void**)&sl.pts, (size_t)c->cfg.max_batch * pts_floats * sizeof(float)) != cudaSuccess ||
            cudaMalloc((void**)&sl.n_pts, (size_t)c->cfg.max_batch * sizeof(int)) != cudaSuccess) {
            cudaGetLastError(); c->err = "cudaMalloc failed (stage slot)"; return RGBL_E_CUDA;
        }
    }
    int max_pts = 0;
    for (int f = 0; f < n_frames; ++f) {
        if (gray && !gray[f]) { c->err = "empty image"; return RGBL_E_EMPTY; }
        if (n_pts[f] < 0 || n_pts[f] > c->cfg.max_points || (n_pts[f] && !pts4xn[f])) { c->err = "bad point cloud"; return RGBL_E_CAPACITY; }
        max_pts = std::max(max_pts, n_pts[f]);
        CU(cudaMemcpy2DAsync(sl.img + (size_t)f * img_bytes, l0.pitch, gray[f], stride, c->cfg.width, c->cfg.height, cudaMemcpyHostToDevice, c->st));
        if (n_pts[f]) CU(cudaMemcpyAsync(sl.pts + (size_t)f * pts_floats, pts4xn[f], (size_t)4 * n_pts[f] * sizeof(float), cudaMemcpyHostToDevice, c->st));
    }
    sl.h_n_pts.assign(n_pts, n_pts + n_frames);
    CU(cudaMemcpyAsync(sl.n_pts, sl.h_n_pts.data(), (size_t)n_frames * sizeof(int), cudaMemcpyHostToDevice, c->st));
    CU(cudaStreamSynchronize(c->st));
    sl.n_frames = n_frames; sl.max_pts = max_pts;
    return RGBL_OK;
}

// staged slot -> the context's working input buffers (device-to-device, ~76 MB per 32 KITTI frames: tens of microseconds)
static int restage(Ctx* c, int slot) {
    const Ctx::StageSlot& sl = c->stage[slot];
    if (!sl.img || sl.n_frames < 1) { c->err = "stage slot is empty"; return RGBL_E_INVALID; }
    const LevelGeom& l0 = c->levels[0];
    const size_t img_bytes = (size_t)l0.pitch * c->cfg.height, pts_floats = (size_t)4 * c->cfg.max_points;
    for (int f = 0; f < sl.n_frames; ++f)
        CU(cudaMemcpyAsync(c->d_pyr + (size_t)f * c->frame_bytes + l0.off, sl.img + (size_t)f * img_bytes, img_bytes, cudaMemcpyDeviceToDevice, c->st));
    CU(cudaMemcpyAsync(c->d_pts, sl.pts, (size_t)sl.n_frames * pts_floats * sizeof(float), cudaMemcpyDeviceToDevice, c->st_aux));
    CU(cudaMemcpyAsync(c->d_n_pts, sl.n_pts, (size_t)sl.n_frames * sizeof(int), cudaMemcpyDeviceToDevice, c->st_aux));
    for (int f = 0; f < sl.n_frames; ++f) c->h_n_pts[f] = sl.h_n_pts[f];
    c->resident_frames = sl.n_frames; c->resident_max_pts = sl.max_pts;
    return RGBL_OK;
}

int rgbl_track_sequence(rgbl_ctx* ctx, const float P[12], const rgbl_depth_params* prm, const rgbl_chain_params* chain, const rgbl_sequence_io* io) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    if (!P || !prm || !chain || !io || !io->poses || !io->n_matches || !io->n_inliers) { c->err = "null argument"; return RGBL_E_INVALID; }
    const int T = io->frames_per_batch, nb = io->n_batches;
    if (T < 1 || T > c->cfg.max_batch || nb < 1) { c->err = "bad batch geometry"; return RGBL_E_INVALID; }
    if (c->chain_pending) { c->err = "a tracking chain is in flight (rgbl_resident_track_end not called)"; return RGBL_E_INVALID; }
    const bool host_in = io->gray != nullptr;
    if (host_in) {
        if (!io->pts4xn || !io->n_pts) { c->err = "null argument"; return RGBL_E_INVALID; }
        int rc = check_batch_args(c, T, io->width, io->height, io->stride); if (rc) return rc;
    } else if (io->n_slots < 1 || io->n_slots > Ctx::kMaxStageSlots) { c->err = "resident mode needs 1..8 staged slots"; return RGBL_E_INVALID; }
    const bool want_frames = io->kps != nullptr;
    if (want_frames && (!io->desc || !io->depth || !io->uright || !io->n_kp || io->cap != c->cap_kp)) {
        c->err = "frame outputs need kps, desc, depth, uright, n_kp and cap == rgbl_keypoint_capacity()"; return RGBL_E_INVALID;
    }
    CU(cudaSetDevice(c->cfg.device));
    rgbl_chain_params cp = *chain;
    int collected = 0;
    auto collect = [&]() -> int {
        const size_t o = (size_t)collected * T;
        const int rc = rgbl_resident_track_end2(ctx, io->poses + o * 7, io->n_matches + o, io->n_inliers + o,
                                                io->n_local_matches ? io->n_local_matches + o : nullptr, nullptr);
        ++collected;
        return rc;
    };
    for (int b = 0; b < nb; ++b) {
        int max_pts = 0, rc;
        if (host_in) {
            rc = upload_rgbl(c, T, io->gray + (size_t)b * T, io->stride, io->pts4xn + (size_t)b * T, io->n_pts + (size_t)b * T, &max_pts);
        } else {
            rc = restage(c, (io->first_slot + b) % io->n_slots); max_pts = c->resident_max_pts;
            if (!rc && c->resident_frames != T) { c->err = "staged slot holds a different number of frames"; rc = RGBL_E_INVALID; }
        }
        if (rc) { while (c->chain_pending) collect(); return rc; }
        rc = process_rgbl(c, T, max_pts, P, prm);
        if (rc < 0) { while (c->chain_pending) collect(); return rc; }
        if (want_frames) {        // whole [T][cap] arrays + counts, asynchronously behind the batch's kernels
            const size_t o = (size_t)b * T, n = (size_t)T * c->cap_kp;
            CU(cudaMemcpyAsync(io->kps + o * io->cap, c->d_kps, n * sizeof(rgbl_keypoint), cudaMemcpyDeviceToHost, c->st));
            CU(cudaMemcpyAsync(io->desc + o * io->cap * 32, c->d_desc, n * 32, cudaMemcpyDeviceToHost, c->st));
            CU(cudaMemcpyAsync(io->depth + o * io->cap, c->d_depth, n * sizeof(float), cudaMemcpyDeviceToHost, c->st));
            CU(cudaMemcpyAsync(io->uright + o * io->cap, c->d_uright, n * sizeof(float), cudaMemcpyDeviceToHost, c->st));
            CU(cudaMemcpyAsync(io->n_kp + o, c->d_n_sel, (size_t)T * sizeof(int), cudaMemcpyDeviceToHost, c->st));
        }
        cp.continue_sequence = (b > 0 || chain->continue_sequence) ? 1 : 0;
        rc = rgbl_resident_track_begin2(ctx, &cp);
        if (rc) { while (c->chain_pending) collect(); return rc; }
        if (c->chain_pending == 2) { rc = collect(); if (rc) { while (c->chain_pending) collect(); return rc; } }
    }
    int rc_all = RGBL_OK;
    while (c->chain_pending) { const int rc = collect(); if (rc && !rc_all) rc_all = rc; }
    CU(cudaStreamSynchronize(c->st));
    CU(cudaStreamSynchronize(c->st_aux));
    prof_collect(c);
    return rc_all;
}

int rgbl_resident_download(rgbl_ctx* ctx, rgbl_keypoint* kps, uint8_t* desc, float* depth, float* uright, int cap, int* n_out) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    if (!kps || !desc || !depth || !uright || !n_out) { c->err = "null argument"; return RGBL_E_INVALID; }
    if (c->last_frames < 1) { c->err = "nothing processed"; return RGBL_E_INVALID; }
    CU(cudaSetDevice(c->cfg.device));
    return download_rgbl(c, c->last_frames, kps, desc, depth, uright, cap, n_out);
}

int rgbl_set_host_quadtree(rgbl_ctx* ctx, int on) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    if (!on && !c->qt_device_ok) {       // the same test rgbl_create made: node capacity, root count and shared memory of every level
        c->err = "this context's level geometry does not fit the device quad-tree (quota + 3 <= 1024, 1 <= nIni <= 64 per level)"; return RGBL_E_UNSUPPORTED;
    }
    c->device_quadtree = !on;
    return RGBL_OK;
}

/* ---- device-side stopwatch on the context's main stream ---- */
int rgbl_timer_mark(rgbl_ctx* ctx, int which) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || (which != 0 && which != 1)) return RGBL_E_INVALID;
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaEventRecord(which == 0 ? c->ev_t0 : c->ev_t1, c->st));
    return RGBL_OK;
}
int rgbl_timer_elapsed_ms(rgbl_ctx* ctx, double* ms) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c || !ms) return RGBL_E_INVALID;
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaEventSynchronize(c->ev_t1));
    float f = 0.f;
    CU(cudaEventElapsedTime(&f, c->ev_t0, c->ev_t1));
    *ms = f;
    return RGBL_OK;
}

/* ---- profiling ---- */
int rgbl_profile_enable(rgbl_ctx* ctx, int on) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    c->prof_on = on != 0;
    c->prof_serial = on == 2;
    return RGBL_OK;
}
int rgbl_profile_reset(rgbl_ctx* ctx) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    for (int i = 0; i < kNumStages; ++i) { c->st_ms[i] = 0; c->st_launches[i] = 0; c->st_calls[i] = 0; }
    c->host_quadtree_ms = 0; c->total_launches = 0;
    return RGBL_OK;
}
int rgbl_profile_num_stages(void) { return kNumStages; }
const char* rgbl_profile_stage_name(int stage) { return (stage >= 0 && stage < kNumStages) ? kStageNames[stage] : ""; }
int rgbl_profile_read(const rgbl_ctx* ctx, int stage, double* total_ms, int64_t* kernel_launches, int64_t* calls) {
    const Ctx* c = reinterpret_cast<const Ctx*>(ctx);
    if (!c || stage < 0 || stage >= kNumStages) return RGBL_E_INVALID;
    if (total_ms) *total_ms = c->st_ms[stage];
    if (kernel_launches) *kernel_launches = c->st_launches[stage];
    if (calls) *calls = c->st_calls[stage];
    return RGBL_OK;
}
int rgbl_profile_totals(const rgbl_ctx* ctx, int64_t* kernel_launches, double* host_quadtree_ms) {
    const Ctx* c = reinterpret_cast<const Ctx*>(ctx);
    if (!c) return RGBL_E_INVALID;
    if (kernel_launches) *kernel_launches = c->total_launches;
    if (host_quadtree_ms) *host_quadtree_ms = c->host_quadtree_ms;
    return RGBL_OK;
}

}  // extern "C"
