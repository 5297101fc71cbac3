// C ABI of the ORB extractor: handle, per-image-size plan, CUDA graph, host staging.
// Replaces ORBextractor (corbslam_client/include/ORBextractor.h:45-112, src/ORBextractor.cc:410-470,1043-1132).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "orb_kernels.cuh"

namespace corb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

static inline int cv_round_f(float v) { return (int)nearbyintf(v); }   // cvRound: round half to even
static inline int cv_round_d(double v) { return (int)nearbyint(v); }

}  // namespace corb

using namespace corb;

struct PatchKey {  // what an import node of an instantiated graph currently points at
    const uint8_t* src = nullptr; int stride = 0; const uint8_t* src1 = nullptr; int stride1 = 0;
};

struct corb_orb {
    // parameters and tables (ORBextractor.cc:410-470)
    int nfeatures, nlevels, ini_th, min_th, device;
    float scale_factor_f;
    double scale_factor;  // the reference stores the float argument in a double member (ORBextractor.h:97)
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> quota;
    int umax[16];

    // plan for the current image size
    int plan_w = 0, plan_h = 0;
    OrbGeom geom;
    OrbBuffers buf;
    TmaMaps tma;
    size_t pyr_bytes = 0;
    int cand_total = 0;
    int key_smem_cap = 0, oct_smem = 0;
    bool h2d_node = getenv("CORB_H2D_NODE") != nullptr;  // corb_orb_set_host_transfer / CORB_H2D_NODE=1: host images enter through a copy-engine memcpy node instead of k_import's
                            // loads from mapped host memory (equal for one blocking frame, ~25 % more frames/s with 8 in flight)
    int tail_base = 0;      // chain graph: levels tail_base + 1 .. are produced by one k_pyramid launch (0 = resize launches only)
    int graph_mode = 0;  // 0: fused pyramid + per-level FAST/quadtree branches, 1: four fused launches, 2: resize chain + branches
    std::vector<void*> dev_allocs;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // two instantiations of the per-frame graph: results stay in HBM (dev) / results copied to pinned host memory (host)
    cudaGraph_t graph[2] = {nullptr, nullptr};
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
    cudaGraphNode_t import_node[2] = {nullptr, nullptr};
    PatchKey import_at[2], pair_at_l[3], pair_at_r[3];  // reset whenever the graphs are re-instantiated
    int kernel_launches = 0;

    // pinned host staging
    uint8_t* h_img = nullptr;       // plan_w * plan_h
    uint8_t* d_stage = nullptr;     // plan_w * plan_h, tightly packed landing buffer of the H2D node (owned by dev_allocs)
    uint8_t* h_pyr = nullptr;       // pyr_bytes
    corb_keypoint* h_kps = nullptr; // kp_cap
    uint8_t* h_desc = nullptr;      // kp_cap * 32
    int* h_scalars = nullptr;       // [0] count, [1] status
    uint8_t* h_out = nullptr;       // pinned mirror of the device output blob (h_kps / h_desc / h_scalars point into it)
    uint8_t* d_out = nullptr;
    size_t out_bytes = 0;
    bool pending = false, pending_pyr = false, pending_empty = false;
    int pending_pair = 0;  // on the left handle of a pair: 1 = corb_orb_extract_pair_submit, 2 = corb_frame_stereo_submit in flight
    bool own_dirty = false;  // work enqueued on the handle's own stream has not been synchronised yet

    // one graph for a stereo pair (this handle = left): [right frame || left frame] (-> stereo matching) (-> D2H)
    //   variant 0: results stay in HBM, 1: + D2H of both result blobs, 2: + ComputeStereoMatches + D2H of everything
    corb_orb* pair_peer = nullptr;
    long long uid = 0, pair_peer_uid = -1;  // handles are identified by a process-unique id, not by address
    int pair_peer_serial = -1, pair_self_serial = -1;
    float pair_mbf = 0.f, pair_mb = 0.f;
    cudaGraph_t pair_graph[3] = {nullptr, nullptr, nullptr};
    cudaGraphExec_t pair_exec[3] = {nullptr, nullptr, nullptr};
    cudaGraphNode_t pair_imp_l[3] = {nullptr, nullptr, nullptr}, pair_imp_r[3] = {nullptr, nullptr, nullptr};
    bool pair_split[3] = {false, false, false};  // the pair graph has separate launches (and import nodes) per image
    int plan_serial = 0;          // bumped whenever the plan (device buffers) is rebuilt
    bool host_count = false;      // the last extraction on this handle also delivered its keypoint count to h_scalars[0]
    cudaEvent_t ev_busy = nullptr; // recorded on the stream that last ran work touching this handle's buffers
    cudaStream_t busy_stream = nullptr;

    // stereo matching outputs (allocated on first use; this handle is the LEFT one)
    float* d_stereo = nullptr;   // [u_right kp_cap | depth kp_cap | best_dist kp_cap (int)]
    float* h_stereo = nullptr;   // pinned [u_right | depth]
    cudaEvent_t ev_peer = nullptr;
    int stereo_cap = 0;
    int* d_bands = nullptr;      // right-keypoint row bands for the stereo matcher (owned by dev_allocs)
    int stereo_bands = 0;
};

static void free_plan(corb_orb* h) {
    for (int v = 0; v < 3; v++) {
        if (h->pair_exec[v]) cudaGraphExecDestroy(h->pair_exec[v]), h->pair_exec[v] = nullptr;
        if (h->pair_graph[v]) cudaGraphDestroy(h->pair_graph[v]), h->pair_graph[v] = nullptr;
    }
    h->pair_peer = nullptr;
    h->plan_serial++;
    for (int v = 0; v < 2; v++) {
        if (h->graph_exec[v]) cudaGraphExecDestroy(h->graph_exec[v]), h->graph_exec[v] = nullptr;
        if (h->graph[v]) cudaGraphDestroy(h->graph[v]), h->graph[v] = nullptr;
    }
    for (void* p : h->dev_allocs) cudaFree(p);
    h->dev_allocs.clear();
    if (h->h_img) cudaFreeHost(h->h_img), h->h_img = nullptr;
    if (h->h_pyr) cudaFreeHost(h->h_pyr), h->h_pyr = nullptr;
    if (h->h_out) cudaFreeHost(h->h_out), h->h_out = nullptr;
    h->h_kps = nullptr; h->h_desc = nullptr; h->h_scalars = nullptr;
    if (h->h_stereo) cudaFreeHost(h->h_stereo), h->h_stereo = nullptr;
    h->d_stereo = nullptr;  // owned by dev_allocs
    h->stereo_cap = 0;
    h->plan_w = h->plan_h = 0;
}

// Geometry of one level; returns false if the reference itself cannot process this size (empty FAST grid, nIni = 0).
static bool level_geometry(const corb_orb* h, int l, int w, int hgt, LevelGeom* L) {
    memset(L, 0, sizeof(*L));
    L->level = l;
    L->w = cv_round_f((float)w * h->inv_scale[l]);  // ORBextractor.cc:1111-1112
    L->h = cv_round_f((float)hgt * h->inv_scale[l]);
    L->max_bx = L->w - kEdge + 3;
    L->max_by = L->h - kEdge + 3;
    const int width_i = L->max_bx - kBorder, height_i = L->max_by - kBorder;
    if (width_i < 30 || height_i < 30) return false;
    const float width = (float)width_i, height = (float)height_i, W = 30;  // :769-787
    L->n_cols = (int)(width / W);
    L->n_rows = (int)(height / W);
    L->w_cell = (int)ceilf(width / L->n_cols);
    L->h_cell = (int)ceilf(height / L->n_rows);
    L->slot = ((L->w_cell + 1) / 2) * ((L->h_cell + 1) / 2);
    L->quota = h->quota[l];
    L->n_ini = (int)roundf((float)width_i / height_i);  // :543
    if (L->n_ini < 1) return false;
    L->h_x = (float)width_i / L->n_ini;
    L->node_cap = L->quota + 3 > 4 * L->n_ini ? L->quota + 3 : 4 * L->n_ini;
    L->scale = h->scale[l];
    L->size = (float)(int)(kPatch * h->scale[l]);  // :837
    return true;
}

static int capacity_for(const corb_orb* h, int w, int hgt) {
    int cap = 0;
    for (int l = 0; l < h->nlevels; l++) {
        LevelGeom L;
        if (!level_geometry(h, l, w, hgt, &L)) return -1;
        cap += L.node_cap;
    }
    return cap;
}

template <typename T>
static int dev_alloc(corb_orb* h, T** p, size_t n) {
    void* q = nullptr;
    CORB_CUDA(cudaMalloc(&q, n * sizeof(T) + 256));
    h->dev_allocs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return CORB_OK;
}

static int record_graph(corb_orb* h);
static int settle(corb_orb* h);

static int make_plan(corb_orb* h, int w, int hgt) {
    if (h->plan_w == w && h->plan_h == hgt) return CORB_OK;
    CORB_CUDA(cudaSetDevice(h->device));
    if (h->stream) {
        int rc0 = settle(h);  // a peer's pair graph may still be using this handle's buffers
        if (rc0 != CORB_OK) return rc0;
        CORB_CUDA(cudaStreamSynchronize(h->stream));
    }
    free_plan(h);
    OrbGeom& g = h->geom;
    memset(&g, 0, sizeof(g));
    g.n_levels = h->nlevels;
    g.ini_th = h->ini_th;
    g.min_th = h->min_th;
    size_t img_off = 0;
    int cell_base = 0, kp_base = 0, xtab = 0, ytab = 0, tiles = 0;
    long long cand_base = 0;
    for (int l = 0; l < h->nlevels; l++) {
        LevelGeom& L = g.lv[l];
        CORB_CHECK(level_geometry(h, l, w, hgt, &L), CORB_ERR_UNSUPPORTED,
                   "image %dx%d: pyramid level %d is too small for the 30 px FAST grid", w, hgt, l);
        CORB_CHECK(L.w <= 32767 && L.h <= 32767 && L.node_cap <= 65535, CORB_ERR_UNSUPPORTED, "image or feature count too large");
        L.pitch = align_up(L.w, 128);
        L.img_off = (int)img_off;
        img_off += (size_t)L.pitch * L.h;
        L.cell_base = cell_base;
        cell_base += L.n_cols * L.n_rows;
        L.cand_base = (int)cand_base;
        cand_base += (long long)L.n_cols * L.n_rows * L.slot;
        L.kp_base = kp_base;
        kp_base += L.node_cap;
        L.xtab_off = xtab;
        L.ytab_off = ytab;
        if (l > 0) { xtab += L.w; ytab += L.h; }
        L.blur_tile_base = tiles;
        L.blur_tiles_x = (L.w + kBlurTW - 1) / kBlurTW;
        tiles += L.blur_tiles_x * ((L.h + kBlurTH - 1) / kBlurTH);
    }
    CORB_CHECK(img_off < (1u << 30) && cand_base < (1 << 24), CORB_ERR_UNSUPPORTED, "image too large");
    g.n_cells = cell_base;
    g.kp_cap = kp_base;
    g.blur_tiles = tiles;
    h->pyr_bytes = img_off;
    h->cand_total = (int)cand_base;

    // resize coefficient tables, built exactly as cv::resize builds them (double -> float -> short)
    std::vector<int> xofs(xtab + 1), yofs(ytab + 1);
    std::vector<short2> alpha(xtab + 1), beta(ytab + 1);
    for (int l = 1; l < h->nlevels; l++) {
        const LevelGeom &S = g.lv[l - 1], &D = g.lv[l];
        const double scale_x = 1.0 / ((double)D.w / S.w), scale_y = 1.0 / ((double)D.h / S.h);
        for (int dx = 0; dx < D.w; dx++) {
            float fx = (float)((dx + 0.5) * scale_x - 0.5);
            int sx = (int)floorf(fx);
            fx -= sx;
            if (sx < 0) { fx = 0; sx = 0; }
            if (sx >= S.w - 1) { fx = 0; sx = S.w - 1; }
            xofs[D.xtab_off + dx] = sx;
            alpha[D.xtab_off + dx] = make_short2((short)cv_round_f((1.f - fx) * 2048), (short)cv_round_f(fx * 2048));
        }
        for (int dy = 0; dy < D.h; dy++) {
            float fy = (float)((dy + 0.5) * scale_y - 0.5);
            int sy = (int)floorf(fy);
            fy -= sy;
            yofs[D.ytab_off + dy] = sy;
            beta[D.ytab_off + dy] = make_short2((short)cv_round_f((1.f - fy) * 2048), (short)cv_round_f(fy * 2048));
        }
    }

    OrbBuffers& b = h->buf;
    memset(&b, 0, sizeof(b));
    int rc;
    int *d_xofs, *d_yofs;
    short2 *d_alpha, *d_beta;
#define A(ptr, n) if ((rc = dev_alloc(h, &(ptr), (n))) != CORB_OK) return rc
    A(b.pyr, h->pyr_bytes);
    A(b.blur, h->pyr_bytes);
    A(d_xofs, xofs.size());
    A(d_alpha, alpha.size());
    A(d_yofs, yofs.size());
    A(d_beta, beta.size());
    A(h->d_stage, (size_t)w * hgt);
    A(b.cell_count, g.n_cells);
    A(b.cand_xy, h->cand_total);
    A(b.cand_ro, h->cand_total);
    A(b.key_node, h->cand_total);
    A(b.lvl_kp, g.kp_cap);
    A(b.level_count, kMaxLevels);
    A(b.level_cand, kMaxLevels);
    A(b.level_cand_out, kMaxLevels);
    // [keypoints | descriptors | count, status] in one allocation so one D2H copy returns everything
    h->out_bytes = align_up_sz(sizeof(corb_keypoint) * (size_t)g.kp_cap, 16) + (size_t)g.kp_cap * 32 + 16;
    uint8_t* out_blob;
    A(out_blob, h->out_bytes);
    b.kps = reinterpret_cast<corb_keypoint*>(out_blob);
    b.desc = out_blob + align_up_sz(sizeof(corb_keypoint) * (size_t)g.kp_cap, 16);
    b.count = reinterpret_cast<int*>(b.desc + (size_t)g.kp_cap * 32);
    b.status = b.count + 1;
    h->d_out = out_blob;
#undef A
#define A2(ptr, n) if ((rc = dev_alloc(h, &(ptr), (n))) != CORB_OK) return rc
    b.xofs = d_xofs; b.alpha = d_alpha; b.yofs = d_yofs; b.beta = d_beta;
    CORB_CUDA(cudaMemcpy(d_xofs, xofs.data(), xofs.size() * sizeof(int), cudaMemcpyHostToDevice));
    CORB_CUDA(cudaMemcpy(d_alpha, alpha.data(), alpha.size() * sizeof(short2), cudaMemcpyHostToDevice));
    CORB_CUDA(cudaMemcpy(d_yofs, yofs.data(), yofs.size() * sizeof(int), cudaMemcpyHostToDevice));
    CORB_CUDA(cudaMemcpy(d_beta, beta.data(), beta.size() * sizeof(short2), cudaMemcpyHostToDevice));
    CORB_CUDA(cudaMemset(b.pyr, 0, h->pyr_bytes));
    CORB_CUDA(cudaMemset(b.blur, 0, h->pyr_bytes));
    CORB_CUDA(cudaMemset(b.status, 0, sizeof(int)));
    CORB_CUDA(cudaMemset(b.count, 0, sizeof(int)));
    CORB_CUDA(cudaMemset(b.level_count, 0, kMaxLevels * sizeof(int)));
    CORB_CUDA(cudaMemset(b.level_cand, 0, kMaxLevels * sizeof(int)));
    CORB_CUDA(cudaMemset(b.level_cand_out, 0, kMaxLevels * sizeof(int)));
    {   // quadtree path tables (k_octtree's closed-form path)
        std::vector<uint16_t> lut;
        for (int l = 0; l < h->nlevels; l++) {
            g.lv[l].lut_off = (int)lut.size();
            lut.resize(lut.size() + oct_lut_entries_host(g.lv[l]) + 2);
            build_oct_lut(g.lv[l], lut.data() + g.lv[l].lut_off);
        }
        uint16_t* d_lut;
        A2(d_lut, lut.size());
        CORB_CUDA(cudaMemcpy(d_lut, lut.data(), lut.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        b.oct_lut = d_lut;
    }

    memset(&h->tma, 0, sizeof(h->tma));
    b.tma_maps = &h->tma;
    b.use_tma = !getenv("CORB_NO_TMA") && encode_tma_maps(g, b.pyr, &h->tma) ? 1 : 0;
    b.oct_fast = getenv("CORB_OCT_GENERIC") ? 0 : 1;
    {
        CUtensorMap* d_maps;
        A2(d_maps, 2 * kMaxLevels);  // m[] then mb[]
        CORB_CUDA(cudaMemcpy(d_maps, &h->tma, sizeof(h->tma), cudaMemcpyHostToDevice));
        b.tma_dev = d_maps;
        std::vector<int> ptab;
        build_pyr_plan(g, xofs.data(), yofs.data(), 0, &b.pyr_plan, &ptab);
        int* d_ptab;
        A2(d_ptab, ptab.size());
        CORB_CUDA(cudaMemcpy(d_ptab, ptab.data(), ptab.size() * sizeof(int), cudaMemcpyHostToDevice));
        b.pyr_plan.tab = d_ptab;
        cudaError_t pe = prepare_pyramid(b.pyr_plan);
        CORB_CHECK(pe == cudaSuccess, CORB_ERR_CUDA, "pyramid kernel shared memory: %s", cudaGetErrorString(pe));
        // the short end of the resize chain (small levels: a launch each costs more than the pixels) in one launch
        const char* te = getenv("CORB_PYR_TAIL");
        h->tail_base = te ? atoi(te) : 0;  // measured: every tail variant loses to the plain chain (131.7 vs 134-147 us)
        if (h->tail_base < 1 || h->tail_base >= h->nlevels - 1) h->tail_base = 0;  // 0 = plain chain
        if (h->tail_base) {
            std::vector<int> ttab;
            build_pyr_plan(g, xofs.data(), yofs.data(), h->tail_base, &b.pyr_tail, &ttab);
            int* d_ttab;
            A2(d_ttab, ttab.size());
            CORB_CUDA(cudaMemcpy(d_ttab, ttab.data(), ttab.size() * sizeof(int), cudaMemcpyHostToDevice));
            b.pyr_tail.tab = d_ttab;
            pe = prepare_pyramid(b.pyr_tail);
            CORB_CHECK(pe == cudaSuccess, CORB_ERR_CUDA, "pyramid kernel shared memory: %s", cudaGetErrorString(pe));
        }
    }
    {   // CORB_GRAPH=fused|hybrid selects the alternative per-frame graph shapes (kept for A/B measurements)
        const char* gm = getenv("CORB_GRAPH");
        h->graph_mode = gm && !strcmp(gm, "fused") ? 1 : gm && !strcmp(gm, "hybrid") ? 0 : 2;
    }
    CORB_CUDA(cudaMallocHost(&h->h_img, (size_t)w * hgt));
    CORB_CUDA(cudaMallocHost(&h->h_pyr, h->pyr_bytes));
    CORB_CUDA(cudaMallocHost(&h->h_out, h->out_bytes));
    h->h_kps = reinterpret_cast<corb_keypoint*>(h->h_out);
    h->h_desc = h->h_out + (b.desc - h->d_out);
    h->h_scalars = reinterpret_cast<int*>(h->h_out + (reinterpret_cast<uint8_t*>(b.count) - h->d_out));

    // quadtree: keep a level's keys in shared memory when they fit (typical: 2-3 k keys), else global scratch
    int max_cand_level = 0;
    for (int l = 0; l < h->nlevels; l++) max_cand_level = std::max(max_cand_level, g.lv[l].n_cols * g.lv[l].n_rows * g.lv[l].slot);
    h->key_smem_cap = std::min(max_cand_level, 4096);
    cudaError_t e = prepare_octtree(g, h->key_smem_cap, &h->oct_smem);
    if (e != cudaSuccess || h->oct_smem > 200 * 1024) {
        set_error("quadtree kernel needs %d B of shared memory (nfeatures too large?): %s", h->oct_smem, cudaGetErrorString(e));
        return e != cudaSuccess ? CORB_ERR_CUDA : CORB_ERR_UNSUPPORTED;
    }
    h->plan_w = w;
    h->plan_h = hgt;
    return record_graph(h);
}

// One CUDA graph per plan. Levels are independent once they exist, so the graph is a set of per-level pipelines
//   level 0:            FAST_0 -> quadtree_0
//   level l: resize_l -> FAST_l -> quadtree_l          (resize_l also feeds resize_{l+1})
// joined by orientation + BRIEF; the Gaussian blur (needed only by BRIEF) runs behind the resize chain. The longest
// pipeline (level 0: 36 % of the cells, the largest quadtree) therefore starts at t = 0 instead of after the chain.
// Enqueues the work of one frame of handle `h` into the capture that is active on `stream`. `ls`/`ev` are scratch
// streams/events (n_levels and 2*n_levels of them) that only shape the captured dependency graph.
// With `peer` (the right handle of a stereo pair, same geometry) every launch processes both images (grid z = 2).
static void capture_frame(corb_orb* h, cudaStream_t stream, bool d2h, std::vector<cudaStream_t>& ls, std::vector<cudaEvent_t>& ev,
                          corb_orb* peer = nullptr, cudaEvent_t ev_after_import = nullptr) {
    const OrbGeom& g = h->geom;
    const OrbBuffers& b = h->buf;
    const OrbBuffers* b1 = peer ? &peer->buf : nullptr;
    const int L = g.n_levels;
    // placeholder sources, patched before every launch. Host images (the variants that also copy the results back)
    // come in through the copy engine: a 2D memcpy node per image, which sustains more of the PCIe bandwidth than SM
    // loads from mapped host memory; device images go through k_import.
    if (d2h && h->h2d_node) {
        // (only 1-D memcpy nodes can be re-pointed in an instantiated graph, so the image lands tightly packed in a
        // staging buffer and k_import re-pitches it from there)
        const LevelGeom& L0 = g.lv[0];
        const size_t bytes = (size_t)L0.w * L0.h;
        cudaMemcpyAsync(h->d_stage, h->h_img, bytes, cudaMemcpyHostToDevice, stream);
        if (peer) cudaMemcpyAsync(peer->d_stage, peer->h_img, bytes, cudaMemcpyHostToDevice, stream);
        launch_import(g, b, h->d_stage, L0.w, stream, b1, peer ? peer->d_stage : nullptr, L0.w);
    } else {
        // host variants read mapped page-locked memory: the flat 128-bit kernel; device images: the row kernel
        launch_import(g, b, h->h_img, g.lv[0].w, stream, b1, peer ? peer->h_img : nullptr, g.lv[0].w, d2h);
    }
    if (ev_after_import) cudaEventRecord(ev_after_import, stream);
    if (h->graph_mode == 0) {
        // import -> pyramid (all levels, one launch) -> per level: FAST_l -> quadtree_l -> join -> orient + BRIEF
        //                                           \-> blur (queued behind the FAST launches) ------/
        launch_pyramid(g, b, stream, b1);
        cudaEventRecord(ev[0], stream);
        for (int l = 0; l < L; l++) {
            cudaStream_t sl = l == 0 ? stream : ls[l];
            if (l > 0) cudaStreamWaitEvent(sl, ev[0], 0);
            launch_fast_cells(g, b, l, sl, b1);
            if (l == 0) cudaEventRecord(ev[2], sl);
            launch_octtree(g, b, l, h->key_smem_cap, h->oct_smem, sl, b1);
            if (l > 0) cudaEventRecord(ev[L + l], sl);
        }
        cudaStreamWaitEvent(ls[L], ev[2], 0);  // the blur (low priority) starts once FAST on level 0 is through
        launch_blur(g, b, ls[L], b1);
        cudaEventRecord(ev[1], ls[L]);
        cudaStreamWaitEvent(stream, ev[1], 0);
        for (int l = 1; l < L; l++) cudaStreamWaitEvent(stream, ev[L + l], 0);
        launch_orient_desc(g, b, stream, b1);
        if (d2h) {
            cudaMemcpyAsync(h->h_out, h->d_out, h->out_bytes, cudaMemcpyDeviceToHost, stream);
            if (peer) cudaMemcpyAsync(peer->h_out, peer->d_out, peer->out_bytes, cudaMemcpyDeviceToHost, stream);
        }
        return;
    }
    if (h->graph_mode == 1) {
        // import -> pyramid (all levels, one launch) -> FAST (all cells of all levels) -> quadtree (one CTA per level)
        //                                 \-> blur ----------------------------------------------------/-> orient + BRIEF
        launch_pyramid(g, b, stream, b1);
        launch_fast_all(g, b, stream, b1);
        // the blur (1 524 CTAs) would queue in front of FAST's CTAs if it started with them; behind FAST it fills the
        // machine while the quadtree occupies 8 SMs
        cudaEventRecord(ev[0], stream);
        cudaStreamWaitEvent(ls[0], ev[0], 0);
        launch_blur(g, b, ls[0], b1);
        cudaEventRecord(ev[1], ls[0]);
        launch_octtree(g, b, -1, h->key_smem_cap, h->oct_smem, stream, b1);
        cudaStreamWaitEvent(stream, ev[1], 0);
        launch_orient_desc(g, b, stream, b1);
        if (d2h) {
            cudaMemcpyAsync(h->h_out, h->d_out, h->out_bytes, cudaMemcpyDeviceToHost, stream);
            if (peer) cudaMemcpyAsync(peer->h_out, peer->d_out, peer->out_bytes, cudaMemcpyDeviceToHost, stream);
        }
        return;
    }
    for (int l = 0; l < L; l++) {
        if (l > 0 && (h->tail_base == 0 || l <= h->tail_base)) launch_resize(g, b, l, stream, b1);
        else if (l > 0 && l == h->tail_base + 1) launch_pyramid(g, b, stream, b1, true);  // levels l .. L - 1 at once
        cudaEventRecord(ev[l], stream);               // level l exists
        cudaStreamWaitEvent(ls[l], ev[l], 0);
        launch_fast_cells(g, b, l, ls[l], b1);
        launch_octtree(g, b, l, h->key_smem_cap, h->oct_smem, ls[l], b1);
        cudaEventRecord(ev[L + l], ls[l]);            // level l distributed
    }
    // the blur is needed only by BRIEF at the very end: least urgent stream, behind the last resize
    cudaEventRecord(ev[2 * L], stream);
    cudaStreamWaitEvent(ls[L], ev[2 * L], 0);
    launch_blur(g, b, ls[L], b1);
    cudaEventRecord(ev[2 * L + 1], ls[L]);
    cudaStreamWaitEvent(stream, ev[2 * L + 1], 0);
    for (int l = 0; l < L; l++) cudaStreamWaitEvent(stream, ev[L + l], 0);
    launch_orient_desc(g, b, stream, b1);
    if (d2h) {
        cudaMemcpyAsync(h->h_out, h->d_out, h->out_bytes, cudaMemcpyDeviceToHost, stream);
        if (peer) cudaMemcpyAsync(peer->h_out, peer->d_out, peer->out_bytes, cudaMemcpyDeviceToHost, stream);
    }
}

struct CaptureScratch {
    std::vector<cudaStream_t> ls;
    std::vector<cudaEvent_t> ev;
    int init(int n_streams, int n_events) {
        ls.assign(n_streams, nullptr);
        ev.assign(n_events, nullptr);
        // kernel nodes inherit the priority of the stream they were captured on: the branches of the coarse levels
        // (index 0 = level 0) get the highest priority, so their CTAs are not queued behind later, cheaper launches
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);  // lo = least urgent (0), hi = most urgent (negative)
        ls.push_back(nullptr);  // one extra, least urgent stream (the blur branch)
        for (size_t i = 0; i < ls.size(); i++) {
            // CORB_PRIO (A/B switch): 0 = no priorities, 1 = level branches from the most urgent level down,
            // 2 (default) = the serial resize chain (main stream, most urgent) above every level branch
            const char* pe = getenv("CORB_PRIO");
            const int mode = pe ? atoi(pe) : 2;
            const int prio = mode == 0 ? lo : i + 1 == ls.size() ? lo : std::min(lo, hi + (mode == 2 ? 1 : 0) + (int)i / 2);
            CORB_CUDA(cudaStreamCreateWithPriority(&ls[i], cudaStreamNonBlocking, prio));
        }
        for (auto& e : ev) CORB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        return CORB_OK;
    }
    ~CaptureScratch() {
        for (auto& s : ls) if (s) cudaStreamDestroy(s);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
    }
};

// finds the k_import node(s) of a captured graph; `dst` selects the one writing to that level-0 buffer
static int find_import_node(cudaGraph_t graph, const uint8_t* dst, cudaGraphNode_t* out) {
    size_t n_nodes = 0;
    CORB_CUDA(cudaGraphGetNodes(graph, nullptr, &n_nodes));
    std::vector<cudaGraphNode_t> nodes(n_nodes);
    CORB_CUDA(cudaGraphGetNodes(graph, nodes.data(), &n_nodes));
    *out = nullptr;
    for (cudaGraphNode_t nd : nodes) {
        cudaGraphNodeType t;
        CORB_CUDA(cudaGraphNodeGetType(nd, &t));
        if (t != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp;
        CORB_CUDA(cudaGraphKernelNodeGetParams(nd, &kp));
        if ((kp.func == import_kernel_ptr(false) || kp.func == import_kernel_ptr(true)) &&
            *reinterpret_cast<uint8_t* const*>(kp.kernelParams[2]) == dst)
            *out = nd;
    }
    CORB_CHECK(*out, CORB_ERR_CUDA, "import node not found in the captured graph");
    return CORB_OK;
}

static int find_h2d_node(cudaGraph_t graph, const uint8_t* dst, cudaGraphNode_t* out);

static int record_graph_variant(corb_orb* h, int variant) {
    const int L = h->geom.n_levels;
    CaptureScratch sc;
    int rc = sc.init(L, 2 * L + 3);
    if (rc != CORB_OK) return rc;
    CORB_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    capture_frame(h, h->stream, variant == 1, sc.ls, sc.ev);
    CORB_CUDA(cudaStreamEndCapture(h->stream, &h->graph[variant]));
    rc = variant == 1 && h->h2d_node ? find_h2d_node(h->graph[variant], h->d_stage, &h->import_node[variant])
                      : find_import_node(h->graph[variant], h->buf.pyr + h->geom.lv[0].img_off, &h->import_node[variant]);
    if (rc != CORB_OK) return rc;
    CORB_CUDA(cudaGraphInstantiate(&h->graph_exec[variant], h->graph[variant], 0));
    h->import_at[variant] = PatchKey();
    return CORB_OK;
}

static int record_graph(corb_orb* h) {
    for (int v = 0; v < 2; v++) {
        int rc = record_graph_variant(h, v);
        if (rc != CORB_OK) return rc;
    }
    h->kernel_launches = h->graph_mode == 2 ? 1 + (h->geom.n_levels - 1) + 2 * h->geom.n_levels + 2
                       : h->graph_mode == 1 ? 6 : 2 + 2 * h->geom.n_levels + 2;
    if (h->graph_mode == 2 && h->tail_base) h->kernel_launches -= h->geom.n_levels - 1 - h->tail_base - 1;
    return CORB_OK;
}

// finds the H2D memcpy node that fills `dst` (a handle's staging buffer)
static int find_h2d_node(cudaGraph_t graph, const uint8_t* dst, cudaGraphNode_t* out) {
    size_t n_nodes = 0;
    CORB_CUDA(cudaGraphGetNodes(graph, nullptr, &n_nodes));
    std::vector<cudaGraphNode_t> nodes(n_nodes);
    CORB_CUDA(cudaGraphGetNodes(graph, nodes.data(), &n_nodes));
    *out = nullptr;
    for (cudaGraphNode_t nd : nodes) {
        cudaGraphNodeType t;
        CORB_CUDA(cudaGraphNodeGetType(nd, &t));
        if (t != cudaGraphNodeTypeMemcpy) continue;
        cudaMemcpy3DParms mp;
        CORB_CUDA(cudaGraphMemcpyNodeGetParams(nd, &mp));
        if (mp.dstPtr.ptr == (void*)dst) *out = nd;
    }
    CORB_CHECK(*out, CORB_ERR_CUDA, "H2D node not found in the captured graph");
    return CORB_OK;
}


static int patch_memcpy_import(cudaGraphExec_t exec, cudaGraphNode_t node, corb_orb* h, const uint8_t* src, int stride) {
    const LevelGeom& L0 = h->geom.lv[0];
    CORB_CHECK(stride == L0.w, CORB_ERR_INVALID, "host images are staged contiguously before the H2D node");
    CORB_CUDA(cudaGraphExecMemcpyNodeSetParams1D(exec, node, h->d_stage, src, (size_t)L0.w * L0.h, cudaMemcpyHostToDevice));
    return CORB_OK;
}

static int patch_import(cudaGraphExec_t exec, cudaGraphNode_t node, corb_orb* h, const uint8_t* src, int stride,
                        corb_orb* peer = nullptr, const uint8_t* src1 = nullptr, int stride1 = 0, cudaGraphNode_t node1 = nullptr,
                        PatchKey* cache = nullptr) {
    // Re-pointing a node costs a few microseconds of driver time in front of every frame: skip it when the node already
    // points at these buffers (a client that reuses its image buffers, or whose pageable images are staged into h_img).
    if (cache) {
        const PatchKey now = {src, stride, src1, stride1};
        if (cache->src == now.src && cache->stride == now.stride && cache->src1 == now.src1 && cache->stride1 == now.stride1) return CORB_OK;
        *cache = now;
    }
    cudaGraphNodeType nt;
    CORB_CUDA(cudaGraphNodeGetType(node, &nt));
    if (nt == cudaGraphNodeTypeMemcpy) {
        int rc = patch_memcpy_import(exec, node, h, src, stride);
        if (rc == CORB_OK && peer) rc = patch_memcpy_import(exec, node1, peer, src1, stride1);
        return rc;
    }
    const LevelGeom& L0 = h->geom.lv[0];
    uint8_t* dst = h->buf.pyr + L0.img_off;
    uint8_t* dst1 = peer ? peer->buf.pyr + L0.img_off : nullptr;
    int pitch = L0.pitch, w = L0.w, hh = L0.h;
    void* args[9] = {(void*)&src, (void*)&stride, (void*)&dst, (void*)&pitch, (void*)&w, (void*)&hh, (void*)&src1, (void*)&stride1,
                     (void*)&dst1};
    cudaKernelNodeParams cur;
    CORB_CUDA(cudaGraphKernelNodeGetParams(node, &cur));  // which of the two import kernels this node runs
    const bool host_src = cur.func == import_kernel_ptr(true);
    cudaKernelNodeParams kp = {};
    kp.func = cur.func;
    import_launch_dims(L0.w, L0.h, peer ? 2 : 1, host_src, &kp.gridDim, &kp.blockDim);
    kp.sharedMemBytes = 0;
    kp.kernelParams = args;
    CORB_CUDA(cudaGraphExecKernelNodeSetParams(exec, node, &kp));
    return CORB_OK;
}

// Work that touched this handle's buffers may have been enqueued on a peer's stream (pair graphs run on the left
// handle's stream): order this handle's own stream behind it before using the buffers again.
static int settle(corb_orb* h) {
    if (h->busy_stream && h->busy_stream != h->stream) {
        CORB_CUDA(cudaEventRecord(h->ev_busy, h->busy_stream));
        CORB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_busy, 0));
    }
    h->busy_stream = nullptr;
    return CORB_OK;
}

// Launches the per-frame graph on (src, stride): src is device memory or page-locked host memory (UVA).
static int launch_frame(corb_orb* h, int variant, const uint8_t* src, int stride) {
    int rc = settle(h);
    if (rc != CORB_OK) return rc;
    rc = patch_import(h->graph_exec[variant], h->import_node[variant], h, src, stride, nullptr, nullptr, 0, nullptr, &h->import_at[variant]);
    if (rc != CORB_OK) return rc;
    CORB_CUDA(cudaGraphLaunch(h->graph_exec[variant], h->stream));
    h->host_count = variant >= 1;  // the host variants end with the D2H of keypoints, descriptors and the count
    h->own_dirty = true;
    return CORB_OK;
}

static void fill_stereo_args(corb_orb* left, corb_orb* right, float mbf, float mb, StereoArgs* a) {
    const OrbGeom& g = left->geom;
    memset(a, 0, sizeof(*a));
    a->kl = left->buf.kps; a->kr = right->buf.kps;
    a->dl = reinterpret_cast<const uint4*>(left->buf.desc); a->dr = reinterpret_cast<const uint4*>(right->buf.desc);
    a->nl = left->buf.count; a->nr = right->buf.count;
    a->pyr_l = left->buf.pyr; a->pyr_r = right->buf.pyr;
    for (int l = 0; l < left->nlevels; l++) { a->scale[l] = left->scale[l]; a->inv_scale[l] = left->inv_scale[l]; }
    a->mbf = mbf; a->mb = mb;
    a->u_right = left->d_stereo; a->depth = left->d_stereo + g.kp_cap;
    a->best_dist = reinterpret_cast<int*>(left->d_stereo + 2 * (size_t)g.kp_cap);
    a->n_rows = g.lv[0].h;
    a->n_bands = left->stereo_bands;
    a->band_cap = g.kp_cap;
    a->band_cnt = left->d_bands;
    a->band_list = left->d_bands + align_up(left->stereo_bands, 64);
    a->rinfo = reinterpret_cast<int4*>(left->d_bands + align_up(left->stereo_bands, 64) + align_up_sz((size_t)left->stereo_bands * g.kp_cap, 4));
}

static int ensure_stereo_buffers(corb_orb* left) {
    const OrbGeom& g = left->geom;
    if (left->stereo_cap != g.kp_cap) {
        int rc = dev_alloc(left, &left->d_stereo, 3 * (size_t)g.kp_cap);
        if (rc != CORB_OK) return rc;
        if (left->h_stereo) cudaFreeHost(left->h_stereo), left->h_stereo = nullptr;
        CORB_CUDA(cudaMallocHost(&left->h_stereo, 2 * (size_t)g.kp_cap * sizeof(float)));
        left->stereo_cap = g.kp_cap;
        // 8-row bands of right keypoints: counters | lists [bands][kp_cap] | per-keypoint records (int4)
        left->stereo_bands = (g.lv[0].h + 7) / 8 + 1;
        const size_t ints = (size_t)align_up(left->stereo_bands, 64) + align_up_sz((size_t)left->stereo_bands * g.kp_cap, 4) + 4 * (size_t)g.kp_cap + 16;
        rc = dev_alloc(left, &left->d_bands, ints);
        if (rc != CORB_OK) return rc;
        CORB_CUDA(cudaMemset(left->d_bands, 0, (size_t)align_up(left->stereo_bands, 64) * sizeof(int)));
    }
    return CORB_OK;
}

static int check_pair(corb_orb* left, corb_orb* right) {
    CORB_CHECK(left && right && left != right, CORB_ERR_INVALID, "two distinct handles are required");
    CORB_CHECK(left->plan_w && left->plan_w == right->plan_w && left->plan_h == right->plan_h && left->nlevels == right->nlevels &&
                   left->device == right->device && left->scale_factor_f == right->scale_factor_f &&
                   left->nfeatures == right->nfeatures && left->ini_th == right->ini_th && left->min_th == right->min_th,
               CORB_ERR_INVALID, "left and right extractor must share device, image size and ORB parameters (Tracking.cc:118-121)");
    return CORB_OK;
}

// One graph for both images of a stereo frame, on the left handle's stream (see corb_orb::pair_graph).
static int ensure_pair_graph(corb_orb* hl, corb_orb* hr, int variant, float mbf, float mb) {
    int rc = check_pair(hl, hr);
    if (rc != CORB_OK) return rc;
    if (hl->pair_peer != hr || hl->pair_peer_uid != hr->uid || hl->pair_peer_serial != hr->plan_serial ||
        hl->pair_self_serial != hl->plan_serial) {
        for (int v = 0; v < 3; v++) {
            if (hl->pair_exec[v]) cudaGraphExecDestroy(hl->pair_exec[v]), hl->pair_exec[v] = nullptr;
            if (hl->pair_graph[v]) cudaGraphDestroy(hl->pair_graph[v]), hl->pair_graph[v] = nullptr;
        }
        hl->pair_peer = hr;
        hl->pair_peer_uid = hr->uid;
        hl->pair_peer_serial = hr->plan_serial;
        hl->pair_self_serial = hl->plan_serial;
    }
    if (variant == 2 && hl->pair_exec[2] && (hl->pair_mbf != mbf || hl->pair_mb != mb)) {
        cudaGraphExecDestroy(hl->pair_exec[2]); hl->pair_exec[2] = nullptr;
        cudaGraphDestroy(hl->pair_graph[2]); hl->pair_graph[2] = nullptr;
    }
    if (hl->pair_exec[variant]) return CORB_OK;
    const int L = hl->geom.n_levels;
    if (variant == 2 && (rc = ensure_stereo_buffers(hl)) != CORB_OK) return rc;
    CaptureScratch sl, sr, sb;
    if ((rc = sl.init(L, 2 * L + 3)) != CORB_OK) return rc;
    // Host images arrive over PCIe one after the other (the link is the bottleneck: ~16 us per image). The host variants
    // give each image its own launches so that the left image's pipeline starts as soon as *its* bytes are in, while
    // the right image is still being read (measured: 145 -> 132 us per frame); device-resident images are processed
    // together in every launch (grid z = 2). CORB_PAIR_MERGED=1 forces the merged form everywhere (A/B switch).
    const bool split = (variant >= 1 || getenv("CORB_PAIR_SPLIT_DEV") != nullptr) && getenv("CORB_PAIR_MERGED") == nullptr && !hl->h2d_node;
    hl->pair_split[variant] = split;
    if (split && ((rc = sr.init(L, 2 * L + 3)) != CORB_OK || (rc = sb.init(1, 2)) != CORB_OK)) return rc;
    CORB_CUDA(cudaStreamBeginCapture(hl->stream, cudaStreamCaptureModeThreadLocal));
    if (split) {
        capture_frame(hl, hl->stream, variant >= 1, sl.ls, sl.ev, nullptr, sb.ev[0]);
        cudaStreamWaitEvent(sb.ls[0], sb.ev[0], 0);  // right import behind the left import
        capture_frame(hr, sb.ls[0], variant >= 1, sr.ls, sr.ev);
        cudaEventRecord(sb.ev[1], sb.ls[0]);
        cudaStreamWaitEvent(hl->stream, sb.ev[1], 0);
    } else {
        capture_frame(hl, hl->stream, variant >= 1, sl.ls, sl.ev, hr);  // both images in every launch (grid z = 2)
    }
    if (variant == 2) {
        StereoArgs a;
        fill_stereo_args(hl, hr, mbf, mb, &a);
        launch_stereo(hl->geom, a, hl->stream);
        cudaMemcpyAsync(hl->h_stereo, hl->d_stereo, 2 * (size_t)hl->geom.kp_cap * sizeof(float), cudaMemcpyDeviceToHost, hl->stream);
        hl->pair_mbf = mbf;
        hl->pair_mb = mb;
    }
    CORB_CUDA(cudaStreamEndCapture(hl->stream, &hl->pair_graph[variant]));
    if (variant >= 1 && hl->h2d_node) {
        if ((rc = find_h2d_node(hl->pair_graph[variant], hl->d_stage, &hl->pair_imp_l[variant])) != CORB_OK) return rc;
        if ((rc = find_h2d_node(hl->pair_graph[variant], hr->d_stage, &hl->pair_imp_r[variant])) != CORB_OK) return rc;
    } else if ((rc = find_import_node(hl->pair_graph[variant], hl->buf.pyr + hl->geom.lv[0].img_off, &hl->pair_imp_l[variant])) != CORB_OK) {
        return rc;
    }
    if (split && (rc = find_import_node(hl->pair_graph[variant], hr->buf.pyr + hr->geom.lv[0].img_off, &hl->pair_imp_r[variant])) != CORB_OK)
        return rc;
    CORB_CUDA(cudaGraphInstantiate(&hl->pair_exec[variant], hl->pair_graph[variant], 0));
    hl->pair_at_l[variant] = hl->pair_at_r[variant] = PatchKey();
    return CORB_OK;
}

// stages a host image for the import kernel: page-locked caller memory is read in place, pageable memory is copied
static void stage_input(corb_orb* h, const uint8_t* img, int w, int hgt, int stride, const uint8_t** src, int* src_stride) {
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, img) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    if (!pinned) cudaGetLastError();
    *src = img;
    *src_stride = stride;
    if (!pinned || (h->h2d_node && stride != w)) {  // the H2D node copies w * hgt contiguous bytes
        if (stride == w) memcpy(h->h_img, img, (size_t)w * hgt);
        else for (int y = 0; y < hgt; y++) memcpy(h->h_img + (size_t)y * w, img + (size_t)y * stride, w);
        *src = h->h_img;
        *src_stride = w;
    }
}

static int launch_pair(corb_orb* hl, corb_orb* hr, int variant, const uint8_t* src_l, int stride_l, const uint8_t* src_r, int stride_r) {
    int rc = settle(hl);
    if (rc != CORB_OK) return rc;
    if (hr->own_dirty) {  // earlier stand-alone work on the right handle's own stream must finish first
        CORB_CUDA(cudaEventRecord(hr->ev_busy, hr->stream));
        CORB_CUDA(cudaStreamWaitEvent(hl->stream, hr->ev_busy, 0));
        hr->own_dirty = false;
    }
    if (hl->pair_split[variant]) {
        if ((rc = patch_import(hl->pair_exec[variant], hl->pair_imp_l[variant], hl, src_l, stride_l, nullptr, nullptr, 0, nullptr,
                               &hl->pair_at_l[variant])) != CORB_OK) return rc;
        if ((rc = patch_import(hl->pair_exec[variant], hl->pair_imp_r[variant], hr, src_r, stride_r, nullptr, nullptr, 0, nullptr,
                               &hl->pair_at_r[variant])) != CORB_OK) return rc;
    } else if ((rc = patch_import(hl->pair_exec[variant], hl->pair_imp_l[variant], hl, src_l, stride_l, hr, src_r, stride_r,
                                  hl->pair_imp_r[variant], &hl->pair_at_l[variant])) != CORB_OK) {
        return rc;
    }
    CORB_CUDA(cudaGraphLaunch(hl->pair_exec[variant], hl->stream));
    hl->host_count = hr->host_count = variant >= 1;
    hl->own_dirty = true;
    hr->busy_stream = hl->stream;
    return CORB_OK;
}

static int collect(corb_orb* h, corb_keypoint* kps, uint8_t* desc, int* n) {
    CORB_CHECK(h->h_scalars[1] == 0, CORB_ERR_CAPACITY, "device-side consistency check %d failed", h->h_scalars[1]);
    const int cnt = h->h_scalars[0];
    CORB_CHECK(cnt >= 0 && cnt <= h->geom.kp_cap, CORB_ERR_CAPACITY, "keypoint count %d out of range", cnt);
    if (n) *n = cnt;
    if (kps) memcpy(kps, h->h_kps, sizeof(corb_keypoint) * cnt);
    if (desc) memcpy(desc, h->h_desc, (size_t)cnt * 32);
    return CORB_OK;
}

static void copy_pyramid(corb_orb* h, uint8_t* const* pyr_out) {
    for (int l = 0; l < h->nlevels; l++) {
        if (!pyr_out[l]) continue;
        const LevelGeom& L = h->geom.lv[l];
        for (int y = 0; y < L.h; y++) memcpy(pyr_out[l] + (size_t)y * L.w, h->h_pyr + L.img_off + (size_t)y * L.pitch, L.w);
    }
}

extern "C" {

const char* corb_last_error(void) { return corb::get_error(); }
int corb_version(void) { return 100; }
int corb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int corb_orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int device, corb_orb** out) {
    CORB_CHECK(out, CORB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    CORB_CHECK(nfeatures >= 1 && nlevels >= 1 && nlevels <= kMaxLevels && scale_factor > 1.f && ini_th >= 1 && min_th >= 1 &&
                   ini_th <= 254 && min_th <= 254,
               CORB_ERR_INVALID, "invalid ORB parameters (nfeatures %d, scale %g, levels %d, FAST %d/%d)", nfeatures,
               (double)scale_factor, nlevels, ini_th, min_th);
    int ndev = 0;
    CORB_CUDA(cudaGetDeviceCount(&ndev));
    CORB_CHECK(device >= 0 && device < ndev, CORB_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
    corb_orb* h = new corb_orb;
    {
        static std::atomic<long long> next_uid{1};
        h->uid = next_uid.fetch_add(1);
    }
    h->nfeatures = nfeatures; h->nlevels = nlevels; h->ini_th = ini_th; h->min_th = min_th; h->device = device;
    h->scale_factor_f = scale_factor;
    h->scale_factor = scale_factor;
    h->scale.resize(nlevels); h->inv_scale.resize(nlevels); h->sigma2.resize(nlevels); h->inv_sigma2.resize(nlevels);
    h->scale[0] = 1.f; h->sigma2[0] = 1.f;
    for (int i = 1; i < nlevels; i++) {
        h->scale[i] = (float)(h->scale[i - 1] * h->scale_factor);
        h->sigma2[i] = h->scale[i] * h->scale[i];
    }
    for (int i = 0; i < nlevels; i++) {
        h->inv_scale[i] = 1.0f / h->scale[i];
        h->inv_sigma2[i] = 1.0f / h->sigma2[i];
    }
    h->quota.resize(nlevels);
    const float factor = (float)(1.0f / h->scale_factor);
    float desired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        h->quota[l] = cv_round_f(desired);
        sum += h->quota[l];
        desired *= factor;
    }
    h->quota[nlevels - 1] = std::max(nfeatures - sum, 0);
    {  // umax (:452-469)
        int v, v0;
        const int vmax = (int)floorf(kHalfPatch * sqrtf(2.f) / 2 + 1), vmin = (int)ceilf(kHalfPatch * sqrtf(2.f) / 2);
        const double hp2 = kHalfPatch * kHalfPatch;
        for (v = 0; v <= vmax; ++v) h->umax[v] = cv_round_d(sqrt(hp2 - v * v));
        for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
            while (h->umax[v0] == h->umax[v0 + 1]) ++v0;
            h->umax[v] = v0;
            ++v0;
        }
    }
    cudaError_t e = cudaSetDevice(device);
    {
        int lo = 0, hi = 0;
        if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, hi);  // level-0 branch runs here
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_busy, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        set_error("CUDA setup on device %d failed: %s", device, cudaGetErrorString(e));
        corb_orb_destroy(h);
        return CORB_ERR_CUDA;
    }
    *out = h;
    return CORB_OK;
}

void corb_orb_destroy(corb_orb* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) {
        settle(h);
        cudaStreamSynchronize(h->stream);
    }
    free_plan(h);
    if (h->ev_peer) cudaEventDestroy(h->ev_peer);
    if (h->ev_busy) cudaEventDestroy(h->ev_busy);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int corb_orb_levels(const corb_orb* h) { return h ? h->nlevels : 0; }
float corb_orb_scale_factor(const corb_orb* h) { return h ? h->scale_factor_f : 0.f; }

int corb_orb_tables(const corb_orb* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* quota, int* umax16) {
    CORB_CHECK(h, CORB_ERR_INVALID, "handle is NULL");
    for (int i = 0; i < h->nlevels; i++) {
        if (scale) scale[i] = h->scale[i];
        if (inv_scale) inv_scale[i] = h->inv_scale[i];
        if (sigma2) sigma2[i] = h->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = h->inv_sigma2[i];
        if (quota) quota[i] = h->quota[i];
    }
    if (umax16) memcpy(umax16, h->umax, sizeof(h->umax));
    return CORB_OK;
}

int corb_orb_level_size(const corb_orb* h, int level, int w, int hgt, int* lw, int* lh) {
    CORB_CHECK(h && level >= 0 && level < h->nlevels && lw && lh, CORB_ERR_INVALID, "bad argument");
    *lw = cv_round_f((float)w * h->inv_scale[level]);
    *lh = cv_round_f((float)hgt * h->inv_scale[level]);
    return CORB_OK;
}

int corb_orb_capacity(const corb_orb* h, int w, int hgt) {
    if (!h || w < 1 || hgt < 1) return -1;
    return capacity_for(h, w, hgt);
}

int corb_orb_extract_submit(corb_orb* h, const uint8_t* img, int w, int hgt, int stride, int want_pyramid) {
    CORB_CHECK(h, CORB_ERR_INVALID, "handle is NULL");
    CORB_CHECK(!h->pending, CORB_ERR_INVALID, "a submitted extraction has not been waited for");
    if (!img || w < 1 || hgt < 1) {  // empty image: the reference returns silently (ORBextractor.cc:1046)
        h->pending = true;
        h->pending_pyr = false;
        h->pending_empty = true;
        return CORB_OK;
    }
    CORB_CHECK(stride >= w, CORB_ERR_INVALID, "stride %d < width %d", stride, w);
    CORB_CUDA(cudaSetDevice(h->device));
    int rc = make_plan(h, w, hgt);
    if (rc != CORB_OK) return rc;

    const uint8_t* src;
    int src_stride;
    stage_input(h, img, w, hgt, stride, &src, &src_stride);
    rc = launch_frame(h, 1, src, src_stride);  // import (PCIe read) -> kernels -> D2H of the result blob, one launch
    if (rc != CORB_OK) return rc;
    if (want_pyramid)
        CORB_CUDA(cudaMemcpyAsync(h->h_pyr, h->buf.pyr, h->pyr_bytes, cudaMemcpyDeviceToHost, h->stream));
    h->pending = true;
    h->pending_pyr = want_pyramid != 0;
    h->pending_empty = false;
    return CORB_OK;
}

int corb_orb_extract_wait(corb_orb* h, corb_keypoint* kps, uint8_t* desc, int* n, uint8_t* const* pyr_out) {
    CORB_CHECK(h && n, CORB_ERR_INVALID, "bad argument");
    CORB_CHECK(h->pending, CORB_ERR_INVALID, "nothing was submitted");
    h->pending = false;
    if (h->pending_empty) {
        *n = 0;
        return CORB_OK;
    }
    CORB_CUDA(cudaSetDevice(h->device));
    CORB_CUDA(cudaStreamSynchronize(h->stream));
    h->own_dirty = false;
    int rc = collect(h, kps, desc, n);
    if (rc != CORB_OK) return rc;
    if (pyr_out) {
        CORB_CHECK(h->pending_pyr, CORB_ERR_INVALID, "pyramid requested at wait but not at submit");
        copy_pyramid(h, pyr_out);
    }
    return CORB_OK;
}

int corb_orb_extract(corb_orb* h, const uint8_t* img, int w, int hgt, int stride, corb_keypoint* kps, uint8_t* desc, int* n,
                     uint8_t* const* pyr_out) {
    int rc = corb_orb_extract_submit(h, img, w, hgt, stride, pyr_out != nullptr);
    if (rc != CORB_OK) return rc;
    return corb_orb_extract_wait(h, kps, desc, n, pyr_out);
}

// Left and right image of one stereo frame from one thread (the reference starts two threads, Frame.cc:78-81): both
// frames are one CUDA graph on the left handle's stream (two parallel branches), i.e. one driver launch per stereo frame.
static int pair_prepare(corb_orb* hl, corb_orb* hr, int w, int hgt) {
    CORB_CHECK(hl && hr && hl != hr, CORB_ERR_INVALID, "two distinct handles are required");
    CORB_CHECK(!hl->pending && !hr->pending && !hl->pending_pair, CORB_ERR_INVALID, "a submitted extraction has not been waited for");
    CORB_CHECK(hl->device == hr->device, CORB_ERR_INVALID, "left and right handle must live on the same device");
    CORB_CUDA(cudaSetDevice(hl->device));
    int rc = make_plan(hl, w, hgt);
    if (rc != CORB_OK) return rc;
    return make_plan(hr, w, hgt);
}

// pending_pair: 0 none, 1 extract_pair submitted, 2 frame_stereo submitted (state lives on the left handle)
int corb_orb_extract_pair_submit(corb_orb* hl, corb_orb* hr, const uint8_t* img_l, const uint8_t* img_r, int w, int hgt, int stride,
                                 int want_pyramid) {
    CORB_CHECK(hl && hr, CORB_ERR_INVALID, "bad argument");
    CORB_CHECK(img_l && img_r && w >= 1 && hgt >= 1, CORB_ERR_INVALID, "the split form needs two non-empty images");
    CORB_CHECK(stride >= w, CORB_ERR_INVALID, "stride %d < width %d", stride, w);
    CORB_CHECK(!hl->pending_pair, CORB_ERR_INVALID, "a submitted stereo pair has not been waited for");
    int rc = pair_prepare(hl, hr, w, hgt);
    if (rc != CORB_OK) return rc;
    if ((rc = ensure_pair_graph(hl, hr, 1, 0.f, 0.f)) != CORB_OK) return rc;
    const uint8_t *sl, *sr;
    int stl, str_;
    stage_input(hl, img_l, w, hgt, stride, &sl, &stl);
    stage_input(hr, img_r, w, hgt, stride, &sr, &str_);
    if ((rc = launch_pair(hl, hr, 1, sl, stl, sr, str_)) != CORB_OK) return rc;
    if (want_pyramid) {
        CORB_CUDA(cudaMemcpyAsync(hl->h_pyr, hl->buf.pyr, hl->pyr_bytes, cudaMemcpyDeviceToHost, hl->stream));
        CORB_CUDA(cudaMemcpyAsync(hr->h_pyr, hr->buf.pyr, hr->pyr_bytes, cudaMemcpyDeviceToHost, hl->stream));
    }
    hl->pending_pair = 1;
    hl->pending_pyr = want_pyramid != 0;
    return CORB_OK;
}

int corb_orb_extract_pair_wait(corb_orb* hl, corb_orb* hr, corb_keypoint* kps_l, uint8_t* desc_l, int* n_l, corb_keypoint* kps_r,
                               uint8_t* desc_r, int* n_r, uint8_t* const* pyr_l, uint8_t* const* pyr_r) {
    CORB_CHECK(hl && hr && n_l && n_r, CORB_ERR_INVALID, "bad argument");
    CORB_CHECK(hl->pending_pair == 1, CORB_ERR_INVALID, "no stereo pair was submitted on these handles");
    CORB_CHECK(!(pyr_l || pyr_r) || hl->pending_pyr, CORB_ERR_INVALID, "pyramid requested at wait but not at submit");
    hl->pending_pair = 0;
    CORB_CUDA(cudaSetDevice(hl->device));
    CORB_CUDA(cudaStreamSynchronize(hl->stream));
    hl->own_dirty = false;
    hr->busy_stream = nullptr;
    int rc;
    if ((rc = collect(hl, kps_l, desc_l, n_l)) != CORB_OK) return rc;
    if ((rc = collect(hr, kps_r, desc_r, n_r)) != CORB_OK) return rc;
    if (pyr_l) copy_pyramid(hl, pyr_l);
    if (pyr_r) copy_pyramid(hr, pyr_r);
    return CORB_OK;
}

int corb_orb_extract_pair(corb_orb* hl, corb_orb* hr, const uint8_t* img_l, const uint8_t* img_r, int w, int hgt, int stride,
                          corb_keypoint* kps_l, uint8_t* desc_l, int* n_l, corb_keypoint* kps_r, uint8_t* desc_r, int* n_r,
                          uint8_t* const* pyr_l, uint8_t* const* pyr_r) {
    CORB_CHECK(hl && hr && n_l && n_r, CORB_ERR_INVALID, "bad argument");
    if (!img_l || !img_r || w < 1 || hgt < 1) {  // empty image(s): fall back to the per-handle path (silent return, :1046)
        int rc = corb_orb_extract(hl, img_l, w, hgt, stride, kps_l, desc_l, n_l, pyr_l);
        if (rc != CORB_OK) return rc;
        return corb_orb_extract(hr, img_r, w, hgt, stride, kps_r, desc_r, n_r, pyr_r);
    }
    int rc = corb_orb_extract_pair_submit(hl, hr, img_l, img_r, w, hgt, stride, pyr_l || pyr_r);
    if (rc != CORB_OK) return rc;
    return corb_orb_extract_pair_wait(hl, hr, kps_l, desc_l, n_l, kps_r, desc_r, n_r, pyr_l, pyr_r);
}

int corb_orb_extract_pair_device(corb_orb* hl, corb_orb* hr, const uint8_t* d_img_l, const uint8_t* d_img_r, int w, int hgt,
                                 int stride) {
    CORB_CHECK(d_img_l && d_img_r && w >= 1 && hgt >= 1 && stride >= w, CORB_ERR_INVALID, "bad argument");
    int rc = pair_prepare(hl, hr, w, hgt);
    if (rc != CORB_OK) return rc;
    if ((rc = ensure_pair_graph(hl, hr, 0, 0.f, 0.f)) != CORB_OK) return rc;
    return launch_pair(hl, hr, 0, d_img_l, stride, d_img_r, stride);
}

// ---- Frame::ComputeStereoMatches (Frame.cc:470-644) on the resident results of the last extraction of `left`/`right`
int corb_stereo_match(corb_orb* left, corb_orb* right, float mbf, float mb, int n_left, float* u_right, float* depth) {
    CORB_CHECK(u_right && depth && n_left >= 0, CORB_ERR_INVALID, "bad argument");
    CORB_CHECK(left && !left->pending && right && !right->pending, CORB_ERR_INVALID, "wait for the submitted extractions first");
    CORB_CHECK(mb > 0.f && mbf > 0.f, CORB_ERR_INVALID, "mbf and mb must be positive");
    int rc = check_pair(left, right);
    if (rc != CORB_OK) return rc;
    CORB_CUDA(cudaSetDevice(left->device));
    if ((rc = ensure_stereo_buffers(left)) != CORB_OK) return rc;
    if ((rc = settle(left)) != CORB_OK) return rc;
    // the right extraction must have finished: order the left stream behind whatever stream produced it
    cudaStream_t rs = right->busy_stream ? right->busy_stream : right->stream;
    if (rs != left->stream) {
        CORB_CUDA(cudaEventRecord(right->ev_busy, rs));
        CORB_CUDA(cudaStreamWaitEvent(left->stream, right->ev_busy, 0));
    }
    StereoArgs a;
    fill_stereo_args(left, right, mbf, mb, &a);
    launch_stereo(left->geom, a, left->stream);
    CORB_CUDA(cudaGetLastError());
    CORB_CUDA(cudaMemcpyAsync(left->h_stereo, left->d_stereo, 2 * (size_t)left->geom.kp_cap * sizeof(float), cudaMemcpyDeviceToHost,
                              left->stream));
    CORB_CUDA(cudaStreamSynchronize(left->stream));
    left->own_dirty = false;
    CORB_CHECK(n_left <= left->geom.kp_cap, CORB_ERR_INVALID, "n_left exceeds the keypoint capacity");
    memcpy(u_right, left->h_stereo, n_left * sizeof(float));
    memcpy(depth, left->h_stereo + left->geom.kp_cap, n_left * sizeof(float));
    return CORB_OK;
}

int corb_frame_stereo_submit(corb_orb* hl, corb_orb* hr, const uint8_t* img_l, const uint8_t* img_r, int w, int hgt, int stride,
                             float mbf, float mb) {
    CORB_CHECK(hl && hr && hl != hr, CORB_ERR_INVALID, "bad argument");
    CORB_CHECK(img_l && img_r && w >= 1 && hgt >= 1, CORB_ERR_INVALID, "the split form needs two non-empty images");
    CORB_CHECK(mb > 0.f && mbf > 0.f, CORB_ERR_INVALID, "mbf and mb must be positive");
    CORB_CHECK(stride >= w, CORB_ERR_INVALID, "stride %d < width %d", stride, w);
    CORB_CHECK(!hl->pending_pair, CORB_ERR_INVALID, "a submitted stereo pair has not been waited for");
    int rc = pair_prepare(hl, hr, w, hgt);
    if (rc != CORB_OK) return rc;
    if ((rc = ensure_pair_graph(hl, hr, 2, mbf, mb)) != CORB_OK) return rc;  // extraction x2 + stereo matching + D2H, one launch
    const uint8_t *sl, *sr;
    int stl, str_;
    stage_input(hl, img_l, w, hgt, stride, &sl, &stl);
    stage_input(hr, img_r, w, hgt, stride, &sr, &str_);
    if ((rc = launch_pair(hl, hr, 2, sl, stl, sr, str_)) != CORB_OK) return rc;
    hl->pending_pair = 2;
    return CORB_OK;
}

int corb_frame_stereo_wait(corb_orb* hl, corb_orb* hr, corb_keypoint* kps_l, uint8_t* desc_l, int* n_l, corb_keypoint* kps_r,
                           uint8_t* desc_r, int* n_r, float* u_right, float* depth) {
    CORB_CHECK(hl && hr && n_l && n_r && u_right && depth, CORB_ERR_INVALID, "bad argument");
    CORB_CHECK(hl->pending_pair == 2, CORB_ERR_INVALID, "no stereo frame was submitted on these handles");
    hl->pending_pair = 0;
    CORB_CUDA(cudaSetDevice(hl->device));
    CORB_CUDA(cudaStreamSynchronize(hl->stream));
    hl->own_dirty = false;
    hr->busy_stream = nullptr;
    int rc;
    if ((rc = collect(hl, kps_l, desc_l, n_l)) != CORB_OK) return rc;
    if ((rc = collect(hr, kps_r, desc_r, n_r)) != CORB_OK) return rc;
    memcpy(u_right, hl->h_stereo, *n_l * sizeof(float));
    memcpy(depth, hl->h_stereo + hl->geom.kp_cap, *n_l * sizeof(float));
    return CORB_OK;
}

int corb_frame_stereo(corb_orb* hl, corb_orb* hr, const uint8_t* img_l, const uint8_t* img_r, int w, int hgt, int stride, float mbf,
                      float mb, corb_keypoint* kps_l, uint8_t* desc_l, int* n_l, corb_keypoint* kps_r, uint8_t* desc_r, int* n_r,
                      float* u_right, float* depth) {
    CORB_CHECK(hl && hr && hl != hr && n_l && n_r && u_right && depth, CORB_ERR_INVALID, "bad argument");
    CORB_CHECK(mb > 0.f && mbf > 0.f, CORB_ERR_INVALID, "mbf and mb must be positive");
    if (!img_l || !img_r || w < 1 || hgt < 1) {
        *n_l = *n_r = 0;
        return CORB_OK;
    }
    int rc = corb_frame_stereo_submit(hl, hr, img_l, img_r, w, hgt, stride, mbf, mb);
    if (rc != CORB_OK) return rc;
    return corb_frame_stereo_wait(hl, hr, kps_l, desc_l, n_l, kps_r, desc_r, n_r, u_right, depth);
}

int corb_orb_extract_device(corb_orb* h, const uint8_t* d_img, int w, int hgt, int stride) {
    CORB_CHECK(h && d_img && w >= 1 && hgt >= 1 && stride >= w, CORB_ERR_INVALID, "bad argument");
    CORB_CHECK(!h->pending, CORB_ERR_INVALID, "a submitted extraction has not been waited for");
    CORB_CUDA(cudaSetDevice(h->device));
    int rc = make_plan(h, w, hgt);
    if (rc != CORB_OK) return rc;
    return launch_frame(h, 0, d_img, stride);
}

int corb_orb_sync(corb_orb* h) {
    CORB_CHECK(h, CORB_ERR_INVALID, "handle is NULL");
    CORB_CUDA(cudaSetDevice(h->device));
    int rc = settle(h);
    if (rc != CORB_OK) return rc;
    CORB_CUDA(cudaStreamSynchronize(h->stream));
    h->own_dirty = false;
    return CORB_OK;
}

int corb_orb_device_results(const corb_orb* h, const corb_keypoint** d_kps, const uint8_t** d_desc, const int** d_count) {
    CORB_CHECK(h && h->plan_w, CORB_ERR_INVALID, "no extraction has run on this handle");
    if (d_kps) *d_kps = h->buf.kps;
    if (d_desc) *d_desc = h->buf.desc;
    if (d_count) *d_count = h->buf.count;
    return CORB_OK;
}

int corb_frame_bow(corb_orb* h, corb_voc* v, int levelsup, corb_bow_store* s) {
    CORB_CHECK(h && v && s && h->plan_w, CORB_ERR_INVALID, "bad argument or no extraction has run on this handle");
    CORB_CHECK(!h->pending && !h->pending_pair, CORB_ERR_INVALID, "wait for the submitted extraction first");
    CORB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->busy_stream ? h->busy_stream : h->stream;
    // the keypoint count of the last extraction: already on the host after a host-result extraction (which has been waited
    // for), else read behind the extraction on its stream (device-resident extractions never told the host)
    int n = 0;
    if (h->host_count && h->h_scalars && !h->own_dirty) {
        n = h->h_scalars[0];
    } else {
        CORB_CUDA(cudaMemcpyAsync(&n, h->buf.count, sizeof(int), cudaMemcpyDeviceToHost, st));
        CORB_CUDA(cudaStreamSynchronize(st));
    }
    CORB_CHECK(n >= 0 && n <= h->geom.kp_cap, CORB_ERR_INVALID, "keypoint count %d out of range", n);
    return corb_bow_store_fill(s, v, h->buf.kps, h->buf.desc, n, levelsup, (void*)st);
}

int corb_orb_host_results(const corb_orb* h, const corb_keypoint** kps, const uint8_t** desc, int* n) {
    CORB_CHECK(h && h->plan_w && h->h_scalars, CORB_ERR_INVALID, "no extraction has run on this handle");
    CORB_CHECK(h->h_scalars[1] == 0, CORB_ERR_CAPACITY, "device-side consistency check %d failed", h->h_scalars[1]);
    const int cnt = h->h_scalars[0];
    CORB_CHECK(cnt >= 0 && cnt <= h->geom.kp_cap, CORB_ERR_CAPACITY, "keypoint count %d out of range", cnt);
    if (kps) *kps = h->h_kps;
    if (desc) *desc = h->h_desc;
    if (n) *n = cnt;
    return CORB_OK;
}

int corb_orb_device_level(const corb_orb* h, int level, int blurred, const uint8_t** d_ptr, int* pitch, int* lw, int* lh) {
    CORB_CHECK(h && h->plan_w && level >= 0 && level < h->nlevels, CORB_ERR_INVALID, "bad argument or no plan");
    const LevelGeom& L = h->geom.lv[level];
    if (d_ptr) *d_ptr = (blurred ? h->buf.blur : h->buf.pyr) + L.img_off;
    if (pitch) *pitch = L.pitch;
    if (lw) *lw = L.w;
    if (lh) *lh = L.h;
    return CORB_OK;
}

// Eager (non-graph) replay of the launches of one extraction on the image currently in level 0, with a CUDA event
// between consecutive launches: per-launch device time for the roofline report. Order: resize 1..L-1, blur,
// fast_cells per level, quadtree per level, orient_desc (the graph runs the per-level launches concurrently).
static int profile_launch_count(const corb_orb* h) { return (h->nlevels - 1) + 1 + 2 * h->nlevels + 1; }

int corb_orb_profile(corb_orb* h, int reps, float* ms, int cap, int* n) {
    CORB_CHECK(h && h->plan_w && ms && n && reps >= 1, CORB_ERR_INVALID, "bad argument or no plan");
    const OrbGeom& g = h->geom;
    const int nk = profile_launch_count(h);
    CORB_CHECK(cap >= nk, CORB_ERR_INVALID, "need room for %d launches", nk);
    CORB_CUDA(cudaSetDevice(h->device));
    {
        int rc = settle(h);
        if (rc != CORB_OK) return rc;
    }
    std::vector<cudaEvent_t> ev(nk + 1);
    for (auto& e : ev) CORB_CUDA(cudaEventCreate(&e));
    for (int i = 0; i < nk; i++) ms[i] = 0.f;
    for (int r = 0; r < reps; r++) {
        int k = 0;
        CORB_CUDA(cudaEventRecord(ev[k++], h->stream));
        for (int l = 1; l < g.n_levels; l++) {
            launch_resize(g, h->buf, l, h->stream);
            CORB_CUDA(cudaEventRecord(ev[k++], h->stream));
        }
        launch_blur(g, h->buf, h->stream);
        CORB_CUDA(cudaEventRecord(ev[k++], h->stream));
        for (int l = 0; l < g.n_levels; l++) {
            launch_fast_cells(g, h->buf, l, h->stream);
            CORB_CUDA(cudaEventRecord(ev[k++], h->stream));
        }
        for (int l = 0; l < g.n_levels; l++) {
            launch_octtree(g, h->buf, l, h->key_smem_cap, h->oct_smem, h->stream);
            CORB_CUDA(cudaEventRecord(ev[k++], h->stream));
        }
        launch_orient_desc(g, h->buf, h->stream);
        CORB_CUDA(cudaEventRecord(ev[k++], h->stream));
        CORB_CUDA(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < nk; i++) {
            float t = 0.f;
            CORB_CUDA(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
            ms[i] += t / reps;
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    *n = nk;
    return CORB_OK;
}

const char* corb_orb_kernel_name(const corb_orb* h, int i) {
    if (!h) return "";
    static thread_local char name[32];
    const int nr = h->nlevels - 1, L = h->nlevels;
    if (i < nr) { snprintf(name, sizeof(name), "k_resize[%d]", i + 1); return name; }
    i -= nr;
    if (i == 0) return "k_blur";
    i -= 1;
    if (i < L) { snprintf(name, sizeof(name), "k_fast_cells[%d]", i); return name; }
    i -= L;
    if (i < L) { snprintf(name, sizeof(name), "k_octtree[%d]", i); return name; }
    i -= L;
    return i == 0 ? "k_orient_desc" : "";
}

void* corb_orb_stream(const corb_orb* h) { return h ? (void*)h->stream : nullptr; }
int corb_orb_launches_per_extract(const corb_orb* h) { return h ? h->kernel_launches : 0; }
int corb_orb_uses_tma(const corb_orb* h) { return h && h->plan_w ? h->buf.use_tma : 0; }

int corb_orb_set_host_transfer(corb_orb* h, int mode) {
    CORB_CHECK(h && (mode == 0 || mode == 1), CORB_ERR_INVALID, "mode must be 0 (mapped reads) or 1 (copy engine)");
    CORB_CHECK(!h->pending && !h->pending_pair, CORB_ERR_INVALID, "wait for the submitted extraction first");
    if (h->h2d_node == (mode == 1)) return CORB_OK;
    CORB_CUDA(cudaSetDevice(h->device));
    int rc = settle(h);
    if (rc != CORB_OK) return rc;
    CORB_CUDA(cudaStreamSynchronize(h->stream));
    h->h2d_node = mode == 1;
    // the captured graphs embed the transfer form: drop them, they are rebuilt on the next call
    for (int v = 0; v < 3; v++) {
        if (h->pair_exec[v]) cudaGraphExecDestroy(h->pair_exec[v]), h->pair_exec[v] = nullptr;
        if (h->pair_graph[v]) cudaGraphDestroy(h->pair_graph[v]), h->pair_graph[v] = nullptr;
    }
    for (int v = 0; v < 2; v++) {
        if (h->graph_exec[v]) cudaGraphExecDestroy(h->graph_exec[v]), h->graph_exec[v] = nullptr;
        if (h->graph[v]) cudaGraphDestroy(h->graph[v]), h->graph[v] = nullptr;
    }
    return CORB_OK;
}
int corb_orb_host_transfer(const corb_orb* h) { return h && h->h2d_node ? 1 : 0; }

int corb_orb_tap(corb_orb* h, int what, int level, void* out, size_t out_bytes, int* n) {
    CORB_CHECK(h && h->plan_w && level >= 0 && level < h->nlevels, CORB_ERR_INVALID, "bad argument or no plan");
    CORB_CUDA(cudaSetDevice(h->device));
    {
        int rc = settle(h);
        if (rc != CORB_OK) return rc;
    }
    CORB_CUDA(cudaStreamSynchronize(h->stream));
    const LevelGeom& L = h->geom.lv[level];
    if (what == CORB_TAP_PYRAMID || what == CORB_TAP_BLURRED) {
        CORB_CHECK(out && out_bytes >= (size_t)L.w * L.h, CORB_ERR_INVALID, "output buffer too small");
        const uint8_t* src = (what == CORB_TAP_BLURRED ? h->buf.blur : h->buf.pyr) + L.img_off;
        CORB_CUDA(cudaMemcpy2D(out, L.w, src, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
        if (n) *n = L.w * L.h;
        return CORB_OK;
    }
    if (what == CORB_TAP_LEVEL_COUNT) {
        CORB_CHECK(n, CORB_ERR_INVALID, "n is NULL");
        CORB_CUDA(cudaMemcpy(n, h->buf.level_count + level, sizeof(int), cudaMemcpyDeviceToHost));
        return CORB_OK;
    }
    if (what == CORB_TAP_CANDIDATES) {
        CORB_CHECK(n, CORB_ERR_INVALID, "n is NULL");
        // the level's list holds the corners in arrival order; the order key restores the reference order
        int total = 0;
        CORB_CUDA(cudaMemcpy(&total, h->buf.level_cand_out + level, sizeof(int), cudaMemcpyDeviceToHost));
        CORB_CHECK(total >= 0 && total <= L.n_cols * L.n_rows * L.slot, CORB_ERR_CAPACITY, "candidate count %d out of range", total);
        *n = total;
        if (!out) return CORB_OK;
        CORB_CHECK(out_bytes >= (size_t)total * 12, CORB_ERR_INVALID, "output buffer too small for %d candidates", total);
        std::vector<uint32_t> xy(total), ro(total);
        if (total) {
            CORB_CUDA(cudaMemcpy(xy.data(), h->buf.cand_xy + L.cand_base, (size_t)total * 4, cudaMemcpyDeviceToHost));
            CORB_CUDA(cudaMemcpy(ro.data(), h->buf.cand_ro + L.cand_base, (size_t)total * 4, cudaMemcpyDeviceToHost));
        }
        std::vector<int> idx(total);
        for (int i = 0; i < total; i++) idx[i] = i;
        std::sort(idx.begin(), idx.end(), [&](int a_, int b_) { return (ro[a_] & 0xffffffu) > (ro[b_] & 0xffffffu); });
        int32_t* o = (int32_t*)out;
        for (int i : idx) {
            *o++ = xy[i] & 0xffff;
            *o++ = xy[i] >> 16;
            *o++ = ro[i] >> 24;
        }
        return CORB_OK;
    }
    set_error("unknown tap %d", what);
    return CORB_ERR_INVALID;
}

}  // extern "C"
