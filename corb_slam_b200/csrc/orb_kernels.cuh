// Launch wrappers of the ORB front-end kernels (definitions in orb_kernels.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/corb_b200.h"
#include "orb_geom.h"

namespace corb {

// one tensor map per pyramid level (u8, {w, h}, row pitch), box = 80 x 66 bytes: the FAST cell ROI
struct TmaMaps {
    CUtensorMap m[kMaxLevels];   // FAST: one 96 x 66 box per grid cell
    CUtensorMap mb[kMaxLevels];  // blur: one 96 x 22 box per 64 x 16 output tile
};

// second image of a pair for k_fast_cells
struct FastImage2 {
    const uint8_t* pyr;
    int* cell_count;
    uint32_t* cand_xy;
    uint32_t* cand_ro;
    int* level_cand;
    int* status;
    const CUtensorMap* tmap_dev;
};

// Plan of the fused pyramid kernel (k_pyramid): level 0 is cut into tw x th tiles; per level and axis the table holds,
// for every tile, the first pixel it owns (n_tiles + 1 entries; ownership = where the chain of left/top source pixels
// ends up in level 0, a monotone map, so the owned ranges partition the level) and the inclusive range it must
// compute (owned pixels plus everything deeper levels of the same tile read).
struct PyrPlan {
    int base;                                 // source level (0 for the whole pyramid)
    int tw, th, ntx, nty;
    int buf_bytes;                            // one shared-memory level buffer
    int tab_smem_ints;                        // staged resize-table slices (all levels)
    const int* tab;                           // device copy of the table
    int xoff[kMaxLevels], yoff[kMaxLevels];   // start of level l's x / y section: own[nt + 1] | need0[nt] | need1[nt]
};

struct OrbBuffers {
    uint8_t* pyr;        // un-blurred pyramid, all levels, pitched
    uint8_t* blur;       // blurred pyramid, same layout
    const int* xofs;     // resize tables (levels >= 1)
    const short2* alpha;
    const int* yofs;
    const short2* beta;
    int* cell_count;     // [n_cells]
    // FAST candidates: every cell appends its corners to its level's list (space reserved with one atomicAdd on
    // level_cand[l]), so the list order varies from run to run; the reference order (cell row, cell column, raster,
    // ORBextractor.cc:789-832) travels with each record as an order key and is all the quadtree needs of it.
    uint32_t* cand_xy;   // [cand_total]  x | y << 16, relative to (16,16); level l's list starts at lv[l].cand_base
    uint32_t* cand_ro;   // [cand_total]  response << 24 | (0xffffff - order), order = cell * slot + raster index in the cell
    int* level_cand;     // [n_levels]    list length; counted up by FAST, read and reset to 0 by the quadtree
    int* level_cand_out; // [n_levels]    candidates per level of the last frame (tap)
    const uint16_t* oct_lut;  // quadtree path tables (see LevelGeom::lut_off)
    uint16_t* key_node;  // [cand_total]  quadtree scratch (used when a level's keys do not fit shared memory)
    uint2* lvl_kp;       // [kp_cap]      kept keypoints per level: .x = X | Y << 16 (level coords), .y = response
    int* level_count;    // [n_levels]
    corb_keypoint* kps;  // [kp_cap]
    uint8_t* desc;       // [kp_cap * 32]
    int* count;          // [1]
    int* status;         // [1] device-side error flag (0 ok)
    const TmaMaps* tma_maps;  // host pointer (passed by value as a __grid_constant__ kernel parameter)
    int use_tma;
    int oct_fast;        // 1: k_octtree may take its closed-form path (CORB_OCT_GENERIC=1 forces the pass-by-pass code)
    const CUtensorMap* tma_dev;  // device copy of the per-level tensor maps (all-level FAST launch)
    PyrPlan pyr_plan;    // whole pyramid from level 0 (CORB_GRAPH=hybrid|fused)
    PyrPlan pyr_tail;    // levels pyr_tail.base + 1 .. from level pyr_tail.base: the short end of the resize chain in one launch
};

// encodes the per-level tensor maps; returns false when the driver entry point is unavailable
bool encode_tma_maps(const OrbGeom& g, uint8_t* pyr, TmaMaps* out);

int octtree_smem_bytes(const OrbGeom& g, int level, int key_smem_cap);
int oct_lut_entries_host(const LevelGeom& L);
void build_oct_lut(const LevelGeom& L, uint16_t* out);

// Every launch takes an optional second image (`b1`, the other handle of a stereo pair with the same geometry): the
// grid gets z = 2 and block z works on b1, so a stereo frame costs one set of launches instead of two.
void launch_import(const OrbGeom& g, const OrbBuffers& b, const uint8_t* src, int stride, cudaStream_t s,
                   const OrbBuffers* b1 = nullptr, const uint8_t* src1 = nullptr, int stride1 = 0, bool host_src = false);
const void* import_kernel_ptr(bool host_src);
void import_launch_dims(int w, int h, int n_img, bool host_src, dim3* grid, dim3* block);
void launch_resize(const OrbGeom& g, const OrbBuffers& b, int level, cudaStream_t s, const OrbBuffers* b1 = nullptr);
// host side of PyrPlan: fills everything but `tab` (returned in tab_host for the caller to upload)
void build_pyr_plan(const OrbGeom& g, const int* xofs, const int* yofs, int base, PyrPlan* plan, std::vector<int>* tab_host);
cudaError_t prepare_pyramid(const PyrPlan& p);
void launch_pyramid(const OrbGeom& g, const OrbBuffers& b, cudaStream_t s, const OrbBuffers* b1 = nullptr, bool tail = false);
void launch_fast_all(const OrbGeom& g, const OrbBuffers& b, cudaStream_t s, const OrbBuffers* b1 = nullptr);
void launch_fast_cells(const OrbGeom& g, const OrbBuffers& b, int level, cudaStream_t s, const OrbBuffers* b1 = nullptr);
void launch_blur(const OrbGeom& g, const OrbBuffers& b, cudaStream_t s, const OrbBuffers* b1 = nullptr);
cudaError_t prepare_octtree(const OrbGeom& g, int key_smem_cap, int* smem_bytes_out);
void launch_octtree(const OrbGeom& g, const OrbBuffers& b, int level, int key_smem_cap, int smem_bytes, cudaStream_t s,
                    const OrbBuffers* b1 = nullptr);
void launch_orient_desc(const OrbGeom& g, const OrbBuffers& b, cudaStream_t s, const OrbBuffers* b1 = nullptr);

// Frame::ComputeStereoMatches on the device-resident results of two extractor handles (stereo.cu)
struct StereoArgs {
    const corb_keypoint *kl, *kr;
    const uint4 *dl, *dr;
    const int *nl, *nr;          // device counts of the two extractions
    const uint8_t *pyr_l, *pyr_r;
    float scale[kMaxLevels], inv_scale[kMaxLevels];
    float mbf, mb;
    float *u_right, *depth;      // [kp_cap]
    int* best_dist;              // [kp_cap] SAD of the accepted match, -1 if none
    int n_rows;
    // right keypoints bucketed by 8-row bands (a keypoint sits in every band its row range touches): the coarse form of
    // the reference's vRowIndices (Frame.cc:487-497); the exact row test is applied per candidate
    int4* rinfo;                 // [kp_cap] (min row, max row, octave, x bits) per right keypoint
    int* band_cnt;               // [n_bands] entries per band; zeroed by k_stereo_outliers for the next frame
    int* band_list;              // [n_bands][band_cap] right keypoint indices
    int n_bands, band_cap;
};

void launch_stereo(const OrbGeom& g, const StereoArgs& a, cudaStream_t s);

}  // namespace corb
