// Launch wrappers of the ORB front-end kernels (definitions in orb_kernels.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/corb_b200.h"
#include "orb_geom.h"

namespace corb {

// one tensor map per pyramid level (u8, {w, h}, row pitch), box = 80 x 66 bytes: the FAST cell ROI
struct TmaMaps {
    CUtensorMap m[kMaxLevels];
};

struct OrbBuffers {
    uint8_t* pyr;        // un-blurred pyramid, all levels, pitched
    uint8_t* blur;       // blurred pyramid, same layout
    const int* xofs;     // resize tables (levels >= 1)
    const short2* alpha;
    const int* yofs;
    const short2* beta;
    int* cell_count;     // [n_cells]
    uint32_t* cand_xy;   // [cand_total]  x | y << 16, relative to (16,16)
    uint8_t* cand_r;     // [cand_total]  FAST response
    int* cell_off;       // [n_cells]     quadtree scratch: exclusive prefix of cell_count per level
    uint32_t* key_xy;    // [cand_total]  quadtree scratch (used when a level's keys do not fit shared memory)
    uint8_t* key_r;      // [cand_total]
    uint16_t* key_node;  // [cand_total]
    uint2* lvl_kp;       // [kp_cap]      kept keypoints per level: .x = X | Y << 16 (level coords), .y = response
    int* level_count;    // [n_levels]
    int* level_cand;     // [n_levels]    number of candidates per level (tap)
    corb_keypoint* kps;  // [kp_cap]
    uint8_t* desc;       // [kp_cap * 32]
    int* count;          // [1]
    int* status;         // [1] device-side error flag (0 ok)
    const TmaMaps* tma_maps;  // host pointer (passed by value as a __grid_constant__ kernel parameter)
    int use_tma;
};

// encodes the per-level tensor maps; returns false when the driver entry point is unavailable
bool encode_tma_maps(const OrbGeom& g, uint8_t* pyr, TmaMaps* out);

int octtree_smem_bytes(const OrbGeom& g, int level, int key_smem_cap);

void launch_import(const OrbGeom& g, const OrbBuffers& b, const uint8_t* src, int stride, cudaStream_t s);
const void* import_kernel_ptr();
void launch_resize(const OrbGeom& g, const OrbBuffers& b, int level, cudaStream_t s);
void launch_fast_cells(const OrbGeom& g, const OrbBuffers& b, int level, cudaStream_t s);
void launch_blur(const OrbGeom& g, const OrbBuffers& b, cudaStream_t s);
cudaError_t prepare_octtree(const OrbGeom& g, int key_smem_cap, int* smem_bytes_out);
void launch_octtree(const OrbGeom& g, const OrbBuffers& b, int level, int key_smem_cap, int smem_bytes, cudaStream_t s);
void launch_orient_desc(const OrbGeom& g, const OrbBuffers& b, cudaStream_t s);

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
};

void launch_stereo(const OrbGeom& g, const StereoArgs& a, cudaStream_t s);

}  // namespace corb
