// Per-(handle, image size) geometry of the ORB front end, passed by value to every kernel.
#pragma once
#include <stdint.h>

namespace corb {

constexpr int kMaxLevels = 16;
constexpr int kEdge = 19;        // EDGE_THRESHOLD  (ORBextractor.cc:74)
constexpr int kBorder = kEdge - 3;  // minBorderX/Y = 16 (:775-776)
constexpr int kHalfPatch = 15;   // HALF_PATCH_SIZE (:73)
constexpr int kPatch = 31;
constexpr int kBlurTW = 64, kBlurTH = 32;  // output tile of the Gaussian blur launch (host tile counts and the kernel agree on it)       // PATCH_SIZE      (:72)
constexpr int kCellRoiMax = 66;  // wCell + 6 < 60 + 6 (W = 30 => wCell = ceil(width / floor(width/30)) < 60)

struct LevelGeom {
    int level;               // index of this level
    int w, h, pitch;         // level size, bytes per row (multiple of 128)
    int img_off;             // byte offset of the level inside the pyramid / blurred buffers
    int max_bx, max_by;      // maxBorderX/Y = w-16, h-16
    int n_cols, n_rows, w_cell, h_cell;
    int cell_base;           // first cell of this level in the flat cell list
    int slot;                // per-cell candidate capacity = ceil(w_cell/2)*ceil(h_cell/2)
    int cand_base;           // first candidate slot of this level (= sum of n_cells*slot of lower levels)
    int quota;               // mnFeaturesPerLevel[level]
    int n_ini;               // initial quadtree nodes
    float h_x;               // (float)width / n_ini
    int node_cap;            // max(quota + 3, 4 * n_ini)
    int lut_off;             // quadtree path tables of this level inside OrbBuffers::oct_lut: xs[w - 32] then ys[h - 32]
    int kp_base;             // first per-level keypoint slot (= sum of node_cap of lower levels)
    int xtab_off, ytab_off;  // offsets into the resize tables (level >= 1)
    int blur_tile_base;      // first tile of this level in the blur launch
    int blur_tiles_x;
    float scale;             // mvScaleFactor[level]
    float size;              // (float)(int)(31 * scale)
};

struct OrbGeom {
    int n_levels;
    int ini_th, min_th;
    int n_cells;      // total over levels
    int kp_cap;       // total over levels of node_cap
    int blur_tiles;   // total over levels
    LevelGeom lv[kMaxLevels];
};

}  // namespace corb
