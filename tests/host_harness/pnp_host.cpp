/* TEST INFRASTRUCTURE ONLY: compiles the per-thread EPnP of the CUDA path (corb_slam_b200/csrc/pnp_core.cuh) for the
 * host, so the `not gpu` suite can check - without a GPU - that the device arithmetic agrees bit for bit with the
 * oracle. Nothing in the product links or loads this. Built by tests/test_pnp_cpu.py with g++ -ffp-contract=off (the
 * CUDA file is compiled with -fmad=false). */
#define CORB_HD
#include "../../corb_slam_b200/csrc/pnp_core.cuh"

#include <vector>

using namespace corb::pnp;

extern "C" {

/* EPnP over the points selected by `mask` (bit i of word i/32), or over the 4 indices in `list` when mask == NULL. */
double harness_epnp_pose(const float* p3d, const float* p2d, int n_pts, const int* list, const uint32_t* mask, double fu, double fv,
                         double uc, double vc, int stride, double* Rt) {
    std::vector<double> buf((size_t)WS_DOUBLES * stride);
    Ws ws{buf.data(), stride};
    PtSet s;
    s.p3d = p3d; s.p2d = p2d; s.mask = mask; s.n_words = (n_pts + 31) / 32;
    if (mask) {
        s.n = 0;
        for (int w = 0; w < s.n_words; w++) s.n += __builtin_popcount(mask[w]);
    } else {
        s.n = 4;
        for (int k = 0; k < 4; k++) s.list[k] = list[k];
    }
    Epnp e;
    e.fu = fu; e.fv = fv; e.uc = uc; e.vc = vc;
    return e.compute_pose(s, ws, Rt);
}

void harness_resolve_draws(int n, const int* r, int* list) { resolve_draws(n, r, list); }

int harness_count_inliers(const double* Rt, double fu, double fv, double uc, double vc, const float* p3d, const float* p2d,
                          const float* max_err, int n, uint8_t* inl) {
    int c = 0;
    for (int i = 0; i < n; i++) {
        inl[i] = is_inlier(Rt, fu, fv, uc, vc, p3d + 3 * i, p2d + 2 * i, max_err[i]);
        c += inl[i];
    }
    return c;
}
}
