// Replacement body of PnPsolver::iterate (corbslam_client/src/PnPsolver.cc:210-300); constructors, SetRansacParameters, find()
// and the class stay the reference's (the EPnP members :420-962 are simply no longer called). A maintainer deletes the
// iterate() body from PnPsolver.cc and adds this file.
//
// The library call is stateless and evaluates every RANSAC hypothesis at once, so the RandomInt values are drawn up front,
// exactly like the loop body would draw them (:236-238): 4 per iteration from N, N-1, N-2, N-3 available indices. They are
// kept per solver object (a file-static map: the header is untouched).
#define private public  // mnIterations, mvP2D, ... are private members of the reference class; the header stays byte-identical
#include "PnPsolver.h"
#undef private

#include <map>
#include <mutex>
#include <vector>

#include "Thirdparty/DBoW2/DUtils/Random.h"
#include "shim_common.h"

using namespace std;

namespace ORB_SLAM2 {

namespace {
std::mutex g_mu;
std::map<const PnPsolver*, vector<int32_t> > g_draws;
}

cv::Mat PnPsolver::iterate(int nIterations, bool& bNoMore, vector<bool>& vbInliers, int& nInliers) {
    bNoMore = false;
    vbInliers.clear();
    nInliers = 0;
    vector<int32_t>* draws;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        draws = &g_draws[this];
    }
    // a solver that has not iterated yet starts a new stream: the destructor (reference code) cannot drop the map entry, and a
    // new object may live at the address of a destroyed one
    if (mnIterations == 0) draws->clear();
    corb_pnp_problem p;
    p.n = N;
    static const float zero3[3] = {0, 0, 0};
    p.p2d = N ? &mvP2D[0].x : zero3;
    p.p3d = N ? &mvP3Dw[0].x : zero3;
    p.max_err = N ? mvMaxError.data() : zero3;
    p.fx = (float)fu; p.fy = (float)fv; p.cx = (float)uc; p.cy = (float)vc;
    p.min_inliers = mRansacMinInliers;
    p.max_its = mRansacMaxIts;
    p.iterations_done = mnIterations;
    p.n_iterations = nIterations;
    if (N >= mRansacMinInliers) {  // :219-223: otherwise the loop is never entered and no number is drawn
        const int itEnd = max(mRansacMaxIts, mnIterations + nIterations);
        for (int it = (int)draws->size() / 4; it < itEnd; ++it)
            for (int k = 0; k < 4; ++k) draws->push_back(DUtils::Random::RandomInt(0, N - k - 1));
    }
    static const int32_t none[4] = {0, 0, 0, 0};
    p.draws = draws->empty() ? none : draws->data();
    corb_pnp_result r;
    vector<uint8_t> inl(max(N, 1));
    uint8_t* ip = inl.data();
    corb_shim::check(corb_pnp_iterate_batch(corb_shim::matcher(), 1, &p, &r, &ip), "corb_pnp_iterate_batch");
    mnIterations = r.iterations;
    bNoMore = r.no_more != 0;
    if (!r.status) return cv::Mat();
    nInliers = r.n_inliers;
    vbInliers = vector<bool>(mvpMapPointMatches.size(), false);
    for (int i = 0; i < N; i++)
        if (inl[i]) vbInliers[mvKeyPointIndices[i]] = true;  // :264-270, :282-288
    return cv::Mat(4, 4, CV_32F, (void*)r.Tcw).clone();
}

}  // namespace ORB_SLAM2
