// What corbslam_server's global-BA thread does with more than one GPU attached (GlobalOptimize.cpp:399,435-547 runs
// Optimizer::GlobalBundleAdjustemnt on ONE detached thread of ONE process): one process, N threads, N ncclComm_t, every thread
// owns one GPU and one landmark shard (landmark l -> rank l % N, poses replicated) and calls corb_ba_solve with an all-reduce
// hook that is a plain ncclAllReduce on the solver's stream - the hook of INTEGRATION.md section 4, in C.
//
//   ba_nccl N problem.bin result.bin [iterations]
// problem.bin (written by tests / bench.py): int32 P, L, E, then pose_q[4P] pose_t[3P] (f64), pose_fixed[P] (u8), pose_cam[5P] (f64),
// point_xyz[3L] (f64), point_fixed[L] (u8), edge_pose[E] edge_point[E] (i32), edge_obs[3E] edge_inv_sigma2[E] (f64).
// result.bin: int32 n_trials, iterations; f64 ms_total (max over ranks), chi2_initial, chi2_final; u8 trial_accepted[n_trials];
// pose_q[4P] pose_t[3P] of rank 0; point_xyz[3L]; f64 rank_spread = max |pose of rank r - pose of rank 0|.
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <thread>
#include <vector>

#include "corb_b200.h"

struct Problem {
    int32_t P, L, E;
    std::vector<double> pose_q, pose_t, pose_cam, point_xyz, edge_obs, edge_info;
    std::vector<uint8_t> pose_fixed, point_fixed;
    std::vector<int32_t> edge_pose, edge_point;
};

template <typename T> static bool rd(FILE* f, std::vector<T>& v, size_t n) { v.resize(n); return n == 0 || fread(v.data(), sizeof(T), n, f) == n; }

static int allreduce_hook(void* user, double* d_buf, size_t n, int op, void* stream) {
    ncclComm_t comm = *(ncclComm_t*)user;
    const ncclRedOp_t o = op == 0 ? ncclSum : op == 1 ? ncclMin : ncclMax;
    return ncclAllReduce(d_buf, d_buf, n, ncclDouble, o, comm, (cudaStream_t)stream) == ncclSuccess ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: ba_nccl N problem.bin result.bin [iterations]\n"); return 2; }
    const int N = atoi(argv[1]), iters = argc > 4 ? atoi(argv[4]) : 10;
    Problem pr;
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(&pr.P, 4, 3, f) != 3) { fprintf(stderr, "cannot read %s\n", argv[2]); return 2; }
    const size_t P = pr.P, L = pr.L, E = pr.E;
    if (!rd(f, pr.pose_q, 4 * P) || !rd(f, pr.pose_t, 3 * P) || !rd(f, pr.pose_fixed, P) || !rd(f, pr.pose_cam, 5 * P) || !rd(f, pr.point_xyz, 3 * L) ||
        !rd(f, pr.point_fixed, L) || !rd(f, pr.edge_pose, E) || !rd(f, pr.edge_point, E) || !rd(f, pr.edge_obs, 3 * E) || !rd(f, pr.edge_info, E)) {
        fprintf(stderr, "short problem file\n");
        return 2;
    }
    fclose(f);
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (ndev < N) { fprintf(stderr, "needs %d GPUs, %d visible\n", N, ndev); return 3; }
    std::vector<ncclComm_t> comms(N);
    std::vector<int> devs(N);
    for (int i = 0; i < N; i++) devs[i] = i;
    if (N > 1 && ncclCommInitAll(comms.data(), N, devs.data()) != ncclSuccess) { fprintf(stderr, "ncclCommInitAll failed\n"); return 3; }

    struct Shard { Problem p; std::vector<int32_t> ids; corb_ba_result res; int rc; double ms; };
    std::vector<Shard> sh(N);
    for (int r = 0; r < N; r++) {  // landmark l -> rank l % N with its edges (edges stay grouped by landmark)
        Problem& s = sh[r].p;
        s.P = pr.P;
        s.pose_q = pr.pose_q; s.pose_t = pr.pose_t; s.pose_fixed = pr.pose_fixed; s.pose_cam = pr.pose_cam;
        std::vector<int32_t> remap(L, -1);
        for (size_t l = r; l < L; l += N) {
            remap[l] = (int32_t)sh[r].ids.size();
            sh[r].ids.push_back((int32_t)l);
            s.point_xyz.insert(s.point_xyz.end(), &pr.point_xyz[3 * l], &pr.point_xyz[3 * l] + 3);
            s.point_fixed.push_back(pr.point_fixed[l]);
        }
        for (size_t e = 0; e < E; e++) {
            const int32_t m = remap[pr.edge_point[e]];
            if (m < 0) continue;
            s.edge_pose.push_back(pr.edge_pose[e]);
            s.edge_point.push_back(m);
            s.edge_obs.insert(s.edge_obs.end(), &pr.edge_obs[3 * e], &pr.edge_obs[3 * e] + 3);
            s.edge_info.push_back(pr.edge_info[e]);
        }
        s.L = (int32_t)s.point_fixed.size();
        s.E = (int32_t)s.edge_pose.size();
    }
    auto run = [&](int r, int its, bool keep) {
        Shard& s = sh[r];
        Problem q = s.p;  // corb_ba_solve updates poses / points in place
        corb_ba_problem cp = {q.P, q.L, q.E, q.pose_q.data(), q.pose_t.data(), q.pose_fixed.data(), q.pose_cam.data(), q.point_xyz.data(),
                              q.point_fixed.data(), q.edge_pose.data(), q.edge_point.data(), q.edge_obs.data(), q.edge_info.data()};
        cudaSetDevice(r);
        const auto t0 = std::chrono::steady_clock::now();
        s.rc = corb_ba_solve(&cp, its, nullptr, 0, r, &s.res, N > 1 ? allreduce_hook : nullptr, N > 1 ? (void*)&comms[r] : nullptr);
        s.ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (keep) { s.p.pose_q = q.pose_q; s.p.pose_t = q.pose_t; s.p.point_xyz = q.point_xyz; }
    };
    auto all = [&](int its, bool keep) {
        std::vector<std::thread> th;
        for (int r = 0; r < N; r++) th.emplace_back(run, r, its, keep);
        for (auto& t : th) t.join();
    };
    all(1, false);  // warm-up: contexts, NCCL channels, the per-device arenas at this size
    all(iters, true);
    double ms = 0, spread = 0;
    for (int r = 0; r < N; r++) {
        if (sh[r].rc != CORB_OK) { fprintf(stderr, "rank %d: corb_ba_solve failed (%d): %s\n", r, sh[r].rc, corb_last_error()); return 4; }
        ms = std::max(ms, sh[r].ms);
        for (size_t i = 0; i < 4 * P; i++) spread = std::max(spread, fabs(sh[r].p.pose_q[i] - sh[0].p.pose_q[i]));
        for (size_t i = 0; i < 3 * P; i++) spread = std::max(spread, fabs(sh[r].p.pose_t[i] - sh[0].p.pose_t[i]));
    }
    std::vector<double> pts(3 * L);
    for (int r = 0; r < N; r++)
        for (size_t i = 0; i < sh[r].ids.size(); i++) memcpy(&pts[3 * (size_t)sh[r].ids[i]], &sh[r].p.point_xyz[3 * i], 24);
    const corb_ba_result& R = sh[0].res;
    FILE* o = fopen(argv[3], "wb");
    if (!o) return 2;
    const int32_t hdr[2] = {R.n_trials, R.iterations};
    const double dh[3] = {ms, R.chi2_initial, R.chi2_final};
    fwrite(hdr, 4, 2, o); fwrite(dh, 8, 3, o); fwrite(R.trial_accepted, 1, std::min(R.n_trials, 256), o);
    fwrite(sh[0].p.pose_q.data(), 8, 4 * P, o); fwrite(sh[0].p.pose_t.data(), 8, 3 * P, o); fwrite(pts.data(), 8, 3 * L, o);
    fwrite(&spread, 8, 1, o);
    fclose(o);
    printf("ba_nccl: %d thread(s) / GPU(s), P=%zu L=%zu E=%zu, %d iterations, %d trials, %.2f ms (max over ranks), chi2 %.6f -> %.6f, rank spread %.3g\n",
           N, P, L, E, R.iterations, R.n_trials, ms, R.chi2_initial, R.chi2_final, spread);
    for (int r = 0; r < N && N > 1; r++) ncclCommDestroy(comms[r]);
    return 0;
}
