// Shared by the shim/*.cc files: the C++ bodies a CORB-SLAM maintainer compiles INSTEAD of the hot-path function bodies of the
// reference. They include the reference's own headers, keep every class interface, and call the C ABI of
// include/corb_b200.h. No CPU fallback: a failing call aborts (the reference has no error path at these seams either).
#ifndef CORB_SHIM_COMMON_H
#define CORB_SHIM_COMMON_H
#include <stdio.h>
#include <stdlib.h>

#include "corb_b200.h"

namespace corb_shim {

inline void check(int rc, const char* what) {
    if (rc != CORB_OK) {
        fprintf(stderr, "corb_b200 shim: %s failed: %s\n", what, corb_last_error());
        abort();
    }
}
// one corbslam_client <-> one GPU: device = clientId - 1, exported by the launcher as CORB_DEVICE (default 0)
inline int device() {
    const char* d = getenv("CORB_DEVICE");
    return d ? atoi(d) : 0;
}
// Tracking, LoopClosing and the server's fuse thread call the matchers concurrently: one handle (stream + staging) per thread
inline corb_matcher* matcher() {
    static thread_local corb_matcher* m = nullptr;
    if (!m) check(corb_matcher_create(device(), &m), "corb_matcher_create");
    return m;
}

}  // namespace corb_shim
#endif
