/* Allocator environment of libref.so — TEST INFRASTRUCTURE ONLY (oracle/refbuild).
 *
 * ORBextractor::DistributeOctTree sorts its split candidates by (key count, ExtractorNode*) (ORBextractor.cc:684 with
 * the pairs built at :591,:627): nodes with equal counts are split in the order of their HEAP ADDRESSES, so the
 * reference's keypoints are a function of the allocator, not only of the image. With glibc
 * malloc the freed list nodes are recycled LIFO and the order is arbitrary (and differs between the two extractor
 * threads of Frame.cc:78-81). To compare the unmodified source with anything, the environment has to be fixed: while
 * `monotone` is on (default), std::list<ExtractorNode> nodes - recognised by their size - come from a per-thread bump
 * pool (only while a glue call that runs the extractor is in flight), so that address order == creation order, which is the canonical tie-break of the oracle and of the CUDA path
 * (DESIGN.md section 2). ref_set_monotone_nodes(0) restores plain malloc; tests/test_ref_cpu.py measures how many
 * keypoints that changes. Every other allocation is malloc / free. (-Bsymbolic binds libref's own calls to these.) */
#include <sys/mman.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <list>
#include <new>

#include "ORBextractor.h"

namespace {
const size_t kNodeBytes = sizeof(std::_List_node<ORB_SLAM2::ExtractorNode>);
const size_t kPoolBytes = (size_t)1 << 30; /* address space only (MAP_NORESERVE) */
std::atomic<int> g_monotone(1);
std::atomic<int> g_active(0); /* > 0 while a glue call that runs ORBextractor::operator() is in flight */
struct Pool {
    char* base;
    size_t off, live;
    ~Pool() { if (base) munmap(base, kPoolBytes); }
};
thread_local Pool t_pool = {0, 0, 0};

inline void* pool_alloc() {
    Pool& p = t_pool;
    if (!p.base) {
        void* m = mmap(0, kPoolBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) { std::fprintf(stderr, "ref_alloc: mmap failed\n"); std::abort(); }
        p.base = (char*)m;
    }
    const size_t sz = (kNodeBytes + 15) & ~(size_t)15;
    if (p.off + sz > kPoolBytes) { std::fprintf(stderr, "ref_alloc: node pool exhausted\n"); std::abort(); }
    void* r = p.base + p.off;
    p.off += sz;
    p.live++;
    return r;
}
inline bool pool_free(void* ptr) {
    Pool& p = t_pool;
    if (!p.base || (char*)ptr < p.base || (char*)ptr >= p.base + kPoolBytes) return false;
    if (--p.live == 0) p.off = 0; /* every node of a DistributeOctTree call is gone: start over */
    return true;
}
} // namespace

void* operator new(size_t n) {
    if (n == kNodeBytes && g_active.load(std::memory_order_relaxed) > 0 && g_monotone.load(std::memory_order_relaxed)) return pool_alloc();
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](size_t n) {
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete(void* p) noexcept { if (p && !pool_free(p)) std::free(p); }
void operator delete(void* p, size_t) noexcept { if (p && !pool_free(p)) std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete[](void* p, size_t) noexcept { std::free(p); }

extern "C" void ref_set_monotone_nodes(int on) { g_monotone.store(on ? 1 : 0); }
/* The pool is only live inside extractor calls: a block of the node size allocated here but released inside libstdc++.so
 * (which keeps its own operator delete when libref is dlopen()ed RTLD_LOCAL) must never come from the pool. */
extern "C" void ref_alloc_scope(int delta) { g_active.fetch_add(delta); }
