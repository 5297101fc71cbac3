// Stand-alone probe: which (box, coordinate) combinations does a u8 2-D TMA tile load accept on sm_100a?
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BW, int BH>
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x, int y, uint8_t* out) {
    __shared__ __align__(128) uint8_t tile[BW * BH];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW * BH) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         smem_u32(tile)),
                     "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(x), "r"(y), "r"(smem_u32(&bar))
                     : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}
template <int BW, int BH>
int run(int W, int H, int pitch, int x, int y) {
    std::vector<uint8_t> img((size_t)pitch * H);
    for (int r = 0; r < H; r++) for (int c = 0; c < pitch; c++) img[(size_t)r * pitch + c] = (uint8_t)((r * 7 + c * 3 + 1) & 0xff);
    uint8_t *d, *o;
    cudaMalloc(&d, img.size()); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&o, BW * BH);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("box %dx%d: encode failed %d\n", BW, BH, (int)r); return 1; }
    k<BW, BH><<<1, 128>>>(m, x, y, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("box %dx%d at (%d,%d): %s\n", BW, BH, x, y, cudaGetErrorString(e)); return 2; }
    std::vector<uint8_t> got(BW * BH); cudaMemcpy(got.data(), o, BW * BH, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < BH; r2++) for (int c = 0; c < BW; c++) {
        int gx = x + c, gy = y + r2;
        uint8_t exp = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? img[(size_t)gy * pitch + gx] : 0;
        bad += got[r2 * BW + c] != exp;
    }
    printf("box %dx%d at (%d,%d): ok, %d mismatches\n", BW, BH, x, y, bad);
    return 0;
}
int main(int argc, char** argv) {
    int variant = atoi(argv[1]), x = atoi(argv[2]), y = atoi(argv[3]);
    if (variant == 0) return run<80, 66>(1242, 375, 1280, x, y);
    if (variant == 1) return run<64, 64>(1242, 375, 1280, x, y);
    if (variant == 2) return run<80, 64>(1242, 375, 1280, x, y);
    if (variant == 3) return run<128, 66>(1242, 375, 1280, x, y);
    if (variant == 4) return run<96, 66>(1242, 375, 1280, x, y);
    return 9;
}
