// Probe: may the rows of a TMA tensor map OVERLAP in memory (stride of dimension 1 smaller than the extent of dimension 0)?
// Needed for a patch-matrix-free first conv layer: a 3x3 stride-2 conv on a C=4 image reads, per output pixel xo and tap row dy,
// the 4 pixels 2xo .. 2xo+3 = 16 halves = 32 contiguous bytes, and consecutive xo are only 16 bytes (2 pixels) apart.
// Image [H=12][W=24][C=4] u16, value = y*1000 + x*10 + c.  Map: dims {16, Wo=10, H=12}, strides {16 B, W*8 B}, box {16, 4, 3 rows via
// traversal stride 2}, no swizzle and 32-byte swizzle.  Prints, per delivered 32-byte row, the (y, x) of its first and last pixel.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, uint16_t* out, int nbytes, int x0, int y0) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), s = (uint32_t)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2048; ++i) ((uint16_t*)sm)[i] = 0xFFFF;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("fence.proxy.async.shared::cta;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(nbytes));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(s), "l"((uint64_t)&tm), "r"(0), "r"(x0), "r"(y0), "r"(b) : "memory");
        uint32_t ok = 0; int spins = 0;
        while (!ok && spins < 2000000) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(b));
            ++spins;
        }
        out[2048] = ok;
        for (int i = 0; i < 2048; ++i) out[i] = ((uint16_t*)sm)[i];
    }
}
int main() {
    const int H = 12, W = 24, C = 4, Wo = 10;
    std::vector<uint16_t> h(H * W * C);
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < C; ++c) h[(y * W + x) * C + c] = y * 1000 + x * 10 + c;
    uint16_t *d, *o; cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 2049 * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    void* fp; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    PFN enc = (PFN)fp;
    CUtensorMapSwizzle sws[2] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_32B};
    for (int si = 0; si < 2; ++si) {
        CUtensorMap tm; cuuint64_t dims[3] = {16, Wo, H}; cuuint64_t str[2] = {16, (cuuint64_t)W * C * 2}; cuuint32_t box[3] = {16, 4, 6}; cuuint32_t es[3] = {1, 1, 2};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sws[si], CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("swizzle %s: overlapping rows (stride 16 B < extent 32 B): encode rc=%d\n", si ? "32B" : "none", (int)r);
        if (r) continue;
        for (int t = 0; t < 2; ++t) {
            const int x0 = t ? 7 : 2, y0 = t ? 9 : 1;       // second case runs off the right / bottom edge (zero fill expected)
            cudaMemset(o, 0, 2049 * 2);
            k<<<1, 32, 8192>>>(tm, o, 4 * 3 * 32, x0, y0);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<uint16_t> g(2049); cudaMemcpy(g.data(), o, 2049 * 2, cudaMemcpyDeviceToHost);
            printf("  box at (xo=%d, y=%d): err=%d completed=%d\n", x0, y0, (int)e, (int)g[2048]);
            for (int row = 0; row < 12; ++row) {
                printf("    smem row %2d:", row);
                for (int ch = 0; ch < 2; ++ch) {          // the two 16-byte chunks of the row, as stored
                    const uint16_t a = g[row * 16 + ch * 8], bq = g[row * 16 + ch * 8 + 4];
                    if (a == 0xFFFF) printf(" [untouched]"); else printf(" px(y=%d,x=%d)+px(y=%d,x=%d)", a / 1000, (a % 1000) / 10, bq / 1000, (bq % 1000) / 10);
                }
                printf("\n");
            }
            if (e) { cudaDeviceReset(); return 0; }
        }
    }
    return 0;
}
