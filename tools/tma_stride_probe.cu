// Probe: semantics of cuTensorMapEncodeTiled elementStrides (traversal stride) on B200.
// Tensor [H=12][W=20][C=16] u16, value = h*1000 + w*10 + (c>0). Box variants loaded at (c0,w0,h0)=(0,1,1); prints the (h,w) each smem pixel got.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, uint16_t* out, int nbytes, int w0, int h0) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), s = (uint32_t)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4096; ++i) ((uint16_t*)sm)[i] = 0xFFFF;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("fence.proxy.async.shared::cta;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(nbytes));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(s), "l"((uint64_t)&tm), "r"(0), "r"(w0), "r"(h0), "r"(b) : "memory");
        uint32_t ok = 0; int spins = 0;
        while (!ok && spins < 2000000) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(b));
            ++spins;
        }
        out[4096] = ok;
        for (int i = 0; i < 4096; ++i) out[i] = ((uint16_t*)sm)[i];
    }
}
int main() {
    const int H = 12, W = 20, C = 16;
    std::vector<uint16_t> h(H * W * C);
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < C; ++c) h[(y * W + x) * C + c] = y * 1000 + x * 10 + (c > 0);
    uint16_t *d, *o; cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 4097 * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    void* fp; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    PFN enc = (PFN)fp;
    struct V { unsigned bw, bh, sw, sh; int w0, h0; } vs[] = {{8, 4, 2, 2, 1, 1}, {16, 8, 2, 2, 1, 1}, {15, 7, 2, 2, 1, 1}, {8, 4, 2, 2, -1, -1}, {8, 4, 1, 1, 1, 1}};
    for (auto v : vs) {
        CUtensorMap tm; cuuint64_t dims[3] = {C, W, H}; cuuint64_t str[2] = {C * 2, (cuuint64_t)W * C * 2}; cuuint32_t box[3] = {C, v.bw, v.bh}; cuuint32_t es[3] = {1, v.sw, v.sh};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box(w=%u,h=%u) stride(%u,%u) at (w0=%d,h0=%d): encode rc=%d\n", v.bw, v.bh, v.sw, v.sh, v.w0, v.h0, (int)r);
        if (r) continue;
        // try both candidate byte counts
        for (int mode = 0; mode < 2; ++mode) {
            int npix = mode == 0 ? v.bw * v.bh : ((v.bw + v.sw - 1) / v.sw) * ((v.bh + v.sh - 1) / v.sh);
            if (mode == 1 && npix == (int)(v.bw * v.bh)) continue;
            cudaMemset(o, 0, 4097 * 2);
            k<<<1, 32, 16384>>>(tm, o, npix * C * 2, v.w0, v.h0);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<uint16_t> g(4097); cudaMemcpy(g.data(), o, 4097 * 2, cudaMemcpyDeviceToHost);
            printf("  expect_tx=%d px -> err=%d completed=%d ; first pixels (h,w): ", npix, (int)e, (int)g[4096]);
            for (int p = 0; p < 40; ++p) { uint16_t val = g[p * C]; if (val == 0xFFFF) { printf("[--] "); continue; } printf("(%d,%d)%s ", val / 1000, (val % 1000) / 10, g[p * C + 1] == val + 1 ? "" : "!"); }
            printf("\n");
            if (e) { cudaDeviceReset(); return 0; }
        }
    }
    return 0;
}
