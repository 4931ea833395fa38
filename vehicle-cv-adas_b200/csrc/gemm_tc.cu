// gemm_tc.cu -- the hot op: fused conv(+folded BN)+bias+SiLU/ReLU(+residual) as a row-shifted
// implicit GEMM on the 5th-gen tensor cores (tcgen05.mma, fp16 x fp16 -> fp32 in TMEM), operands
// staged by TMA (cp.async.bulk.tensor.2d, 128B swizzle) through an mbarrier ring.
//
// Replaces: the opaque conv stacks ONNXRuntime/TensorRT execute behind
//   coreEngine.py:150-157 (TensorRTEngine.engine_inference) / :184-186 (OnnxEngine.engine_inference).
//
// Tile: BM = 128 output rows (pixels of the padded NHWC grid) x BN output channels (runtime,
// multiple of 16, <= 256) x BK = 64 channels per k-block.  A 3x3 stride-1 conv runs 9 taps x
// (Cin/64) k-blocks, each A tile being the SAME 2-D activation matrix loaded at row offset
// m0 + dy*(W+2) + dx (the zero halo of the padded layout supplies the conv padding, TMA's
// out-of-bounds zero fill covers the matrix ends).
//
// Warp roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator +
// MMA issuer (one elected lane), warps 2..5 = epilogue (TMEM lane quarter = warp_idx % 4):
// tcgen05.ld -> +bias -> activation -> (+residual) -> fp16/fp32 vector stores of interior rows.
// One output tile per CTA; smem is sized so two CTAs share an SM and one CTA's epilogue overlaps
// the other's main loop.
#include "common.h"

namespace adas {

static constexpr int BM = 128;
static constexpr int BK = 64;                       // fp16 elements per k-block = 128 bytes = one swizzle row
static constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KiB
static constexpr int NUM_THREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address, 16-byte units
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset: 8 rows * 128 B
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // layout: SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float act_apply(float x, int act) {
    if (act == 1) return __fdividef(x, 1.0f + __expf(-x));   // SiLU
    if (act == 2) return fmaxf(x, 0.0f);                     // ReLU
    return x;
}

__global__ void __launch_bounds__(NUM_THREADS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_holder;
    __shared__ float s_bias[256];

    const int warp_idx = threadIdx.x >> 5;   // warp-uniform
    const int lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * p.BN;
    const int m0 = blockIdx.y * BM;
    const int stages = p.stages;
    const int b_stage_bytes = ((p.BN * BK * 2) + 1023) & ~1023;
    const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int num_kb = p.ntaps * p.kpt;

    if (warp_idx == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(&tmem_full_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)p.BN) tmem_cols <<= 1;
    if (warp_idx == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp_idx >= 2) {
        // stage the bias slice for this N tile
        for (int j = threadIdx.x - 64; j < p.BN; j += 128) {
            float b = 0.f;
            if (!p.transposed && p.bias != nullptr && (n0 + j) < p.N) b = p.bias[n0 + j];
            s_bias[j] = b;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_holder;

    if (warp_idx == 0) {
        if (lane == 0) {
            // ================= TMA producer =================
            const uint32_t tx_bytes = A_STAGE_BYTES + p.BN * BK * 2;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % stages;
                const uint32_t ph = (uint32_t)(kb / stages) & 1u;
                mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
                const uint32_t fb = smem_u32(&full_bar[s]);
                mbar_expect_tx(fb, tx_bytes);
                const int tap = kb / p.kpt;
                const int kc = kb - tap * p.kpt;
                int shift = 0;
                if (p.ntaps == 9) shift = (tap / 3 - 1) * p.Wp + (tap % 3 - 1);
                const uint32_t a_dst = smem_base + s * stage_bytes;
                tma_load_2d(a_dst, &tmA, kc * BK, m0 + shift, fb);
                tma_load_2d(a_dst + A_STAGE_BYTES, &tmB, tap * p.Kc + kc * BK, n0, fb);
            }
        }
    } else if (warp_idx == 1) {
        if (lane == 0) {
            // ================= MMA issuer =================
            // instruction descriptor: D=f32, A=B=f16, both K-major, N = BN, M = 128
            const uint32_t idesc = (1u << 4) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % stages;
                const uint32_t ph = (uint32_t)(kb / stages) & 1u;
                mbar_wait(smem_u32(&full_bar[s]), ph);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_base + s * stage_bytes;
                const uint32_t b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t ad = make_smem_desc(a_addr + k * 32);
                    const uint64_t bd = make_smem_desc(b_addr + k * 32);
                    umma_f16(tmem_base, ad, bd, idesc, (uint32_t)((kb | k) != 0));
                }
                umma_commit(smem_u32(&empty_bar[s]));   // frees the smem slot when these MMAs retire
            }
            umma_commit(smem_u32(&tmem_full_bar));      // accumulator complete
        }
    } else {
        // ================= epilogue =================
        const int q = warp_idx & 3;                       // TMEM lane quarter this warp may access
        const int row = m0 + q * 32 + lane;
        bool row_ok = row < p.M;
        if (p.mask_H > 0 && row_ok) {
            const int Wp = p.mask_W + 2;
            const int img = (p.mask_H + 2) * Wp;
            const int pp = row % img;
            const int yy = pp / Wp;
            const int xx = pp - yy * Wp;
            row_ok = (yy >= 1) && (yy <= p.mask_H) && (xx >= 1) && (xx <= p.mask_W);
        }
        float row_bias = 0.f;
        if (p.transposed && p.bias != nullptr && row < p.M) row_bias = p.bias[row];
        mbar_wait(smem_u32(&tmem_full_bar), 0);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int c = 0; c < p.BN; c += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + (uint32_t)c, v);
            tmem_ld_wait();
            const int n = n0 + c;
            if (!p.transposed) {
                if (row_ok && n < p.N) {
                    float f[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + s_bias[c + j];
                    const bool second = (n + 16) <= p.N;   // N % 8 == 0: either 8 or 16 valid columns
                    if (p.res != nullptr && p.res_ld < 0) {
                        // res_ld < 0 encodes "add residual BEFORE the activation" (ResNet BasicBlock)
                        const __half* rp = p.res + (size_t)row * (size_t)(-p.res_ld) + n;
                        uint4 r0 = *reinterpret_cast<const uint4*>(rp);
                        const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) { float2 t = __half22float2(h0[j]); f[2 * j] += t.x; f[2 * j + 1] += t.y; }
                        if (second) {
                            uint4 r1 = *reinterpret_cast<const uint4*>(rp + 8);
                            const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
                            for (int j = 0; j < 4; ++j) { float2 t = __half22float2(h1[j]); f[8 + 2 * j] += t.x; f[8 + 2 * j + 1] += t.y; }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = act_apply(f[j], p.act);
                    if (p.res != nullptr && p.res_ld > 0) {
                        // residual AFTER the activation (YOLO Bottleneck shortcut)
                        const __half* rp = p.res + (size_t)row * (size_t)p.res_ld + n;
                        uint4 r0 = *reinterpret_cast<const uint4*>(rp);
                        const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) { float2 t = __half22float2(h0[j]); f[2 * j] += t.x; f[2 * j + 1] += t.y; }
                        if (second) {
                            uint4 r1 = *reinterpret_cast<const uint4*>(rp + 8);
                            const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
                            for (int j = 0; j < 4; ++j) { float2 t = __half22float2(h1[j]); f[8 + 2 * j] += t.x; f[8 + 2 * j + 1] += t.y; }
                        }
                    }
                    if (p.out_f32) {
                        float* op = reinterpret_cast<float*>(p.out) + (size_t)row * (size_t)p.out_ld + n;
                        *reinterpret_cast<float4*>(op) = make_float4(f[0], f[1], f[2], f[3]);
                        *reinterpret_cast<float4*>(op + 4) = make_float4(f[4], f[5], f[6], f[7]);
                        if (second) {
                            *reinterpret_cast<float4*>(op + 8) = make_float4(f[8], f[9], f[10], f[11]);
                            *reinterpret_cast<float4*>(op + 12) = make_float4(f[12], f[13], f[14], f[15]);
                        }
                    } else {
                        __half* op = reinterpret_cast<__half*>(p.out) + (size_t)row * (size_t)p.out_ld + n;
                        uint4 o0, o1;
                        __half2* q0 = reinterpret_cast<__half2*>(&o0);
                        __half2* q1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            q0[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
                            q1[j] = __floats2half2_rn(f[8 + 2 * j], f[8 + 2 * j + 1]);
                        }
                        *reinterpret_cast<uint4*>(op) = o0;
                        if (second) *reinterpret_cast<uint4*>(op + 8) = o1;
                    }
                }
            } else {
                // swap-AB FC: rows are output features, columns are batch entries
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int col = n + j;
                        if (col < p.N) {
                            float x = act_apply(__uint_as_float(v[j]) + row_bias, p.act);
                            if (p.out_f32) reinterpret_cast<float*>(p.out)[(size_t)col * (size_t)p.out_ld + row] = x;
                            else reinterpret_cast<__half*>(p.out)[(size_t)col * (size_t)p.out_ld + row] = __float2half_rn(x);
                        }
                    }
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

int gemm_tc_smem_bytes(int BN, int stages) {
    const int b_stage = ((BN * BK * 2) + 1023) & ~1023;
    return stages * (A_STAGE_BYTES + b_stage) + 1024;
}

int gemm_tc_pick_stages(int BN, int num_kb) {
    // keep two CTAs resident per SM (<= ~113 KiB each) so epilogues overlap main loops
    const int b_stage = ((BN * BK * 2) + 1023) & ~1023;
    int s = (112 * 1024 - 1024) / (A_STAGE_BYTES + b_stage);
    if (s < 2) s = 4;                 // BN = 256: one CTA per SM with a deeper ring
    if (BN > 160 && s < 4) s = 4;
    if (s > 6) s = 6;
    if (s > num_kb) s = num_kb < 2 ? 2 : num_kb;
    return s;
}

int gemm_tc_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t st) {
    ADAS_CHECK(p.BN % 16 == 0 && p.BN >= 16 && p.BN <= 256, "gemm_tc: bad BN %d", p.BN);
    ADAS_CHECK(p.N % 8 == 0 || p.transposed, "gemm_tc: N %d must be a multiple of 8", p.N);
    ADAS_CHECK(p.stages >= 2 && p.stages <= 8, "gemm_tc: bad stage count %d", p.stages);
    const int smem = gemm_tc_smem_bytes(p.BN, p.stages);
    static int max_set = 0;
    if (smem > max_set) {
        ADAS_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
        max_set = 227 * 1024;
    }
    dim3 grid((p.N + p.BN - 1) / p.BN, (p.M + BM - 1) / BM, 1);
    gemm_tc_kernel<<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, p);
    count_launch();
    ADAS_CUDA(cudaGetLastError());
    return 0;
}

// ---- TMA descriptor (driver entry point fetched through the runtime; no libcuda link) ----------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    return fn;
}

int make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_rows) {
    PFN_encodeTiled fn = get_encode_fn();
    ADAS_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
    ADAS_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16-byte aligned", base);
    ADAS_CHECK((row_stride_bytes & 15) == 0, "TMA row stride %llu not a multiple of 16", (unsigned long long)row_stride_bytes);
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ADAS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed: %d (inner=%llu rows=%llu stride=%llu box=%u,%u)", (int)r,
               (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_stride_bytes, box_inner, box_rows);
    return 0;
}

// ---- SIMT validation kernel: same contract, CUDA cores, fp32 accumulate ---------------------------
__global__ void gemm_simt_kernel(const GemmParams p) {
    const int row = blockIdx.x * 64 + (threadIdx.x >> 2);       // 64 rows per block
    const int ng = blockIdx.y * 4 + (threadIdx.x & 3);          // group of 8 output columns
    const int n = ng * 8;
    if (row >= p.M || n >= p.N) return;
    bool row_ok = true;
    if (p.mask_H > 0) {
        const int Wp = p.mask_W + 2;
        const int img = (p.mask_H + 2) * Wp;
        const int pp = row % img;
        const int yy = pp / Wp;
        const int xx = pp - yy * Wp;
        row_ok = (yy >= 1) && (yy <= p.mask_H) && (xx >= 1) && (xx <= p.mask_W);
    }
    if (!row_ok) return;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int ncols = (p.N - n) < 8 ? (p.N - n) : 8;
    for (int tap = 0; tap < p.ntaps; ++tap) {
        int shift = 0;
        if (p.ntaps == 9) shift = (tap / 3 - 1) * p.Wp + (tap % 3 - 1);
        const long ar = (long)row + shift;
        if (ar < 0 || ar >= p.M) continue;
        const __half* a = p.A + (size_t)ar * p.a_ld;
        for (int c = 0; c < p.Kc; ++c) {
            const float av = __half2float(a[c]);
            for (int j = 0; j < ncols; ++j)
                acc[j] = fmaf(av, __half2float(p.Wt[(size_t)(n + j) * p.w_ld + tap * p.Kc + c]), acc[j]);
        }
    }
    for (int j = 0; j < ncols; ++j) {
        float x = acc[j];
        if (p.bias) x += p.transposed ? p.bias[row] : p.bias[n + j];
        if (p.res != nullptr && p.res_ld < 0) x += __half2float(p.res[(size_t)row * (size_t)(-p.res_ld) + n + j]);
        x = act_apply(x, p.act);
        if (p.res != nullptr && p.res_ld > 0) x += __half2float(p.res[(size_t)row * (size_t)p.res_ld + n + j]);
        const size_t o = p.transposed ? ((size_t)(n + j) * p.out_ld + row) : ((size_t)row * p.out_ld + n + j);
        if (p.out_f32) reinterpret_cast<float*>(p.out)[o] = x;
        else reinterpret_cast<__half*>(p.out)[o] = __float2half_rn(x);
    }
}

int gemm_simt_launch(const GemmParams& p, cudaStream_t st) {
    dim3 grid((p.M + 63) / 64, (p.N + 31) / 32, 1);
    gemm_simt_kernel<<<grid, 256, 0, st>>>(p);
    count_launch();
    ADAS_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace adas
