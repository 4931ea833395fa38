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
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace adas {

}  // namespace adas
#include "tc_common.cuh"
namespace adas {

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
                if ((p.dbg & 1) && kb >= stages) break;      // DEBUG: operands loaded once, MMA rate only
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
                if (!((p.dbg & 1) && kb >= stages)) mbar_wait(smem_u32(&full_bar[s]), ph);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_base + s * stage_bytes;
                const uint32_t b_addr = a_addr + A_STAGE_BYTES;
                if (!(p.dbg & 2)) {                          // DEBUG bit 1: skip the MMAs, TMA rate only
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ad = make_smem_desc(a_addr + k * 32);
                        const uint64_t bd = make_smem_desc(b_addr + k * 32);
                        umma_f16(tmem_base, ad, bd, idesc, (uint32_t)((kb | k) != 0));
                    }
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
            if (p.dbg & 16) break;
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
                    for (int j = 0; j < 16; ++j) f[j] = act_apply(f[j], (p.dbg & 8) ? 0 : p.act);
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
                    if ((p.dbg & 4) && f[0] != 123456.f) continue;
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


// =====================================================================================================
// v2: persistent, operand-traffic-aware variant (the product path).
//
// Why: the first kernel moved 32-44 operand bytes-per-kMAC tiles (128 x BN) and was bound by L2->SM operand
// traffic, not by the tensor pipe (profiles/launches_r01_v1.md).  v2 raises MACs per operand byte:
//   * MT = 2 sub-tiles of 128 rows share every weight tile when BN <= 128 (BM = 256);
//   * "slab" mode for 3x3 convs: one 136-row activation slab per (dy, k-block) feeds the three dx taps -- the
//     UMMA descriptor simply starts dx rows into the slab (matrix-descriptor base offset = row phase), so the
//     activation tile is fetched 3x instead of 9x;
//   * persistent CTAs (one per SM) with two TMEM accumulator stages: the epilogue of tile i (8 warps) overlaps
//     the main loop of tile i+1.
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM alloc + MMA issuer, warps 2..9 epilogue.
static constexpr int V2_THREADS = 320;
static constexpr int SLAB_ROWS = 136;
static constexpr int SLAB_BYTES = SLAB_ROWS * BK * 2;   // 17408 = 17 * 1024

__device__ __forceinline__ uint64_t make_smem_desc_off(uint32_t smem_addr) {
    uint64_t d = make_smem_desc(smem_addr);
    d |= (uint64_t)((smem_addr >> 7) & 7u) << 49;        // base offset: row phase inside the 8-row swizzle atom
    return d;
}

struct GemmV2 {
    GemmParams p;
    int MT;          // 1..4 sub-tiles of 128 rows per CTA tile (they share every weight tile)
    int sub_cols;    // TMEM columns per sub-tile accumulator
    int acc_stages;  // 2 when two whole accumulator sets fit in the 512 TMEM columns (epilogue overlaps the next tile), else 1
    int slab;        // 1: ntaps == 9 handled as 3 dy-steps x 3 dx-taps from a shared slab
    int a_sub_bytes; // bytes of one A sub-tile in a stage
    int b_bytes;     // bytes of one B tile (BN rows), 1024-aligned
    int stage_bytes;
    int n_tiles, m_tiles, total_tiles;
    int use_base_offset;
    int pair;        // 1: cta_group::2 -- the two CTAs of a cluster run ONE M=256 MMA per step, each staging half of the weight tile
    int mc;          // 1: CTA pairs (cluster 2x1x1) on adjacent M tiles share every weight tile through TMA multicast
    int pdl;         // 1: launched with programmatic stream serialisation (griddepcontrol in the kernel)
    int work_items;  // scheduler items: tiles, or tile pairs when mc
};


// kCluster = false: plain launch, no cluster / cta_group::2 instructions in the binary (a kernel that contains them must be
// launched with a cluster attribute).  kCluster = true: CTA pairs (TMA multicast or cta_group::2 MMA).
// kEpiWarps = 8 (product) or 16 (EXPERIMENTAL, ADAS_B200_EPI16=1): four epilogue warps per scheduler instead of two, 16-column
// batches so that the register budget of a 576-thread CTA (112 / thread) holds.  ncu on the 8-warp kernel: the epilogue issues
// ~370 instructions per 32-column batch but takes ~3x the issue time -- latency-bound with two warps per scheduler.
// kHoist (EXPERIMENTAL, ADAS_B200_HOIST=1 with the 8-warp epilogue): output rows / halo masks computed before the accumulator wait.
template <bool kCluster, int kEpiWarps = 8, bool kHoist = false>
__global__ void __launch_bounds__(64 + 32 * kEpiWarps, 1)
gemm_tc_v2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmV2 g) {
    const int g_pair = kCluster ? g.pair : 0;
    const int g_mc = kCluster ? g.mc : 0;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ uint32_t tmem_holder;
    __shared__ float s_bias[2][256];
    __shared__ __align__(16) float s_bias_al[2][256];     // 16-warp variant only (vector loads); unused -> dropped elsewhere

    const GemmParams& p = g.p;
    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int stages = p.stages;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int taps_per_step = g.slab ? 3 : 1;
    const int ksteps = (g.slab ? 3 : p.ntaps) * p.kpt;
    const int BMT = BM * g.MT;
    const int mt_cols = g.sub_cols;               // TMEM column offset between sub-tiles
    const int acc_stride = g.MT * g.sub_cols;     // TMEM columns per accumulator stage
    const bool acc2 = g.acc_stages == 2;
    // work distribution: item w -> (n tile, m tile).  With multicast pairs both CTAs of a cluster walk the same items and
    // take the even / odd M tile of the pair, so they need the same weight tiles at the same time.
    const int cl2 = (g_mc | g_pair);
    const int crank = cl2 ? (int)cluster_ctarank() : 0;
    const int w_first = cl2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int w_step = cl2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
#define TILE_OF(w, nt_, mt_) const int nt_ = (w) % g.n_tiles; const int mt_ = cl2 ? 2 * ((w) / g.n_tiles) + crank : (w) / g.n_tiles;

    if (warp_idx == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), g_mc ? 2 : 1);     // multicast: both CTAs' MMA warps release a stage
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tfull_bar[s]), 1);
            mbar_init(smem_u32(&tempty_bar[s]), kEpiWarps * (g_pair ? 2 : 1));   // one arrive per epilogue warp; pair: both CTAs' epilogues drain the leader's MMA
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp_idx == 1) {
        if (g_pair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (cl2) cluster_sync_all();          // peer barriers are initialised before any multicast / remote arrive
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_holder;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the
    // tail of the previous kernel in the stream; its results may only be touched after this wait.  The next kernel is
    // allowed to start its own prologue as soon as SMs free up.
    if (g.pdl) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }

    if (p.dbg & 64) {
        // DEBUG: launch skeleton only (prologue + dependency wait + teardown)
    } else if (warp_idx == 0) {
        if (lane == 0) {
            // ================= TMA producer =================
            const uint32_t tx_bytes = p.s2 ? (uint32_t)(p.s2_bw * p.s2_bh * BK * 2 + p.BN * BK * 2)
                                           : (uint32_t)(g.MT * (g.slab ? SLAB_BYTES : A_STAGE_BYTES) + taps_per_step * p.BN * BK * 2);
            uint32_t it = 0;
            const int half_rows = p.BN >> 1;
            // pair mode: this CTA stages its own A rows + half of each weight tile; all bytes of both CTAs are accounted on the
            // LEADER's full barrier (the leader expects 2x), the follower only issues its loads
            const uint32_t pair_bytes = (uint32_t)(g.MT * (g.slab ? SLAB_BYTES : A_STAGE_BYTES) + taps_per_step * half_rows * BK * 2);
            for (int w = w_first; w < g.work_items; w += w_step) {
                if (p.dbg & 32) break;                     // DEBUG: no loads at all (MMA-rate experiment)
                TILE_OF(w, n_t, m_t)
                const int n0 = n_t * p.BN;
                const int m0 = m_t * BMT;
                for (int ks = 0; ks < ksteps; ++ks, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (it / stages) & 1u;
                    mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
                    uint32_t fb = smem_u32(&full_bar[s]);
                    if (g_pair) {
                        if (crank == 0) mbar_expect_tx(fb, 2u * pair_bytes);
                        fb = mapa_rank(fb, 0);               // cluster address of the leader's barrier
                    } else {
                        mbar_expect_tx(fb, tx_bytes);
                    }
                    // K order is (dy, k-block, dx) in BOTH modes so every output element accumulates in the same order
                    // whatever tile shape the autotuner picks (bit-identical results across batch sizes).
                    int grp, kc;
                    if (g.slab || p.ntaps != 9) { grp = ks / p.kpt; kc = ks - grp * p.kpt; }
                    else { const int dy = ks / (3 * p.kpt); const int r = ks - dy * 3 * p.kpt; kc = r / 3; grp = dy * 3 + (r - kc * 3); }
                    const uint32_t a_dst = smem_base + s * g.stage_bytes;
                    const uint32_t b_dst = a_dst + g.MT * g.a_sub_bytes;
                    if (g.slab) {
                        const int r0 = m0 + (grp - 1) * p.Wp - 1;
                        for (int mt = 0; mt < g.MT; ++mt) {
                            if (g_pair) tma_load_2d_pair(a_dst + mt * g.a_sub_bytes, &tmA, kc * BK, r0 + mt * BM, fb);
                            else tma_load_2d(a_dst + mt * g.a_sub_bytes, &tmA, kc * BK, r0 + mt * BM, fb);
                        }
                        for (int dx = 0; dx < 3; ++dx) {
                            if (g_pair) tma_load_2d_pair(b_dst + dx * g.b_bytes, &tmB, (grp * 3 + dx) * p.Kc + kc * BK, n0 + crank * half_rows, fb);
                            else if (g_mc) tma_load_2d_mc(b_dst + dx * g.b_bytes + crank * half_rows * 128, &tmB, (grp * 3 + dx) * p.Kc + kc * BK,
                                                     n0 + crank * half_rows, fb, (uint16_t)3);
                            else tma_load_2d(b_dst + dx * g.b_bytes, &tmB, (grp * 3 + dx) * p.Kc + kc * BK, n0, fb);
                        }
                    } else if (p.s2) {
                        // stride-2 conv: tile = bw x bh output pixels of image b; input pixel of tap (dy,dx) is (2*yo+dy, 2*xo+dx)
                        // in padded coordinates, fetched by one 4-D TMA box with traversal stride 2 in x and y
                        const int mt_idx = m_t;
                        const int per_img = p.s2_tw * p.s2_th;
                        const int b = mt_idx / per_img;
                        const int rem = mt_idx - b * per_img;
                        const int ty = rem / p.s2_tw, tx = rem - ty * p.s2_tw;
                        const int dy = p.ntaps == 9 ? grp / 3 : 1, dx = p.ntaps == 9 ? grp % 3 : 1;
                        tma_load_4d(a_dst, &tmA, kc * BK, 2 * tx * p.s2_bw + dx, 2 * ty * p.s2_bh + dy, b, fb);
                        tma_load_2d(b_dst, &tmB, grp * p.Kc + kc * BK, n0, fb);
                    } else {
                        int shift = 0;
                        if (p.ntaps == 9) shift = (grp / 3 - 1) * p.Wp + (grp % 3 - 1);
                        else if (p.ntaps == 4) shift = (grp - 2) * p.Wp;          // stem: row pairs yo-1 .. yo+2 (plan.py stem7x7s2)
                        for (int mt = 0; mt < g.MT; ++mt) {
                            if (g_pair) tma_load_2d_pair(a_dst + mt * g.a_sub_bytes, &tmA, kc * BK, m0 + shift + mt * BM, fb);
                            else tma_load_2d(a_dst + mt * g.a_sub_bytes, &tmA, kc * BK, m0 + shift + mt * BM, fb);
                        }
                        if (g_pair) tma_load_2d_pair(b_dst, &tmB, grp * p.Kc + kc * BK, n0 + crank * half_rows, fb);
                        else if (g_mc) tma_load_2d_mc(b_dst + crank * half_rows * 128, &tmB, grp * p.Kc + kc * BK, n0 + crank * half_rows, fb, (uint16_t)3);
                        else tma_load_2d(b_dst, &tmB, grp * p.Kc + kc * BK, n0, fb);
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        if (!(g_pair && crank != 0)) {
            // ================= MMA issuer (pair mode: leader CTA only, M = 256 across both SMs) =================
            // The WHOLE warp walks the loop and waits on the barriers; one elected lane issues tcgen05.mma / commit.  Keeping the
            // control flow warp-uniform lets the compiler hold descriptors and addresses in uniform registers -- with a
            // single-lane loop every UTCHMMA operand went through R2UR and the issue rate, not the tensor pipe, set the pace.
            const uint32_t idesc = (1u << 4) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)((g_pair ? 2 * BM : BM) >> 4) << 24);
            const uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            const uint32_t a_step = (uint32_t)g.a_sub_bytes >> 4, b_step = (uint32_t)g.b_bytes >> 4;
            const int n_dx = taps_per_step, n_mt = g.MT;
            const bool skip_mma = (p.dbg & 2) != 0, no_wait = (p.dbg & 32) != 0;
            uint32_t it = 0, tile_it = 0;
            for (int w = w_first; w < g.work_items; w += w_step, ++tile_it) {
                const int as = acc2 ? (int)(tile_it & 1) : 0;
                mbar_wait(smem_u32(&tempty_bar[as]), ((acc2 ? (tile_it >> 1) : tile_it) & 1u) ^ 1u);     // epilogue drained this accumulator stage
                tcgen05_fence_after();
                const uint32_t d_base = tmem_base + (uint32_t)(as * acc_stride);
                for (int ks = 0; ks < ksteps; ++ks, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (it / stages) & 1u;
                    if (!no_wait) mbar_wait(smem_u32(&full_bar[s]), ph);
                    tcgen05_fence_after();
                    const uint32_t a_lo = ((smem_base + s * g.stage_bytes) & 0x3FFFFu) >> 4;      // 16-byte units
                    const uint32_t b_lo = a_lo + (uint32_t)n_mt * a_step;
                    if (elect_one()) {
                        if (!skip_mma) {
                            for (int dx = 0; dx < n_dx; ++dx) {
                                for (int mt = 0; mt < n_mt; ++mt) {
                                    const uint32_t a_sub = a_lo + (uint32_t)mt * a_step + (g.slab ? (uint32_t)dx * 8u : 0u);   // +dx rows of 128 B
                                    const uint32_t b_sub = b_lo + (uint32_t)dx * b_step;
                                    const uint32_t d = d_base + (uint32_t)(mt * mt_cols);
#pragma unroll
                                    for (int k = 0; k < BK / 16; ++k) {
                                        const uint64_t ad = desc_hi | (uint64_t)(a_sub + 2u * k);
                                        const uint64_t bd = desc_hi | (uint64_t)(b_sub + 2u * k);
                                        if (g_pair) umma_f16_pair(d, ad, bd, idesc, (uint32_t)((ks | dx | k) != 0));
                                        else umma_f16(d, ad, bd, idesc, (uint32_t)((ks | dx | k) != 0));
                                    }
                                }
                            }
                        }
                        if (g_pair) umma_commit_pair(smem_u32(&empty_bar[s]), (uint16_t)3);
                        else if (g_mc) umma_commit_mc(smem_u32(&empty_bar[s]), (uint16_t)3);   // frees the stage in both CTAs of the pair
                        else umma_commit(smem_u32(&empty_bar[s]));
                    }
                    __syncwarp();
                }
                if (elect_one()) {
                    if (g_pair) umma_commit_pair(smem_u32(&tfull_bar[as]), (uint16_t)3);      // both epilogues may read their TMEM halves
                    else umma_commit(smem_u32(&tfull_bar[as]));
                }
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue (8 warps) =================
        // Each warp owns TMEM lane quarter q; the two warps sharing a quarter alternate 32-column batches.
        // Per batch: two tcgen05.ld.x16 are issued, the bias (smem broadcast) and residual (global) operands are
        // fetched while they are in flight, then one wait, the math, and 16-byte stores of the thread's row.
        const int q = warp_idx & 3;
        const int half = (warp_idx - 2) >> 2;
        const int et = threadIdx.x - 64;
        // 32-byte stores need 32-byte aligned rows
        const bool wide_st = !(p.dbg & 512) && !p.out_f32 && (p.out_ld % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 31u) == 0);
        uint32_t tile_it = 0;
        for (int w = w_first; w < g.work_items; w += w_step, ++tile_it) {
            const int as = acc2 ? (int)(tile_it & 1) : 0;
            const int bs = (int)(tile_it & 1);        // bias staging buffer (alternates even with one accumulator stage)
            TILE_OF(w, n_t, m_t)
            const int n0 = n_t * p.BN;
            const int m0 = m_t * BMT;
            if (!p.transposed) {
                for (int j = et; j < p.BN; j += 32 * kEpiWarps) {
                    const float bv = (p.bias != nullptr && (n0 + j) < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
                    if constexpr (kEpiWarps == 16) s_bias_al[bs][j] = bv;
                    else s_bias[bs][j] = bv;
                }
            }
            // (16-warp variant) the output row of every sub-tile and its halo mask need two integer divisions: computed here, while
            // the main loop is still running, instead of after the accumulator is ready (ncu source page of the 8-warp kernel:
            // this block holds ~40 % as many stall samples as a whole 32-column batch)
            int pre_row[4] = {0, 0, 0, 0};
            bool pre_ok[4] = {false, false, false, false};
            if constexpr (kEpiWarps == 16 || kHoist) {
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
                    if (mt < g.MT) {
                        int row = m0 + mt * BM + q * 32 + lane;
                        bool row_ok = row < p.M;
                        if (p.s2) {
                            const int r = q * 32 + lane;
                            const int per_img = p.s2_tw * p.s2_th;
                            const int b = m_t / per_img;
                            const int rem = m_t - b * per_img;
                            const int ty = rem / p.s2_tw, tx = rem - ty * p.s2_tw;
                            const int j = r / p.s2_bw, i = r - j * p.s2_bw;
                            const int yo = ty * p.s2_bh + j, xo = tx * p.s2_bw + i;
                            row_ok = (r < p.s2_bw * p.s2_bh) && (yo < p.s2_Ho) && (xo < p.s2_Wo);
                            row = (b * (p.s2_Ho + 2) + yo + 1) * (p.s2_Wo + 2) + xo + 1;
                        } else if (p.mask_H > 0 && row_ok) {
                            const int Wp = p.mask_W + 2;
                            const int img = (p.mask_H + 2) * Wp;
                            const int pp = row % img;
                            const int yy = pp / Wp;
                            const int xx = pp - yy * Wp;
                            row_ok = (yy >= 1) && (yy <= p.mask_H) && (xx >= 1) && (xx <= p.mask_W);
                        }
                        pre_row[mt] = row;
                        pre_ok[mt] = row_ok;
                    }
                }
            }
            if constexpr (kEpiWarps == 16) asm volatile("bar.sync 1, 512;" ::: "memory");
            else asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(smem_u32(&tfull_bar[as]), (acc2 ? (tile_it >> 1) : tile_it) & 1u);
            tcgen05_fence_after();
            for (int mt = 0; mt < g.MT; ++mt) {
                int row = m0 + mt * BM + q * 32 + lane;
                bool row_ok = row < p.M;
                if constexpr (kEpiWarps == 16 || kHoist) {
                    row = mt == 0 ? pre_row[0] : mt == 1 ? pre_row[1] : mt == 2 ? pre_row[2] : pre_row[3];
                    row_ok = mt == 0 ? pre_ok[0] : mt == 1 ? pre_ok[1] : mt == 2 ? pre_ok[2] : pre_ok[3];
                } else
                if (p.s2) {
                    const int r = q * 32 + lane;
                    const int mt_idx = m_t;
                    const int per_img = p.s2_tw * p.s2_th;
                    const int b = mt_idx / per_img;
                    const int rem = mt_idx - b * per_img;
                    const int ty = rem / p.s2_tw, tx = rem - ty * p.s2_tw;
                    const int j = r / p.s2_bw, i = r - j * p.s2_bw;
                    const int yo = ty * p.s2_bh + j, xo = tx * p.s2_bw + i;
                    row_ok = (r < p.s2_bw * p.s2_bh) && (yo < p.s2_Ho) && (xo < p.s2_Wo);
                    row = (b * (p.s2_Ho + 2) + yo + 1) * (p.s2_Wo + 2) + xo + 1;
                } else if (p.mask_H > 0 && row_ok) {
                    const int Wp = p.mask_W + 2;
                    const int img = (p.mask_H + 2) * Wp;
                    const int pp = row % img;
                    const int yy = pp / Wp;
                    const int xx = pp - yy * Wp;
                    row_ok = (yy >= 1) && (yy <= p.mask_H) && (xx >= 1) && (xx <= p.mask_W);
                }
                float row_bias = 0.f;
                if (p.transposed && p.bias != nullptr && row < p.M) row_bias = p.bias[row];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * acc_stride + mt * mt_cols);
                const size_t res_ld = (size_t)(p.res_ld < 0 ? -p.res_ld : p.res_ld);
                if constexpr (kEpiWarps == 16) {
                    // four warps per TMEM lane quarter take interleaved 16-column batches
                    const int part = (warp_idx - 2) >> 2;
                    for (int c = part * 16; c < p.BN; c += 64) {
                        if (p.dbg & 16) break;
                        uint32_t v[16];
                        tmem_ld16(taddr + (uint32_t)c, v);
                        const int n = n0 + c;
                        const int ncols = min(16, p.N - n);
                        uint4 rr[2];
                        const bool has_res = (p.res != nullptr) && row_ok && !p.transposed;
                        if (has_res) {
                            const __half* rp = p.res + (size_t)row * res_ld + n;
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                if (k * 8 < ncols) rr[k] = *reinterpret_cast<const uint4*>(rp + k * 8);
                        }
                        tmem_ld_wait();
                        if (!p.transposed) {
                            if (row_ok && ncols > 0) {
                                float f[16];
                                const float4* sb4 = reinterpret_cast<const float4*>(&s_bias_al[bs][c]);      // c is a multiple of 16
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float4 b4 = sb4[j];
                                    f[4 * j] = __uint_as_float(v[4 * j]) + b4.x; f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
                                    f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z; f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
                                }
                                if (has_res && p.res_ld < 0) {
#pragma unroll
                                    for (int k = 0; k < 2; ++k)
                                        if (k * 8 < ncols) {
                                            const __half2* h = reinterpret_cast<const __half2*>(&rr[k]);
#pragma unroll
                                            for (int j = 0; j < 4; ++j) { float2 tt = __half22float2(h[j]); f[k * 8 + 2 * j] += tt.x; f[k * 8 + 2 * j + 1] += tt.y; }
                                        }
                                }
                                const int act = (p.dbg & 8) ? 0 : p.act;
                                if (act == 1) {
#pragma unroll
                                    for (int j = 0; j < 16; j += 4) silu4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                                } else if (act == 2) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
                                }
                                if (has_res && p.res_ld > 0) {
#pragma unroll
                                    for (int k = 0; k < 2; ++k)
                                        if (k * 8 < ncols) {
                                            const __half2* h = reinterpret_cast<const __half2*>(&rr[k]);
#pragma unroll
                                            for (int j = 0; j < 4; ++j) { float2 tt = __half22float2(h[j]); f[k * 8 + 2 * j] += tt.x; f[k * 8 + 2 * j + 1] += tt.y; }
                                        }
                                }
                                if ((p.dbg & 4) && f[0] != 123456.f) continue;
                                if (p.out_f32) {
                                    float* op = reinterpret_cast<float*>(p.out) + (size_t)row * (size_t)p.out_ld + n;
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        if (k * 4 < ncols) *reinterpret_cast<float4*>(op + k * 4) = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
                                } else {
                                    __half* op = reinterpret_cast<__half*>(p.out) + (size_t)row * (size_t)p.out_ld + n;
                                    uint32_t o[8];
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        const __half2 h = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
                                        o[j] = *reinterpret_cast<const uint32_t*>(&h);
                                    }
                                    if (wide_st && ncols == 16) {
                                        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(op), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
                                                     "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
                                    } else {
                                        *reinterpret_cast<uint4*>(op) = make_uint4(o[0], o[1], o[2], o[3]);
                                        if (ncols > 8) *reinterpret_cast<uint4*>(op + 8) = make_uint4(o[4], o[5], o[6], o[7]);
                                    }
                                }
                            }
                        } else if (row_ok) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int col = n + j;
                                if (col < p.N) {
                                    const float x = act_apply(__uint_as_float(v[j]) + row_bias, p.act);
                                    if (p.out_f32) reinterpret_cast<float*>(p.out)[(size_t)col * (size_t)p.out_ld + row] = x;
                                    else reinterpret_cast<__half*>(p.out)[(size_t)col * (size_t)p.out_ld + row] = __float2half_rn(x);
                                }
                            }
                        }
                    }
                } else
                for (int c = half * 32; c < p.BN; c += 64) {
                    if (p.dbg & 16) break;
                    const bool two = (c + 16) < p.BN;          // BN is a multiple of 16: a batch is 32 or 16 columns
                    uint32_t v0[16], v1[16];
                    tmem_ld16(taddr + (uint32_t)c, v0);
                    if (two) tmem_ld16(taddr + (uint32_t)(c + 16), v1);
                    const int n = n0 + c;
                    const int ncols = min(two ? 32 : 16, p.N - n);     // valid columns in this batch (multiple of 8, may be <= 0)
                    // operands fetched under the TMEM load latency
                    uint4 rr[4];
                    const bool has_res = (p.res != nullptr) && row_ok && !p.transposed;
                    if (has_res) {
                        const __half* rp = p.res + (size_t)row * res_ld + n;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (k * 8 < ncols) rr[k] = *reinterpret_cast<const uint4*>(rp + k * 8);
                    }
                    tmem_ld_wait();
                    if (!p.transposed) {
                        if (row_ok && ncols > 0) {
                            float f[32];
#pragma unroll
                            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v0[j]) + s_bias[bs][c + j];
                            if (two) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) f[16 + j] = __uint_as_float(v1[j]) + s_bias[bs][c + 16 + j];
                            }
                            if (has_res && p.res_ld < 0) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if (k * 8 < ncols) {
                                        const __half2* h = reinterpret_cast<const __half2*>(&rr[k]);
#pragma unroll
                                        for (int j = 0; j < 4; ++j) { float2 tt = __half22float2(h[j]); f[k * 8 + 2 * j] += tt.x; f[k * 8 + 2 * j + 1] += tt.y; }
                                    }
                            }
                            const int act = (p.dbg & 8) ? 0 : p.act;
                            if (act == 1) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4) silu4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                            } else if (act == 2) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                            }
                            if (has_res && p.res_ld > 0) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if (k * 8 < ncols) {
                                        const __half2* h = reinterpret_cast<const __half2*>(&rr[k]);
#pragma unroll
                                        for (int j = 0; j < 4; ++j) { float2 tt = __half22float2(h[j]); f[k * 8 + 2 * j] += tt.x; f[k * 8 + 2 * j + 1] += tt.y; }
                                    }
                            }
                            if ((p.dbg & 4) && f[0] != 123456.f) continue;
                            if (p.out_f32) {
                                float* op = reinterpret_cast<float*>(p.out) + (size_t)row * (size_t)p.out_ld + n;
#pragma unroll
                                for (int k = 0; k < 8; ++k)
                                    if (k * 4 < ncols) *reinterpret_cast<float4*>(op + k * 4) = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
                            } else {
                                __half* op = reinterpret_cast<__half*>(p.out) + (size_t)row * (size_t)p.out_ld + n;
                                if (wide_st && ncols == 32) {
                                    // two 32-byte stores: every lane writes whole sectors of its row
                                    uint32_t o[16];
#pragma unroll
                                    for (int j = 0; j < 16; ++j) {
                                        const __half2 h = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
                                        o[j] = *reinterpret_cast<const uint32_t*>(&h);
                                    }
                                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(op), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
                                                 "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
                                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(op + 16), "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]),
                                                 "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15]) : "memory");
                                } else
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if (k * 8 < ncols) {
                                        uint4 o;
                                        __half2* qh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                                        for (int j = 0; j < 4; ++j) qh[j] = __floats2half2_rn(f[k * 8 + 2 * j], f[k * 8 + 2 * j + 1]);
                                        *reinterpret_cast<uint4*>(op + k * 8) = o;
                                    }
                            }
                        }
                    } else if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = n + j;
                            if (col < p.N && (j < 16 || two)) {
                                const float x = act_apply(__uint_as_float(j < 16 ? v0[j & 15] : v1[j & 15]) + row_bias, p.act);
                                if (p.out_f32) reinterpret_cast<float*>(p.out)[(size_t)col * (size_t)p.out_ld + row] = x;
                                else reinterpret_cast<__half*>(p.out)[(size_t)col * (size_t)p.out_ld + row] = __float2half_rn(x);
                            }
                        }
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (g_pair) mbar_arrive_cluster(mapa_rank(smem_u32(&tempty_bar[as]), 0));   // the leader's MMA waits for both epilogues
                else mbar_arrive(smem_u32(&tempty_bar[as]));      // this warp is done reading accumulator stage `as`
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (cl2) cluster_sync_all();           // the peer may still multicast into this CTA's shared memory / arrive on its barriers
    if (warp_idx == 1) {
        __syncwarp();
        if (g_pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
#undef TILE_OF
}

static int g_num_sms = 0;

int gemm_tc_v2_config(const GemmParams& p_in, GemmV2* g) {
    GemmParams p = p_in;
    static int force_bn = -1;
    if (force_bn < 0) { const char* f = getenv("ADAS_B200_BN"); force_bn = f ? atoi(f) : 0; }
    if (force_bn >= 16 && force_bn <= 256 && force_bn % 16 == 0 && force_bn <= ((p.N + 15) / 16) * 16 && !p.transposed) p.BN = force_bn;   // test hook
    g->p = p;
    g->sub_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256;
    g->MT = p.mt_hint >= 1 ? p.mt_hint : ((p.BN <= 128) ? 2 : 1);
    static int force_mt = -1;
    if (force_mt < 0) { const char* f = getenv("ADAS_B200_MT"); force_mt = f ? atoi(f) : 0; }
    if (force_mt >= 1 && force_mt <= 4 && force_mt * g->sub_cols <= 512) g->MT = force_mt;     // test hook: exercise every sub-tile count
    if (p.s2) g->MT = 1;
    if (g->MT > 4 || g->MT * g->sub_cols > 512) return 1;
    g->acc_stages = (2 * g->MT * g->sub_cols <= 512) ? 2 : 1;
    const int want_pair = (p.mc_hint == 2 && !p.s2 && p.BN % 32 == 0) ? 1 : 0;
    const int b_bytes = ((((want_pair ? p.BN / 2 : p.BN)) * BK * 2) + 1023) & ~1023;
    g->b_bytes = b_bytes;
    const int budget = 218 * 1024;
    g->slab = 0;
    if (p.ntaps == 9 && !p.s2) {
        const int slab_stage = g->MT * SLAB_BYTES + 3 * b_bytes;
        if (2 * slab_stage <= budget) g->slab = 1;
    }
    g->a_sub_bytes = g->slab ? SLAB_BYTES : A_STAGE_BYTES;
    g->stage_bytes = g->MT * g->a_sub_bytes + (g->slab ? 3 : 1) * b_bytes;
    int stages = budget / g->stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) return 1;
    g->p.stages = stages;
    const int BMT = BM * g->MT;
    g->n_tiles = (p.N + p.BN - 1) / p.BN;
    g->m_tiles = (p.M + BMT - 1) / BMT;
    g->total_tiles = g->n_tiles * g->m_tiles;
    g->mc = (p.mc_hint == 1 && !p.s2 && g->m_tiles >= 2) ? 1 : 0;
    g->pair = want_pair;
    g->pdl = 0;
    g->work_items = (g->mc | g->pair) ? g->n_tiles * ((g->m_tiles + 1) / 2) : g->total_tiles;
    const char* bo = getenv("ADAS_B200_BASEOFF");
    g->use_base_offset = (bo && bo[0] == '1');   // measured on B200: the swizzle phase follows the absolute smem address, the field stays 0
    static int dbg = -1;
    if (dbg < 0) { const char* d = getenv("ADAS_B200_DBG"); dbg = d ? atoi(d) : 0; }
    g->p.dbg = dbg;
    const char* ns = getenv("ADAS_B200_NOSLAB");
    if (ns && ns[0] == '1' && g->slab) {
        g->slab = 0;
        g->a_sub_bytes = A_STAGE_BYTES;
        g->stage_bytes = g->MT * g->a_sub_bytes + b_bytes;
        stages = budget / g->stage_bytes;
        if (stages > 8) stages = 8;
        g->p.stages = stages;
    }
    return 0;
}

int gemm_tc_v2_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmV2& g, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        ADAS_CUDA(cudaFuncSetAttribute(gemm_tc_v2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096));
        ADAS_CUDA(cudaFuncSetAttribute(gemm_tc_v2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096));
        int dev = 0;
        ADAS_CUDA(cudaGetDevice(&dev));
        ADAS_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
        attr = true;
    }
    const int smem = g.p.stages * g.stage_bytes + 1024;
    static int epi16 = -1, hoist = 0;
    if (epi16 < 0) {
        const char* ev = getenv("ADAS_B200_EPI16");
        epi16 = (ev && ev[0] == '1') ? 1 : 0;
        if (epi16) ADAS_CUDA(cudaFuncSetAttribute(gemm_tc_v2_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096));
        const char* hv = getenv("ADAS_B200_HOIST");
        hoist = (!epi16 && hv && hv[0] == '1') ? 1 : 0;
        if (hoist) ADAS_CUDA(cudaFuncSetAttribute(gemm_tc_v2_kernel<false, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096));
    }
    static int pdl = -1;
    if (pdl < 0) { const char* pe = getenv("ADAS_B200_PDL"); pdl = (pe && pe[0] == '0') ? 0 : 1; }
    if (pdl && !(g.mc | g.pair)) {
        GemmV2 gp = g;
        gp.pdl = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(g.total_tiles < g_num_sms ? g.total_tiles : g_num_sms, 1, 1);
        cfg.blockDim = dim3(epi16 ? 64 + 32 * 16 : V2_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr1;
        attr1.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr1.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr1;
        cfg.numAttrs = 1;
        if (epi16) ADAS_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_v2_kernel<false, 16>, tmA, tmB, gp));
        else if (hoist) ADAS_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_v2_kernel<false, 8, true>, tmA, tmB, gp));
        else ADAS_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_v2_kernel<false>, tmA, tmB, gp));
        count_launch();
        return 0;
    }
    if (g.mc | g.pair) {
        const int pairs_max = g_num_sms / 2;
        const int pairs = g.work_items < pairs_max ? g.work_items : pairs_max;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs, 1, 1);
        cfg.blockDim = dim3(V2_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attrs[2];
        attrs[0].id = cudaLaunchAttributeClusterDimension;
        attrs[0].val.clusterDim.x = 2; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
        attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attrs;
        cfg.numAttrs = pdl ? 2 : 1;
        GemmV2 gp = g;
        gp.pdl = pdl ? 1 : 0;
        ADAS_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_v2_kernel<true>, tmA, tmB, gp));
        count_launch();
        return 0;
    }
    int grid = g.total_tiles < g_num_sms ? g.total_tiles : g_num_sms;
    gemm_tc_v2_kernel<false><<<grid, V2_THREADS, smem, st>>>(tmA, tmB, g);
    count_launch();
    ADAS_CUDA(cudaGetLastError());
    return 0;
}


// Per-layer tile choice.  Measured on B200 (profiles/r01_gemm_analysis.md): one SM ingests ~32 operand bytes per
// clock from L2 while its tensor pipe retires 4096 MACs per clock, so a tile is modelled by max(operand bytes / 32,
// MMA clocks, epilogue clocks) and layers are charged whole waves of 148 persistent CTAs.
int gemm_tc_v2_candidates(int M, int N, int Kc, int ntaps, int max_out, int* BN_out, int* mt_hint_out, int pair) {
    // (the multicast on/off variants of each returned candidate are tried by the autotuner in engine.cu)
    const int cand[] = {256, 192, 160, 128, 96, 80, 64, 48, 32, 16};
    struct C { double t; int BN, mt; } list[48];
    int n = 0;
    const int kpt = (Kc + 63) / 64;
    for (int ci = 0; ci < 10; ++ci) {
        int BN = cand[ci];
        if (BN > N) { if (ci + 1 < 10 && cand[ci + 1] >= N) continue; BN = (N + 15) / 16 * 16; }
        if (BN > 256) continue;
        if (BN < 64 && N >= 64 && max_out > 1) continue;       // narrow tiles only ever win on narrow layers
        const int n_tiles = (N + BN - 1) / BN;
        if ((double)n_tiles * BN > 1.35 * N) continue;          // too much padded-N work
        bool dup = false;
        for (int k = 0; k < n; ++k) dup = dup || (list[k].BN == BN);
        if (dup) continue;
        for (int mt = 1; mt <= 4; ++mt) {
            if (mt >= 2 && BN > 128) continue;
            GemmParams p;
            memset(&p, 0, sizeof(p));
            p.M = M; p.N = N; p.Kc = Kc; p.ntaps = ntaps; p.kpt = kpt; p.BN = BN; p.mt_hint = mt;
            p.mc_hint = pair ? 2 : 0;
            if (pair && BN % 32 != 0) continue;
            GemmV2 g;
            if (gemm_tc_v2_config(p, &g)) continue;
            const double tiles = (double)g.total_tiles;
            const double ksteps = (g.slab ? 3.0 : (double)ntaps) * kpt;
            const double bytes = ksteps * (g.MT * (g.slab ? SLAB_BYTES : A_STAGE_BYTES) + (g.slab ? 3 : 1) * (pair ? BN / 2 : BN) * 128.0);
            const double mma = (double)g.MT * ntaps * kpt * 2.0 * BN;
            const double epi = (double)g.MT * 128.0 * BN * 0.55;
            const double per_tile = (g.acc_stages == 2 ? fmax(fmax(bytes / 32.0, mma), epi) : fmax(bytes / 32.0, mma) + epi) + 600.0;
            const double waves = ceil((pair ? 2.0 * g.work_items : tiles) / 148.0);
            const double t = waves * per_tile + epi * 0.5 + 2500.0;
            if (n < 48) { list[n].t = t; list[n].BN = BN; list[n].mt = mt; ++n; }
        }
    }
    // insertion sort by modelled time
    for (int i = 1; i < n; ++i) { C c = list[i]; int j = i - 1; while (j >= 0 && list[j].t > c.t) { list[j + 1] = list[j]; --j; } list[j + 1] = c; }
    if (n == 0) { list[0].BN = N <= 256 ? (N + 15) / 16 * 16 : 256; list[0].mt = list[0].BN <= 128 ? 2 : 1; n = 1; }
    if (n > max_out) n = max_out;
    for (int i = 0; i < n; ++i) { BN_out[i] = list[i].BN; mt_hint_out[i] = list[i].mt; }
    return n;
}

void gemm_tc_v2_choose(int M, int N, int Kc, int ntaps, int* BN_out, int* mt_hint_out) {
    gemm_tc_v2_candidates(M, N, Kc, ntaps, 1, BN_out, mt_hint_out);
}

struct GemmV2Launch {
    CUtensorMap tmA, tmB;
    GemmV2 g;
};

int gemm_tc_v2_prepare(const GemmParams& p, const void* a_base, uint64_t a_inner, uint64_t a_rows, uint64_t a_stride_bytes,
                       const void* b_base, uint64_t b_inner, uint64_t b_rows, uint64_t b_stride_bytes, void** opaque) {
    ADAS_CHECK(p.BN % 16 == 0 && p.BN >= 16 && p.BN <= 256, "gemm_tc_v2: bad BN %d", p.BN);
    ADAS_CHECK(p.N % 8 == 0 || p.transposed, "gemm_tc_v2: N %d must be a multiple of 8", p.N);
    GemmV2Launch* L = new GemmV2Launch();
    if (gemm_tc_v2_config(p, &L->g)) { delete L; ADAS_CHECK(false, "gemm_tc_v2: tile does not fit in shared memory (BN %d)", p.BN); }
    const uint32_t a_box_rows = L->g.slab ? SLAB_ROWS : BM;
    if (make_tmap_2d(&L->tmA, a_base, a_inner, a_rows, a_stride_bytes, 64, a_box_rows) ||
        make_tmap_2d(&L->tmB, b_base, b_inner, b_rows, b_stride_bytes, 64, (uint32_t)((L->g.mc | L->g.pair) ? L->g.p.BN / 2 : L->g.p.BN))) {
        delete L;
        return 1;
    }
    *opaque = L;
    return 0;
}

int gemm_tc_v2_prepare_s2(const GemmParams& p, const void* a_base, uint64_t a_C, uint64_t a_Wp, uint64_t a_Hp, uint64_t a_B, uint64_t a_ld,
                          const void* b_base, uint64_t b_inner, uint64_t b_rows, uint64_t b_stride_bytes, void** opaque) {
    ADAS_CHECK(p.s2 && p.BN % 16 == 0 && p.BN >= 16 && p.BN <= 256 && p.N % 8 == 0, "gemm_tc_v2_s2: bad tile (BN %d)", p.BN);
    GemmV2Launch* L = new GemmV2Launch();
    if (gemm_tc_v2_config(p, &L->g)) { delete L; ADAS_CHECK(false, "gemm_tc_v2_s2: tile does not fit in shared memory"); }
    if (make_tmap_4d_s2(&L->tmA, a_base, a_C, a_Wp, a_Hp, a_B, a_ld, 2u * (uint32_t)p.s2_bw, 2u * (uint32_t)p.s2_bh) ||
        make_tmap_2d(&L->tmB, b_base, b_inner, b_rows, b_stride_bytes, 64, (uint32_t)L->g.p.BN)) {
        delete L;
        return 1;
    }
    *opaque = L;
    return 0;
}

int gemm_tc_v2_run(void* opaque, cudaStream_t st) {
    GemmV2Launch* L = static_cast<GemmV2Launch*>(opaque);
    return gemm_tc_v2_launch(L->tmA, L->tmB, L->g, st);
}

void gemm_tc_v2_free(void* opaque) { delete static_cast<GemmV2Launch*>(opaque); }
int gemm_tc_v2_grid(const void* opaque) {
    const GemmV2& g = static_cast<const GemmV2Launch*>(opaque)->g;
    const int n = (g.mc | g.pair) ? 2 * g.work_items : g.total_tiles;
    return n < 148 ? n : 148;
}

void gemm_tc_v2_describe(const void* opaque, char* out, int cap) {
    const GemmV2& g = static_cast<const GemmV2Launch*>(opaque)->g;
    snprintf(out, (size_t)cap, "M=%d N=%d K=%d taps=%d act=%d res=%d f32=%d s2=%d tr=%d | BN=%d MT=%d slab=%d stages=%d acc=%d tiles=%d pair=%d", g.p.M, g.p.N,
             g.p.Kc * g.p.ntaps, g.p.ntaps, g.p.act, g.p.res ? (g.p.res_ld < 0 ? -1 : 1) : 0, g.p.out_f32, g.p.s2, g.p.transposed, g.p.BN, g.MT, g.slab,
             g.p.stages, g.acc_stages, g.total_tiles, g.pair | (g.mc << 1));
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

int gemm_tc_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p_in, cudaStream_t st) {
    GemmParams p = p_in;
    static int dbg = -1;
    if (dbg < 0) { const char* d = getenv("ADAS_B200_DBG"); dbg = d ? atoi(d) : 0; }
    p.dbg = dbg;
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

int make_tmap_4d_s2(CUtensorMap* tm, const void* base, uint64_t C, uint64_t Wp, uint64_t Hp, uint64_t B, uint64_t ld_elems,
                    uint32_t box_w_src, uint32_t box_h_src) {
    PFN_encodeTiled fn = get_encode_fn();
    ADAS_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
    ADAS_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld_elems * 2) % 16 == 0, "TMA 4-D map alignment");
    ADAS_CHECK(box_w_src <= 256 && box_h_src <= 256, "TMA 4-D box too large (%u x %u)", box_w_src, box_h_src);
    cuuint64_t dims[4] = {C, Wp, Hp, B};
    cuuint64_t strides[3] = {ld_elems * 2, Wp * ld_elems * 2, Hp * Wp * ld_elems * 2};
    cuuint32_t box[4] = {64, box_w_src, box_h_src, 1};
    cuuint32_t estr[4] = {1, 2, 2, 1};          // traversal stride 2 in x and y: the box delivers ceil(box/2) pixels per axis
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ADAS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4-D, stride 2) failed: %d", (int)r);
    return 0;
}

// plain 4-D tiled map over [C, W, H, B] (row pitch Wp pixels, image pitch Hp rows): used for TMA STORES of output patches into the
// interior of a padded NHWC grid (the tensor extent is the interior, so the unit clips partial patches and never touches the halo)
int make_tmap_4d(CUtensorMap* tm, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t B, uint64_t ld_elems, uint64_t Wp, uint64_t Hp,
                 uint32_t box_c, uint32_t box_w, uint32_t box_h) {
    PFN_encodeTiled fn = get_encode_fn();
    ADAS_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
    ADAS_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld_elems * 2) % 16 == 0, "TMA 4-D map alignment");
    ADAS_CHECK(box_w <= 256 && box_h <= 256 && box_c * 2 <= 128, "TMA 4-D box too large (%u x %u x %u)", box_c, box_w, box_h);
    cuuint64_t dims[4] = {C, W, H, B};
    cuuint64_t strides[3] = {ld_elems * 2, Wp * ld_elems * 2, Hp * Wp * ld_elems * 2};
    cuuint32_t box[4] = {box_c, box_w, box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ADAS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4-D) failed: %d (C=%llu W=%llu H=%llu B=%llu)", (int)r, (unsigned long long)C,
               (unsigned long long)W, (unsigned long long)H, (unsigned long long)B);
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
        else if (p.ntaps == 4) shift = (tap - 2) * p.Wp;
        long ar = (long)row + shift;
        if (p.s2) {
            // out row -> (b, yo, xo); input pixel (2*yo+dy, 2*xo+dx) in the input's padded grid
            const int Wop = p.mask_W + 2, img = (p.mask_H + 2) * Wop;
            const int b = row / img, pp = row - b * img;
            const int yo = pp / Wop - 1, xo = pp % Wop - 1;
            const int dy = p.ntaps == 9 ? tap / 3 : 1, dx = p.ntaps == 9 ? tap % 3 : 1;
            ar = ((long)b * p.s2_Hp_in + 2 * yo + dy) * p.Wp + 2 * xo + dx;
        } else if (ar < 0 || ar >= p.M) continue;
        if (ar < 0) continue;
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
