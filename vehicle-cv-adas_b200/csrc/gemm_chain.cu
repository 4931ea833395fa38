// gemm_chain.cu -- a CHAIN of same-shape stride-1 conv layers in one persistent launch.
//
// Replaces: the same opaque conv stacks as gemm_v3.cu (coreEngine.py:150-157 / :184-186); this is the launch-count side of it.
//
// Why (profiles/r02_*): at batch 8 a YOLOv8l + UFLDv2 step is 139 GEMM launches of ~24 us; every launch pays ~1.8 us of launch
// skeleton, ~1.2 us until its first operands arrive and a ~2 us epilogue tail during which the tensor pipe is idle -- about a
// quarter of the step.  The bottleneck convs of a C2f block (ultralytics `Bottleneck`: cv1 -> cv2 (+x)) and the BasicBlocks of a ResNet
// stage (torchvision: conv1 -> conv2 (+identity)) are runs of IDENTICAL 3x3 GEMMs, each reading what the previous one wrote.  This
// kernel runs such a run as one launch:
//   * work items = (layer, tile), claimed in order from a global atomic counter (dynamic scheduling), so an item only ever waits
//     for items claimed before it by CTAs that are running -- dead-lock free whatever number of CTAs is resident;
//   * per layer and 128-row block a completion counter in global memory: a tile of layer l+1 is loaded once the row blocks it reads
//     (its own rows +- one padded image row) have been stored by ALL column tiles of layer l.  The residual of a layer (the input of the
//     previous layer in both network families) needs no counter of its own: the previous layer's tiles over the same rows had
//     waited for it;
//   * TMA stores -> cp.async.bulk.wait_group -> fence.proxy.async + __threadfence -> atomicAdd(release) on the producer side,
//     ld.acquire.gpu -> fence.proxy.async on the consumer side order the async-proxy writes before the async-proxy reads;
//   * everything else (operand ring, TMEM double buffering, 16-warp staged epilogue, residual through TMA) is gemm_v3.cu's; the
//     accumulator / staging / ring state simply carries over from one layer to the next, so the epilogue of a layer's last tile
//     overlaps the main loop of the next layer's first tile.
// Per-layer tensor maps live in a device array (ChainLayer), the tile shape is common to the chain.
#include "common.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>
#include "gemm_v3.h"

namespace adas {

struct alignas(128) ChainLayer {
    CUtensorMap tmA, tmB, tmC, tmR;
    const float* bias;
    int res_mode;        // 0: none, +1: residual added after the activation (YOLO shortcut), -1: before it (ResNet)
    int pad_[29];
};
static_assert(sizeof(ChainLayer) == 640, "ChainLayer layout");

static constexpr int SCHED_SLOTS = 2;

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ctl[0] = next item to claim, ctl[1] = CTAs that have finished, ctl[2] = epoch (launches completed so far)
__global__ void __launch_bounds__(V3_THREADS, 1)
conv_chain_v3_kernel(const GemmV3 g, const ChainLayer* __restrict__ layers, int n_layers, uint32_t* __restrict__ flags, uint32_t* __restrict__ ctl,
                     int n_rb) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t tfull_bar[2];
    __shared__ __align__(8) uint64_t tempty_bar[2];
    __shared__ __align__(8) uint64_t res_bar[V3_STG_BUFS];
    __shared__ __align__(8) uint64_t sched_full[SCHED_SLOTS];
    __shared__ __align__(8) uint64_t sched_empty[SCHED_SLOTS];
    __shared__ int sched_item[SCHED_SLOTS];
    __shared__ uint32_t tmem_holder;
    __shared__ __align__(16) float s_bias[2][256];

    const GemmParams& p = g.p;
    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int stages = g.stages;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int taps_per_step = g.slab ? 3 : 1;
    const int BMT = BM * g.MT;
    const int mt_cols = g.sub_cols;
    const int acc_stride = g.MT * g.sub_cols;
    const bool acc2 = g.acc_stages == 2;
    const int total_items = n_layers * g.total_tiles;

    if (warp_idx == 0 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tfull_bar[s]), 1);
            mbar_init(smem_u32(&tempty_bar[s]), V3_EPI_WARPS);
        }
        for (int s = 0; s < V3_STG_BUFS; ++s) mbar_init(smem_u32(&res_bar[s]), 1);
        for (int s = 0; s < SCHED_SLOTS; ++s) {
            mbar_init(smem_u32(&sched_full[s]), 1);
            mbar_init(smem_u32(&sched_empty[s]), 1 + V3_EPI_WARPS);       // MMA warp + every epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp_idx == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_holder;
    if (g.pdl) {
        asm volatile("griddepcontrol.wait;" ::: "memory");          // the previous kernel (incl. the previous launch of this chain) is complete
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    // The tensor maps of a chain live in GLOBAL memory (one set per layer), not in kernel parameters.  A chain object that is freed and
    // another one allocated at the same address (the build-time comparison of candidate chains does exactly that) changes descriptors
    // under the SMs' descriptor caches; the tensormap proxy fence makes every CTA re-read them.  (Hypothesis for the rare device hang
    // of DESIGN.md 4.2: a stale weight-tile descriptor with a different box height delivers fewer bytes than the mbarrier expects.)
    for (int i = threadIdx.x; i < n_layers * 4; i += blockDim.x) {
        const ChainLayer* Lf = layers + (i >> 2);
        const void* tm = (i & 3) == 0 ? (const void*)&Lf->tmA : (i & 3) == 1 ? (const void*)&Lf->tmB : (i & 3) == 2 ? (const void*)&Lf->tmC : (const void*)&Lf->tmR;
        asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
    }
    __syncthreads();
    const uint32_t epoch = ld_acquire_u32(ctl + 2);                // counters of all earlier launches are already in `flags`
    const uint32_t flag_target = (epoch + 1u) * (uint32_t)g.n_tiles;

    if (warp_idx == 0) {
        if (lane == 0) {
            // ================= scheduler + TMA producer =================
            const uint32_t a_bytes = (uint32_t)(g.slab ? V3_SLAB_BYTES : A_STAGE_BYTES);
            const uint32_t tx_bytes = (uint32_t)g.MT * a_bytes + (uint32_t)(taps_per_step * p.BN * BK * 2);
            const int n_grp = g.slab ? 3 : p.ntaps;
            const bool dx_inner = (!g.slab && p.ntaps == 9);
            const int o_cnt = dx_inner ? 3 : n_grp;
            const int i_cnt = dx_inner ? 3 : 1;
            const int halo = p.ntaps == 9 ? p.Wp + 1 : 0;
            uint32_t s = 0, ph = 0, ss = 0, sph = 0;
            int item = (int)atomicAdd(ctl, 1u);
            while (true) {
                const int l = item / g.total_tiles, t = item - l * g.total_tiles;
                const int n_t = t % g.n_tiles, m_t = t / g.n_tiles;
                const int n0 = n_t * p.BN, m0 = m_t * BMT;
                const ChainLayer* L = layers + (item < total_items ? l : 0);
                if (item < total_items && l > 0) {
                    // the row blocks this tile reads have been stored by every column tile of the previous layer
                    const uint32_t* fl = flags + (size_t)(l - 1) * n_rb;
                    int lo = (m0 - halo) / BM, hi = (m0 + BMT - 1 + halo) / BM;
                    if (m0 - halo < 0) lo = 0;
                    if (hi > n_rb - 1) hi = n_rb - 1;
                    for (int rb = lo; rb <= hi; ++rb)
                        while ((int32_t)(ld_acquire_u32(fl + rb) - flag_target) < 0) __nanosleep(32);
                    fence_proxy_async_all();
                }
                // hand the item to the MMA and epilogue warps -- only now: the epilogue requests the residual tile (the previous layer's
                // input) as soon as it sees the item, and that data is only guaranteed once this tile's row blocks are complete
                mbar_wait(smem_u32(&sched_empty[ss]), sph ^ 1u);
                *reinterpret_cast<volatile int*>(&sched_item[ss]) = item;
                mbar_arrive(smem_u32(&sched_full[ss]));
                if (++ss == SCHED_SLOTS) { ss = 0; sph ^= 1u; }
                if (item >= total_items) break;
                const int next = (int)atomicAdd(ctl, 1u);            // claimed now, so its latency hides behind this item's loads
                for (int o = 0; o < o_cnt; ++o) {
                    for (int kc = 0; kc < p.kpt; ++kc) {
                        for (int in = 0; in < i_cnt; ++in) {
                            const int grp = dx_inner ? o * 3 + in : o;
                            mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
                            const uint32_t fb = smem_u32(&full_bar[s]);
                            mbar_expect_tx(fb, tx_bytes);
                            const uint32_t a_dst = smem_base + s * g.stage_bytes;
                            const uint32_t b_dst = a_dst + g.MT * g.a_sub_bytes;
                            if (g.slab) {
                                const int r0 = m0 + (grp - 1) * p.Wp - 1;
                                for (int mt = 0; mt < g.MT; ++mt) tma_load_2d(a_dst + mt * g.a_sub_bytes, &L->tmA, kc * BK, r0 + mt * BM, fb);
                                for (int dx = 0; dx < 3; ++dx) tma_load_2d(b_dst + dx * g.b_bytes, &L->tmB, (grp * 3 + dx) * p.Kc + kc * BK, n0, fb);
                            } else {
                                int shift = 0;
                                if (p.ntaps == 9) shift = (grp / 3 - 1) * p.Wp + (grp % 3 - 1);
                                for (int mt = 0; mt < g.MT; ++mt) tma_load_2d(a_dst + mt * g.a_sub_bytes, &L->tmA, kc * BK, m0 + shift + mt * BM, fb);
                                tma_load_2d(b_dst, &L->tmB, grp * p.Kc + kc * BK, n0, fb);
                            }
                            if (++s == (uint32_t)stages) { s = 0; ph ^= 1u; }
                        }
                    }
                }
                item = next;
            }
        }
    } else if (warp_idx == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = (1u << 4) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        const uint32_t a_step = (uint32_t)g.a_sub_bytes >> 4, b_step = (uint32_t)g.b_bytes >> 4;
        const int n_dx = taps_per_step, n_mt = g.MT;
        const int ksteps = (g.slab ? 3 : p.ntaps) * p.kpt;
        uint32_t s = 0, ph = 0, tile_it = 0, ss = 0, sph = 0;
        while (true) {
            mbar_wait(smem_u32(&sched_full[ss]), sph);
            const int item = *reinterpret_cast<volatile int*>(&sched_item[ss]);
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sched_empty[ss]));
            if (++ss == SCHED_SLOTS) { ss = 0; sph ^= 1u; }
            if (item >= total_items) break;
            const int as = acc2 ? (int)(tile_it & 1) : 0;
            mbar_wait(smem_u32(&tempty_bar[as]), ((acc2 ? (tile_it >> 1) : tile_it) & 1u) ^ 1u);
            tcgen05_fence_after();
            const uint32_t d_base = tmem_base + (uint32_t)(as * acc_stride);
            for (int ks = 0; ks < ksteps; ++ks) {
                mbar_wait(smem_u32(&full_bar[s]), ph);
                tcgen05_fence_after();
                const uint32_t a_lo = ((smem_base + s * g.stage_bytes) & 0x3FFFFu) >> 4;
                const uint32_t b_lo = a_lo + (uint32_t)n_mt * a_step;
                if (elect_one()) {
                    for (int dx = 0; dx < n_dx; ++dx) {
                        for (int mt = 0; mt < n_mt; ++mt) {
                            const uint32_t a_sub = a_lo + (uint32_t)mt * a_step + (g.slab ? (uint32_t)dx * 8u : 0u);
                            const uint32_t b_sub = b_lo + (uint32_t)dx * b_step;
                            const uint32_t d = d_base + (uint32_t)(mt * mt_cols);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                const uint64_t ad = desc_hi | (uint64_t)(a_sub + 2u * k);
                                const uint64_t bd = desc_hi | (uint64_t)(b_sub + 2u * k);
                                umma_f16(d, ad, bd, idesc, (uint32_t)((ks | dx | k) != 0));
                            }
                        }
                    }
                    umma_commit(smem_u32(&empty_bar[s]));
                }
                __syncwarp();
                if (++s == (uint32_t)stages) { s = 0; ph ^= 1u; }
            }
            if (elect_one()) umma_commit(smem_u32(&tfull_bar[as]));
            __syncwarp();
            ++tile_it;
        }
    } else {
        // ================= epilogue (16 warps), staged TMA-store path of gemm_v3.cu =================
        const int ew = warp_idx - 2;
        const int q = warp_idx & 3;
        const int part = ew >> 2;
        const int et = threadIdx.x - 64;
        const int r = q * 32 + lane;
        const bool issuer = (et == 0);
        const uint32_t stg_base = smem_base + (uint32_t)g.stg_off;
        const int n_chunks = p.BN >> 6;
        const int chunks_per_tile = g.MT * n_chunks;
        uint32_t tile_it = 0, chunk_it = 0, res_it = 0, ss = 0, sph = 0;
        while (true) {
            mbar_wait(smem_u32(&sched_full[ss]), sph);
            const int item = *reinterpret_cast<volatile int*>(&sched_item[ss]);
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&sched_empty[ss]));
            if (++ss == SCHED_SLOTS) { ss = 0; sph ^= 1u; }
            if (item >= total_items) break;
            const int l = item / g.total_tiles, t = item - l * g.total_tiles;
            const int n_t = t % g.n_tiles, m_t = t / g.n_tiles;
            const int n0 = n_t * p.BN, m0 = m_t * BMT;
            const ChainLayer* L = layers + l;
            const int res_mode = L->res_mode;
            const int as = acc2 ? (int)(tile_it & 1) : 0;
            const int bs = (int)(tile_it & 1);
            for (int j = et; j < p.BN; j += 32 * V3_EPI_WARPS) s_bias[bs][j] = (L->bias != nullptr && (n0 + j) < p.N) ? __ldg(L->bias + n0 + j) : 0.f;
            // residual tiles of the first two chunks: requested while the main loop of this tile is still running (every earlier
            // store has completed -- the issuer waited for all of them before it signalled the previous tile)
            const uint32_t res_base = res_it, chunk_base = chunk_it;       // counters at the start of this tile
            auto issue_res = [&](int c_of_tile) {
                const int mt2 = c_of_tile / n_chunks, cc2 = c_of_tile - mt2 * n_chunks;
                const uint32_t bar = smem_u32(&res_bar[(res_base + (uint32_t)c_of_tile) % V3_STG_BUFS]);
                const uint32_t dst = stg_base + ((chunk_base + (uint32_t)c_of_tile) % V3_STG_BUFS) * (uint32_t)V3_STG_BYTES;
                mbar_expect_tx(bar, (uint32_t)V3_STG_BYTES);
                tma_load_2d(dst, &L->tmR, n0 + cc2 * 64, m0 + mt2 * BM, bar);
            };
            if (res_mode != 0 && issuer) {
                fence_proxy_async_all();          // ordered after the producer warp's acquire of this tile's row-block counters
                issue_res(0);
                if (chunks_per_tile > 1) issue_res(1);
            }
            uint32_t okmask = 0;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                if (mt < g.MT) {
                    const int row = m0 + mt * BM + r;
                    bool ok = row < p.M;
                    if (p.mask_H > 0 && ok) {
                        const int Wp = p.mask_W + 2;
                        const int pp = row - fast_div(row, g.fd_img) * g.fd_img.d;
                        const int yy = fast_div(pp, g.fd_wp);
                        const int xx = pp - yy * Wp;
                        ok = (yy >= 1) && (yy <= p.mask_H) && (xx >= 1) && (xx <= p.mask_W);
                    }
                    okmask |= ok ? (1u << mt) : 0u;
                }
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            mbar_wait(smem_u32(&tfull_bar[as]), (acc2 ? (tile_it >> 1) : tile_it) & 1u);
            tcgen05_fence_after();
            int c_local = 0;
            for (int mt = 0; mt < g.MT; ++mt) {
                const bool row_ok = (okmask >> mt) & 1u;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * acc_stride + mt * mt_cols);
                for (int cc = 0; cc < n_chunks; ++cc, ++c_local) {
                    const int c = cc * 64 + part * 16;
                    const bool last_ld = (mt == g.MT - 1) && (cc == n_chunks - 1);
                    uint32_t v[16];
                    tmem_ld16(taddr + (uint32_t)c, v);
                    uint4 rr[2];
                    const uint32_t buf = chunk_it % V3_STG_BUFS;
                    const uint32_t stg = stg_base + buf * (uint32_t)V3_STG_BYTES + (uint32_t)r * 128u;
                    const uint32_t sw = (uint32_t)(r & 7);
                    const uint32_t slot0 = stg + ((((uint32_t)(2 * part)) ^ sw) << 4), slot1 = stg + ((((uint32_t)(2 * part + 1)) ^ sw) << 4);
                    if (res_mode != 0) {
                        mbar_wait(smem_u32(&res_bar[res_it % V3_STG_BUFS]), (res_it / V3_STG_BUFS) & 1u);
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(rr[0].x), "=r"(rr[0].y), "=r"(rr[0].z), "=r"(rr[0].w) : "r"(slot0));
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(rr[1].x), "=r"(rr[1].y), "=r"(rr[1].z), "=r"(rr[1].w) : "r"(slot1));
                    }
                    tmem_ld_wait();
                    if (last_ld) {
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[as]));
                    }
                    float f[16];
                    const float4* sb4 = reinterpret_cast<const float4*>(&s_bias[bs][c & 255]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 b4 = sb4[j];
                        f[4 * j] = __uint_as_float(v[4 * j]) + b4.x; f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
                        f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z; f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
                    }
                    if (res_mode < 0) {
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const __half2* h = reinterpret_cast<const __half2*>(&rr[k]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) { float2 tt = __half22float2(h[j]); f[k * 8 + 2 * j] += tt.x; f[k * 8 + 2 * j + 1] += tt.y; }
                        }
                    }
                    if (p.act == 1) {
#pragma unroll
                        for (int j = 0; j < 16; j += 2) silu2(f[j], f[j + 1]);
                    } else if (p.act == 2) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
                    }
                    if (res_mode > 0) {
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const __half2* h = reinterpret_cast<const __half2*>(&rr[k]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) { float2 tt = __half22float2(h[j]); f[k * 8 + 2 * j] += tt.x; f[k * 8 + 2 * j + 1] += tt.y; }
                        }
                    }
                    uint32_t o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const __half2 h = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
                        o[j] = row_ok ? *reinterpret_cast<const uint32_t*>(&h) : 0u;
                    }
                    st_shared_v4(slot0, o[0], o[1], o[2], o[3]);
                    st_shared_v4(slot1, o[4], o[5], o[6], o[7]);
                    fence_async_smem();
                    if (issuer) bulk_wait_read1();          // stores <= i-2 have left their buffers: (i+1) % 3 may be rewritten after the barrier
                    __syncwarp();
                    asm volatile("bar.sync 1, 512;" ::: "memory");
                    if (issuer) {
                        if (n0 + cc * 64 < p.N && m0 + mt * BM < p.M) tma_store_2d(&L->tmC, stg_base + buf * (uint32_t)V3_STG_BYTES, n0 + cc * 64, m0 + mt * BM);
                        bulk_commit();
                        if (res_mode != 0 && c_local + 2 < chunks_per_tile) {
                            bulk_wait_read1();              // store (i-1) has left buffer (i+2) % 3
                            issue_res(c_local + 2);
                        }
                    }
                    __syncwarp();
                    ++chunk_it;
                    if (res_mode != 0) ++res_it;
                }
            }
            // publish the tile: all of its rows are in global memory before the row-block counters move
            if (issuer) {
                bulk_wait_all();
                fence_proxy_async_all();
                __threadfence();
                uint32_t* fl = flags + (size_t)l * n_rb;
                for (int mt = 0; mt < g.MT; ++mt) {
                    const int rb = m_t * g.MT + mt;
                    if (rb < n_rb) atomicAdd(fl + rb, 1u);
                }
            }
            __syncwarp();
            ++tile_it;
        }
        if (issuer) bulk_wait_all();
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
    // the last CTA to leave re-arms the scheduler for the next launch of this chain (flags keep counting: epoch + 1)
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t done = atomicAdd(ctl + 1, 1u) + 1u;
        if (done == gridDim.x) {
            ctl[0] = 0u;
            ctl[1] = 0u;
            __threadfence();
            atomicAdd(ctl + 2, 1u);
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
struct GemmChain {
    GemmV3 g;
    int n_layers = 0, n_rb = 0;
    ChainLayer* d_layers = nullptr;
    uint32_t* d_flags = nullptr;      // [n_layers][n_rb]
    uint32_t* d_ctl = nullptr;        // [4]
};

static std::mutex g_chain_mu;
static bool g_chain_attr[64] = {false};

int gemm_chain_prepare(void* const* layer_opaques, int n_layers, void** out) {
    ADAS_CHECK(n_layers >= 2 && n_layers <= 64, "gemm_chain: %d layers", n_layers);
    std::vector<ChainLayer> hl((size_t)n_layers);
    const GemmV3Launch* L0 = static_cast<const GemmV3Launch*>(layer_opaques[0]);
    for (int i = 0; i < n_layers; ++i) {
        const GemmV3Launch* L = static_cast<const GemmV3Launch*>(layer_opaques[i]);
        const GemmV3 &a = L0->g, &b = L->g;
        ADAS_CHECK(b.tma_st && !b.p.s2 && !b.p.transposed && a.p.M == b.p.M && a.p.N == b.p.N && a.p.Kc == b.p.Kc && a.p.ntaps == b.p.ntaps &&
                   a.p.BN == b.p.BN && a.MT == b.MT && a.slab == b.slab && a.stages == b.stages && a.stage_bytes == b.stage_bytes &&
                   a.stg_off == b.stg_off && a.p.act == b.p.act && a.p.Wp == b.p.Wp && a.p.mask_H == b.p.mask_H && a.p.mask_W == b.p.mask_W,
                   "gemm_chain: layer %d does not share the chain's shape / tile configuration", i);
        ADAS_CHECK(b.p.res == nullptr || b.res_tma, "gemm_chain: layer %d has a residual that cannot go through TMA", i);
        memset(&hl[i], 0, sizeof(ChainLayer));
        hl[i].tmA = L->tmA; hl[i].tmB = L->tmB; hl[i].tmC = L->tmC; hl[i].tmR = L->tmR;
        hl[i].bias = b.p.bias;
        hl[i].res_mode = b.p.res == nullptr ? 0 : (b.p.res_ld < 0 ? -1 : 1);
    }
    GemmChain* c = new GemmChain();
    c->g = L0->g;
    c->n_layers = n_layers;
    c->n_rb = (L0->g.p.M + BM - 1) / BM;
    ADAS_CUDA(cudaMalloc(&c->d_layers, sizeof(ChainLayer) * (size_t)n_layers));
    ADAS_CUDA(cudaMemcpy(c->d_layers, hl.data(), sizeof(ChainLayer) * (size_t)n_layers, cudaMemcpyHostToDevice));
    ADAS_CUDA(cudaMalloc(&c->d_flags, sizeof(uint32_t) * (size_t)n_layers * c->n_rb));
    ADAS_CUDA(cudaMemset(c->d_flags, 0, sizeof(uint32_t) * (size_t)n_layers * c->n_rb));
    ADAS_CUDA(cudaMalloc(&c->d_ctl, 16));
    ADAS_CUDA(cudaMemset(c->d_ctl, 0, 16));
    *out = c;
    return 0;
}

int gemm_chain_run(void* opaque, cudaStream_t st) {
    GemmChain* c = static_cast<GemmChain*>(opaque);
    int num_sms = 0;
    if (v3_num_sms(&num_sms)) return 1;
    {
        int dev = 0;
        ADAS_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(g_chain_mu);
        if (dev >= 0 && dev < 64 && !g_chain_attr[dev]) {
            ADAS_CUDA(cudaFuncSetAttribute(conv_chain_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V3_DYN_SMEM_MAX));
            g_chain_attr[dev] = true;
        }
    }
    static const bool pdl = !(getenv("ADAS_B200_PDL") && getenv("ADAS_B200_PDL")[0] == '0');
    GemmV3 gp = c->g;
    gp.pdl = pdl ? 1 : 0;
    const int items = c->n_layers * gp.total_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(items < num_sms ? items : num_sms, 1, 1);
    cfg.blockDim = dim3(V3_THREADS, 1, 1);
    cfg.dynamicSmemBytes = gp.stages * gp.stage_bytes + V3_STG_BUFS * V3_STG_BYTES + 1024;
    cfg.stream = st;
    cudaLaunchAttribute attr1;
    attr1.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr1.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr1;
    cfg.numAttrs = pdl ? 1 : 0;
    ADAS_CUDA(cudaLaunchKernelEx(&cfg, conv_chain_v3_kernel, gp, (const ChainLayer*)c->d_layers, c->n_layers, c->d_flags, c->d_ctl, c->n_rb));
    count_launch();
    return 0;
}

void gemm_chain_free(void* opaque) {
    GemmChain* c = static_cast<GemmChain*>(opaque);
    if (!c) return;
    cudaFree(c->d_layers); cudaFree(c->d_flags); cudaFree(c->d_ctl);
    delete c;
}

void gemm_chain_describe(const void* opaque, char* out, int cap) {
    const GemmChain* c = static_cast<const GemmChain*>(opaque);
    const GemmV3& g = c->g;
    snprintf(out, (size_t)cap, "M=%d N=%d K=%d taps=%d act=%d | chain of %d layers BN=%d MT=%d slab=%d stages=%d acc=%d tiles=%d", g.p.M, g.p.N * c->n_layers,
             g.p.Kc * g.p.ntaps, g.p.ntaps, g.p.act, c->n_layers, g.p.BN, g.MT, g.slab, g.stages, g.acc_stages, g.total_tiles);
}

}  // namespace adas
