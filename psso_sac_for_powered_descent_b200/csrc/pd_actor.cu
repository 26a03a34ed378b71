// pd_actor.cu - shared-weight SAC actor over a large env batch (BASELINE config 4).
//
// Reference: Actor.forward / Actor.sample, src/agents/sac_pytorch.py:129-179
//     h1 = relu(W1 obs + b1); h2 = relu(W2 h1 + b2); mean = Wm h2 + bm;
//     log_std = clamp(Ws h2 + bs, -20, 2); action = tanh(mean + exp(log_std) * eps) * max_action
// with the script default hidden = 256, two hidden layers (sac_pytorch_powered_descent.py:55-56).
//
// Only the 256x256 layer is a real GEMM (2*65536 flop/env of 133 120).  It runs on the
// 5th-generation tensor cores:
//   * a persistent CTA per SM owns a 128-env row tile at a time (UMMA M = 128, N = 256);
//   * W2 is converted once to fp16 and stored in global memory as the exact shared-memory
//     image (K-major, 128-byte swizzle, four 64-wide K blocks); each CTA pulls the 128 KB image
//     with 1-D TMA bulk copies (cp.async.bulk, completes on an mbarrier) once per launch;
//   * layer 1 (K = 2 or 5: plain FMAs) is computed by the 128 threads straight into the
//     swizzled A tile in shared memory as fp16;
//   * one elected thread issues 16 tcgen05.mma (K = 16 each) accumulating fp32 into 256 TMEM
//     columns, commits to an mbarrier;
//   * the epilogue reads each env's row back with tcgen05.ld (one TMEM lane per thread), applies
//     bias + ReLU, the two 256->A heads (GEMV, fp32 FMAs), Philox noise, tanh - h2 never
//     leaves TMEM/registers.
// An fp32 CUDA-core variant (actor_fp32_kernel) serves other hidden sizes and is the numerics
// reference for the tensor-core path in the GPU tests.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pd_actor.h"

namespace pd {

__device__ __forceinline__ void philox4x32_a(unsigned int c0, unsigned int c1, unsigned int c2,
                                             unsigned int c3, unsigned int k0, unsigned int k1,
                                             unsigned int out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        unsigned int n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// N(0,1) x 4 for (env, step)
__device__ __forceinline__ void normal4(unsigned long long seed, unsigned int env, unsigned int step,
                                        float z[4]) {
    unsigned int r[4];
    philox4x32_a(env, step, 0x41435452u, 0u, (unsigned int)seed, (unsigned int)(seed >> 32), r);
    float u0 = ((r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f), u1 = ((r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float u2 = ((r[2] >> 8) + 0.5f) * (1.0f / 16777216.0f), u3 = ((r[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
    float s, c;
    sincospif(2.0f * u1, &s, &c);
    z[0] = ra * c; z[1] = ra * s;
    sincospif(2.0f * u3, &s, &c);
    z[2] = rb * c; z[3] = rb * s;
}

template <int A>
__device__ __forceinline__ void head_and_sample(const ActorArgs &p, const float *mean, const float *lstd,
                                                unsigned int env, float *act_out, float *mean_out) {
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (!p.deterministic) normal4(p.seed, env, p.step, z);
#pragma unroll
    for (int a = 0; a < A; ++a) {
        float ls = fminf(fmaxf(lstd[a], -20.0f), 2.0f);
        float pre = p.deterministic ? mean[a] : fmaf(expf(ls), z[a], mean[a]);
        act_out[(size_t)env * A + a] = tanhf(pre) * p.max_action;
        if (mean_out) mean_out[(size_t)env * A + a] = mean[a];
    }
}

// ------------------------------------------------------------------ fp32 CUDA-core variant
template <int O, int A>
__global__ void __launch_bounds__(128) actor_fp32_kernel(ActorArgs p, const float *__restrict__ obs,
                                                         float *__restrict__ act, float *mean_out, int B) {
    extern __shared__ float sh[];      // h1[H] per thread would be too big: recompute per tile of k
    const int H = p.hidden;
    int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= B) return;
    float o[O];
#pragma unroll
    for (int k = 0; k < O; ++k) o[k] = obs[(size_t)env * O + k];
    float mean[A], lstd[A];
#pragma unroll
    for (int a = 0; a < A; ++a) { mean[a] = __ldg(p.bm + a); lstd[a] = __ldg(p.bs + a); }
    // h2[n] = relu(b2[n] + sum_k W2[n][k] h1[k]); h1 recomputed on the fly (K <= 5 FMAs each)
    for (int n = 0; n < H; ++n) {
        float acc = __ldg(p.b2 + n);
        for (int k = 0; k < H; ++k) {
            float h = __ldg(p.b1 + k);
#pragma unroll
            for (int j = 0; j < O; ++j) h = fmaf(__ldg(p.w1 + k * O + j), o[j], h);
            acc = fmaf(__ldg(p.w2 + (size_t)n * H + k), fmaxf(h, 0.f), acc);
        }
        float h2 = fmaxf(acc, 0.f);
#pragma unroll
        for (int a = 0; a < A; ++a) {
            mean[a] = fmaf(__ldg(p.wm + a * H + n), h2, mean[a]);
            lstd[a] = fmaf(__ldg(p.ws + a * H + n), h2, lstd[a]);
        }
    }
    head_and_sample<A>(p, mean, lstd, (unsigned)env, act, mean_out);
}

// Operand precision.  The 256 x 256 layer's operands are IEEE half (11-bit significand), not bfloat16
// (8 bits): h1 = relu(.) of normalised observations and the weights of a tanh-headed policy are
// O(1), far inside half's range, so the 8x finer rounding is free - measured max error of the
// pre-tanh mean against torch fp32: ~3e-4 of the activation scale (bfloat16: 2e-3).  fp32 accuracy
// (1e-6) needs a 3-term operand split whose operands (W2 alone: 2 x 128 KB) do not fit the 227 KB of
// shared memory next to the A tile; `fp32_path` selects the exact CUDA-core kernel instead.
// ------------------------------------------------------------------ W2 -> fp16 smem image
// image[kb][n][swizzled 128 B row]: K block kb (64 k's), row n of W2 (N = 256 rows, K-major),
// 16-byte chunk c stored at chunk position c ^ (n & 7)  (UMMA SWIZZLE_128B canonical layout).
__global__ void actor_prep_w2_kernel(const float *__restrict__ w2, __half *__restrict__ img) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;      // one 16-byte chunk (8 elements) each
    if (idx >= 4 * 256 * 8) return;
    int c = idx & 7, n = (idx >> 3) & 255, kb = idx >> 11;
    const float *src = w2 + (size_t)n * 256 + kb * 64 + c * 8;
    __half *dst = img + (size_t)kb * 256 * 64 + (size_t)n * 64 + ((c ^ (n & 7)) * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = __float2half_rn(src[j]);
}

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start address >> 4 in [0,14), LBO in [16,30) (unused for swizzled K-major), SBO >> 4 in
// [32,46) = 1024 B between 8-row groups, version 1 in [46,48), layout type 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D = F32 (bits 4-5 = 1), A = B = F16 (format fields at bits 7-9 and
// 10-12 = 0; 1 would be F16), both K-major, N >> 3 at bit 17, M >> 4 at bit 24: M = 128, N = 256
__device__ __forceinline__ uint32_t umma_idesc_f16_m128_n256() {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_fp16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
          "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
          "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
          "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ tensor-core actor kernel
constexpr int ACT_H = 256;
constexpr int TILE_M = 128;
constexpr uint32_t A_BYTES = TILE_M * ACT_H * 2;        // 64 KB: 4 K-blocks x 16 KB
constexpr uint32_t B_BYTES = ACT_H * ACT_H * 2;         // 128 KB: 4 K-blocks x 32 KB

// 128 x ACT_GROUPS threads = groups of 4 warps.  Row (env) of a thread = tid & 127; group h = tid >> 7
// takes 4 / ACT_GROUPS K blocks of layer 1 and 256 / ACT_GROUPS accumulator columns of the epilogue
// (a warp may only touch TMEM lanes 32 (warp % 4) .. +31, so the groups share lanes and split
// columns).  The tile loop is bound by this CUDA-core work at few warps per SM, not by the MMA
// (profiles/r2_actor_tc_kernel.txt: 2 groups = 8 warps: issue 34 %, tensor pipe 20 %).  The fp32 accumulator is double-buffered in TMEM (2 x 256 of the 512 columns): the
// 16 UMMAs of tile i are issued asynchronously and run on the tensor core while all 8 warps do the
// epilogue of tile i-1; only then does the CTA wait for tile i's commit (its A tile may not be
// overwritten earlier).  Per tile the CTA therefore pays layer 1 + epilogue, each at half the
// former per-thread work, and no longer the MMA latency.
#ifndef PD_ACT_GROUPS
#define PD_ACT_GROUPS 4
#endif
constexpr int ACT_GROUPS = PD_ACT_GROUPS;            // column groups of 4 warps: 2 (256 threads) or 4 (512 threads)
constexpr int ACT_THREADS = 128 * ACT_GROUPS;
constexpr int ACT_COLS = ACT_H / ACT_GROUPS;         // accumulator columns per group
constexpr int ACT_KB = 4 / ACT_GROUPS;               // layer-1 K blocks per group

template <int O, int A>
__global__ void __launch_bounds__(ACT_THREADS, 1)
actor_tc_kernel(ActorArgs p, const float *__restrict__ obs, float *__restrict__ act, float *mean_out,
                int B, int n_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                                  // [4][128][64] fp16, swizzled
    uint8_t *sB = smem + A_BYTES;                        // [4][256][64] fp16, swizzled (W2 image)
    // Per-hidden-unit records, padded to float4 so that a (warp-uniform, broadcast) LDS.128 brings a
    // whole record: layer 1 {b1[k], w1[k][0..O)} and epilogue {b2[n], wm[0..A)[n], ws[0..A)[n]}.  With
    // scalar loads these were 3 LDS per accumulator column and per hidden unit - 46 % + 27 % of the
    // kernel's instructions sat in those two loops (profiles/r2_actor_tc_kernel.txt).
    constexpr int L1W = (O + 1 + 3) / 4 * 4, EPW = (1 + 2 * A + 3) / 4 * 4;
    float *s_l1 = (float *)(smem + A_BYTES + B_BYTES);   // [256][L1W]
    float *s_ep = s_l1 + ACT_H * L1W;                    // [256][EPW]
    float *s_part = s_ep + ACT_H * EPW;                  // [ACT_GROUPS - 1][128][2A] head partial sums of groups 1..
    uint64_t *bars = (uint64_t *)(s_part + (ACT_GROUPS - 1) * TILE_M * 2 * A);   // [0] W2 landed, [1] MMA done
    uint32_t *s_tmem = (uint32_t *)(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (TILE_M - 1), half = tid >> 7;
    const uint32_t bar_w2 = smem_u32(bars), bar_mma = smem_u32(bars + 1);

    if (tid == 0) {
        mbar_init(bar_w2, 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {      // one warp allocates all 512 TMEM columns: two 128 x 256 fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(s_tmem)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int k = tid; k < ACT_H; k += ACT_THREADS) {
        s_l1[k * L1W] = p.b1[k];
#pragma unroll
        for (int q = 0; q < O; ++q) s_l1[k * L1W + 1 + q] = p.w1[k * O + q];
#pragma unroll
        for (int q = O + 1; q < L1W; ++q) s_l1[k * L1W + q] = 0.f;
        s_ep[k * EPW] = p.b2[k];
#pragma unroll
        for (int a = 0; a < A; ++a) { s_ep[k * EPW + 1 + a] = p.wm[a * ACT_H + k]; s_ep[k * EPW + 1 + A + a] = p.ws[a * ACT_H + k]; }
#pragma unroll
        for (int q = 1 + 2 * A; q < EPW; ++q) s_ep[k * EPW + q] = 0.f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;

    if (tid == 0) {       // W2 image: 8 TMA bulk copies of 16 KB, one mbarrier
        mbar_expect_tx(bar_w2, B_BYTES);
        const uint8_t *src = (const uint8_t *)p.w2_img;
#pragma unroll
        for (int i = 0; i < 8; ++i) tma_bulk_g2s(smem_u32(sB) + i * 16384, src + (size_t)i * 16384, 16384, bar_w2);
    }
    const uint32_t idesc = umma_idesc_f16_m128_n256();
    uint32_t mma_phase = 0;
    bool w2_ready = false;

    // epilogue of one finished tile: bias + ReLU on this half's 128 accumulator columns, partial
    // head dot products, combine the halves through shared memory, sample
    auto epilogue = [&](int tile, uint32_t buf) {
        float mean[A], lstd[A];
#pragma unroll
        for (int a = 0; a < A; ++a) { mean[a] = 0.f; lstd[a] = 0.f; }
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + buf * 256 + half * ACT_COLS;
#pragma unroll 1
        for (int cb = 0; cb < ACT_COLS / 32; ++cb) {
            uint32_t v[32];
            tmem_ld32(lane_base + cb * 32, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = half * ACT_COLS + cb * 32 + j;
                float rec[EPW];
#pragma unroll
                for (int q = 0; q < EPW; q += 4)
                    *reinterpret_cast<float4 *>(rec + q) = *reinterpret_cast<const float4 *>(s_ep + n * EPW + q);
                const float h2 = fmaxf(__uint_as_float(v[j]) + rec[0], 0.f);
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    mean[a] = fmaf(rec[1 + a], h2, mean[a]);
                    lstd[a] = fmaf(rec[1 + A + a], h2, lstd[a]);
                }
            }
        }
        if (half > 0) {
            float *dst = s_part + ((half - 1) * TILE_M + row) * 2 * A;
#pragma unroll
            for (int a = 0; a < A; ++a) { dst[a] = mean[a]; dst[A + a] = lstd[a]; }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();      // partial sums visible; all TMEM reads of this buffer are done
        if (half == 0) {
            const int env = tile * TILE_M + row;
#pragma unroll
            for (int a = 0; a < A; ++a) {
#pragma unroll
                for (int g = 0; g < ACT_GROUPS - 1; ++g) {
                    mean[a] += s_part[(g * TILE_M + row) * 2 * A + a];
                    lstd[a] += s_part[(g * TILE_M + row) * 2 * A + A + a];
                }
                mean[a] += p.bm[a];
                lstd[a] += p.bs[a];
            }
            if (env < B) head_and_sample<A>(p, mean, lstd, (unsigned)env, act, mean_out);
        }
    };

    int prev_tile = -1;
    uint32_t buf = 0;
    // the observations of the next tile are fetched while this tile's epilogue runs
    float o_next[O];
    {
        const int env0 = blockIdx.x * TILE_M + row;
#pragma unroll
        for (int k = 0; k < O; ++k) o_next[k] = (blockIdx.x < n_tiles && env0 < B) ? obs[(size_t)env0 * O + k] : 0.f;
    }
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        float o[O];
#pragma unroll
        for (int k = 0; k < O; ++k) o[k] = o_next[k];
        // layer 1 -> fp16 -> swizzled A tile (row = tid & 127), this half's two K blocks
        const uint32_t row_off = (uint32_t)(row >> 3) * 1024 + (uint32_t)(row & 7) * 128;
#pragma unroll 1
        for (int kq = 0; kq < ACT_KB; ++kq) {
            const int kb = half * ACT_KB + kq;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint32_t packed[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float h[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int k = kb * 64 + c * 8 + j * 2 + e;
                        float rec[L1W];
#pragma unroll
                        for (int q = 0; q < L1W; q += 4)
                            *reinterpret_cast<float4 *>(rec + q) = *reinterpret_cast<const float4 *>(s_l1 + k * L1W + q);
                        float acc = rec[0];
#pragma unroll
                        for (int q = 0; q < O; ++q) acc = fmaf(rec[1 + q], o[q], acc);
                        h[e] = fmaxf(acc, 0.f);
                    }
                    __half2 b2v = __floats2half2_rn(h[0], h[1]);
                    packed[j] = *reinterpret_cast<uint32_t *>(&b2v);
                }
                uint8_t *dst = sA + kb * 16384 + row_off + ((c ^ (row & 7)) * 16);
                *reinterpret_cast<uint4 *>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
        }
        // generic-proxy smem writes -> visible to the tensor core (async proxy)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            if (!w2_ready) mbar_wait(bar_w2, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {      // UMMA_K = 16 fp16 = 32 bytes inside the 128 B atom
                    uint64_t ad = umma_desc_sw128(smem_u32(sA) + kb * 16384 + k * 32);
                    uint64_t bd = umma_desc_sw128(smem_u32(sB) + kb * 32768 + k * 32);
                    umma_fp16(tmem_base + buf * 256, ad, bd, idesc, (kb | k) ? 1u : 0u);
                }
            }
            umma_commit(bar_mma);
        }
        w2_ready = true;
        {
            const int nt = tile + gridDim.x, envn = nt * TILE_M + row;
#pragma unroll
            for (int k = 0; k < O; ++k) o_next[k] = (nt < n_tiles && envn < B) ? obs[(size_t)envn * O + k] : 0.f;
        }
        // the previous tile's epilogue runs while the tensor core works on this tile
        if (prev_tile >= 0) epilogue(prev_tile, buf ^ 1u);
        mbar_wait(bar_mma, mma_phase);      // tile done: the A tile may be rewritten
        mma_phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        prev_tile = tile;
        buf ^= 1u;
    }
    if (prev_tile >= 0) epilogue(prev_tile, buf ^ 1u);
    if (!w2_ready && tid == 0) mbar_wait(bar_w2, 0);      // never leave with a TMA in flight
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

size_t actor_tc_smem_bytes(int O, int A) {
    const int l1w = (O + 1 + 3) / 4 * 4, epw = (1 + 2 * A + 3) / 4 * 4;
    return 1024 + A_BYTES + B_BYTES + sizeof(float) * (ACT_H * l1w + ACT_H * epw + (ACT_GROUPS - 1) * TILE_M * 2 * A) + 64;
}

int actor_prep_w2(const float *w2, void *img, cudaStream_t st) {
    actor_prep_w2_kernel<<<(4 * 256 * 8 + 255) / 256, 256, 0, st>>>(w2, (__half *)img);
    return cudaGetLastError() != cudaSuccess;
}

template <int O, int A>
static int launch_t(const ActorArgs &p, const float *obs, float *act, float *mean_out, int B, int use_tc,
                    int n_sm, cudaStream_t st) {
    if (use_tc) {
        size_t smem = actor_tc_smem_bytes(O, A);
        static bool attr_set = false;
        if (!attr_set) {
            if (cudaFuncSetAttribute(actor_tc_kernel<O, A>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem) != cudaSuccess) return 1;
            attr_set = true;
        }
        int n_tiles = (B + TILE_M - 1) / TILE_M;
        int grid = n_tiles < n_sm ? n_tiles : n_sm;
        actor_tc_kernel<O, A><<<grid, ACT_THREADS, smem, st>>>(p, obs, act, mean_out, B, n_tiles);
    } else {
        actor_fp32_kernel<O, A><<<(B + 127) / 128, 128, 0, st>>>(p, obs, act, mean_out, B);
    }
    return cudaGetLastError() != cudaSuccess;
}

int actor_launch(int O, int A, const ActorArgs &p, const float *obs, float *act, float *mean_out, int B,
                 int use_tc, int n_sm, cudaStream_t st) {
    if (O == 2 && A == 1) return launch_t<2, 1>(p, obs, act, mean_out, B, use_tc, n_sm, st);
    if (O == 5 && A == 4) return launch_t<5, 4>(p, obs, act, mean_out, B, use_tc, n_sm, st);
    if (O == 8 && A == 2) return launch_t<8, 2>(p, obs, act, mean_out, B, use_tc, n_sm, st);   // ascent
    if (O == 4 && A == 1) return launch_t<4, 1>(p, obs, act, mean_out, B, use_tc, n_sm, st);   // ballistic arc
    if (O == 1 && A == 1) return launch_t<1, 1>(p, obs, act, mean_out, B, use_tc, n_sm, st);   // P-control
    return 1;
}

}  // namespace pd
