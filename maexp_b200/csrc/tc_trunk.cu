// Conv trunk of the acting / evaluation batches and of the C51 / DQN update passes (TRAIN keeps the pooled map and the pooling
// argmax for the backward pass) with the second convolution on the 5th-generation tensor cores
// (core/agents/models/customs/convolution.py: conv 2x2 -> max_pool2d(2, 2) -> conv 2x2 -> flatten; visible_range 7, 32 and
// 64 filters).  conv2 is 70 % of the trunk's arithmetic and, as an implicit GEMM, a dense contraction:
//
//     flat[b][oc * 4 + pos] = b2[oc] + sum_{ic, tap} pooled[b][ic][cell(pos, tap)] * W2[oc][ic][tap]        M = 4 B, N = 64, K = 128
//
// A persistent CTA (20 warps) owns tiles of 32 observations = 128 GEMM rows (row = 4 * observation + output position):
//   warps 4-19  conv1 + max-pool on the CUDA cores exactly as trunk_fused_fwd_kernel does it (one warp per observation, lane =
//               conv1 channel, the agent's conv1 weights in shared memory); every lane then writes the four im2col chunks of its
//               channel (the 2x2 taps under each output position are 4 consecutive K values = one 16-byte chunk), split into
//               TF32 hi / lo, straight into the K-major SWIZZLE_128B operand tile -- the pooled map never exists in memory
//   warp 0      one thread issues 32 tcgen05.mma.cta_group::1.kind::tf32 per tile (4 K slabs x 4 K steps x {a_hi [b_hi; b_lo] with
//               N = 128, a_lo b_hi with N = 64}: the three TF32 terms, fp32-level accuracy) into a 128-column TMEM accumulator
//               (the epilogue adds the two halves) and commits to an mbarrier.  conv2's weights are pre-split hi / lo panels,
//               loaded ONCE per CTA by TMA (eight 8 KB boxes); the tile's observations arrive by one TMA bulk copy
//   warps 0-3   epilogue: tcgen05.ld of their TMEM lane quadrant, bias, stores in PyTorch's flatten order (the four output
//               positions of an observation sit in four neighbouring lanes: 16 contiguous bytes per output channel)
// The conv warps start the next tile's observations as soon as the MMAs have released the operand tile, while the epilogue
// warps drain the accumulator.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "tc_common.cuh"
#include "tc_tma_host.h"
#include "tc_trunk.h"

using namespace maexp_tc;

namespace {

constexpr int kV = 7, kH1 = 6, kHp = 3, kCells = 9, kF1 = 32, kF2 = 64, kK2 = 128, kTileObs = 32;
constexpr int kEpiWarps = 4, kConvWarps = 16, kThreads = (kEpiWarps + kConvWarps) * 32;
constexpr uint32_t kASlab = kBM * 128;   // 16 KB: 128 rows x 32 floats
constexpr uint32_t kWSlab = kF2 * 128;   //  8 KB:  64 rows x 32 floats
constexpr int kSlabs = kK2 / kBK;        // 4

struct alignas(64) TrunkArgs {
    CUtensorMap tm_w2hi, tm_w2lo;  // [G * 64][128]
    const float *obs, *params;
    float *flat;
    long long stride, g_obs, P, o_c1w, o_c1b, o_c2b, g_flat;
    int B, C;
    float *p_out;      // TRAIN: pooled map [G][.][9][32] (NHWC) and pooling argmax the backward pass reads (group pitch g_p)
    uint8_t *arg_out;
    long long g_p;
    long long *dbg;  // MAEXP_TC_TRUNK_DBG: cycle stamps of CTA (0, 0)
};

// TC > 0: the number of observation channels at compile time (the per-channel cell reads become 128-bit loads at static offsets)
template <int TC, bool TRAIN>
__global__ void __launch_bounds__(kThreads, 1) tc_trunk_kernel(const __grid_constant__ TrunkArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t wbar, mbar, obar;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_b1[kF1], s_b2[kF2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = blockIdx.y;
    const int C = TC ? TC : a.C, K1 = C * 4;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
    uint8_t *a_hi = smem, *a_lo = smem + kSlabs * kASlab;
    // conv2's weight slabs: the hi and the lo panel of a slab sit back to back (128 rows), so a_hi x [W_hi; W_lo] is ONE UMMA with
    // N = 128 (columns 0-63: hi x hi, 64-127: hi x lo) -- an N = 64 UMMA takes the same ~64 cycles as an N = 128 one
    uint8_t *w_hi = smem + 2 * kSlabs * kASlab, *w_lo = w_hi + kWSlab;  // slab s at + s * 2 * kWSlab
    float *w1t = reinterpret_cast<float *>(w_hi + 2 * kSlabs * kWSlab);  // [K1][32] conv1 weight, k-major
    float *s_tile = w1t + K1 * 32;  // the tile's 32 observation rows as they sit in HBM ([32][stride]), one bulk copy per tile
    const float *src = a.obs + (size_t)g * a.g_obs;
    const int ntiles = (a.B + kTileObs - 1) / kTileObs;
    auto request_obs = [&](int tile) {  // one thread: TMA bulk copy of the tile's rows (contiguous: row pitch = ring slot pitch)
        const uint32_t bytes = (uint32_t)(min(kTileObs, a.B - tile * kTileObs) * a.stride * 4);
        mbar_expect_tx(&obar, bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s_tile)),
                     "l"(src + (size_t)tile * kTileObs * a.stride), "r"(bytes), "r"(smem_u32(&obar))
                     : "memory");
    };

    const float *W = a.params + (size_t)g * a.P;
    for (int i = tid; i < K1 * 32; i += kThreads) w1t[i] = W[a.o_c1w + (size_t)(i & 31) * K1 + (i >> 5)];
    if (tid < kF1) s_b1[tid] = W[a.o_c1b + tid];
    if (tid < kF2) s_b2[tid] = W[a.o_c2b + tid];
    if (tid == 0) {
        mbar_init(&wbar, 1);
        mbar_init(&mbar, 1);
        mbar_init(&obar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if ((int)blockIdx.x < ntiles) request_obs(blockIdx.x);
        mbar_expect_tx(&wbar, 2u * kSlabs * kWSlab);
        for (int s = 0; s < kSlabs; ++s) {
            tma_2d(smem_u32(w_hi + s * 2 * kWSlab), &a.tm_w2hi, &wbar, s * kBK, g * kF2);
            tma_2d(smem_u32(w_lo + s * 2 * kWSlab), &a.tm_w2lo, &wbar, s * kBK, g * kF2);
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = idesc_tf32_m128(kF2), idesc2 = idesc_tf32_m128(2 * kF2);
    float *dst = a.flat + (size_t)g * a.g_flat;

    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        if (warp >= kEpiWarps) {
            // ---------------------------------------------------------- conv1 + max-pool + im2col, one observation per pass
            const long long c0 = clock64();
            mbar_wait(&obar, it & 1);  // this tile's observations have landed
            if (a.dbg && blockIdx.x == 0 && g == 0 && tid == kEpiWarps * 32 && it < 16) { a.dbg[it * 8] = c0; a.dbg[it * 8 + 1] = clock64(); }
            for (int o = warp - kEpiWarps; o < kTileObs; o += kConvWarps) {
                const int b = tile * kTileObs + o;
                if (b >= a.B) break;  // rows past the batch keep stale operands: their outputs are never stored
                const float *s_obs = s_tile + (size_t)o * a.stride;
                // channel by channel: the 49 cells of the channel and the lane's 4 taps in registers, then all 36 conv positions
                // (9 pooling windows x 4) as independent accumulators -- 53 shared-memory loads per 144 FMAs, no dependent chains
                float acc[kCells][4];
#pragma unroll
                for (int cell = 0; cell < kCells; ++cell) acc[cell][0] = acc[cell][1] = acc[cell][2] = acc[cell][3] = 0.0f;
#pragma unroll
                for (int c = 0; c < (TC ? TC : C); ++c) {
                    const float *w = w1t + (c * 4) * 32 + lane;
                    const float w0 = w[0], w1 = w[32], w2 = w[64], w3 = w[96];
                    float ob[kV * kV];
                    if constexpr (TC > 0) {  // the channel's 49 floats through 13 aligned 128-bit loads (static offsets)
                        // (the loop over c is unrolled for TC > 0, so c0 / sh are constants)
                        const int first = c * kV * kV, c0 = first & ~3, sh = first - c0;
                        float buf[56];
#pragma unroll
                        for (int v4 = 0; v4 < (kV * kV + 3 + 3) / 4; ++v4) {
                            const float4 t = *reinterpret_cast<const float4 *>(s_obs + c0 + v4 * 4);
                            buf[v4 * 4] = t.x; buf[v4 * 4 + 1] = t.y; buf[v4 * 4 + 2] = t.z; buf[v4 * 4 + 3] = t.w;
                        }
#pragma unroll
                        for (int i = 0; i < kV * kV; ++i) ob[i] = buf[sh + i];
                    } else {
#pragma unroll
                        for (int i = 0; i < kV * kV; ++i) ob[i] = s_obs[c * kV * kV + i];
                    }
#pragma unroll
                    for (int cell = 0; cell < kCells; ++cell) {
                        const int y0 = (cell / kHp) * 2, x0 = (cell % kHp) * 2;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {  // conv position (y0 + q / 2, x0 + q % 2) of the pooling window
                            const int y = y0 + (q >> 1), x = x0 + (q & 1);
                            float t = acc[cell][q];
                            t = fmaf(ob[y * kV + x], w0, t);
                            t = fmaf(ob[y * kV + x + 1], w1, t);
                            t = fmaf(ob[(y + 1) * kV + x], w2, t);
                            t = fmaf(ob[(y + 1) * kV + x + 1], w3, t);
                            acc[cell][q] = t;
                        }
                    }
                }
                float p[kCells];
                if constexpr (TRAIN) {
                    // the learner's online(s) pass keeps what the backward pass reads: the pooled map (NHWC) and the pooling argmax
                    // -- bias added before the comparison, first maximum wins, as F.max_pool2d on the biased conv output
                    const float bb = s_b1[lane];
#pragma unroll
                    for (int cell = 0; cell < kCells; ++cell) {
                        float best = acc[cell][0] + bb;
                        int bi = 0;
                        if (acc[cell][1] + bb > best) { best = acc[cell][1] + bb; bi = 1; }
                        if (acc[cell][2] + bb > best) { best = acc[cell][2] + bb; bi = 2; }
                        if (acc[cell][3] + bb > best) { best = acc[cell][3] + bb; bi = 3; }
                        p[cell] = best;
                        const size_t pi = (size_t)g * a.g_p + ((size_t)b * kCells + cell) * kF1 + lane;
                        a.p_out[pi] = best;
                        a.arg_out[pi] = (uint8_t)bi;
                    }
                } else {
#pragma unroll
                    for (int cell = 0; cell < kCells; ++cell)  // max(x_i) + b == max(x_i + b)
                        p[cell] = fmaxf(fmaxf(acc[cell][0], acc[cell][1]), fmaxf(acc[cell][2], acc[cell][3])) + s_b1[lane];
                }
                // K = ic * 4 + tap (conv2 weight layout [oc][ic][kh][kw]): lane = ic owns chunk ic of every row of this observation
                const int slab = lane >> 3, chunk = lane & 7;
#pragma unroll
                for (int pos = 0; pos < 4; ++pos) {
                    const int c0 = (pos >> 1) * kHp + (pos & 1);
                    stage_chunk<3>(a_hi + slab * kASlab, a_lo + slab * kASlab, o * 4 + pos, chunk,
                                   make_float4(p[c0], p[c0 + 1], p[c0 + kHp], p[c0 + kHp + 1]));
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
            if (a.dbg && blockIdx.x == 0 && g == 0 && tid == kEpiWarps * 32 && it < 16) a.dbg[it * 8 + 2] = clock64();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // operand tile complete; the previous tile's accumulator has been read
        if (tid == 0) {
            if (tile + (int)gridDim.x < ntiles) request_obs(tile + gridDim.x);  // the conv warps are done with the buffer
            if (it == 0) mbar_wait(&wbar, 0);  // conv2 weight panels have landed
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int s = 0; s < kSlabs; ++s) {
                const uint32_t aA = smem_u32(a_hi + s * kASlab), aAlo = smem_u32(a_lo + s * kASlab);
                const uint32_t aW = smem_u32(w_hi + s * 2 * kWSlab);
#pragma unroll
                for (int k = 0; k < kBK / 8; ++k) {
                    const uint32_t ko = (uint32_t)k * 32;
                    umma_tf32(tmem, smem_desc(aA + ko), smem_desc(aW + ko), idesc2, (s > 0 || k > 0) ? 1u : 0u);  // hi x [hi; lo]
                    umma_tf32(tmem, smem_desc(aAlo + ko), smem_desc(aW + ko), idesc, 1u);                       // lo x hi
                }
            }
            umma_commit(&mbar);
        }
        if (a.dbg && blockIdx.x == 0 && g == 0 && tid == kEpiWarps * 32 && it < 16) a.dbg[it * 8 + 3] = clock64();
        mbar_wait(&mbar, it & 1);  // the MMAs have finished: the operand tile is free, the accumulator is complete
        if (a.dbg && blockIdx.x == 0 && g == 0 && tid == kEpiWarps * 32 && it < 16) a.dbg[it * 8 + 4] = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp < kEpiWarps) {
            // ---------------------------------------------------------- epilogue: row = 4 * observation + position
            const int row = warp * 32 + lane, b = tile * kTileObs + (row >> 2), pos = row & 3;
#pragma unroll
            for (int c0 = 0; c0 < kF2; c0 += 16) {
                uint32_t r[16], r2[16];
                tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
                tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kF2 + c0), r2);
                if (b < a.B) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        dst[(size_t)b * (kF2 * 4) + (size_t)(c0 + j) * 4 + pos] = (__uint_as_float(r[j]) + __uint_as_float(r2[j])) + s_b2[c0 + j];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

__global__ void tc_trunk_prepare_kernel(const float *params, long long P, long long o_c2w, float *hi, float *lo) {
    const int g = blockIdx.y;
    const float *src = params + (size_t)g * P + o_c2w;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kF2 * kK2; i += gridDim.x * blockDim.x) {
        const float v = src[i], h = tf32_rna(v);
        hi[(size_t)g * kF2 * kK2 + i] = h;
        lo[(size_t)g * kF2 * kK2 + i] = v - h;
    }
}

}  // namespace

namespace maexp {

size_t tc_trunk_workspace_floats(int G) { return 2 * (size_t)G * kF2 * kK2 + 8; }

bool tc_trunk_forward(cudaStream_t st, int G, const TcTrunkFwd &f) {
    if (f.V != kV || f.F1 != kF1 || f.F2 != kF2 || f.C < 1 || f.C > 16 || f.B < 1 || !f.workspace) return false;
    if (reinterpret_cast<uintptr_t>(f.workspace) & 15) return false;
    TrunkArgs a = {};
    float *hi = f.workspace, *lo = f.workspace + (size_t)G * kF2 * kK2;
    if (!make_map_f32(&a.tm_w2hi, hi, kK2, (uint64_t)G * kF2, 0, kK2 * 4, 0, kF2)) return false;
    if (!make_map_f32(&a.tm_w2lo, lo, kK2, (uint64_t)G * kF2, 0, kK2 * 4, 0, kF2)) return false;
    a.obs = f.obs; a.params = f.params; a.flat = f.flat;
    a.stride = f.stride; a.g_obs = f.g_obs; a.P = f.P; a.o_c1w = f.o_c1w; a.o_c1b = f.o_c1b; a.o_c2b = f.o_c2b; a.g_flat = f.g_flat;
    a.B = f.B; a.C = f.C;
    if ((f.stride & 3) || (f.g_obs & 3) || (reinterpret_cast<uintptr_t>(f.obs) & 15) || f.stride < f.C * kV * kV + 4) return false;
    const size_t smem = 2 * (size_t)kSlabs * kASlab + 2 * (size_t)kSlabs * kWSlab +
                        sizeof(float) * ((size_t)f.C * 4 * 32 + (size_t)kTileObs * f.stride) + 1024;
    if (smem > 227 * 1024) return false;  // (7 channels: 44 KB of observations per tile do not fit next to the operand tiles)
    const bool train = f.p_out != nullptr && f.arg_out != nullptr;
    auto fn = train ? (f.C == 3 ? tc_trunk_kernel<3, true> : tc_trunk_kernel<0, true>) : (f.C == 3 ? tc_trunk_kernel<3, false> : tc_trunk_kernel<0, false>);
    static size_t attr[4] = {0, 0, 0, 0};
    const int ai = (f.C == 3) + 2 * train;
    if (smem > attr[ai]) {
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
        attr[ai] = smem;
    }
    a.p_out = f.p_out; a.arg_out = f.arg_out; a.g_p = f.g_p;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = (f.B + kTileObs - 1) / kTileObs;
    int per_group = sms / G;  // one persistent CTA per SM
    if (per_group < 1) per_group = 1;
    if (per_group > ntiles) per_group = ntiles;
    static long long *dbg_buf = nullptr;
    if (getenv("MAEXP_TC_TRUNK_DBG")) {
        if (!dbg_buf) { cudaMalloc(&dbg_buf, 256 * sizeof(long long)); cudaMemset(dbg_buf, 0, 256 * sizeof(long long)); }
        a.dbg = dbg_buf;
    }
    tc_trunk_prepare_kernel<<<dim3(8, G), 256, 0, st>>>(f.params, f.P, f.o_c2w, hi, lo);
    fn<<<dim3(per_group, G), kThreads, smem, st>>>(a);
    if (a.dbg) {  // measured: wait obs 500 | conv + im2col 7 600 | barrier 300 | wait MMA 3 050 | period 11 700 cycles per 32 observations
        long long h[256];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "tc_trunk CTA(0,0), a conv warp, cycles per tile: wait obs | conv + im2col | to barrier | wait MMA | tile period\n");
        for (int it = 1; it < 12; ++it)
            fprintf(stderr, "  %2d %7lld %7lld %7lld %7lld %8lld\n", it, h[it * 8 + 1] - h[it * 8], h[it * 8 + 2] - h[it * 8 + 1],
                    h[it * 8 + 3] - h[it * 8 + 2], h[it * 8 + 4] - h[it * 8 + 3], h[it * 8] - h[(it - 1) * 8]);
    }
    return cudaGetLastError() == cudaSuccess;
}

}  // namespace maexp
