// tcgen05 (5th-generation tensor core) kernels of the learner's update batches: the two-layer MLP heads of
// core/agents/models/mlp.py:28-34 (fc_V / fc_A of the distributional networks, `mlp` of ConvMLP) for B >= 128 rows.
//
//   tc_mlp_fwd_kernel   out = relu(x W1^T + b1) W2^T + b2 for one (agent, head, 128-row tile) per CTA.  Both contractions run
//                       on tcgen05.mma.cta_group::1.kind::tf32 with fp32 accumulators in tensor memory; every product is
//                       evaluated as three TF32 terms (a_hi b_hi + a_lo b_hi + a_hi b_lo), which keeps the result within
//                       ~2^-21 of the fp32 product sum (the learner's 1e-5 parity bar).  The hidden activation never goes
//                       back to HBM between the layers: the first accumulator is read from TMEM slab by slab
//                       (tcgen05.ld), gets bias + ReLU, is split into hi / lo and becomes the A operand of the second
//                       contraction in shared memory (the training pass also writes it out for the backward pass).
//
// STATUS (round 2, B200): numerically green (c51 / qr B = 256 and the large-batch tests within 1e-5) but SLOWER than the SIMT
// kernel it would replace: 45 us against 33 us for the C51 heads at B = 256 x 8 agents.  A CTA's twelve K slabs are a serial
// chain of global-load latency -> register split -> shared-memory staging -> proxy fence -> CTA barrier -> single-thread MMA
// issue -> commit; the tensor pipe works for ~6 us of the 45 (globaltimer stamps and the ncu source page are in
// profiles/r2_tc_mlp_fwd_stamps.txt / r2_tc_mlp_fwd_ncu.csv: long-scoreboard stalls on the operand loads dominate every
// phase).  32 CTAs of 8 warps cannot hide that latency; the SIMT kernel spreads the same rows over 128 CTAs.  The learner
// therefore keeps the SIMT kernels by default (MAEXP_TC_MLP=1 selects this one); what it needs next is a TMA-fed multi-stage
// ring with pre-split operands so that no thread touches the operands at all.
//
// Operand tiles are K-major SWIZZLE_128B slabs of 32 floats staged through registers (the weights sit at arbitrary float
// offsets of the flat per-agent parameter vector and change every update, so there is nothing for a TMA descriptor to
// point at that would not have to be rewritten per step).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <type_traits>

#include "tc_common.cuh"
#include "tc_mlp.h"

using namespace maexp_tc;

namespace {

constexpr int kThreads = 256;
constexpr uint32_t kTmemCols = 512;  // D1 (<= 128 columns) at column 0, D2 (<= 256 columns) at column 256

__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// instruction descriptor: D = F32 (bits 4-5 = 1), A and B = TF32 (bits 7-9 and 10-12 = 2), both K-major, N >> 3 in bits
// 17-22, M >> 4 in bits 24-28
__device__ __forceinline__ uint32_t idesc_tf32(int npad) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}
// the twelve UMMAs of one 32-float K slab: four K steps of 8, three TF32 terms each
__device__ __forceinline__ void issue_slab(uint32_t tmem_d, uint32_t aA, uint32_t aAlo, uint32_t aW, uint32_t aWlo, uint32_t idesc,
                                           bool first) {
#pragma unroll
    for (int k = 0; k < kBK / 8; ++k) {
        const uint32_t ko = (uint32_t)k * 32;
        umma_tf32(tmem_d, smem_desc(aA + ko), smem_desc(aW + ko), idesc, (!first || k > 0) ? 1u : 0u);
        umma_tf32(tmem_d, smem_desc(aAlo + ko), smem_desc(aW + ko), idesc, 1u);
        umma_tf32(tmem_d, smem_desc(aA + ko), smem_desc(aWlo + ko), idesc, 1u);
    }
}

// ----------------------------------------------------------------------------------------------------- forward
// Pipeline: two shared-memory operand buffers (A hi/lo 32 KB + W hi/lo 64 KB each) and two register sets of prefetched
// global chunks, so that while the tensor core works on slab s the threads stage slab s + 1 and the loads of slab s + 2 are
// in flight.  One mbarrier per buffer receives the tcgen05.commit of the MMAs that read it.
struct Pipe {
    uint64_t *bars;
    uint32_t issued[2];  // commits issued on each buffer
    __device__ __forceinline__ void wait_free(int b) {  // the MMAs that last read buffer b have completed
        if (issued[b]) mbar_wait(&bars[b], (issued[b] - 1) & 1);
    }
};

__global__ void __launch_bounds__(kThreads, 1) tc_mlp_fwd_kernel(const maexp::TcMlpFwd a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int head = blockIdx.y, g = blockIdx.z, m0 = blockIdx.x * kBM;
    const int H = a.H, O = a.out[head], npad1 = (H + 15) / 16 * 16, npad2 = (O + 15) / 16 * 16;
    const float *Wg = a.params + (size_t)g * a.P;
    const float *X = a.x + (size_t)g * a.x_g, *W1 = Wg + a.w1[head], *B1 = Wg + a.b1[head], *W2 = Wg + a.w2[head],
                *B2 = Wg + a.b2[head];
    float *hout = a.hbuf[head] ? a.hbuf[head] + (size_t)g * a.rows_max * H : nullptr;
    float *out = a.outp[head] + (size_t)g * a.rows_max * O;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
    constexpr uint32_t kBufBytes = 2 * kBM * 128 + 2 * 256 * 128;  // A hi | A lo | W hi | W lo
    auto bufA = [&](int b) { return smem + (size_t)b * kBufBytes; };
    auto bufAlo = [&](int b) { return bufA(b) + kBM * 128; };
    auto bufW = [&](int b) { return bufA(b) + 2 * kBM * 128; };
    auto bufWlo = [&](int b) { return bufW(b) + 256 * 128; };

    unsigned long long ts[8] = {};
    auto stamp = [&](int i) { if (a.dbg & 8) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts[i])); };
    stamp(0);
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_base_smem, kTmemCols);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d1 = tmem_base_smem, tmem_d2 = tmem_base_smem + 256;
    Pipe pipe = {bars, {0u, 0u}};
    stamp(1);

    const int nkb1 = (a.din + kBK - 1) / kBK, nkb2 = (npad1 + kBK - 1) / kBK;
    const bool x_vec = (a.x_row & 3) == 0 && (a.x_g & 3) == 0 && ((uintptr_t)a.x & 15) == 0;
    const bool w1_vec = (a.din & 3) == 0 && ((uintptr_t)W1 & 15) == 0;
    const bool w2_vec = (H & 3) == 0 && ((uintptr_t)W2 & 15) == 0;
    // register sets of prefetched chunks: slab s lives in set s & 1
    float4 ra[2][kBM * 8 / kThreads], rw[2][256 * 8 / kThreads];
    // global slab index s: layer 1 slabs 0 .. nkb1-1 (x and W1), then layer 2 slabs (W2 only; A comes out of tensor memory)
    auto prefetch = [&](int s, int set) {
        if (s < nkb1) {
            const int k0 = s * kBK;
#pragma unroll
            for (int j = 0; j < kBM * 8 / kThreads; ++j) {
                const int i = tid + j * kThreads, r = i >> 3, c = i & 7, m = m0 + r;
                ra[set][j] = m < a.rows ? load_chunk(X + (size_t)m * a.x_row + k0 + c * 4, k0 + c * 4, a.din, x_vec)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 128 * 8 / kThreads; ++j) {
                const int i = tid + j * kThreads, r = i >> 3, c = i & 7;
                if (i < npad1 * 8)
                    rw[set][j] = r < H ? load_chunk(W1 + (size_t)r * a.din + k0 + c * 4, k0 + c * 4, a.din, w1_vec)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else if (s < nkb1 + nkb2) {
            const int k0 = (s - nkb1) * kBK;
#pragma unroll
            for (int j = 0; j < 256 * 8 / kThreads; ++j) {
                const int i = tid + j * kThreads, r = i >> 3, c = i & 7;
                if (i < npad2 * 8)
                    rw[set][j] = r < O ? load_chunk(W2 + (size_t)r * H + k0 + c * 4, k0 + c * 4, H, w2_vec)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    const uint32_t idesc1 = idesc_tf32(npad1), idesc2 = idesc_tf32(npad2);
    const int quad = warp & 3, half = warp >> 2, row = quad * 32 + lane, m = m0 + row;

    // one pipeline step; SET is the compile-time register set (= s & 1) so that the arrays stay in registers
    auto step = [&](int s, auto set_tag) {
        constexpr int SET = decltype(set_tag)::value;
        const int b = s & 1;
        pipe.wait_free(b);
        uint8_t *sA = bufA(b), *sAlo = bufAlo(b), *sW = bufW(b), *sWlo = bufWlo(b);
        if (s < nkb1) {
#pragma unroll
            for (int j = 0; j < kBM * 8 / kThreads; ++j) {
                const int i = tid + j * kThreads;
                stage_chunk<3>(sA, sAlo, i >> 3, i & 7, ra[SET][j]);
            }
#pragma unroll
            for (int j = 0; j < 128 * 8 / kThreads; ++j) {
                const int i = tid + j * kThreads;
                if (i < npad1 * 8) stage_chunk<3>(sW, sWlo, i >> 3, i & 7, rw[SET][j]);
            }
        } else {
            // layer 2: the A slab is relu(D1 + b1), read from tensor memory (layer 1 is complete: its last commit was waited
            // for by wait_free of this or the previous step — see below)
            const int kb = s - nkb1, n0 = kb * kBK + half * 16;  // this thread's 16 hidden units of the slab
            float v[16];
            if (n0 < npad1) {
                uint32_t r[16];
                tmem_ld16(tmem_d1 + ((uint32_t)(quad * 32) << 16) + (uint32_t)n0, r);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = (n0 + j < H) ? fmaxf(__uint_as_float(r[j]) + B1[n0 + j], 0.0f) : 0.0f;
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.0f;
            }
            if (hout && m < a.rows) {  // the training pass keeps the hidden activation for the backward pass
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const int n = n0 + j4 * 4;
                    if ((H & 3) == 0 && n + 3 < H) {
                        *reinterpret_cast<float4 *>(hout + (size_t)m * H + n) =
                            make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (n + j < H) hout[(size_t)m * H + n + j] = v[j4 * 4 + j];
                    }
                }
            }
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
                stage_chunk<3>(sA, sAlo, row, half * 4 + j4, make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]));
#pragma unroll
            for (int j = 0; j < 256 * 8 / kThreads; ++j) {
                const int i = tid + j * kThreads;
                if (i < npad2 * 8) stage_chunk<3>(sW, sWlo, i >> 3, i & 7, rw[SET][j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (s < nkb1) issue_slab(tmem_d1, smem_u32(sA), smem_u32(sAlo), smem_u32(sW), smem_u32(sWlo), idesc1, s == 0);
            else issue_slab(tmem_d2, smem_u32(sA), smem_u32(sAlo), smem_u32(sW), smem_u32(sWlo), idesc2, s == nkb1);
            umma_commit(&bars[b]);
        }
        ++pipe.issued[b];
        prefetch(s + 2, SET);  // this set's registers are free again
        if (s + 1 == nkb1) {
            // the first layer-2 step reads D1: every layer-1 MMA must have completed (commits complete in issue order, so
            // the last one covers both buffers)
            mbar_wait(&bars[b], (pipe.issued[b] - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
    };
    prefetch(0, 0);
    prefetch(1, 1);
    const int nslabs = nkb1 + nkb2;
    for (int s = 0; s < nslabs; s += 2) {
        step(s, std::integral_constant<int, 0>{});
        if (s + 1 < nslabs) step(s + 1, std::integral_constant<int, 1>{});
    }
    stamp(2);
    {  // all MMAs complete
        const int b = (nslabs - 1) & 1;
        mbar_wait(&bars[b], (pipe.issued[b] - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    stamp(3);

    // ---- epilogue: out = D2 + b2 through shared memory (both operand buffers are free now) so that the global stores are
    //      coalesced rows instead of 32 scattered 16-byte pieces per instruction
    float *tile = reinterpret_cast<float *>(smem);
    const int pitch = npad2 + 1;  // odd pitch: lanes (= rows) hit distinct banks
    for (int n0 = half * 16; n0 < npad2; n0 += 32) {
        uint32_t r[16];
        tmem_ld16(tmem_d2 + ((uint32_t)(quad * 32) << 16) + (uint32_t)n0, r);
#pragma unroll
        for (int j = 0; j < 16; ++j) tile[row * pitch + n0 + j] = (n0 + j < O) ? __uint_as_float(r[j]) + B2[n0 + j] : 0.0f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    stamp(4);
    const int rows_here = min(kBM, a.rows - m0);
    float *obase = out + (size_t)m0 * O;  // the tile's rows are one contiguous run of rows_here * O floats
    const int total = rows_here * O;
    if ((((uintptr_t)obase) & 15) == 0) {
        for (int i = tid * 4; i < total; i += kThreads * 4) {
            float w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = i + j, r = e / O, c = e - r * O;
                w[j] = e < total ? tile[r * pitch + c] : 0.0f;
            }
            if (i + 3 < total) *reinterpret_cast<float4 *>(obase + i) = make_float4(w[0], w[1], w[2], w[3]);
            else
                for (int j = 0; j < 4 && i + j < total; ++j) obase[i + j] = w[j];
        }
    } else {
        for (int i = tid; i < total; i += kThreads) {
            const int r = i / O, c = i - r * O;
            obase[i] = tile[r * pitch + c];
        }
    }
    stamp(5);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_smem), "r"(kTmemCols) : "memory");
    stamp(6);
    if ((a.dbg & 8) && tid == 0 && (blockIdx.x | blockIdx.z) == 0)
        printf("tc_mlp head %d (O %d): alloc %llu ns, slabs %llu, drain %llu, tmem->smem %llu, store %llu, dealloc %llu\n", head, O,
               ts[1] - ts[0], ts[2] - ts[1], ts[3] - ts[2], ts[4] - ts[3], ts[5] - ts[4], ts[6] - ts[5]);
}

}  // namespace

namespace maexp {

bool tc_mlp_forward(cudaStream_t st, int G, const TcMlpFwd &a) {
    if (a.nheads < 1 || a.nheads > 2 || a.rows < 1 || a.H < 4 || a.H > 128 || a.din < 1) return false;
    for (int k = 0; k < a.nheads; ++k)
        if (a.out[k] < 1 || a.out[k] > 256) return false;
    const size_t smem = 2 * (2 * (size_t)kBM * 128 + 2 * 256 * 128) + 1024;  // two operand buffers
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(tc_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
        configured = true;
    }
    tc_mlp_fwd_kernel<<<dim3((a.rows + kBM - 1) / kBM, a.nheads, G), kThreads, smem, st>>>(a);
    return true;
}

}  // namespace maexp
