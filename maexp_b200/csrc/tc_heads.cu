// TMA-fed, warp-specialised tcgen05 kernels for the two-layer MLP heads of core/agents/models/mlp.py:28-34 -- fc_V and fc_A of
// the distributional networks, `mlp` of ConvMLP -- as ONE kernel per forward pass.  Three entry points of the same pipeline:
//   tc_heads_kernel          one CTA per 128-row tile; raw head outputs (+ the hidden activation for a backward pass) or the
//                            fused greedy-value epilogue; up to three passes / heads as grid slices (the update's forward passes)
//   tc_heads_acting_kernel   persistent form for the acting batches of the 4-action dueling distributional networks: the
//                            epilogue of tile i overlaps the main loop of tile i + 1 (further down)
//   quantile-mean collapse   QR-DQN acting: the second layer shrinks to 1 + A rows (plan_pass)
//
//
//     hidden_k = relu(x W1_k^T + b1_k)   (all heads share x: one UMMA with N = heads * 128 columns)
//     out_k    = hidden_k W2_k^T + b2_k  (the hidden activation goes from tensor memory through bias + ReLU back into shared
//                                         memory as the A operand of the second contraction; it never touches HBM)
//
// Every fp32 product is evaluated as three TF32 terms (a_hi b_hi + a_lo b_hi + a_hi b_lo, a_hi = rna-tf32(a)), which keeps
// the Q-values inside the learner's 1e-5 parity bar (tests/test_learner_gpu.py large-batch tests).
//
// Roles of tc_heads_kernel (320 threads, one CTA per 128-row tile and agent, one CTA per SM: it owns all 512 TMEM columns):
//   warp 0      producer: one thread issues the TMA loads (cp.async.bulk.tensor, SWIZZLE_128B boxes of 32 floats x rows) of the
//               raw fp32 x slab and of the PRE-SPLIT weight slabs (hi / lo panels written once per call by
//               tc_heads_prepare_kernel, so no thread ever touches a weight) into a two-stage ring, mbarrier expect_tx
//   warp 1      one thread issues tcgen05.mma.cta_group::1.kind::tf32 (12 per 32-float K slab: 4 K steps x 3 terms) and
//               tcgen05.commit, whose mbarrier arrival hands the stage back to the producer
//   warps 2-5   layer 1: split the x slab into hi / lo in place (shared -> registers -> shared, no global latency on this
//               path); layer 2: read 32 hidden columns of their TMEM lane quadrant (tcgen05.ld), bias + ReLU, split, store
//               the swizzled A slab
//   warps 2-9   epilogue, two threads per row: accumulators -> shared memory (odd pitch), then either coalesced row stores of the
//               raw outputs or the dueling combination + expectation (greedy_values_row)
// Tried and dropped in the persistent kernel: eight operand warps (two threads per row in layer 2) -- 0.63 against 0.61 ms.
// Stage hand-offs: full[s] (TMA bytes landed) -> conv[s] (operands split, fence.proxy.async) -> MMA -> empty[s] (commit).
// TMEM: columns [0, heads*128) hold the first accumulator; the widest head's output sits at column 256, the other head's
// output re-uses the first-processed head's hidden columns once they have been consumed.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "tc_common.cuh"
#include "tc_heads.h"
#include "tc_tma_host.h"

using namespace maexp_tc;

namespace {

constexpr int kThreads = 320, kStages = 2, kHidden = 128;  // producer, MMA issuer, 4 operand warps, 4 more epilogue warps
constexpr uint32_t kATile = kBM * 128;          // 16 KB: 128 rows x 32 floats
constexpr uint32_t kWTile = 256 * 128;          // 32 KB: up to 256 rows x 32 floats
constexpr uint32_t kStageBytes = 2 * kATile + 2 * kWTile;

struct alignas(64) HeadsArgs {
    CUtensorMap tm_x, tm_w1hi, tm_w1lo, tm_w2hi[2], tm_w2lo[2];
    const float *b1cat;   // [G][heads * 128]
    const float *b2p[2];  // [G][npad_k], zero padded
    float *out[2];
    long long out_g[2];
    float *hbuf[2];       // training pass: relu(x W1^T + b1) of every head, [G][.][128] (group pitch hbuf_g), or nullptr
    long long hbuf_g;
    int out_n[2], npad[2];
    int order[2];         // heads in the order their second layer runs
    uint32_t d2col[2];    // TMEM column of every head's output accumulator
    int rows, din, nheads;
    // fused greedy-value epilogue of the dueling distributional heads (head 0 = fc_V [Z], head 1 = fc_A [A][Z]):
    // epi 1 = categorical expectation (softmax over the atoms, sum p z), epi 2 = mean over the quantiles, 0 = raw head outputs
    int epi, A, Z, rs;    // rs: odd row pitch (floats) of the shared-memory copy of the two accumulators
    const float *support; // [Z] atom values (epi 1)
    float *q_out;         // [G][rows][A]
    long long q_g;
    long long *dbg;       // MAEXP_TC_HEADS_DBG: cycle counts of CTA (0, 0) per role and phase
};

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }  // the eight epilogue warps
__device__ __forceinline__ float fast_exp(float x) {  // x <= 0: one MUFU; the terms that matter (x near 0) are good to ~1e-7
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
// Greedy values of one row from its two accumulators in shared memory (my[0..npadV) = fc_V, adv = fc_A [A][Z], both without
// bias): dueling combination v + (adv - mean over actions), then the categorical expectation (softmax over the atoms,
// sum p z) or the mean over the quantiles (get_action of categorical_dqn_brain.py / qr_dqn_brain.py).  TWO threads per row
// (half = 0 / 1, in different warps): the dueling pass is split over the atoms, the expectation over the actions (half, half
// + 2, ...), epi_bar() in between.  TA > 0: the number of actions at compile time -- the per-action chains are the only
// instruction-level parallelism a row has.  Four interleaved partial sums per action: the expectation sum p z cancels (atoms
// of both signs), and a single running sum over the 51 atoms loses ~2.5x more than the tree reduction of dist_q_kernel
// (7.6e-6 against 3.0e-6 mean relative error of the C51 Q-values against the fp32 oracle; tools/exp_tc_accuracy.py).
template <int TA>
__device__ __forceinline__ void greedy_values_row(float *my, float *adv, const float *bV, const float *bA, const float *sup, int A_rt,
                                                  int Z, int categorical, float *q_row, bool store, int half) {
    const int A = TA ? TA : A_rt;
    constexpr int NA = TA ? TA : 8, NH = (NA + 1) / 2;
    const int zh = (Z + 1) / 2, z_lo = half * zh, z_hi = min(Z, z_lo + zh);
    for (int z = z_lo; z < z_hi; ++z) {
        float x[NA], sum = 0.0f;
#pragma unroll
        for (int k = 0; k < NA; ++k)
            if (k < A) { x[k] = adv[k * Z + z] + bA[k * Z + z]; sum += x[k]; }
        const float c = (my[z] + bV[z]) - sum / (float)A;
#pragma unroll
        for (int k = 0; k < NA; ++k)
            if (k < A) adv[k * Z + z] = x[k] + c;
    }
    epi_bar();
    float mx[NH], se[NH][4], sq[NH][4];
#pragma unroll
    for (int i = 0; i < NH; ++i) {
        mx[i] = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) se[i][j] = sq[i][j] = 0.0f;
    }
    if (categorical) {
        for (int z = 0; z < Z; ++z) {
#pragma unroll
            for (int i = 0; i < NH; ++i)
                if (half + 2 * i < A) mx[i] = fmaxf(mx[i], adv[(half + 2 * i) * Z + z]);
        }
    }
    for (int z0 = 0; z0 < Z; z0 += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int z = z0 + j;
            if (z < Z) {
                const float sz = sup[z];
#pragma unroll
                for (int i = 0; i < NH; ++i)
                    if (half + 2 * i < A) {
                        const float d = adv[(half + 2 * i) * Z + z];
                        if (categorical) {
                            const float e = fast_exp(d - mx[i]);
                            se[i][j] += e;
                            sq[i][j] += e * sz;
                        } else {
                            sq[i][j] += d;
                        }
                    }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NH; ++i)
        if (half + 2 * i < A && store) {
            const float num = (sq[i][0] + sq[i][1]) + (sq[i][2] + sq[i][3]), den = (se[i][0] + se[i][1]) + (se[i][2] + se[i][3]);
            q_row[half + 2 * i] = categorical ? num / den : num / (float)Z;
        }
}

// up to three passes (their own input, parameters and outputs; same shapes) as grid slices of one launch: the forward passes of
// an update.  The whole block sits in the kernel's parameter space, so the tensor maps of every pass can be addressed there.
struct alignas(64) HeadsArgsN {
    HeadsArgs p[3];
};

__global__ void __launch_bounds__(kThreads, 1) tc_heads_kernel(const __grid_constant__ HeadsArgsN aa) {
    const HeadsArgs &a = aa.p[blockIdx.z];
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t full[kStages], conv[kStages], empty[kStages], d1_done, all_done;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_b1[2 * kHidden];
    __shared__ float s_b2[512], s_sup[256];  // fused epilogue: both heads' output biases (padded widths back to back), atom values
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.y, m0 = blockIdx.x * kBM;
    for (int i = tid; i < a.nheads * kHidden; i += kThreads) s_b1[i] = a.b1cat[(size_t)g * a.nheads * kHidden + i];
    for (int i = tid; i < a.npad[0] + (a.nheads > 1 ? a.npad[1] : 0); i += kThreads)
        s_b2[i] = i < a.npad[0] ? a.b2p[0][(size_t)g * a.npad[0] + i] : a.b2p[1][(size_t)g * a.npad[1] + i - a.npad[0]];
    if (a.epi)
        for (int i = tid; i < a.Z; i += kThreads) s_sup[i] = a.support ? a.support[i] : 0.0f;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
    auto a_hi = [&](int s) { return smem + (size_t)s * kStageBytes; };
    auto a_lo = [&](int s) { return smem + (size_t)s * kStageBytes + kATile; };
    auto w_hi = [&](int s) { return smem + (size_t)s * kStageBytes + 2 * kATile; };
    auto w_lo = [&](int s) { return smem + (size_t)s * kStageBytes + 2 * kATile + kWTile; };
    const int ns1 = a.din / kBK, per_head = kHidden / kBK, T = ns1 + per_head * a.nheads;
    const int n1 = a.nheads * kHidden;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&conv[s], 128);
            mbar_init(&empty[s], 1);
        }
        mbar_init(&d1_done, 1);
        mbar_init(&all_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            for (int t = 0; t < T; ++t) {
                const int s = t % kStages, n = t / kStages;
                const long long c0 = clock64();
                if (t >= kStages) mbar_wait(&empty[s], (n - 1) & 1);
                if (a.dbg && blockIdx.x == 0 && g == 0) a.dbg[t] = clock64() - c0;  // producer: wait for the stage
                if (t < ns1) {
                    mbar_expect_tx(&full[s], kATile + 2u * (uint32_t)n1 * 128u);
                    tma_3d(smem_u32(a_hi(s)), &a.tm_x, &full[s], t * kBK, m0, g);
                    tma_2d(smem_u32(w_hi(s)), &a.tm_w1hi, &full[s], t * kBK, g * n1);
                    tma_2d(smem_u32(w_lo(s)), &a.tm_w1lo, &full[s], t * kBK, g * n1);
                } else {
                    const int h = a.order[(t - ns1) / per_head], j = (t - ns1) % per_head;
                    mbar_expect_tx(&full[s], 2u * (uint32_t)a.npad[h] * 128u);
                    tma_2d(smem_u32(w_hi(s)), &a.tm_w2hi[h], &full[s], j * kBK, g * a.npad[h]);
                    tma_2d(smem_u32(w_lo(s)), &a.tm_w2lo[h], &full[s], j * kBK, g * a.npad[h]);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            for (int t = 0; t < T; ++t) {
                const int s = t % kStages, n = t / kStages;
                const long long c0 = clock64();
                mbar_wait(&full[s], n & 1);  // the weight slabs (and, in layer 1, the raw x slab) have landed
                const long long c1 = clock64();
                mbar_wait(&conv[s], n & 1);  // the A operand is split and visible to the async proxy
                if (a.dbg && blockIdx.x == 0 && g == 0) { a.dbg[32 + t] = c1 - c0; a.dbg[64 + t] = clock64() - c1; a.dbg[96 + t] = c0; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t d = tmem, idesc = idesc_tf32_m128(n1);
                bool first = t == 0;
                if (t >= ns1) {
                    const int h = a.order[(t - ns1) / per_head], j = (t - ns1) % per_head;
                    d = tmem + a.d2col[h];
                    idesc = idesc_tf32_m128(a.npad[h]);
                    first = j == 0;
                }
                const uint32_t aA = smem_u32(a_hi(s)), aAlo = smem_u32(a_lo(s)), aW = smem_u32(w_hi(s)), aWlo = smem_u32(w_lo(s));
#pragma unroll
                for (int k = 0; k < kBK / 8; ++k) {  // UMMA_K = 8 TF32 = 32 bytes inside the 128-byte swizzle row
                    const uint32_t ko = (uint32_t)k * 32;
                    umma_tf32(d, smem_desc(aA + ko), smem_desc(aW + ko), idesc, (first && k == 0) ? 0u : 1u);
                    umma_tf32(d, smem_desc(aAlo + ko), smem_desc(aW + ko), idesc, 1u);
                    umma_tf32(d, smem_desc(aA + ko), smem_desc(aWlo + ko), idesc, 1u);
                }
                umma_commit(&empty[s]);
                if (t == ns1 - 1) umma_commit(&d1_done);
                if (t == T - 1) umma_commit(&all_done);
            }
        }
    } else {
        // ------------------------------------------------------------------ operand conversion (warps 2-5) + epilogue (warps 2-9)
        const int ct = (tid - 64) & 127, quad = warp & 3, row = quad * 32 + lane;  // a warp reads the TMEM lane quadrant warp % 4
        const int half = warp >= 6;  // second thread of the row in the epilogue
        for (int t = 0; t < (half ? 0 : T); ++t) {
            const int s = t % kStages, n = t / kStages;
            if (t < ns1) {
                mbar_wait(&full[s], n & 1);
                float4 *hi4 = reinterpret_cast<float4 *>(a_hi(s)), *lo4 = reinterpret_cast<float4 *>(a_lo(s));
#pragma unroll
                for (int i = 0; i < (int)(kATile / 16 / 128); ++i) {  // element-wise: the swizzled position is kept
                    const float4 v = hi4[ct + 128 * i];
                    const float4 hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
                    hi4[ct + 128 * i] = hi;
                    lo4[ct + 128 * i] = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                }
            } else {
                if (t == ns1) {
                    mbar_wait(&d1_done, 0);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                if (t >= kStages) mbar_wait(&empty[s], (n - 1) & 1);  // the MMAs that read this stage's A slab have finished
                const int h = a.order[(t - ns1) / per_head], j = (t - ns1) % per_head;
                uint32_t r[32];
                tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(h * kHidden + j * kBK), r);
                const float *b1 = s_b1 + h * kHidden + j * kBK;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 b = *reinterpret_cast<const float4 *>(b1 + c * 4);
                    const float4 v = make_float4(fmaxf(__uint_as_float(r[c * 4]) + b.x, 0.0f), fmaxf(__uint_as_float(r[c * 4 + 1]) + b.y, 0.0f),
                                                 fmaxf(__uint_as_float(r[c * 4 + 2]) + b.z, 0.0f), fmaxf(__uint_as_float(r[c * 4 + 3]) + b.w, 0.0f));
                    stage_chunk<3>(a_hi(s), a_lo(s), row, c, v);
                    if (a.hbuf[h] && m0 + row < a.rows)  // the backward pass and its ReLU mask read the hidden activation
                        *reinterpret_cast<float4 *>(a.hbuf[h] + (size_t)g * a.hbuf_g + (size_t)(m0 + row) * kHidden + j * kBK + c * 4) = v;
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
            mbar_arrive(&conv[s]);
        }
        // ---- epilogue: bias, row stores (thread = row)
        const long long e0 = clock64();
        mbar_wait(&all_done, 0);
        if (a.dbg && blockIdx.x == 0 && g == 0 && tid == 64) { a.dbg[128] = clock64() - e0; a.dbg[129] = clock64(); }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int m = m0 + row;
        // every row's accumulators go to shared memory first (odd pitch: conflict-free for row-per-thread access); the two
        // threads of a row take alternate 16-column groups
        float *s_acc = reinterpret_cast<float *>(smem);
        float *my = s_acc + (size_t)row * a.rs;
        for (int h = 0; h < a.nheads; ++h) {
            const int off = h ? a.npad[0] : 0;
            for (int c0 = half * 16; c0 < a.npad[h]; c0 += 32) {
                uint32_t r[16];
                tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + a.d2col[h] + (uint32_t)c0, r);
#pragma unroll
                for (int j = 0; j < 16; ++j) my[off + c0 + j] = __uint_as_float(r[j]);
            }
        }
        epi_bar();
        if (a.epi == 3) {
            // collapsed heads: one value and A advantages per row
            if (!half && m < a.rows) {
                const float *adv = my + a.npad[0], *bA = s_b2 + a.npad[0];
                float sum = 0.0f;
                for (int k = 0; k < a.A; ++k) sum += adv[k] + bA[k];
                const float c = (my[0] + s_b2[0]) - sum / (float)a.A;
                for (int k = 0; k < a.A; ++k) a.q_out[(size_t)g * a.q_g + (size_t)m * a.A + k] = (adv[k] + bA[k]) + c;
            }
        } else if (a.epi) {
            // dueling combination + expectation in shared memory: the 255 head outputs per row never reach HBM
            float *q_row = a.q_out + (size_t)g * a.q_g + (size_t)m * a.A;
            if (a.A == 4) greedy_values_row<4>(my, my + a.npad[0], s_b2, s_b2 + a.npad[0], s_sup, 4, a.Z, a.epi == 1, q_row, m < a.rows, half);
            else if (a.A == 5) greedy_values_row<5>(my, my + a.npad[0], s_b2, s_b2 + a.npad[0], s_sup, 5, a.Z, a.epi == 1, q_row, m < a.rows, half);
            else greedy_values_row<0>(my, my + a.npad[0], s_b2, s_b2 + a.npad[0], s_sup, a.A, a.Z, a.epi == 1, q_row, m < a.rows, half);
        } else {
            // raw head outputs: a warp per row, lanes along the columns -> coalesced row stores
            const int ew = warp - 2;
            for (int h = 0; h < a.nheads; ++h) {
                const int off = h ? a.npad[0] : 0, on = a.out_n[h];
                float *obase = a.out[h] + (size_t)g * a.out_g[h];
                for (int r = ew; r < kBM && m0 + r < a.rows; r += 8)
                    for (int c = lane; c < on; c += 32) obase[(size_t)(m0 + r) * on + c] = s_acc[(size_t)r * a.rs + off + c] + s_b2[off + c];
            }
        }
    }
    if (a.dbg && blockIdx.x == 0 && g == 0 && tid == 64) a.dbg[130] = clock64();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent form for the acting batches of the dueling distributional networks with 4 actions (fused greedy values):
// a CTA walks the 128-row tiles of one agent, and the epilogue of tile i runs on its own eight warps WHILE the producer / MMA /
// operand warps are already in tile i + 1.  What makes that possible:
//   * the advantage head's output columns are atom-major (column z * 4 + k; tc_heads_prepare_kernel permutes W2 / b2), so a
//     16-column tcgen05.ld holds all four actions of four atoms: the dueling mean needs no shared-memory copy of the accumulator
//     and the epilogue reads tensor memory directly (the operand ring stays with the main loop)
//   * the value head's output (64 columns inside the first accumulator's region) is moved to shared memory first thing
//     (d1_free), after which the next tile's first layer may overwrite it; the advantage head's accumulator (columns 256..)
//     is only needed by the next tile's SECOND layer (d2_free), eight K slabs later
// TMEM, roles and the three-term TF32 arithmetic are those of tc_heads_kernel.
constexpr int kThreadsP = 448;  // producer, MMA issuer, 4 operand warps, 8 epilogue warps
constexpr int kCPitch = 53;     // floats per row of the c_z = v_z - mean_z exchange buffer (Z <= 52)

__global__ void __launch_bounds__(kThreadsP, 1) tc_heads_acting_kernel(const __grid_constant__ HeadsArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t full[kStages], conv[kStages], empty[kStages], d1_done, d2_full, d1_free, d2_free;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_b1[2 * kHidden];
    __shared__ float s_b2[288], s_sup[64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = blockIdx.y;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    auto a_hi = [&](int s) { return smem + (size_t)s * kStageBytes; };
    auto a_lo = [&](int s) { return smem + (size_t)s * kStageBytes + kATile; };
    auto w_hi = [&](int s) { return smem + (size_t)s * kStageBytes + 2 * kATile; };
    auto w_lo = [&](int s) { return smem + (size_t)s * kStageBytes + 2 * kATile + kWTile; };
    float *s_c = reinterpret_cast<float *>(smem + (size_t)kStages * kStageBytes);  // [128][kCPitch]
    const int ns1 = a.din / kBK, per_head = kHidden / kBK, T = ns1 + 2 * per_head, n1 = 2 * kHidden;
    const int ntiles = (a.rows + kBM - 1) / kBM;

    for (int i = tid; i < n1; i += kThreadsP) s_b1[i] = a.b1cat[(size_t)g * n1 + i];
    for (int i = tid; i < a.npad[0] + a.npad[1]; i += kThreadsP)
        s_b2[i] = i < a.npad[0] ? a.b2p[0][(size_t)g * a.npad[0] + i] : a.b2p[1][(size_t)g * a.npad[1] + i - a.npad[0]];
    for (int i = tid; i < a.Z; i += kThreadsP) s_sup[i] = a.support ? a.support[i] : 0.0f;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&conv[s], 128);
            mbar_init(&empty[s], 1);
        }
        mbar_init(&d1_done, 1);
        mbar_init(&d2_full, 1);
        mbar_init(&d1_free, 256);
        mbar_init(&d2_free, 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            int u = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int t = 0; t < T; ++t, ++u) {
                    const int s = u % kStages, n = u / kStages;
                    if (u >= kStages) mbar_wait(&empty[s], (n - 1) & 1);
                    if (t < ns1) {
                        mbar_expect_tx(&full[s], kATile + 2u * (uint32_t)n1 * 128u);
                        tma_3d(smem_u32(a_hi(s)), &a.tm_x, &full[s], t * kBK, tile * kBM, g);
                        tma_2d(smem_u32(w_hi(s)), &a.tm_w1hi, &full[s], t * kBK, g * n1);
                        tma_2d(smem_u32(w_lo(s)), &a.tm_w1lo, &full[s], t * kBK, g * n1);
                    } else {
                        const int h = a.order[(t - ns1) / per_head], j = (t - ns1) % per_head;
                        mbar_expect_tx(&full[s], 2u * (uint32_t)a.npad[h] * 128u);
                        tma_2d(smem_u32(w_hi(s)), &a.tm_w2hi[h], &full[s], j * kBK, g * a.npad[h]);
                        tma_2d(smem_u32(w_lo(s)), &a.tm_w2lo[h], &full[s], j * kBK, g * a.npad[h]);
                    }
                }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int u = 0, it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it)
                for (int t = 0; t < T; ++t, ++u) {
                    const int s = u % kStages, n = u / kStages;
                    mbar_wait(&full[s], n & 1);
                    mbar_wait(&conv[s], n & 1);
                    if (it > 0 && t == 0) mbar_wait(&d1_free, (it - 1) & 1);    // the previous tile's value head has been read
                    if (it > 0 && t == ns1) mbar_wait(&d2_free, (it - 1) & 1);  // ... and its advantage head
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint32_t d = tmem, idesc = idesc_tf32_m128(n1);
                    bool first = t == 0;
                    if (t >= ns1) {
                        const int h = a.order[(t - ns1) / per_head], j = (t - ns1) % per_head;
                        d = tmem + a.d2col[h];
                        idesc = idesc_tf32_m128(a.npad[h]);
                        first = j == 0;
                    }
                    const uint32_t aA = smem_u32(a_hi(s)), aAlo = smem_u32(a_lo(s)), aW = smem_u32(w_hi(s)), aWlo = smem_u32(w_lo(s));
#pragma unroll
                    for (int k = 0; k < kBK / 8; ++k) {
                        const uint32_t ko = (uint32_t)k * 32;
                        umma_tf32(d, smem_desc(aA + ko), smem_desc(aW + ko), idesc, (first && k == 0) ? 0u : 1u);
                        umma_tf32(d, smem_desc(aAlo + ko), smem_desc(aW + ko), idesc, 1u);
                        umma_tf32(d, smem_desc(aA + ko), smem_desc(aWlo + ko), idesc, 1u);
                    }
                    umma_commit(&empty[s]);
                    if (t == ns1 - 1) umma_commit(&d1_done);
                    if (t == T - 1) umma_commit(&d2_full);
                }
        }
    } else if (warp < 6) {
        // ------------------------------------------------------------------ operand warps
        const int ct = tid - 64, quad = warp & 3, row = quad * 32 + lane;
        int u = 0, it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it)
            for (int t = 0; t < T; ++t, ++u) {
                const int s = u % kStages, n = u / kStages;
                if (t < ns1) {
                    mbar_wait(&full[s], n & 1);
                    float4 *hi4 = reinterpret_cast<float4 *>(a_hi(s)), *lo4 = reinterpret_cast<float4 *>(a_lo(s));
#pragma unroll
                    for (int i = 0; i < (int)(kATile / 16 / 128); ++i) {
                        const float4 v = hi4[ct + 128 * i];
                        const float4 hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
                        hi4[ct + 128 * i] = hi;
                        lo4[ct + 128 * i] = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                    }
                } else {
                    if (t == ns1) {
                        mbar_wait(&d1_done, it & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    if (u >= kStages) mbar_wait(&empty[s], (n - 1) & 1);
                    const int h = a.order[(t - ns1) / per_head], j = (t - ns1) % per_head;
                    uint32_t r[32];
                    tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(h * kHidden + j * kBK), r);
                    const float *b1 = s_b1 + h * kHidden + j * kBK;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 b = *reinterpret_cast<const float4 *>(b1 + c * 4);
                        const float4 v = make_float4(fmaxf(__uint_as_float(r[c * 4]) + b.x, 0.0f), fmaxf(__uint_as_float(r[c * 4 + 1]) + b.y, 0.0f),
                                                     fmaxf(__uint_as_float(r[c * 4 + 2]) + b.z, 0.0f), fmaxf(__uint_as_float(r[c * 4 + 3]) + b.w, 0.0f));
                        stage_chunk<3>(a_hi(s), a_lo(s), row, c, v);
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&conv[s]);
            }
    } else {
        // ------------------------------------------------------------------ epilogue warps (two threads per row)
        const int quad = warp & 3, row = quad * 32 + lane, half = warp >= 10;
        const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);
        const int Z = a.Z, npadA = a.npad[1];
        const float *bV = s_b2, *bA = s_b2 + a.npad[0];
        float *c_row = s_c + (size_t)row * kCPitch;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            mbar_wait(&d2_full, it & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // value head -> shared memory (two 16-column groups per thread), then the first accumulator's region is free
            for (int c0 = half * 16; c0 < a.npad[0]; c0 += 32) {
                uint32_t r[16];
                tmem_ld16(lane_base + a.d2col[0] + (uint32_t)c0, r);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < Z) c_row[c0 + j] = __uint_as_float(r[j]) + bV[c0 + j];
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&d1_free);
            epi_bar();
            // dueling pass, split over the atoms: c_z = v_z - mean over the actions (columns z * 4 + k)
            {
                const int nchunks = npadA / 16, c_lo = half ? nchunks / 2 : 0, c_hi = half ? nchunks : nchunks / 2;
                for (int c = c_lo; c < c_hi; ++c) {
                    uint32_t r[16];
                    tmem_ld16(lane_base + a.d2col[1] + (uint32_t)(c * 16), r);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int z = c * 4 + q;
                        if (z < Z) {
                            const float *b = bA + c * 16 + q * 4;
                            const float sum = ((__uint_as_float(r[q * 4]) + b[0]) + (__uint_as_float(r[q * 4 + 1]) + b[1])) +
                                              ((__uint_as_float(r[q * 4 + 2]) + b[2]) + (__uint_as_float(r[q * 4 + 3]) + b[3]));
                            c_row[z] -= sum * 0.25f;
                        }
                    }
                }
            }
            epi_bar();
            // expectation (C51) / quantile mean (QR-DQN), split over the actions: this thread owns actions half and half + 2
            float mx[2] = {-INFINITY, -INFINITY}, se[2][4], sq[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) se[i][j] = sq[i][j] = 0.0f;
            const int nchunks = npadA / 16;
            if (a.epi == 1) {
                for (int c = 0; c < nchunks; ++c) {
                    uint32_t r[16];
                    tmem_ld16(lane_base + a.d2col[1] + (uint32_t)(c * 16), r);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int z = c * 4 + q;
                        if (z < Z) {
                            const float cz = c_row[z];
#pragma unroll
                            for (int i = 0; i < 2; ++i)
                                mx[i] = fmaxf(mx[i], (__uint_as_float(half ? r[q * 4 + 1 + 2 * i] : r[q * 4 + 2 * i]) + bA[c * 16 + q * 4 + half + 2 * i]) + cz);
                        }
                    }
                }
            }
            for (int c = 0; c < nchunks; ++c) {
                uint32_t r[16];
                tmem_ld16(lane_base + a.d2col[1] + (uint32_t)(c * 16), r);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int z = c * 4 + q;
                    if (z < Z) {
                        const float cz = c_row[z], sz = s_sup[z];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const float d = (__uint_as_float(half ? r[q * 4 + 1 + 2 * i] : r[q * 4 + 2 * i]) + bA[c * 16 + q * 4 + half + 2 * i]) + cz;
                            if (a.epi == 1) {
                                const float e = fast_exp(d - mx[i]);
                                se[i][q] += e;
                                sq[i][q] += e * sz;
                            } else {
                                sq[i][q] += d;
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&d2_free);
            const int m = tile * kBM + row;
            if (m < a.rows) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float num = (sq[i][0] + sq[i][1]) + (sq[i][2] + sq[i][3]), den = (se[i][0] + se[i][1]) + (se[i][2] + se[i][3]);
                    a.q_out[(size_t)g * a.q_g + (size_t)m * 4 + half + 2 * i] = a.epi == 1 ? num / den : num / (float)Z;
                }
            }
            epi_bar();  // c_row is rewritten by the next tile
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// Splits the heads' weights into TF32 hi / lo panels with TMA-friendly strides (rows of din / 128 floats, 16-byte aligned,
// zero rows up to the padded width) and gathers the biases.  One launch per forward call: ~0.4 MB per agent.
struct PrepArgs {
    const float *params;
    long long P;
    long long w1[2], b1[2], w2[2], b2[2];
    int out[2], npad[2], nheads, din;
    float *w1hi, *w1lo, *w2hi[2], *w2lo[2], *b1cat, *b2p[2];
    int perm_A, perm_Z;  // > 0: head 1's output rows go atom-major (row z * A + k <- k * Z + z), see tc_heads_acting_kernel
    int mean_Z;          // > 0: output row r is the MEAN of the mean_Z source rows r * mean_Z .. (quantile-mean collapse, see below)
};
__global__ void tc_heads_prepare_kernel(const PrepArgs p) {
    const int g = blockIdx.y;
    const float *src = p.params + (size_t)g * p.P;
    const int n1 = p.nheads * kHidden;
    // (32-bit index arithmetic: the panels of one agent are < 2^31 elements -- checked by the host; 64-bit divisions made this
    // kernel 10 us, on the chain of every forward pass of an update)
    const int w1n = n1 * p.din;
    int total = w1n + n1;
    for (int h = 0; h < p.nheads; ++h) total += p.npad[h] * (kHidden + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int j = i;
        if (j < w1n) {
            const int r = j / p.din, k = j - r * p.din, h = r / kHidden;
            const float v = src[p.w1[h] + (long long)(r - h * kHidden) * p.din + k], hi = tf32_rna(v);
            p.w1hi[(size_t)g * w1n + j] = hi;
            p.w1lo[(size_t)g * w1n + j] = v - hi;
            continue;
        }
        j -= w1n;
        if (j < n1) {
            const int h = j / kHidden;
            p.b1cat[(size_t)g * n1 + j] = src[p.b1[h] + (j - h * kHidden)];
            continue;
        }
        j -= n1;
        for (int h = 0; h < p.nheads; ++h) {
            const int wn = p.npad[h] * kHidden;
            if (j < wn) {
                const int r = j / kHidden, k = j - r * kHidden;
                const int rs = (h == 1 && p.perm_A && r < p.out[h]) ? (r % p.perm_A) * p.perm_Z + r / p.perm_A : r;  // source row
                float v = 0.0f;
                if (r < p.out[h]) {
                    if (p.mean_Z) {
                        for (int z = 0; z < p.mean_Z; ++z) v += src[p.w2[h] + ((long long)r * p.mean_Z + z) * kHidden + k];
                        v /= (float)p.mean_Z;
                    } else {
                        v = src[p.w2[h] + (long long)rs * kHidden + k];
                    }
                }
                const float hi = tf32_rna(v);
                p.w2hi[h][(size_t)g * wn + j] = hi;
                p.w2lo[h][(size_t)g * wn + j] = v - hi;
                break;
            }
            j -= wn;
            if (j < p.npad[h]) {
                const int js = (h == 1 && p.perm_A && j < p.out[h]) ? (j % p.perm_A) * p.perm_Z + j / p.perm_A : j;
                float b = 0.0f;
                if (j < p.out[h]) {
                    if (p.mean_Z) {
                        for (int z = 0; z < p.mean_Z; ++z) b += src[p.b2[h] + j * p.mean_Z + z];
                        b /= (float)p.mean_Z;
                    } else {
                        b = src[p.b2[h] + js];
                    }
                }
                p.b2p[h][(size_t)g * p.npad[h] + j] = b;
                break;
            }
            j -= p.npad[h];
        }
    }
}

}  // namespace

namespace maexp {

size_t tc_heads_workspace_floats(int G, int nheads, int din, const int *out) {
    size_t n = 2 * (size_t)G * nheads * kHidden * din + (size_t)G * nheads * kHidden;
    for (int h = 0; h < nheads; ++h) {
        const size_t npad = (size_t)(out[h] + 15) / 16 * 16;
        n += 2 * (size_t)G * npad * kHidden + (size_t)G * npad;
    }
    return n + 64;  // 16-byte alignment slack between the panels
}

namespace {
// arguments of one pass: panel layout, tensor maps, TMEM plan, epilogue mode; false: shape outside what the kernels handle
bool plan_pass(int G, const TcHeadsFwd &f, HeadsArgs &a, PrepArgs &p, bool &persistent) {
    if (f.nheads < 1 || f.nheads > 2 || f.H != kHidden || f.din < kBK || f.din % kBK || f.rows < 1 || !f.workspace) return false;
    if ((reinterpret_cast<uintptr_t>(f.x) & 15) || (f.x_row & 3) || (f.x_g & 3) || f.din > (1 << 20)) return false;
    a = HeadsArgs();
    p = PrepArgs();
    a.nheads = p.nheads = f.nheads;
    a.rows = f.rows;
    a.hbuf_g = f.rows_max * kHidden;
    a.din = p.din = f.din;
    // Quantile-mean collapse (QR-DQN acting): the greedy value is the MEAN over the quantiles of v + (adv - mean over actions)
    // (qr_dqn_brain.py get_action), and the mean commutes with the linear second layer -- so the second layer shrinks from
    // Z + A Z output rows (100 + 400) to 1 + A rows whose weights are the quantile means of the original rows.
    const bool collapse = f.q_out && !f.categorical && f.nheads == 2 && f.Z >= 1 && f.A >= 1 && f.A <= 8 && f.out[0] == f.Z &&
                          f.out[1] == f.A * f.Z && !f.hbuf[0] && !f.hbuf[1];
    if (collapse) p.mean_Z = f.Z;
    for (int h = 0; h < f.nheads; ++h) {
        const int eff_out = collapse ? (h ? f.A : 1) : f.out[h];
        a.out_n[h] = p.out[h] = eff_out;
        a.npad[h] = p.npad[h] = (eff_out + 15) / 16 * 16;
        if (f.out[h] < 1 || a.npad[h] > 256) return false;
        a.out[h] = f.outp[h];
        a.out_g[h] = f.rows_max * f.out[h];
        a.hbuf[h] = f.hbuf[h];
        if (f.hbuf[h] && (reinterpret_cast<uintptr_t>(f.hbuf[h]) & 15)) return false;
        p.w1[h] = f.w1[h]; p.b1[h] = f.b1[h]; p.w2[h] = f.w2[h]; p.b2[h] = f.b2[h];
    }
    // TMEM plan: first accumulator in columns [0, heads * 128); the widest head's output at column 256, the other one in the
    // hidden columns of the head processed first (free once its four slabs have been read)
    if (f.nheads == 2) {
        const int big = a.npad[1] >= a.npad[0] ? 1 : 0, small = 1 - big;
        if (a.npad[small] > kHidden) return false;
        a.order[0] = big; a.order[1] = small;
        a.d2col[big] = 256;
        a.d2col[small] = (uint32_t)(big * kHidden);
    } else {
        a.order[0] = 0;
        a.d2col[0] = 256;
    }
    // workspace carve-up (16-byte aligned panels)
    float *w = f.workspace;
    auto take = [&](size_t n) { float *r = w; w += (n + 3) & ~(size_t)3; return r; };
    const int n1 = f.nheads * kHidden;
    p.params = f.params; p.P = f.P;
    p.w1hi = take((size_t)G * n1 * f.din);
    p.w1lo = take((size_t)G * n1 * f.din);
    p.b1cat = take((size_t)G * n1);
    for (int h = 0; h < f.nheads; ++h) {
        p.w2hi[h] = take((size_t)G * a.npad[h] * kHidden);
        p.w2lo[h] = take((size_t)G * a.npad[h] * kHidden);
        p.b2p[h] = take((size_t)G * a.npad[h]);
        a.b2p[h] = p.b2p[h];
    }
    a.b1cat = p.b1cat;
    if (reinterpret_cast<uintptr_t>(f.workspace) & 15) return false;
    bool ok = make_map_f32(&a.tm_x, f.x, (uint64_t)f.din, (uint64_t)f.rows, (uint64_t)G, (uint64_t)f.x_row * 4, (uint64_t)f.x_g * 4, kBM);
    ok = ok && make_map_f32(&a.tm_w1hi, p.w1hi, (uint64_t)f.din, (uint64_t)G * n1, 0, (uint64_t)f.din * 4, 0, (uint32_t)n1);
    ok = ok && make_map_f32(&a.tm_w1lo, p.w1lo, (uint64_t)f.din, (uint64_t)G * n1, 0, (uint64_t)f.din * 4, 0, (uint32_t)n1);
    for (int h = 0; h < f.nheads && ok; ++h) {
        ok = ok && make_map_f32(&a.tm_w2hi[h], p.w2hi[h], kHidden, (uint64_t)G * a.npad[h], 0, kHidden * 4, 0, (uint32_t)a.npad[h]);
        ok = ok && make_map_f32(&a.tm_w2lo[h], p.w2lo[h], kHidden, (uint64_t)G * a.npad[h], 0, kHidden * 4, 0, (uint32_t)a.npad[h]);
    }
    if (!ok) return false;
    a.rs = (a.npad[0] + (f.nheads > 1 ? a.npad[1] : 0)) | 1;  // the epilogue's shared-memory copy of the accumulators
    if ((size_t)a.rs * kBM * 4 > (size_t)kStages * kStageBytes || a.rs > 512) return false;
    if (collapse) {
        a.epi = 3;
        a.A = f.A; a.Z = 1;
        a.support = nullptr;
        a.q_out = f.q_out;
        a.q_g = f.q_g;
    } else if (f.q_out && f.nheads == 2 && f.A >= 1 && f.A <= 8 && f.Z >= 1 && f.Z <= 256 && f.out[0] == f.Z && f.out[1] == f.A * f.Z) {
        a.epi = f.categorical ? 1 : 2;
        a.A = f.A; a.Z = f.Z;
        a.support = f.support;
        a.q_out = f.q_out;
        a.q_g = f.q_g;
    }
    // acting batches of the 4-action dueling distributional networks: persistent kernel, epilogue overlapped with the next tile
    static const bool acting_off = getenv("MAEXP_TC_HEADS_PERSISTENT") && atoi(getenv("MAEXP_TC_HEADS_PERSISTENT")) == 0;
    persistent = (a.epi == 1 || a.epi == 2) && f.A == 4 && f.Z <= 52 && a.npad[0] <= 64 && a.npad[0] + a.npad[1] <= 288 && a.order[0] == 1 &&
                            !f.hbuf[0] && !f.hbuf[1] && !acting_off;
    if (persistent) { p.perm_A = 4; p.perm_Z = f.Z; }
    if (f.fused_q_done) *f.fused_q_done = a.epi != 0;
    return true;
}
bool set_smem_attr() {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(tc_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kStages * kStageBytes + 1024)) != cudaSuccess)
            return false;
        attr_set = true;
    }
    return true;
}
}  // namespace

// the forward passes of an update (same shapes, their own inputs / parameters / outputs) as ONE launch: blockIdx.z = pass
// prep != nullptr: the weight-panel kernels (which read parameters only) run on prep->stream behind prep->after -- a point of the
// pass before its trunk -- and the heads kernel waits for them: in a captured update they become a branch beside the trunk
// instead of 4 us per head on the pass's chain
bool tc_heads_forward_multi(cudaStream_t st, int G, const TcHeadsFwd *f, int n, const TcHeadsPrepBranch *prep) {
    if (n < 1 || n > 3) return false;
    HeadsArgsN aa = {};
    PrepArgs p[3];
    for (int q = 0; q < n; ++q) {
        bool persistent = false;
        if (f[q].rows != f[0].rows || f[q].din != f[0].din || f[q].nheads != f[0].nheads) return false;
        if (!plan_pass(G, f[q], aa.p[q], p[q], persistent)) return false;
        if (persistent) {  // (acting batches go through tc_heads_forward)
            p[q].perm_A = p[q].perm_Z = 0;
        }
    }
    if (!set_smem_attr()) return false;
    if (prep) {
        cudaStreamWaitEvent(prep->stream, prep->after, 0);
        for (int q = 0; q < n; ++q) tc_heads_prepare_kernel<<<dim3(128, G), 256, 0, prep->stream>>>(p[q]);
        cudaEventRecord(prep->done, prep->stream);
        cudaStreamWaitEvent(st, prep->done, 0);
    } else {
        for (int q = 0; q < n; ++q) tc_heads_prepare_kernel<<<dim3(128, G), 256, 0, st>>>(p[q]);
    }
    tc_heads_kernel<<<dim3((f[0].rows + kBM - 1) / kBM, G, n), kThreads, (size_t)kStages * kStageBytes + 1024, st>>>(aa);
    return cudaGetLastError() == cudaSuccess;
}

bool tc_heads_forward(cudaStream_t st, int G, const TcHeadsFwd &f) {
    HeadsArgsN aa = {};
    HeadsArgs &a = aa.p[0];
    PrepArgs p;
    bool persistent = false;
    if (!plan_pass(G, f, a, p, persistent)) return false;
    static long long *dbg_buf = nullptr;
    if (getenv("MAEXP_TC_HEADS_DBG")) {
        if (!dbg_buf) { cudaMalloc(&dbg_buf, 256 * sizeof(long long)); cudaMemset(dbg_buf, 0, 256 * sizeof(long long)); }
        a.dbg = dbg_buf;
    }
    const size_t smem = (size_t)kStages * kStageBytes + 1024;
    if (!set_smem_attr()) return false;
    tc_heads_prepare_kernel<<<dim3(128, G), 256, 0, st>>>(p);
    if (persistent) {
        const size_t smem_p = (size_t)kStages * kStageBytes + 1024 + sizeof(float) * kBM * kCPitch;
        static bool attr_p = false;
        if (!attr_p) {
            if (cudaFuncSetAttribute(tc_heads_acting_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p) != cudaSuccess) return false;
            attr_p = true;
        }
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int ntiles = (f.rows + kBM - 1) / kBM;
        int per_group = sms / G < 1 ? 1 : sms / G;
        if (per_group > ntiles) per_group = ntiles;
        tc_heads_acting_kernel<<<dim3(per_group, G), kThreadsP, smem_p, st>>>(a);
        return cudaGetLastError() == cudaSuccess;
    }
    tc_heads_kernel<<<dim3((f.rows + kBM - 1) / kBM, G), kThreads, smem, st>>>(aa);
    if (a.dbg) {
        long long host[256];
        cudaStreamSynchronize(st);
        cudaMemcpy(host, dbg_buf, sizeof(host), cudaMemcpyDeviceToHost);
        const int T = f.din / kBK + 4 * f.nheads;
        fprintf(stderr, "tc_heads CTA(0,0) cycles, %d slabs: slab | producer waits empty | MMA waits full | MMA waits conv | MMA start (rel)\n", T);
        for (int t = 0; t < T && t < 32; ++t)
            fprintf(stderr, "  %2d %8lld %8lld %8lld %9lld\n", t, host[t], host[32 + t], host[64 + t], host[96 + t] - host[96]);
        fprintf(stderr, "  epilogue waited %lld, epilogue took %lld, total from first MMA wait %lld\n", host[128], host[130] - host[129], host[130] - host[96]);
    }
    return cudaGetLastError() == cudaSuccess;
}

}  // namespace maexp
