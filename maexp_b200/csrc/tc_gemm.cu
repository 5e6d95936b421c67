// tcgen05 (5th-generation tensor core) GEMM for the large acting / evaluation batches:
//     C[M,N] = epi( A[M,K] * W[N,K]^T + bias ),   fp32 in, fp32 out, accumulators in tensor memory (TMEM).
// The operands are fp32; the tensor cores take them as TF32.  With split = 3 every product is evaluated as
// a_hi*b_hi + a_lo*b_hi + a_hi*b_lo (a_hi = rna-tf32(a), a_lo = a - a_hi), which keeps the result within ~2^-21 of the
// fp32 product sum, so the actor's Q-values stay inside the 1e-5 bar of the fp32 learner.  split = 1 is plain TF32.
//
// One CTA (8 warps) owns a 128-row tile and all N <= 256 columns:
//   * all threads stage a 128 x 32 slab of A and an Npad x 32 slab of W into shared memory in the K-major
//     SWIZZLE_128B canonical layout (rows of 128 bytes, 16-byte chunk c of row r stored at chunk c ^ (r & 7));
//     fence.proxy.async makes the generic-proxy stores visible to the tensor core.  One slab per CTA (16-64 KB), so
//     several CTAs are resident per SM and one CTA's loads overlap another's MMAs
//   * one thread issues tcgen05.mma.cta_group::1.kind::tf32 (UMMA 128 x Npad x 8, four per slab and term) and
//     tcgen05.commit, whose mbarrier arrival releases the slab for the next load
//   * epilogue: each warp reads its 32 TMEM lanes (= rows) with tcgen05.ld.32x32b, adds bias / ReLU, stores rows.
// No TMA descriptors: the weights sit at arbitrary float offsets of the flat per-agent parameter vector.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <string>

#include "maexp_b200.h"
#include "tc_common.cuh"

namespace maexp {
int fail(const std::string &msg);  // env.cu
}

using namespace maexp_tc;

namespace {

struct TcArgs {
    const float *A, *W, *bias;
    float *C;
    int M, N, K, npad;
    long long lda, ldw, ldc, gA, gW, gBias, gC;
    int relu, split;
};

// dynamic shared memory: [A_hi | A_lo | W_hi | W_lo], each tile 1024-byte aligned
constexpr int kThreads = 256;

template <int SPLIT>
__global__ void __launch_bounds__(kThreads) tc_gemm_kernel(const TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = blockIdx.y;
    // wide layers: blockIdx.z walks column tiles of 256; a.N / a.npad below are this tile's
    const int nt0 = blockIdx.z * 256, tileN = min(256, a.N - nt0), npad = (tileN + 15) / 16 * 16;
    const int m0 = blockIdx.x * kBM;
    const float *A = a.A + (size_t)g * a.gA, *W = a.W + (size_t)g * a.gW + (size_t)nt0 * a.ldw;
    const uint32_t a_bytes = kBM * 128, w_bytes = (uint32_t)npad * 128;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
    uint8_t *sA = smem, *sAlo = sA + a_bytes;
    uint8_t *sW = sA + (SPLIT == 3 ? 2 : 1) * a_bytes, *sWlo = sW + w_bytes;
    uint32_t ncols = 32;
    while (ncols < (uint32_t)npad) ncols <<= 1;

    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_smem;
    // instruction descriptor: D = F32 (bits 4-5 = 1), A and B = TF32 (bits 7-9 and 10-12 = 2), both K-major,
    // N >> 3 in bits 17-22, M >> 4 in bits 24-28
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

    const int num_kb = (a.K + kBK - 1) / kBK;
    const bool a_vec = (a.lda & 3) == 0 && (a.gA & 3) == 0 && ((uintptr_t)a.A & 15) == 0;
    const bool w_vec = (a.ldw & 3) == 0 && (a.gW & 3) == 0 && ((uintptr_t)a.W & 15) == 0;
    // register prefetch: the global loads of slab kb + 1 are issued right after slab kb's MMAs and stay in flight
    // while those run and while this CTA waits for them (4 A chunks and up to 8 W chunks of 16 bytes per thread)
    float4 ra[kBM * 8 / kThreads], rw[256 * 8 / kThreads];
    auto prefetch = [&](int k0) {
#pragma unroll
        for (int j = 0; j < kBM * 8 / kThreads; ++j) {
            const int i = tid + j * kThreads, r = i >> 3, c = i & 7, m = m0 + r;
            ra[j] = m < a.M ? load_chunk(A + (size_t)m * a.lda + k0 + c * 4, k0 + c * 4, a.K, a_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 256 * 8 / kThreads; ++j) {
            const int i = tid + j * kThreads, r = i >> 3, c = i & 7;
            if (i < npad * 8)
                rw[j] = r < tileN ? load_chunk(W + (size_t)r * a.ldw + k0 + c * 4, k0 + c * 4, a.K, w_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    prefetch(0);
    for (int kb = 0; kb < num_kb; ++kb) {
        if (kb >= 1) mbar_wait(&bar, (kb - 1) & 1);  // the MMAs of the previous slab have finished reading shared memory
#pragma unroll
        for (int j = 0; j < kBM * 8 / kThreads; ++j) {
            const int i = tid + j * kThreads;
            stage_chunk<SPLIT>(sA, sAlo, i >> 3, i & 7, ra[j]);
        }
#pragma unroll
        for (int j = 0; j < 256 * 8 / kThreads; ++j) {
            const int i = tid + j * kThreads;
            if (i < npad * 8) stage_chunk<SPLIT>(sW, sWlo, i >> 3, i & 7, rw[j]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t aA = smem_u32(sA), aAlo = smem_u32(sAlo), aW = smem_u32(sW), aWlo = smem_u32(sWlo);
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k) {  // UMMA_K = 8 TF32 = 32 bytes: advance the start address inside the swizzle row
                const uint32_t ko = (uint32_t)k * 32;
                umma_tf32(tmem_d, smem_desc(aA + ko), smem_desc(aW + ko), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                if (SPLIT == 3) {
                    umma_tf32(tmem_d, smem_desc(aAlo + ko), smem_desc(aW + ko), idesc, 1u);
                    umma_tf32(tmem_d, smem_desc(aA + ko), smem_desc(aWlo + ko), idesc, 1u);
                }
            }
            umma_commit(&bar);  // arrives when the MMAs above have finished
        }
        if (kb + 1 < num_kb) prefetch((kb + 1) * kBK);
    }
    mbar_wait(&bar, (num_kb - 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: warp w owns TMEM lanes 32w .. 32w+31 (= rows), 8 columns per tcgen05.ld
    // ---- (a warp may only read the TMEM lane quadrant warp % 4: warps 0-3 take the even 16-column groups, 4-7 the odd
    //      ones).  Sixteen consecutive columns per load = 64 contiguous bytes of the thread's output row.
    const int quad = warp & 3, m = m0 + quad * 32 + lane;
    float *crow = a.C + (size_t)g * a.gC + (size_t)m * a.ldc + nt0;
    const float *bias = a.bias ? a.bias + (size_t)g * a.gBias + nt0 : nullptr;
    const bool c_vec = (a.ldc & 3) == 0 && (a.gC & 3) == 0 && ((uintptr_t)a.C & 15) == 0;
    for (int n0 = (warp >> 2) * 16; n0 < npad; n0 += 32) {
        uint32_t r[16];
        const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)n0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < a.M) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = n0 + j;
                float x = __uint_as_float(r[j]);
                if (bias && n < tileN) x += bias[n];
                if (a.relu) x = fmaxf(x, 0.0f);
                v[j] = x;
            }
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const int n = n0 + j4 * 4;
                if (c_vec && n + 3 < tileN) {
                    *reinterpret_cast<float4 *>(crow + n) = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n + j < tileN) crow[n + j] = v[j4 * 4 + j];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(ncols) : "memory");
}


// ---- weight-gradient GEMMs of the update (the batch is the K dimension): C[M,N] = sum_k A[k][m] * B[k][n], both operands
// ---- stored k-outermost (A = dY [rows][M], B = X [rows][N]).  The tensor core wants both K-major, so the staging transposes:
// ---- a thread loads a 4 (k) x 4 (m) block as four 16-byte row pieces and stores four 16-byte K-chunks, one per row of the
// ---- swizzled tile (lanes of a quarter warp hold the 8 chunks of one row: conflict-free stores, 64-byte global segments).
// ---- Three TF32 terms per product; split-K slices write dense partial tiles part[slice][g][M][N] for the learner's
// ---- fixed-order reduction (deterministic, like its CUDA-core split-K GEMM).
struct TnArgs {
    const float *A, *B;
    float *C;  // ksplit == 1: C[m][n] (row stride ldc, group stride gC); else the partial tiles
    int M, N, K, tileN, ntiles, ksplit, kslice;
    long long lda, ldb, ldc, gA, gB, gC;
};

__device__ __forceinline__ void load_block_t(const float *base, long long ld, int col, int limit, bool vec, int k0, int kend,
                                             float4 (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int k = k0 + i;
        if (k < kend && col < limit) {
            const float *src = base + (size_t)k * ld + col;
            if (vec && col + 3 < limit) v[i] = *reinterpret_cast<const float4 *>(src);
            else {
                v[i].x = src[0];
                if (col + 1 < limit) v[i].y = src[1];
                if (col + 2 < limit) v[i].z = src[2];
                if (col + 3 < limit) v[i].w = src[3];
            }
        }
    }
}
__device__ __forceinline__ void stage_block_t(uint8_t *hi, uint8_t *lo, int q, int c, const float4 (&v)[4]) {
    stage_chunk<3>(hi, lo, 4 * q + 0, c, make_float4(v[0].x, v[1].x, v[2].x, v[3].x));
    stage_chunk<3>(hi, lo, 4 * q + 1, c, make_float4(v[0].y, v[1].y, v[2].y, v[3].y));
    stage_chunk<3>(hi, lo, 4 * q + 2, c, make_float4(v[0].z, v[1].z, v[2].z, v[3].z));
    stage_chunk<3>(hi, lo, 4 * q + 3, c, make_float4(v[0].w, v[1].w, v[2].w, v[3].w));
}

__global__ void __launch_bounds__(kThreads, 3) tc_gemm_tn_kernel(const TnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = blockIdx.y;
    const int nt = blockIdx.z % a.ntiles, slice = blockIdx.z / a.ntiles;
    const int n0t = nt * a.tileN, tileN = min(a.tileN, a.N - n0t), npad = (tileN + 15) / 16 * 16;
    const int m0 = blockIdx.x * kBM, tileM = min(kBM, a.M - m0);
    const int kbeg = slice * a.kslice, kend = min(a.K, kbeg + a.kslice);
    const float *A = a.A + (size_t)g * a.gA + m0, *B = a.B + (size_t)g * a.gB + n0t;
    const uint32_t a_bytes = kBM * 128, w_bytes = (uint32_t)npad * 128;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sA = smem, *sAlo = sA + a_bytes, *sW = sA + 2 * a_bytes, *sWlo = sW + w_bytes;
    uint32_t ncols = 32;
    while (ncols < (uint32_t)npad) ncols <<= 1;
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_smem;
    const uint32_t idesc = idesc_tf32_m128(npad);
    const int num_kb = (kend - kbeg + kBK - 1) / kBK;
    const bool a_vec = (a.lda & 3) == 0 && (a.gA & 3) == 0 && ((uintptr_t)a.A & 15) == 0;
    const bool b_vec = (a.ldb & 3) == 0 && (a.gB & 3) == 0 && ((uintptr_t)a.B & 15) == 0;
    // one 4 x 4 block of A per thread (32 row quads x 8 K-chunks), up to two of B (<= 64 row quads x 8 K-chunks)
    const int c = tid & 7, q = tid >> 3;
    const int nbq = npad >> 2;  // row quads of the B tile
    float4 ra[4], rb0[4], rb1[4];
    auto prefetch = [&](int k0) {
        load_block_t(A, a.lda, 4 * q, tileM, a_vec, k0 + 4 * c, kend, ra);
        if (q < nbq) load_block_t(B, a.ldb, 4 * q, tileN, b_vec, k0 + 4 * c, kend, rb0);
        if (q + 32 < nbq) load_block_t(B, a.ldb, 4 * (q + 32), tileN, b_vec, k0 + 4 * c, kend, rb1);
    };
    if (num_kb > 0) prefetch(kbeg);
    for (int kb = 0; kb < num_kb; ++kb) {
        if (kb >= 1) mbar_wait(&bar, (kb - 1) & 1);
        stage_block_t(sA, sAlo, q, c, ra);
        if (q < nbq) stage_block_t(sW, sWlo, q, c, rb0);
        if (q + 32 < nbq) stage_block_t(sW, sWlo, q + 32, c, rb1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t aA = smem_u32(sA), aAlo = smem_u32(sAlo), aW = smem_u32(sW), aWlo = smem_u32(sWlo);
#pragma unroll
            for (int k = 0; k < kBK / 8; ++k) {
                const uint32_t ko = (uint32_t)k * 32;
                umma_tf32(tmem_d, smem_desc(aA + ko), smem_desc(aW + ko), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                umma_tf32(tmem_d, smem_desc(aAlo + ko), smem_desc(aW + ko), idesc, 1u);
                umma_tf32(tmem_d, smem_desc(aA + ko), smem_desc(aWlo + ko), idesc, 1u);
            }
            umma_commit(&bar);
        }
        if (kb + 1 < num_kb) prefetch(kbeg + (kb + 1) * kBK);
    }
    if (num_kb > 0) mbar_wait(&bar, (num_kb - 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warp w reads TMEM lane quadrant w % 4 (rows), warps 0-3 the even 16-column groups, 4-7 the odd ones
    const int quad = warp & 3, ml = quad * 32 + lane, m = m0 + ml;
    float *crow;
    long long ldc;
    if (a.ksplit > 1) {
        crow = a.C + (((size_t)slice * gridDim.y + g) * a.M + m) * a.N + n0t;
        ldc = a.N;
    } else {
        crow = a.C + (size_t)g * a.gC + (size_t)m * a.ldc + n0t;
        ldc = a.ldc;
    }
    const bool c_vec = (ldc & 3) == 0 && ((a.ksplit > 1) ? (((size_t)a.M * a.N) & 3) == 0 : (a.gC & 3) == 0) && ((uintptr_t)a.C & 15) == 0;
    for (int n0 = (warp >> 2) * 16; n0 < npad; n0 += 32) {
        uint32_t r[16];
        if (num_kb > 0) tmem_ld16(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)n0, r);
        else {
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = 0u;
        }
        if (ml < tileM) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const int n = n0 + j4 * 4;
                if (c_vec && n + 3 < tileN) {
                    *reinterpret_cast<float4 *>(crow + n) = make_float4(__uint_as_float(r[j4 * 4]), __uint_as_float(r[j4 * 4 + 1]),
                                                                        __uint_as_float(r[j4 * 4 + 2]), __uint_as_float(r[j4 * 4 + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n + j < tileN) crow[n + j] = __uint_as_float(r[j4 * 4 + j]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(ncols) : "memory");
}

}  // namespace

namespace maexp {

// grouped launch used by the learner's inference path; returns false when the shape is outside what the kernel handles
bool tc_gemm_launch(cudaStream_t st, int G, const float *A, const float *W, const float *bias, float *C, int M, int N, int K,
                    long long lda, long long ldw, long long ldc, long long gA, long long gW, long long gBias, long long gC, int relu,
                    int split) {
    if (N < 1 || M < 1 || K < 1) return false;
    TcArgs a;
    a.A = A; a.W = W; a.bias = bias; a.C = C;
    a.M = M; a.N = N; a.K = K; a.npad = ((N < 256 ? N : 256) + 15) / 16 * 16;  // widest column tile (sizes shared memory)
    a.lda = lda; a.ldw = ldw; a.ldc = ldc; a.gA = gA; a.gW = gW; a.gBias = gBias; a.gC = gC;
    a.relu = relu; a.split = split;
    const size_t smem = (size_t)(split == 3 ? 2 : 1) * ((size_t)kBM * 128 + (size_t)a.npad * 128) + 1024;
    if (smem > 200 * 1024) return false;
    auto fn = split == 3 ? tc_gemm_kernel<3> : tc_gemm_kernel<1>;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    fn<<<dim3((M + kBM - 1) / kBM, G, (N + 255) / 256), kThreads, smem, st>>>(a);
    return true;
}


// grouped weight-gradient launch: C[g][M,N] = sum_k A[g][k][m] B[g][k][n].  *ksplit_out > 1: the kernel wrote partial tiles
// part[slice][g][M][N] into `part` (the caller reduces them in slice order); returns false when the shape is outside what the
// kernel handles (the caller's CUDA-core GEMM then runs)
bool tc_gemm_tn_launch(cudaStream_t st, int G, const float *A, const float *B, float *C, int M, int N, int K, long long lda,
                       long long ldb, long long ldc, long long gA, long long gB, long long gC, float *part, long long part_floats,
                       int *ksplit_out) {
    if (N < 16 || M < 16 || K < 64) return false;
    TnArgs a;
    a.A = A; a.B = B; a.M = M; a.N = N; a.K = K;
    a.ntiles = (N + 255) / 256;
    a.tileN = ((N + a.ntiles - 1) / a.ntiles + 15) / 16 * 16;
    a.lda = lda; a.ldb = ldb; a.ldc = ldc; a.gA = gA; a.gB = gB; a.gC = gC;
    // split K into slices of >= 12 slabs when one CTA's chain would be long (the conv2 weight gradient: K = B * positions)
    // MAEXP_TC_WGRAD_SLABS = K slabs (of 32) per CTA: a single-buffered CTA pays a chain of latencies per slab, so the
    // slices are short and many (default 12 slabs for K >= 1024, no split below)
    static const int slabs_env = getenv("MAEXP_TC_WGRAD_SLABS") ? atoi(getenv("MAEXP_TC_WGRAD_SLABS")) : 0;
    int ksplit = 1;
    if (part && slabs_env > 0) {
        ksplit = (K + 32 * slabs_env - 1) / (32 * slabs_env);
        if (ksplit > 32) ksplit = 32;
        while (ksplit > 1 && (long long)ksplit * G * M * N > part_floats) --ksplit;
    } else if (part && K >= 1024) {
        ksplit = K / 384;
        if (ksplit > 8) ksplit = 8;
        while (ksplit > 1 && (long long)ksplit * G * M * N > part_floats) --ksplit;
    }
    a.ksplit = ksplit;
    a.kslice = ((K + ksplit - 1) / ksplit + kBK - 1) / kBK * kBK;
    a.C = ksplit > 1 ? part : C;
    const int npad = (a.tileN + 15) / 16 * 16;
    const size_t smem = 2 * ((size_t)kBM * 128 + (size_t)npad * 128) + 1024;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(tc_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (kBM * 128 + 256 * 128) + 1024) != cudaSuccess)
            return false;
        attr = true;
    }
    tc_gemm_tn_kernel<<<dim3((M + kBM - 1) / kBM, G, a.ntiles * ksplit), kThreads, smem, st>>>(a);
    *ksplit_out = ksplit;
    return true;
}
}  // namespace maexp

extern "C" int maexp_tc_gemm(const float *A, const float *W, const float *bias, float *C, int32_t M, int32_t N, int32_t K,
                             int64_t lda, int64_t ldw, int64_t ldc, int32_t relu, int32_t split, void *stream) {
    if (!A || !W || !C) return maexp::fail("maexp_tc_gemm: null argument");
    if (split != 1 && split != 3) return maexp::fail("maexp_tc_gemm: split must be 1 (TF32) or 3 (3xTF32)");
    if (!maexp::tc_gemm_launch(static_cast<cudaStream_t>(stream), 1, A, W, bias, C, M, N, K, lda, ldw, ldc, 0, 0, 0, 0, relu, split))
        return maexp::fail("maexp_tc_gemm: unsupported shape");
    if (cudaGetLastError() != cudaSuccess) return maexp::fail("maexp_tc_gemm: launch failed");
    return 0;
}

extern "C" int maexp_tc_gemm_tn(const float *A, const float *B, float *C, float *partials, int64_t partials_floats, int32_t M, int32_t N,
                                int32_t K, int64_t lda, int64_t ldb, int64_t ldc, int32_t *ksplit_out, void *stream) {
    if (!A || !B || !C) return maexp::fail("maexp_tc_gemm_tn: null argument");
    int ks = 1;
    if (!maexp::tc_gemm_tn_launch(static_cast<cudaStream_t>(stream), 1, A, B, C, M, N, K, lda, ldb, ldc, 0, 0, 0, partials, partials_floats, &ks))
        return maexp::fail("maexp_tc_gemm_tn: unsupported shape (M, N >= 16, K >= 64)");
    if (cudaGetLastError() != cudaSuccess) return maexp::fail("maexp_tc_gemm_tn: launch failed");
    if (ksplit_out) *ksplit_out = ks;
    return 0;
}
