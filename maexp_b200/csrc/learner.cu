// Data-parallel learner update for the per-agent value networks, sm_100a.
//
// Replaces (for G = num_agents independent networks at once, grouped in every launch):
//   DA3Brain.learn / DQNBrain.learn   core/agents/brains/da3_brain.py:58-89, dqn_brain.py:20-49   (TD target + MSE)
//   CategoricalDQNBrain.learn         core/agents/brains/categorical_dqn_brain.py:42-89           (C51 projection + CE)
//   QRDQNBrain.learn                  core/agents/brains/qr_dqn_brain.py:15-51,81-119             (quantile Huber)
//   Conv / MLP / ConvMLP / CategoricalDQN / QRDQN forward+backward
//                                      core/agents/models/convolution.py:14-55, mlp.py:14-34,
//                                      customs/conv_mlp.py:21-66, categorical_dqn.py:28-79, qr_dqn.py:20-65
//   clip_grad_norm_ + optim.RMSprop   brains/*.py, core/agents/optimizers/__init__.py:28-33
//   hard_update                        core/utils/updates.py:25-34
//
// Numerics: fp32 FMA-free SIMT arithmetic in a fixed summation order (deterministic).  The parity
// bar is 1e-5 relative on losses and Q-values against torch fp32 eager, which
// rules out TF32 tensor-core math at these sizes; the networks are 42k-140k parameters and the
// update is launch-latency bound, so every layer is ONE grouped launch for
// all agents and the whole update is CUDA-graph capturable (no host sync, no allocation).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/maexp_b200.h"
#include "tc_heads.h"
#include "tc_mlp.h"
#include "tc_trunk.h"

namespace maexp {
int fail(const std::string &msg);  // env.cu
// tc_gemm.cu: tcgen05 tensor-core GEMM (3xTF32) for the large inference batches
bool tc_gemm_launch(cudaStream_t st, int G, const float *A, const float *W, const float *bias, float *C, int M, int N, int K,
                    long long lda, long long ldw, long long ldc, long long gA, long long gW, long long gBias, long long gC, int relu,
                    int split);
// tc_gemm.cu: the weight-gradient products (batch = K, both operands k-outermost) on tcgen05, transposed while staging
bool tc_gemm_tn_launch(cudaStream_t st, int G, const float *A, const float *B, float *C, int M, int N, int K, long long lda,
                       long long ldb, long long ldc, long long gA, long long gB, long long gC, float *part, long long part_floats,
                       int *ksplit_out);
#define MAEXP_CUDA_L(call)                                                                        \
    do {                                                                                          \
        cudaError_t err__ = (call);                                                               \
        if (err__ != cudaSuccess)                                                                 \
            return fail(std::string(#call) + ": " + cudaGetErrorString(err__));                   \
    } while (0)

// ------------------------------------------------------------------------------ grouped GEMM
// C[g][m,n] (=|+=) epi( sum_k A[g][m,k] * B[g][k,n] + bias[g][n] ), arbitrary element strides.
struct GemmArgs {
    const float *A, *B, *bias, *aux;
    float *C;
    int M, N, K;
    long long sam, sak, sbk, sbn, scm, scn;  // element strides
    long long gA, gB, gC, gBias, gAux;       // per-group offsets
    int relu;        // max(x, 0)
    int accumulate;  // C += result
    int mask_aux;    // result *= (aux[m,n] > 0)   (ReLU backward; aux uses C's strides)
};

constexpr int kTile = 64, kBK = 16;
constexpr int kSides = 4;  // side streams: forward passes 1 and 2 use the first two, weight gradients rotate over all
constexpr int kTrunkBwdBlocks = 64;  // most blocks per agent of trunk_fused_bwd_kernel (= partial sums per gradient element)

// Deterministic split-K: slice z of the K range accumulates into a dense partial tile
// part[slice][g][M][N]; splitk_reduce_kernel then adds the slices in a fixed order and applies the
// epilogue.  With ksplit == 1 the epilogue runs here and C is written directly.
__global__ void __launch_bounds__(256) gemm_kernel(const GemmArgs a, int ksplit, int kslice, float *part) {
    __shared__ __align__(16) float As[kBK][kTile + 4];
    __shared__ __align__(16) float Bs[kBK][kTile + 4];
    const int g = blockIdx.z / ksplit, slice = blockIdx.z - g * ksplit;
    const float *A = a.A + (size_t)g * a.gA, *B = a.B + (size_t)g * a.gB;
    const int m0 = blockIdx.y * kTile, n0 = blockIdx.x * kTile;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int kbeg = slice * kslice, kend = min(a.K, kbeg + kslice);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    const bool a_k_fast = a.sak == 1, b_n_fast = a.sbn == 1;
    // per-thread tile coordinates of the 4 A and 4 B elements it stages (coalesced along the unit stride)
    int amm[4], akk[4], bnn[4], bkk[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int idx = tid + r * 256;
        if (a_k_fast) { akk[r] = idx & 15; amm[r] = idx >> 4; } else { amm[r] = idx & 63; akk[r] = idx >> 6; }
        if (b_n_fast) { bnn[r] = idx & 63; bkk[r] = idx >> 6; } else { bkk[r] = idx & 15; bnn[r] = idx >> 4; }
    }
    float ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int m = m0 + amm[r], k = k0 + akk[r];
            ra[r] = (m < a.M && k < kend) ? A[(size_t)m * a.sam + (size_t)k * a.sak] : 0.0f;
            const int n = n0 + bnn[r], k2 = k0 + bkk[r];
            rb[r] = (n < a.N && k2 < kend) ? B[(size_t)k2 * a.sbk + (size_t)n * a.sbn] : 0.0f;
        }
    };
    fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += kBK) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            As[akk[r]][amm[r]] = ra[r];
            Bs[bkk[r]][bnn[r]] = rb[r];
        }
        __syncthreads();
        if (k0 + kBK < kend) fetch(k0 + kBK);  // next tile's loads overlap this tile's math
#pragma unroll
        for (int kk = 0; kk < kBK; ++kk) {
            // 128-bit shared loads and fused multiply-adds (the 1e-5 learner bar does not need separately rounded products;
            // the env kernel's bit-exact noise arithmetic is what --fmad=false is for)
            const float4 a4 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *C = a.C + (size_t)g * a.gC;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= a.N) continue;
            float v = acc[i][j];
            if (ksplit > 1) {
                part[(((size_t)slice * gridDim.z / ksplit + g) * a.M + m) * a.N + n] = v;
                continue;
            }
            if (a.bias) v = __fadd_rn(v, a.bias[(size_t)g * a.gBias + n]);
            if (a.relu) v = fmaxf(v, 0.0f);
            const size_t ci = (size_t)m * a.scm + (size_t)n * a.scn;
            if (a.mask_aux) v = (a.aux[(size_t)g * a.gAux + ci] > 0.0f) ? v : 0.0f;
            if (a.accumulate) v = __fadd_rn(C[ci], v);
            C[ci] = v;
        }
    }
}

// Large-M variant for the acting / evaluation batches (thousands of rows per agent): 128 x 64 x 16 tiles, 8 x 4 outputs per
// thread, 128-bit global loads along K for both operands (activations [M,K] and PyTorch-layout weights [N,K] are K-major),
// register prefetch of the next tile.  Same epilogue as gemm_kernel; no split-K (the grid is large already).
constexpr int kBigM = 128, kBigN = 64, kBigK = 16;
__global__ void __launch_bounds__(256) gemm_big_kernel(const GemmArgs a) {
    __shared__ __align__(16) float As[kBigK][kBigM + 4];
    __shared__ __align__(16) float Bs[kBigK][kBigN + 4];
    const int g = blockIdx.z;
    const float *A = a.A + (size_t)g * a.gA, *B = a.B + (size_t)g * a.gB;
    const int m0 = blockIdx.y * kBigM, n0 = blockIdx.x * kBigN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;  // thread tile: rows ty*8.., cols tx*4..
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    // staging: A tile 128 x 16 = 512 float4 (two per thread), B tile 64 x 16 = 256 float4 (one per thread)
    const int am[2] = {tid >> 2, (tid + 256) >> 2}, ak = (tid & 3) * 4;
    const int bn = tid >> 2, bk = (tid & 3) * 4;
    float4 ra[2], rb;
    auto fetch = [&](int k0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int m = m0 + am[r];
            ra[r] = (m < a.M && k0 + ak < a.K) ? *reinterpret_cast<const float4 *>(A + (size_t)m * a.sam + k0 + ak)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const int n = n0 + bn;  // weights sit at arbitrary offsets of the flat parameter vector: scalar loads
        if (n < a.N && k0 + bk < a.K) {
            const float *w = B + (size_t)n * a.sbn + k0 + bk;
            rb = make_float4(w[0], w[1], w[2], w[3]);
        } else {
            rb = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < a.K; k0 += kBigK) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            As[ak + 0][am[r]] = ra[r].x; As[ak + 1][am[r]] = ra[r].y; As[ak + 2][am[r]] = ra[r].z; As[ak + 3][am[r]] = ra[r].w;
        }
        Bs[bk + 0][bn] = rb.x; Bs[bk + 1][bn] = rb.y; Bs[bk + 2][bn] = rb.z; Bs[bk + 3][bn] = rb.w;
        __syncthreads();
        if (k0 + kBigK < a.K) fetch(k0 + kBigK);
#pragma unroll
        for (int kk = 0; kk < kBigK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8 + 4]);
            const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *C = a.C + (size_t)g * a.gC;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= a.N) continue;
            float v = acc[i][j];
            if (a.bias) v = __fadd_rn(v, a.bias[(size_t)g * a.gBias + n]);
            if (a.relu) v = fmaxf(v, 0.0f);
            const size_t ci = (size_t)m * a.scm + (size_t)n * a.scn;
            if (a.mask_aux) v = (a.aux[(size_t)g * a.gAux + ci] > 0.0f) ? v : 0.0f;
            if (a.accumulate) v = __fadd_rn(C[ci], v);
            C[ci] = v;
        }
    }
}

__global__ void splitk_reduce_kernel(const GemmArgs a, int ksplit, int G, const float *part) {
    const int g = blockIdx.y;
    const long long total = (long long)a.M * a.N;
    float *C = a.C + (size_t)g * a.gC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / a.N), n = (int)(i - (long long)m * a.N);
        float v = 0.0f;
        int s = 0;
        for (; s + 8 <= ksplit; s += 8) {  // eight independent loads in flight, summed in slice order
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = part[((size_t)(s + j) * G + g) * total + i];
#pragma unroll
            for (int j = 0; j < 8; ++j) v = __fadd_rn(v, t[j]);
        }
        for (; s < ksplit; ++s) v = __fadd_rn(v, part[((size_t)s * G + g) * total + i]);
        if (a.bias) v = __fadd_rn(v, a.bias[(size_t)g * a.gBias + n]);
        if (a.relu) v = fmaxf(v, 0.0f);
        const size_t ci = (size_t)m * a.scm + (size_t)n * a.scn;
        if (a.mask_aux) v = (a.aux[(size_t)g * a.gAux + ci] > 0.0f) ? v : 0.0f;
        if (a.accumulate) v = __fadd_rn(C[ci], v);
        C[ci] = v;
    }
}

// column sums: out[g][n] = sum_m X[g][m,n]  (bias gradients), one warp per column, fixed order
__global__ void colsum_kernel(const float *X, float *out, int M, int N, long long gX, long long gOut) {
    const int g = blockIdx.y, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= N) return;
    const float *x = X + (size_t)g * gX + warp;
    float s = 0.0f;
    for (int m = lane; m < M; m += 32) s = __fadd_rn(s, x[(size_t)m * N]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = __fadd_rn(s, __shfl_down_sync(0xFFFFFFFFu, s, o));
    if (lane == 0) out[(size_t)g * gOut + warp] = s;
}

// im2col for k=2, s=1: col[g][(b,oy,ox)][c*4 + ky*2 + kx] = src[g][b*sb + c*sc + (oy+ky)*sy + (ox+kx)*sx]
__global__ void im2col_kernel(const float *src, float *col, int B, int Cn, int Ho, int Wo, long long sb, long long sc,
                              long long sy, long long sx, long long gSrc, long long gCol) {
    const int g = blockIdx.y;
    const long long total = (long long)B * Ho * Wo * Cn * 4;
    if (total < (1ll << 31)) {  // 32-bit index arithmetic (the 64-bit divisions below cost ~10x: 24 -> 6 us for conv2's im2col at B = 256)
        const unsigned K = (unsigned)Cn * 4u, HW = (unsigned)Ho * (unsigned)Wo, n = (unsigned)total;
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const unsigned row = i / K, kk = i - row * K, b = row / HW, pos = row - b * HW, oy = pos / (unsigned)Wo, ox = pos - oy * (unsigned)Wo;
            const unsigned c = kk >> 2, ky = (kk >> 1) & 1u, kx = kk & 1u;
            col[(size_t)g * gCol + i] = src[(size_t)g * gSrc + (long long)b * sb + (long long)c * sc + (long long)(oy + ky) * sy + (long long)(ox + kx) * sx];
        }
        return;
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(i % (Cn * 4));
        const long long row = i / (Cn * 4);
        const int ox = (int)(row % Wo), oy = (int)((row / Wo) % Ho);
        const long long b = row / ((long long)Wo * Ho);
        const int c = kk >> 2, ky = (kk >> 1) & 1, kx = kk & 1;
        col[(size_t)g * gCol + i] = src[(size_t)g * gSrc + b * sb + c * sc + (oy + ky) * sy + (ox + kx) * sx];
    }
}

// F.max_pool2d(kernel 2, stride 2) on NHWC [B][H*W][F]; first maximum wins (strict >), index kept for backward
__global__ void pool_fwd_kernel(const float *y, float *p, uint8_t *arg, int B, int H, int W, int F, long long gY,
                                long long gP) {
    const int g = blockIdx.y, Hp = H / 2, Wp = W / 2;
    const long long total = (long long)B * Hp * Wp * F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i % F);
        const long long cell = i / F;
        const int px = (int)(cell % Wp), py = (int)((cell / Wp) % Hp);
        const long long b = cell / ((long long)Wp * Hp);
        const float *src = y + (size_t)g * gY + (size_t)b * H * W * F + f;
        float best = src[((size_t)(2 * py) * W + 2 * px) * F];
        int bi = 0;
#pragma unroll
        for (int d = 1; d < 4; ++d) {
            const float v = src[((size_t)(2 * py + (d >> 1)) * W + 2 * px + (d & 1)) * F];
            if (v > best) { best = v; bi = d; }
        }
        p[(size_t)g * gP + i] = best;
        arg[(size_t)g * gP + i] = (uint8_t)bi;
    }
}

__global__ void pool_bwd_kernel(const float *dp, const uint8_t *arg, float *dy, int B, int H, int W, int F, long long gY,
                                long long gP) {
    const int g = blockIdx.y, Hp = H / 2, Wp = W / 2;
    const long long total = (long long)B * H * W * F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i % F);
        const long long cell = i / F;
        const int x = (int)(cell % W), yy = (int)((cell / W) % H);
        const long long b = cell / ((long long)W * H);
        float v = 0.0f;
        const int py = yy >> 1, px = x >> 1;
        if (py < Hp && px < Wp) {
            const size_t pi = (size_t)g * gP + (((size_t)b * Hp + py) * Wp + px) * F + f;
            if (arg[pi] == (uint8_t)(((yy & 1) << 1) | (x & 1))) v = dp[pi];
        }
        dy[(size_t)g * gY + i] = v;
    }
}

// col2im (gather form) for k=2, s=1 on NHWC: dp[b][y][x][o] = sum over (ky,kx) of dcol[(b,y-ky,x-kx)][o*4+ky*2+kx]
__global__ void col2im_kernel(const float *dcol, float *dp, int B, int Hp, int Wp, int F, long long gCol, long long gP) {
    const int g = blockIdx.y, Ho = Hp - 1, Wo = Wp - 1;
    const long long total = (long long)B * Hp * Wp * F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(i % F);
        const long long cell = i / F;
        const int x = (int)(cell % Wp), y = (int)((cell / Wp) % Hp);
        const long long b = cell / ((long long)Wp * Hp);
        float s = 0.0f;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int ky = d >> 1, kx = d & 1, qy = y - ky, qx = x - kx;
            if (qy >= 0 && qy < Ho && qx >= 0 && qx < Wo)
                s = __fadd_rn(s, dcol[(size_t)g * gCol + (((size_t)b * Ho + qy) * Wo + qx) * (F * 4) + o * 4 + d]);
        }
        dp[(size_t)g * gP + i] = s;
    }
}

// NHWC [B][Q][F] <-> flattened NCHW [B][F*Q]  (out.view(B, -1) of conv_mlp.py:44)
__global__ void nhwc_to_flat_kernel(const float *y, float *flat, int B, int Q, int F, long long gs, int inverse) {
    const int g = blockIdx.y;
    const long long total = (long long)B * Q * F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i % F);
        const int q = (int)((i / F) % Q);
        const long long b = i / ((long long)F * Q);
        const size_t fi = (size_t)g * gs + (size_t)b * F * Q + (size_t)f * Q + q;
        if (!inverse) flat[fi] = y[(size_t)g * gs + i];
        else flat[(size_t)g * gs + i] = y[fi];  // here `y` is the flat gradient, `flat` the NHWC gradient
    }
}

// NHWC [B][Q][F] <-> a slice of a wider feature row: flat[b*row + off + f*Q + q]  (concatenated embeddings of the
// merged models, conv_mlp.py:87-98)
__global__ void nhwc_flat_strided_kernel(const float *y, float *flat, int B, int Q, int F, long long gY, long long row,
                                         long long off, long long gFlat, int inverse) {
    const int g = blockIdx.y;
    const long long total = (long long)B * Q * F;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i % F);
        const int q = (int)((i / F) % Q);
        const long long b = i / ((long long)F * Q);
        const size_t fi = (size_t)g * gFlat + (size_t)b * row + off + (size_t)f * Q + q;
        if (!inverse) flat[fi] = y[(size_t)g * gY + i];
        else flat[(size_t)g * gY + i] = y[fi];  // `y` is the gradient of the wide row, `flat` the NHWC gradient
    }
}

// ------------------------------------------------------------------------------ losses
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}

// fixed-order mean over the batch of per-sample losses: one block per group
__global__ void mean_kernel(const float *per_sample, float *loss, int B) {
    __shared__ float part[256];
    const int g = blockIdx.x;
    float s = 0.0f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) s = __fadd_rn(s, per_sample[(size_t)g * B + b]);
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) part[threadIdx.x] = __fadd_rn(part[threadIdx.x], part[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[g] = __fdiv_rn(part[0], (float)B);
}

// DQN TD target + MSE (da3_brain.py:69-87): one thread per (group, sample)
__global__ void dqn_loss_kernel(const float *q, const float *qn, const int32_t *act, const float *rew, const uint8_t *done,
                                float gamma, int B, int A, float *dq, float *per_sample) {
    const int g = blockIdx.y, b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const size_t i = (size_t)g * B + b;
    const float *qs = q + i * A, *qt = qn + i * A;
    float v = qt[0];
    for (int a = 1; a < A; ++a) v = fmaxf(v, qt[a]);
    const bool d = done[i] != 0;
    if (d) v = 0.0f;
    const float y = __fadd_rn(rew[i], __fmul_rn(__fmul_rn(gamma, d ? 0.0f : 1.0f), v));
    const int a_taken = act[i];
    const float diff = __fadd_rn(qs[a_taken], -y);
    per_sample[i] = __fmul_rn(diff, diff);
    for (int a = 0; a < A; ++a) dq[i * A + a] = (a == a_taken) ? __fdiv_rn(__fmul_rn(2.0f, diff), (float)B) : 0.0f;
}

// Distributional heads share the dueling combination  out[a][j] = V[j] + A[a][j] - mean_a A[a][j]
// (categorical_dqn.py:60-63, qr_dqn.py:46-53).  One WARP per (group, sample); lanes stride over atoms.
constexpr int kMaxAtomsPerLane = 4;  // up to 128 atoms / quantiles
constexpr int kMaxActions = 8;

struct DistArgs {
    const float *v_s, *a_s;    // online(s):   V [G][B][Z], A [G][B][A*Z]
    const float *v_n, *a_n;    // online(s'):  double-DQN action selection
    const float *v_t, *a_t;    // target(s')
    const int32_t *act;
    const float *rew;
    const float *support;      // C51: z_j ; QR: tau_hat_j
    float *dv, *da;            // gradients w.r.t. raw head outputs of online(s)
    float *per_sample;
    int B, A, Z;
    float gamma, v_min, v_max, delta_z;
};

__device__ __forceinline__ float dueling(const float *v, const float *adv, int A, int Z, int a, int j) {
    if (!v) return adv[a * Z + j];  // stand-alone loss entry points: adv holds the network's combined output
    float s = 0.0f;
    for (int k = 0; k < A; ++k) s = __fadd_rn(s, adv[k * Z + j]);
    const float mean = __fdiv_rn(s, (float)A);
    return __fadd_rn(v[j], __fadd_rn(adv[a * Z + j], -mean));
}

// C51 (categorical_dqn_brain.py:51-87).  One warp per (agent, sample); a lane owns atoms lane, lane + 32, ...  The dueling
// rows of a network are formed once per atom (the same operation order as dueling()), the projection scatters into a
// per-warp shared-memory row, and every head output is read exactly once.
// NA = most actions, NT = atoms per lane the instantiation is unrolled for: every warp runs the code once, so the kernel
// was bound by instruction fetch (ncu: "no instruction" 4.4 of 15 stall cycles per issue, 1 724 instructions per warp of which
// three quarters were predicated off at 4 actions x 51 atoms, 16.5 us); <4, 2> covers the reference's configurations.  All
// three networks' rows are loaded up front (one round of global latency instead of three).
template <int NA, int NT>
__global__ void __launch_bounds__(256) c51_loss_kernel(const DistArgs p, float *m_scratch) {
    __shared__ float s_m[8][32 * NT];
    const int g = blockIdx.y, wid = threadIdx.x >> 5, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= p.B) return;
    (void)m_scratch;
    const int A = p.A, Z = p.Z;
    const size_t i = (size_t)g * p.B + warp;
    const bool duel = p.v_s != nullptr;
    const float *vs = duel ? p.v_s + i * Z : nullptr, *as = p.a_s + i * A * Z;
    const float *vn = duel ? p.v_n + i * Z : nullptr, *an = p.a_n + i * A * Z;
    const float *vt = duel ? p.v_t + i * Z : nullptr, *at = p.a_t + i * A * Z;
    const int a_taken = p.act[i];
    const float r = p.rew[i];
    // all action rows of one network for this lane's atoms: v + (adv[a] - mean_a adv), sum over actions in index order
    auto rows = [&](const float *v, const float *adv, float (&d)[NA][NT]) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int jj = lane + 32 * t;
#pragma unroll
            for (int k = 0; k < NA; ++k) d[k][t] = 0.0f;
            if (jj >= Z) continue;
            float ad[NA], sum = 0.0f;
#pragma unroll
            for (int k = 0; k < NA; ++k)
                if (k < A) { ad[k] = adv[k * Z + jj]; sum = __fadd_rn(sum, ad[k]); }
            if (!v) {
#pragma unroll
                for (int k = 0; k < NA; ++k)
                    if (k < A) d[k][t] = ad[k];
            } else {
                const float mean = __fdiv_rn(sum, (float)A), vj = v[jj];
#pragma unroll
                for (int k = 0; k < NA; ++k)
                    if (k < A) d[k][t] = __fadd_rn(vj, __fadd_rn(ad[k], -mean));
            }
        }
    };
    float sup[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) sup[t] = (lane + 32 * t < Z) ? p.support[lane + 32 * t] : 0.0f;

    float dn[NA][NT], dt[NA][NT], ds[NA][NT];
    rows(vn, an, dn);
    rows(vt, at, dt);
    rows(vs, as, ds);
    // best action of the ONLINE net on s': argmax_a sum_j softmax(out[a])_j z_j, first maximum wins
    int best = 0;
    float best_q = 0.0f;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
        if (a >= A) break;
        float mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < NT; ++t)
            if (lane + 32 * t < Z) mx = fmaxf(mx, dn[a][t]);
        mx = warp_max(mx);
        float se = 0.0f, sq = 0.0f;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (lane + 32 * t < Z) {
                const float e = expf(__fadd_rn(dn[a][t], -mx));
                se = __fadd_rn(se, e);
                sq = __fadd_rn(sq, __fmul_rn(e, sup[t]));
            }
        }
        se = warp_sum(se);
        sq = warp_sum(sq);
        const float qv = __fdiv_rn(sq, se);
        if (a == 0 || qv > best_q) { best_q = qv; best = a; }
    }
    // p' = softmax(target(s')[best])
    float pt[NT], mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int jj = lane + 32 * t;
        float row = dt[0][t];
#pragma unroll
        for (int k = 1; k < NA; ++k) row = (k == best) ? dt[k][t] : row;
        pt[t] = jj < Z ? row : -INFINITY;
        mx = fmaxf(mx, pt[t]);
    }
    mx = warp_max(mx);
    float se = 0.0f;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        pt[t] = (lane + 32 * t < Z) ? expf(__fadd_rn(pt[t], -mx)) : 0.0f;
        se = __fadd_rn(se, pt[t]);
    }
    se = warp_sum(se);
    // projection onto the support: two scatter-adds per atom (index_add_, :80-81) into this warp's shared-memory row
    float *m = s_m[wid];
    for (int jj = lane; jj < Z; jj += 32) m[jj] = 0.0f;
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int jj = lane + 32 * t;
        if (jj >= Z) continue;
        const float prob = __fdiv_rn(pt[t], se);
        float tz = __fadd_rn(r, __fmul_rn(p.gamma, sup[t]));
        tz = fminf(fmaxf(tz, p.v_min), p.v_max);
        const float bb = __fdiv_rn(__fadd_rn(tz, -p.v_min), p.delta_z);
        int lo = (int)floorf(bb), up = (int)ceilf(bb);
        if (lo == up && lo > 0) lo -= 1;            // :73
        if (lo == up && up < Z - 1) up += 1;        // :75
        atomicAdd(&m[lo], __fmul_rn(prob, __fadd_rn((float)up, -bb)));
        atomicAdd(&m[up], __fmul_rn(prob, __fadd_rn(bb, -(float)lo)));
    }
    __syncwarp();
    // log_softmax(online(s)[a_taken]), loss and its gradient
    float l[NT];
    mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int jj = lane + 32 * t;
        float row = ds[0][t];
#pragma unroll
        for (int k = 1; k < NA; ++k) row = (k == a_taken) ? ds[k][t] : row;
        l[t] = jj < Z ? row : -INFINITY;
        mx = fmaxf(mx, l[t]);
    }
    mx = warp_max(mx);
    float sum_e = 0.0f;
#pragma unroll
    for (int t = 0; t < NT; ++t)
        if (lane + 32 * t < Z) sum_e = __fadd_rn(sum_e, expf(__fadd_rn(l[t], -mx)));
    sum_e = warp_sum(sum_e);
    const float lse = __fadd_rn(mx, logf(sum_e));
    float loss = 0.0f, msum = 0.0f;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int jj = lane + 32 * t;
        if (jj < Z) {
            loss = __fadd_rn(loss, __fmul_rn(m[jj], __fadd_rn(l[t], -lse)));
            msum = __fadd_rn(msum, m[jj]);
        }
    }
    loss = warp_sum(loss);
    msum = warp_sum(msum);
    if (lane == 0) p.per_sample[i] = -loss;
    const float invB = __fdiv_rn(1.0f, (float)p.B);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int jj = lane + 32 * t;
        if (jj >= Z) continue;
        const float sm = expf(__fadd_rn(l[t], -lse));
        const float gl = __fmul_rn(__fadd_rn(__fmul_rn(sm, msum), -m[jj]), invB);  // d loss / d out[a_taken][j]
        if (duel) p.dv[i * Z + jj] = gl;
        const float shared = duel ? __fdiv_rn(gl, (float)A) : 0.0f;
        for (int k = 0; k < A; ++k)
            p.da[i * A * Z + k * Z + jj] = __fadd_rn(k == a_taken ? gl : 0.0f, -shared);
    }
}

// QR-DQN (qr_dqn_brain.py:23-51,91-117): element-wise quantile Huber, kappa = 1
__global__ void qr_loss_kernel(const DistArgs p) {
    const int g = blockIdx.y, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= p.B) return;
    const int A = p.A, Z = p.Z;
    const size_t i = (size_t)g * p.B + warp;
    const bool duel = p.v_s != nullptr;
    const float *vs = duel ? p.v_s + i * Z : nullptr, *as = p.a_s + i * A * Z;
    const float *vn = duel ? p.v_n + i * Z : nullptr, *an = p.a_n + i * A * Z;
    const float *vt = duel ? p.v_t + i * Z : nullptr, *at = p.a_t + i * A * Z;
    const int a_taken = p.act[i];
    int best = 0;
    float best_q = 0.0f;
    for (int a = 0; a < A; ++a) {  // argmax_a mean_j quantiles(s')[a][j] of the online net
        float s = 0.0f;
        for (int j = lane; j < Z; j += 32) s = __fadd_rn(s, dueling(vn, an, A, Z, a, j));
        s = __fdiv_rn(warp_sum(s), (float)Z);
        if (a == 0 || s > best_q) { best_q = s; best = a; }
    }
    const float r = p.rew[i], invB = __fdiv_rn(1.0f, (float)p.B);
    float loss = 0.0f;
    for (int j = lane; j < Z; j += 32) {
        const float target = __fadd_rn(r, __fmul_rn(p.gamma, dueling(vt, at, A, Z, best, j)));
        const float theta = dueling(vs, as, A, Z, a_taken, j);
        const float delta = __fadd_rn(target, -theta);
        const float ad = fabsf(delta);
        const float huber = ad <= 1.0f ? __fmul_rn(0.5f, __fmul_rn(delta, delta)) : __fadd_rn(ad, -0.5f);
        const float w = fabsf(__fadd_rn(p.support[j], delta < 0.0f ? -1.0f : 0.0f));
        loss = __fadd_rn(loss, __fmul_rn(w, huber));
        const float dh = ad <= 1.0f ? delta : (delta > 0.0f ? 1.0f : -1.0f);
        const float gl = __fmul_rn(-__fmul_rn(w, dh), invB);  // d loss / d theta_j
        if (duel) p.dv[i * Z + j] = gl;
        const float shared = duel ? __fdiv_rn(gl, (float)A) : 0.0f;
        for (int k = 0; k < A; ++k)
            p.da[i * A * Z + k * Z + j] = __fadd_rn(k == a_taken ? gl : 0.0f, -shared);
    }
    loss = warp_sum(loss);
    if (lane == 0) p.per_sample[i] = loss;
}

// greedy-policy values Q[g][b][a] from raw head outputs (get_action of the brains)
__global__ void dist_q_kernel(const float *v, const float *adv, const float *support, int B, int Bmax, int A, int Z,
                              int categorical, float *q) {
    const int g = blockIdx.y, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const size_t i = (size_t)g * Bmax + warp;  // head outputs are strided by the maximum batch
    const float *vrow = v + i * Z, *arow = adv + i * A * Z;
    // each lane keeps its atoms of every action: v + (adv - mean over actions), computed once (same operation order as dueling())
    float d[kMaxActions][kMaxAtomsPerLane];
#pragma unroll
    for (int t = 0; t < kMaxAtomsPerLane; ++t) {
        const int j = lane + 32 * t;
        if (j >= Z) continue;
        float ad[kMaxActions], s = 0.0f;
#pragma unroll
        for (int k = 0; k < kMaxActions; ++k)
            if (k < A) { ad[k] = arow[k * Z + j]; s = __fadd_rn(s, ad[k]); }
        const float mean = __fdiv_rn(s, (float)A), vj = vrow[j];
#pragma unroll
        for (int k = 0; k < kMaxActions; ++k)
            if (k < A) d[k][t] = __fadd_rn(vj, __fadd_rn(ad[k], -mean));
    }
#pragma unroll
    for (int a = 0; a < kMaxActions; ++a) {
        if (a >= A) break;
        float out;
        if (categorical) {
            float mx = -INFINITY;
#pragma unroll
            for (int t = 0; t < kMaxAtomsPerLane; ++t)
                if (lane + 32 * t < Z) mx = fmaxf(mx, d[a][t]);
            mx = warp_max(mx);
            float se = 0.0f, sq = 0.0f;
#pragma unroll
            for (int t = 0; t < kMaxAtomsPerLane; ++t) {
                const int j = lane + 32 * t;
                if (j >= Z) continue;
                const float e = expf(__fadd_rn(d[a][t], -mx));
                se = __fadd_rn(se, e);
                sq = __fadd_rn(sq, __fmul_rn(e, support[j]));
            }
            out = __fdiv_rn(warp_sum(sq), warp_sum(se));
        } else {
            float s = 0.0f;
#pragma unroll
            for (int t = 0; t < kMaxAtomsPerLane; ++t)
                if (lane + 32 * t < Z) s = __fadd_rn(s, d[a][t]);
            out = __fdiv_rn(warp_sum(s), (float)Z);
        }
        if (lane == 0) q[((size_t)g * B + warp) * A + a] = out;
    }
}


// ------------------------------------------------------------------------------ IQN pieces
// CosineEmbeddingNetwork input (customs/iqn.py:32-40): cos(tau * (pi * k)), k = 1..ncos, fp32 products as torch forms them
__global__ void iqn_cos_kernel(const float *taus, float *cosf_out, int B, int Nq, int ncos, int per_sample, long long gT,
                               long long gC) {
    const int g = blockIdx.y;
    const long long total = (long long)B * Nq * ncos;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % ncos);
        const long long r = i / ncos;
        const int q = (int)(r % Nq);
        const long long b = r / Nq;
        const float tau = per_sample ? taus[(size_t)g * gT + b * Nq + q] : taus[(size_t)g * gT + q];
        const float ipi = __fmul_rn(3.14159274101257324f, (float)(k + 1));
        cosf_out[(size_t)g * gC + i] = cosf(__fmul_rn(tau, ipi));
    }
}

// e[(b,q)][d] = psi[b][d] * phi[(b,q)][d]   (customs/iqn.py:123)
__global__ void iqn_hadamard_fwd_kernel(const float *psi, const float *phi, float *e, int B, int Nq, int D, long long gPsi,
                                        long long gRow) {
    const int g = blockIdx.y;
    const long long total = (long long)B * Nq * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const long long b = i / ((long long)D * Nq);
        e[(size_t)g * gRow + i] = __fmul_rn(psi[(size_t)g * gPsi + b * D + d], phi[(size_t)g * gRow + i]);
    }
}

// dphi = de * psi masked by relu (phi > 0); dpsi[b][d] = sum_q de[(b,q)][d] * phi[(b,q)][d]  (fixed order over q)
__global__ void iqn_hadamard_bwd_kernel(const float *de, const float *psi, const float *phi, float *dphi, float *dpsi, int B,
                                        int Nq, int D, long long gPsi, long long gRow) {
    const int g = blockIdx.y;
    const long long total = (long long)B * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const long long b = i / D;
        const float ps = psi[(size_t)g * gPsi + i];
        float acc = 0.0f;
        for (int q = 0; q < Nq; ++q) {
            const size_t r = (size_t)g * gRow + ((size_t)b * Nq + q) * D + d;
            const float ph = phi[r], g_e = de[r];
            acc = __fadd_rn(acc, __fmul_rn(g_e, ph));
            dphi[r] = ph > 0.0f ? __fmul_rn(g_e, ps) : 0.0f;
        }
        dpsi[(size_t)g * gPsi + i] = acc;
    }
}

struct IqnArgs {
    const float *v0, *a0;  // online(s, taus):        V [G][B*Nq], A [G][B*Nq][A]
    const float *v1, *a1;  // online(s', tau_sel):    action selection (iqn_brain.py:84-85, customs/iqn.py:88-94)
    const float *v2, *a2;  // target(s', tau_dashes): [G][B*Nd], [G][B*Nd][A]
    const float *taus;     // [G][B][Nq]
    const int32_t *act;
    const float *rew;
    float *dv, *da, *per_sample;
    int B, A, Nq, Nd;
    long long gV, gA, gV2, gA2, gT;  // group strides
    float gamma;
};

__device__ __forceinline__ float iqn_quantile(const float *v, const float *adv, int A, int r, int a) {
    if (!v) return adv[(size_t)r * A + a];  // stand-alone entry point: adv holds get_quantiles' output [B, N, A]
    float s = 0.0f;
    for (int k = 0; k < A; ++k) s = __fadd_rn(s, adv[(size_t)r * A + k]);
    return __fadd_rn(v[r], __fadd_rn(adv[(size_t)r * A + a], -__fdiv_rn(s, (float)A)));  // customs/iqn.py:133-134
}

// IQNBrain.learn's loss (iqn_brain.py:23-50,73-113): pairwise quantile Huber, one warp per (group, sample)
__global__ void iqn_loss_kernel(const IqnArgs p) {
    const int g = blockIdx.y, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= p.B) return;
    const int A = p.A, Nq = p.Nq, Nd = p.Nd, b = warp;
    const bool duel = p.v0 != nullptr;
    const float *v0 = duel ? p.v0 + (size_t)g * p.gV : nullptr, *a0 = p.a0 + (size_t)g * p.gA;
    const float *v1 = duel ? p.v1 + (size_t)g * p.gV : nullptr, *a1 = p.a1 + (size_t)g * p.gA;
    const float *v2 = duel ? p.v2 + (size_t)g * p.gV2 : nullptr, *a2 = p.a2 + (size_t)g * p.gA2;
    const size_t sb = (size_t)g * p.B + b;
    const int a_taken = p.act[sb];
    int best = 0;
    float best_q = 0.0f;
    for (int a = 0; a < A; ++a) {  // q_values = quantiles.mean(dim=1); argmax, first maximum
        float s = 0.0f;
        for (int i = lane; i < Nq; i += 32) s = __fadd_rn(s, iqn_quantile(v1, a1, A, b * Nq + i, a));
        s = __fdiv_rn(warp_sum(s), (float)Nq);
        if (a == 0 || s > best_q) { best_q = s; best = a; }
    }
    const float r = p.rew[sb];
    const float scale = __fdiv_rn(1.0f, __fmul_rn((float)Nd, (float)p.B));
    float loss = 0.0f;
    for (int i = lane; i < Nq; i += 32) {
        const float theta = iqn_quantile(v0, a0, A, b * Nq + i, a_taken);
        const float tau = p.taus[(size_t)g * p.gT + (size_t)b * Nq + i];
        float grad = 0.0f;
        for (int j = 0; j < Nd; ++j) {
            const float target = __fadd_rn(r, __fmul_rn(p.gamma, iqn_quantile(v2, a2, A, b * Nd + j, best)));
            const float delta = __fadd_rn(target, -theta);
            const float ad = fabsf(delta);
            const float hub = ad <= 1.0f ? __fmul_rn(0.5f, __fmul_rn(delta, delta)) : __fadd_rn(ad, -0.5f);
            const float w = fabsf(__fadd_rn(tau, delta < 0.0f ? -1.0f : 0.0f));
            loss = __fadd_rn(loss, __fmul_rn(w, hub));
            const float dh = ad <= 1.0f ? delta : (delta > 0.0f ? 1.0f : -1.0f);
            grad = __fadd_rn(grad, -__fmul_rn(w, dh));
        }
        const float gl = __fmul_rn(grad, scale);  // d loss / d theta_i
        const size_t row = (size_t)b * Nq + i;
        if (duel) p.dv[(size_t)g * p.gV + row] = gl;
        const float shared = duel ? __fdiv_rn(gl, (float)A) : 0.0f;
        for (int k = 0; k < A; ++k)
            p.da[(size_t)g * p.gA + row * A + k] = __fadd_rn(k == a_taken ? gl : 0.0f, -shared);
    }
    loss = warp_sum(loss);
    if (lane == 0) p.per_sample[sb] = __fdiv_rn(loss, (float)Nd);  // .sum(dim=1).mean(dim=1)
}

// Q[g][b][a] = mean over quantiles (customs/iqn.py:138-139)
__global__ void iqn_q_kernel(const float *v, const float *adv, int B, int A, int Nq, long long gV, long long gA, float *q) {
    const int g = blockIdx.y, b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int a = 0; a < A; ++a) {
        float s = 0.0f;
        for (int i = 0; i < Nq; ++i) s = __fadd_rn(s, iqn_quantile(v + (size_t)g * gV, adv + (size_t)g * gA, A, b * Nq + i, a));
        q[((size_t)g * B + b) * A + a] = __fdiv_rn(s, (float)Nq);
    }
}


// ------------------------------------------------------------------------------ DA3 (ViT) pieces
// PatchEmbed (vit.py:34-47) is a stride-ps convolution: gather its im2col rows
// patches[g][(b,tx,ty)][(c,i,j)] = src[g][b*sb + c*sc + (tx*ps+i)*Win + (ty*ps+j)], tokens ordered tx*Wt + ty (flatten(2))
__global__ void vit_patch_gather_kernel(const float *src, float *patches, int B, int Cin, int Ht, int Wt, int Win, int ps,
                                        long long sb, long long sc, long long gSrc, long long gP) {
    const int g = blockIdx.y;
    const int K = Cin * ps * ps, nt = Ht * Wt;
    const long long total = (long long)B * nt * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        const int t = (int)((i / K) % nt);
        const long long b = i / ((long long)K * nt);
        const int c = k / (ps * ps), ij = k - c * ps * ps, pi = ij / ps, pj = ij - pi * ps;
        const int tx = t / Wt, ty = t - tx * Wt;
        patches[(size_t)g * gP + i] = src[(size_t)g * gSrc + b * sb + (long long)c * sc + (long long)(tx * ps + pi) * Win + (ty * ps + pj)];
    }
}

// x0[b][t] = (t == 0 ? token0 : tok[b][t-1]) + pos_embed[t]     (da3.py:77-83); token0 is the saliency_vector parameter,
// or, for DA6's local stage, a per-sample vector tok0_ext [B][E] (da6.py:113-116)
__global__ void da3_tokens_kernel(const float *tok, const float *sal, const float *tok0_ext, const float *pos, float *x0, int B,
                                  int Tn, int E, long long gTok, long long gX, long long gT0, long long P) {
    const int g = blockIdx.y;
    const long long total = (long long)B * Tn * E;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % E);
        const int t = (int)((i / E) % Tn);
        const long long b = i / ((long long)E * Tn);
        float base;
        if (t == 0) base = tok0_ext ? tok0_ext[(size_t)g * gT0 + b * E + d] : sal[(size_t)g * P + d];
        else base = tok[(size_t)g * gTok + (b * (Tn - 1) + (t - 1)) * E + d];
        x0[(size_t)g * gX + i] = __fadd_rn(base, pos[(size_t)g * P + (size_t)t * E + d]);
    }
}

// nn.LayerNorm(E), eps 1e-5, one warp per row; rows may be strided (final norm reads token 0 of every sample)
__global__ void layernorm_fwd_kernel(const float *x, long long x_row, const float *gamma, const float *beta, float *y,
                                     float *mean_out, float *rstd_out, int rows, int E, long long gX, long long gY,
                                     long long gS, long long P) {
    const int g = blockIdx.y, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const float *xr = x + (size_t)g * gX + (size_t)warp * x_row;
    float s = 0.0f;
    for (int d = lane; d < E; d += 32) s = __fadd_rn(s, xr[d]);
    const float mean = __fdiv_rn(warp_sum(s), (float)E);
    float v = 0.0f;
    for (int d = lane; d < E; d += 32) {
        const float c = __fadd_rn(xr[d], -mean);
        v = __fadd_rn(v, __fmul_rn(c, c));
    }
    const float var = __fdiv_rn(warp_sum(v), (float)E);
    const float rstd = __fdiv_rn(1.0f, sqrtf(__fadd_rn(var, 1e-5f)));
    for (int d = lane; d < E; d += 32) {
        const float xh = __fmul_rn(__fadd_rn(xr[d], -mean), rstd);
        y[(size_t)g * gY + (size_t)warp * E + d] = __fadd_rn(__fmul_rn(xh, gamma[(size_t)g * P + d]), beta[(size_t)g * P + d]);
    }
    if (lane == 0) {
        mean_out[(size_t)g * gS + warp] = mean;
        rstd_out[(size_t)g * gS + warp] = rstd;
    }
}

// dx (+)= LN backward; gx = dy * xhat (for d gamma = column sum of gx, d beta = column sum of dy)
__global__ void layernorm_bwd_kernel(const float *dy, const float *x, long long x_row, const float *gamma, const float *mean_in,
                                     const float *rstd_in, float *dx, long long dx_row, int accumulate, float *gx, int rows,
                                     int E, long long gDy, long long gX, long long gDx, long long gS, long long P) {
    const int g = blockIdx.y, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const float *xr = x + (size_t)g * gX + (size_t)warp * x_row, *dyr = dy + (size_t)g * gDy + (size_t)warp * E;
    const float mean = mean_in[(size_t)g * gS + warp], rstd = rstd_in[(size_t)g * gS + warp];
    float s1 = 0.0f, s2 = 0.0f;
    for (int d = lane; d < E; d += 32) {
        const float xh = __fmul_rn(__fadd_rn(xr[d], -mean), rstd);
        const float dg = __fmul_rn(dyr[d], gamma[(size_t)g * P + d]);
        s1 = __fadd_rn(s1, dg);
        s2 = __fadd_rn(s2, __fmul_rn(dg, xh));
        gx[(size_t)g * gDy + (size_t)warp * E + d] = __fmul_rn(dyr[d], xh);
    }
    s1 = __fdiv_rn(warp_sum(s1), (float)E);
    s2 = __fdiv_rn(warp_sum(s2), (float)E);
    float *dxr = dx + (size_t)g * gDx + (size_t)warp * dx_row;
    for (int d = lane; d < E; d += 32) {
        const float xh = __fmul_rn(__fadd_rn(xr[d], -mean), rstd);
        const float dg = __fmul_rn(dyr[d], gamma[(size_t)g * P + d]);
        const float v = __fmul_rn(rstd, __fadd_rn(__fadd_rn(dg, -s1), -__fmul_rn(xh, s2)));
        dxr[d] = accumulate ? __fadd_rn(dxr[d], v) : v;
    }
}

constexpr int kMaxTokens = 96, kMaxHeadDim = 32;

// Attention.forward (vit.py:116-132) for one (group, sample, head) per block; thread t owns query token t
// `shrink`: HardShrinkAttention (hard_shrink_attention.py:52-80) zeroes the softmax weights that are not above
// model.af_lambd before they multiply V; the plain Attention passes -1 (every weight is >= 0 and kept).  `probs`
// keeps the softmax output itself, which the backward pass needs.
__global__ void __launch_bounds__(128) da3_attn_fwd_kernel(const float *qkv, float *probs, float *out, int B, int Tn, int E,
                                                           int heads, float scale, long long gQ, long long gPr, long long gO,
                                                           float shrink) {
    __shared__ float sk[kMaxTokens][kMaxHeadDim + 1], sv[kMaxTokens][kMaxHeadDim + 1];
    const int g = blockIdx.z, b = blockIdx.y, h = blockIdx.x, dh = E / heads, t = threadIdx.x;
    const float *base = qkv + (size_t)g * gQ + (size_t)b * Tn * 3 * E;
    for (int i = threadIdx.x; i < Tn * dh; i += blockDim.x) {
        const int u = i / dh, d = i - u * dh;
        sk[u][d] = base[(size_t)u * 3 * E + E + h * dh + d];
        sv[u][d] = base[(size_t)u * 3 * E + 2 * E + h * dh + d];
    }
    __syncthreads();
    if (t >= Tn) return;
    float q[kMaxHeadDim];
    for (int d = 0; d < dh; ++d) q[d] = base[(size_t)t * 3 * E + h * dh + d];
    float *pr = probs + (size_t)g * gPr + (((size_t)b * heads + h) * Tn + t) * Tn;
    float mx = -INFINITY;
    for (int u = 0; u < Tn; ++u) {
        float s = 0.0f;
        for (int d = 0; d < dh; ++d) s = __fadd_rn(s, __fmul_rn(q[d], sk[u][d]));
        s = __fmul_rn(s, scale);
        pr[u] = s;
        mx = fmaxf(mx, s);
    }
    float se = 0.0f;
    for (int u = 0; u < Tn; ++u) {
        const float e = expf(__fadd_rn(pr[u], -mx));
        pr[u] = e;
        se = __fadd_rn(se, e);
    }
    float o[kMaxHeadDim];
    for (int d = 0; d < dh; ++d) o[d] = 0.0f;
    for (int u = 0; u < Tn; ++u) {
        const float pn = __fdiv_rn(pr[u], se);
        pr[u] = pn;
        const float pu = pn > shrink ? pn : 0.0f;
        for (int d = 0; d < dh; ++d) o[d] = __fadd_rn(o[d], __fmul_rn(pu, sv[u][d]));
    }
    float *orow = out + (size_t)g * gO + ((size_t)b * Tn + t) * E + h * dh;
    for (int d = 0; d < dh; ++d) orow[d] = o[d];
}

// The same attention for the large acting batches: identical arithmetic, but the score rows stay in shared memory instead
// of using the (then unneeded) probability buffer in HBM as scratch — 0.3 GB less traffic per 8192 observations.
__global__ void __launch_bounds__(96) da3_attn_infer_kernel(const float *qkv, float *out, int B, int Tn, int E, int heads,
                                                            float scale, long long gQ, long long gO, float shrink) {
    __shared__ float sk[kMaxTokens][kMaxHeadDim + 1], sv[kMaxTokens][kMaxHeadDim + 1];
    extern __shared__ float ss[];  // [Tn][Tn + 1] scores / probabilities, one row per thread
    const int g = blockIdx.z, b = blockIdx.y, h = blockIdx.x, dh = E / heads, t = threadIdx.x;
    const float *base = qkv + (size_t)g * gQ + (size_t)b * Tn * 3 * E;
    for (int i = threadIdx.x; i < Tn * dh; i += blockDim.x) {
        const int u = i / dh, d = i - u * dh;
        sk[u][d] = base[(size_t)u * 3 * E + E + h * dh + d];
        sv[u][d] = base[(size_t)u * 3 * E + 2 * E + h * dh + d];
    }
    __syncthreads();
    if (t >= Tn) return;
    float q[kMaxHeadDim];
    for (int d = 0; d < dh; ++d) q[d] = base[(size_t)t * 3 * E + h * dh + d];
    float *pr = ss + t * (Tn + 1);
    float mx = -INFINITY;
    for (int u = 0; u < Tn; ++u) {
        float sc = 0.0f;
        for (int d = 0; d < dh; ++d) sc = __fadd_rn(sc, __fmul_rn(q[d], sk[u][d]));
        sc = __fmul_rn(sc, scale);
        pr[u] = sc;
        mx = fmaxf(mx, sc);
    }
    float se = 0.0f;
    for (int u = 0; u < Tn; ++u) {
        const float e = expf(__fadd_rn(pr[u], -mx));
        pr[u] = e;
        se = __fadd_rn(se, e);
    }
    float o[kMaxHeadDim];
    for (int d = 0; d < dh; ++d) o[d] = 0.0f;
    for (int u = 0; u < Tn; ++u) {
        const float pn = __fdiv_rn(pr[u], se);
        const float pu = pn > shrink ? pn : 0.0f;
        for (int d = 0; d < dh; ++d) o[d] = __fadd_rn(o[d], __fmul_rn(pu, sv[u][d]));
    }
    float *orow = out + (size_t)g * gO + ((size_t)b * Tn + t) * E + h * dh;
    for (int d = 0; d < dh; ++d) orow[d] = o[d];
}

// backward of the attention core: d(out) -> d(qkv); two phases with a barrier, fixed summation order
__global__ void __launch_bounds__(128) da3_attn_bwd_kernel(const float *qkv, const float *probs, const float *dout,
                                                           float *dqkv, int B, int Tn, int E, int heads, float scale,
                                                           long long gQ, long long gPr, long long gO, float shrink) {
    extern __shared__ float smem_attn[];
    const int g = blockIdx.z, b = blockIdx.y, h = blockIdx.x, dh = E / heads, t = threadIdx.x;
    float *sq = smem_attn;                  // [Tn][dh]
    float *sk = sq + Tn * dh;               // [Tn][dh]
    float *sv = sk + Tn * dh;               // [Tn][dh]
    float *sdo = sv + Tn * dh;              // [Tn][dh]
    float *sds = sdo + Tn * dh;             // [Tn][Tn]  dS
    const float *base = qkv + (size_t)g * gQ + (size_t)b * Tn * 3 * E;
    const float *dob = dout + (size_t)g * gO + (size_t)b * Tn * E;
    for (int i = threadIdx.x; i < Tn * dh; i += blockDim.x) {
        const int u = i / dh, d = i - u * dh;
        sq[i] = base[(size_t)u * 3 * E + h * dh + d];
        sk[i] = base[(size_t)u * 3 * E + E + h * dh + d];
        sv[i] = base[(size_t)u * 3 * E + 2 * E + h * dh + d];
        sdo[i] = dob[(size_t)u * E + h * dh + d];
    }
    __syncthreads();
    const float *pr = probs + (size_t)g * gPr + (((size_t)b * heads + h) * Tn) * Tn;
    float *dq_out = dqkv + (size_t)g * gQ + (size_t)b * Tn * 3 * E;
    if (t < Tn) {
        // dP[t,u] = dO_t . v_u where P[t,u] > shrink, else 0 (Hardshrink passes the gradient of the kept weights only);
        // dS = P * (dP - sum_u dP*P) ; dq_t = scale * sum_u dS[t,u] k_u
        float dot = 0.0f;
        for (int u = 0; u < Tn; ++u) {
            float dp = 0.0f;
            for (int d = 0; d < dh; ++d) dp = __fadd_rn(dp, __fmul_rn(sdo[t * dh + d], sv[u * dh + d]));
            if (!(pr[t * Tn + u] > shrink)) dp = 0.0f;
            sds[t * Tn + u] = dp;
            dot = __fadd_rn(dot, __fmul_rn(dp, pr[t * Tn + u]));
        }
        float dq[kMaxHeadDim];
        for (int d = 0; d < dh; ++d) dq[d] = 0.0f;
        for (int u = 0; u < Tn; ++u) {
            const float ds = __fmul_rn(pr[t * Tn + u], __fadd_rn(sds[t * Tn + u], -dot));
            sds[t * Tn + u] = ds;
            for (int d = 0; d < dh; ++d) dq[d] = __fadd_rn(dq[d], __fmul_rn(ds, sk[u * dh + d]));
        }
        for (int d = 0; d < dh; ++d) dq_out[(size_t)t * 3 * E + h * dh + d] = __fmul_rn(dq[d], scale);
    }
    __syncthreads();
    if (t < Tn) {
        // key/value token u = t: dk_u = scale * sum_t' dS[t',u] q_t' ; dv_u = sum_t' P[t',u] dO_t'
        const int u = t;
        float dk[kMaxHeadDim], dv[kMaxHeadDim];
        for (int d = 0; d < dh; ++d) dk[d] = dv[d] = 0.0f;
        for (int tq = 0; tq < Tn; ++tq) {
            const float ds = sds[tq * Tn + u], pn = pr[tq * Tn + u], pu = pn > shrink ? pn : 0.0f;
            for (int d = 0; d < dh; ++d) {
                dk[d] = __fadd_rn(dk[d], __fmul_rn(ds, sq[tq * dh + d]));
                dv[d] = __fadd_rn(dv[d], __fmul_rn(pu, sdo[tq * dh + d]));
            }
        }
        for (int d = 0; d < dh; ++d) {
            dq_out[(size_t)u * 3 * E + E + h * dh + d] = __fmul_rn(dk[d], scale);
            dq_out[(size_t)u * 3 * E + 2 * E + h * dh + d] = dv[d];
        }
    }
}

// nn.Hardshrink on non-negative softmax weights, in place (attention-map export of the hard-shrink blocks)
__global__ void hard_shrink_kernel(float *p, long long n, float shrink) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(p[i] > shrink)) p[i] = 0.0f;
}

// nn.GELU() (exact erf form) forward / backward
__global__ void gelu_fwd_kernel(const float *x, float *y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i];
        y[i] = __fmul_rn(__fmul_rn(0.5f, v), __fadd_rn(1.0f, erff(__fmul_rn(v, 0.70710678118654752440f))));
    }
}
__global__ void gelu_bwd_kernel(const float *x, const float *dy, float *dx, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i];
        const float cdf = __fmul_rn(0.5f, __fadd_rn(1.0f, erff(__fmul_rn(v, 0.70710678118654752440f))));
        const float pdf = __fmul_rn(0.39894228040143267794f, expf(__fmul_rn(-0.5f, __fmul_rn(v, v))));
        dx[i] = __fmul_rn(dy[i], __fadd_rn(cdf, __fmul_rn(v, pdf)));
    }
}

// dx2[b][t] = (t == 0) ? d_token0[b] : 0   (only the saliency token feeds the head, da3.py:90-91)
__global__ void da3_scatter_token0_kernel(const float *dtok0, float *dx, int B, int Tn, int E, long long gT, long long gX) {
    const int g = blockIdx.y;
    const long long total = (long long)B * Tn * E;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % E);
        const int t = (int)((i / E) % Tn);
        const long long b = i / ((long long)E * Tn);
        dx[(size_t)g * gX + i] = t == 0 ? dtok0[(size_t)g * gT + b * E + d] : 0.0f;
    }
}

// rows t >= 1 of dx0 [B][Tn][E] -> dtok [B*(Tn-1)][E]
__global__ void da3_token_rows_kernel(const float *dx0, float *dtok, int B, int Tn, int E, long long gX, long long gTok) {
    const int g = blockIdx.y;
    const long long total = (long long)B * (Tn - 1) * E;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % E);
        const int t = (int)((i / E) % (Tn - 1));
        const long long b = i / ((long long)E * (Tn - 1));
        dtok[(size_t)g * gTok + i] = dx0[(size_t)g * gX + (b * Tn + t + 1) * E + d];
    }
}

// ------------------------------------------------------------------------------ optimizer
// stage 1: per-block partial sums of squares (fixed order); stage 2 inside the apply kernel
__global__ void sqnorm_partial_kernel(const float *grads, float *partial, long long P, float scale) {
    __shared__ float part[256];
    const int g = blockIdx.y;
    const float *x = grads + (size_t)g * P;
    float s = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        const float v = __fmul_rn(x[i], scale);
        s = __fadd_rn(s, __fmul_rn(v, v));
    }
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) part[threadIdx.x] = __fadd_rn(part[threadIdx.x], part[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)g * gridDim.x + blockIdx.x] = part[0];
}

// nn.utils.clip_grad_norm_(max_norm) + optim.RMSprop(lr, alpha, eps).step()  (optimizers/__init__.py:28-33)
// optim.Adam single-tensor update (torch/optim/adam.py _single_tensor_adam, no amsgrad / weight decay) after the same clip:
// exp_avg.lerp_(g, 1-b1); exp_avg_sq = b2*exp_avg_sq + (1-b2)*g*g; p -= (lr / (1-b1^t)) * exp_avg / (sqrt(exp_avg_sq)/sqrt(1-b2^t) + eps)
// with the scalar factors evaluated in double, as Python does.  steps[g] was incremented by adam_step_kernel.
__global__ void clip_adam_kernel(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, const float *partial,
                                 int nparts, long long P, float scale, float max_norm, float lr, float beta1, float beta2,
                                 float eps, const uint8_t *active, const int *steps, float *norm_out) {
    const int g = blockIdx.y;
    if (active && !active[g]) return;
    float total = 0.0f;
    for (int i = 0; i < nparts; ++i) total = __fadd_rn(total, partial[(size_t)g * nparts + i]);
    const float norm = sqrtf(total);
    const float coef = fminf(__fdiv_rn(max_norm, __fadd_rn(norm, 1e-6f)), 1.0f);
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) norm_out[g] = norm;
    const double t = (double)steps[g];
    const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
    const float step_size = (float)((double)lr / bc1), bc2_sqrt = (float)sqrt(bc2);
    const float w1 = (float)(1.0 - (double)beta1), w2 = (float)(1.0 - (double)beta2);
    float *p = params + (size_t)g * P, *m = exp_avg + (size_t)g * P, *v = exp_avg_sq + (size_t)g * P;
    const float *gr = grads + (size_t)g * P;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        const float gi = __fmul_rn(__fmul_rn(gr[i], scale), coef);
        const float mi = __fadd_rn(m[i], __fmul_rn(w1, __fadd_rn(gi, -m[i])));
        const float vi = __fadd_rn(__fmul_rn(v[i], beta2), __fmul_rn(__fmul_rn(w2, gi), gi));
        m[i] = mi;
        v[i] = vi;
        const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2_sqrt), eps);
        p[i] = __fadd_rn(p[i], -__fmul_rn(step_size, __fdiv_rn(mi, denom)));
    }
}

__global__ void adam_step_kernel(int *steps, const uint8_t *active, int G) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < G && (!active || active[g])) steps[g] += 1;
}

__global__ void clip_rmsprop_kernel(float *params, const float *grads, float *sq, const float *partial, int nparts,
                                    long long P, float scale, float max_norm, float lr, float alpha, float one_minus_alpha, float eps,
                                    const uint8_t *active, float *norm_out) {
    const int g = blockIdx.y;
    if (active && !active[g]) return;
    float total = 0.0f;
    for (int i = 0; i < nparts; ++i) total = __fadd_rn(total, partial[(size_t)g * nparts + i]);
    const float norm = sqrtf(total);
    const float coef = fminf(__fdiv_rn(max_norm, __fadd_rn(norm, 1e-6f)), 1.0f);
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) norm_out[g] = norm;
    float *p = params + (size_t)g * P, *s = sq + (size_t)g * P;
    const float *gr = grads + (size_t)g * P;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        const float gi = __fmul_rn(__fmul_rn(gr[i], scale), coef);
        const float v = __fadd_rn(__fmul_rn(s[i], alpha), __fmul_rn(__fmul_rn(one_minus_alpha, gi), gi));
        s[i] = v;
        const float avg = __fadd_rn(sqrtf(v), eps);
        p[i] = __fadd_rn(p[i], -__fmul_rn(lr, __fdiv_rn(gi, avg)));
    }
}

// Extra canvas channels behind the `base` decoded ones (customs/da3_iqn.py:195-211, da6.py:166-179, iqn.py:149-160,
// conv_mlp.py:56-64): objects [T,X,Y] from the snapshot's bit rows, one-hot agent cells [N,X,Y]
// (ObservationHandler.decode_agents_channel, observation_handler.py:218-250), the agent's destination area [1,X,Y].
__global__ void canvas_extra_kernel(float *canvas, long long rows, int base, int T, int Na, int Dc, int X, int Y,
                                    const uint32_t *obj, const int8_t *agents, const uint32_t *dest) {
    const int Cx = T + Na + Dc, Ct = base + Cx;
    const long long XY = (long long)X * Y, total = rows * Cx * XY;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i % Y), x = (int)((i / Y) % X), c = (int)((i / XY) % Cx);
        const long long row = i / (Cx * XY);
        float v;
        if (c < T) v = (float)((obj[(row * T + c) * X + x] >> y) & 1u);
        else if (c < T + Na) {
            const int8_t *a = agents + (row * Na + (c - T)) * 2;
            v = (a[0] == x && a[1] == y) ? 1.0f : 0.0f;
        } else v = (float)((dest[row * X + x] >> y) & 1u);
        canvas[(row * Ct + base + c) * XY + (long long)x * Y + y] = v;
    }
}

// ---------------------------------------------------------------------------------- data-parallel learner: fused collective
// One cooperative kernel per update replaces  NCCL all-reduce(sum) -> grads *= 1/world -> clip_grad_norm_ -> RMSprop  of the
// data-parallel learner.  Every rank's flat gradient buffer [G*P] lives in peer-mapped (symmetric) memory, so
// the exchange is plain 128-bit loads / stores over NVLink inside the kernel:
//   phase 1  reduce-scatter + all-gather: rank r sums slice r of all ranks' buffers in rank order, scales by 1/world and
//            writes the reduced slice back into EVERY rank's buffer; block partials of the per-agent sum of squares
//   phase 2  (after a grid sync) block 0 adds the block partials in a fixed order and publishes this rank's per-agent
//            partial norms to every rank; cross-GPU barrier on the signal pads
//   phase 3  every rank: norm_g = sqrt(sum of the world partials in rank order), clip coefficient, RMSprop on all G*P
//            parameters from the reduced gradients now sitting in its own buffer.
// All ranks apply bit-identical gradients and norms (each value is computed by exactly one rank), so replicas stay equal.
struct DpPeers {
    float *grads[8];        // peer-mapped gradient buffers, [G*P] floats each, index = rank
    float *partials[8];     // peer-mapped [world][G] partial sums of squares, index = rank
    uint32_t *signals[8];   // peer-mapped signal pads (>= 8 words used), index = rank
    float *mc_grads;        // NVLS multicast mapping of the gradient buffers (nullptr: peer loads / stores)
    int world, rank;
};

// in-switch reduction / broadcast over the NVLS multicast mapping (one request instead of `world` peer accesses)
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float *mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float *mc, const float4 &v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ void dp_barrier(const DpPeers &pe, uint32_t epoch) {
    // called by block 0 after a grid sync: thread t < world signals rank t and waits for rank t's signal
    if (threadIdx.x < pe.world) {
        __threadfence_system();
        const int t = threadIdx.x;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pe.signals[t] + pe.rank), "r"(epoch) : "memory");
        uint32_t seen = 0;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(pe.signals[pe.rank] + t) : "memory");
            if (clock64() - t0 > (1ll << 33)) __trap();  // a rank that never arrives must not wedge the GPU (~4 s)
        } while ((int32_t)(seen - epoch) < 0);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) dp_reduce_apply_kernel(const DpPeers pe, float *params, float *sq, long long P, int G,
                                                             uint32_t *epoch_counter, float *block_part, float max_norm, float lr,
                                                             float alpha, float one_minus_alpha, float eps, const uint8_t *active,
                                                             float *norm_out) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ float red[256];
    __shared__ uint32_t s_epoch;
    const long long N = (long long)G * P, N4 = N >> 2;  // P is a multiple of 4 floats
    const int W = pe.world, r = pe.rank;
    const float inv_world = __fdiv_rn(1.0f, (float)W);
    // every rank's gradients must be complete before anyone reads them (this rank's are, by stream order)
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) s_epoch = *epoch_counter;  // advanced at the end; two barriers per call
        __syncthreads();
        dp_barrier(pe, s_epoch + 1);
    }
    grid.sync();

    // ---- phase 1: my slice, reduced in rank order and broadcast
    const long long per = (N4 + W - 1) / W, lo = per * r, hi = min(N4, lo + per);
    float acc[8];  // sum of squares per agent touched by this thread (a slice spans at most a few agents)
    const int g_first = (int)((lo * 4) / P);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
    {
        const long long stride1 = (long long)gridDim.x * blockDim.x;
        auto reduce_one = [&](long long i, float4 (&u)[8]) {
            float4 v = u[0];
            for (int q = 1; q < W; ++q) {
                v.x = __fadd_rn(v.x, u[q].x); v.y = __fadd_rn(v.y, u[q].y); v.z = __fadd_rn(v.z, u[q].z); v.w = __fadd_rn(v.w, u[q].w);
            }
            v.x = __fmul_rn(v.x, inv_world); v.y = __fmul_rn(v.y, inv_world); v.z = __fmul_rn(v.z, inv_world); v.w = __fmul_rn(v.w, inv_world);
            for (int q = 0; q < W; ++q) *reinterpret_cast<float4 *>(pe.grads[q] + i * 4) = v;
            const int k = (int)((i * 4) / P) - g_first;  // a float4 never straddles agents (P % 4 == 0)
            const float ss = __fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fadd_rn(__fmul_rn(v.z, v.z), __fmul_rn(v.w, v.w)));
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j == k) acc[j] = __fadd_rn(acc[j], ss);
        };
        auto finish_one = [&](long long i, float4 v) {  // NVLS: v is already the sum over ranks
            v.x = __fmul_rn(v.x, inv_world); v.y = __fmul_rn(v.y, inv_world); v.z = __fmul_rn(v.z, inv_world); v.w = __fmul_rn(v.w, inv_world);
            multimem_st(pe.mc_grads + i * 4, v);
            const int k = (int)((i * 4) / P) - g_first;
            const float ss = __fadd_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)), __fadd_rn(__fmul_rn(v.z, v.z), __fmul_rn(v.w, v.w)));
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j == k) acc[j] = __fadd_rn(acc[j], ss);
        };
        if (pe.mc_grads) {
            for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += 2 * stride1) {
                const long long i2 = i + stride1;
                const float4 va = multimem_ld_reduce_add(pe.mc_grads + i * 4);
                float4 vb = va;
                if (i2 < hi) vb = multimem_ld_reduce_add(pe.mc_grads + i2 * 4);
                finish_one(i, va);
                if (i2 < hi) finish_one(i2, vb);
            }
        } else
        for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += 2 * stride1) {
            const long long i2 = i + stride1;
            float4 ua[8], ub[8];  // all peer loads of two elements in flight before the first add
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q < W) ua[q] = *reinterpret_cast<const float4 *>(pe.grads[q] + i * 4);
            if (i2 < hi) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (q < W) ub[q] = *reinterpret_cast<const float4 *>(pe.grads[q] + i2 * 4);
            }
            reduce_one(i, ua);
            if (i2 < hi) reduce_one(i2, ub);
        }
    }
    // block partials per agent slot (fixed tree order), then block 0 adds the blocks in order: deterministic
    for (int k = 0; k < 8; ++k) {
        red[threadIdx.x] = acc[k];
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) red[threadIdx.x] = __fadd_rn(red[threadIdx.x], red[threadIdx.x + o]);
            __syncthreads();
        }
        if (threadIdx.x == 0) block_part[(size_t)blockIdx.x * 8 + k] = red[0];
        __syncthreads();
    }
    __threadfence_system();  // the reduced slice must be visible on every GPU before the barrier below
    grid.sync();
    if (blockIdx.x == 0) {
        // warp k adds agent slot k over the blocks: lane-strided partial sums, then a fixed shuffle tree (deterministic)
        const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
        float part = 0.0f;
        for (unsigned b = lane; b < gridDim.x; b += 32) part = __fadd_rn(part, block_part[(size_t)b * 8 + k]);
        for (int o = 16; o > 0; o >>= 1) part = __fadd_rn(part, __shfl_down_sync(0xFFFFFFFFu, part, o));
        if (lane == 0) red[k] = part;
        __syncthreads();
        if (threadIdx.x < G) {
            const int g = threadIdx.x, slot = g - g_first;
            const float total = (slot >= 0 && slot < 8) ? red[slot] : 0.0f;  // 0 for agents outside my slice
            for (int q = 0; q < W; ++q) pe.partials[q][(size_t)r * G + g] = total;
        }
        __syncthreads();
        dp_barrier(pe, s_epoch + 2);
        if (threadIdx.x == 0) *epoch_counter = s_epoch + 2;
    }
    grid.sync();

    // ---- phase 3: clip + RMSprop on every parameter (clip_grad_norm_ after the reduce, as a single GPU on the global batch)
    __shared__ float s_coef[256];
    const float *gr = pe.grads[r];
    const float *part = pe.partials[r];
    if (threadIdx.x < G) {  // per-agent clip coefficient, once per block: world partials added in rank order
        const int g = threadIdx.x;
        float total = 0.0f;
        for (int q = 0; q < W; ++q) total = __fadd_rn(total, part[(size_t)q * G + g]);
        const float norm = sqrtf(total);
        s_coef[g] = (active && !active[g]) ? -1.0f : fminf(__fdiv_rn(max_norm, __fadd_rn(norm, 1e-6f)), 1.0f);
        if (norm_out && blockIdx.x == 0) norm_out[g] = norm;
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    auto update = [&](long long i, const float4 &g4, const float4 &s4, const float4 &p4) {
        const float coef = s_coef[(int)((i * 4) / P)];
        if (coef < 0.0f) return;  // inactive agent (agent_tasks "-1")
        const float gi[4] = {__fmul_rn(g4.x, coef), __fmul_rn(g4.y, coef), __fmul_rn(g4.z, coef), __fmul_rn(g4.w, coef)};
        float sv[4] = {s4.x, s4.y, s4.z, s4.w}, pv[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sv[j] = __fadd_rn(__fmul_rn(sv[j], alpha), __fmul_rn(__fmul_rn(one_minus_alpha, gi[j]), gi[j]));
            pv[j] = __fadd_rn(pv[j], -__fmul_rn(lr, __fdiv_rn(gi[j], __fadd_rn(sqrtf(sv[j]), eps))));
        }
        *reinterpret_cast<float4 *>(sq + i * 4) = make_float4(sv[0], sv[1], sv[2], sv[3]);
        *reinterpret_cast<float4 *>(params + i * 4) = make_float4(pv[0], pv[1], pv[2], pv[3]);
    };
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N4; i += 2 * stride) {  // two float4 in flight
        const long long i2 = i + stride;
        const float4 ga = *reinterpret_cast<const float4 *>(gr + i * 4), sa = *reinterpret_cast<const float4 *>(sq + i * 4),
                     pa = *reinterpret_cast<const float4 *>(params + i * 4);
        float4 gb = ga, sb = sa, pb = pa;
        if (i2 < N4) {
            gb = *reinterpret_cast<const float4 *>(gr + i2 * 4);
            sb = *reinterpret_cast<const float4 *>(sq + i2 * 4);
            pb = *reinterpret_cast<const float4 *>(params + i2 * 4);
        }
        update(i, ga, sa, pa);
        if (i2 < N4) update(i2, gb, sb, pb);
    }
}

int decode_relative_launch(const float *obs, const int8_t *positions, float *canvas, long long rows, int channels, int visible_range,
                           int size_x, int size_y, int obs_stride, int only_area_plus_one, long long canvas_stride,
                           const int8_t *los, cudaStream_t st);  // agents.cu

}  // namespace maexp

using namespace maexp;

// --------------------------------------------------------------------------------- handle
struct ParamEntry {
    std::string name;
    long long offset, count;
};
struct HeadLayout {
    long long w1, b1, w2, b2;
    int out;
};

struct maexp_learner {
    maexp_learner_cfg cfg;
    int G, B, C, V, A, H, F1, F2, Z, D;  // D = flattened conv output (or C*V*V for the plain MLP)
    int H1, Hp, H2;                      // conv1 out side, pooled side, conv2 out side
    bool has_conv, dist, iqn, da3, da6 = false;
    bool use_tc = true;  // tensor-core GEMMs for large inference batches (MAEXP_TC=0 turns them off)
    bool batched_fwd = false;      // all forward passes of an update as slices of one launch (MAEXP_BATCHED_FWD=1; measured slower than one stream each: 263 vs 234 us)
    bool fused_mlp = true;         // one-kernel MLP heads for the update batches (MAEXP_FUSED_MLP=0: GEMM by GEMM)
    // MAEXP_TC_WGRAD=1: the weight-gradient GEMMs of >= 128-row update batches on tcgen05 (tc_gemm_tn_kernel).  Opt-in: measured
    // SLOWER inside the update (184 against 175 us per C51 team update) although its kernels are ~5x shorter -- their 80-100 KB
    // of shared memory cannot sit beside the three resident CTAs per SM of trunk_fused_bwd_kernel, so they run after the
    // data-gradient chain instead of beside it (stamps: gradients joined 22 us after the chain instead of 1 us)
    bool tc_wgrad = false;
    // MAEXP_MLP_BWD_SPLIT=1: the heads' data-gradient chain as two kernels (dh once per head, then dx) instead of one whose
    // column tiles each repeat the dout W2 product (twice at din = 256).  Opt-in: measured slower inside the update (183 against
    // 175 us) -- one more launch on the chain, and the dh kernel's 128 CTAs walk the same 8 K steps
    bool mlp_bwd_split = false;
    bool tc_mlp = true;            // the MLP heads of >= 128-row update batches on tcgen05 (csrc/tc_mlp.cu; MAEXP_TC_MLP=0: the SIMT kernel).
                                   // First measured slower than the SIMT kernel (45 vs 33 us at B = 256; profiles/r2_tc_mlp_fwd_*.txt): its
                                   // operand staging went through GENERIC stores (the 1024-byte alignment of the dynamic shared memory was
                                   // computed through uintptr_t, which loses the address space).  With shared-space stores: C51 team update
                                   // 239 -> 224 us
    bool tc_heads = true;          // acting batches: both layers of all heads in one TMA-fed tcgen05 kernel (csrc/tc_heads.cu; MAEXP_TC_HEADS=0: GEMM by GEMM)
    bool tc_trunk = true;          // acting batches: conv2 of the trunk as an implicit GEMM on tcgen05 (csrc/tc_trunk.cu; MAEXP_TC_TRUNK=0: CUDA cores)
    float *tc_trunk_ws = nullptr;
    struct { const float *key; float *ws; } tc_trunk_upd[4] = {};  // the update's passes run concurrently: one panel buffer each
    float *tc_heads_ws = nullptr;  // its pre-split weight panels
    size_t tc_heads_ws_floats = 0;
    // the same kernel for the forward passes of an update (online(s), online(s'), target(s') run concurrently on their own
    // streams): one panel buffer per pass, keyed by the pass's output buffer
    struct { const float *key; float *ws; size_t floats; } tc_heads_upd[8] = {};
    bool fused_trunk = true;       // one-kernel conv trunk for every local-view pass (MAEXP_FUSED_TRUNK=0: layer by layer)
    const float *bwd_obs = nullptr;  // observations of the last training pass (conv1's im2col in trunk_backward)
    float *trunk_part = nullptr;      // block partials of trunk_fused_bwd_kernel
    bool col2_ready = false, dy2_ready = false;  // backward(): conv2's im2col is in flight on a side stream (ev_col2) / the
    cudaEvent_t ev_col2 = nullptr;               // fused MLP backward wrote d_y2 itself
    bool fused_trunk_acts = false; // the last training pass kept p / arg only: trunk_backward builds col1 / col2 itself
    int last_q_batch = 0;
    int E_, heads_, Fm, L_, Tn;   // DA3 / DA6: embed_dim, num_heads, mlp hidden, block_loop, max tokens of any stack
    float attn_shrink = -1.0f;    // model.af_lambd of HardShrinkAttention, -1: plain softmax attention
    long long o_nfw, o_nfb, o_hw, o_hb, o_f1w = 0, o_f1b = 0;  // shared final norm, drl_head, DA6 fc1
    struct BlockLayout { long long n1w, n1b, qkvw, projw, projb, n2w, n2b, f1w, f1b, f2w, f2b; };
    int Nq, Nd, ncos, rows_max;  // IQN: quantiles, target quantiles, cosines, B * max(Nq, Nd)
    HeadLayout se;               // IQN state_embedder (customs/iqn.py:51-53)
    long long o_cw, o_cb;        // IQN cosine_net.net.0
    long long P;                         // parameters per agent
    long long o_c1w, o_c1b, o_c2w, o_c2b;
    std::vector<HeadLayout> heads;
    std::vector<ParamEntry> table;
    std::vector<void *> allocs;
    float *online = nullptr, *target = nullptr, *grads = nullptr, *sq = nullptr, *exp_avg = nullptr;
    int *adam_steps = nullptr;
    float *support = nullptr, *ones = nullptr, *splitk_ws = nullptr;
    // the three forward passes of an update are independent: passes 1 and 2 run on side streams (fork / join with
    // events, which also works under stream capture) with their own split-K workspaces
    float *splitk_side[kSides] = {};
    cudaStream_t side[kSides] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[kSides] = {}, ev_w[kSides] = {};
    cudaEvent_t ev_mark[3] = {}, ev_prep[3] = {};
    float *mean_dst = nullptr;  // where backward() writes the batch-mean loss (set by the update just before it)
    bool zero_live = false, mlp_bwd_head_par = true;  // ev_fork of THIS update exists until its backward pass is issued
    cudaEvent_t ev_zero = nullptr;
    bool fork_live = false, prep_branch = true;  // ev_fork was recorded in this update; weight-panel kernels beside the trunk   // backward(): points of the main stream the deferred weight-gradient work waits for
    bool critical_first = true;    // MAEXP_CRITICAL_FIRST=0: weight-gradient kernels are issued where their inputs become ready
    float *d_h2 = nullptr, *d_hrow2 = nullptr;  // second hidden-gradient buffers: weight-gradient GEMMs of head k read
                                                // theirs on a side stream while head k+1's data gradients are computed
    // activations of the three passes (0: online(s) kept for backward, 1: online(s'), 2: target(s'))
    struct Acts {
        float *col1, *y1, *p, *col2, *y2, *flat, *h[2], *out[2];
        uint8_t *arg;
    } acts[3];
    struct IqnActs {
        float *se_h, *psi, *cosf, *phi, *e, *hrow[2], *out[2];
    } iq[3];
    float *d_e = nullptr, *d_phi = nullptr, *d_psi = nullptr, *d_hrow = nullptr;
    struct Da3Block { float *n1, *m1, *r1, *qkv, *probs, *ao, *x1, *n2, *m2, *r2, *f1, *g1, *x2; };
    struct Da3Acts { float *patches, *tok, *x0, *nf, *mf, *rf, *q; std::vector<Da3Block> blk; };
    // one transformer stack: patch embedding of a [Cin, Hin, Win] image, token 0, pre-LN blocks (vit.py)
    struct VitStack {
        int Cin = 0, Hin = 0, Win = 0, ps = 1, Ht = 0, Wt = 0, Tn = 0, K = 0;  // Tn = Ht*Wt + 1 tokens, K = Cin*ps*ps
        long long o_sal = -1, o_pos = 0, o_pew = 0, o_peb = 0;
        std::vector<BlockLayout> blocks;
        Da3Acts a[3];
    } vloc, vrel;  // local crop (DA3, MergedDA3, DA6) and relative canvas (DA6)
    float *sal2[3] = {nullptr, nullptr, nullptr}, *d_sal2 = nullptr;  // DA6: fc1(LN(relative token 0)) and its gradient
    // generalized conv trunks (rectangular input, any source): relative / merged models
    struct GTrunk {
        int Cin, Hin, Win, H1, W1, Hp, Wp, H2, W2, D;
        long long c1w, c1b, c2w, c2b;
        struct A { float *col1, *y1, *p, *col2, *y2; uint8_t *arg; } a[3];
        float *d_y2, *d_col2, *d_p, *d_y1;
    } gt_rel, gt_loc;
    int area_mask = 0, X = 0, Y = 0, Drel = 0, Dloc = 0;
    bool merged_conv = false, relative_conv = false, merged_da3 = false, relative_vit = false;
    HeadLayout rse, lse;            // relative_state_embedder (MergedDA3, MergedIQN), local_state_embedder (MergedIQN)
    float *lse_h[3] = {nullptr, nullptr, nullptr}, *lse_out[3] = {nullptr, nullptr, nullptr}, *d_locsal = nullptr, *d_locflat = nullptr;
    const int8_t *pos = nullptr, *next_pos = nullptr;  // bound by maexp_learner_bind_positions
    // extra canvas channels behind the decoded observation (objects_channel / agents_channel / destination_channel):
    // ex_o = T, ex_a = num_agents, ex_d = 1 where the model reads them (0 otherwise); Cx = their sum; set 0 belongs to the
    // state batch (h->pos), set 1 to the new-state batch (h->next_pos); bound by maexp_learner_bind_canvas_inputs
    int ex_o = 0, ex_a = 0, ex_d = 0, Cx = 0, Tn_obj = 0, Nteam = 0;
    struct CanvasIn { const uint32_t *obj = nullptr; const int8_t *agents = nullptr; const uint32_t *dest = nullptr; } cin[2];
    float *own_grads = nullptr;       // the library's gradient buffer (h->grads may be re-bound to symmetric memory)
    uint32_t *dp_epoch = nullptr;     // barrier epoch of the fused data-parallel apply
    float *dp_block_part = nullptr;   // its block partials
    int dp_blocks = 0;
    int8_t *los = nullptr;  // [X,Y,V,V] line-of-sight table (maexp_learner_set_los): the merged models' noise-free area channel
    float *canvas[3] = {nullptr, nullptr, nullptr}, *feat[3] = {nullptr, nullptr, nullptr}, *d_feat = nullptr;
    float *rel_flat[3] = {nullptr, nullptr, nullptr}, *rse_h[3] = {nullptr, nullptr, nullptr}, *rse_out[3] = {nullptr, nullptr, nullptr};
    float *d_relsal = nullptr, *d_relflat = nullptr;
    float *da_dx = nullptr, *da_dx1 = nullptr, *da_dn = nullptr, *da_gx = nullptr, *da_dqkv = nullptr, *da_dao = nullptr,
          *da_dg1 = nullptr, *da_df1 = nullptr, *da_dnf = nullptr, *da_dtok0 = nullptr, *da_dtok = nullptr;
    // backward workspace
    float *d_out[2], *d_h, *d_flat, *d_y2, *d_col2, *d_p, *d_y1, *per_sample, *m_scratch, *partial, *q_tmp;
    int nparts;
};

namespace {

// Lazily allocated kernel workspaces (weight panels every reader's prepare kernel overwrites completely): NO memset -- cudaMemset
// runs on the legacy default stream, which does not order against the non-blocking side streams the update's passes run on, and
// zeroed a panel after its prepare kernel had written it (a first-call race: gradients off by 40 % once in a few runs).
template <typename T>
int lalloc_scratch(maexp_learner *h, T **ptr, size_t count) {
    void *d = nullptr;
    MAEXP_CUDA_L(cudaMalloc(&d, (count ? count : 1) * sizeof(T)));
    h->allocs.push_back(d);
    *ptr = static_cast<T *>(d);
    return 0;
}

template <typename T>
int lalloc(maexp_learner *h, T **ptr, size_t count) {
    void *d = nullptr;
    const size_t bytes = (count ? count : 1) * sizeof(T);
    MAEXP_CUDA_L(cudaMalloc(&d, bytes));
    MAEXP_CUDA_L(cudaMemset(d, 0, bytes));
    h->allocs.push_back(d);
    *ptr = static_cast<T *>(d);
    return 0;
}

inline dim3 grid1d(long long total, int g, int threads = 256, int cap = 1024) {
    long long blocks = (total + threads - 1) / threads;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return dim3((unsigned)blocks, (unsigned)g);
}

constexpr long long kSplitWorkspaceFloats = 8ll << 20;  // 32 MB of partial tiles
float *g_splitk_ws = nullptr;                            // set per call from the handle (single host thread)
// set by maexp_learner_q_values around the forward pass of the dueling distributional networks: the heads kernel may write the
// greedy values straight to q_out (and then sets done), which replaces dist_q_kernel and the 255 head outputs per row
struct FusedQ { float *q_out = nullptr; long long q_g = 0; int categorical = 0; bool done = false; } g_fused_q;
// MAEXP_LEARNER_STAMPS=1: one-thread kernels that write %globaltimer at the phase boundaries of an update, on the stream
// that reaches them -- the timeline of the graph-replayed update WITHOUT a profiler attached (CUPTI serialises branches)
__global__ void stamp_kernel(unsigned long long *slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *slot = t;
}
unsigned long long *g_stamps = nullptr;
void stamp(cudaStream_t st, int slot) {
    static const bool on = getenv("MAEXP_LEARNER_STAMPS") != nullptr;
    if (!on) return;
    // MAEXP_LEARNER_STAMPS=1: every slot; any other value: a bit mask of the slots to keep (every stamp kernel on the chain costs
    // ~2.5 us and changes which of the kernels that are ready together starts first, so a sparse mask perturbs the update less)
    static const unsigned long long mask = strtoull(getenv("MAEXP_LEARNER_STAMPS"), nullptr, 0);
    if (mask != 1ull && !((mask >> slot) & 1ull)) return;
    if (!g_stamps) {
        if (cudaMalloc(&g_stamps, 64 * sizeof(unsigned long long)) != cudaSuccess) return;
        cudaMemset(g_stamps, 0, 64 * sizeof(unsigned long long));
        cudaDeviceSynchronize();
    }
    stamp_kernel<<<1, 1, 0, st>>>(g_stamps + slot);
}
bool g_inference_tc = false;                             // set by the q_values entry points around their forward pass
bool g_wgrad_tc = false;                                 // set by backward(): the weight-gradient GEMMs of this update on tcgen05
constexpr int kAttnMapsBatch = 256;                      // q_values batches up to this size keep their attention maps
struct InferenceScope {                                  // (every model family: conv heads, transformer projections)
    explicit InferenceScope(bool on) { g_inference_tc = on; }
    ~InferenceScope() { g_inference_tc = false; }
};

void gemm(cudaStream_t st, int G, const float *A, const float *Bm, float *C, const float *bias, int M, int N, int K,
          long long sam, long long sak, long long sbk, long long sbn, long long scm, long long scn, long long gA,
          long long gB, long long gC, long long gBias, int relu = 0, int accumulate = 0, const float *aux = nullptr,
          long long gAux = 0) {
    GemmArgs a;
    a.A = A; a.B = Bm; a.C = C; a.bias = bias; a.aux = aux;
    a.M = M; a.N = N; a.K = K;
    a.sam = sam; a.sak = sak; a.sbk = sbk; a.sbn = sbn; a.scm = scm; a.scn = scn;
    a.gA = gA; a.gB = gB; a.gC = gC; a.gBias = gBias; a.gAux = gAux;
    a.relu = relu; a.accumulate = accumulate; a.mask_aux = aux != nullptr;
    // thousands of rows, both operands K-major and 16-byte aligned along K: the large-tile kernels
    auto aligned16 = [](const void *ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    // acting / evaluation forward passes (g_inference_tc): tensor cores, three TF32 terms per product (fp32-level accuracy)
    if (g_inference_tc && M >= 512 && K >= 32 && sak == 1 && sbk == 1 && scn == 1 && !accumulate && !aux &&
        tc_gemm_launch(st, G, A, Bm, bias, C, M, N, K, sam, sbn, scm, gA, gB, gBias, gC, relu, 3))
        return;
    // weight gradients of an update (C = dY^T X: both operands k-outermost, the batch is K): tensor cores, three TF32 terms
    if (g_wgrad_tc && sam == 1 && sbn == 1 && scn == 1 && !bias && !relu && !accumulate && !aux && K >= 128 && M >= 32 && N >= 64) {
        int ks = 1;
        static int dbg_slot = 14;  // MAEXP_LEARNER_STAMPS: when each tensor-core weight-gradient kernel started / ended on its side stream
        if (dbg_slot >= 38) dbg_slot = 14;
        const int s0 = dbg_slot;
        dbg_slot += 2;
        stamp(st, s0);
        struct StampAfter { cudaStream_t st; int slot; ~StampAfter() { stamp(st, slot); } } stamp_after{st, s0 + 1};
        if (tc_gemm_tn_launch(st, G, A, Bm, C, M, N, K, sak, sbk, scm, gA, gB, gC, g_splitk_ws, g_splitk_ws ? kSplitWorkspaceFloats : 0,
                              &ks)) {
            if (ks > 1) {
                const long long total = (long long)M * N;
                splitk_reduce_kernel<<<dim3((unsigned)((total + 255) / 256 > 256 ? 256 : (total + 255) / 256), G), 256, 0, st>>>(
                    a, ks, G, g_splitk_ws);
            }
            return;
        }
    }
    if (M >= 2048 && sak == 1 && sbk == 1 && (K & 3) == 0 && (sam & 3) == 0 && (gA & 3) == 0 && aligned16(A)) {
        dim3 grid((N + kBigN - 1) / kBigN, (M + kBigM - 1) / kBigM, G);
        gemm_big_kernel<<<grid, 256, 0, st>>>(a);
        return;
    }
    const int tm = (M + kTile - 1) / kTile, tn = (N + kTile - 1) / kTile;
    // split K (powers of two, slices >= 64) until the grid covers two waves of 148 SMs; shapes only → deterministic
    int ksplit = 1;
    while ((long long)tm * tn * G * ksplit < 296 && K / (ksplit * 2) >= 64 &&
           (long long)ksplit * 2 * G * M * N <= kSplitWorkspaceFloats && g_splitk_ws)
        ksplit *= 2;
    const int kslice = ((K + ksplit - 1) / ksplit + kBK - 1) / kBK * kBK;
    dim3 grid(tn, tm, G * ksplit);
    gemm_kernel<<<grid, 256, 0, st>>>(a, ksplit, kslice, g_splitk_ws);
    if (ksplit > 1) {
        const long long total = (long long)M * N;
        splitk_reduce_kernel<<<dim3((unsigned)((total + 255) / 256 > 256 ? 256 : (total + 255) / 256), G), 256, 0, st>>>(
            a, ksplit, G, g_splitk_ws);
    }
}

// bias gradient: column sums as ones[1,M] x dY[M,N] through the split-K GEMM (fixed order)
void colsum(cudaStream_t st, int G, const float *ones, const float *X, float *out, int M, int N, long long gX,
            long long gOut, int accumulate = 0) {
    gemm(st, G, ones, X, out, nullptr, 1, N, M, 0, 0, N, 1, N, 1, 0, gX, gOut, 0, 0, accumulate);
}

// All bias gradients of the MLP heads in one launch: column sums of d_out (second layers) and d_h (first layers) over the batch
// rows, one thread per column, four interleaved partial sums combined in a fixed order (deterministic).  Replaces four
// ones[1,B] x dY GEMMs and their split-K reductions (8 launches, ~60 us of kernel time per C51 team update).
struct BiasGradArgs {
    const float *src[4];
    long long src_g[4], dst_off[4];
    int cols[4], n;
};
__global__ void __launch_bounds__(256) bias_grads_kernel(const BiasGradArgs a, float *grads, long long P, int rows) {
    // 32 columns x 8 row groups per CTA: row group q sums rows q, q + 8, ... with four interleaved partial sums; the eight
    // group sums meet in shared memory and are added in the order q = 0..7 (fixed order: deterministic)
    __shared__ float part[8][33];
    const int g = blockIdx.z, t = blockIdx.y, lane = threadIdx.x & 31, q = threadIdx.x >> 5, c = blockIdx.x * 32 + lane;
    if (t >= a.n) return;
    const int stride = a.cols[t];
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    if (c < stride) {
        const float *s = a.src[t] + (size_t)g * a.src_g[t] + c;
        int r = q;
        for (; r + 24 < rows; r += 32) {
            s0 += s[(size_t)r * stride];
            s1 += s[(size_t)(r + 8) * stride];
            s2 += s[(size_t)(r + 16) * stride];
            s3 += s[(size_t)(r + 24) * stride];
        }
        for (; r < rows; r += 8) s0 += s[(size_t)r * stride];
    }
    part[q][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (q == 0 && c < stride) {
        float sum = part[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) sum += part[k][lane];
        grads[(size_t)g * P + a.dst_off[t] + c] = sum;
    }
}

// Weight gradients of an update as ONE grouped launch: C[m][n] = sum_k A[k][m] B[k][n] for up to five problems (both operands
// "K-outer": row = batch sample, the layout d_out / d_h / activations already have), 64 x 64 output tiles, the whole K range in
// one CTA in ascending order (deterministic, no split-K workspace or reduction launch), register-prefetched 16-row slabs.
// Replaces the strided split-K GEMM + reduce pairs (2 launches per problem) on the side streams of backward().
struct WgProblem {
    const float *A, *B;
    float *C;
    long long gA, gB;
    int M, N, K, lda, ldb, ldc, tiles_n;
};
struct WgArgs {
    WgProblem p[5];
    int tile0[6];  // prefix sums of the tile counts
    int n;
};
__global__ void __launch_bounds__(256) wgrad_group_kernel(const WgArgs a, long long gC) {
    __shared__ __align__(16) float As[16][64 + 4];
    __shared__ __align__(16) float Bs[16][64 + 4];
    int pi = 0;
    while (pi + 1 < a.n && (int)blockIdx.x >= a.tile0[pi + 1]) ++pi;
    const WgProblem &P = a.p[pi];
    const int t = blockIdx.x - a.tile0[pi], tm = t / P.tiles_n, tn = t - tm * P.tiles_n, g = blockIdx.y;
    const int m0 = tm * 64, n0 = tn * 64, tid = threadIdx.x;
    const float *A = P.A + (size_t)g * P.gA, *B = P.B + (size_t)g * P.gB;
    const int lr = tid >> 4, lc = (tid & 15) * 4;  // this thread's slab element group: row lr, columns lc..lc+3
    const bool va = (P.lda & 3) == 0 && (P.gA & 3) == 0 && (reinterpret_cast<uintptr_t>(P.A) & 15) == 0;
    const bool vb = (P.ldb & 3) == 0 && (P.gB & 3) == 0 && (reinterpret_cast<uintptr_t>(P.B) & 15) == 0;
    auto fetch = [&](const float *src, int ld, int lim, int c0, int k, bool vec) {
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (k < P.K) {
            const float *s = src + (size_t)k * ld + c0 + lc;
            if (vec && c0 + lc + 3 < lim) v = *reinterpret_cast<const float4 *>(s);
            else {
                if (c0 + lc < lim) v.x = s[0];
                if (c0 + lc + 1 < lim) v.y = s[1];
                if (c0 + lc + 2 < lim) v.z = s[2];
                if (c0 + lc + 3 < lim) v.w = s[3];
            }
        }
        return v;
    };
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads, a 4 x 4 micro-tile each
    float4 ra = fetch(A, P.lda, P.M, m0, lr, va), rb = fetch(B, P.ldb, P.N, n0, lr, vb);
    for (int k0 = 0; k0 < P.K; k0 += 16) {
        *reinterpret_cast<float4 *>(&As[lr][lc]) = ra;
        *reinterpret_cast<float4 *>(&Bs[lr][lc]) = rb;
        __syncthreads();
        if (k0 + 16 < P.K) {
            ra = fetch(A, P.lda, P.M, m0, k0 + 16 + lr, va);
            rb = fetch(B, P.ldb, P.N, n0, k0 + 16 + lr, vb);
        }
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *C = P.C + (size_t)g * gC;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= P.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < P.N) C[(size_t)m * P.ldc + n] = acc[i][j];
        }
    }
}
void wgrad_group(cudaStream_t st, int G, WgArgs &a, long long gC) {
    a.tile0[0] = 0;
    for (int i = 0; i < a.n; ++i) {
        a.p[i].tiles_n = (a.p[i].N + 63) / 64;
        a.tile0[i + 1] = a.tile0[i] + ((a.p[i].M + 63) / 64) * a.p[i].tiles_n;
    }
    wgrad_group_kernel<<<dim3(a.tile0[a.n], G), 256, 0, st>>>(a, gC);
}

// Weight / bias gradients never feed the data-gradient chain, so inside a backward pass that opts in (g_wg.h set) they
// are issued on the side streams, in rotation, each behind an event recorded on the main stream at that point.
struct WGradCtx {
    maexp_learner *h = nullptr;
    cudaStream_t main_stream = nullptr;
    int next = 0;
    bool used[kSides] = {};
} g_wg;

cudaStream_t wgrad_begin(cudaStream_t st) {
    if (!g_wg.h) return st;
    maexp_learner *h = g_wg.h;
    const int k = g_wg.next;
    g_wg.next = (k + 1) % kSides;
    cudaEventRecord(h->ev_w[k], st);
    cudaStreamWaitEvent(h->side[k], h->ev_w[k], 0);
    g_splitk_ws = h->splitk_side[k];
    g_wg.used[k] = true;
    return h->side[k];
}
// the same, behind an event recorded earlier on the main stream (the work is issued later than its inputs became ready)
cudaStream_t wgrad_begin_at(cudaStream_t st, cudaEvent_t ev) {
    if (!g_wg.h) return st;
    maexp_learner *h = g_wg.h;
    const int k = g_wg.next;
    g_wg.next = (k + 1) % kSides;
    cudaStreamWaitEvent(h->side[k], ev, 0);
    g_splitk_ws = h->splitk_side[k];
    g_wg.used[k] = true;
    return h->side[k];
}
void wgrad_end() {
    if (g_wg.h) g_splitk_ws = g_wg.h->splitk_ws;
}
void wgrad_open(maexp_learner *h, cudaStream_t st) { g_wg = WGradCtx(); g_wg.h = h; g_wg.main_stream = st; }
void wgrad_close() {  // join the side streams back into the main stream
    maexp_learner *h = g_wg.h;
    if (!h) return;
    for (int k = 0; k < kSides; ++k)
        if (g_wg.used[k]) {
            cudaEventRecord(h->ev_join[k], h->side[k]);
            cudaStreamWaitEvent(g_wg.main_stream, h->ev_join[k], 0);
        }
    g_splitk_ws = h->splitk_ws;
    g_wg = WGradCtx();
}

// Inference-only conv trunk (acting / evaluation batches): conv 2x2 -> max_pool2d(2,2) -> conv 2x2 -> flatten in ONE
// kernel, one warp per sample, so no intermediate ever reaches HBM (the layer-by-layer path moves ~12 KB per sample).
// Lanes are output channels: conv1 + pooling with lane = conv1 channel (F1 <= 32), conv2 with lane = output channel
// (and lane + 32 when F2 > 32).  The agent's conv weights sit in shared memory: conv1 k-major, conv2 as eight
// consecutive floats per (input channel, lane) so that one input channel costs two 128-bit loads; the pooled map is
// stored channel-major so that a channel's cells are three 128-bit broadcast loads.  V is a template parameter: every
// cell / tap index is a compile-time constant.  flat[b][oc*H2*W2 + pos] is PyTorch's flatten order (convolution.py:44-55).
// TRAIN (the learner's online(s) pass) also keeps what the backward pass reads: the pooled map p[b][cell][F1] (NHWC) and
// the pooling argmax (bias added before the comparison, first maximum wins: F.max_pool2d on the biased conv output).
// blockIdx.z selects one of up to three passes of an update (online(s), online(s'), target(s')), so that all of them are one
// launch and share the SMs instead of queueing behind each other's shared-memory footprint.
// one warp copies a row of n floats from HBM into shared memory with eight loads in flight per lane (a plain load -> store
// loop is one global round trip per element)
__device__ __forceinline__ void warp_copy_row(float *dst, const float *src, int n, int lane) {
    for (int base = lane; base < n; base += 32 * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (base + 32 * u < n) ? __ldg(src + base + 32 * u) : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (base + 32 * u < n) dst[base + 32 * u] = v[u];
    }
}
struct TrunkPasses {
    const float *obs[3], *params[3];
    float *flat[3];
    int train[3];
};
template <int V, bool TRAIN>
__global__ void __launch_bounds__(512) trunk_fused_fwd_kernel(const __grid_constant__ TrunkPasses tp, int B, int C,
                                                              int F1, int F2, long long stride, long long gObs, long long P,
                                                              long long o_c1w, long long o_c1b, long long o_c2w,
                                                              long long o_c2b, long long gFlat, float *p_out, uint8_t *arg_out,
                                                              long long gP) {
    constexpr int H1 = V - 1, Hp = H1 / 2, H2 = Hp - 1, Q = H2 * H2, NC = Hp * Hp, PC = (NC + 3) / 4 * 4;
    extern __shared__ __align__(16) float sm[];
    const float *obs = tp.obs[blockIdx.z], *params = tp.params[blockIdx.z];
    float *flat = tp.flat[blockIdx.z];
    const bool train = TRAIN && tp.train[blockIdx.z];
    const int g = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int K1 = C * 4, CVV = C * V * V, D = F2 * Q;
    float *w2t = sm;                      // [F1][32][8]: (tap, half) of output channels lane / lane + 32
    float *w1t = w2t + F1 * 256;          // [K1][32]    conv1 weight, k-major
    float *b1 = w1t + K1 * 32;            // [32]
    float *b2 = b1 + 32;                  // [64]
    float *s_p = b2 + 64 + warp * (32 * PC + (CVV + 3) / 4 * 4);  // per warp: pooled [ic][PC] (16-byte aligned), then the observation
    float *s_obs = s_p + 32 * PC;
    const float *W = params + (size_t)g * P;
    for (int base = threadIdx.x; base < F1 * 256; base += blockDim.x * 16) {  // sixteen loads in flight per thread
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int i = base + u * blockDim.x, ic = i >> 8, l = (i >> 3) & 31, tap = (i >> 1) & 3, oc = l + 32 * (i & 1);
            v[u] = (i < F1 * 256 && oc < F2) ? __ldg(W + o_c2w + ((size_t)oc * F1 + ic) * 4 + tap) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int i = base + u * blockDim.x;
            if (i < F1 * 256) w2t[i] = v[u];
        }
    }
    for (int i = threadIdx.x; i < K1 * 32; i += blockDim.x) {
        const int k = i >> 5, f = i & 31;
        w1t[i] = f < F1 ? W[o_c1w + (size_t)f * K1 + k] : 0.0f;
    }
    for (int i = threadIdx.x; i < 32; i += blockDim.x) b1[i] = i < F1 ? W[o_c1b + i] : 0.0f;
    for (int i = threadIdx.x; i < 64; i += blockDim.x) b2[i] = i < F2 ? W[o_c2b + i] : 0.0f;
    for (int i = lane; i < 32 * PC; i += 32) s_p[i] = 0.0f;
    __syncthreads();
    const float *src = obs + (size_t)g * gObs;
    float *dst = flat + (size_t)g * gFlat;
    for (int b = blockIdx.x * nwarps + warp; b < B; b += gridDim.x * nwarps) {
        warp_copy_row(s_obs, src + (size_t)b * stride, CVV, lane);
        __syncwarp();
        // conv1 + max-pool, lane = conv1 output channel: the 3x3 input patch of a pooled cell is read once per input
        // channel and feeds the four conv positions under the pooling window
#pragma unroll
        for (int cell = 0; cell < NC; ++cell) {
            constexpr int dummy = 0;
            (void)dummy;
            const int py = cell / Hp, px = cell % Hp;
            float a00 = 0.0f, a01 = 0.0f, a10 = 0.0f, a11 = 0.0f;
            for (int c = 0; c < C; ++c) {
                const float *o = s_obs + c * V * V + (py * 2) * V + px * 2;
                const float *w = w1t + (c * 4) * 32 + lane;
                const float w0 = w[0], w1 = w[32], w2 = w[64], w3 = w[96];
                const float o00 = o[0], o01 = o[1], o02 = o[2], o10 = o[V], o11 = o[V + 1], o12 = o[V + 2],
                            o20 = o[2 * V], o21 = o[2 * V + 1], o22 = o[2 * V + 2];
                a00 = fmaf(o00, w0, a00); a00 = fmaf(o01, w1, a00); a00 = fmaf(o10, w2, a00); a00 = fmaf(o11, w3, a00);
                a01 = fmaf(o01, w0, a01); a01 = fmaf(o02, w1, a01); a01 = fmaf(o11, w2, a01); a01 = fmaf(o12, w3, a01);
                a10 = fmaf(o10, w0, a10); a10 = fmaf(o11, w1, a10); a10 = fmaf(o20, w2, a10); a10 = fmaf(o21, w3, a10);
                a11 = fmaf(o11, w0, a11); a11 = fmaf(o12, w1, a11); a11 = fmaf(o21, w2, a11); a11 = fmaf(o22, w3, a11);
            }
            if (train) {
                const float bb = b1[lane];
                float best = a00 + bb;
                int bi = 0;
                if (a01 + bb > best) { best = a01 + bb; bi = 1; }
                if (a10 + bb > best) { best = a10 + bb; bi = 2; }
                if (a11 + bb > best) { best = a11 + bb; bi = 3; }
                s_p[lane * PC + cell] = best;
                if (lane < F1) {
                    const size_t pi = (size_t)g * gP + ((size_t)b * NC + cell) * F1 + lane;
                    p_out[pi] = best;
                    arg_out[pi] = (uint8_t)bi;
                }
            } else {
                s_p[lane * PC + cell] = fmaxf(fmaxf(a00, a01), fmaxf(a10, a11)) + b1[lane];  // max(x_i) + b == max(x_i + b)
            }
        }
        __syncwarp();
        // conv2, lane = output channel (two per lane when F2 > 32)
        float acc0[Q], acc1[Q];
#pragma unroll
        for (int pos = 0; pos < Q; ++pos) acc0[pos] = acc1[pos] = 0.0f;
        for (int ic = 0; ic < F1; ++ic) {
            float pv[PC];
#pragma unroll
            for (int v4 = 0; v4 < PC / 4; ++v4) {
                const float4 t = *reinterpret_cast<const float4 *>(s_p + ic * PC + v4 * 4);
                pv[v4 * 4] = t.x; pv[v4 * 4 + 1] = t.y; pv[v4 * 4 + 2] = t.z; pv[v4 * 4 + 3] = t.w;
            }
            const float4 wa = *reinterpret_cast<const float4 *>(w2t + (ic * 32 + lane) * 8);
            const float4 wb = *reinterpret_cast<const float4 *>(w2t + (ic * 32 + lane) * 8 + 4);
            const float w0[4] = {wa.x, wa.z, wb.x, wb.z}, w1[4] = {wa.y, wa.w, wb.y, wb.w};  // [tap] for oc = lane, lane + 32
#pragma unroll
            for (int pos = 0; pos < Q; ++pos) {
#pragma unroll
                for (int tap = 0; tap < 4; ++tap) {
                    const float p = pv[(pos / H2 + (tap >> 1)) * Hp + pos % H2 + (tap & 1)];
                    acc0[pos] = fmaf(p, w0[tap], acc0[pos]);
                    acc1[pos] = fmaf(p, w1[tap], acc1[pos]);
                }
            }
        }
#pragma unroll
        for (int pos = 0; pos < Q; ++pos) {
            if (lane < F2) dst[(size_t)b * D + (size_t)lane * Q + pos] = acc0[pos] + b2[lane];
            if (lane + 32 < F2) dst[(size_t)b * D + (size_t)(lane + 32) * Q + pos] = acc1[pos] + b2[lane + 32];
        }
        __syncwarp();
    }
}

// Backward of the conv trunk below conv2's output, for the activations the fused forward kept.  One warp per sample,
// lane = conv1 channel (= conv2 input channel): d_p = conv2^T(d_flat) with the agent's conv2 weights in shared memory,
// routed through the pooling argmax straight into conv1's weight / bias gradient (the pooled gradient is one non-zero per
// 2x2 window, so conv1's [B*H1*H1, F1] gradient matrix and the 9 216-row reduction over it never exist), plus conv2's bias
// gradient.  Sums run in a fixed order: samples per warp, warps per block in shared memory, blocks in
// trunk_bwd_reduce_kernel — deterministic.  part[g][block][K1*32 + 32 + 64].
template <int V>
__global__ void __launch_bounds__(256) trunk_fused_bwd_kernel(const float *d_flat, const float *obs, const float *params,
                                                              const uint8_t *arg, float *part, int B, int C, int F1, int F2,
                                                              long long stride, long long gObs, long long P, long long o_c2w,
                                                              long long gFlat, long long gP) {
    constexpr int H1 = V - 1, Hp = H1 / 2, H2 = Hp - 1, Q = H2 * H2, NC = Hp * Hp;
    extern __shared__ __align__(16) float sm[];
    const int g = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int K1 = C * 4, CVV = C * V * V, D = F2 * Q, Dp = (D + 3) / 4 * 4, CVVp = (CVV + 3) / 4 * 4;
    const int per_warp = Dp + CVVp + K1 * 32 + 96;
    float *w2s = sm;                                // [F2][32][4]: taps of (oc, ic = lane)
    float *s_d = w2s + F2 * 128 + warp * per_warp;  // this sample's d_flat row
    float *s_obs = s_d + Dp;
    float *gw = s_obs + CVVp;                       // [K1][32] conv1 weight gradient of this warp's samples, then [32] + [64] biases
    const float *W = params + (size_t)g * P;
    // sixteen loads in flight per thread: as a plain load -> store loop this staging was one global round trip per element
    // (32 in a row per thread), 60 % of the kernel's stall samples (ncu, profiles/r2_bwd_chain_hot_lines.txt)
    for (int base = threadIdx.x; base < F2 * 128; base += blockDim.x * 16) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int i = base + u * blockDim.x, oc = i >> 7, ic = (i >> 2) & 31, tap = i & 3;
            v[u] = (i < F2 * 128 && ic < F1) ? __ldg(W + o_c2w + ((size_t)oc * F1 + ic) * 4 + tap) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int i = base + u * blockDim.x;
            if (i < F2 * 128) w2s[i] = v[u];
        }
    }
    for (int i = lane; i < K1 * 32 + 96; i += 32) gw[i] = 0.0f;
    __syncthreads();
    float gb1 = 0.0f, gb2a = 0.0f, gb2b = 0.0f;
    for (int b = blockIdx.x * nwarps + warp; b < B; b += gridDim.x * nwarps) {
        const float *dsrc = d_flat + (size_t)g * gFlat + (size_t)b * D;
        const float *osrc = obs + (size_t)g * gObs + (size_t)b * stride;
        // the pooling argmax of this sample (read after the conv2 products): loaded now, so that its latency hides behind them
        int amax[NC];
#pragma unroll
        for (int cell = 0; cell < NC; ++cell) amax[cell] = lane < F1 ? arg[(size_t)g * gP + ((size_t)b * NC + cell) * F1 + lane] : 0;
        warp_copy_row(s_d, dsrc, D, lane);
        warp_copy_row(s_obs, osrc, CVV, lane);
        __syncwarp();
        float acc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = 0.0f;
        for (int oc = 0; oc < F2; ++oc) {
            const float4 w4 = *reinterpret_cast<const float4 *>(w2s + (oc * 32 + lane) * 4);
            const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int pos = 0; pos < Q; ++pos) {
                const float dv = s_d[oc * Q + pos];
#pragma unroll
                for (int tap = 0; tap < 4; ++tap) {
                    const int cell = (pos / H2 + (tap >> 1)) * Hp + pos % H2 + (tap & 1);
                    acc[cell] = fmaf(dv, w[tap], acc[cell]);
                }
            }
        }
        // conv2 bias gradient: lanes are output channels here
        if (lane < F2)
            for (int pos = 0; pos < Q; ++pos) gb2a += s_d[lane * Q + pos];
        if (lane + 32 < F2)
            for (int pos = 0; pos < Q; ++pos) gb2b += s_d[(lane + 32) * Q + pos];
        // pooling argmax -> the conv1 output position that received each pooled gradient
        int off[NC];
#pragma unroll
        for (int cell = 0; cell < NC; ++cell) {
            const int a = amax[cell];
            off[cell] = (2 * (cell / Hp) + (a >> 1)) * V + 2 * (cell % Hp) + (a & 1);
            gb1 += acc[cell];
        }
        for (int k = 0; k < K1; ++k) {
            const float *o = s_obs + (k >> 2) * V * V + ((k >> 1) & 1) * V + (k & 1);
            float sum = 0.0f;
#pragma unroll
            for (int cell = 0; cell < NC; ++cell) sum = fmaf(acc[cell], o[off[cell]], sum);
            gw[k * 32 + lane] += sum;
        }
        __syncwarp();
    }
    gw[K1 * 32 + lane] = gb1;
    gw[K1 * 32 + 32 + lane] = gb2a;
    gw[K1 * 32 + 64 + lane] = gb2b;
    __syncthreads();
    const int n = K1 * 32 + 96;
    float *dst = part + ((size_t)g * gridDim.x + blockIdx.x) * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float sum = 0.0f;
        for (int w = 0; w < nwarps; ++w) sum += sm[F2 * 128 + w * per_warp + Dp + CVVp + i];
        dst[i] = sum;
    }
}

// block partials -> conv1 weight / bias and conv2 bias gradients (fixed order)
__global__ void trunk_bwd_reduce_kernel(const float *part, float *grads, int nblk, int K1, int F1, int F2, long long P,
                                        long long o_c1w, long long o_c1b, long long o_c2b) {
    const int g = blockIdx.y, n = K1 * 32 + 96;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float sum = 0.0f;
    for (int blk = 0; blk < nblk; ++blk) sum += part[((size_t)g * nblk + blk) * n + i];
    float *G = grads + (size_t)g * P;
    if (i < K1 * 32) {
        const int k = i >> 5, f = i & 31;
        if (f < F1) G[o_c1w + (size_t)f * K1 + k] = sum;
    } else {
        const int j = i - K1 * 32;
        if (j < 32) { if (j < F1) G[o_c1b + j] = sum; }
        else if (j - 32 < F2) G[o_c2b + j - 32] = sum;
    }
}

// One-hidden-layer MLP head (mlp.py:28-34) in ONE kernel for up to two heads of every agent: a CTA owns 32 rows of one
// (agent, head), computes hidden = relu(x W1^T + b1) (hidden <= 128 units) with the tile kept in shared memory as the A
// operand of the second layer, and streams W1 / W2 (PyTorch [out][in] layout, K-major) through shared memory in 32-wide
// K chunks with a register prefetch of the next chunk.  Replaces gemm + split-K reduce + gemm + split-K reduce per head
// (four launches of ~10 us each at the update batch sizes); the sums run k = 0..K-1 in order (deterministic).
// blockIdx.y = head + nheads * pass: up to three passes of an update (their own input, parameter set and outputs) share
// one launch.
struct MlpHeads {
    long long w1[2], b1[2], w2[2], b2[2];
    int out[2], nheads;
    int vec_w1[2], vec_w2[2];  // 128-bit loads along K are legal for this head's weights
    const float *x[3], *params[3];
    float *hbuf[3][2], *outp[3][2];
    int vec_x[3];
};
constexpr int kMlpTM = 32, kMlpTN = 128, kMlpBK = 32, kMlpMaxIn = 1024;

__global__ void __launch_bounds__(256) mlp_fused_fwd_kernel(long long x_row, long long x_g, long long P,
                                                            const __grid_constant__ MlpHeads hd, int rows,
                                                            long long rows_max, int din, int H) {
    constexpr int TM = kMlpTM, TN = kMlpTN, BK = kMlpBK;
    __shared__ __align__(16) float As[BK][TM + 4];
    __shared__ __align__(16) float Bs[BK][TN + 4];
    __shared__ __align__(16) float Hs[TN][TM + 4];
    const int g = blockIdx.z, pass = blockIdx.y / hd.nheads, head = blockIdx.y - pass * hd.nheads, m0 = blockIdx.x * TM;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const float *X = hd.x[pass] + (size_t)g * x_g;
    const float *Wg = hd.params[pass] + (size_t)g * P;
    const int vec_x = hd.vec_x[pass];
    const int am = tid >> 3, ak = (tid & 7) * 4;  // this thread's A element group: row am, k offset ak..ak+3
    float4 ra, rb[4];
    auto fetch_a = [&](int k0) {
        const int m = m0 + am, k = k0 + ak;
        ra = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (m < rows) {
            const float *src = X + (size_t)m * x_row + k;
            if (vec_x && k + 3 < din) ra = *reinterpret_cast<const float4 *>(src);
            else {
                if (k < din) ra.x = src[0];
                if (k + 1 < din) ra.y = src[1];
                if (k + 2 < din) ra.z = src[2];
                if (k + 3 < din) ra.w = src[3];
            }
        }
    };
    auto fetch_b = [&](const float *W, int N, int K, int n0, int k0, int vec) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int idx = tid + r * 256, n = n0 + (idx >> 3), k = k0 + (idx & 7) * 4;
            rb[r] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (n < N) {
                const float *src = W + (size_t)n * K + k;
                if (vec && k + 3 < K) rb[r] = *reinterpret_cast<const float4 *>(src);
                else {
                    if (k < K) rb[r].x = src[0];
                    if (k + 1 < K) rb[r].y = src[1];
                    if (k + 2 < K) rb[r].z = src[2];
                    if (k + 3 < K) rb[r].w = src[3];
                }
            }
        }
    };
    auto store_a = [&]() { As[ak][am] = ra.x; As[ak + 1][am] = ra.y; As[ak + 2][am] = ra.z; As[ak + 3][am] = ra.w; };
    auto store_b = [&]() {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int idx = tid + r * 256, n = idx >> 3, k = (idx & 7) * 4;
            Bs[k][n] = rb[r].x; Bs[k + 1][n] = rb[r].y; Bs[k + 2][n] = rb[r].z; Bs[k + 3][n] = rb[r].w;
        }
    };
    float acc[4][4];
    auto clear = [&]() {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    };
    auto mma = [&](const float (*A)[TM + 4]) {
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&A[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    };
    // ---- layer 1: hidden[32][H] = relu(X W1^T + b1)
    {
        const float *W1 = Wg + hd.w1[head];
        const int vec = hd.vec_w1[head];
        clear();
        fetch_a(0);
        fetch_b(W1, H, din, 0, 0, vec);
        for (int k0 = 0; k0 < din; k0 += BK) {
            store_a();
            store_b();
            __syncthreads();
            if (k0 + BK < din) { fetch_a(k0 + BK); fetch_b(W1, H, din, 0, k0 + BK, vec); }
            mma(As);
            __syncthreads();
        }
        const float *b1 = Wg + hd.b1[head];
        float *hb = hd.hbuf[pass][head] + (size_t)g * rows_max * H;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = tx * 4 + j;
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = n < H ? fmaxf(__fadd_rn(acc[i][j], b1[n]), 0.0f) : 0.0f;
            *reinterpret_cast<float4 *>(&Hs[n][ty * 4]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();
        // hidden activations to HBM (the backward pass and its ReLU mask read them), coalesced along the hidden index
        for (int i = tid; i < TM * TN; i += 256) {
            const int m = i >> 7, n = i & (TN - 1);
            if (m0 + m < rows && n < H) hb[(size_t)(m0 + m) * H + n] = Hs[n][m];
        }
    }
    // ---- layer 2: out[32][O] = hidden W2^T + b2, 128 output columns at a time
    const float *W2 = Wg + hd.w2[head], *b2 = Wg + hd.b2[head];
    const int O = hd.out[head], vec2 = hd.vec_w2[head];
    float *outp = hd.outp[pass][head] + (size_t)g * rows_max * O;
    for (int n0 = 0; n0 < O; n0 += TN) {
        clear();
        fetch_b(W2, O, H, n0, 0, vec2);
        for (int k0 = 0; k0 < H; k0 += BK) {
            store_b();
            __syncthreads();
            if (k0 + BK < H) fetch_b(W2, O, H, n0, k0 + BK, vec2);
            mma(Hs + k0);
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + ty * 4 + i;
            if (m >= rows) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tx * 4 + j;
                if (n < O) outp[(size_t)m * O + n] = __fadd_rn(acc[i][j], b2[n]);
            }
        }
    }
}

// Data-gradient chain of the MLP heads in ONE kernel: for every head, dh = (dout W2) * [h > 0] (kept in shared memory as the
// A operand of the next product and written out for the weight-gradient GEMMs), then dx (+)= dh W1, summed over the heads
// in registers.  A CTA owns 32 rows and one 128-column tile of dx (blockIdx.y); both weight matrices are read in their
// PyTorch layout, which is already [k][n] for these two products.  Replaces 2 GEMMs + 2 split-K reductions per head on the
// critical path of the backward pass.  dx == nullptr: only dh (plain MLP on the observation).
// three resident CTAs per SM (<= 80 registers, no spills): at 127 registers two CTAs fill an SM's register file and nothing else
// (a weight-gradient CTA of the side streams) fits beside them
#ifndef MAEXP_MLP_BWD_MINB
#define MAEXP_MLP_BWD_MINB 3
#endif
struct MlpBwdHeads {
    long long w1[2], w2[2];
    int out[2];
    const float *dout[2], *hbuf[2];
    float *dh[2];
};

__global__ void __launch_bounds__(256, MAEXP_MLP_BWD_MINB) mlp_fused_bwd_kernel(const float *params, long long P, const MlpBwdHeads hd, int nheads,
                                                            int rows, long long rows_max, int din, int H, float *dx,
                                                            int vec_w1, float *dx_nhwc, int Q, int accumulate, int dh_only) {
    // dh_only == 1: blockIdx.y is the head, the CTA stops after writing dh (the split chain: mlp_dx_kernel then forms dx from dh
    // without every column tile repeating this product)
    // dh_only == 2 (heads in parallel): blockIdx.y = head * column tiles + column tile; a CTA walks ONE head's chain and adds its
    // dx tile into the zeroed dx / dx_nhwc with atomicAdd -- two addends per element (0 + a + b = 0 + b + a exactly), so the
    // result does not depend on the order the CTAs arrive in
    const int head_par = dh_only == 2, col_tiles = head_par ? (int)gridDim.y / nheads : (int)gridDim.y;
    const int my_head = head_par ? (int)blockIdx.y / col_tiles : (int)blockIdx.y, my_tile = head_par ? (int)blockIdx.y % col_tiles : (int)blockIdx.y;
    if (head_par) dh_only = 0;
    constexpr int TM = kMlpTM, TN = kMlpTN, BK = kMlpBK;
    __shared__ __align__(16) float As[BK][TM + 4];
    __shared__ __align__(16) float Bs[BK][TN + 4];
    __shared__ __align__(16) float Hs[TN][TM + 4];
    const int g = blockIdx.z, m0 = blockIdx.x * TM, n0 = dh_only ? 0 : my_tile * TN;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const float *Wg = params + (size_t)g * P;
    const int am = tid >> 3, ak = (tid & 7) * 4;
    float4 ra, rb[4];
    auto fetch_a = [&](const float *A, int K, int k0) {  // A[m][k], K-major rows (dout), scalar loads: K = 51 / 204 / 4 ...
        const int m = m0 + am, k = k0 + ak;
        ra = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (m < rows) {
            const float *src = A + (size_t)m * K + k;
            if (k < K) ra.x = src[0];
            if (k + 1 < K) ra.y = src[1];
            if (k + 2 < K) ra.z = src[2];
            if (k + 3 < K) ra.w = src[3];
        }
    };
    auto fetch_b = [&](const float *W, int K, int N, int k0, int nbase, int vec) {  // W[k][n] row-major
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int idx = tid + r * 256, k = k0 + (idx >> 5), n = nbase + (idx & 31) * 4;
            rb[r] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (k < K) {
                const float *src = W + (size_t)k * N + n;
                if (vec && n + 3 < N) rb[r] = *reinterpret_cast<const float4 *>(src);
                else {
                    if (n < N) rb[r].x = src[0];
                    if (n + 1 < N) rb[r].y = src[1];
                    if (n + 2 < N) rb[r].z = src[2];
                    if (n + 3 < N) rb[r].w = src[3];
                }
            }
        }
    };
    auto store_a = [&]() { As[ak][am] = ra.x; As[ak + 1][am] = ra.y; As[ak + 2][am] = ra.z; As[ak + 3][am] = ra.w; };
    auto store_b = [&]() {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int idx = tid + r * 256;
            *reinterpret_cast<float4 *>(&Bs[idx >> 5][(idx & 31) * 4]) = rb[r];
        }
    };
    float acc[4][4], accx[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) accx[i][j] = 0.0f;
    auto mma = [&](const float (*A)[TM + 4], float (&c)[4][4]) {
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&A[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) c[i][j] = fmaf(av[i], bv[j], c[i][j]);
        }
    };
    const int vec_h = (H & 3) == 0;  // W2 rows are H floats: 16-byte aligned when the tensor is (checked by the host for W1 only)
    for (int head = (dh_only || head_par) ? my_head : 0; head < ((dh_only || head_par) ? my_head + 1 : nheads); ++head) {
        const int O = head == 0 ? hd.out[0] : hd.out[1];
        const float *dout = (head == 0 ? hd.dout[0] : hd.dout[1]) + (size_t)g * rows_max * O;
        const float *hb = (head == 0 ? hd.hbuf[0] : hd.hbuf[1]) + (size_t)g * rows_max * H;
        float *dh = (head == 0 ? hd.dh[0] : hd.dh[1]) + (size_t)g * rows_max * H;
        const float *W2 = Wg + (head == 0 ? hd.w2[0] : hd.w2[1]);
        const float *W1 = Wg + (head == 0 ? hd.w1[0] : hd.w1[1]);
        const int vec2 = vec_h && ((reinterpret_cast<uintptr_t>(W2) & 15) == 0);
        // ---- dh[32][H] = dout[32][O] W2[O][H], masked by the forward ReLU
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
        fetch_a(dout, O, 0);
        fetch_b(W2, O, H, 0, 0, vec2);
        for (int k0 = 0; k0 < O; k0 += BK) {
            store_a();
            store_b();
            __syncthreads();
            if (k0 + BK < O) { fetch_a(dout, O, k0 + BK); fetch_b(W2, O, H, k0 + BK, 0, vec2); }
            mma(As, acc);
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = tx * 4 + j;
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = m0 + ty * 4 + i;
                v[i] = (n < H && m < rows && hb[(size_t)m * H + n] > 0.0f) ? acc[i][j] : 0.0f;
            }
            *reinterpret_cast<float4 *>(&Hs[n][ty * 4]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();
        if (my_tile == 0 || dh_only)
            for (int i = tid; i < TM * TN; i += 256) {
                const int m = i >> 7, n = i & (TN - 1);
                if (m0 + m < rows && n < H) dh[(size_t)(m0 + m) * H + n] = Hs[n][m];
            }
        if (!dx || dh_only) { __syncthreads(); continue; }
        // ---- dx tile += dh[32][H] W1[H][din]
        fetch_b(W1, H, din, 0, n0, vec_w1);
        for (int k0 = 0; k0 < H; k0 += BK) {
            store_b();
            __syncthreads();
            if (k0 + BK < H) fetch_b(W1, H, din, k0 + BK, n0, vec_w1);
            mma(Hs + k0, accx);
            __syncthreads();
        }
    }
    if (!dx || dh_only) return;
    float *dst = dx + (size_t)g * rows_max * din;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < din) {
                if (head_par) {
                    atomicAdd(&dst[(size_t)m * din + n], accx[i][j]);
                    if (dx_nhwc) atomicAdd(&dx_nhwc[(size_t)g * rows_max * din + ((size_t)m * Q + n % Q) * (din / Q) + n / Q], accx[i][j]);
                    continue;
                }
                if (accumulate) accx[i][j] = __fadd_rn(dst[(size_t)m * din + n], accx[i][j]);  // dx += (the caller's earlier heads)
                dst[(size_t)m * din + n] = accx[i][j];
                // the same gradient as conv2's NHWC output gradient [m][pos][oc] (flatten order is oc*Q + pos): what the
                // conv2 weight-gradient GEMM reads
                if (dx_nhwc) dx_nhwc[(size_t)g * rows_max * din + ((size_t)m * Q + n % Q) * (din / Q) + n / Q] = accx[i][j];
            }
        }
    }
}

// Second half of the split data-gradient chain: dx[rows][din] = sum over the heads of dh_head[rows][H] W1_head[H][din], dh read
// back from HBM (written by mlp_fused_bwd_kernel with dh_only).  A CTA owns 32 rows and one 128-column tile; every output is
// the same fmaf chain (head 0, k ascending, then head 1) as in the one-kernel form, so both give the same bits.
__global__ void __launch_bounds__(256) mlp_dx_kernel(const float *params, long long P, const MlpBwdHeads hd, int nheads, int rows,
                                                     long long rows_max, int din, int H, float *dx, int vec_w1, float *dx_nhwc,
                                                     int Q) {
    constexpr int TM = kMlpTM, TN = kMlpTN, BK = kMlpBK;
    __shared__ __align__(16) float As[2][BK][TM + 4];
    __shared__ __align__(16) float Bs[2][BK][TN + 4];
    const int g = blockIdx.z, m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const float *Wg = params + (size_t)g * P;
    const int am = tid >> 3, ak = (tid & 7) * 4;
    const bool vec_h = (H & 3) == 0 && ((rows_max * H) & 3) == 0;
    float4 ra, rb[4];
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    const int nkb = (H + BK - 1) / BK, steps = nheads * nkb;
    auto fetch = [&](int s) {
        const int head = s / nkb, k0 = (s - head * nkb) * BK;
        const float *dh = (head == 0 ? hd.dh[0] : hd.dh[1]) + (size_t)g * rows_max * H;
        const float *W1 = Wg + (head == 0 ? hd.w1[0] : hd.w1[1]);
        const int m = m0 + am, k = k0 + ak;
        ra = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (m < rows && k < H) {
            const float *src = dh + (size_t)m * H + k;
            if (vec_h && ((reinterpret_cast<uintptr_t>(dh) & 15) == 0) && k + 3 < H) ra = *reinterpret_cast<const float4 *>(src);
            else {
                ra.x = src[0];
                if (k + 1 < H) ra.y = src[1];
                if (k + 2 < H) ra.z = src[2];
                if (k + 3 < H) ra.w = src[3];
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int idx = tid + r * 256, kk = k0 + (idx >> 5), n = n0 + (idx & 31) * 4;
            rb[r] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (kk < H) {
                const float *src = W1 + (size_t)kk * din + n;
                if (vec_w1 && n + 3 < din) rb[r] = *reinterpret_cast<const float4 *>(src);
                else {
                    if (n < din) rb[r].x = src[0];
                    if (n + 1 < din) rb[r].y = src[1];
                    if (n + 2 < din) rb[r].z = src[2];
                    if (n + 3 < din) rb[r].w = src[3];
                }
            }
        }
    };
    auto store = [&](int buf) {
        As[buf][ak][am] = ra.x; As[buf][ak + 1][am] = ra.y; As[buf][ak + 2][am] = ra.z; As[buf][ak + 3][am] = ra.w;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int idx = tid + r * 256;
            *reinterpret_cast<float4 *>(&Bs[buf][idx >> 5][(idx & 31) * 4]) = rb[r];
        }
    };
    fetch(0);
    store(0);
    __syncthreads();
    for (int s = 0; s < steps; ++s) {  // double-buffered tiles: one barrier per K step
        const int buf = s & 1;
        if (s + 1 < steps) fetch(s + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (s + 1 < steps) store(buf ^ 1);
        __syncthreads();
    }
    float *dst = dx + (size_t)g * rows_max * din;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < din) {
                dst[(size_t)m * din + n] = acc[i][j];
                if (dx_nhwc) dx_nhwc[(size_t)g * rows_max * din + ((size_t)m * Q + n % Q) * (din / Q) + n / Q] = acc[i][j];
            }
        }
    }
}

// the fused trunk for n <= 3 passes in one launch; at most one of them (train[k] != 0) keeps p / arg for the backward pass
bool trunk_forward_multi(maexp_learner *h, cudaStream_t st, int n, const float *const *params, const float *const *obs,
                         maexp_learner::Acts *const *acts, const int *train, int B) {
    const int G = h->G, C = h->C, V = h->V, F1 = h->F1, F2 = h->F2, Hp = h->Hp;
    const long long P = h->P, Bmax = h->B, stride = h->cfg.obs_stride;
    if (!h->fused_trunk || F1 > 32 || F2 > 64 || !(V == 5 || V == 7 || V == 9) || n < 1 || n > 3) return false;
    // acting / evaluation batches: conv1 + pooling on the CUDA cores, conv2 as an implicit GEMM on the tensor cores
    // ... and the passes of an update (>= 128 rows; the training pass also keeps the pooled map and the argmax)
    const bool tc_acting = g_inference_tc && !train[0] && B >= 2048, tc_update = !g_inference_tc && h->tc_mlp && B >= 128 &&
               (h->cfg.kind == MAEXP_LEARNER_C51 || h->cfg.kind == MAEXP_LEARNER_DQN_CONV_MLP);  // (the quantile-Huber gradients of QR-DQN / IQN amplify the 2e-6 past the gradient bar)
    if (h->use_tc && h->tc_trunk && n == 1 && (tc_acting || tc_update) && V == 7 && F1 == 32 && F2 == 64) {
        float *panel = nullptr;
        if (tc_acting) {
            if (!h->tc_trunk_ws) {
                float *ws = nullptr;
                if (lalloc_scratch(h, &ws, tc_trunk_workspace_floats(G)) == 0) h->tc_trunk_ws = ws;
            }
            panel = h->tc_trunk_ws;
        } else {
            for (int k = 0; k < 4 && !panel; ++k) {
                auto &e = h->tc_trunk_upd[k];
                if (e.key == acts[0]->flat || !e.key) {
                    if (!e.ws) {
                        float *ws = nullptr;
                        if (lalloc_scratch(h, &ws, tc_trunk_workspace_floats(G))) break;
                        e.ws = ws;
                    }
                    e.key = acts[0]->flat;
                    panel = e.ws;
                }
            }
        }
        if (panel) {
            TcTrunkFwd t = {};
            t.obs = obs[0]; t.params = params[0]; t.flat = acts[0]->flat;
            t.stride = stride; t.g_obs = (long long)B * stride; t.P = P;
            t.o_c1w = h->o_c1w; t.o_c1b = h->o_c1b; t.o_c2w = h->o_c2w; t.o_c2b = h->o_c2b; t.g_flat = Bmax * h->D;
            t.B = B; t.C = C; t.V = V; t.F1 = F1; t.F2 = F2;
            t.workspace = panel;
            if (train[0]) { t.p_out = acts[0]->p; t.arg_out = acts[0]->arg; t.g_p = Bmax * Hp * Hp * F1; }
            if (tc_trunk_forward(st, G, t)) {
                if (train[0]) { h->bwd_obs = obs[0]; h->fused_trunk_acts = true; }
                return true;
            }
        }
    }
    // the 32 KB of conv2 weights are staged once per CTA: 16 warps per CTA at the update batch sizes, so that the three
    // forward passes of an update (3 x 128 CTAs of ~69 KB at B = 256, G = 8) are resident together
    const int warps = (B >= 128 && B < 512) ? 16 : B >= 64 ? 8 : 4, pc = (Hp * Hp + 3) / 4 * 4;
    const size_t smem = sizeof(float) * ((size_t)F1 * 256 + (size_t)C * 4 * 32 + 96 +
                                         (size_t)warps * (32 * pc + (C * V * V + 3) / 4 * 4));
    if (smem > 200 * 1024) return false;
    TrunkPasses tp = {};
    maexp_learner::Acts *kept = nullptr;
    for (int k = 0; k < n; ++k) {
        tp.obs[k] = obs[k]; tp.params[k] = params[k]; tp.flat[k] = acts[k]->flat; tp.train[k] = train[k];
        if (train[k]) { kept = acts[k]; h->bwd_obs = obs[k]; }
    }
    auto fn = !kept ? (V == 5 ? trunk_fused_fwd_kernel<5, false> : V == 7 ? trunk_fused_fwd_kernel<7, false>
                                                                          : trunk_fused_fwd_kernel<9, false>)
                    : (V == 5 ? trunk_fused_fwd_kernel<5, true> : V == 7 ? trunk_fused_fwd_kernel<7, true>
                                                                         : trunk_fused_fwd_kernel<9, true>);
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int blocks = (B + warps - 1) / warps;
    if (blocks > 148 * 4) blocks = 148 * 4;
    fn<<<dim3(blocks, G, n), warps * 32, smem, st>>>(tp, B, C, F1, F2, stride, (long long)B * stride, P, h->o_c1w, h->o_c1b,
                                                    h->o_c2w, h->o_c2b, Bmax * h->D, kept ? kept->p : nullptr,
                                                    kept ? kept->arg : nullptr, Bmax * Hp * Hp * F1);
    if (kept) h->fused_trunk_acts = true;
    return true;
}

// ---- conv trunk: obs [G][B][obs_stride] -> a.flat [G][Bmax][D]
void trunk_forward(maexp_learner *h, cudaStream_t st, const float *params, const float *obs, int B, maexp_learner::Acts &a,
                   bool inference = false) {
    const int G = h->G, C = h->C, V = h->V, F1 = h->F1, F2 = h->F2;
    const long long P = h->P, Bmax = h->B, stride = h->cfg.obs_stride;
    const int H1 = h->H1, Hp = h->Hp, H2 = h->H2;
    // one fused kernel for the whole trunk; the training pass (inference == false) also keeps p / arg, and trunk_backward
    // builds the two im2col matrices its weight-gradient GEMMs read on their side streams (h->fused_trunk_acts)
    if (!inference) h->fused_trunk_acts = false;
    {
        const float *ps[1] = {params}, *os[1] = {obs};
        maexp_learner::Acts *as[1] = {&a};
        const int tr[1] = {inference ? 0 : 1};
        if (trunk_forward_multi(h, st, 1, ps, os, as, tr, B)) return;
    }
    const long long gcol1 = Bmax * H1 * H1 * C * 4, gy1 = Bmax * H1 * H1 * F1, gp = Bmax * Hp * Hp * F1;
    const long long gcol2 = Bmax * H2 * H2 * F1 * 4, gy2 = Bmax * H2 * H2 * F2;
    im2col_kernel<<<grid1d((long long)B * H1 * H1 * C * 4, G), 256, 0, st>>>(
        obs, a.col1, B, C, H1, H1, stride, (long long)V * V, V, 1, (long long)B * stride, gcol1);
    gemm(st, G, a.col1, params + h->o_c1w, a.y1, params + h->o_c1b, B * H1 * H1, F1, C * 4, C * 4, 1, 1, C * 4, F1, 1,
         gcol1, P, gy1, P);
    pool_fwd_kernel<<<grid1d((long long)B * Hp * Hp * F1, G), 256, 0, st>>>(a.y1, a.p, a.arg, B, H1, H1, F1, gy1, gp);
    im2col_kernel<<<grid1d((long long)B * H2 * H2 * F1 * 4, G), 256, 0, st>>>(
        a.p, a.col2, B, F1, H2, H2, (long long)Hp * Hp * F1, 1, (long long)Hp * F1, F1, gp, gcol2);
    gemm(st, G, a.col2, params + h->o_c2w, a.y2, params + h->o_c2b, B * H2 * H2, F2, F1 * 4, F1 * 4, 1, 1, F1 * 4, F2,
         1, gcol2, P, gy2, P);
    nhwc_to_flat_kernel<<<grid1d((long long)B * H2 * H2 * F2, G), 256, 0, st>>>(a.y2, a.flat, B, H2 * H2, F2, gy2, 0);
}

// ---- conv trunk backward from h->d_flat [G][B][D] (activations of pass 0)
void trunk_backward(maexp_learner *h, cudaStream_t st, bool defer_conv2_wgrad = false, bool only_conv2_wgrad = false, cudaEvent_t after = nullptr) {
    const int G = h->G, B = h->B, C = h->C, F1 = h->F1, F2 = h->F2;
    const long long P = h->P;
    maexp_learner::Acts &a = h->acts[0];
    const int H1 = h->H1, Hp = h->Hp, H2 = h->H2;
    const long long gcol1 = (long long)B * H1 * H1 * C * 4, gy1 = (long long)B * H1 * H1 * F1, gp = (long long)B * Hp * Hp * F1;
    const long long gcol2 = (long long)B * H2 * H2 * F1 * 4, gy2 = (long long)B * H2 * H2 * F2;
    const int M2 = B * H2 * H2;
    if (h->fused_trunk_acts && h->trunk_part) {
        const int V = h->V, warps = 8, K1 = C * 4, D = F2 * H2 * H2;
        const size_t smem = sizeof(float) * ((size_t)F2 * 128 + (size_t)warps * ((D + 3) / 4 * 4 + (C * V * V + 3) / 4 * 4 + K1 * 32 + 96));
        // samples per warp (MAEXP_TRUNK_BWD_SPW, default 2): 16 blocks per agent at B = 256
        static const int spw = getenv("MAEXP_TRUNK_BWD_SPW") ? (atoi(getenv("MAEXP_TRUNK_BWD_SPW")) < 1 ? 1 : atoi(getenv("MAEXP_TRUNK_BWD_SPW"))) : 2;
        int nblk = (B + warps * spw - 1) / (warps * spw);
        if (nblk > kTrunkBwdBlocks) nblk = kTrunkBwdBlocks;
        if (smem <= 200 * 1024) {
            if (!defer_conv2_wgrad) {   // conv2's weight gradient stays a GEMM, off the critical path
                cudaStream_t ws = only_conv2_wgrad ? wgrad_begin_at(st, after) : wgrad_begin(st);
                if (!h->dy2_ready)  // else the fused MLP backward wrote d_y2 next to d_flat
                    nhwc_to_flat_kernel<<<grid1d((long long)B * H2 * H2 * F2, G), 256, 0, ws>>>(h->d_flat, h->d_y2, B, H2 * H2, F2, gy2, 1);
                if (h->col2_ready) cudaStreamWaitEvent(ws, h->ev_col2, 0);
                else
                    im2col_kernel<<<grid1d((long long)B * H2 * H2 * F1 * 4, G), 256, 0, ws>>>(
                        a.p, a.col2, B, F1, H2, H2, (long long)Hp * Hp * F1, 1, (long long)Hp * F1, F1, gp, gcol2);
                h->col2_ready = h->dy2_ready = false;
                gemm(ws, G, h->d_y2, a.col2, h->grads + h->o_c2w, nullptr, F2, F1 * 4, M2, 1, F2, F1 * 4, 1, F1 * 4, 1, gy2, gcol2, P, 0);
                wgrad_end();
                if (only_conv2_wgrad) return;
            }
            auto fn = V == 5 ? trunk_fused_bwd_kernel<5> : V == 7 ? trunk_fused_bwd_kernel<7> : trunk_fused_bwd_kernel<9>;
            cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            fn<<<dim3(nblk, G), warps * 32, smem, st>>>(h->d_flat, h->bwd_obs, h->online, a.arg, h->trunk_part, B, C, F1, F2,
                                                       h->cfg.obs_stride, (long long)B * h->cfg.obs_stride, P, h->o_c2w,
                                                       (long long)B * D, gp);
            stamp(st, 13);
            trunk_bwd_reduce_kernel<<<dim3((K1 * 32 + 96 + 127) / 128, G), 128, 0, st>>>(h->trunk_part, h->grads, nblk, K1, F1, F2, P,
                                                                                       h->o_c1w, h->o_c1b, h->o_c2b);
            return;
        }
    }
    nhwc_to_flat_kernel<<<grid1d((long long)B * H2 * H2 * F2, G), 256, 0, st>>>(h->d_flat, h->d_y2, B, H2 * H2, F2, gy2, 1);
    {
        cudaStream_t ws = wgrad_begin(st);
        if (h->col2_ready) cudaStreamWaitEvent(ws, h->ev_col2, 0);
        else if (h->fused_trunk_acts)  // the fused forward kept the pooled map only
            im2col_kernel<<<grid1d((long long)B * H2 * H2 * F1 * 4, G), 256, 0, ws>>>(
                a.p, a.col2, B, F1, H2, H2, (long long)Hp * Hp * F1, 1, (long long)Hp * F1, F1, gp, gcol2);
        h->col2_ready = h->dy2_ready = false;
        gemm(ws, G, h->d_y2, a.col2, h->grads + h->o_c2w, nullptr, F2, F1 * 4, M2, 1, F2, F1 * 4, 1, F1 * 4, 1, gy2, gcol2, P, 0);
        colsum(ws, G, h->ones, h->d_y2, h->grads + h->o_c2b, M2, F2, gy2, P);
        wgrad_end();
    }
    gemm(st, G, h->d_y2, h->online + h->o_c2w, h->d_col2, nullptr, M2, F1 * 4, F2, F2, 1, F1 * 4, 1, F1 * 4, 1, gy2, P,
         gcol2, 0);
    col2im_kernel<<<grid1d((long long)B * Hp * Hp * F1, G), 256, 0, st>>>(h->d_col2, h->d_p, B, Hp, Hp, F1, gcol2, gp);
    pool_bwd_kernel<<<grid1d((long long)B * H1 * H1 * F1, G), 256, 0, st>>>(h->d_p, a.arg, h->d_y1, B, H1, H1, F1, gy1, gp);
    const int M1 = B * H1 * H1;
    {   // the two conv1 reductions over B*H1*H1 rows run side by side
        cudaStream_t ws = wgrad_begin(st);
        if (h->fused_trunk_acts)
            im2col_kernel<<<grid1d((long long)B * H1 * H1 * C * 4, G), 256, 0, ws>>>(
                h->bwd_obs, a.col1, B, C, H1, H1, h->cfg.obs_stride, (long long)h->V * h->V, h->V, 1,
                (long long)B * h->cfg.obs_stride, gcol1);
        gemm(ws, G, h->d_y1, a.col1, h->grads + h->o_c1w, nullptr, F1, C * 4, M1, 1, F1, C * 4, 1, C * 4, 1, gy1, gcol1, P, 0);
        wgrad_end();
        ws = wgrad_begin(st);
        colsum(ws, G, h->ones, h->d_y1, h->grads + h->o_c1b, M1, F1, gy1, P);
        wgrad_end();
    }
}

// ---- MLP with one hidden ReLU layer on `rows` rows per group (mlp.py:28-34); buffers strided by rows_max
struct MlpIn {
    const float *x;
    long long x_row, x_g;  // row stride and group stride of x
    int rows;
    long long rows_max;
    int din;
};

// all heads of a network that share the input (fc_V / fc_A of the distributional nets), for n <= 3 passes with the same
// shapes (their own input, parameters and outputs), in one fused launch
bool mlp_fused_forward_multi(maexp_learner *h, cudaStream_t st, int n, const float *const *params, const HeadLayout *hds,
                             int nheads, const MlpIn *ins, float *const (*hbufs)[2], float *const (*outs)[2]) {
    const int H = h->H;
    const MlpIn &in = ins[0];
    // acting / evaluation batches: layer 1 of all heads, bias + ReLU and layer 2 in one TMA-fed tcgen05 kernel
    if (g_inference_tc && h->use_tc && h->tc_heads && in.rows >= 512 && n == 1 && nheads >= 1 && nheads <= 2 && H == 128 &&
        in.din >= 32 && in.din % 32 == 0) {
        int outs_n[2] = {hds[0].out, nheads > 1 ? hds[1].out : 0};
        const size_t need = tc_heads_workspace_floats(h->G, nheads, in.din, outs_n);
        if (need > h->tc_heads_ws_floats) {  // first call (or a wider layer): the old panel buffer stays in h->allocs until destroy
            float *ws = nullptr;
            if (lalloc_scratch(h, &ws, need) == 0) { h->tc_heads_ws = ws; h->tc_heads_ws_floats = need; }
        }
        if (need <= h->tc_heads_ws_floats) {
            TcHeadsFwd t = {};
            t.x = in.x; t.x_row = in.x_row; t.x_g = in.x_g;
            t.params = params[0]; t.P = h->P;
            t.nheads = nheads;
            for (int k = 0; k < nheads; ++k) {
                t.w1[k] = hds[k].w1; t.b1[k] = hds[k].b1; t.w2[k] = hds[k].w2; t.b2[k] = hds[k].b2; t.out[k] = hds[k].out;
                t.outp[k] = outs[0][k];
            }
            t.rows = in.rows; t.rows_max = in.rows_max; t.din = in.din; t.H = H;
            t.workspace = h->tc_heads_ws;
            if (g_fused_q.q_out && h->dist && !h->iqn && nheads == 2) {
                t.q_out = g_fused_q.q_out; t.q_g = g_fused_q.q_g; t.A = h->A; t.Z = h->Z;
                t.categorical = g_fused_q.categorical; t.support = h->support;
                t.fused_q_done = &g_fused_q.done;
            }
            if (tc_heads_forward(st, h->G, t)) return true;
            g_fused_q.done = false;
        }
    }
    // large acting batches keep the tensor-core path; the fused kernel serves the update batches and small inference calls
    // and a wide input (the relative-canvas embedders: 7 744 .. 11 520 features on a few rows) needs the split-K GEMM's CTAs
    if (!h->fused_mlp || nheads < 1 || nheads > 2 || n < 1 || n > 3 || H > kMlpTN || (g_inference_tc && in.rows >= 512) ||
        in.din > kMlpMaxIn)
        return false;
    auto aligned16 = [](const void *ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    const long long P = h->P;
    // update batches of >= 128 rows: both layers on the tensor cores (tcgen05, three TF32 terms per product), the hidden
    // activation handed from the first accumulator to the second contraction inside the SM (csrc/tc_mlp.cu)
    // (not IQN: its quantile-Huber gradients amplify the 2e-6 of the three-term TF32 products past the 1e-4 gradient bar)
    if (h->tc_mlp && h->tc_heads && !h->iqn && !g_inference_tc && in.rows >= 128 && H == 128 && in.din >= 32 && in.din % 32 == 0) {
        // TMA-fed kernel (csrc/tc_heads.cu).  A single pass with two heads (what the update's three streams issue) runs one head
        // per CTA -- two grid slices, each with its own 128-column first accumulator: an update batch is only two row tiles per
        // agent, so the kernel's time is one CTA's latency, and a head's 12 K slabs are shorter than both heads' 16
        // (the forward phase of a C51 update ends at 58 instead of 65 us, the update takes 182 instead of 189 us;
        // MAEXP_TC_HEADS_SPLIT=0: both heads in one CTA).  The x slab is loaded and split twice.
        static const bool split_heads = !(getenv("MAEXP_TC_HEADS_SPLIT") && atoi(getenv("MAEXP_TC_HEADS_SPLIT")) == 0);
        if (n == 1 && nheads == 2 && split_heads) {
            bool ok1 = true;
            TcHeadsFwd ent[2] = {};
            for (int k = 0; k < 2 && ok1; ++k) {
                int outs_n[2] = {hds[k].out, 0};
                const size_t need = tc_heads_workspace_floats(h->G, 1, in.din, outs_n);
                int slot = -1;
                for (int j = 0; j < 8 && slot < 0; ++j)
                    if (h->tc_heads_upd[j].key == outs[0][k] || !h->tc_heads_upd[j].key) slot = j;
                if (slot < 0) { ok1 = false; break; }
                auto &e = h->tc_heads_upd[slot];
                if (e.floats < need) {
                    float *ws = nullptr;
                    if (lalloc_scratch(h, &ws, need)) { ok1 = false; break; }
                    e.ws = ws; e.floats = need;
                }
                e.key = outs[0][k];
                TcHeadsFwd &t = ent[k];
                t.x = ins[0].x; t.x_row = ins[0].x_row; t.x_g = ins[0].x_g;
                t.params = params[0]; t.P = P;
                t.nheads = 1;
                t.w1[0] = hds[k].w1; t.b1[0] = hds[k].b1; t.w2[0] = hds[k].w2; t.b2[0] = hds[k].b2; t.out[0] = hds[k].out;
                t.outp[0] = outs[0][k]; t.hbuf[0] = hbufs[0][k];
                t.rows = ins[0].rows; t.rows_max = ins[0].rows_max; t.din = ins[0].din; t.H = H;
                t.workspace = e.ws;
            }
            // inside an update the weight-panel kernels of this pass run on a free side stream behind the update's fork point
            // (beside the pass's trunk): MAEXP_PREP_BRANCH=0 keeps them on the pass's chain
            TcHeadsPrepBranch pb = {}, *prep = nullptr;
            if (ok1 && h->fork_live && h->prep_branch) {
                int pass = 0;
                for (int q = 0; q < 3; ++q)
                    if (outs[0][0] == h->acts[q].out[0]) pass = q;
                pb.stream = h->side[2 + (pass & 1)];
                pb.after = h->ev_fork;
                pb.done = h->ev_prep[pass];
                prep = &pb;
            }
            if (ok1 && tc_heads_forward_multi(st, h->G, ent, 2, prep)) return true;
        }
        bool ok = true;
        TcHeadsFwd passes[3] = {};
        for (int q = 0; q < n && ok; ++q) {
            int outs_n[2] = {hds[0].out, nheads > 1 ? hds[1].out : 0};
            const size_t need = tc_heads_workspace_floats(h->G, nheads, in.din, outs_n);
            int slot = -1;
            for (int k = 0; k < 8 && slot < 0; ++k)
                if (h->tc_heads_upd[k].key == outs[q][0] || !h->tc_heads_upd[k].key) slot = k;
            if (slot < 0) { ok = false; break; }
            auto &e = h->tc_heads_upd[slot];
            if (e.floats < need) {
                float *ws = nullptr;
                if (lalloc_scratch(h, &ws, need)) { ok = false; break; }
                e.ws = ws; e.floats = need;
            }
            e.key = outs[q][0];
            TcHeadsFwd &t = passes[q];
            t.x = ins[q].x; t.x_row = ins[q].x_row; t.x_g = ins[q].x_g;
            t.params = params[q]; t.P = P;
            t.nheads = nheads;
            for (int k = 0; k < nheads; ++k) {
                t.w1[k] = hds[k].w1; t.b1[k] = hds[k].b1; t.w2[k] = hds[k].w2; t.b2[k] = hds[k].b2; t.out[k] = hds[k].out;
                t.outp[k] = outs[q][k]; t.hbuf[k] = hbufs[q][k];
            }
            t.rows = ins[q].rows; t.rows_max = ins[q].rows_max; t.din = ins[q].din; t.H = H;
            t.workspace = e.ws;
        }
        if (ok && tc_heads_forward_multi(st, h->G, passes, n)) return true;  // all passes as grid slices of one launch
    }
    if (h->tc_mlp && !h->iqn && in.rows >= 128 && H <= 128 && (H & 3) == 0) {
        bool ok = true;
        for (int q = 0; q < n && ok; ++q) {
            TcMlpFwd t = {};
            t.x = ins[q].x; t.x_row = ins[q].x_row; t.x_g = ins[q].x_g;
            t.params = params[q]; t.P = P;
            t.nheads = nheads;
            for (int k = 0; k < nheads; ++k) {
                t.w1[k] = hds[k].w1; t.b1[k] = hds[k].b1; t.w2[k] = hds[k].w2; t.b2[k] = hds[k].b2; t.out[k] = hds[k].out;
                t.hbuf[k] = hbufs[q][k]; t.outp[k] = outs[q][k];
            }
            t.rows = ins[q].rows; t.rows_max = ins[q].rows_max; t.din = ins[q].din; t.H = H;
            if (const char *dbg = getenv("MAEXP_TC_MLP_DBG")) t.dbg = atoi(dbg);
            ok = tc_mlp_forward(st, h->G, t);
        }
        if (ok) return true;
    }
    MlpHeads a = {};
    a.nheads = nheads;
    for (int k = 0; k < nheads; ++k) {
        a.w1[k] = hds[k].w1; a.b1[k] = hds[k].b1; a.w2[k] = hds[k].w2; a.b2[k] = hds[k].b2;
        a.out[k] = hds[k].out;
        a.vec_w1[k] = (in.din & 3) == 0 && (P & 3) == 0;
        a.vec_w2[k] = (H & 3) == 0 && (P & 3) == 0;
        for (int q = 0; q < n; ++q) {
            a.vec_w1[k] = a.vec_w1[k] && aligned16(params[q] + hds[k].w1);
            a.vec_w2[k] = a.vec_w2[k] && aligned16(params[q] + hds[k].w2);
            a.hbuf[q][k] = hbufs[q][k]; a.outp[q][k] = outs[q][k];
        }
    }
    for (int q = 0; q < n; ++q) {
        a.x[q] = ins[q].x; a.params[q] = params[q];
        a.vec_x[q] = (in.x_row & 3) == 0 && (in.x_g & 3) == 0 && aligned16(ins[q].x);
    }
    dim3 grid((in.rows + kMlpTM - 1) / kMlpTM, nheads * n, h->G);
    mlp_fused_fwd_kernel<<<grid, 256, 0, st>>>(in.x_row, in.x_g, P, a, in.rows, in.rows_max, in.din, H);
    return true;
}

bool mlp_fused_forward(maexp_learner *h, cudaStream_t st, const float *params, const HeadLayout *hds, int nheads, const MlpIn &in,
                       float *const *hbufs, float *const *outs) {
    if (nheads < 1 || nheads > 2) return false;
    float *const hb[1][2] = {{hbufs[0], nheads > 1 ? hbufs[1] : nullptr}};
    float *const ob[1][2] = {{outs[0], nheads > 1 ? outs[1] : nullptr}};
    return mlp_fused_forward_multi(h, st, 1, &params, hds, nheads, &in, hb, ob);
}

void mlp_forward(maexp_learner *h, cudaStream_t st, const float *params, const HeadLayout &hd, const MlpIn &in, float *hbuf,
                 float *out) {
    const int G = h->G, H = h->H;
    const long long P = h->P;
    if (mlp_fused_forward(h, st, params, &hd, 1, in, &hbuf, &out)) return;
    gemm(st, G, in.x, params + hd.w1, hbuf, params + hd.b1, in.rows, H, in.din, in.x_row, 1, 1, in.din, H, 1, in.x_g, P,
         in.rows_max * H, P, /*relu=*/1);
    gemm(st, G, hbuf, params + hd.w2, out, params + hd.b2, in.rows, hd.out, H, H, 1, 1, H, hd.out, 1, in.rows_max * H, P,
         in.rows_max * hd.out, P);
}

// gradients of one MLP's parameters into h->grads; dx (optional, [rows][din], group stride rows_max*din) = or += dh W1
void mlp_backward(maexp_learner *h, cudaStream_t st, const HeadLayout &hd, const MlpIn &in, const float *dout,
                  const float *hbuf, float *dhbuf, float *dx, int accumulate_dx) {
    const int G = h->G, H = h->H, O = hd.out, rows = in.rows;
    const long long P = h->P, gO = in.rows_max * O, gH = in.rows_max * H;
    {
        cudaStream_t ws = wgrad_begin(st);
        gemm(ws, G, dout, hbuf, h->grads + hd.w2, nullptr, O, H, rows, 1, O, H, 1, H, 1, gO, gH, P, 0);
        colsum(ws, G, h->ones, dout, h->grads + hd.b2, rows, O, gO, P);
        wgrad_end();
    }
    if (h->fused_mlp && H <= kMlpTN && in.din <= kMlpMaxIn) {  // the data-gradient chain of this head in one kernel
        MlpBwdHeads bh = {};
        bh.w1[0] = hd.w1; bh.w2[0] = hd.w2; bh.out[0] = O; bh.dout[0] = dout; bh.hbuf[0] = hbuf; bh.dh[0] = dhbuf;
        const int vec_w1 = (in.din & 3) == 0 && (P & 3) == 0 && (reinterpret_cast<uintptr_t>(h->online + hd.w1) & 15) == 0;
        dim3 grid((rows + kMlpTM - 1) / kMlpTM, dx ? (in.din + kMlpTN - 1) / kMlpTN : 1, G);
        mlp_fused_bwd_kernel<<<grid, 256, 0, st>>>(h->online, P, bh, 1, rows, in.rows_max, in.din, H, dx, vec_w1, nullptr, 1,
                                                   accumulate_dx, 0);
        cudaStream_t ws = wgrad_begin(st);
        gemm(ws, G, dhbuf, in.x, h->grads + hd.w1, nullptr, H, in.din, rows, 1, H, in.x_row, 1, in.din, 1, gH, in.x_g, P, 0);
        colsum(ws, G, h->ones, dhbuf, h->grads + hd.b1, rows, H, gH, P);
        wgrad_end();
        return;
    }
    gemm(st, G, dout, h->online + hd.w2, dhbuf, nullptr, rows, H, O, O, 1, H, 1, H, 1, gO, P, gH, 0, 0, 0, hbuf, gH);
    {
        cudaStream_t ws = wgrad_begin(st);
        gemm(ws, G, dhbuf, in.x, h->grads + hd.w1, nullptr, H, in.din, rows, 1, H, in.x_row, 1, in.din, 1, gH, in.x_g, P, 0);
        colsum(ws, G, h->ones, dhbuf, h->grads + hd.b1, rows, H, gH, P);
        wgrad_end();
    }
    if (dx)
        gemm(st, G, dhbuf, h->online + hd.w1, dx, nullptr, rows, in.din, H, H, 1, in.din, 1, in.din, 1, gH, P,
             in.rows_max * in.din, 0, 0, accumulate_dx);
}

// forward of one (non-IQN) network for all groups
void forward(maexp_learner *h, cudaStream_t st, const float *params, const float *obs, int B, maexp_learner::Acts &a,
             bool inference = false) {
    MlpIn in;
    if (h->has_conv) {
        trunk_forward(h, st, params, obs, B, a, inference);
        in = {a.flat, h->D, (long long)h->B * h->D, B, h->B, h->D};
    } else {
        in = {obs, h->cfg.obs_stride, (long long)B * h->cfg.obs_stride, B, h->B, h->D};
    }
    if (h->heads.size() == 2 && mlp_fused_forward(h, st, params, h->heads.data(), 2, in, a.h, a.out)) return;
    for (size_t k = 0; k < h->heads.size(); ++k) mlp_forward(h, st, params, h->heads[k], in, a.h[k], a.out[k]);
}

// All forward passes of an update (online(s) kept for the backward pass, target(s'), and online(s') for the distributional
// nets' double-DQN action selection) as two launches: every pass is a slice of the grid of the fused trunk kernel and of the
// fused MLP kernel.  Separate streams gave the same overlap in principle, but each pass's CTAs queued behind the others'
// shared-memory footprint and ran at 8 warps per SM.
bool forward_batched(maexp_learner *h, cudaStream_t st, const float *obs, const float *next_obs, int B) {
    if (!h->batched_fwd || !h->fused_mlp || h->heads.empty() || h->heads.size() > 2) return false;
    const int n = h->dist ? 3 : 2;
    const float *params[3] = {h->online, h->target, h->online};
    const float *inputs[3] = {obs, next_obs, next_obs};
    maexp_learner::Acts *acts[3] = {&h->acts[0], &h->acts[2], &h->acts[1]};
    const int train[3] = {1, 0, 0};
    MlpIn ins[3];
    if (h->has_conv) {
        h->fused_trunk_acts = false;
        if (!trunk_forward_multi(h, st, n, params, inputs, acts, train, B)) return false;
        for (int q = 0; q < n; ++q) ins[q] = {acts[q]->flat, h->D, (long long)h->B * h->D, B, h->B, h->D};
    } else {
        for (int q = 0; q < n; ++q) ins[q] = {inputs[q], h->cfg.obs_stride, (long long)B * h->cfg.obs_stride, B, h->B, h->D};
    }
    float *const hb[3][2] = {{acts[0]->h[0], acts[0]->h[1]}, {acts[1]->h[0], acts[1]->h[1]}, {acts[2]->h[0], acts[2]->h[1]}};
    float *const ob[3][2] = {{acts[0]->out[0], acts[0]->out[1]}, {acts[1]->out[0], acts[1]->out[1]}, {acts[2]->out[0], acts[2]->out[1]}};
    if (mlp_fused_forward_multi(h, st, n, params, h->heads.data(), (int)h->heads.size(), ins, hb, ob)) return true;
    // (shapes the fused MLP does not take: the heads pass by pass on this stream)
    for (int q = 0; q < n; ++q)
        for (size_t k = 0; k < h->heads.size(); ++k) mlp_forward(h, st, params[q], h->heads[k], ins[q], acts[q]->h[k], acts[q]->out[k]);
    return true;
}

// backward of the online net through the activations of pass 0; d_out[k] already filled
void backward(maexp_learner *h, cudaStream_t st, const float *obs) {
    maexp_learner::Acts &a = h->acts[0];
    const int B = h->B;
    MlpIn in;
    if (h->has_conv) in = {a.flat, h->D, (long long)B * h->D, B, B, h->D};
    else in = {obs, h->cfg.obs_stride, (long long)B * h->cfg.obs_stride, B, B, h->D};
    wgrad_open(h, st);
    const int nheads = (int)h->heads.size(), H = h->H, G = h->G;
    // C51 / DQN only, like the tensor-core trunk: the 2e-6 of the three-term products is inside their gradient bar
    struct WgradTcScope {
        explicit WgradTcScope(bool on) { g_wgrad_tc = on; }
        ~WgradTcScope() { g_wgrad_tc = false; }
    } wgrad_tc_scope(h->use_tc && h->tc_wgrad && B >= 128 &&
                     (h->cfg.kind == MAEXP_LEARNER_C51 || h->cfg.kind == MAEXP_LEARNER_DQN_CONV_MLP));
    if (h->fused_mlp && nheads <= 2 && H <= kMlpTN && in.din <= kMlpMaxIn) {
        // weight gradients on the side streams around ONE kernel for the data-gradient chain of all heads
        const long long P = h->P, gH = (long long)B * H;
        MlpBwdHeads bh;
        float *dhs[2] = {h->d_h, h->d_h2};
        // critical_first: the data-gradient chain (mlp_fused_bwd -> trunk_fused_bwd -> reduce) is issued BEFORE the weight-gradient
        // kernels that branch off it, which wait for events marked at the points their inputs are ready.  In a captured graph
        // the kernels that become ready together are launched in creation order: with the side work first the chain's kernels
        // started 19 us (after the loss) and 10 us (after mlp_fused_bwd) late (globaltimer stamps, tools/exp_learner_stamps.py)
        const bool to_trunk0 = h->has_conv && h->fused_trunk_acts && h->trunk_part;
        const bool cfirst = h->critical_first && g_wg.h && to_trunk0 && (h->V == 5 || h->V == 7 || h->V == 9);
        if (cfirst) cudaEventRecord(h->ev_mark[0], st);
        for (int k = 0; k < nheads; ++k) {
            const HeadLayout &hd = h->heads[k];
            const int O = hd.out;
            bh.w1[k] = hd.w1; bh.w2[k] = hd.w2; bh.out[k] = O;
            bh.dout[k] = h->d_out[k]; bh.hbuf[k] = h->acts[0].h[k]; bh.dh[k] = dhs[k];
            if (cfirst) continue;
            cudaStream_t ws = wgrad_begin(st);
            gemm(ws, G, h->d_out[k], bh.hbuf[k], h->grads + hd.w2, nullptr, O, H, B, 1, O, H, 1, H, 1, (long long)B * O, gH, P, 0);
            colsum(ws, G, h->ones, h->d_out[k], h->grads + hd.b2, B, O, (long long)B * O, P);
            wgrad_end();
        }
        auto aligned16 = [](const void *ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
        const int vec_w1 = (in.din & 3) == 0 && (P & 3) == 0 && aligned16(h->online + h->heads[0].w1) &&
                           (nheads < 2 || aligned16(h->online + h->heads[1].w1));
        float *dx = h->has_conv ? h->d_flat : nullptr;
        dim3 grid((B + kMlpTM - 1) / kMlpTM, dx ? (in.din + kMlpTN - 1) / kMlpTN : 1, G);
        const bool to_trunk = h->has_conv && h->fused_trunk_acts && h->trunk_part;
        if (to_trunk && !cfirst) {  // conv2's im2col only needs the forward pass: issue it now, on a side stream
            maexp_learner::Acts &a0 = h->acts[0];
            const int Hp = h->Hp, H2 = h->H2, F1 = h->F1;
            cudaStream_t ws = wgrad_begin(st);
            im2col_kernel<<<grid1d((long long)B * H2 * H2 * F1 * 4, G), 256, 0, ws>>>(
                a0.p, a0.col2, B, F1, H2, H2, (long long)Hp * Hp * F1, 1, (long long)Hp * F1, F1, (long long)B * Hp * Hp * F1,
                (long long)B * H2 * H2 * F1 * 4);
            cudaEventRecord(h->ev_col2, ws);
            wgrad_end();
            h->col2_ready = h->dy2_ready = true;
        }
        stamp(st, 11);
        // MAEXP_MLP_BWD_SPLIT=1: dh once per head, then dx from it -- in the one-kernel form every one of the din / 128 column
        // tiles repeats the dout W2 product (2x at din = 256, the flattened trunk output at V = 7)
        // heads in parallel (default for two heads inside an update; MAEXP_MLP_BWD_HEADS_PARALLEL=0: one CTA walks both): at
        // B = 256 the grid is 8 row tiles x 2 column tiles x 8 agents = 128 CTAs of 8 warps, one per SM, so the kernel's time is
        // one CTA's chain of 18 K steps; with a head per CTA the longer chain is 12 steps and an SM holds two CTAs
        const bool head_par = h->mlp_bwd_head_par && dx && nheads == 2 && h->zero_live && !h->mlp_bwd_split;
        if (head_par) {
            const size_t bytes = sizeof(float) * (size_t)G * B * in.din;
            cudaStream_t zs = h->side[3];  // zero dx (and its NHWC copy) on a branch behind the update's fork point: off the chain
            cudaStreamWaitEvent(zs, h->ev_fork, 0);
            cudaMemsetAsync(dx, 0, bytes, zs);
            if (to_trunk) cudaMemsetAsync(h->d_y2, 0, bytes, zs);
            cudaEventRecord(h->ev_zero, zs);
            cudaStreamWaitEvent(st, h->ev_zero, 0);
            mlp_fused_bwd_kernel<<<dim3(grid.x, grid.y * nheads, G), 256, 0, st>>>(h->online, P, bh, nheads, B, B, in.din, H, dx, vec_w1,
                                                                                  to_trunk ? h->d_y2 : nullptr, h->H2 * h->H2, 0, 2);
            if (cfirst) cudaEventRecord(h->ev_mark[2], st);
        } else if (h->mlp_bwd_split && dx && grid.y > 1) {
            mlp_fused_bwd_kernel<<<dim3(grid.x, nheads, G), 256, 0, st>>>(h->online, P, bh, nheads, B, B, in.din, H, dx, vec_w1, nullptr,
                                                                         1, 0, 1);
            if (cfirst) cudaEventRecord(h->ev_mark[2], st);  // dh is ready: the first-layer weight gradients may start
            mlp_dx_kernel<<<grid, 256, 0, st>>>(h->online, P, bh, nheads, B, B, in.din, H, dx, vec_w1, to_trunk ? h->d_y2 : nullptr,
                                                h->H2 * h->H2);
        } else {
            mlp_fused_bwd_kernel<<<grid, 256, 0, st>>>(h->online, P, bh, nheads, B, B, in.din, H, dx, vec_w1,
                                                       to_trunk ? h->d_y2 : nullptr, h->H2 * h->H2, 0, 0);
            if (cfirst) cudaEventRecord(h->ev_mark[2], st);
        }
        stamp(st, 12);
        if (cfirst) {
            cudaEventRecord(h->ev_mark[1], st);
            h->dy2_ready = true;  // the fused MLP backward wrote d_y2 next to d_flat
            stamp(st, 7);
            trunk_backward(h, st, /*defer_conv2_wgrad=*/true);
            stamp(st, 8);
            // ---- the deferred side work, each behind the point of the chain that produced its input
            // MAEXP_WGRAD_GROUPED=1: the weight gradients as two grouped launches without split-K (wgrad_group_kernel).  Measured
            // SLOWER (224 against 198 us per C51 team update): 224 CTAs with K = 256 .. 1 024 each are long serial chains, the
            // split-K GEMMs spread the same work over ~4x more CTAs.  Kept as the deterministic single-pass reference.
            const bool grouped = getenv("MAEXP_WGRAD_GROUPED") != nullptr;
            // when the second-layer weight gradients and the im2col start: 0 = right after the loss (beside mlp_fused_bwd), 1 = after it
            // (measured: 200 us per C51 team update against 220 us -- the early start wins although it slows mlp_fused_bwd down)
            static const int late = getenv("MAEXP_WGRAD_LATE") ? atoi(getenv("MAEXP_WGRAD_LATE")) : 0;
            cudaEvent_t ev_w2 = h->ev_mark[late ? 1 : 0];
            // what reads dh only (first-layer weight gradients, bias gradients) starts when the split chain's dh kernel is done
            // (MAEXP_DH_EARLY=0: after the whole MLP chain, as before the split)
            static const int dh_early = getenv("MAEXP_DH_EARLY") ? atoi(getenv("MAEXP_DH_EARLY")) : 1;
            cudaEvent_t ev_dh = h->ev_mark[dh_early ? 2 : 1];
            if (h->mean_dst) {  // the loss scalar of this update (reads per_sample, written by the loss kernel before ev_mark[0])
                cudaStream_t ws = wgrad_begin_at(st, h->ev_mark[0]);
                mean_kernel<<<G, 256, 0, ws>>>(h->per_sample, h->mean_dst, B);
                wgrad_end();
                h->mean_dst = nullptr;
            }
            if (grouped) {  // after the loss: second-layer weight gradients of all heads, one grouped launch
                WgArgs wa = {};
                for (int k = 0; k < nheads; ++k) {
                    const HeadLayout &hd = h->heads[k];
                    wa.p[wa.n++] = {h->d_out[k], bh.hbuf[k], h->grads + hd.w2, (long long)B * hd.out, gH, hd.out, H, B, hd.out, H, H, 0};
                }
                cudaStream_t ws = wgrad_begin_at(st, ev_w2);
                wgrad_group(ws, G, wa, P);
                wgrad_end();
            } else
            for (int k = 0; k < nheads; ++k) {  // after the loss: second-layer weight / bias gradients
                const HeadLayout &hd = h->heads[k];
                const int O = hd.out;
                cudaStream_t ws = wgrad_begin_at(st, ev_w2);
                gemm(ws, G, h->d_out[k], bh.hbuf[k], h->grads + hd.w2, nullptr, O, H, B, 1, O, H, 1, H, 1, (long long)B * O, gH, P, 0);
                wgrad_end();
            }
            {   // after the loss: conv2's im2col (forward activations only)
                maexp_learner::Acts &a0 = h->acts[0];
                const int Hp = h->Hp, H2 = h->H2, F1 = h->F1;
                cudaStream_t ws = wgrad_begin_at(st, ev_w2);
                im2col_kernel<<<grid1d((long long)B * H2 * H2 * F1 * 4, G), 256, 0, ws>>>(
                    a0.p, a0.col2, B, F1, H2, H2, (long long)Hp * Hp * F1, 1, (long long)Hp * F1, F1, (long long)B * Hp * Hp * F1,
                    (long long)B * H2 * H2 * F1 * 4);
                cudaEventRecord(h->ev_col2, ws);
                wgrad_end();
                h->col2_ready = true;
            }
            if (grouped) {  // after mlp_fused_bwd (d_h, d_y2) and the im2col: first-layer and conv2 weight gradients, one launch
                WgArgs wa = {};
                for (int k = 0; k < nheads; ++k) {
                    const HeadLayout &hd = h->heads[k];
                    wa.p[wa.n++] = {dhs[k], in.x, h->grads + hd.w1, gH, in.x_g, H, in.din, B, H, (int)in.x_row, in.din, 0};
                }
                const int F1 = h->F1, F2 = h->F2, H2 = h->H2, M2 = B * H2 * H2;
                wa.p[wa.n++] = {h->d_y2, h->acts[0].col2, h->grads + h->o_c2w, (long long)M2 * F2, (long long)M2 * F1 * 4, F2, F1 * 4, M2,
                                F2, F1 * 4, F1 * 4, 0};
                cudaStream_t ws = wgrad_begin_at(st, h->ev_mark[1]);
                cudaStreamWaitEvent(ws, h->ev_col2, 0);
                h->col2_ready = h->dy2_ready = false;
                wgrad_group(ws, G, wa, P);
                wgrad_end();
            } else
            for (int k = 0; k < nheads; ++k) {  // after dh: first-layer weight / bias gradients
                const HeadLayout &hd = h->heads[k];
                cudaStream_t ws = wgrad_begin_at(st, ev_dh);
                gemm(ws, G, dhs[k], in.x, h->grads + hd.w1, nullptr, H, in.din, B, 1, H, in.x_row, 1, in.din, 1, gH, in.x_g, P, 0);
                wgrad_end();
            }
            {   // after mlp_fused_bwd: every bias gradient of the heads in one launch
                BiasGradArgs ba = {};
                int maxc = 0;
                for (int k = 0; k < nheads; ++k) {
                    const HeadLayout &hd = h->heads[k];
                    ba.src[ba.n] = h->d_out[k]; ba.src_g[ba.n] = (long long)B * hd.out; ba.cols[ba.n] = hd.out; ba.dst_off[ba.n] = hd.b2; ++ba.n;
                    ba.src[ba.n] = dhs[k]; ba.src_g[ba.n] = gH; ba.cols[ba.n] = H; ba.dst_off[ba.n] = hd.b1; ++ba.n;
                    maxc = hd.out > maxc ? hd.out : maxc;
                    maxc = H > maxc ? H : maxc;
                }
                cudaStream_t ws = wgrad_begin_at(st, ev_dh);
                bias_grads_kernel<<<dim3((maxc + 31) / 32, ba.n, G), 256, 0, ws>>>(ba, h->grads, P, B);
                wgrad_end();
            }
            if (!grouped) trunk_backward(h, st, false, /*only_conv2_wgrad=*/true, h->ev_mark[1]);  // after mlp_fused_bwd (d_y2) and the im2col
            wgrad_close();
            return;
        }
        for (int k = 0; k < nheads; ++k) {
            const HeadLayout &hd = h->heads[k];
            cudaStream_t ws = wgrad_begin(st);
            gemm(ws, G, dhs[k], in.x, h->grads + hd.w1, nullptr, H, in.din, B, 1, H, in.x_row, 1, in.din, 1, gH, in.x_g, P, 0);
            colsum(ws, G, h->ones, dhs[k], h->grads + hd.b1, B, H, gH, P);
            wgrad_end();
        }
    } else {
        for (size_t k = 0; k < h->heads.size(); ++k)
            mlp_backward(h, st, h->heads[k], in, h->d_out[k], a.h[k], k & 1 ? h->d_h2 : h->d_h, h->has_conv ? h->d_flat : nullptr,
                         k > 0);
    }
    stamp(st, 7);
    if (h->has_conv) trunk_backward(h, st);
    stamp(st, 8);
    wgrad_close();
}

// ---- IQN (customs/iqn.py:88-139): quantiles of `nq` taus per sample; taus per sample [G][B][nq] or shared [G][nq]
// ---- DA3 (customs/da3.py:22-96, vit.py): patch embedding, saliency token, pre-LN blocks, final LN, head
inline void ln_fwd(maexp_learner *h, cudaStream_t st, const float *params, const float *x, long long x_row, long long gX,
                   long long wo, long long bo, float *y, float *m, float *r, int rows, long long rows_max) {
    layernorm_fwd_kernel<<<dim3((rows * 32 + 255) / 256, h->G), 256, 0, st>>>(x, x_row, params + wo, params + bo, y, m, r, rows,
                                                                              h->E_, gX, rows_max * h->E_, rows_max, h->P);
}

// one stack: src image [g][b*sb + c*sc + x*Win + y] -> s.a[pass].nf = LN(last block output)[:, 0]
void vit_forward(maexp_learner *h, cudaStream_t st, const float *params, maexp_learner::VitStack &s, const float *src,
                 long long sb, long long sc, long long gSrc, int B, int pass, const float *tok0_ext) {
    maexp_learner::Da3Acts &a = s.a[pass];
    const int G = h->G, E = h->E_, Tn = s.Tn, Fm = h->Fm, K = s.K;
    const long long P = h->P, Bmax = h->B, Rm = Bmax * Tn, Rt = Bmax * (Tn - 1);
    const int rows = B * Tn, rowst = B * (Tn - 1);
    vit_patch_gather_kernel<<<grid1d((long long)rowst * K, G), 256, 0, st>>>(src, a.patches, B, s.Cin, s.Ht, s.Wt, s.Win, s.ps, sb,
                                                                             sc, gSrc, Rt * K);
    gemm(st, G, a.patches, params + s.o_pew, a.tok, params + s.o_peb, rowst, E, K, K, 1, 1, K, E, 1, Rt * K, P, Rt * E, P);
    da3_tokens_kernel<<<grid1d((long long)rows * E, G), 256, 0, st>>>(a.tok, tok0_ext ? nullptr : params + s.o_sal, tok0_ext,
                                                                      params + s.o_pos, a.x0, B, Tn, E, Rt * E, Rm * E, Bmax * E, P);
    const float *x = a.x0;
    const float scale = 1.0f / sqrtf((float)(E / h->heads_));
    for (int l = 0; l < h->L_; ++l) {
        const maexp_learner::BlockLayout &w = s.blocks[l];
        maexp_learner::Da3Block &k = a.blk[l];
        ln_fwd(h, st, params, x, E, Rm * E, w.n1w, w.n1b, k.n1, k.m1, k.r1, rows, Rm);
        gemm(st, G, k.n1, params + w.qkvw, k.qkv, nullptr, rows, 3 * E, E, E, 1, 1, E, 3 * E, 1, Rm * E, P, Rm * 3 * E, 0);
        if (g_inference_tc && B > kAttnMapsBatch) {  // acting batches: no attention maps kept (maexp_learner_attention refuses)
            cudaFuncSetAttribute(da3_attn_infer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
            da3_attn_infer_kernel<<<dim3(h->heads_, B, G), (Tn + 31) / 32 * 32, sizeof(float) * Tn * (Tn + 1), st>>>(
                k.qkv, k.ao, B, Tn, E, h->heads_, scale, Rm * 3 * E, Rm * E, h->attn_shrink);
        } else {
            da3_attn_fwd_kernel<<<dim3(h->heads_, B, G), 128, 0, st>>>(k.qkv, k.probs, k.ao, B, Tn, E, h->heads_, scale, Rm * 3 * E,
                                                                       Bmax * h->heads_ * Tn * Tn, Rm * E, h->attn_shrink);
        }
        cudaMemcpyAsync(k.x1, x, sizeof(float) * (size_t)G * Rm * E, cudaMemcpyDeviceToDevice, st);
        gemm(st, G, k.ao, params + w.projw, k.x1, params + w.projb, rows, E, E, E, 1, 1, E, E, 1, Rm * E, P, Rm * E, P, 0,
             /*accumulate=*/1);
        ln_fwd(h, st, params, k.x1, E, Rm * E, w.n2w, w.n2b, k.n2, k.m2, k.r2, rows, Rm);
        gemm(st, G, k.n2, params + w.f1w, k.f1, params + w.f1b, rows, Fm, E, E, 1, 1, E, Fm, 1, Rm * E, P, Rm * Fm, P);
        gelu_fwd_kernel<<<grid1d((long long)G * Rm * Fm, 1), 256, 0, st>>>(k.f1, k.g1, (long long)G * Rm * Fm);
        cudaMemcpyAsync(k.x2, k.x1, sizeof(float) * (size_t)G * Rm * E, cudaMemcpyDeviceToDevice, st);
        gemm(st, G, k.g1, params + w.f2w, k.x2, params + w.f2b, rows, E, Fm, Fm, 1, 1, Fm, E, 1, Rm * Fm, P, Rm * E, P, 0, 1);
        x = k.x2;
    }
    // final LayerNorm on the saliency token only (da3.py:59-60,90-91)
    layernorm_fwd_kernel<<<dim3((B * 32 + 255) / 256, G), 256, 0, st>>>(x, (long long)Tn * E, params + h->o_nfw, params + h->o_nfb,
                                                                        a.nf, a.mf, a.rf, B, E, Rm * E, Bmax * E, Bmax, P);
}

// gradients of one stack from h->da_dnf = d LN(token 0); norm_acc: the shared final norm was already written by
// another stack (DA6), so add.  d_tok0_ext != nullptr: token 0 was a per-sample input, return its gradient [G][B][E].
void vit_backward(maexp_learner *h, cudaStream_t st, maexp_learner::VitStack &s, float *d_tok0_ext, int norm_acc) {
    maexp_learner::Da3Acts &a = s.a[0];
    const int G = h->G, B = h->B, E = h->E_, Tn = s.Tn, Fm = h->Fm, K = s.K;
    const long long P = h->P, Rm = (long long)B * Tn, Rt = (long long)B * (Tn - 1);
    const int rows = B * Tn, rowst = B * (Tn - 1);
    float *grads = h->grads;
    const float *W = h->online;
    const float scale = 1.0f / sqrtf((float)(E / h->heads_));
    auto ln_bwd = [&](const float *dy, const float *x, long long x_row, long long gX, long long wo, const float *m,
                      const float *r, float *dx, long long dx_row, long long gDx, int acc, int nrows, long long nrows_max) {
        layernorm_bwd_kernel<<<dim3((nrows * 32 + 255) / 256, G), 256, 0, st>>>(dy, x, x_row, W + wo, m, r, dx, dx_row, acc,
                                                                                h->da_gx, nrows, E, nrows_max * E, gX, gDx,
                                                                                nrows_max, P);
    };
    const float *x_last = h->L_ ? a.blk[h->L_ - 1].x2 : a.x0;
    ln_bwd(h->da_dnf, x_last, (long long)Tn * E, Rm * E, h->o_nfw, a.mf, a.rf, h->da_dtok0, E, (long long)B * E, 0, B, B);
    colsum(st, G, h->ones, h->da_gx, grads + h->o_nfw, B, E, (long long)B * E, P, norm_acc);
    colsum(st, G, h->ones, h->da_dnf, grads + h->o_nfb, B, E, (long long)B * E, P, norm_acc);
    da3_scatter_token0_kernel<<<grid1d((long long)rows * E, G), 256, 0, st>>>(h->da_dtok0, h->da_dx, B, Tn, E, (long long)B * E,
                                                                              Rm * E);
    for (int l = h->L_ - 1; l >= 0; --l) {
        const maexp_learner::BlockLayout &w = s.blocks[l];
        maexp_learner::Da3Block &k = a.blk[l];
        const float *x_in = l ? a.blk[l - 1].x2 : a.x0;
        // mlp branch: x2 = x1 + fc2(gelu(fc1(LN2(x1))))
        gemm(st, G, h->da_dx, k.g1, grads + w.f2w, nullptr, E, Fm, rows, 1, E, Fm, 1, Fm, 1, Rm * E, Rm * Fm, P, 0);
        colsum(st, G, h->ones, h->da_dx, grads + w.f2b, rows, E, Rm * E, P);
        gemm(st, G, h->da_dx, W + w.f2w, h->da_dg1, nullptr, rows, Fm, E, E, 1, Fm, 1, Fm, 1, Rm * E, P, Rm * Fm, 0);
        gelu_bwd_kernel<<<grid1d((long long)G * Rm * Fm, 1), 256, 0, st>>>(k.f1, h->da_dg1, h->da_df1, (long long)G * Rm * Fm);
        gemm(st, G, h->da_df1, k.n2, grads + w.f1w, nullptr, Fm, E, rows, 1, Fm, E, 1, E, 1, Rm * Fm, Rm * E, P, 0);
        colsum(st, G, h->ones, h->da_df1, grads + w.f1b, rows, Fm, Rm * Fm, P);
        gemm(st, G, h->da_df1, W + w.f1w, h->da_dn, nullptr, rows, E, Fm, Fm, 1, E, 1, E, 1, Rm * Fm, P, Rm * E, 0);
        ln_bwd(h->da_dn, k.x1, E, Rm * E, w.n2w, k.m2, k.r2, h->da_dx, E, Rm * E, /*accumulate=*/1, rows, Rm);
        colsum(st, G, h->ones, h->da_gx, grads + w.n2w, rows, E, Rm * E, P);
        colsum(st, G, h->ones, h->da_dn, grads + w.n2b, rows, E, Rm * E, P);
        // attention branch: x1 = x_in + proj(attn(LN1(x_in)))        (da_dx now holds d x1)
        gemm(st, G, h->da_dx, k.ao, grads + w.projw, nullptr, E, E, rows, 1, E, E, 1, E, 1, Rm * E, Rm * E, P, 0);
        colsum(st, G, h->ones, h->da_dx, grads + w.projb, rows, E, Rm * E, P);
        gemm(st, G, h->da_dx, W + w.projw, h->da_dao, nullptr, rows, E, E, E, 1, E, 1, E, 1, Rm * E, P, Rm * E, 0);
        const size_t smem = sizeof(float) * (size_t)(4 * Tn * (E / h->heads_) + Tn * Tn);
        da3_attn_bwd_kernel<<<dim3(h->heads_, B, G), 128, smem, st>>>(k.qkv, k.probs, h->da_dao, h->da_dqkv, B, Tn, E, h->heads_,
                                                                      scale, Rm * 3 * E, (long long)B * h->heads_ * Tn * Tn, Rm * E,
                                                                      h->attn_shrink);
        gemm(st, G, h->da_dqkv, k.n1, grads + w.qkvw, nullptr, 3 * E, E, rows, 1, 3 * E, E, 1, E, 1, Rm * 3 * E, Rm * E, P, 0);
        gemm(st, G, h->da_dqkv, W + w.qkvw, h->da_dn, nullptr, rows, E, 3 * E, 3 * E, 1, E, 1, E, 1, Rm * 3 * E, P, Rm * E, 0);
        ln_bwd(h->da_dn, x_in, E, Rm * E, w.n1w, k.m1, k.r1, h->da_dx, E, Rm * E, 1, rows, Rm);
        colsum(st, G, h->ones, h->da_gx, grads + w.n1w, rows, E, Rm * E, P);
        colsum(st, G, h->ones, h->da_dn, grads + w.n1b, rows, E, Rm * E, P);
    }
    // embeddings: pos_embed, token 0 (parameter or per-sample input), patch projection (da_dx = d x0)
    colsum(st, G, h->ones, h->da_dx, grads + s.o_pos, B, Tn * E, Rm * E, P);
    if (d_tok0_ext)  // rows (g, b, token 0) of da_dx; B == batch_size in the backward pass, so (g*B + b) * Tn * E is uniform
        cudaMemcpy2DAsync(d_tok0_ext, sizeof(float) * E, h->da_dx, sizeof(float) * (size_t)Tn * E, sizeof(float) * E,
                          (size_t)G * B, cudaMemcpyDeviceToDevice, st);
    else
        gemm(st, G, h->ones, h->da_dx, grads + s.o_sal, nullptr, 1, E, B, 0, 0, (long long)Tn * E, 1, E, 1, 0, Rm * E, P, 0);
    da3_token_rows_kernel<<<grid1d((long long)rowst * E, G), 256, 0, st>>>(h->da_dx, h->da_dtok, B, Tn, E, Rm * E, Rt * E);
    gemm(st, G, h->da_dtok, a.patches, grads + s.o_pew, nullptr, E, K, rowst, 1, E, K, 1, K, 1, Rt * E, Rt * K, P, 0);
    colsum(st, G, h->ones, h->da_dtok, grads + s.o_peb, rowst, E, Rt * E, P);
}

// DA3 / MergedDA3's local stack on the batch's V x V crops, optionally followed by the drl_head (da3.py:64-96)
int decode_canvas(maexp_learner *h, cudaStream_t st, const float *obs, const int8_t *pos, int B, int pass, int only_area);

int da3_forward(maexp_learner *h, cudaStream_t st, const float *params, const float *obs, int B, int pass,
                bool with_head = true, const int8_t *pos = nullptr) {
    const long long P = h->P, Bmax = h->B;
    const int E = h->E_;
    if (h->relative_vit) {  // relative view: the transformer reads the decoded [C, X, Y] canvas (da3.py:98-109, da3_iqn.py:194-210)
        const long long XY = (long long)h->X * h->Y;
        if (decode_canvas(h, st, obs, pos, B, pass, 0)) return -1;
        const long long Ct = h->C + h->Cx;
        vit_forward(h, st, params, h->vloc, h->canvas[pass], Ct * XY, XY, (long long)B * Ct * XY, B, pass, nullptr);
    } else {
        vit_forward(h, st, params, h->vloc, obs, h->cfg.obs_stride, (long long)h->V * h->V, (long long)B * h->cfg.obs_stride, B,
                    pass, nullptr);
    }
    if (with_head)
        gemm(st, h->G, h->vloc.a[pass].nf, params + h->o_hw, h->vloc.a[pass].q, params + h->o_hb, B, h->A, E, E, 1, 1, E, h->A, 1,
             Bmax * E, P, Bmax * h->A, P);
    return 0;
}

// drl_head gradients and d LN(token 0) into h->da_dnf
void head_backward(maexp_learner *h, cudaStream_t st, const float *nf) {
    const int G = h->G, B = h->B, E = h->E_, A = h->A;
    const long long P = h->P;
    const float *dq = h->d_out[0];
    gemm(st, G, dq, nf, h->grads + h->o_hw, nullptr, A, E, B, 1, A, E, 1, E, 1, (long long)B * A, (long long)B * E, P, 0);
    colsum(st, G, h->ones, dq, h->grads + h->o_hb, B, A, (long long)B * A, P);
    gemm(st, G, dq, h->online + h->o_hw, h->da_dnf, nullptr, B, E, A, A, 1, E, 1, E, 1, (long long)B * A, P, (long long)B * E, 0);
}

void da3_backward(maexp_learner *h, cudaStream_t st, bool with_head = true) {
    if (with_head) head_backward(h, st, h->vloc.a[0].nf);  // the merged model fills da_dnf from its own concatenated head
    vit_backward(h, st, h->vloc, nullptr, 0);
}

int decode_canvas(maexp_learner *h, cudaStream_t st, const float *obs, const int8_t *pos, int B, int pass, int only_area);

// DA6 (customs/da6.py:98-126): relative stack on the area canvas -> LN -> fc1 -> token 0 of the local stack -> LN -> drl_head
int da6_forward(maexp_learner *h, cudaStream_t st, const float *params, const float *obs, const int8_t *pos, int B, int pass,
                bool with_head = true) {
    const long long P = h->P, Bmax = h->B, XY = (long long)h->X * h->Y;
    const int E = h->E_, G = h->G;
    if (decode_canvas(h, st, obs, pos, B, pass, 1)) return -1;
    const long long Ct = 1 + h->Cx;
    vit_forward(h, st, params, h->vrel, h->canvas[pass], Ct * XY, XY, (long long)B * Ct * XY, B, pass, nullptr);
    gemm(st, G, h->vrel.a[pass].nf, params + h->o_f1w, h->sal2[pass], params + h->o_f1b, B, E, E, E, 1, 1, E, E, 1, Bmax * E, P,
         Bmax * E, P);
    vit_forward(h, st, params, h->vloc, obs, h->cfg.obs_stride, (long long)h->V * h->V, (long long)B * h->cfg.obs_stride, B, pass,
                h->sal2[pass]);
    if (with_head)
        gemm(st, G, h->vloc.a[pass].nf, params + h->o_hw, h->vloc.a[pass].q, params + h->o_hb, B, h->A, E, E, 1, 1, E, h->A, 1,
             Bmax * E, P, Bmax * h->A, P);
    return 0;
}

void da6_backward(maexp_learner *h, cudaStream_t st, bool with_head = true) {
    const int G = h->G, B = h->B, E = h->E_;
    const long long P = h->P, BE = (long long)B * E;
    if (with_head) head_backward(h, st, h->vloc.a[0].nf);  // else da_dnf already holds d LN(token 0) (IQN head)
    vit_backward(h, st, h->vloc, h->d_sal2, 0);
    // fc1: sal2 = nf_rel W^T + b
    gemm(st, G, h->d_sal2, h->vrel.a[0].nf, h->grads + h->o_f1w, nullptr, E, E, B, 1, E, E, 1, E, 1, BE, BE, P, 0);
    colsum(st, G, h->ones, h->d_sal2, h->grads + h->o_f1b, B, E, BE, P);
    gemm(st, G, h->d_sal2, h->online + h->o_f1w, h->da_dnf, nullptr, B, E, E, E, 1, E, 1, E, 1, BE, P, BE, 0);
    vit_backward(h, st, h->vrel, nullptr, /*norm_acc=*/1);  // the final norm is shared by both stages (da6.py:94,110,121)
}

// ---- generalized conv trunk: src[g][b*sb + c*sc + y*sy + x*sx] -> flat slice [b*flat_row + flat_off + ...]
void gtrunk_forward(maexp_learner *h, cudaStream_t st, const float *params, maexp_learner::GTrunk &t, int pass, const float *src,
                    long long sb, long long sc, long long sy, long long sx, long long gSrc, int B, float *flat,
                    long long flat_row, long long flat_off, long long gFlat) {
    const int G = h->G, F1 = h->F1, F2 = h->F2;
    const long long P = h->P, Bmax = h->B;
    auto &a = t.a[pass];
    const long long n1 = (long long)t.H1 * t.W1, np = (long long)t.Hp * t.Wp, n2 = (long long)t.H2 * t.W2;
    const long long gcol1 = Bmax * n1 * t.Cin * 4, gy1 = Bmax * n1 * F1, gp = Bmax * np * F1, gcol2 = Bmax * n2 * F1 * 4,
                    gy2 = Bmax * n2 * F2;
    im2col_kernel<<<grid1d((long long)B * n1 * t.Cin * 4, G), 256, 0, st>>>(src, a.col1, B, t.Cin, t.H1, t.W1, sb, sc, sy, sx, gSrc,
                                                                            gcol1);
    gemm(st, G, a.col1, params + t.c1w, a.y1, params + t.c1b, (int)(B * n1), F1, t.Cin * 4, t.Cin * 4, 1, 1, t.Cin * 4, F1, 1,
         gcol1, P, gy1, P);
    pool_fwd_kernel<<<grid1d((long long)B * np * F1, G), 256, 0, st>>>(a.y1, a.p, a.arg, B, t.H1, t.W1, F1, gy1, gp);
    im2col_kernel<<<grid1d((long long)B * n2 * F1 * 4, G), 256, 0, st>>>(a.p, a.col2, B, F1, t.H2, t.W2, np * F1, 1,
                                                                         (long long)t.Wp * F1, F1, gp, gcol2);
    gemm(st, G, a.col2, params + t.c2w, a.y2, params + t.c2b, (int)(B * n2), F2, F1 * 4, F1 * 4, 1, 1, F1 * 4, F2, 1, gcol2, P,
         gy2, P);
    nhwc_flat_strided_kernel<<<grid1d((long long)B * n2 * F2, G), 256, 0, st>>>(a.y2, flat, B, (int)n2, F2, gy2, flat_row, flat_off,
                                                                                gFlat, 0);
}

void gtrunk_backward(maexp_learner *h, cudaStream_t st, maexp_learner::GTrunk &t, const float *dflat, long long dflat_row,
                     long long dflat_off, long long gDflat) {
    const int G = h->G, B = h->B, F1 = h->F1, F2 = h->F2;
    const long long P = h->P;
    auto &a = t.a[0];
    const long long n1 = (long long)t.H1 * t.W1, np = (long long)t.Hp * t.Wp, n2 = (long long)t.H2 * t.W2;
    const long long gcol1 = B * n1 * t.Cin * 4, gy1 = B * n1 * F1, gp = B * np * F1, gcol2 = B * n2 * F1 * 4, gy2 = B * n2 * F2;
    nhwc_flat_strided_kernel<<<grid1d((long long)B * n2 * F2, G), 256, 0, st>>>(dflat, t.d_y2, B, (int)n2, F2, gy2, dflat_row,
                                                                                dflat_off, gDflat, 1);
    const int M2 = (int)(B * n2), M1 = (int)(B * n1);
    gemm(st, G, t.d_y2, a.col2, h->grads + t.c2w, nullptr, F2, F1 * 4, M2, 1, F2, F1 * 4, 1, F1 * 4, 1, gy2, gcol2, P, 0);
    colsum(st, G, h->ones, t.d_y2, h->grads + t.c2b, M2, F2, gy2, P);
    gemm(st, G, t.d_y2, h->online + t.c2w, t.d_col2, nullptr, M2, F1 * 4, F2, F2, 1, F1 * 4, 1, F1 * 4, 1, gy2, P, gcol2, 0);
    col2im_kernel<<<grid1d((long long)B * np * F1, G), 256, 0, st>>>(t.d_col2, t.d_p, B, t.Hp, t.Wp, F1, gcol2, gp);
    pool_bwd_kernel<<<grid1d((long long)B * n1 * F1, G), 256, 0, st>>>(t.d_p, a.arg, t.d_y1, B, t.H1, t.W1, F1, gy1, gp);
    gemm(st, G, t.d_y1, a.col1, h->grads + t.c1w, nullptr, F1, t.Cin * 4, M1, 1, F1, t.Cin * 4, 1, t.Cin * 4, 1, gy1, gcol1, P, 0);
    colsum(st, G, h->ones, t.d_y1, h->grads + t.c1b, M1, F1, gy1, P);
}

// ---- relative / merged models (conv_mlp.py:56-108, da3.py:99-225): the relative branch reads the decoded canvas
int decode_canvas(maexp_learner *h, cudaStream_t st, const float *obs, const int8_t *pos, int B, int pass, int only_area) {
    if (!pos) return fail("maexp_learner: relative / merged models need maexp_learner_bind_positions");
    const long long rows = (long long)h->G * B, XY = (long long)h->X * h->Y;
    const int base = only_area ? 1 : h->C, Ct = base + h->Cx;
    if (decode_relative_launch(obs, pos, h->canvas[pass], rows, h->C, h->V, h->X, h->Y, h->cfg.obs_stride, only_area, Ct * XY,
                               only_area ? h->los : nullptr, st))
        return -1;
    if (h->Cx) {  // state["objects"], decode_agents_channel(state), state["destination"] behind the decoded channels
        const maexp_learner::CanvasIn &in = h->cin[(pos == h->next_pos && pos != h->pos) ? 1 : 0];
        if ((h->ex_o && !in.obj) || (h->ex_a && !in.agents) || (h->ex_d && !in.dest))
            return fail("maexp_learner: objects_channel / agents_channel / destination_channel need maexp_learner_bind_canvas_inputs");
        const long long total = rows * h->Cx * XY;
        canvas_extra_kernel<<<grid1d(total, 1, 256, 148 * 16), 256, 0, st>>>(h->canvas[pass], rows, base, h->ex_o, h->ex_a, h->ex_d,
                                                                             h->X, h->Y, in.obj, in.agents, in.dest);
    }
    return 0;
}

int rel_forward(maexp_learner *h, cudaStream_t st, const float *params, const float *obs, const int8_t *pos, int B, int pass) {
    const int V = h->V, C = h->C, X = h->X, Y = h->Y;
    const long long Bmax = h->B;
    if (h->relative_conv) {  // ConvMLP on the full canvas [C, X, Y] (conv_mlp.py:56-66)
        if (decode_canvas(h, st, obs, pos, B, pass, 0)) return -1;
        const long long Ct = C + h->Cx;
        gtrunk_forward(h, st, params, h->gt_rel, pass, h->canvas[pass], Ct * X * Y, (long long)X * Y, Y, 1,
                       (long long)B * Ct * X * Y, B, h->feat[pass], h->Drel, 0, Bmax * h->Drel);
    } else {  // merged: conv(area canvas + 1) ++ conv(local)  (conv_mlp.py:87-108)
        if (decode_canvas(h, st, obs, pos, B, pass, 1)) return -1;
        const long long Dt = h->Drel + h->Dloc;
        gtrunk_forward(h, st, params, h->gt_rel, pass, h->canvas[pass], (long long)X * Y, (long long)X * Y, Y, 1,
                       (long long)B * X * Y, B, h->merged_da3 ? h->rel_flat[pass] : h->feat[pass],
                       h->merged_da3 ? h->Drel : Dt, 0, Bmax * (h->merged_da3 ? h->Drel : Dt));
        if (!h->merged_da3)
            gtrunk_forward(h, st, params, h->gt_loc, pass, obs, h->cfg.obs_stride, (long long)V * V, V, 1,
                           (long long)B * h->cfg.obs_stride, B, h->feat[pass], Dt, h->Drel, Bmax * Dt);
    }
    return 0;
}

int merged_forward(maexp_learner *h, cudaStream_t st, const float *params, const float *obs, const int8_t *pos, int B, int pass,
                   float **q_out, bool with_head = true) {
    const long long Bmax = h->B, P = h->P;
    if (rel_forward(h, st, params, obs, pos, B, pass)) return -1;
    if (!h->merged_da3) {
        const int D = h->D;
        const MlpIn in = {h->feat[pass], D, Bmax * D, B, Bmax, D};
        mlp_forward(h, st, params, h->heads[0], in, h->acts[pass].h[0], h->acts[pass].out[0]);
        *q_out = h->acts[pass].out[0];
        return 0;
    }
    // MergedDA3: [relative_state_embedder(conv(area)), LN(local ViT)[:, 0]] -> drl_head   (da3.py:187-215)
    const int E = h->E_, Dr = h->Drel, Dt = Dr + E;
    const MlpIn rin = {h->rel_flat[pass], Dr, Bmax * Dr, B, Bmax, Dr};
    mlp_forward(h, st, params, h->rse, rin, h->rse_h[pass], h->rse_out[pass]);
    da3_forward(h, st, params, obs, B, pass, /*with_head=*/false);
    for (int g = 0; g < h->G; ++g) {  // concatenate per group (2-D copies are graph-capturable)
        cudaMemcpy2DAsync(h->feat[pass] + (size_t)g * Bmax * Dt, sizeof(float) * Dt, h->rse_out[pass] + (size_t)g * Bmax * Dr,
                          sizeof(float) * Dr, sizeof(float) * Dr, B, cudaMemcpyDeviceToDevice, st);
        cudaMemcpy2DAsync(h->feat[pass] + (size_t)g * Bmax * Dt + Dr, sizeof(float) * Dt, h->vloc.a[pass].nf + (size_t)g * Bmax * E,
                          sizeof(float) * E, sizeof(float) * E, B, cudaMemcpyDeviceToDevice, st);
    }
    if (!with_head) return 0;  // MergedDA3_IQN: h->feat[pass] is the state embedding of the IQN head
    gemm(st, h->G, h->feat[pass], params + h->o_hw, h->vloc.a[pass].q, params + h->o_hb, B, h->A, Dt, Dt, 1, 1, Dt, h->A, 1, Bmax * Dt, P,
         Bmax * h->A, P);
    *q_out = h->vloc.a[pass].q;
    return 0;
}

void merged_backward(maexp_learner *h, cudaStream_t st, bool with_head = true) {
    const int G = h->G, B = h->B, A = h->A;
    const long long P = h->P;
    if (!h->merged_da3) {
        const int D = h->D;
        const MlpIn in = {h->feat[0], D, (long long)B * D, B, B, D};
        mlp_backward(h, st, h->heads[0], in, h->d_out[0], h->acts[0].h[0], h->d_h, h->d_feat, 0);
        if (h->relative_conv) {
            gtrunk_backward(h, st, h->gt_rel, h->d_feat, D, 0, (long long)B * D);
        } else {
            gtrunk_backward(h, st, h->gt_rel, h->d_feat, D, 0, (long long)B * D);
            gtrunk_backward(h, st, h->gt_loc, h->d_feat, D, h->Drel, (long long)B * D);
        }
        return;
    }
    const int E = h->E_, Dr = h->Drel, Dt = Dr + E;
    if (with_head) {  // else d_feat already holds the gradient of the concatenated embedding (IQN head)
        const float *dq = h->d_out[0];
        gemm(st, G, dq, h->feat[0], h->grads + h->o_hw, nullptr, A, Dt, B, 1, A, Dt, 1, Dt, 1, (long long)B * A, (long long)B * Dt, P, 0);
        colsum(st, G, h->ones, dq, h->grads + h->o_hb, B, A, (long long)B * A, P);
        gemm(st, G, dq, h->online + h->o_hw, h->d_feat, nullptr, B, Dt, A, A, 1, Dt, 1, Dt, 1, (long long)B * A, P, (long long)B * Dt, 0);
    }
    for (int g = 0; g < G; ++g) {  // split the concatenated gradient
        cudaMemcpy2DAsync(h->d_relsal + (size_t)g * B * Dr, sizeof(float) * Dr, h->d_feat + (size_t)g * B * Dt, sizeof(float) * Dt,
                          sizeof(float) * Dr, B, cudaMemcpyDeviceToDevice, st);
        cudaMemcpy2DAsync(h->da_dnf + (size_t)g * B * E, sizeof(float) * E, h->d_feat + (size_t)g * B * Dt + Dr, sizeof(float) * Dt,
                          sizeof(float) * E, B, cudaMemcpyDeviceToDevice, st);
    }
    const MlpIn rin = {h->rel_flat[0], Dr, (long long)B * Dr, B, B, Dr};
    mlp_backward(h, st, h->rse, rin, h->d_relsal, h->rse_h[0], h->d_h, h->d_relflat, 0);
    gtrunk_backward(h, st, h->gt_rel, h->d_relflat, Dr, 0, (long long)B * Dr);
    da3_backward(h, st, /*with_head=*/false);
}

// fork: side stream k continues from `st`; returns the stream pass k+1 is issued on and selects its split-K workspace
cudaStream_t fork_pass(maexp_learner *h, cudaStream_t st, int k, bool first) {
    if (first) cudaEventRecord(h->ev_fork, st);
    cudaStreamWaitEvent(h->side[k], h->ev_fork, 0);
    g_splitk_ws = h->splitk_side[k];
    return h->side[k];
}
void join_pass(maexp_learner *h, cudaStream_t st, int k) {
    cudaEventRecord(h->ev_join[k], h->side[k]);
    cudaStreamWaitEvent(st, h->ev_join[k], 0);
    g_splitk_ws = h->splitk_ws;
}

// ---- IQN head on a state embedding psi [G][B][D] (customs/iqn.py:96-136, customs/da3_iqn.py:26-84).  The embedding comes
//      from conv + state_embedder (IQN), the local transformer (DA3_IQN), relative embedder ++ local transformer
//      (MergedDA3_IQN) or the two-stage transformer (DA6_IQN)
int iqn_forward(maexp_learner *h, cudaStream_t st, const float *params, const float *obs, const int8_t *pos, int B,
                const float *taus, int nq, int per_sample, int pass) {
    maexp_learner::Acts &a = h->acts[pass];
    maexp_learner::IqnActs &q = h->iq[pass];
    const int G = h->G, D = h->D;
    const long long P = h->P, Bmax = h->B, R = h->rows_max;
    const float *psi = q.psi;
    if (h->da6) {
        if (da6_forward(h, st, params, obs, pos, B, pass, /*with_head=*/false)) return -1;
        psi = h->vloc.a[pass].nf;
    } else if (h->merged_da3) {
        float *unused = nullptr;
        if (merged_forward(h, st, params, obs, pos, B, pass, &unused, /*with_head=*/false)) return -1;
        psi = h->feat[pass];
    } else if (h->da3) {
        if (da3_forward(h, st, params, obs, B, pass, /*with_head=*/false, pos)) return -1;
        psi = h->vloc.a[pass].nf;
    } else if (h->relative_conv) {  // IQN on the relative view: conv(decoded canvas) -> state_embedder (customs/iqn.py:88-94,141-160)
        if (rel_forward(h, st, params, obs, pos, B, pass)) return -1;
        const MlpIn se_in = {h->feat[pass], D, Bmax * D, B, Bmax, D};
        mlp_forward(h, st, params, h->se, se_in, q.se_h, q.psi);
    } else if (h->merged_conv) {  // MergedIQN (customs/iqn.py:203-216): [rel embedder(conv(area)), local embedder(conv(crop))]
        if (rel_forward(h, st, params, obs, pos, B, pass)) return -1;
        const int Dr = h->Drel, Dl = h->Dloc;
        const MlpIn rin = {h->feat[pass], D, Bmax * D, B, Bmax, Dr};
        const MlpIn lin = {h->feat[pass] + Dr, D, Bmax * D, B, Bmax, Dl};
        mlp_forward(h, st, params, h->rse, rin, h->rse_h[pass], h->rse_out[pass]);
        mlp_forward(h, st, params, h->lse, lin, h->lse_h[pass], h->lse_out[pass]);
        // one 2-D copy per branch: rows (g, b) are uniform because the strides are per maximum batch
        cudaMemcpy2DAsync(q.psi, sizeof(float) * D, h->rse_out[pass], sizeof(float) * Dr, sizeof(float) * Dr, (size_t)G * Bmax,
                          cudaMemcpyDeviceToDevice, st);
        cudaMemcpy2DAsync(q.psi + Dr, sizeof(float) * D, h->lse_out[pass], sizeof(float) * Dl, sizeof(float) * Dl, (size_t)G * Bmax,
                          cudaMemcpyDeviceToDevice, st);
    } else {
        trunk_forward(h, st, params, obs, B, a, /*inference=*/pass != 0);  // only pass 0 feeds a backward pass
        const MlpIn se_in = {a.flat, D, Bmax * D, B, Bmax, D};
        mlp_forward(h, st, params, h->se, se_in, q.se_h, q.psi);  // state_embedder
    }
    const int rows = B * nq;
    iqn_cos_kernel<<<grid1d((long long)rows * h->ncos, G), 256, 0, st>>>(taus, q.cosf, B, nq, h->ncos, per_sample,
                                                                         per_sample ? (long long)B * nq : (long long)nq,
                                                                         R * h->ncos);
    // phi = relu(cos W^T + b)   (cosine_net.net.0 + ReLU, customs/iqn.py:24)
    gemm(st, G, q.cosf, params + h->o_cw, q.phi, params + h->o_cb, rows, D, h->ncos, h->ncos, 1, 1, h->ncos, D, 1,
         R * h->ncos, P, R * D, P, /*relu=*/1);
    iqn_hadamard_fwd_kernel<<<grid1d((long long)rows * D, G), 256, 0, st>>>(psi, q.phi, q.e, B, nq, D, Bmax * D, R * D);
    const MlpIn e_in = {q.e, D, R * D, rows, R, D};
    if (!mlp_fused_forward(h, st, params, h->heads.data(), 2, e_in, q.hrow, q.out))
        for (int k = 0; k < 2; ++k) mlp_forward(h, st, params, h->heads[k], e_in, q.hrow[k], q.out[k]);
    return 0;
}

void iqn_backward(maexp_learner *h, cudaStream_t st, int nq) {
    maexp_learner::Acts &a = h->acts[0];
    maexp_learner::IqnActs &q = h->iq[0];
    const int G = h->G, B = h->B, D = h->D, rows = B * nq;
    const long long P = h->P, R = h->rows_max;
    const float *psi = h->merged_da3 ? h->feat[0] : h->da3 ? h->vloc.a[0].nf : q.psi;
    const MlpIn e_in = {q.e, D, R * D, rows, R, D};
    const bool side_wgrads = !h->da3 && !h->relative_conv && !h->merged_conv;  // plain IQN: buffers audited for the overlap
    if (side_wgrads) wgrad_open(h, st);
    for (int k = 0; k < 2; ++k)
        mlp_backward(h, st, h->heads[k], e_in, h->d_out[k], q.hrow[k], k & 1 ? h->d_hrow2 : h->d_hrow, h->d_e, k > 0);
    // h->d_psi aliases the buffer the embedding's producer reads its output gradient from (set in create)
    iqn_hadamard_bwd_kernel<<<grid1d((long long)B * D, G), 256, 0, st>>>(h->d_e, psi, q.phi, h->d_phi, h->d_psi, B, nq, D,
                                                                         (long long)B * D, R * D);
    // cosine net: dW[D,ncos] = dphi^T cos ; db = colsum(dphi)
    gemm(st, G, h->d_phi, q.cosf, h->grads + h->o_cw, nullptr, D, h->ncos, rows, 1, D, h->ncos, 1, h->ncos, 1, R * D,
         R * h->ncos, P, 0);
    colsum(st, G, h->ones, h->d_phi, h->grads + h->o_cb, rows, D, R * D, P);
    if (h->da6) {
        da6_backward(h, st, /*with_head=*/false);
    } else if (h->merged_da3) {
        merged_backward(h, st, /*with_head=*/false);
    } else if (h->da3) {
        da3_backward(h, st, /*with_head=*/false);
    } else if (h->relative_conv) {
        const MlpIn se_in = {h->feat[0], D, (long long)B * D, B, B, D};
        mlp_backward(h, st, h->se, se_in, h->d_psi, q.se_h, h->d_h, h->d_feat, 0);
        gtrunk_backward(h, st, h->gt_rel, h->d_feat, D, 0, (long long)B * D);
    } else if (h->merged_conv) {
        const int Dr = h->Drel, Dl = h->Dloc;
        cudaMemcpy2DAsync(h->d_relsal, sizeof(float) * Dr, h->d_psi, sizeof(float) * D, sizeof(float) * Dr, (size_t)G * B,
                          cudaMemcpyDeviceToDevice, st);
        cudaMemcpy2DAsync(h->d_locsal, sizeof(float) * Dl, h->d_psi + Dr, sizeof(float) * D, sizeof(float) * Dl, (size_t)G * B,
                          cudaMemcpyDeviceToDevice, st);
        const MlpIn rin = {h->feat[0], D, (long long)B * D, B, B, Dr};
        const MlpIn lin = {h->feat[0] + Dr, D, (long long)B * D, B, B, Dl};
        mlp_backward(h, st, h->rse, rin, h->d_relsal, h->rse_h[0], h->d_h, h->d_relflat, 0);
        gtrunk_backward(h, st, h->gt_rel, h->d_relflat, Dr, 0, (long long)B * Dr);
        mlp_backward(h, st, h->lse, lin, h->d_locsal, h->lse_h[0], h->d_h, h->d_locflat, 0);
        gtrunk_backward(h, st, h->gt_loc, h->d_locflat, Dl, 0, (long long)B * Dl);
    } else {
        const MlpIn se_in = {a.flat, D, (long long)B * D, B, B, D};
        mlp_backward(h, st, h->se, se_in, h->d_psi, q.se_h, h->d_h, h->d_flat, 0);
        trunk_backward(h, st);
    }
    if (side_wgrads) wgrad_close();
}

}  // namespace

extern "C" {

int maexp_learner_create(const maexp_learner_cfg *cfg, maexp_learner **out) {
    if (!cfg || !out) return fail("maexp_learner_create: null argument");
    if (cfg->kind < MAEXP_LEARNER_DQN_CONV_MLP || cfg->kind > MAEXP_LEARNER_DA6)
        return fail("maexp_learner_create: unknown kind");
    if (cfg->num_agents < 1 || cfg->batch_size < 1) return fail("maexp_learner_create: bad num_agents / batch_size");
    if (cfg->optimizer < 0 || cfg->optimizer > 1) return fail("maexp_learner_create: optimizer must be 0 (rmsprop) or 1 (adam)");
    if (cfg->num_actions < 1 || cfg->num_actions > kMaxActions) return fail("maexp_learner_create: num_actions must be in 1..8");
    maexp_learner *h = new maexp_learner();
    h->cfg = *cfg;
    if (const char *tc = getenv("MAEXP_TC")) h->use_tc = atoi(tc) != 0;
    if (const char *ft = getenv("MAEXP_FUSED_TRUNK")) h->fused_trunk = atoi(ft) != 0;
    if (const char *fm = getenv("MAEXP_FUSED_MLP")) h->fused_mlp = atoi(fm) != 0;
    if (const char *tm = getenv("MAEXP_TC_MLP")) h->tc_mlp = atoi(tm) != 0;
    if (const char *th = getenv("MAEXP_TC_HEADS")) h->tc_heads = atoi(th) != 0;
    if (const char *tw = getenv("MAEXP_TC_WGRAD")) h->tc_wgrad = atoi(tw) != 0;
    if (const char *ms = getenv("MAEXP_MLP_BWD_SPLIT")) h->mlp_bwd_split = atoi(ms) != 0;
    if (const char *pb = getenv("MAEXP_PREP_BRANCH")) h->prep_branch = atoi(pb) != 0;
    if (const char *hp = getenv("MAEXP_MLP_BWD_HEADS_PARALLEL")) h->mlp_bwd_head_par = atoi(hp) != 0;
    if (const char *tt = getenv("MAEXP_TC_TRUNK")) h->tc_trunk = atoi(tt) != 0;
    if (const char *bf = getenv("MAEXP_BATCHED_FWD")) h->batched_fwd = atoi(bf) != 0;
    h->G = cfg->num_agents; h->B = cfg->batch_size; h->C = cfg->channels; h->V = cfg->visible_range;
    h->A = cfg->num_actions; h->H = cfg->hidden; h->F1 = cfg->conv1_channels; h->F2 = cfg->conv2_channels;
    const int trunk = cfg->kind == MAEXP_LEARNER_IQN ? cfg->iqn_trunk : 0;  // MAEXP_IQN_TRUNK_*
    if (trunk < 0 || trunk > MAEXP_IQN_TRUNK_DA6) { delete h; return fail("maexp_learner_create: unknown iqn_trunk"); }
    h->da6 = cfg->kind == MAEXP_LEARNER_DA6 || trunk == MAEXP_IQN_TRUNK_DA6;
    h->da3 = cfg->kind == MAEXP_LEARNER_DA3 || h->da6 || trunk != 0;  // da3: "has transformer stacks"
    h->area_mask = cfg->area_mask; h->X = cfg->size_x; h->Y = cfg->size_y;
    // conv trunks on the decoded canvas: ConvMLP / IQN on the relative view, MergedConvMLP / MergedIQN on the merged view
    const bool conv_kind = cfg->kind == MAEXP_LEARNER_DQN_CONV_MLP || (cfg->kind == MAEXP_LEARNER_IQN && trunk == MAEXP_IQN_TRUNK_CONV);
    h->relative_conv = conv_kind && cfg->area_mask == MAEXP_MASK_RELATIVE;
    h->merged_conv = conv_kind && cfg->area_mask == MAEXP_MASK_MERGED;
    h->merged_da3 = (cfg->kind == MAEXP_LEARNER_DA3 && cfg->area_mask == MAEXP_MASK_MERGED) || trunk == MAEXP_IQN_TRUNK_MERGED_DA3;
    if (trunk == MAEXP_IQN_TRUNK_DA3 && cfg->area_mask == MAEXP_MASK_MERGED) { delete h; return fail("maexp_learner_create: the DA3 IQN trunk reads the local or relative view"); }
    h->relative_vit = h->da3 && !h->da6 && !h->merged_da3 && cfg->area_mask == MAEXP_MASK_RELATIVE;
    if (h->relative_vit && (h->X < 4 || h->Y < 4 || cfg->patch_size < 1)) { delete h; return fail("maexp_learner_create: relative DA3 needs size_x / size_y / patch_size"); }
    if (trunk == MAEXP_IQN_TRUNK_MERGED_DA3 && cfg->area_mask != MAEXP_MASK_MERGED) { delete h; return fail("maexp_learner_create: the MergedDA3 IQN trunk reads the merged view"); }
    const bool gmodel = h->relative_conv || h->merged_conv || h->merged_da3;
    // in_chans the reference adds to each model family's (relative) input: ConvMLP and DA3 objects only (conv_mlp.py:25-27,
    // da3.py:31-33), IQN objects + agents (iqn.py:55-60), DA3_IQN all three (da3_iqn.py:96-104); DA6 / DA6_IQN read
    // objects + agents on their relative stack (da6.py:32-37, da6_iqn.py:42-47)
    if ((cfg->objects_channel && (cfg->type_objects < 1 || cfg->type_objects > MAEXP_MAX_TYPES)) ||
        (cfg->agents_channel && (cfg->team_agents < 1 || cfg->team_agents > MAEXP_MAX_AGENTS))) {
        delete h;
        return fail("maexp_learner_create: objects_channel / agents_channel need type_objects / team_agents");
    }
    // CategoricalDQN, QRDQN and the plain MLP never read the flags
    const bool reads_flags = cfg->kind == MAEXP_LEARNER_DQN_CONV_MLP || cfg->kind == MAEXP_LEARNER_IQN ||
                             cfg->kind == MAEXP_LEARNER_DA3 || cfg->kind == MAEXP_LEARNER_DA6;
    const int xo = reads_flags && cfg->objects_channel ? cfg->type_objects : 0,
              xa = reads_flags && cfg->agents_channel ? cfg->team_agents : 0,
              xd = reads_flags && cfg->destination_channel ? 1 : 0;
    const bool iqn_kind = cfg->kind == MAEXP_LEARNER_IQN;
    const bool vit_iqn = iqn_kind && trunk != MAEXP_IQN_TRUNK_CONV;  // DA3_IQN and its subclasses
    // what the family's constructor sizes its relative-view input with (also where a subclass never uses it)
    const int parent_extra = xo + ((iqn_kind) ? xa : 0) + (vit_iqn ? xd : 0);
    if (h->relative_conv || h->relative_vit) { h->ex_o = xo; h->ex_a = iqn_kind ? xa : 0; h->ex_d = vit_iqn ? xd : 0; }
    if (h->da6) { h->ex_o = xo; h->ex_a = xa; h->ex_d = 0; }
    h->Cx = h->ex_o + h->ex_a + h->ex_d;
    if (cfg->area_mask == MAEXP_MASK_LOCAL && parent_extra) {
        delete h;  // the reference sizes the first layer for C + extras but feeds it the C-channel crop: a shape error
        return fail("maexp_learner_create: objects_channel / agents_channel / destination_channel need the relative or merged view");
    }
    if (h->da6 && (cfg->area_mask != MAEXP_MASK_MERGED || h->X < 4 || h->Y < 4)) { delete h; return fail("maexp_learner_create: DA6 reads the merged view (area_mask merged, size_x / size_y set)"); }
    if (cfg->area_mask != MAEXP_MASK_LOCAL && !gmodel && !h->da6 && !h->relative_vit) { delete h; return fail("maexp_learner_create: relative / merged views are implemented for conv_mlp (dqn), iqn, da3 and da6"); }
    if (gmodel && (h->X < 4 || h->Y < 4)) { delete h; return fail("maexp_learner_create: size_x / size_y missing for a relative / merged model"); }
    h->has_conv = cfg->kind != MAEXP_LEARNER_DQN_MLP && !h->da3 && !gmodel;
    h->dist = cfg->kind == MAEXP_LEARNER_C51 || cfg->kind == MAEXP_LEARNER_QR;
    h->iqn = cfg->kind == MAEXP_LEARNER_IQN;
    h->Nq = h->iqn ? cfg->num_atoms : 0;
    h->Nd = h->iqn ? cfg->num_quantiles_dash : 0;
    h->ncos = h->iqn ? cfg->num_cosines : 0;
    if (h->iqn && (h->Nq < 1 || h->Nd < 1 || h->ncos < 1)) { delete h; return fail("maexp_learner_create: IQN needs num_atoms (quantiles), num_quantiles_dash, num_cosines >= 1"); }
    h->rows_max = h->B * (h->Nq > h->Nd ? h->Nq : h->Nd);
    h->Z = h->dist ? cfg->num_atoms : 1;
    if (h->dist && (h->Z < 1 || h->Z > 32 * kMaxAtomsPerLane)) { delete h; return fail("maexp_learner_create: num_atoms must be in 1..128"); }
    if (cfg->obs_stride < h->C * h->V * h->V) { delete h; return fail("maexp_learner_create: obs_stride < C*V*V"); }
    long long off = 0;
    auto add = [&](const std::string &name, long long count) {
        off = (off + 3) & ~3ll;  // every tensor starts on a 16-byte boundary of the flat vector (vector loads of weights)
        h->table.push_back({name, off, count});
        const long long o = off;
        off += count;
        return o;
    };
    auto set_trunk = [&](maexp_learner::GTrunk &t, const std::string &prefix, int cin, int hin, int win) {
        t.Cin = cin; t.Hin = hin; t.Win = win;
        t.H1 = hin - 1; t.W1 = win - 1; t.Hp = t.H1 / 2; t.Wp = t.W1 / 2; t.H2 = t.Hp - 1; t.W2 = t.Wp - 1;
        t.D = h->F2 * t.H2 * t.W2;
        t.c1w = add(prefix + ".conv_list.0.weight", (long long)h->F1 * cin * 4);
        t.c1b = add(prefix + ".conv_list.0.bias", h->F1);
        t.c2w = add(prefix + ".conv_post.weight", (long long)h->F2 * h->F1 * 4);
        t.c2b = add(prefix + ".conv_post.bias", h->F2);
    };
    if (h->relative_conv) {  // ConvMLP on the decoded [C, X, Y] canvas (conv_mlp.py:21-66)
        set_trunk(h->gt_rel, "conv", h->C + h->Cx, h->X, h->Y);
        h->Drel = h->gt_rel.D; h->Dloc = 0; h->D = h->Drel;
        h->H1 = h->Hp = h->H2 = 0;
    } else if (h->merged_conv) {  // MergedConvMLP (conv_mlp.py:69-108); the parent's `conv` keeps unused parameters
        maexp_learner::GTrunk unused;
        set_trunk(unused, "conv", h->C + parent_extra, h->X, h->Y);
        if (cfg->kind == MAEXP_LEARNER_IQN) {  // IQN.__init__'s state_embedder on the unused trunk (customs/iqn.py:51-53,167)
            add("state_embedder.fc_list.0.weight", (long long)h->H * unused.D); add("state_embedder.fc_list.0.bias", h->H);
            add("state_embedder.fc_post.weight", (long long)unused.D * h->H); add("state_embedder.fc_post.bias", unused.D);
        }
        set_trunk(h->gt_loc, "local_conv", h->C, h->V, h->V);
        set_trunk(h->gt_rel, "relative_conv", 1, h->X, h->Y);
        h->Drel = h->gt_rel.D; h->Dloc = h->gt_loc.D; h->D = h->Drel + h->Dloc;
        h->H1 = h->Hp = h->H2 = 0;
    }
    if (h->da3) {
        h->E_ = cfg->embed_dim; h->heads_ = cfg->num_heads; h->Fm = cfg->mlp_hidden; h->L_ = cfg->block_loop;
        if (cfg->attention_hard) {
            if (!(cfg->af_lambd >= 0.0f)) { delete h; return fail("maexp_learner_create: af_lambd must be >= 0 (nn.Hardshrink)"); }
            h->attn_shrink = cfg->af_lambd;
        }
        h->H1 = h->Hp = h->H2 = 0;
        h->D = h->E_;
        auto set_stack = [&](maexp_learner::VitStack &s, int cin, int hin, int win, int ps) {
            s.Cin = cin; s.Hin = hin; s.Win = win; s.ps = ps > 0 ? ps : 1;
            s.Ht = hin / s.ps; s.Wt = win / s.ps; s.Tn = s.Ht * s.Wt + 1; s.K = cin * s.ps * s.ps;
        };
        if (h->relative_vit) set_stack(h->vloc, h->C + h->Cx, h->X, h->Y, cfg->patch_size);
        else set_stack(h->vloc, h->C, h->V, h->V, h->da6 ? cfg->local_patch_size : 1);
        if (h->da6) set_stack(h->vrel, 1 + h->Cx, h->X, h->Y, cfg->relative_patch_size);
        h->Tn = h->da6 && h->vrel.Tn > h->vloc.Tn ? h->vrel.Tn : h->vloc.Tn;
        if (h->E_ < 1 || h->heads_ < 1 || h->E_ % h->heads_ || h->E_ / h->heads_ > kMaxHeadDim || h->Tn > kMaxTokens || h->Fm < 1 ||
            h->vloc.Tn < 2 || (h->da6 && h->vrel.Tn < 2) ||
            h->L_ < 0 || (4 * h->Tn * (h->E_ / h->heads_) + h->Tn * h->Tn) * 4 > 48 * 1024) {
            delete h;
            return fail("maexp_learner_create: unsupported DA3 / DA6 shape (2 <= tokens <= 96, head_dim <= 32, attention tile <= 48 KB)");
        }
        auto add_blocks = [&](maexp_learner::VitStack &s, const std::string &prefix) {
            for (int l = 0; l < h->L_; ++l) {
                const std::string b = prefix + "blocks." + std::to_string(l) + ".";
                maexp_learner::BlockLayout w;
                w.n1w = add(b + "norm1.weight", h->E_); w.n1b = add(b + "norm1.bias", h->E_);
                w.qkvw = add(b + "attn.qkv.weight", 3ll * h->E_ * h->E_);
                w.projw = add(b + "attn.proj.weight", (long long)h->E_ * h->E_); w.projb = add(b + "attn.proj.bias", h->E_);
                w.n2w = add(b + "norm2.weight", h->E_); w.n2b = add(b + "norm2.bias", h->E_);
                w.f1w = add(b + "mlp.fc1.weight", (long long)h->Fm * h->E_); w.f1b = add(b + "mlp.fc1.bias", h->Fm);
                w.f2w = add(b + "mlp.fc2.weight", (long long)h->E_ * h->Fm); w.f2b = add(b + "mlp.fc2.bias", h->E_);
                s.blocks.push_back(w);
            }
        };
        // own parameters first, then submodules in registration order (customs/da3.py:33-60, customs/da6.py:40-96)
        auto add_unused_parent = [&](bool with_saliency) {  // what DA3.__init__ / DA3_IQN.__init__ creates and a subclass never uses
            const int ps = cfg->patch_size > 0 ? cfg->patch_size : 1;
            if (with_saliency) add("saliency_vector", h->E_);
            add("pos_embed", (long long)((h->X / ps) * (h->Y / ps) + 1) * h->E_);
            add("patch_embed.proj.weight", (long long)h->E_ * (h->C + parent_extra) * ps * ps);
            add("patch_embed.proj.bias", h->E_);
            maexp_learner::VitStack unused;
            add_blocks(unused, "");
        };
        if (h->da6) {
            if (h->iqn) add_unused_parent(false);  // DA6_IQN(DA3_IQN) re-assigns saliency_vector (da6_iqn.py:25-28,66)
            h->vrel.o_sal = add("saliency_vector", h->E_);
            h->vrel.o_pos = add("relative_pos_embed", (long long)h->vrel.Tn * h->E_);
            h->vloc.o_pos = add("local_pos_embed", (long long)h->vloc.Tn * h->E_);
            h->vrel.o_pew = add("relative_patch_embed.proj.weight", (long long)h->E_ * h->vrel.K);
            h->vrel.o_peb = add("relative_patch_embed.proj.bias", h->E_);
            h->vloc.o_pew = add("local_patch_embed.proj.weight", (long long)h->E_ * h->vloc.K);
            h->vloc.o_peb = add("local_patch_embed.proj.bias", h->E_);
            add_blocks(h->vrel, "relative_");
            add_blocks(h->vloc, "local_");
        } else {
        const std::string lp = h->merged_da3 ? "local_" : "";
        if (h->merged_da3) add_unused_parent(true);  // da3.py:99-103, da3_iqn.py:214-218
        h->vloc.o_sal = add(lp + "saliency_vector", h->E_);
        h->vloc.o_pos = add(lp + "pos_embed", (long long)h->vloc.Tn * h->E_);
        h->vloc.o_pew = add(lp + "patch_embed.proj.weight", (long long)h->E_ * h->vloc.K);
        h->vloc.o_peb = add(lp + "patch_embed.proj.bias", h->E_);
        add_blocks(h->vloc, lp);
        }
        h->o_nfw = add("norm.weight", h->E_); h->o_nfb = add("norm.bias", h->E_);
        if (h->da6) { h->o_f1w = add("fc1.weight", (long long)h->E_ * h->E_); h->o_f1b = add("fc1.bias", h->E_); }
        if (h->merged_da3) {
            set_trunk(h->gt_rel, "relative_conv", 1, h->X, h->Y);
            h->Drel = h->gt_rel.D; h->Dloc = 0;
        }
        if (h->iqn) h->D = h->E_ + h->Drel;  // embedding_dim of IQN_Head (da3_iqn.py:142-144,263-265)
        else { h->o_hw = add("drl_head.weight", (long long)h->A * (h->E_ + h->Drel)); h->o_hb = add("drl_head.bias", h->A); }
    } else if (gmodel) {
        // layout set above
    } else if (h->has_conv) {
        h->H1 = h->V - 1; h->Hp = h->H1 / 2; h->H2 = h->Hp - 1;
        if (h->H2 < 1) { delete h; return fail("maexp_learner_create: visible_range too small for the conv trunk"); }
        h->D = h->F2 * h->H2 * h->H2;
        // state_dict order of convolution.py:21-42
        h->o_c1w = add("conv.conv_list.0.weight", (long long)h->F1 * h->C * 4);
        h->o_c1b = add("conv.conv_list.0.bias", h->F1);
        h->o_c2w = add("conv.conv_post.weight", (long long)h->F2 * h->F1 * 4);
        h->o_c2b = add("conv.conv_post.bias", h->F2);
    } else {
        h->H1 = h->Hp = h->H2 = 0;
        h->D = h->C * h->V * h->V;
    }
    auto add_head = [&](const std::string &prefix, int outn) {
        HeadLayout hd;
        hd.w1 = add(prefix + ".fc_list.0.weight", (long long)h->H * h->D);
        hd.b1 = add(prefix + ".fc_list.0.bias", h->H);
        hd.w2 = add(prefix + ".fc_post.weight", (long long)outn * h->H);
        hd.b2 = add(prefix + ".fc_post.bias", outn);
        hd.out = outn;
        h->heads.push_back(hd);
    };
    if (h->merged_da3) {  // relative_state_embedder: MLP(Drel -> H -> Drel)  (da3.py:160-162)
        const int keepD = h->D;
        h->D = h->Drel;
        add_head("relative_state_embedder", h->Drel);
        h->rse = h->heads.back();
        h->heads.pop_back();
        h->D = keepD;
    }
    if (h->da3 && h->iqn) {  // IQN_Head (da3_iqn.py:26-40)
        h->o_cw = add("drl_head.cosine_net.net.0.weight", (long long)h->D * h->ncos);
        h->o_cb = add("drl_head.cosine_net.net.0.bias", h->D);
        add_head("drl_head.fc_V", 1);
        add_head("drl_head.fc_A", h->A);
    } else if (h->da3) {
        // no MLP heads: the drl_head is part of the layout above
    } else if ((h->relative_conv || h->merged_conv) && !h->iqn) {
        add_head("mlp", h->A);
    } else if (h->iqn) {  // state_dict order of customs/iqn.py:43-67
        if (h->merged_conv) {  // MergedIQN: one embedder per branch (customs/iqn.py:183-188)
            const int keepD = h->D;
            h->D = h->Dloc; add_head("local_state_embedder", h->Dloc); h->lse = h->heads.back(); h->heads.pop_back();
            h->D = h->Drel; add_head("relative_state_embedder", h->Drel); h->rse = h->heads.back(); h->heads.pop_back();
            h->D = keepD;
        } else {
            add_head("state_embedder", h->D);
            h->se = h->heads.back();
            h->heads.pop_back();
        }
        h->o_cw = add("cosine_net.net.0.weight", (long long)h->D * h->ncos);
        h->o_cb = add("cosine_net.net.0.bias", h->D);
        add_head("fc_V", 1);
        add_head("fc_A", h->A);
    } else if (h->dist) {
        add_head("fc_V", h->Z);           // categorical_dqn.py:47-52, qr_dqn.py:37-42
        add_head("fc_A", h->A * h->Z);
    } else if (h->has_conv) {
        add_head("mlp", h->A);            // conv_mlp.py:41
    } else {
        add_head("", h->A);               // mlp.py:19-26 (flattened observation)
        for (auto &e : h->table) e.name = e.name.substr(1);
    }
    h->P = (off + 3) & ~3ll;  // padding elements stay zero: zero gradient, zero optimizer update
    const long long G = h->G, B = h->B, P = h->P;
    int rc = 0;
    rc |= lalloc(h, &h->online, G * P);
    rc |= lalloc(h, &h->target, G * P);
    rc |= lalloc(h, &h->grads, G * P);
    rc |= lalloc(h, &h->sq, G * P);
    if (cfg->optimizer == 1) {
        rc |= lalloc(h, &h->exp_avg, G * P);
        rc |= lalloc(h, &h->adam_steps, G);
    }
    rc |= lalloc(h, &h->support, 128);
    rc |= lalloc(h, &h->ones, 4);
    rc |= lalloc(h, &h->splitk_ws, (size_t)kSplitWorkspaceFloats);
    for (int k = 0; k < kSides; ++k) {
        rc |= lalloc(h, &h->splitk_side[k], (size_t)kSplitWorkspaceFloats);
        {   // MAEXP_SIDE_PRIORITY=low: the weight-gradient / inference branches yield to the data-gradient chain on the main stream
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            const char *pr = getenv("MAEXP_SIDE_PRIORITY");
            const int prio = (pr && pr[0] == 'l') ? lo : (pr && pr[0] == 'h') ? hi : 0;
            rc |= cudaStreamCreateWithPriority(&h->side[k], cudaStreamNonBlocking, prio) != cudaSuccess;
        }
        rc |= cudaEventCreateWithFlags(&h->ev_join[k], cudaEventDisableTiming) != cudaSuccess;
        rc |= cudaEventCreateWithFlags(&h->ev_w[k], cudaEventDisableTiming) != cudaSuccess;
    }
    rc |= cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess;
    rc |= cudaEventCreateWithFlags(&h->ev_col2, cudaEventDisableTiming) != cudaSuccess;
    for (int k = 0; k < 3; ++k) rc |= cudaEventCreateWithFlags(&h->ev_mark[k], cudaEventDisableTiming) != cudaSuccess;
    for (int k = 0; k < 3; ++k) rc |= cudaEventCreateWithFlags(&h->ev_prep[k], cudaEventDisableTiming) != cudaSuccess;
    rc |= cudaEventCreateWithFlags(&h->ev_zero, cudaEventDisableTiming) != cudaSuccess;
    if (const char *cf = getenv("MAEXP_CRITICAL_FIRST")) h->critical_first = atoi(cf) != 0;
    for (int k = 0; k < 3; ++k) {
        maexp_learner::Acts &a = h->acts[k];
        memset(&a, 0, sizeof(a));
        if (h->has_conv) {
            rc |= lalloc(h, &a.col1, G * B * h->H1 * h->H1 * h->C * 4);
            rc |= lalloc(h, &a.y1, G * B * h->H1 * h->H1 * h->F1);
            rc |= lalloc(h, &a.p, G * B * h->Hp * h->Hp * h->F1);
            rc |= lalloc(h, &a.arg, G * B * h->Hp * h->Hp * h->F1);
            rc |= lalloc(h, &a.col2, G * B * h->H2 * h->H2 * h->F1 * 4);
            rc |= lalloc(h, &a.y2, G * B * h->D);
            rc |= lalloc(h, &a.flat, G * B * h->D);
        }
        for (size_t j = 0; j < h->heads.size(); ++j) {
            rc |= lalloc(h, &a.h[j], G * B * h->H);
            rc |= lalloc(h, &a.out[j], G * B * h->heads[j].out);
        }
    }
    const long long R = h->iqn ? h->rows_max : B;
    if (h->da3) {
        const long long Rm = B * h->Tn, R49 = B * (h->Tn - 1), E = h->E_;
        auto alloc_stack = [&](maexp_learner::VitStack &s) {
            const long long Rs = B * s.Tn, Rt = B * (s.Tn - 1);
            for (int k = 0; k < 3; ++k) {
                maexp_learner::Da3Acts &a = s.a[k];
                rc |= lalloc(h, &a.patches, G * Rt * s.K);
                rc |= lalloc(h, &a.tok, G * Rt * E);
                rc |= lalloc(h, &a.x0, G * Rs * E);
                rc |= lalloc(h, &a.nf, G * B * E);
                rc |= lalloc(h, &a.mf, G * B);
                rc |= lalloc(h, &a.rf, G * B);
                rc |= lalloc(h, &a.q, G * B * h->A);
                a.blk.resize(h->L_);
                for (int l = 0; l < h->L_; ++l) {
                    maexp_learner::Da3Block &q = a.blk[l];
                    rc |= lalloc(h, &q.n1, G * Rs * E); rc |= lalloc(h, &q.m1, G * Rs); rc |= lalloc(h, &q.r1, G * Rs);
                    rc |= lalloc(h, &q.qkv, G * Rs * 3 * E);
                    rc |= lalloc(h, &q.probs, G * B * h->heads_ * s.Tn * s.Tn);
                    rc |= lalloc(h, &q.ao, G * Rs * E); rc |= lalloc(h, &q.x1, G * Rs * E);
                    rc |= lalloc(h, &q.n2, G * Rs * E); rc |= lalloc(h, &q.m2, G * Rs); rc |= lalloc(h, &q.r2, G * Rs);
                    rc |= lalloc(h, &q.f1, G * Rs * h->Fm); rc |= lalloc(h, &q.g1, G * Rs * h->Fm);
                    rc |= lalloc(h, &q.x2, G * Rs * E);
                }
            }
        };
        alloc_stack(h->vloc);
        if (h->da6) {
            alloc_stack(h->vrel);
            for (int k = 0; k < 3; ++k) {
                rc |= lalloc(h, &h->sal2[k], G * B * E);
                rc |= lalloc(h, &h->canvas[k], G * B * (1 + h->Cx) * h->X * h->Y);
            }
            rc |= lalloc(h, &h->d_sal2, G * B * E);
        }
        if (h->relative_vit)
            for (int k = 0; k < 3; ++k) rc |= lalloc(h, &h->canvas[k], G * B * (h->C + h->Cx) * h->X * h->Y);
        rc |= lalloc(h, &h->d_out[0], G * B * h->A);
        rc |= lalloc(h, &h->da_dx, G * Rm * E); rc |= lalloc(h, &h->da_dn, G * Rm * E); rc |= lalloc(h, &h->da_gx, G * Rm * E);
        rc |= lalloc(h, &h->da_dqkv, G * Rm * 3 * E); rc |= lalloc(h, &h->da_dao, G * Rm * E);
        rc |= lalloc(h, &h->da_dg1, G * Rm * h->Fm); rc |= lalloc(h, &h->da_df1, G * Rm * h->Fm);
        rc |= lalloc(h, &h->da_dnf, G * B * E); rc |= lalloc(h, &h->da_dtok0, G * B * E); rc |= lalloc(h, &h->da_dtok, G * R49 * E);
    }
    for (size_t j = 0; j < h->heads.size(); ++j) rc |= lalloc(h, &h->d_out[j], G * R * h->heads[j].out);
    if (h->iqn) {
        for (int k = 0; k < 3; ++k) {
            maexp_learner::IqnActs &q = h->iq[k];
            rc |= lalloc(h, &q.se_h, G * B * h->H);
            rc |= lalloc(h, &q.psi, G * B * h->D);
            rc |= lalloc(h, &q.cosf, G * R * h->ncos);
            rc |= lalloc(h, &q.phi, G * R * h->D);
            rc |= lalloc(h, &q.e, G * R * h->D);
            for (int j = 0; j < 2; ++j) {
                rc |= lalloc(h, &q.hrow[j], G * R * h->H);
                rc |= lalloc(h, &q.out[j], G * R * h->heads[j].out);
            }
        }
        rc |= lalloc(h, &h->d_e, G * R * h->D);
        rc |= lalloc(h, &h->d_phi, G * R * h->D);
        rc |= lalloc(h, &h->d_psi, G * B * h->D);

        rc |= lalloc(h, &h->d_hrow, G * R * h->H);
        rc |= lalloc(h, &h->d_hrow2, G * R * h->H);
    }
    rc |= lalloc(h, &h->d_h, G * B * h->H);
    rc |= lalloc(h, &h->d_h2, G * B * h->H);
    if (h->has_conv) {
        rc |= lalloc(h, &h->d_flat, G * B * h->D);
        rc |= lalloc(h, &h->d_y2, G * B * h->D);
        rc |= lalloc(h, &h->d_col2, G * B * h->H2 * h->H2 * h->F1 * 4);
        rc |= lalloc(h, &h->trunk_part, (size_t)G * kTrunkBwdBlocks * (h->C * 4 * 32 + 96));
        rc |= lalloc(h, &h->d_p, G * B * h->Hp * h->Hp * h->F1);
        rc |= lalloc(h, &h->d_y1, G * B * h->H1 * h->H1 * h->F1);
    }
    if (gmodel) {
        auto alloc_trunk = [&](maexp_learner::GTrunk &t) {
            const long long n1 = (long long)t.H1 * t.W1, np = (long long)t.Hp * t.Wp, n2 = (long long)t.H2 * t.W2;
            for (int k = 0; k < 3; ++k) {
                rc |= lalloc(h, &t.a[k].col1, G * B * n1 * t.Cin * 4);
                rc |= lalloc(h, &t.a[k].y1, G * B * n1 * h->F1);
                rc |= lalloc(h, &t.a[k].p, G * B * np * h->F1);
                rc |= lalloc(h, &t.a[k].arg, G * B * np * h->F1);
                rc |= lalloc(h, &t.a[k].col2, G * B * n2 * h->F1 * 4);
                rc |= lalloc(h, &t.a[k].y2, G * B * n2 * h->F2);
            }
            rc |= lalloc(h, &t.d_y2, G * B * n2 * h->F2);
            rc |= lalloc(h, &t.d_col2, G * B * n2 * h->F1 * 4);
            rc |= lalloc(h, &t.d_p, G * B * np * h->F1);
            rc |= lalloc(h, &t.d_y1, G * B * n1 * h->F1);
        };
        alloc_trunk(h->gt_rel);
        if (h->merged_conv) alloc_trunk(h->gt_loc);
        const long long Dt = h->merged_da3 ? h->Drel + h->E_ : h->D;
        for (int k = 0; k < 3; ++k) {
            rc |= lalloc(h, &h->canvas[k], G * B * (h->relative_conv ? h->C + h->Cx : 1) * h->X * h->Y);
            rc |= lalloc(h, &h->feat[k], G * B * Dt);
            if (h->merged_da3 || (h->iqn && h->merged_conv)) {
                rc |= lalloc(h, &h->rel_flat[k], G * B * h->Drel);
                rc |= lalloc(h, &h->rse_h[k], G * B * h->H);
                rc |= lalloc(h, &h->rse_out[k], G * B * h->Drel);
            }
            if (h->iqn && h->merged_conv) {
                rc |= lalloc(h, &h->lse_h[k], G * B * h->H);
                rc |= lalloc(h, &h->lse_out[k], G * B * h->Dloc);
            }
        }
        rc |= lalloc(h, &h->d_feat, G * B * Dt);
        if (h->merged_da3 || (h->iqn && h->merged_conv)) {
            rc |= lalloc(h, &h->d_relsal, G * B * h->Drel);
            rc |= lalloc(h, &h->d_relflat, G * B * h->Drel);
        }
        if (h->iqn && h->merged_conv) {
            rc |= lalloc(h, &h->d_locsal, G * B * h->Dloc);
            rc |= lalloc(h, &h->d_locflat, G * B * h->Dloc);
        }
    }
    if (h->iqn && h->da3)  // the embedding's producer reads its output gradient from its own buffer
        h->d_psi = h->merged_da3 ? h->d_feat : h->da_dnf;
    rc |= lalloc(h, &h->per_sample, G * B);
    rc |= lalloc(h, &h->m_scratch, G * B * 128);
    rc |= lalloc(h, &h->q_tmp, G * B * h->A);
    h->nparts = 32;
    rc |= lalloc(h, &h->partial, G * h->nparts);
    if (rc) { maexp_learner_destroy(h); return -1; }
    const float one = 1.0f;
    if (cudaMemcpy(h->ones, &one, sizeof(one), cudaMemcpyHostToDevice) != cudaSuccess) { maexp_learner_destroy(h); return fail("maexp_learner_create: memcpy"); }
    *out = h;
    return 0;
}

void maexp_learner_destroy(maexp_learner *h) {
    if (!h) return;
    for (int k = 0; k < kSides; ++k) {
        if (h->side[k]) { cudaStreamSynchronize(h->side[k]); cudaStreamDestroy(h->side[k]); }
        if (h->ev_join[k]) cudaEventDestroy(h->ev_join[k]);
        if (h->ev_w[k]) cudaEventDestroy(h->ev_w[k]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_col2) cudaEventDestroy(h->ev_col2);
    if (h->ev_zero) cudaEventDestroy(h->ev_zero);
    for (int k = 0; k < 3; ++k) {
        if (h->ev_mark[k]) cudaEventDestroy(h->ev_mark[k]);
        if (h->ev_prep[k]) cudaEventDestroy(h->ev_prep[k]);
    }
    for (void *d : h->allocs) cudaFree(d);
    delete h;
}

int64_t maexp_learner_num_params(const maexp_learner *h) { return h ? h->P : -1; }
int maexp_learner_num_tensors(const maexp_learner *h) { return h ? (int)h->table.size() : -1; }
int maexp_learner_tensor(const maexp_learner *h, int index, char *name, int name_len, int64_t *offset, int64_t *count) {
    if (!h || index < 0 || index >= (int)h->table.size()) return fail("maexp_learner_tensor: bad index");
    const ParamEntry &e = h->table[index];
    if (name && name_len > 0) {
        strncpy(name, e.name.c_str(), name_len - 1);
        name[name_len - 1] = 0;
    }
    if (offset) *offset = e.offset;
    if (count) *count = e.count;
    return 0;
}

float *maexp_learner_buffer(maexp_learner *h, int which) {
    if (!h) return nullptr;
    switch (which) {
    case MAEXP_BUF_ONLINE: return h->online;
    case MAEXP_BUF_TARGET: return h->target;
    case MAEXP_BUF_GRADS: return h->grads;
    case MAEXP_BUF_SQUARE_AVG: return h->sq;
    case MAEXP_BUF_SUPPORT: return h->support;
    case MAEXP_BUF_EXP_AVG: return h->exp_avg;
    case MAEXP_BUF_ADAM_STEPS: return reinterpret_cast<float *>(h->adam_steps);
    }
    return nullptr;
}

int maexp_learner_bind_positions(maexp_learner *h, const int8_t *pos, const int8_t *next_pos) {
    if (!h) return fail("maexp_learner_bind_positions: null handle");
    h->pos = pos;
    h->next_pos = next_pos;
    return 0;
}

int maexp_learner_bind_canvas_inputs(maexp_learner *h, const uint32_t *obj_rows, const uint32_t *next_obj_rows,
                                     const int8_t *agents, const int8_t *next_agents, const uint32_t *dest_rows,
                                     const uint32_t *next_dest_rows) {
    if (!h) return fail("maexp_learner_bind_canvas_inputs: null handle");
    h->cin[0].obj = obj_rows; h->cin[0].agents = agents; h->cin[0].dest = dest_rows;
    h->cin[1].obj = next_obj_rows; h->cin[1].agents = next_agents; h->cin[1].dest = next_dest_rows;
    return 0;
}

int maexp_learner_set_los(maexp_learner *h, const int8_t *los_host) {
    if (!h) return fail("maexp_learner_set_los: null handle");
    if (!los_host) { h->los = nullptr; return 0; }
    if (h->area_mask != MAEXP_MASK_MERGED) return fail("maexp_learner_set_los: only the merged models read the noise-free area channel");
    const size_t n = (size_t)h->X * h->Y * h->V * h->V;
    if (!h->los && lalloc(h, &h->los, n)) return -1;
    MAEXP_CUDA_L(cudaMemcpy(h->los, los_host, n, cudaMemcpyHostToDevice));
    return 0;
}

int maexp_learner_sync_target(maexp_learner *h, void *stream) {
    if (!h) return fail("maexp_learner_sync_target: null handle");
    MAEXP_CUDA_L(cudaMemcpyAsync(h->target, h->online, sizeof(float) * h->G * h->P, cudaMemcpyDeviceToDevice,
                                 static_cast<cudaStream_t>(stream)));
    return 0;
}

int maexp_learner_q_values(maexp_learner *h, const float *obs, int batch, int use_target, float *q_out, void *stream) {
    if (!h || !obs || !q_out) return fail("maexp_learner_q_values: null argument");
    if (batch < 1 || batch > h->B) return fail("maexp_learner_q_values: batch must be in 1..batch_size");
    if (h->iqn) return fail("maexp_learner_q_values: use maexp_learner_q_values_iqn for IQN");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    maexp_learner::Acts &a = h->acts[1];
    g_splitk_ws = h->splitk_ws;
    const InferenceScope tensor_cores(h->use_tc);
    h->last_q_batch = batch;
    if (h->relative_conv || h->merged_conv || h->merged_da3) {
        float *q = nullptr;
        if (merged_forward(h, st, use_target ? h->target : h->online, obs, h->pos, batch, 1, &q)) return -1;
        MAEXP_CUDA_L(cudaMemcpy2DAsync(q_out, sizeof(float) * batch * h->A, q, sizeof(float) * h->B * h->A,
                                       sizeof(float) * batch * h->A, h->G, cudaMemcpyDeviceToDevice, st));
        MAEXP_CUDA_L(cudaGetLastError());
        return 0;
    }
    if (h->da3) {
        if (h->da6) {
            if (da6_forward(h, st, use_target ? h->target : h->online, obs, h->pos, batch, 1)) return -1;
        } else {
            if (da3_forward(h, st, use_target ? h->target : h->online, obs, batch, 1, true, h->pos)) return -1;
        }
        MAEXP_CUDA_L(cudaMemcpy2DAsync(q_out, sizeof(float) * batch * h->A, h->vloc.a[1].q, sizeof(float) * h->B * h->A,
                                       sizeof(float) * batch * h->A, h->G, cudaMemcpyDeviceToDevice, st));
        MAEXP_CUDA_L(cudaGetLastError());
        return 0;
    }
    g_fused_q = FusedQ();
    if (h->dist) {
        g_fused_q.q_out = q_out;
        g_fused_q.q_g = (long long)batch * h->A;
        g_fused_q.categorical = h->cfg.kind == MAEXP_LEARNER_C51;
    }
    forward(h, st, use_target ? h->target : h->online, obs, batch, a, /*inference=*/true);
    const bool fused_q = g_fused_q.done;
    g_fused_q = FusedQ();
    if (fused_q) {
        // the heads kernel wrote the greedy values
    } else if (!h->dist) {
        // heads are strided by the maximum batch: compact to [G][batch][A]
        MAEXP_CUDA_L(cudaMemcpy2DAsync(q_out, sizeof(float) * batch * h->A, a.out[0], sizeof(float) * h->B * h->A,
                                       sizeof(float) * batch * h->A, h->G, cudaMemcpyDeviceToDevice, st));
    } else {
        dist_q_kernel<<<dim3((batch * 32 + 255) / 256, h->G), 256, 0, st>>>(a.out[0], a.out[1], h->support, batch, h->B, h->A,
                                                                            h->Z, h->cfg.kind == MAEXP_LEARNER_C51, q_out);
    }
    MAEXP_CUDA_L(cudaGetLastError());
    return 0;
}

int maexp_learner_attention_tokens(const maexp_learner *h, int stack) {
    if (!h || !h->da3 || stack < 0 || stack > 1 || (stack == 1 && !h->da6)) return -1;
    return stack ? h->vrel.Tn : h->vloc.Tn;
}

int maexp_learner_attention(maexp_learner *h, int stack, int layer, int batch, float *out, void *stream) {
    if (!h || !out) return fail("maexp_learner_attention: null argument");
    if (!h->da3) return fail("maexp_learner_attention: not an attention model");
    if (stack < 0 || stack > 1 || (stack == 1 && !h->da6)) return fail("maexp_learner_attention: no such transformer stack");
    if (layer < 0 || layer >= h->L_) return fail("maexp_learner_attention: no such block");
    if (batch < 1 || batch > h->B) return fail("maexp_learner_attention: batch must be in 1..batch_size");
    if (h->last_q_batch > kAttnMapsBatch && h->use_tc)
        return fail("maexp_learner_attention: attention maps are kept for q_values batches of at most 256 observations");
    maexp_learner::VitStack &s = stack ? h->vrel : h->vloc;
    const size_t per = (size_t)h->heads_ * s.Tn * s.Tn * sizeof(float);  // one sample
    MAEXP_CUDA_L(cudaMemcpy2DAsync(out, per * batch, s.a[1].blk[layer].probs, per * h->B, per * batch, h->G,
                                   cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    if (h->attn_shrink >= 0.0f) {  // HardShrinkAttention.forward_attn returns the weights after the shrink
        const long long n = (long long)h->G * batch * h->heads_ * s.Tn * s.Tn;
        hard_shrink_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, h->attn_shrink);
        MAEXP_CUDA_L(cudaGetLastError());
    }
    return 0;
}

int maexp_learner_grads(maexp_learner *h, const float *obs, const float *next_obs, const int32_t *actions,
                        const float *rewards, const uint8_t *dones, float *loss_out, void *stream) {
    if (!h || !obs || !next_obs || !actions || !rewards || !dones) return fail("maexp_learner_grads: null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (h->iqn) return fail("maexp_learner_grads: use maexp_learner_grads_iqn for IQN");
    const int G = h->G, B = h->B, A = h->A, Z = h->Z;
    g_splitk_ws = h->splitk_ws;
    if (h->relative_conv || h->merged_conv || h->merged_da3) {  // TD target + MSE on the relative / merged models
        float *q0 = nullptr, *q2 = nullptr;
        cudaStream_t s2 = fork_pass(h, st, 0, true);
        if (merged_forward(h, s2, h->target, next_obs, h->next_pos, B, 2, &q2)) return -1;
        g_splitk_ws = h->splitk_ws;
        if (merged_forward(h, st, h->online, obs, h->pos, B, 0, &q0)) return -1;
        join_pass(h, st, 0);
        dqn_loss_kernel<<<dim3((B + 127) / 128, G), 128, 0, st>>>(q0, q2, actions, rewards, dones, h->cfg.gamma, B, A, h->d_out[0],
                                                                  h->per_sample);
        mean_kernel<<<G, 256, 0, st>>>(h->per_sample, loss_out ? loss_out : h->q_tmp, B);
        merged_backward(h, st);
        MAEXP_CUDA_L(cudaGetLastError());
        return 0;
    }
    if (h->da3) {  // DA3Brain.learn: TD target + MSE on the attention network (da3_brain.py:58-89)
        cudaStream_t s2 = fork_pass(h, st, 0, true);
        if (h->da6) {  // DA6Brain.learn (da6_brain.py:44-75): the same TD target + MSE
            if (da6_forward(h, s2, h->target, next_obs, h->next_pos, B, 2)) return -1;
            g_splitk_ws = h->splitk_ws;
            if (da6_forward(h, st, h->online, obs, h->pos, B, 0)) return -1;
        } else {
            if (da3_forward(h, s2, h->target, next_obs, B, 2, true, h->next_pos)) return -1;
            g_splitk_ws = h->splitk_ws;
            if (da3_forward(h, st, h->online, obs, B, 0, true, h->pos)) return -1;
        }
        join_pass(h, st, 0);
        dqn_loss_kernel<<<dim3((B + 127) / 128, G), 128, 0, st>>>(h->vloc.a[0].q, h->vloc.a[2].q, actions, rewards, dones,
                                                                  h->cfg.gamma, B, A, h->d_out[0], h->per_sample);
        mean_kernel<<<G, 256, 0, st>>>(h->per_sample, loss_out ? loss_out : h->q_tmp, B);
        if (h->da6) da6_backward(h, st);
        else da3_backward(h, st);
        MAEXP_CUDA_L(cudaGetLastError());
        return 0;
    }
    if (!forward_batched(h, st, obs, next_obs, B)) {
        // the training pass is the longest chain: it is issued first (its nodes come first in a captured graph), the side
        // passes fork from the point before it
        stamp(st, 0);
        cudaEventRecord(h->ev_fork, st);
        h->fork_live = h->zero_live = true;
        forward(h, st, h->online, obs, B, h->acts[0]);
        stamp(st, 1);
        cudaStream_t s2 = fork_pass(h, st, 0, false);
        stamp(s2, 2);
        // The distributional nets run two inference passes, target(s') and online(s') (double-DQN action selection), each on its
        // own side stream.  MAEXP_MERGE_INFERENCE_PASSES=1 sends both down ONE branch as grid slices of one trunk launch and one
        // heads launch: under the CUPTI timeline the last of three branches starts ~55 us late and merging looks attractive, but
        // the unprofiled update gets slower (247 against 220 us per C51 team update), so the three branches stay.
        bool both = false;
        if (h->dist && h->has_conv && h->fused_trunk && h->fused_mlp && h->heads.size() == 2 && getenv("MAEXP_MERGE_INFERENCE_PASSES")) {
            const float *params2[2] = {h->target, h->online}, *inputs2[2] = {next_obs, next_obs};
            maexp_learner::Acts *acts2[2] = {&h->acts[2], &h->acts[1]};
            const int train2[2] = {0, 0};
            if (trunk_forward_multi(h, s2, 2, params2, inputs2, acts2, train2, B)) {
                MlpIn ins2[2];
                for (int q = 0; q < 2; ++q) ins2[q] = {acts2[q]->flat, h->D, (long long)h->B * h->D, B, h->B, h->D};
                float *const hb2[2][2] = {{acts2[0]->h[0], acts2[0]->h[1]}, {acts2[1]->h[0], acts2[1]->h[1]}};
                float *const ob2[2][2] = {{acts2[0]->out[0], acts2[0]->out[1]}, {acts2[1]->out[0], acts2[1]->out[1]}};
                if (!mlp_fused_forward_multi(h, s2, 2, params2, h->heads.data(), 2, ins2, hb2, ob2))
                    for (int q = 0; q < 2; ++q)
                        for (size_t k = 0; k < h->heads.size(); ++k)
                            mlp_forward(h, s2, params2[q], h->heads[k], ins2[q], acts2[q]->h[k], acts2[q]->out[k]);
                both = true;
            }
        }
        if (!both) {
            forward(h, s2, h->target, next_obs, B, h->acts[2], /*inference=*/true);  // passes 1 and 2 never feed a backward pass
            stamp(s2, 3);
            if (h->dist) {  // double-DQN action selection by the online net
                cudaStream_t s1 = fork_pass(h, st, 1, false);
                stamp(s1, 4);
                forward(h, s1, h->online, next_obs, B, h->acts[1], /*inference=*/true);
                stamp(s1, 5);
            }
        }
        join_pass(h, st, 0);
        if (h->dist && !both) join_pass(h, st, 1);
        h->fork_live = false;
    }
    if (!h->dist) {
        dqn_loss_kernel<<<dim3((B + 127) / 128, G), 128, 0, st>>>(h->acts[0].out[0], h->acts[2].out[0], actions, rewards,
                                                                  dones, h->cfg.gamma, B, A, h->d_out[0], h->per_sample);
    } else {
        DistArgs p;
        p.v_s = h->acts[0].out[0]; p.a_s = h->acts[0].out[1];
        p.v_n = h->acts[1].out[0]; p.a_n = h->acts[1].out[1];
        p.v_t = h->acts[2].out[0]; p.a_t = h->acts[2].out[1];
        p.act = actions; p.rew = rewards; p.support = h->support;
        p.dv = h->d_out[0]; p.da = h->d_out[1]; p.per_sample = h->per_sample;
        p.B = B; p.A = A; p.Z = Z;
        p.gamma = h->cfg.gamma; p.v_min = h->cfg.v_min; p.v_max = h->cfg.v_max; p.delta_z = h->cfg.delta_z;
        const dim3 grid((B * 32 + 255) / 256, G);
        if (h->cfg.kind == MAEXP_LEARNER_C51) {
            if (A <= 4 && Z <= 64) c51_loss_kernel<4, 2><<<grid, 256, 0, st>>>(p, h->m_scratch);
            else c51_loss_kernel<kMaxActions, kMaxAtomsPerLane><<<grid, 256, 0, st>>>(p, h->m_scratch);
        }
        else qr_loss_kernel<<<grid, 256, 0, st>>>(p);
    }
    // the batch mean of the loss feeds nothing in the backward pass: backward() issues it on a side stream (off the chain)
    h->mean_dst = loss_out ? loss_out : h->q_tmp;
    stamp(st, 6);
    backward(h, st, obs);
    h->zero_live = false;
    if (h->mean_dst) {
        mean_kernel<<<G, 256, 0, st>>>(h->per_sample, h->mean_dst, B);
        h->mean_dst = nullptr;
    }
    stamp(st, 9);
    MAEXP_CUDA_L(cudaGetLastError());
    return 0;
}

int maexp_learner_q_values_iqn(maexp_learner *h, const float *obs, int batch, int use_target, const float *tau_sel,
                               float *q_out, void *stream) {
    if (!h || !obs || !tau_sel || !q_out) return fail("maexp_learner_q_values_iqn: null argument");
    if (!h->iqn) return fail("maexp_learner_q_values_iqn: not an IQN learner");
    if (batch < 1 || batch > h->B) return fail("maexp_learner_q_values_iqn: batch must be in 1..batch_size");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    g_splitk_ws = h->splitk_ws;
    const InferenceScope tensor_cores(h->use_tc);
    h->last_q_batch = batch;
    if (iqn_forward(h, st, use_target ? h->target : h->online, obs, h->pos, batch, tau_sel, h->Nq, /*per_sample=*/0, 1)) return -1;
    iqn_q_kernel<<<dim3((batch + 127) / 128, h->G), 128, 0, st>>>(h->iq[1].out[0], h->iq[1].out[1], batch, h->A, h->Nq,
                                                                  h->rows_max, h->rows_max * h->A, q_out);
    MAEXP_CUDA_L(cudaGetLastError());
    return 0;
}

int maexp_learner_grads_iqn(maexp_learner *h, const float *obs, const float *next_obs, const int32_t *actions,
                            const float *rewards, const uint8_t *dones, const float *taus, const float *tau_dashes,
                            const float *tau_sel, float *loss_out, void *stream) {
    if (!h || !obs || !next_obs || !actions || !rewards || !taus || !tau_dashes || !tau_sel)
        return fail("maexp_learner_grads_iqn: null argument");
    if (!h->iqn) return fail("maexp_learner_grads_iqn: not an IQN learner");
    (void)dones;  // IQNBrain.learn never masks terminal states (iqn_brain.py:104-106)
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int G = h->G, B = h->B;
    g_splitk_ws = h->splitk_ws;
    cudaEventRecord(h->ev_fork, st);  // training pass first (see maexp_learner_grads)
    if (iqn_forward(h, st, h->online, obs, h->pos, B, taus, h->Nq, 1, 0)) return -1;
    cudaStream_t s1 = fork_pass(h, st, 0, false);
    if (iqn_forward(h, s1, h->online, next_obs, h->next_pos, B, tau_sel, h->Nq, 0, 1)) return -1;
    cudaStream_t s2 = fork_pass(h, st, 1, false);
    if (iqn_forward(h, s2, h->target, next_obs, h->next_pos, B, tau_dashes, h->Nd, 1, 2)) return -1;
    join_pass(h, st, 0);
    join_pass(h, st, 1);
    IqnArgs p;
    p.v0 = h->iq[0].out[0]; p.a0 = h->iq[0].out[1];
    p.v1 = h->iq[1].out[0]; p.a1 = h->iq[1].out[1];
    p.v2 = h->iq[2].out[0]; p.a2 = h->iq[2].out[1];
    p.taus = taus; p.act = actions; p.rew = rewards;
    p.dv = h->d_out[0]; p.da = h->d_out[1]; p.per_sample = h->per_sample;
    p.B = B; p.A = h->A; p.Nq = h->Nq; p.Nd = h->Nd;
    p.gV = h->rows_max; p.gA = h->rows_max * h->A; p.gV2 = h->rows_max; p.gA2 = h->rows_max * h->A;
    p.gT = (long long)B * h->Nq;
    p.gamma = h->cfg.gamma;
    iqn_loss_kernel<<<dim3((B * 32 + 255) / 256, G), 256, 0, st>>>(p);
    mean_kernel<<<G, 256, 0, st>>>(h->per_sample, loss_out ? loss_out : h->q_tmp, B);
    iqn_backward(h, st, h->Nq);
    MAEXP_CUDA_L(cudaGetLastError());
    return 0;
}

int maexp_learner_apply(maexp_learner *h, float grad_scale, const uint8_t *active, float *grad_norm_out, void *stream) {
    if (!h) return fail("maexp_learner_apply: null handle");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (g_stamps) {  // an eager call prints the stamps of the update before it (e.g. the last graph replay)
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        if (cap == cudaStreamCaptureStatusNone) {
            unsigned long long t[40];
            cudaStreamSynchronize(st);
            cudaMemcpy(t, g_stamps, sizeof(t), cudaMemcpyDeviceToHost);
            const char *name[14] = {"start", "train pass fwd done", "target pass start", "target pass fwd done", "online(s') pass start",
                                    "online(s') pass fwd done", "loss done", "mlp data-grad chain done", "trunk backward done",
                                    "all gradients joined", "clip + optimizer done", "mlp_fused_bwd launch", "mlp_fused_bwd done",
                                    "trunk_fused_bwd done"};
            for (int i = 0; i < 14; ++i)
                if (t[i] >= t[0])  // (slots outside the mask stay zero)
                    fprintf(stderr, "  stamp %2d %8.1f us  %s\n", i, (double)((long long)(t[i] - t[0])) * 1e-3, name[i]);
            for (int i = 14; i < 38; i += 2)
                if (t[i] > t[0] && t[i + 1] > t[0])
                    fprintf(stderr, "  stamp %2d %8.1f -> %8.1f us  tensor-core weight-gradient GEMM (+ its split-K reduction)\n", i,
                            (double)((long long)(t[i] - t[0])) * 1e-3, (double)((long long)(t[i + 1] - t[0])) * 1e-3);
        }
    }
    sqnorm_partial_kernel<<<dim3(h->nparts, h->G), 256, 0, st>>>(h->grads, h->partial, h->P, grad_scale);
    if (h->cfg.optimizer == 1) {
        adam_step_kernel<<<(h->G + 63) / 64, 64, 0, st>>>(h->adam_steps, active, h->G);
        clip_adam_kernel<<<dim3(64, h->G), 256, 0, st>>>(h->online, h->grads, h->exp_avg, h->sq, h->partial, h->nparts, h->P,
                                                         grad_scale, h->cfg.max_norm, h->cfg.learning_rate, h->cfg.beta1,
                                                         h->cfg.beta2, h->cfg.eps, active, h->adam_steps, grad_norm_out);
    } else {
        clip_rmsprop_kernel<<<dim3(64, h->G), 256, 0, st>>>(h->online, h->grads, h->sq, h->partial, h->nparts, h->P, grad_scale,
                                                            h->cfg.max_norm, h->cfg.learning_rate, h->cfg.alpha,
                                                            h->cfg.one_minus_alpha, h->cfg.eps, active, grad_norm_out);
    }
    stamp(st, 10);
    MAEXP_CUDA_L(cudaGetLastError());
    return 0;
}

// ---- stand-alone loss / optimizer entry points on caller-owned network outputs (the same kernels as the learner's)
namespace {
struct StreamScratch {  // stream-ordered scratch: safe with concurrent streams and under graph capture
    cudaStream_t st;
    std::vector<void *> ptrs;
    explicit StreamScratch(cudaStream_t s) : st(s) {}
    float *get(size_t floats) {
        void *d = nullptr;
        if (cudaMallocAsync(&d, sizeof(float) * (floats ? floats : 1), st) != cudaSuccess) return nullptr;
        ptrs.push_back(d);
        return static_cast<float *>(d);
    }
    ~StreamScratch() { for (void *d : ptrs) cudaFreeAsync(d, st); }
};
int loss_shape_ok(const char *who, int G, int B, int A, int Z) {
    if (G < 1 || B < 1 || A < 1 || A > kMaxActions) return fail(std::string(who) + ": bad shape (1 <= num_actions <= 8)");
    if (Z < 1 || Z > 32 * kMaxAtomsPerLane) return fail(std::string(who) + ": atoms / quantiles must be in 1..128");
    return 0;
}
}  // namespace

int maexp_loss_dqn(const float *q, const float *q_next_target, const int32_t *actions, const float *rewards, const uint8_t *dones,
                   float gamma, int32_t G, int32_t B, int32_t A, float *loss_out, float *dq_out, void *stream) {
    if (!q || !q_next_target || !actions || !rewards || !dones || !loss_out || !dq_out) return fail("maexp_loss_dqn: null argument");
    if (loss_shape_ok("maexp_loss_dqn", G, B, A, 1)) return -1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamScratch ws(st);
    float *per = ws.get((size_t)G * B);
    if (!per) return fail("maexp_loss_dqn: scratch allocation failed");
    dqn_loss_kernel<<<dim3((B + 127) / 128, G), 128, 0, st>>>(q, q_next_target, actions, rewards, dones, gamma, B, A, dq_out, per);
    mean_kernel<<<G, 256, 0, st>>>(per, loss_out, B);
    MAEXP_CUDA_L(cudaGetLastError());
    return 0;
}

static int dist_loss(const char *who, bool c51, const float *out_s, const float *out_next_online, const float *out_next_target,
                     const int32_t *actions, const float *rewards, const float *support, float gamma, float v_min, float v_max,
                     float delta_z, int G, int B, int A, int Z, float *loss_out, float *d_out, void *stream) {
    if (!out_s || !out_next_online || !out_next_target || !actions || !rewards || !support || !loss_out || !d_out)
        return fail(std::string(who) + ": null argument");
    if (loss_shape_ok(who, G, B, A, Z)) return -1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamScratch ws(st);
    float *per = ws.get((size_t)G * B), *m = c51 ? ws.get((size_t)G * B * Z) : per;
    if (!per || !m) return fail(std::string(who) + ": scratch allocation failed");
    DistArgs p;
    p.v_s = p.v_n = p.v_t = nullptr;  // combined outputs: no dueling inside the kernel
    p.a_s = out_s; p.a_n = out_next_online; p.a_t = out_next_target;
    p.act = actions; p.rew = rewards; p.support = support;
    p.dv = nullptr; p.da = d_out; p.per_sample = per;
    p.B = B; p.A = A; p.Z = Z;
    p.gamma = gamma; p.v_min = v_min; p.v_max = v_max; p.delta_z = delta_z;
    const dim3 grid((B * 32 + 255) / 256, G);
    if (c51) {
        if (p.A <= 4 && p.Z <= 64) c51_loss_kernel<4, 2><<<grid, 256, 0, st>>>(p, m);
        else c51_loss_kernel<kMaxActions, kMaxAtomsPerLane><<<grid, 256, 0, st>>>(p, m);
    }
    else qr_loss_kernel<<<grid, 256, 0, st>>>(p);
    mean_kernel<<<G, 256, 0, st>>>(per, loss_out, B);
    MAEXP_CUDA_L(cudaGetLastError());
    return 0;
}

int maexp_loss_c51(const float *logits, const float *logits_next_online, const float *logits_next_target, const int32_t *actions,
                   const float *rewards, const float *support, float gamma, float v_min, float v_max, float delta_z, int32_t G,
                   int32_t B, int32_t A, int32_t Z, float *loss_out, float *d_logits_out, void *stream) {
    return dist_loss("maexp_loss_c51", true, logits, logits_next_online, logits_next_target, actions, rewards, support, gamma,
                     v_min, v_max, delta_z, G, B, A, Z, loss_out, d_logits_out, stream);
}

int maexp_loss_qr(const float *quantiles, const float *quantiles_next_online, const float *quantiles_next_target,
                  const int32_t *actions, const float *rewards, const float *tau_hats, float gamma, int32_t G, int32_t B, int32_t A,
                  int32_t Z, float *loss_out, float *d_quantiles_out, void *stream) {
    return dist_loss("maexp_loss_qr", false, quantiles, quantiles_next_online, quantiles_next_target, actions, rewards, tau_hats,
                     gamma, 0.0f, 0.0f, 1.0f, G, B, A, Z, loss_out, d_quantiles_out, stream);
}

int maexp_loss_iqn(const float *quantiles, const float *quantiles_next_online, const float *quantiles_next_target,
                   const float *taus, const int32_t *actions, const float *rewards, float gamma, int32_t G, int32_t B, int32_t A,
                   int32_t Nq, int32_t Nd, float *loss_out, float *d_quantiles_out, void *stream) {
    if (!quantiles || !quantiles_next_online || !quantiles_next_target || !taus || !actions || !rewards || !loss_out || !d_quantiles_out)
        return fail("maexp_loss_iqn: null argument");
    if (loss_shape_ok("maexp_loss_iqn", G, B, A, 1) || Nq < 1 || Nd < 1) return fail("maexp_loss_iqn: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamScratch ws(st);
    float *per = ws.get((size_t)G * B);
    if (!per) return fail("maexp_loss_iqn: scratch allocation failed");
    IqnArgs p;
    p.v0 = p.v1 = p.v2 = nullptr;
    p.a0 = quantiles; p.a1 = quantiles_next_online; p.a2 = quantiles_next_target;
    p.taus = taus; p.act = actions; p.rew = rewards;
    p.dv = nullptr; p.da = d_quantiles_out; p.per_sample = per;
    p.B = B; p.A = A; p.Nq = Nq; p.Nd = Nd;
    p.gV = 0; p.gA = (long long)B * Nq * A; p.gV2 = 0; p.gA2 = (long long)B * Nd * A; p.gT = (long long)B * Nq;
    p.gamma = gamma;
    iqn_loss_kernel<<<dim3((B * 32 + 255) / 256, G), 256, 0, st>>>(p);
    mean_kernel<<<G, 256, 0, st>>>(per, loss_out, B);
    MAEXP_CUDA_L(cudaGetLastError());
    return 0;
}

int maexp_clip_rmsprop_step(float *params, const float *grads, float *square_avg, int64_t count, int32_t G, float grad_scale,
                            float max_norm, float learning_rate, float alpha, float eps, const uint8_t *active,
                            float *grad_norm_out, void *stream) {
    if (!params || !grads || !square_avg) return fail("maexp_clip_rmsprop_step: null argument");
    if (count < 1 || G < 1) return fail("maexp_clip_rmsprop_step: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamScratch ws(st);
    const int nparts = 32;
    float *partial = ws.get((size_t)G * nparts);
    if (!partial) return fail("maexp_clip_rmsprop_step: scratch allocation failed");
    sqnorm_partial_kernel<<<dim3(nparts, G), 256, 0, st>>>(grads, partial, count, grad_scale);
    clip_rmsprop_kernel<<<dim3(64, G), 256, 0, st>>>(params, grads, square_avg, partial, nparts, count, grad_scale, max_norm,
                                                     learning_rate, alpha, 1.0f - alpha, eps, active, grad_norm_out);
    MAEXP_CUDA_L(cudaGetLastError());
    return 0;
}

int maexp_learner_bind_grads(maexp_learner *h, float *external) {
    if (!h) return fail("maexp_learner_bind_grads: null handle");
    if (!h->own_grads) h->own_grads = h->grads;
    if (external && (reinterpret_cast<uintptr_t>(external) & 15)) return fail("maexp_learner_bind_grads: the buffer must be 16-byte aligned");
    h->grads = external ? external : h->own_grads;
    return 0;
}

int maexp_learner_apply_data_parallel(maexp_learner *h, const maexp_dp_peers *peers, const uint8_t *active, float *grad_norm_out,
                                      void *stream) {
    if (!h || !peers) return fail("maexp_learner_apply_data_parallel: null argument");
    if (h->cfg.optimizer != 0) return fail("maexp_learner_apply_data_parallel: implemented for RMSprop (use the NCCL path for Adam)");
    if (peers->world_size < 1 || peers->world_size > 8 || peers->rank < 0 || peers->rank >= peers->world_size)
        return fail("maexp_learner_apply_data_parallel: world size must be 1..8");
    if (h->G > 256) return fail("maexp_learner_apply_data_parallel: at most 256 agents");
    const long long N4 = (long long)h->G * h->P / 4, per = (N4 + peers->world_size - 1) / peers->world_size;
    if ((per * 4 + h->P - 1) / h->P + 1 > 8) return fail("maexp_learner_apply_data_parallel: a rank's slice may span at most 8 agents");
    DpPeers pe = {};
    pe.world = peers->world_size; pe.rank = peers->rank;
    for (int q = 0; q < pe.world; ++q) {
        if (!peers->grads[q] || !peers->partials[q] || !peers->signals[q]) return fail("maexp_learner_apply_data_parallel: null peer pointer");
        pe.grads[q] = peers->grads[q]; pe.partials[q] = peers->partials[q]; pe.signals[q] = peers->signals[q];
    }
    pe.mc_grads = peers->multicast_grads;
    if (pe.grads[pe.rank] != h->grads) return fail("maexp_learner_apply_data_parallel: bind this rank's gradient buffer first (maexp_learner_bind_grads)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!h->dp_epoch) {
        int dev = 0, sms = 148, per_sm = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dp_reduce_apply_kernel, 256, 0);
        h->dp_blocks = sms * (per_sm < 2 ? per_sm : 2);  // cooperative: every block resident
        if (h->dp_blocks < 1) return fail("maexp_learner_apply_data_parallel: no resident block");
        if (lalloc(h, &h->dp_epoch, 1) || lalloc(h, &h->dp_block_part, (size_t)h->dp_blocks * 8)) return -1;
    }
    float *params = h->online, *sq = h->sq, *bp = h->dp_block_part, *norm_out = grad_norm_out;
    long long P = h->P;
    int G = h->G;
    uint32_t *epoch = h->dp_epoch;
    float max_norm = h->cfg.max_norm, lr = h->cfg.learning_rate, alpha = h->cfg.alpha, oma = h->cfg.one_minus_alpha, eps = h->cfg.eps;
    void *args[] = {&pe, &params, &sq, &P, &G, &epoch, &bp, &max_norm, &lr, &alpha, &oma, &eps, (void *)&active, &norm_out};
    MAEXP_CUDA_L(cudaLaunchCooperativeKernel((const void *)dp_reduce_apply_kernel, dim3(h->dp_blocks), dim3(256), args, 0, st));
    return 0;
}

int maexp_learner_update(maexp_learner *h, const float *obs, const float *next_obs, const int32_t *actions,
                         const float *rewards, const uint8_t *dones, float *loss_out, void *stream) {
    if (maexp_learner_grads(h, obs, next_obs, actions, rewards, dones, loss_out, stream)) return -1;
    return maexp_learner_apply(h, 1.0f, nullptr, nullptr, stream);
}

}  // extern "C"
