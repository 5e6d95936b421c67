// Host interface of the TMA-fed tcgen05 MLP-heads kernel (tc_heads.cu) used by learner.cu for the acting / evaluation
// batches (>= 512 rows per agent).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace maexp {

struct TcHeadsFwd {
    const float *x;          // input rows [G][rows][x_row] (16-byte aligned, x_row and x_g multiples of 4 floats)
    long long x_row, x_g;
    const float *params;     // flat parameters, group stride P
    long long P;
    long long w1[2], b1[2], w2[2], b2[2];  // offsets of fc_list.0 / fc_post of every head (mlp.py:19-26)
    int out[2], nheads;
    float *outp[2];          // [G][rows_max][out]
    float *hbuf[2];          // relu(x W1^T + b1) [G][rows_max][128] of every head (training pass), or nullptr
    int rows;
    long long rows_max;
    int din, H;
    float *workspace;        // tc_heads_workspace_floats() floats, 16-byte aligned: the pre-split weight panels
    // optional fused greedy-value epilogue of the dueling distributional heads (head 0 = fc_V [Z], head 1 = fc_A [A][Z]):
    // q_out [G][rows][A] (group stride q_g) instead of the raw head outputs; *fused_q_done tells the caller whether it ran
    float *q_out;
    long long q_g;
    int A, Z, categorical;
    const float *support;
    bool *fused_q_done;
};
size_t tc_heads_workspace_floats(int G, int nheads, int din, const int *out);
// false: shape outside what the kernel handles (the caller falls back to the per-layer tensor-core GEMMs)
bool tc_heads_forward(cudaStream_t st, int G, const TcHeadsFwd &a);
// n <= 3 passes of the same shapes as grid slices of one launch (the forward passes of an update)
struct TcHeadsPrepBranch { cudaStream_t stream; cudaEvent_t after, done; };  // where the weight-panel kernels run (see tc_heads.cu)
bool tc_heads_forward_multi(cudaStream_t st, int G, const TcHeadsFwd *passes, int n, const TcHeadsPrepBranch *prep = nullptr);

}  // namespace maexp
