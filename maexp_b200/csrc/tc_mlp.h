// Host interface of the tcgen05 MLP kernels (tc_mlp.cu) used by learner.cu for update batches of >= 128 rows.
#pragma once
#include <cuda_runtime.h>

namespace maexp {

struct TcMlpFwd {
    const float *x;          // input rows [G][rows][x_row]
    long long x_row, x_g;
    const float *params;     // flat parameters, group stride P
    long long P;
    long long w1[2], b1[2], w2[2], b2[2];  // offsets of fc_list.0 / fc_post of every head (mlp.py:19-26)
    int out[2], nheads;
    float *hbuf[2];          // relu(x W1^T + b1) [G][rows_max][H] of every head, or nullptr (inference passes)
    float *outp[2];          // [G][rows_max][out]
    int rows;
    long long rows_max;
    int din, H;
    int dbg;                 // experiment switches (MAEXP_TC_MLP_DBG): 1 no MMA issue, 2 no staging stores, 4 no proxy fence
};
// false: shape outside what the kernel handles (the caller falls back to the SIMT kernel)
bool tc_mlp_forward(cudaStream_t st, int G, const TcMlpFwd &a);

}  // namespace maexp
