// Host interface of the tensor-core conv trunk (tc_trunk.cu) used by learner.cu for the acting / evaluation batches.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace maexp {

struct TcTrunkFwd {
    const float *obs;     // [G][B][stride] observations (C * V * V floats used per row)
    const float *params;  // flat parameters, group stride P
    float *flat;          // [G][.][F2 * 4] conv2 output in PyTorch's flatten order, group stride g_flat
    long long stride, g_obs, P, o_c1w, o_c1b, o_c2w, o_c2b, g_flat;
    int B, C, V, F1, F2;
    float *workspace;     // tc_trunk_workspace_floats() floats, 16-byte aligned: the pre-split conv2 weight panels
    float *p_out;         // training pass: pooled map [G][.][9][32] and pooling argmax for the backward pass (group pitch g_p), or nullptr
    unsigned char *arg_out;
    long long g_p;
};
size_t tc_trunk_workspace_floats(int G);
// false: shape outside what the kernel handles (visible_range 7, 32 / 64 filters); the caller falls back to the CUDA-core kernel
bool tc_trunk_forward(cudaStream_t st, int G, const TcTrunkFwd &a);

}  // namespace maexp
