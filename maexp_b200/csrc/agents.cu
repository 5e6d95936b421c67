// Learner-side observation decode and batched epsilon-greedy action selection (sm_100a).
//
//   maexp_decode_relative   ObservationHandler.decode_relative_state  (observation_handler.py:150-216):
//                           pastes each V x V crop onto an [C, X, Y] canvas (0 everywhere, last channel -1);
//                           the reference loops over the batch in Python.
//   maexp_select_actions    DefaultAgent.get_action / get_random_action (default_agent.py:12-19,
//                           abstract_agent.py:31-34) for every (env, agent) at once.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/maexp_b200.h"
#include "philox.cuh"

namespace maexp {
int fail(const std::string &msg);
constexpr uint32_t kStreamEpsilon = 8;

// heatmap_observation (default_environment.py:93,238-247): heat[n][k] += sum over environments of obs[e][n][k], skipping
// the environments that were reset by this call (the reset observation goes through observation_handler.reset, which does
// not count).  Thread = one element k of one agent, coalesced along k; blockIdx.z = a slice of the environments; one
// atomicAdd per thread.  The sums are exact in any order while the observations hold small integers (every view without
// gaussian noise, below 2^24 counts per element); with sensing noise the order of the slices is not fixed.
__global__ void __launch_bounds__(128) heatmap_observation_kernel(const float *obs, long long E, int K, long long env_stride,
                                                                  long long agent_stride, const uint8_t *skip, float *heat) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
    if (k >= K) return;
    const long long per = (E + gridDim.z - 1) / gridDim.z, e0 = per * blockIdx.z, e1 = min(E, e0 + per);
    const float *src = obs + n * agent_stride + k;
    float acc = 0.0f;
#pragma unroll 4
    for (long long e = e0; e < e1; ++e)
        if (!skip || !skip[e]) acc += __ldcs(src + e * env_stride);
    if (e1 > e0) atomicAdd(heat + (size_t)n * K + k, acc);
}

// The same sum with 128-bit loads, for rows that start on 16-byte boundaries (ring slots: slot_floats is a multiple of 8
// and the padding behind C*V*V is readable; canvases with C*X*Y % 4 == 0).  A CTA covers CW float4 columns x RW rows per
// iteration (CW * RW <= 256); every thread keeps four loads of 16 bytes in flight, the RW partial sums meet in shared
// memory and column owners issue the atomics.  HBM-bound: one pass over the slot (310 MB at 65 536 envs x 8 agents).
__global__ void __launch_bounds__(256) heatmap_observation_vec_kernel(const float4 *obs, long long E, int K, int K4, int CW, int RW,
                                                                      long long env_stride4, long long agent_stride4,
                                                                      const uint8_t *skip, float *heat) {
    __shared__ float4 part[256];
    const int c = threadIdx.x % CW, r = threadIdx.x / CW, col = blockIdx.x * CW + c, n = blockIdx.y;
    const bool live = r < RW && col < K4;
    const long long per = (E + gridDim.z - 1) / gridDim.z, e0 = per * blockIdx.z, e1 = min(E, e0 + per);
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (live) {
        const float4 *src = obs + n * agent_stride4 + col;
#pragma unroll 4
        for (long long e = e0 + r; e < e1; e += RW) {
            if (skip && __ldg(skip + e)) continue;
            const float4 v = __ldcs(src + e * env_stride4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    if (live && r == 0) {
        for (int q = 1; q < RW; ++q) {
            const float4 v = part[q * CW + c];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float *dst = heat + (size_t)n * K + 4 * col;
        const float a[4] = {acc.x, acc.y, acc.z, acc.w};
        for (int j = 0; j < 4; ++j)
            if (4 * col + j < K) atomicAdd(dst + j, a[j]);
    }
}

// one thread per canvas element, coalesced along y; HBM-bound: writes 4*C*X*Y bytes per observation.  canvas_stride: floats
// between consecutive rows of the canvas (>= Cout*X*Y: extra channels may follow).  los (only_area only): the [X,Y,V,V]
// line-of-sight table; the area channel is then table - 1 instead of the crop's last channel, i.e. the NOISE-FREE "relative"
// key the merged models read (merged_observtion_handler.py:27,41-47) whatever noise the stored crop carries.
__global__ void decode_relative_kernel(const float *obs, const int8_t *pos, float *canvas, long long rows, int C, int V,
                                       int X, int Y, int obs_stride, int only_area, long long canvas_stride, const int8_t *los) {
    const int R = V / 2, Cout = only_area ? 1 : C;
    const long long total = rows * Cout * X * Y;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i % Y), x = (int)((i / Y) % X);
        const int co = (int)((i / ((long long)X * Y)) % Cout);
        const long long row = i / ((long long)Cout * X * Y);
        const int c = only_area ? C - 1 : co;
        const int px = pos[row * 2], py = pos[row * 2 + 1];
        const int a = x - px + R, b = y - py + R;  // window cell of map cell (x, y); window∩map is what the slices copy
        float v = (c == C - 1) ? -1.0f : 0.0f;
        if (a >= 0 && a < V && b >= 0 && b < V) {
            if (los) v = (float)los[(((long long)px * Y + py) * V + a) * V + b] - 1.0f;  // observation_handler.py:107-108
            else v = obs[row * obs_stride + ((long long)c * V + a) * V + b];
        }
        // merged models use canvas[:, -1:] + 1 (conv_mlp.py:100-108): 1 where visible
        canvas[row * canvas_stride + ((long long)co * X + x) * Y + y] = only_area ? v + 1.0f : v;
    }
}

int decode_relative_launch(const float *obs, const int8_t *positions, float *canvas, long long rows, int channels, int visible_range,
                           int size_x, int size_y, int obs_stride, int only_area_plus_one, long long canvas_stride,
                           const int8_t *los, cudaStream_t st) {
    const long long total = rows * (only_area_plus_one ? 1 : channels) * size_x * size_y;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    decode_relative_kernel<<<(unsigned)blocks, 256, 0, st>>>(obs, positions, canvas, rows, channels, visible_range, size_x, size_y,
                                                             obs_stride, only_area_plus_one, canvas_stride, los);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("maexp_decode_relative: ") + cudaGetErrorString(e));
    return 0;
}

__global__ void select_actions_kernel(const float *q, int8_t *actions, int G, int batch, int A, int env0, int num_envs_total,
                                      float epsilon, uint32_t seed, uint32_t gid0, uint32_t tick, const uint8_t *no_task) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch * G) return;
    const int e = i / G, k = i - e * G;
    const uint4 w = philox4x32_10(seed, gid0 + (uint32_t)(env0 + e), tick, kStreamEpsilon, (uint32_t)k, 0);
    int action;
    if ((no_task && no_task[k]) || unit_float(w.x) < epsilon) {
        action = (int)scale(w.y, (uint32_t)A);  // int(random() * act_size)
    } else {
        const float *row = q + ((size_t)k * batch + e) * A;
        action = 0;
        float best = row[0];
        for (int a = 1; a < A; ++a)
            if (row[a] > best) { best = row[a]; action = a; }  // argmax, first maximum
    }
    actions[(size_t)(env0 + e) * G + k] = (int8_t)action;
    (void)num_envs_total;
}
// U[0,1) floats with torch.rand's mantissa convention (24 random bits): four per Philox call, keyed (seed, 'TAUS'), counter
// (tick, stream 8, quad index, 0)
__global__ void fill_uniform_kernel(float *out, long long n, uint32_t seed, uint32_t tick) {
    const long long quads = (n + 3) / 4;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (long long)gridDim.x * blockDim.x) {
        const uint4 w = philox4x32_10(seed, 0x54415553u /*'TAUS'*/, tick, 8u, (uint32_t)q, (uint32_t)(q >> 32));
        const float v[4] = {unit_float(w.x), unit_float(w.y), unit_float(w.z), unit_float(w.w)};
        for (int j = 0; j < 4 && q * 4 + j < n; ++j) out[q * 4 + j] = v[j];
    }
}
}  // namespace maexp

extern "C" {
using namespace maexp;

int maexp_fill_uniform(float *out, int64_t count, uint32_t seed, uint32_t tick, void *stream) {
    if (!out || count < 1) return fail("maexp_fill_uniform: null or empty buffer");
    long long blocks = ((count + 3) / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    fill_uniform_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, count, seed, tick);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("maexp_fill_uniform: ") + cudaGetErrorString(e));
    return 0;
}

int maexp_decode_relative(const float *obs, const int8_t *positions, float *canvas, int64_t rows, int channels,
                          int visible_range, int size_x, int size_y, int obs_stride, int only_area_plus_one, void *stream) {
    if (!obs || !positions || !canvas) return fail("maexp_decode_relative: null argument");
    if (rows < 1 || channels < 1) return fail("maexp_decode_relative: bad shape");
    return decode_relative_launch(obs, positions, canvas, rows, channels, visible_range, size_x, size_y, obs_stride,
                                  only_area_plus_one, (long long)(only_area_plus_one ? 1 : channels) * size_x * size_y, nullptr,
                                  static_cast<cudaStream_t>(stream));
}

int maexp_heatmap_observation_add(const float *obs, int64_t num_envs, int num_agents, int elems, int64_t env_stride,
                                  int64_t agent_stride, const uint8_t *skip, float *heat, void *stream) {
    if (!obs || !heat) return fail("maexp_heatmap_observation_add: null argument");
    if (num_envs < 1 || num_agents < 1 || elems < 1) return fail("maexp_heatmap_observation_add: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int K4 = (elems + 3) / 4;
    if ((reinterpret_cast<uintptr_t>(obs) & 15) == 0 && env_stride % 4 == 0 && agent_stride % 4 == 0 &&
        (long long)K4 * 4 <= env_stride) {  // 128-bit path
        const int CW = K4 < 64 ? K4 : 64, RW = 256 / CW, xt = (K4 + CW - 1) / CW;
        long long slices = (148LL * 8 + (long long)xt * num_agents - 1) / ((long long)xt * num_agents);
        const long long max_slices = (num_envs + RW * 8 - 1) / (RW * 8);  // at least 8 iterations per thread
        if (slices > max_slices) slices = max_slices;
        if (slices < 1) slices = 1;
        if (slices > 65535) slices = 65535;
        heatmap_observation_vec_kernel<<<dim3((unsigned)xt, (unsigned)num_agents, (unsigned)slices), 256, 0, st>>>(
            reinterpret_cast<const float4 *>(obs), num_envs, elems, K4, CW, RW, env_stride / 4, agent_stride / 4, skip, heat);
        const cudaError_t ev = cudaGetLastError();
        if (ev != cudaSuccess) return fail(std::string("maexp_heatmap_observation_add: ") + cudaGetErrorString(ev));
        return 0;
    }
    // env slices: enough CTAs to fill the 148 SMs a few times over, at least 64 environments per slice
    const int tiles = (elems + 127) / 128;
    long long slices = (148LL * 8 + (long long)tiles * num_agents - 1) / ((long long)tiles * num_agents);
    if (slices > (num_envs + 63) / 64) slices = (num_envs + 63) / 64;
    if (slices < 1) slices = 1;
    if (slices > 65535) slices = 65535;
    heatmap_observation_kernel<<<dim3((unsigned)tiles, (unsigned)num_agents, (unsigned)slices), 128, 0, st>>>(
        obs, num_envs, elems, env_stride, agent_stride, skip, heat);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("maexp_heatmap_observation_add: ") + cudaGetErrorString(e));
    return 0;
}

int maexp_select_actions(const float *q_values, int8_t *actions, int num_agents, int batch, int num_actions, int env_offset,
                         float epsilon, uint32_t seed, uint32_t first_global_env_id, uint32_t tick, const uint8_t *no_task,
                         void *stream) {
    if (!q_values || !actions) return fail("maexp_select_actions: null argument");
    select_actions_kernel<<<(batch * num_agents + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        q_values, actions, num_agents, batch, num_actions, env_offset, 0, epsilon, seed, first_global_env_id, tick, no_task);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("maexp_select_actions: ") + cudaGetErrorString(e));
    return 0;
}
}
