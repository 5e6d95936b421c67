// GPU experiment (not part of the product): can the env step's stores hide behind its sequential phase?
// Launch shape of env_step_kernel at the benchmark size.  The sequential phase is emulated by a dependent shared-memory
// pointer chase on 16 lanes (latency-bound, goes through the same LSU pipe as the stores), CHASE steps per agent.
//   mode 0: chase, then the agent's 20 store instructions in one burst      (what the kernel does today)
//   mode 1: the previous agent's stores dripped into the chase, one store instruction every CHASE/20 steps
//   mode 2: chase only      mode 3: stores only
//   mode 4: chase, then the agent's 9728 bytes leave as four 2432-byte bulk async copies (cp.async.bulk, TMA engine) from a
//           per-warp shared-memory staging tile written with 128-bit shared stores     mode 5: mode 4 without the chase
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128, 7) overlap(float *dst, int E, int N, int slot, int chase, int mode, int *sink) {
    __shared__ int ring[4][64];
    __shared__ __align__(128) float stage[4][608];  // 2432 bytes per warp
    __shared__ __align__(128) float stage2[4][1216];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < 64; i += 32) ring[warp][i] = (i * 37 + 11) & 63;
    __syncwarp();
    const int groups = E / 16;
    const float4 v = make_float4(1.0f, 0.0f, -1.0f, 0.0f);
    int idx = lane;
    auto store = [&](float *base, int k) {  // store instruction k of 20 of one agent
        if (k < 16) __stcs(reinterpret_cast<float4 *>(base + k * slot + lane * 4), v);
        else {
            const int r = lane & 7, e = (lane >> 3) * 4 + (k - 16);
            if (r < slot / 4 - 32) __stcs(reinterpret_cast<float4 *>(base + e * slot + (32 + r) * 4), v);
        }
    };
    for (int g = blockIdx.x * 4 + warp; g < groups; g += gridDim.x * 4) {
        for (int i = 0; i <= N; ++i) {
            float *cur = dst + ((size_t)i * E + (size_t)g * 16) * slot;
            float *prev = dst + ((size_t)(i - 1) * E + (size_t)g * 16) * slot;
            if (mode == 1) {
                const int per = chase / 20;
                for (int k = 0; k < 20; ++k) {
                    if (i < N && lane < 16)
                        for (int s = 0; s < per; ++s) idx = ring[warp][idx];
                    __syncwarp();
                    if (i > 0) store(prev, k);
                }
            } else if (mode >= 6) {
                // mode 6: chase; envs 0..11 by direct stores; envs 12..15 staged in shared memory and sent as ONE bulk async
                // copy that is only waited for before the next agent's staging writes: it drains behind the next chase
                // mode 7: the same with 8 staged envs (needs twice the staging tile)
                if (i == N) continue;
                // mode 8: only the 4 staged envs are written (25 % of the bytes, all asynchronous); mode 9: only the 12 direct ones
                const int staged = (mode == 6 || mode >= 8) ? 4 : 8, direct = 16 - staged;
                const bool do_direct = mode != 8, do_bulk = mode != 9;
                if (lane < 16)
                    for (int s = 0; s < chase; ++s) idx = ring[warp][idx];
                __syncwarp();
                if (do_direct)
                    for (int e = 0; e < direct; ++e) __stcs(reinterpret_cast<float4 *>(cur + e * slot + lane * 4), v);
                if (do_direct && (lane & 7) < 6)
                    for (int e = (lane >> 3); e < direct; e += 4) __stcs(reinterpret_cast<float4 *>(cur + e * slot + (32 + (lane & 7)) * 4), v);
                if (!do_bulk) continue;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                for (int q = lane; q < staged * 38; q += 32) reinterpret_cast<float4 *>(stage2[warp])[q] = v;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    const unsigned src = (unsigned)__cvta_generic_to_shared(stage2[warp]);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(cur + direct * slot), "r"(src),
                                 "r"(staged * 608) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else if (mode >= 4) {
                if (i == N) continue;
                if (mode == 4 && lane < 16)
                    for (int s = 0; s < chase; ++s) idx = ring[warp][idx];
                __syncwarp();
                for (int c = 0; c < 4; ++c) {  // four envs per chunk: 152 quads on 32 lanes
                    for (int q = lane; q < 152; q += 32) reinterpret_cast<float4 *>(stage[warp])[q] = v;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        const unsigned src = (unsigned)__cvta_generic_to_shared(stage[warp]);
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(cur + c * 608), "r"(src),
                                     "r"(2432) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    __syncwarp();
                }
            } else if (i < N) {
                if (mode != 3 && lane < 16)
                    for (int s = 0; s < chase; ++s) idx = ring[warp][idx];
                __syncwarp();
                if (mode != 2)
                    for (int k = 0; k < 20; ++k) store(cur, k);
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (idx == 12345) *sink = idx;
}

int main(int argc, char **argv) {
    const int E = 65536, N = 8, slot = argc > 1 ? atoi(argv[1]) : 152;
    const int ctas_per_sm = argc > 2 ? atoi(argv[2]) : 7;
    float *buf[4];
    int *sink;
    const size_t bytes = (size_t)E * N * slot * 4;
    for (auto &b : buf) cudaMalloc(&b, bytes);
    cudaMalloc(&sink, 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int chase : {260})
        for (int mode : {2, 3, 0}) {
            for (int k = 0; k < 5; ++k) overlap<<<148 * ctas_per_sm, 128>>>(buf[k & 3], E, N, slot, chase, mode, sink);
            cudaEventRecord(a);
            const int reps = 50;
            for (int k = 0; k < reps; ++k) overlap<<<148 * ctas_per_sm, 128>>>(buf[k & 3], E, N, slot, chase, mode, sink);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            printf("chase %3d mode %d (%s): %.2f us\n", chase, mode,
                   mode == 0 ? "chase then burst" : mode == 1 ? "dripped stores  " : mode == 2 ? "chase only      " : mode == 3 ? "stores only     " : mode == 4 ? "chase then bulk " : mode == 5 ? "bulk only       " : mode == 6 ? "12 direct+4 bulk" : mode == 7 ? "8 direct+8 bulk " : mode == 8 ? "4 bulk only     " : "12 direct only  ",
                   ms * 1e3 / reps);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
