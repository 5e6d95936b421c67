// GPU experiment (not part of the product): what is the write ceiling of the env step's store pattern?
// Same launch shape as env_step_kernel at the benchmark size (persistent CTAs of 4 warps, 7 per SM, one group of 16 envs
// per warp): every warp writes, for each of 8 agents, the 16 consecutive 608-byte agent slots of its group with the same
// 128-bit streaming stores (lane = quad), with WAIT_NS of idle time before each agent's burst (the sequential phase).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_exp/store_pattern tools/exp_store_pattern.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128, 7) pattern(float *dst, int E, int N, int slot, int wait_ns, int mode) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int groups = E / 16;
    const float4 v = make_float4(1.0f, 0.0f, -1.0f, 0.0f);
    for (int g = blockIdx.x * 4 + warp; g < groups; g += gridDim.x * 4) {
        for (int i = 0; i < N; ++i) {
            if (wait_ns) {
                unsigned long long t0, t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < (unsigned long long)wait_ns);
            }
            float *base = dst + ((size_t)i * E + (size_t)g * 16) * slot;
            if (mode == 0) {  // the kernel's pattern: lane owns a quad, 16 envs each; 6 trailing quads on 24 lanes
                for (int e = 0; e < 16; ++e) __stcs(reinterpret_cast<float4 *>(base + e * slot + lane * 4), v);
                const int r = lane & 7, e0 = (lane >> 3) * 4;
                if (r < 6)
                    for (int e = 0; e < 4; ++e) __stcs(reinterpret_cast<float4 *>(base + (e0 + e) * slot + (32 + r) * 4), v);
            } else {  // flat: the group's 16 slots are one contiguous run of 16*slot floats
                const int quads = 16 * slot / 4;
                for (int q = lane; q < quads; q += 32) __stcs(reinterpret_cast<float4 *>(base) + q, v);
            }
        }
    }
}

int main(int argc, char **argv) {
    const int E = 65536, N = 8, slot = 152;
    float *buf[4];
    const size_t bytes = (size_t)E * N * slot * 4;
    for (auto &b : buf) cudaMalloc(&b, bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 2; ++mode)
        for (int wait : {0, 1000, 2000, 3000, 4000, 5000}) {
            for (int ctas : {148 * 7, 148 * 4}) {
                for (int k = 0; k < 5; ++k) pattern<<<ctas, 128>>>(buf[k & 3], E, N, slot, wait, mode);
                cudaEventRecord(a);
                const int reps = 50;
                for (int k = 0; k < reps; ++k) pattern<<<ctas, 128>>>(buf[k & 3], E, N, slot, wait, mode);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms = 0;
                cudaEventElapsedTime(&ms, a, b);
                const double us = ms * 1e3 / reps;
                printf("mode %d wait_ns %5d ctas %4d: %.2f us  (%.0f GB/s; minus the %d waits: %.2f us)\n", mode, wait, ctas, us,
                       bytes / us / 1e3, N, us - N * wait / 1e3);
            }
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
