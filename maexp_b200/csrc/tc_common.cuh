// Shared device helpers of the tcgen05 (5th-generation tensor core) kernels: mbarrier, TF32 split, shared-memory matrix
// descriptors (K-major, SWIZZLE_128B), UMMA issue / commit, staging of fp32 chunks into the swizzled operand tiles.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace maexp_tc {

constexpr int kBM = 128, kBK = 32;  // rows per CTA, K floats per slab (= one 128-byte swizzle row)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity));
        if (spin > (1u << 26)) __trap();  // a lost arrival must not hang the GPU
    }
}
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp): start address >> 4 in bits
// [0,14), leading byte offset [16,30) (unused for swizzled K-major), stride byte offset [32,46) = 1024 B between
// 8-row groups, version 1 in [46,48), layout type 2 (SWIZZLE_128B) in [61,64)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// ---- TMA (cp.async.bulk.tensor) loads into SWIZZLE_128B tiles, completion on an mbarrier; the tensor map lives in the
// ---- kernel's parameter space (const __grid_constant__), its address is what the instruction takes
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const void *tm, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_3d(uint32_t dst, const void *tm, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// instruction descriptor of tcgen05.mma.kind::tf32, M = 128: D = F32 (bits 4-5 = 1), A and B = TF32 (bits 7-9 and 10-12 = 2), both
// K-major, N >> 3 in bits 17-22, M >> 4 in bits 24-28
__device__ __forceinline__ uint32_t idesc_tf32_m128(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// one 16-byte chunk (4 consecutive K values of one row) -> swizzled shared memory, optionally split into hi / lo
template <int SPLIT>
__device__ __forceinline__ void stage_chunk(uint8_t *hi_tile, uint8_t *lo_tile, int r, int c, float4 v) {
    const uint32_t off = (uint32_t)r * 128 + (uint32_t)((c ^ (r & 7)) * 16);
    if (SPLIT == 3) {
        const float4 hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
        *reinterpret_cast<float4 *>(hi_tile + off) = hi;
        *reinterpret_cast<float4 *>(lo_tile + off) = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
    } else {
        *reinterpret_cast<float4 *>(hi_tile + off) = v;
    }
}
__device__ __forceinline__ float4 load_chunk(const float *src, int k, int K, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && k + 3 < K) return *reinterpret_cast<const float4 *>(src);
    if (k + 0 < K) v.x = src[0];
    if (k + 1 < K) v.y = src[1];
    if (k + 2 < K) v.z = src[2];
    if (k + 3 < K) v.w = src[3];
    return v;
}


}  // namespace maexp_tc
