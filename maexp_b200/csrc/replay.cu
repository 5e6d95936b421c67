// Fused replay sampling + gather on the device-resident ring (sm_100a).
//
// Replaces ReplayBuffer.sample + RLDataset.__iter__ + DataLoader default collate
// (core/utils/buffer.py:60-63, core/utils/dataset.py:24-29, abstract_trainer.py:93-97): one launch
// picks `batch` whole-team transitions and writes the per-agent batches the learner consumes.
// One warp per (sample, agent); rows are moved with 128-bit loads/stores.  The picks are DISTINCT, as
// np.random.choice(len, batch, replace=False) in buffer.py:61: every CTA recomputes the whole pick list in shared memory
// (one Philox call per sample, a duplicate scan of batch^2 / 2 compares spread over the CTA) -- sequential rejection
// sampling, sample b redraws while it equals a pick of a lower index -- so that no second launch, index buffer or
// grid-wide barrier is needed and the result depends on (seed, tick) only.  Batches above kUniqueMaxBatch keep the
// recompute-per-warp path and sample with replacement.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/maexp_b200.h"
#include "philox.cuh"

namespace maexp {
int fail(const std::string &msg);

constexpr uint32_t kStreamReplay = 7;
constexpr int kMaxRedraws = 64;
constexpr int kMaxProbes = 4096;
constexpr int kUniqueMaxBatch = 1024;  // distinct picks up to this batch size (the reference's batches are 32 .. 256)
constexpr int kMaxUniqueRounds = 256;  // redraws of duplicates per launch before the sampler gives up and says so

// state slot of the transition of age `age` (state = head-1-age, next = head-age); false when the next slot is a reset
__device__ __forceinline__ bool valid_transition(const maexp_replay_view &r, int age, int env, int *slot) {
    int s = (r.head - 1 - age) % r.num_slots;
    if (s < 0) s += r.num_slots;
    *slot = s;
    return !r.was_reset[(size_t)((s + 1) % r.num_slots) * r.num_envs + env];
}

// sample b's next valid transition, continuing its draw sequence at d (d-th Philox draw of the sample: transition age a in
// [0, count-2] -- state slot = head-1-a, next slot = head-a (mod S) -- and environment); false: neither the redraws nor the
// fixed walk found a transition
__device__ bool draw_pick(const maexp_replay_view &r, uint32_t seed, uint32_t tick, int b, uint32_t &d, int &slot, int &env) {
    bool ok = false;
    int age = 0;
    for (; d < (uint32_t)kMaxRedraws && !ok; ++d) {
        const uint4 w = philox4x32_10(seed, 0x52504C59u /*'RPLY'*/, tick, kStreamReplay, d, (uint32_t)b);
        age = (int)scale(w.x, (uint32_t)(r.count - 1));
        env = (int)scale(w.y, (uint32_t)r.num_envs);
        ok = valid_transition(r, age, env, &slot);
    }
    // every redraw hit a transition into a reset slot (tiny rings whose envs reset in lock-step): walk the
    // (age, env) pairs after the last draw in a fixed order
    for (int probe = 0; probe < kMaxProbes && !ok; ++probe) {
        if (++age >= r.count - 1) {
            age = 0;
            if (++env >= r.num_envs) env = 0;
        }
        ok = valid_transition(r, age, env, &slot);
    }
    return ok;
}

__global__ void replay_sample_kernel(const maexp_replay_view r, int batch, uint32_t seed, uint32_t tick,
                                     const int32_t *indices, const maexp_batch out) {
    // unique path: [batch] keys (slot << 32 | env) of every sample, [cap] hash table (cap = the power of two >= 2 * batch), [batch]
    // draw counters
    extern __shared__ unsigned long long s_key[];
    __shared__ int s_first, s_flags, s_any;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int N = r.num_agents;
    const bool unique = !indices && batch <= kUniqueMaxBatch;
    if (unique) {  // (before the early exit below: every thread of the CTA takes part)
        int cap = 64;
        while (cap < 2 * batch) cap <<= 1;
        unsigned long long *s_tab = s_key + batch;
        uint32_t *s_d = reinterpret_cast<uint32_t *>(s_tab + cap);
        if (threadIdx.x == 0) s_flags = s_any = 0;
        for (int q = threadIdx.x; q < cap; q += blockDim.x) s_tab[q] = ~0ull;
        __syncthreads();
        for (int q = threadIdx.x; q < batch; q += blockDim.x) {
            uint32_t d = 0;
            int sl = 0, en = 0;
            if (draw_pick(r, seed, tick, q, d, sl, en)) s_key[q] = ((unsigned long long)(uint32_t)sl << 32) | (uint32_t)en;
            else {
                s_key[q] = ~0ull - (unsigned long long)q;  // no transition in the ring: flagged, never equal to another key
                atomicOr(&s_flags, MAEXP_REPLAY_ERR_NO_VALID);
            }
            s_d[q] = d;
        }
        __syncthreads();
        // does any pick repeat?  (whether one does is independent of the insertion order; almost always none does, and the
        // batch^2 / 2 compares below are skipped)
        for (int q = threadIdx.x; q < batch; q += blockDim.x) {
            const unsigned long long k = s_key[q];
            if ((k >> 32) >= (unsigned long long)r.num_slots) continue;  // (the no-transition marker)
            uint32_t hsh = ((uint32_t)(k >> 32) * 0x9E3779B1u + (uint32_t)k * 0x85EBCA6Bu) & (uint32_t)(cap - 1);
            for (;;) {
                const unsigned long long prev = atomicCAS(&s_tab[hsh], ~0ull, k);
                if (prev == ~0ull) break;
                if (prev == k) { s_any = 1; break; }
                hsh = (hsh + 1) & (uint32_t)(cap - 1);
            }
        }
        __syncthreads();
        for (int round = 0; s_any && round <= kMaxUniqueRounds; ++round) {
            if (threadIdx.x == 0) s_first = batch;
            __syncthreads();
            for (int q = threadIdx.x; q < batch; q += blockDim.x) {  // the lowest sample that repeats a pick of a lower index
                const unsigned long long k = s_key[q];
                bool dup = false;
                for (int j = 0; j < q && !dup; ++j) dup = s_key[j] == k;
                if (dup) atomicMin(&s_first, q);
            }
            __syncthreads();
            const int first = s_first;
            if (first >= batch) break;
            if (round == kMaxUniqueRounds) {  // fewer distinct transitions than samples (np.random.choice raises here)
                if (threadIdx.x == 0) s_flags |= MAEXP_REPLAY_ERR_NOT_UNIQUE;
                __syncthreads();
                break;
            }
            if (threadIdx.x == 0) {  // every sample below `first` is final: `first` draws again
                uint32_t d = s_d[first];
                int sl = 0, en = 0;
                if (draw_pick(r, seed, tick, first, d, sl, en)) s_key[first] = ((unsigned long long)(uint32_t)sl << 32) | (uint32_t)en;
                s_d[first] = d;
            }
            __syncthreads();
        }
    }
    if (warp >= batch * N) return;
    const int b = warp / N, agent = warp - b * N;
    int slot = 0, env = 0, flags = 0;
    if (unique) {
        const unsigned long long k = s_key[b];
        flags = s_flags;
        if ((k >> 32) < (unsigned long long)r.num_slots) {  // (else the no-transition marker: slot 0, env 0, flagged)
            slot = (int)(k >> 32);
            env = (int)(uint32_t)k;
        }
    } else if (indices) {
        slot = indices[2 * b];
        env = indices[2 * b + 1];
        if (slot < 0 || slot >= r.num_slots || env < 0 || env >= r.num_envs) {  // never read outside the ring
            flags |= MAEXP_REPLAY_ERR_INDEX_RANGE;
            slot = min(max(slot, 0), r.num_slots - 1);
            env = min(max(env, 0), r.num_envs - 1);
        }
        if (r.was_reset[(size_t)((slot + 1) % r.num_slots) * r.num_envs + env]) flags |= MAEXP_REPLAY_ERR_INDEX_INVALID;
    } else {  // batches above kUniqueMaxBatch: the pick recomputed per warp (the N warps of a sample take the same draws)
        uint32_t d = 0;
        if (!draw_pick(r, seed, tick, b, d, slot, env)) flags |= MAEXP_REPLAY_ERR_NO_VALID;  // the row below is NOT a transition
    }
    if (flags && lane == 0 && out.error_flag) atomicOr(out.error_flag, flags);
    const int nxt = (slot + 1) % r.num_slots;
    const size_t row = (size_t)r.num_envs * N * r.slot_floats;
    // ring slots are agent-major: [S][N][E][slot_floats]
    const float4 *s0 = reinterpret_cast<const float4 *>(r.obs + (size_t)slot * row + ((size_t)agent * r.num_envs + env) * r.slot_floats);
    const float4 *s1 = reinterpret_cast<const float4 *>(r.obs + (size_t)nxt * row + ((size_t)agent * r.num_envs + env) * r.slot_floats);
    const size_t orow = ((size_t)agent * batch + b) * r.slot_floats;
    float4 *d0 = reinterpret_cast<float4 *>(out.obs + orow), *d1 = reinterpret_cast<float4 *>(out.next_obs + orow);
    // all loads of two strides in flight before the first store (the rows are random 600-byte gathers: latency-bound
    // without memory-level parallelism)
    const int n4 = r.slot_floats / 4;
    for (int q = lane; q < n4; q += 64) {
        const int q2 = q + 32;
        const float4 a0 = __ldg(s0 + q), b0 = __ldg(s1 + q);
        float4 a1 = a0, b1 = b0;
        if (q2 < n4) {
            a1 = __ldg(s0 + q2);
            b1 = __ldg(s1 + q2);
        }
        d0[q] = a0;
        d1[q] = b0;
        if (q2 < n4) {
            d0[q2] = a1;
            d1[q2] = b1;
        }
    }
    if (out.obj_rows && r.objects_bits) {  // agent's objects_matrix snapshot of both slots: T*X strided words each
        const int TX = r.type_objects * r.size_x;
        const size_t per_slot = (size_t)N * TX * r.num_envs;
        const uint32_t *s0b = r.objects_bits + (size_t)slot * per_slot + (size_t)agent * TX * r.num_envs + env;
        const uint32_t *s1b = r.objects_bits + (size_t)nxt * per_slot + (size_t)agent * TX * r.num_envs + env;
        const size_t ob = ((size_t)agent * batch + b) * TX;
        for (int q = lane; q < TX; q += 32) {
            out.obj_rows[ob + q] = __ldg(s0b + (size_t)q * r.num_envs);
            out.next_obj_rows[ob + q] = __ldg(s1b + (size_t)q * r.num_envs);
        }
    }
    if (out.dest_rows && r.destination && r.dest_table) {
        const int d0 = r.destination[((size_t)slot * r.num_envs + env) * N + agent];
        const int d1 = r.destination[((size_t)nxt * r.num_envs + env) * N + agent];
        const size_t ob = ((size_t)agent * batch + b) * r.size_x;
        for (int q = lane; q < r.size_x; q += 32) {
            out.dest_rows[ob + q] = __ldg(r.dest_table + (size_t)d0 * r.size_x + q);
            out.next_dest_rows[ob + q] = __ldg(r.dest_table + (size_t)d1 * r.size_x + q);
        }
    }
    if (lane == 0) {
        const size_t m = ((size_t)nxt * r.num_envs + env) * N + agent, o = (size_t)agent * batch + b;
        out.actions[o] = max((int)r.actions[m], 0);  // -1 marks a reset slot (flagged above): the loss kernels index with it
        out.rewards[o] = r.rewards[m];
        out.dones[o] = r.dones[m];
        if (out.pos && r.positions) {
            const size_t ps = (((size_t)slot * r.num_envs + env) * N + agent) * 2, pn = (((size_t)nxt * r.num_envs + env) * N + agent) * 2;
            out.pos[o * 2] = r.positions[ps];
            out.pos[o * 2 + 1] = r.positions[ps + 1];
            out.next_pos[o * 2] = r.positions[pn];
            out.next_pos[o * 2 + 1] = r.positions[pn + 1];
        }
        if (out.agents && r.positions) {  // every agent's cell (decode_agents_channel, observation_handler.py:218-250)
            const int8_t *ps = r.positions + ((size_t)slot * r.num_envs + env) * N * 2;
            const int8_t *pn = r.positions + ((size_t)nxt * r.num_envs + env) * N * 2;
            for (int k = 0; k < 2 * N; ++k) {
                out.agents[o * 2 * N + k] = ps[k];
                out.next_agents[o * 2 * N + k] = pn[k];
            }
        }
        if (out.picked && agent == 0) {
            out.picked[2 * b] = slot;
            out.picked[2 * b + 1] = env;
        }
    }
}
}  // namespace maexp

extern "C" int maexp_replay_sample(const maexp_replay_view *ring, int32_t batch, uint32_t seed, uint32_t tick,
                                   const int32_t *indices, const maexp_batch *out, void *stream) {
    using namespace maexp;
    if (!ring || !out) return fail("maexp_replay_sample: null argument");
    if (batch < 1) return fail("maexp_replay_sample: batch < 1");
    if (!indices && ring->count < 2) return fail("maexp_replay_sample: the ring holds fewer than 2 slots");
    if (ring->slot_floats % 4) return fail("maexp_replay_sample: slot_floats must be a multiple of 4");
    if (!ring->obs || !ring->actions || !ring->rewards || !ring->dones || !ring->was_reset || !out->obs ||
        !out->next_obs || !out->actions || !out->rewards || !out->dones)
        return fail("maexp_replay_sample: null buffer");
    const int warps = batch * ring->num_agents;
    size_t smem = 0;  // unique path: keys + hash table + draw counters
    if (!indices && batch <= kUniqueMaxBatch) {
        int cap = 64;
        while (cap < 2 * batch) cap <<= 1;
        smem = (size_t)batch * 12 + (size_t)cap * 8;
    }
    replay_sample_kernel<<<(warps * 32 + 255) / 256, 256, smem, static_cast<cudaStream_t>(stream)>>>(*ring, batch, seed, tick,
                                                                                                    indices, *out);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string("maexp_replay_sample: ") + cudaGetErrorString(e));
    return 0;
}
