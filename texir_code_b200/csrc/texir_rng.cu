// texir_rng.cu -- the reference's Cranley-Patterson shift stream, regenerated on the device.
//
// The reference draws the shift of every generate_dir call as  torch.rand(b, 1, 2)  on the CPU default generator and
// moves it to the GPU (utils/sample_util.py:102).  At the Mat stage's 98 304 pixels that is 196 608 serial mt19937 draws,
// 0.3 ... 0.7 ms of host time per optimisation step -- more than the host spends on everything else of a 16-spp step.
// k_mt19937_uniform produces THE SAME float32 numbers from the same generator state: at::mt19937 (624-word state;
// next_state() = three data-parallel phases of 227 + 227 + 169 words and one last word; tempering) and
// at::uniform_real_distribution<float> ((r & (2^24 - 1)) * 2^-24), one CTA, the state in shared memory; it also
// returns the advanced state so that the host generator can be set to where torch.rand would have left it.
// Restated (and pinned against torch itself) in oracle/mt19937_oracle.py.
#include <cuda_runtime.h>

#include <string>

#include "../../include/texir_b200.h"
#include "texir_host.h"

using texir_host::fail;
using texir_host::finish_launch;

namespace {
constexpr int kN = 624, kM = 397;

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
    return (((u & 0x80000000u) | (v & 0x7FFFFFFFu)) >> 1) ^ ((v & 1u) ? 0x9908B0DFu : 0u);
}
__device__ __forceinline__ float mt_uniform(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= y >> 18;
    return (float)(y & 0xFFFFFFu) * 5.9604644775390625e-8f;  // 2^-24
}

__global__ void __launch_bounds__(256) k_mt19937_uniform(const uint32_t* __restrict__ state_in, int left, int next, int64_t n,
                                                         float* __restrict__ out, uint32_t* __restrict__ state_out) {
    __shared__ uint32_t st[kN];
    const int t = threadIdx.x;
    for (int i = t; i < kN; i += 256) st[i] = state_in[i];
    __syncthreads();
    int64_t produced = 0;
    {  // what is left of the current block of 624 words
        const int avail = left > 1 ? left - 1 : 0;
        const int k = (int)(n < avail ? n : avail);
        for (int i = t; i < k; i += 256) out[i] = mt_uniform(st[next + i]);
        produced = k;
        left -= k;
        next += k;
    }
    while (produced < n) {
        // next_state(): new[i] = x[i + 397 mod 624 (already new for i >= 227)] ^ twist(old[i], old[i + 1])
        uint32_t v = 0;
        if (t < kN - kM) v = st[t + kM] ^ mt_twist(st[t], st[t + 1]);                      // i in [0, 227)
        __syncthreads();
        if (t < kN - kM) st[t] = v;
        __syncthreads();
        if (t < kN - kM) v = st[t] ^ mt_twist(st[t + (kN - kM)], st[t + (kN - kM) + 1]);   // i in [227, 454)
        __syncthreads();
        if (t < kN - kM) st[t + (kN - kM)] = v;
        __syncthreads();
        if (t < kM - (kN - kM) - 1) v = st[t + (kN - kM)] ^ mt_twist(st[t + 2 * (kN - kM)], st[t + 2 * (kN - kM) + 1]);  // [454, 623)
        __syncthreads();
        if (t < kM - (kN - kM) - 1) st[t + 2 * (kN - kM)] = v;
        __syncthreads();
        if (t == 0) st[kN - 1] = st[kM - 1] ^ mt_twist(st[kN - 1], st[0]);                 // i = 623
        __syncthreads();
        const int m = (int)(n - produced < kN ? n - produced : kN);
        for (int i = t; i < m; i += 256) out[produced + i] = mt_uniform(st[i]);
        produced += m;
        left = kN + 1 - m;
        next = m;
    }
    for (int i = t; i < kN; i += 256) state_out[i] = st[i];
    if (t == 0) {
        state_out[kN] = (uint32_t)left;
        state_out[kN + 1] = (uint32_t)next;
    }
}
}  // namespace

extern "C" int texir_mt19937_uniform(const uint32_t* state_in, int left, int next, int64_t n, float* out, uint32_t* state_out,
                                     void* stream) {
    if (!state_in || !out || !state_out || n <= 0) return fail("texir_mt19937_uniform: bad argument");
    if (left < 0 || left > kN + 1 || next < 0 || next > kN || (left > 1 && next + left - 1 > kN))
        return fail("texir_mt19937_uniform: inconsistent generator state");
    texir_host::ScopedDevice guard(texir_host::device_of(out));
    k_mt19937_uniform<<<1, 256, 0, (cudaStream_t)stream>>>(state_in, left, next, n, out, state_out);
    return finish_launch("k_mt19937_uniform");
}
