// select_kernels.cu -- K3: exact k-th smallest entropy by 3-pass radix select
// (11|11|10 bits of a monotone 32-bit key) + keep-mask emit.
//
// Replaces src/agent.py:403-407 (concatenate on the host, numpy sort, index
// k = int(n*left_ratio)) and the strict-< mask of src/agent.py:148-150,229-232.
// The order is numpy.sort's: ascending, -0 == +0, NaN last.  Nothing is sorted:
// each pass histograms one digit of the keys that still match the prefix found
// so far (HBM/L2-bound streaming reads, 4 B per entropy per pass) and a
// single-block scan picks the digit that contains rank k.  The state lives in
// device memory so the three passes chain on one stream without a host round
// trip; on several GPUs the per-rank histograms are all-gathered (NCCL) between
// the histogram and the scan, which is exact because the counts are integers.
#include "kernels.cuh"
#include "so3_math.cuh"

namespace suhpe {

namespace {

constexpr int kSelThreads = 512;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void digit_of(uint32_t key, int pass, uint32_t prefix, bool& match, int& bin) {
    if (pass == 1)      { match = true;                              bin = key >> kHistShift1; }
    else if (pass == 2) { match = (key >> kHistShift1) == prefix;    bin = (key >> kHistShift2) & (kHistBins2 - 1); }
    else                { match = (key >> kHistShift2) == prefix;    bin = key & (kHistBins3 - 1); }
}

// warp-aggregated shared-memory histogram update: entropies cluster in a few
// bins, so lanes with equal bins elect one leader that adds the group size
__device__ __forceinline__ void hist_add(unsigned int* hist_s, bool match, int bin) {
    const unsigned active = __ballot_sync(kFull, match);
    if (match) {
        const unsigned peers = __match_any_sync(active, bin);
        if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist_s[bin], __popc(peers));
    }
}

__global__ void __launch_bounds__(kSelThreads)
select_hist_kernel(const float* __restrict__ e, long long n, int pass, const SelectState* __restrict__ state,
                   unsigned long long* __restrict__ hist, bool vec_ok) {
    __shared__ unsigned int hist_s[kHistBinsMax];
    for (int i = threadIdx.x; i < kHistBinsMax; i += kSelThreads) hist_s[i] = 0u;
    __syncthreads();
    const uint32_t prefix = (pass > 1) ? state->prefix : 0u;
    const long long tid = (long long)blockIdx.x * kSelThreads + threadIdx.x;
    const long long stride = (long long)gridDim.x * kSelThreads;
    if (vec_ok) {
        const long long n4 = n >> 2;
        const float4* e4 = reinterpret_cast<const float4*>(e);
        // every lane of a warp runs the same number of iterations (ballots inside)
        const long long iters = (n4 + stride - 1) / stride;
        for (long long it = 0; it < iters; ++it) {
            const long long i = tid + it * stride;
            const bool in = i < n4;
            float4 v = in ? __ldg(e4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                bool match; int bin;
                digit_of(entropy_key(vals[c]), pass, prefix, match, bin);
                hist_add(hist_s, match && in, bin);
            }
        }
        // tail (n % 4 elements), handled by the first warp of block 0
        if (blockIdx.x == 0 && threadIdx.x < 32) {
            const long long i = (n4 << 2) + threadIdx.x;
            const bool in = i < n;
            bool match; int bin;
            digit_of(entropy_key(in ? e[i] : 0.f), pass, prefix, match, bin);
            hist_add(hist_s, match && in, bin);
        }
    } else {
        const long long iters = (n + stride - 1) / stride;
        for (long long it = 0; it < iters; ++it) {
            const long long i = tid + it * stride;
            const bool in = i < n;
            bool match; int bin;
            digit_of(entropy_key(in ? __ldg(e + i) : 0.f), pass, prefix, match, bin);
            hist_add(hist_s, match && in, bin);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kHistBinsMax; i += kSelThreads) {
        const unsigned int c = hist_s[i];
        if (c) atomicAdd(hist + i, (unsigned long long)c);
    }
}

// One block: sum the gathered histograms, inclusive scan, locate rank k.
__global__ void __launch_bounds__(1024)
select_scan_kernel(const unsigned long long* __restrict__ hist_parts, int parts, int pass, SelectState* state) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long warp_off[32];
    const int bins = (pass == 3) ? kHistBins3 : kHistBins1;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // read the state before the barriers: the one thread that locates k rewrites it below
    const unsigned long long k = state->k_remaining;
    const unsigned int old_prefix = state->prefix;
    // each thread owns two adjacent bins
    unsigned long long c0 = 0, c1 = 0;
    for (int p = 0; p < parts; ++p) {
        const unsigned long long* h = hist_parts + (size_t)p * kHistBinsMax;
        if (2 * t < bins)     c0 += h[2 * t];
        if (2 * t + 1 < bins) c1 += h[2 * t + 1];
    }
    unsigned long long incl = c0 + c1;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long o = __shfl_up_sync(kFull, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = warp_tot[lane], s = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long o = __shfl_up_sync(kFull, s, off);
            if (lane >= off) s += o;
        }
        warp_off[lane] = s - w;
    }
    __syncthreads();
    const unsigned long long end1 = warp_off[warp] + incl;   // count of keys in bins <= 2t+1
    const unsigned long long beg0 = end1 - c0 - c1;          // count of keys in bins < 2t
    int hit = -1;
    unsigned long long below = 0;
    if (k >= beg0 && k < beg0 + c0)          { hit = 2 * t;     below = beg0; }
    else if (k >= beg0 + c0 && k < end1)     { hit = 2 * t + 1; below = beg0 + c0; }
    if (hit >= 0) {
        const int bits = (pass == 3) ? 10 : 11;
        const unsigned int prefix = (pass == 1) ? (unsigned)hit : ((old_prefix << bits) | (unsigned)hit);
        state->prefix = prefix;
        state->k_remaining = k - below;
        state->pass = (unsigned)pass;
        if (pass == 3) {
            state->threshold_key = prefix;
            state->threshold = key_entropy(prefix);
        }
    }
}

__global__ void select_init_kernel(SelectState* state, unsigned long long k) {
    state->k_remaining = k;
    state->prefix = 0u;
    state->pass = 0u;
    state->threshold_key = 0u;
    state->threshold = 0.f;
    state->kept = 0ull;
}

// mask[i] = e[i] < thr (strict, IEEE: false for NaN on either side) + kept count
__global__ void __launch_bounds__(kSelThreads)
mask_kernel(const float* __restrict__ e, long long n, const float* __restrict__ thr_dev, float thr_host,
            uint8_t* __restrict__ mask, unsigned long long* __restrict__ kept, bool vec_ok) {
    const float thr = thr_dev ? *thr_dev : thr_host;
    const long long tid = (long long)blockIdx.x * kSelThreads + threadIdx.x;
    const long long stride = (long long)gridDim.x * kSelThreads;
    unsigned int local = 0;
    if (vec_ok) {
        const long long n4 = n >> 2;
        const float4* e4 = reinterpret_cast<const float4*>(e);
        uchar4* m4 = reinterpret_cast<uchar4*>(mask);
        for (long long i = tid; i < n4; i += stride) {
            const float4 v = __ldg(e4 + i);
            uchar4 m;
            m.x = v.x < thr; m.y = v.y < thr; m.z = v.z < thr; m.w = v.w < thr;
            if (mask) m4[i] = m;
            local += m.x + m.y + m.z + m.w;
        }
        for (long long i = (n4 << 2) + tid; i < n; i += stride) {
            const uint8_t m = e[i] < thr;
            if (mask) mask[i] = m;
            local += m;
        }
    } else {
        for (long long i = tid; i < n; i += stride) {
            const uint8_t m = __ldg(e + i) < thr;
            if (mask) mask[i] = m;
            local += m;
        }
    }
    if (kept) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) local += __shfl_xor_sync(kFull, local, off);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(kept, (unsigned long long)local);
    }
}

int stream_blocks(long long items_per_block_pass, long long n) {
    const int sms = device_sm_count();
    long long want = (n + items_per_block_pass - 1) / items_per_block_pass;
    const long long cap = (long long)sms * 4;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

}  // namespace

cudaError_t launch_select_init(SelectState* state, unsigned long long k, cudaStream_t stream) {
    select_init_kernel<<<1, 1, 0, stream>>>(state, k);
    return cudaGetLastError();
}

cudaError_t launch_select_hist(const float* e, long long n, int pass, const SelectState* state,
                               unsigned long long* hist, cudaStream_t stream) {
    cudaError_t err = cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * kHistBinsMax, stream);
    if (err != cudaSuccess) return err;
    if (n <= 0) return cudaSuccess;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(e) & 15u) == 0;
    const int blocks = stream_blocks((long long)kSelThreads * 4 * 4, n);
    select_hist_kernel<<<blocks, kSelThreads, 0, stream>>>(e, n, pass, state, hist, vec_ok);
    return cudaGetLastError();
}

cudaError_t launch_select_scan(const unsigned long long* hist_parts, int parts, int pass,
                               SelectState* state, cudaStream_t stream) {
    select_scan_kernel<<<1, 1024, 0, stream>>>(hist_parts, parts, pass, state);
    return cudaGetLastError();
}

cudaError_t launch_mask(const float* e, long long n, const float* thr_dev, float thr_host,
                        uint8_t* mask, unsigned long long* kept, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(e) & 15u) == 0 &&
                        (mask == nullptr || (reinterpret_cast<uintptr_t>(mask) & 3u) == 0);
    const int blocks = stream_blocks((long long)kSelThreads * 4 * 4, n);
    mask_kernel<<<blocks, kSelThreads, 0, stream>>>(e, n, thr_dev, thr_host, mask, kept, vec_ok);
    return cudaGetLastError();
}

}  // namespace suhpe
