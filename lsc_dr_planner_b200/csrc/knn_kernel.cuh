// knn_kernel.cuh -- neighbour selection for the closed loop (what MultiSyncSimulator::broadcastMsgs does,
// src/multi_sync_simulator.cpp:305-352): the obstacles of agent qi are the other agents, restricted to those whose
// current position is within the Chebyshev communication range (:319-328).  The batched kernels take at most
// max_obs (<= 40) obstacles per agent, so the K nearest are kept (in-range agents first, then -- only when fewer than K
// are in range -- the nearest out-of-range ones as padding; their planes never bind).
//
// One CTA per agent.  Squared distances to all n_total agents go to shared memory as order-preserving unsigned keys
// (non-negative float bits; bit 31 set for out-of-range agents, 0xFFFFFFFF for the agent itself).  The K-th smallest key
// is built bit by bit (32 counting passes over the keys: no atomics -- a radix histogram serialises on the handful of
// exponent values distances share), and an ordered compaction writes the selected ids in ascending agent order (the order
// broadcastMsgs emits them), ties at the threshold broken by the lower id: deterministic.
#pragma once

namespace lscqp {

struct KnnParams {
    int n_total, lo, n_local, K;
    float comm_range;              // <= 0: no range filter
    const float* state;            // [n_total][9] position first
    int* obs_index;                // [n_local][K]
};

constexpr int KNN_THREADS = 128;
// keys (skewed by one word per 32) + per-warp / per-thread counters
inline size_t knn_smem_bytes(int n_total) { return ((size_t) n_total + (n_total >> 5) + 1 + 2 * KNN_THREADS + 8) * sizeof(unsigned); }

__global__ void __launch_bounds__(KNN_THREADS) knn_select_kernel(const KnnParams p) {
#ifdef LSCQP_CUDA_EMUL
    unsigned* knn_smem = reinterpret_cast<unsigned*>(emu_dyn_smem);
#else
    extern __shared__ unsigned knn_smem[];
#endif
    unsigned* keys = knn_smem;                         // [n_total]
    unsigned* cnt = knn_smem + p.n_total + (p.n_total >> 5) + 1;   // [2 * KNN_THREADS + 4]
    auto skew = [](int j) { return j + (j >> 5); };    // one pad word per 32 keys: the compaction's strided reads hit distinct banks
    const int tid = threadIdx.x, a = p.lo + blockIdx.x, N = p.n_total;
    if ((int) blockIdx.x >= p.n_local) return;
    const float ax = p.state[(size_t) a * 9], ay = p.state[(size_t) a * 9 + 1], az = p.state[(size_t) a * 9 + 2];
    for (int j = tid; j < N; j += KNN_THREADS) {
        const float dx = p.state[(size_t) j * 9] - ax, dy = p.state[(size_t) j * 9 + 1] - ay, dz = p.state[(size_t) j * 9 + 2] - az;
        unsigned key = __float_as_uint(dx * dx + dy * dy + dz * dz) & 0x7FFFFFFFu;
        if (p.comm_range > 0.f && fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz)) > p.comm_range) key |= 0x80000000u;
        if (j == a) key = 0xFFFFFFFFu;
        keys[skew(j)] = key;
    }
    __syncthreads();
    // count(keys < t) over the CTA: per-thread count, warp shuffle sum, one shared-memory hop
    auto count_below = [&](unsigned t) {
        int c = 0;
        for (int j = tid; j < N; j += KNN_THREADS) c += keys[skew(j)] < t;
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((tid & 31) == 0) cnt[tid >> 5] = (unsigned) c;
        __syncthreads();
        int tot = 0;
        for (int w = 0; w < KNN_THREADS / 32; w++) tot += (int) cnt[w];
        __syncthreads();
        return tot;
    };
    // the K-th smallest key is the largest T with count(keys < T) < K: set its bits from the top
    unsigned prefix = 0;
    for (int bit = 31; bit >= 0; bit--) {
        const unsigned test = prefix | (1u << bit);
        if (count_below(test) < p.K) prefix = test;
    }
    const int want = p.K - count_below(prefix);        // how many of the keys equal to the threshold are taken
    const unsigned thr = prefix;                       // the K-th smallest key; `want` of the keys equal to it are taken
    // ordered compaction: thread t owns the contiguous id range [t * chunk, (t + 1) * chunk)
    const int chunk = (N + KNN_THREADS - 1) / KNN_THREADS;
    const int j0 = tid * chunk, j1 = (j0 + chunk < N) ? j0 + chunk : N;
    unsigned nl = 0, ne = 0;
    for (int j = j0; j < j1; j++) { const unsigned k = keys[skew(j)]; nl += k < thr; ne += k == thr; }
    cnt[4 + tid] = nl; cnt[4 + KNN_THREADS + tid] = ne;
    __syncthreads();
    unsigned base_l = 0, base_e = 0, tot_l = 0;
    for (int t = 0; t < KNN_THREADS; t++) {
        if (t < tid) { base_l += cnt[4 + t]; base_e += cnt[4 + KNN_THREADS + t]; }
        tot_l += cnt[4 + t];
    }
    // a key below the threshold takes slot (#below before it) + (#taken equal before it); equal keys are taken while
    // their running count is within `want`
    int* out = p.obs_index + (size_t) blockIdx.x * p.K;
    unsigned il = base_l, ie = base_e;
    for (int j = j0; j < j1; j++) {
        const unsigned k = keys[skew(j)];
        if (k < thr) { out[il + (ie < (unsigned) want ? ie : (unsigned) want)] = j; il++; }
        else if (k == thr) { if (ie < (unsigned) want) out[il + ie] = j; ie++; }
    }
    (void) tot_l;
}

}  // namespace lscqp
