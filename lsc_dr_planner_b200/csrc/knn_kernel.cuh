// knn_kernel.cuh -- neighbour selection for the closed loop (what MultiSyncSimulator::broadcastMsgs does,
// src/multi_sync_simulator.cpp:305-352): the obstacles of agent qi are the other agents whose current position is
// within the Chebyshev communication range (:319-328; every other agent when the range is <= 0), in ascending id order.
// The output is a ragged CSR list holding exactly that set.  The solve kernels hold at most max_obs obstacles per agent:
// an agent with more than K in range keeps its K nearest and is *flagged* (overflow_out = its in-range count), so the
// caller can raise max_obs or report it -- nothing is padded and nothing is dropped silently.
//
// knn_select_kernel: one CTA per agent.  Squared distances to all n_total agents go to shared memory as order-preserving
// unsigned keys (non-negative float bits; bit 31 set for out-of-range agents, 0xFFFFFFFF for the agent itself).  When more
// than K agents are in range the K-th smallest key is built bit by bit (32 counting passes over the keys: no atomics -- a
// radix histogram serialises on the handful of exponent values distances share); an ordered compaction then writes the
// selected ids in ascending agent order (the order broadcastMsgs emits them), ties at the threshold broken by the lower
// id: deterministic.  It writes a fixed-stride row per agent plus its count; knn_csr_kernel (one CTA: scan of the counts,
// then a copy) turns the rows into the CSR arrays the assembly / solve kernels read.
#pragma once

namespace lscqp {

struct KnnParams {
    int n_total, lo, n_local, K;
    double comm_range;             // <= 0: no range filter (a double like Param::communication_range)
    const float* state;            // [n_total][9] position first
    int* obs_index;                // [n_local][K]  selected ids, ascending; entries beyond count[a] are not written
    int* count;                    // [n_local]     number of selected ids (<= K)
    int* overflow;                 // [n_local]     0, or the in-range count when it exceeds K (may be null)
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
        // LInfinityDistance of the float positions, compared in double like multi_sync_simulator.cpp:321-326
        if (p.comm_range > 0.0 && (double) fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz)) > p.comm_range) key |= 0x80000000u;
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
    const int n_in = count_below(0x80000000u);         // agents in range (the agent's own key is 0xFFFFFFFF)
    if (tid == 0) {
        p.count[blockIdx.x] = n_in < p.K ? n_in : p.K;
        if (p.overflow) p.overflow[blockIdx.x] = n_in > p.K ? n_in : 0;
    }
    unsigned thr = 0x80000000u;                        // everything in range ...
    int want = 0;
    if (n_in > p.K) {
        // ... unless that exceeds the capacity: the K-th smallest key is the largest T with count(keys < T) < K
        unsigned prefix = 0;
        for (int bit = 31; bit >= 0; bit--) {
            const unsigned test = prefix | (1u << bit);
            if (count_below(test) < p.K) prefix = test;
        }
        want = p.K - count_below(prefix);              // how many of the keys equal to the threshold are taken
        thr = prefix;
    }
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

// rows of knn_select_kernel -> CSR: offsets = exclusive scan of the counts, ids copied in order.  One CTA.
struct KnnCsrParams {
    int n_local, K;
    const int* rows;               // [n_local][K]
    const int* count;              // [n_local]
    int* obs_offsets;              // [n_local + 1]
    int* obs_index;                // [sum count]
};

constexpr int KNN_CSR_THREADS = 1024;

__global__ void __launch_bounds__(KNN_CSR_THREADS) knn_csr_kernel(const KnnCsrParams p) {
#ifdef LSCQP_CUDA_EMUL
    int* s_warp = reinterpret_cast<int*>(emu_dyn_smem);
#else
    __shared__ int s_warp[KNN_CSR_THREADS / 32 + 1];
#endif
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int carry = 0;
    for (int base = 0; base < p.n_local; base += KNN_CSR_THREADS) {
        const int a = base + tid;
        const int c = a < p.n_local ? p.count[a] : 0;
        int incl = c;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = lane < KNN_CSR_THREADS / 32 ? s_warp[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            s_warp[lane] = w;                                   // inclusive totals of the warps
        }
        __syncthreads();
        const int off = carry + incl - c + (warp > 0 ? s_warp[warp - 1] : 0);
        if (a < p.n_local) {
            p.obs_offsets[a] = off;
            for (int j = 0; j < c; j++) p.obs_index[off + j] = p.rows[(size_t) a * p.K + j];
        }
        carry += s_warp[KNN_CSR_THREADS / 32 - 1];
        __syncthreads();
    }
    if (tid == 0) p.obs_offsets[p.n_local] = carry;
}

}  // namespace lscqp
