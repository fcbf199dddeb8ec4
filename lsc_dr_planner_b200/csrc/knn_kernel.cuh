// knn_kernel.cuh -- neighbour selection for the closed loop (what MultiSyncSimulator::broadcastMsgs does,
// src/multi_sync_simulator.cpp:305-352): the obstacles of agent qi are the other agents, restricted to those whose
// current position is within the Chebyshev communication range (:319-328).  The batched kernels take at most
// max_obs (<= 40) obstacles per agent, so the K nearest are kept (in-range agents first, then -- only when fewer than K
// are in range -- the nearest out-of-range ones as padding; their planes never bind).
//
// One CTA per agent.  Squared distances to all n_total agents go to shared memory as order-preserving unsigned keys
// (non-negative float bits; bit 31 set for out-of-range agents, 0xFFFFFFFF for the agent itself), a 4-pass radix select
// finds the K-th smallest key, and an ordered compaction writes the selected ids in ascending agent order (the order
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

__global__ void __launch_bounds__(KNN_THREADS) knn_select_kernel(const KnnParams p) {
#ifdef LSCQP_CUDA_EMUL
    unsigned* knn_smem = reinterpret_cast<unsigned*>(emu_dyn_smem);
#else
    extern __shared__ unsigned knn_smem[];
#endif
    unsigned* keys = knn_smem;                         // [n_total]
    unsigned* hist = knn_smem + p.n_total;             // [256]
    unsigned* cnt = hist + 256;                        // [2 * KNN_THREADS + 4]
    const int tid = threadIdx.x, a = p.lo + blockIdx.x, N = p.n_total;
    if ((int) blockIdx.x >= p.n_local) return;
    const float ax = p.state[(size_t) a * 9], ay = p.state[(size_t) a * 9 + 1], az = p.state[(size_t) a * 9 + 2];
    for (int j = tid; j < N; j += KNN_THREADS) {
        const float dx = p.state[(size_t) j * 9] - ax, dy = p.state[(size_t) j * 9 + 1] - ay, dz = p.state[(size_t) j * 9 + 2] - az;
        unsigned key = __float_as_uint(dx * dx + dy * dy + dz * dz) & 0x7FFFFFFFu;
        if (p.comm_range > 0.f && fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz)) > p.comm_range) key |= 0x80000000u;
        if (j == a) key = 0xFFFFFFFFu;
        keys[j] = key;
    }
    // radix select: after the pass over byte b, `prefix` holds the top bytes of the K-th smallest key and `want` the
    // rank still to be located inside that prefix
    unsigned prefix = 0, mask = 0;
    int want = p.K;                                    // 1-based rank
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int e = tid; e < 256; e += KNN_THREADS) hist[e] = 0;
        __syncthreads();
        for (int j = tid; j < N; j += KNN_THREADS) {
            const unsigned k = keys[j];
            if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0, b = 0;
            for (; b < 255; b++) {
                if (acc + (int) hist[b] >= want) break;
                acc += (int) hist[b];
            }
            cnt[0] = (unsigned) b; cnt[1] = (unsigned) (want - acc);
        }
        __syncthreads();
        prefix |= cnt[0] << shift; mask |= 255u << shift; want = (int) cnt[1];
        __syncthreads();
    }
    const unsigned thr = prefix;                       // the K-th smallest key; `want` of the keys equal to it are taken
    // ordered compaction: thread t owns the contiguous id range [t * chunk, (t + 1) * chunk)
    const int chunk = (N + KNN_THREADS - 1) / KNN_THREADS;
    const int j0 = tid * chunk, j1 = (j0 + chunk < N) ? j0 + chunk : N;
    unsigned nl = 0, ne = 0;
    for (int j = j0; j < j1; j++) { const unsigned k = keys[j]; nl += k < thr; ne += k == thr; }
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
        const unsigned k = keys[j];
        if (k < thr) { out[il + (ie < (unsigned) want ? ie : (unsigned) want)] = j; il++; }
        else if (k == thr) { if (ie < (unsigned) want) out[il + ie] = j; ie++; }
    }
    (void) tot_l;
}

}  // namespace lscqp
