// step_kernel.cuh -- closed-loop glue between two replans, on the device.
//
//   * narrowing of the QP solution to the float trajectory the reference stores
//     (result.desired_traj[m][i] = point3d(...), src/traj_optimizer.cpp:71-83; z := world_z_2d in 2-D)
//   * AgentManager::doStep (src/agent_manager.cpp:29-50): next state = desired_traj.getStateAt(step)
//     (Trajectory::getStateAt / getPointAt / derivative, src/trajectory.cpp:111-199), float arithmetic
//   * previous-solution shift used for both the own initial trajectory and the neighbours'
//     predictions (src/traj_planner.cpp:287-297, 402-411)
#pragma once
#include <math.h>

namespace lscqp {

// Peer exchange block of the sharded closed loop (BASELINE config 5; stands in for MultiSyncSimulator::broadcastMsgs,
// src/multi_sync_simulator.cpp:305-352, which copies every agent's state + prev_traj to every other agent in-process).
// Every rank owns one block in its own HBM, mapped into all peers over NVLink (CUDA IPC):
//     inbox[2][n_total][row]   row = M*18 shifted-trajectory floats + 9 state floats, double buffered by step parity
//     flags[world]             flags[r] = number of steps rank r has published into THIS block
//     seq, done, err, failed   local counters (step number, finished CTAs of the running publish, time-out flag, QP failures)
// step_kernel of step s writes the new rows of its agents straight into inbox[(s+1)&1] of every rank (remote stores),
// the last CTA then fences (system scope) and raises flags[rank] = s+1 everywhere.  exchange_begin_kernel of step s+1
// spins until all flags reach s+1 and copies inbox[(s+1)&1] into the rank's trajectory / state arrays.  A rank can only
// start publishing step s+1 after it has seen every peer's flag s+1, i.e. after every peer finished reading
// inbox[s&1] -- the buffer step s+1 overwrites -- so the two buffers suffice and no second barrier is needed.
struct ExchangeBlock {
    float* inbox;                        // [2][n_total][row]
    unsigned long long* flags;           // [world]
    unsigned long long* ctl;             // [4]: seq, done, err, failed
};
constexpr int EXCHANGE_MAX_WORLD = 16;
struct ExchangePeers {
    int world, rank, n_total, row;       // row = M * 18 + 9
    ExchangeBlock blk[EXCHANGE_MAX_WORLD];   // blk[rank] is the local block, the others are IPC mappings
};

struct StepParams {
    int n_agents, dim;
    double dt, step, z_2d;
    const double* ctrl;      // [n][dim][M][6]
    float* traj_out;         // [n][M][6][3] (may be null)
    float* state_out;        // [n][9]      (may be null)
    float* shifted_out;      // [n][M][6][3] (may be null)
    // failsafe of TrajPlanner::trajOptimization (src/traj_planner.cpp:767-797): an agent whose QP did not return OK
    // keeps initial_traj.  Both null: ctrl is taken as it is.
    const int*   status;     // [n]
    const float* fallback;   // [n][M][6][3] initial_traj
    // publish to the peers' inboxes (null: no exchange); lo = global index of local agent 0
    const ExchangePeers* peers;
    int lo;
};

__device__ __forceinline__ double ipow(double x, int e) {
    double r = 1.0;
    for (int i = 0; i < e; i++) r *= x;
    return r;
}
__device__ __forceinline__ double binom_small(int n, int k) {
    double r = 1.0;
    for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
    return r;
}

// Trajectory::getPointAt: segment index and normalised time of `time` (src/trajectory.cpp:111-141); m = -1: out of range
template <int M>
__device__ __forceinline__ void locate(double dt, double time, int& m, double& t_norm) {
    m = -1; t_norm = 0.0;
    double seg_end = 0.0;
    if (time < 0) return;
    for (int idx = 0; idx < M; idx++) {
        seg_end += dt;
        if (time < seg_end) { m = idx; t_norm = 1 - (seg_end - time) / dt; return; }
    }
    if (time < seg_end + 1e-5) { m = M - 1; t_norm = 1.0; }                     // :130-134
}

// One warp per agent: the lanes stride over the M*18 floats of a trajectory (coalesced loads and stores), lanes 0..8
// evaluate the state component (derivative order d = lane / 3, axis k = lane % 3) -- no per-thread arrays, no local memory.
constexpr int STEP_WARPS = 8;
template <int M>
__global__ void __launch_bounds__(STEP_WARPS * 32)
step_kernel(const StepParams p) {
#ifdef LSCQP_CUDA_EMUL
    float* s_traj = reinterpret_cast<float*>(emu_dyn_smem);
#else
    __shared__ float s_traj[STEP_WARPS * M * 18];
#endif
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int agent = blockIdx.x * STEP_WARPS + warp;
    const bool on = agent < p.n_agents;
    float* c = s_traj + warp * M * 18;                                          // [M][6][3] float control points
    if (on) {
        const bool keep_initial = p.status && p.status[agent] != 0;
        const double* x = p.ctrl + (size_t) agent * p.dim * M * 6;
        const float* fb = p.fallback ? p.fallback + (size_t) agent * M * 18 : nullptr;
        for (int e = lane; e < M * 18; e += 32) {
            const int k = e % 3, mi = e / 3;
            float v;
            if (keep_initial) v = fb[e];                                        // desired_traj = initial_traj, :795-797
            else v = k < p.dim ? (float) x[k * M * 6 + mi] : (float) p.z_2d;    // traj_optimizer.cpp:71-83
            c[e] = v;
            if (p.traj_out) p.traj_out[(size_t) agent * M * 18 + e] = v;
        }
        if (keep_initial && p.peers && lane == 0) atomicAdd(p.peers->blk[p.peers->rank].ctl + 3, 1ull);
    }
    __syncwarp();
    float st = 0.f;
    if (on && (p.state_out || p.peers) && lane < 9) {
        // Trajectory::getStateAt (trajectory.cpp:156-170) through derivative() (:183-199):
        // (c[i+1] - c[i]) * (float)(deg / segment_time), then the Bernstein sum of getPointAt in float
        const int d = lane / 3, k = lane % 3;
        int m; double t;
        locate<M>(p.dt, p.step, m, t);
        if (m >= 0) {
            float q[6];
#pragma unroll
            for (int i = 0; i < 6; i++) q[i] = c[(m * 6 + i) * 3 + k];
            const float s1 = (float) (5 / p.dt), s2 = (float) (4 / p.dt);
            if (d >= 1) {
#pragma unroll
                for (int i = 0; i < 5; i++) q[i] = __fmul_rn(__fsub_rn(q[i + 1], q[i]), s1);
            }
            if (d >= 2) {
#pragma unroll
                for (int i = 0; i < 4; i++) q[i] = __fmul_rn(__fsub_rn(q[i + 1], q[i]), s2);
            }
            const int deg = 5 - d;
#pragma unroll
            for (int i = 0; i < 6; i++) {
                if (i > deg) break;
                const float b = (float) (binom_small(deg, i) * ipow(t, i) * ipow(1 - t, deg - i));   // polynomial.hpp:22-24
                st = __fadd_rn(st, __fmul_rn(q[i], b));
            }
        }
        if (p.dim == 2 && lane == 2) st = (float) p.z_2d;                       // agent_manager.cpp:42-44
        if (p.state_out) p.state_out[(size_t) agent * 9 + lane] = st;
    }
    if (on && (p.shifted_out || p.peers)) {
        // previous-solution shift (traj_planner.cpp:287-297, 402-411): segment m <- m + 1, last segment = last point
        const int world = p.peers ? p.peers->world : 0;
        unsigned long long seq = 0;
        if (p.peers) seq = *reinterpret_cast<volatile unsigned long long*>(p.peers->blk[p.peers->rank].ctl);
        const size_t slot = p.peers ? ((size_t) ((seq + 1) & 1) * p.peers->n_total + (size_t) (p.lo + agent)) * p.peers->row : 0;
        for (int e = lane; e < M * 18; e += 32) {
            const int k = e % 3, m = e / 18;
            const float v = (m == M - 1) ? c[((M - 1) * 6 + 5) * 3 + k] : c[e + 18];
            if (p.shifted_out) p.shifted_out[(size_t) agent * M * 18 + e] = v;
            for (int r = 0; r < world; r++) p.peers->blk[r].inbox[slot + e] = v;
        }
        if (lane < 9) for (int r = 0; r < world; r++) p.peers->blk[r].inbox[slot + M * 18 + lane] = st;
    }
    if (p.peers) {
        // publish: when the last CTA is done, fence the remote stores and raise this rank's flag in every block
        const ExchangePeers& P = *p.peers;
        unsigned long long* ctl = P.blk[P.rank].ctl;
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long done = atomicAdd(ctl + 1, 1ull) + 1;
            if (done == gridDim.x) {
                __threadfence_system();
                const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(ctl) + 1;
                ctl[1] = 0;
                *reinterpret_cast<volatile unsigned long long*>(ctl) = seq;
                __threadfence_system();
                for (int r = 0; r < P.world; r++) *reinterpret_cast<volatile unsigned long long*>(P.blk[r].flags + P.rank) = seq;
            }
        }
    }
}

// Start of a closed-loop step: wait until every rank has published the step this rank is about to read, then copy
// that inbox buffer into the rank's replicated trajectory / state arrays.  Every CTA polls the flags on its own (local
// memory, <= 16 words).  A peer that never arrives would hang the device: the wait gives up after `timeout_cycles`
// and raises ctl[2] (lscqp_exchange_status), leaving the arrays untouched.
struct ExchangeBeginParams {
    const ExchangePeers* peers;
    float* traj;               // [n_total][M*18]
    float* state;              // [n_total][9]
    long long timeout_cycles;
};

__global__ void __launch_bounds__(256) exchange_begin_kernel(const ExchangeBeginParams p) {
    const ExchangePeers& P = *p.peers;
    const ExchangeBlock& me = P.blk[P.rank];
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(me.ctl);
    if (seq == 0) return;                                       // nothing published yet: the initial arrays are replicated
#ifdef LSCQP_CUDA_EMUL
    int* s_ok = reinterpret_cast<int*>(emu_dyn_smem);
#else
    __shared__ int s_ok[1];
#endif
    if (threadIdx.x == 0) {
        int ok = 1;
#ifndef LSCQP_CUDA_EMUL
        const long long t0 = clock64();
#endif
        for (int r = 0; r < P.world; r++) {
            while (*reinterpret_cast<volatile unsigned long long*>(me.flags + r) < seq) {
#ifdef LSCQP_CUDA_EMUL
                ok = 0; break;
#else
                if (clock64() - t0 > p.timeout_cycles) { ok = 0; break; }
                __nanosleep(100);
#endif
            }
            if (!ok) break;
        }
        if (!ok) *reinterpret_cast<volatile unsigned long long*>(me.ctl + 2) = 1ull;
        __threadfence_system();
        s_ok[0] = ok;
    }
    __syncthreads();
    if (!s_ok[0]) return;
    const float* src = me.inbox + (size_t) (seq & 1) * P.n_total * P.row;
    const int tw = P.row - 9;                                   // trajectory floats per row
    const size_t total = (size_t) P.n_total * P.row;
    for (size_t e = (size_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t) gridDim.x * blockDim.x) {
        const size_t a = e / P.row; const int j = (int) (e % P.row);
        const float v = src[e];
        if (j < tw) p.traj[a * tw + j] = v; else p.state[a * 9 + (j - tw)] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// TrajPlanner::isSolValid (src/traj_planner.cpp:990-1045), the check trajOptimization applies to a solution before
// accepting it in DLSC mode (:763-766): SFC containment of the float control points (segment 0: points phi..n only;
// Box::isPointInBox with its SP_EPSILON_FLOAT slack, src/collision_constraints.cpp:81-111) when world_use_octomap, and
// the velocity / acceleration of the state at the replanning period against the limits with 1 % tolerance.
// (The LSC check is commented out in the reference, :1012-1027.)  Thread per agent.
struct ValidateParams {
    int n_agents, M, dim, use_sfc;
    const float*  traj;      // [n][M][6][3]  result.desired_traj (float)
    const float*  state;     // [n][9]        desired_traj.getStateAt(multisim_time_step)
    const double* limits;    // [n][8]        max_vel[3], max_acc[3], ...
    const float*  sfc;       // [n][M][6]     (use_sfc)
    int* valid_out;          // [n]  1 valid | 0 not valid
};

__global__ void __launch_bounds__(128) validate_kernel(const ValidateParams p) {
    const int agent = blockIdx.x * blockDim.x + threadIdx.x;
    if (agent >= p.n_agents) return;
    bool ok = true;
    if (p.use_sfc) {
        for (int m = 0; m < p.M; m++) {
            const float* box = p.sfc + ((size_t) agent * p.M + m) * 6;
            for (int i = (m == 0 ? 3 : 0); i < 6; i++) {                         // :995-1007
                const float* c = p.traj + (((size_t) agent * p.M + m) * 6 + i) * 3;
                for (int k = 0; k < 3; k++)
                    ok = ok && ((double) c[k] > (double) box[k] - 1e-5) && ((double) c[k] < (double) box[3 + k] + 1e-5);
            }
        }
    }
    const double tol = 1.0 + 0.01;                                               // dyn_err_tol_ratio :1030
    for (int k = 0; k < p.dim; k++) {
        const double v = fabs((double) p.state[(size_t) agent * 9 + 3 + k]), a = fabs((double) p.state[(size_t) agent * 9 + 6 + k]);
        if (v > p.limits[(size_t) agent * 8 + k] * tol) ok = false;              // :1033-1036
        if (a > p.limits[(size_t) agent * 8 + 3 + k] * tol) ok = false;          // :1037-1040
    }
    p.valid_out[agent] = ok ? 1 : 0;
}

}  // namespace lscqp
