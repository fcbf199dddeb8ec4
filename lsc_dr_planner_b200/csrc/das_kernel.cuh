// das_kernel.cuh -- batched dual active-set solve of the per-agent trajectory QP: the first pass of lscqp_solve_batch.
//
// Same model as pdip_kernel.cuh (the QP TrajOptimizer::populatebyrow builds, src/traj_optimizer.cpp:216-514, solved by the
// CPLEX call :66), same analytic elimination of the equalities, same exact presolve, same agent-local coordinates -- but
// solved by the Goldfarb-Idnani dual active-set method (Math. Programming 27, 1983) instead of an interior point:
//   * The reduced Hessian H = Z'(blkdiag 2 w_c Q_base + terminal weights)Z does not depend on the agent: it is block
//     diagonal over the dimensions and has one value per number of terminal segments (traj_optimizer.cpp:530-538), so its
//     inverse and the inverse Cholesky factor J = L^-T come from a host-built table (host_common.hpp:build_das_tables).
//   * Start at the unconstrained minimiser y = -H^-1 g.  Each iteration evaluates every row q_r(c) >= 0 at the current
//     point (no per-row state: the rows are recomputed from the control points), takes the most violated one (normal n)
//     and moves along z = (H^-1 - J1 J1') n in the primal / r = S J1' n in the dual space, where the columns of J1 are
//     H-orthonormal and span H^-1 x the active normals and S = R^-1 is the inverse of their triangular factor
//     (J1' N = R).  Either the row becomes feasible -- it joins the active set: J1 gains the column z / sqrt(n'z), S
//     the column (-r, 1) / sqrt(n'z) -- or a multiplier reaches zero first: that row leaves (Givens rotations on the
//     columns of S and J1).  Only J1 is stored (the complement J2 of the textbook method never is: J2 J2' = H^-1 - J1 J1'),
//     the dual direction is a lane-parallel product with the stored inverse factor, and lane j owns active row j.
//     On the trajectory QPs of the forest workload 13-30 rows end up active (of 39 variables) and almost every
//     iteration is an add: ~24 iterations of a row sweep and a few short dot products, instead of ~9 interior-point
//     iterations of four row sweeps, a Hessian assembly, a factorisation and two solves.
//   * The result is checked before it is accepted: primal feasibility of every row at the returned point (the loop's
//     exit test; the active rows, which the method keeps at zero only up to rounding, are re-evaluated too), multipliers
//     >= 0 (invariant of the method), stationarity recomputed from scratch.  Anything else -- more kept obstacles or
//     active rows than this kernel holds, an iteration cap, an infeasible row, a residual above tolerance -- flags the
//     agent in klass[] and leaves it to the interior-point pass (pdip_kernel.cuh, klass_mode 2).
// One warp per agent, 32 threads per CTA, no CTA barrier, no atomics: results are bit-reproducible.
#pragma once
#include "pdip_kernel.cuh"

#ifndef LSCQP_DAS_MINCTAS
#define LSCQP_DAS_MINCTAS 12
#endif
#ifndef LSCQP_DAS_QMAX
#define LSCQP_DAS_QMAX 32
#endif
#ifndef LSCQP_DAS_KPT
#define LSCQP_DAS_KPT 10
#endif

namespace lscqp {

// KPT_: kept obstacles the instance holds.  Up to 16 their row constants live in registers (the throughput instance);
// the large instance (host_common.hpp:Instance::DAS_BIG_KPT) keeps them in shared memory and is run second, over the
// agents the first flagged with "more kept obstacles than I hold" only.
template <class C, int KPT_ = LSCQP_DAS_KPT>
struct Das {
    static constexpr int M = C::M, D = C::D, NCP = C::NCP, NV = C::NV, NR = C::NR, NZS = C::NZS;
    static constexpr int N1 = NR / D;                       // reduced variables per dimension
    static constexpr int RPL = (NR + 31) / 32;              // reduced variables per lane
    static constexpr int VPT = (NV + 31) / 32;              // full-space variables (box-row owners) per lane
    static constexpr int CPL = (NCP + 31) / 32;             // control points (LSC-row owners) per lane
    static constexpr int KPT = KPT_;                        // kept obstacles this kernel holds
    static constexpr bool BIG = KPT_ > 16;                  // row constants in shared memory
    static constexpr int NLSC = CPL * KPT;                  // row slots of a lane: LSC rows first, then 6 per variable
    static constexpr int NSLOT = NLSC + VPT * 6;
    // active rows the instance holds: one lane each in the throughput instance, every reduced variable (two per lane) in the large one
    static constexpr int QMAX = BIG ? (NR < 64 ? NR : 64) : LSCQP_DAS_QMAX;
    static constexpr int QSL = (QMAX + 31) / 32;            // active rows per lane
    static constexpr int LDJ = QMAX | 1;
    static constexpr int KRAW = C::KRAW;
    static_assert(NSLOT <= 64, "row slots do not fit the 64-bit masks");
    static_assert(NR * LDJ >= NV, "full-space scratch does not fit under J1");
    // shared memory (doubles)
    static constexpr int O_C = 0;
    static constexpr int O_Y = O_C + NV;
    static constexpr int O_X0 = O_Y + NR;
    static constexpr int O_VLIM = O_X0 + 3 * D;
    static constexpr int O_ALIM = O_VLIM + D;
    static constexpr int O_GOAL = O_ALIM + D;
    static constexpr int O_ORG = O_GOAL + D;
    static constexpr int O_LB = O_ORG + D;
    static constexpr int O_UB = O_LB + D * M;
    static constexpr int O_TERMW = O_UB + D * M;            // [M] terminal weights, [M] the number of terminal segments
    static constexpr int O_NRM = O_TERMW + M + 1;           // [KPT][M][3]
    static constexpr int O_RB = O_NRM + KPT * M * 3;        // [KPT][CPL][32] row constants (BIG only)
    static constexpr int O_D = O_RB + (BIG ? KPT * CPL * 32 : 0);   // d1 = J1'n [QMAX]
    static constexpr int O_J = O_D + QMAX;                  // J1 [NR][LDJ]  (start-up and epilogue: full-space scratch [NV])
    static constexpr int O_R = O_J + NR * LDJ;              // S = R^-1, upper triangular, packed by columns: S(j, k) at k (k + 1) / 2 + j
    static constexpr int O_INT = O_R + QMAX * (QMAX + 1) / 2;   // int ids[QMAX] (active rows: slot * 32 + lane), act[KRAW], keep[KRAW + 2]
    static constexpr int O_END = O_INT + (QMAX + 2 * KRAW + 2 + 1) / 2;
    // checkpoint of the throughput instance (SolveParams::das_ckpt), in doubles: q, it, y[NR], u[QF], ids[QF], J1[NR][QF], S packed
    static constexpr int QF = LSCQP_DAS_QMAX, KF = LSCQP_DAS_KPT;
    static constexpr int CK_Y = 2, CK_U = CK_Y + NR, CK_ID = CK_U + QF, CK_J = CK_ID + QF, CK_S = CK_J + NR * QF, CK_STRIDE = CK_S + QF * (QF + 1) / 2;
    static constexpr int SMEM_BYTES = O_END * 8;
};

// reduced index <-> (dimension, index within the dimension)
template <class C>
__device__ __forceinline__ void das_decode(int r, int& k, int& r1) {
    if (C::TERM && r >= (C::M - 1) * C::NZS) { k = r - (C::M - 1) * C::NZS; r1 = 3 * (C::M - 1); }
    else { k = (r % C::NZS) / 3; r1 = 3 * (r / C::NZS) + r % 3; }
}
template <class C>
__device__ __forceinline__ int das_encode(int k, int r1) {
    if (C::TERM && r1 == 3 * (C::M - 1)) return (C::M - 1) * C::NZS + k;
    return (r1 / 3) * C::NZS + k * 3 + r1 % 3;
}

// Row (owner lane lp, slot) in the full space: at most three (dimension, control point, coefficient) entries, unused ones   // @phase row_decode
// with a zero coefficient.  LSC row of control point cp and kept obstacle j: normal n on (k, cp), k < D
// (traj_optimizer.cpp:414-421); box rows of variable v: 0 lb, 1 ub, 2 vel+, 3 vel-, 4 acc+, 5 acc- with unit-coefficient
// stencils (:238-270, :440-474).
template <class C, int KPT_>
__device__ __forceinline__ void das_row_full(int lp, int slot, const double* s_nrm, int* fk, int* fcp, double* fa) {
    using A = Das<C, KPT_>;
    constexpr int M = C::M, D = C::D, NCP = C::NCP;
    if (slot < A::NLSC) {
        const int cp = lp + 32 * (slot / A::KPT), j = slot % A::KPT;
#pragma unroll
        for (int k = 0; k < 3; k++) { fk[k] = k < D ? k : 0; fcp[k] = cp; fa[k] = k < D ? s_nrm[(j * M + cp / 6) * 3 + k] : 0.0; }
    } else {
        const int b = slot - A::NLSC, u = b / 6, e = b % 6;
        const int v = lp + 32 * u, k = v / NCP, cp = v % NCP;
        const double sg = (e & 1) ? -1.0 : 1.0;
#pragma unroll
        for (int i = 0; i < 3; i++) { fk[i] = k; fcp[i] = cp; fa[i] = 0.0; }
        if (e < 2) { fa[0] = sg; }                                                       // c - lb >= 0, ub - c >= 0
        else if (e < 4) { fa[0] = sg; fa[1] = -sg; fcp[1] = cp + 1; }                     // vlim -+ (c[i+1] - c[i]) >= 0
        else { fa[0] = -sg; fa[1] = 2.0 * sg; fa[2] = -sg; fcp[1] = cp + 1; fcp[2] = cp + 2; }   // alim -+ (c[i+2] - 2 c[i+1] + c[i]) >= 0
    }
}

// ... and in the reduced space: at most 9 (reduced variable pr, coefficient pv) pairs through the continuity map, with
// the variable's dimension pk and index pc within the dimension (das_decode)
template <class C, int KPT_>
__device__ __forceinline__ void das_row_pairs(int lp, int slot, const double* s_nrm, int* pr, double* pv, int* pk, int* pc) {
    int fk[3], fcp[3];
    double fa[3];
    das_row_full<C, KPT_>(lp, slot, s_nrm, fk, fcp, fa);
#pragma unroll
    for (int f = 0; f < 3; f++) {
        const int m = fcp[f] / 6, i = fcp[f] % 6;
#pragma unroll
        for (int j = 0; j < 3; j++) { pr[f * 3 + j] = 0; pv[f * 3 + j] = 0.0; pk[f * 3 + j] = fk[f]; pc[f * 3 + j] = 0; }
        if (i >= 3) {
            pr[f * 3] = ridx<C>(m, fk[f], i - 3); pv[f * 3] = fa[f];
            pc[f * 3] = (C::TERM && m == C::M - 1) ? 3 * (C::M - 1) : 3 * m + i - 3;
        } else if (m >= 1) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                pr[f * 3 + j] = ridx<C>(m - 1, fk[f], j); pv[f * 3 + j] = fa[f] * tcoef(i, j); pc[f * 3 + j] = 3 * (m - 1) + j;
            }
        }
    }
}

// exact warp arg-min of a double per lane (ties: the lowest lane): two redux.sync.min passes over the halves of an   // @phase argmin
// order-preserving 64-bit key.  Returns the winning lane.
__device__ __forceinline__ int warp_argmin(double v) {
    long long b = __double_as_longlong(v);
    const unsigned long long key = (unsigned long long) b ^ ((unsigned long long) (b >> 63) | 0x8000000000000000ull);
    const unsigned hi = (unsigned) (key >> 32), lo = (unsigned) key;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    return __ffs(__ballot_sync(0xffffffffu, hi == mhi && lo == mlo)) - 1;
}

template <class C, int KPT_ = LSCQP_DAS_KPT>   // @phase setup
__global__ void __launch_bounds__(32, (KPT_ > 16 ? 6 : LSCQP_DAS_MINCTAS))
das_solve_kernel(const SolveParams p) {
    using A = Das<C, KPT_>;
    constexpr int M = C::M, D = C::D, NCP = C::NCP, NV = C::NV, NR = C::NR;
    constexpr int RPL = A::RPL, VPT = A::VPT, CPL = A::CPL, KPT = A::KPT, LDJ = A::LDJ, N1 = A::N1, QMAX = A::QMAX, QSL = A::QSL;
    constexpr unsigned FULL = 0xffffffffu;
    LSCQP_DYN_SMEM(sm);
    const int lane = threadIdx.x & 31;
    const int agent = blockIdx.x;
    if (agent >= p.n_agents) return;

    double* s_c = sm + A::O_C;
    double* s_y = sm + A::O_Y;
    double* s_x0 = sm + A::O_X0;
    double* s_vlim = sm + A::O_VLIM;
    double* s_alim = sm + A::O_ALIM;
    double* s_goal = sm + A::O_GOAL;
    double* s_org = sm + A::O_ORG;
    double* s_lb = sm + A::O_LB;
    double* s_ub = sm + A::O_UB;
    double* s_termw = sm + A::O_TERMW;
    double* s_nrm = sm + A::O_NRM;
    double* s_d = sm + A::O_D;
    double* s_J = sm + A::O_J;
    double* s_S = sm + A::O_R;
    double* s_full = s_J;                                   // [NV] full-space scratch (start-up and epilogue only)
    int* s_ids = reinterpret_cast<int*>(sm + A::O_INT);
    int* s_act = s_ids + A::QMAX;
    int* s_keep = s_act + A::KRAW;
    const double* sQ2 = p.Q2;                               // (kernel parameter: constant bank)

    // every exit that does not deliver a checked solution leaves the agent to the interior-point pass
    // (klass: 0 solved here; else the reason -- 1 obstacle list beyond the ABI capacity, 2 more kept obstacles than KPT,
    //  3 iteration cap, 4 infeasible row, 5 NaN, 6 feasibility / stationarity / multiplier check failed, 7 more active rows than QMAX)
    auto defer = [&](int why) { if (lane == 0) p.klass[agent] = why; };
    const int kl_in = p.klass_mode == 3 ? p.klass[agent] : 0;            // (large instance: the first pass's verdict, + checkpoint slot)

    if (p.klass_mode == 3 && (kl_in & 15) != 2 && (kl_in & 15) != 7) return;   // large instance: only what the first pass could not hold
    const int obs0 = p.obs_offsets[agent];
    int K = p.obs_offsets[agent + 1] - obs0;
    if (K < 0 || K > A::KRAW || K > p.max_obs) { defer(1); return; }      // (reported as ST_CAPACITY by the other pass)
    if (p.rsfc) K = 0;                                                     // see SolveParams::rsfc

    // ---- per-agent constants (as pdip_kernel.cuh: local coordinates, unit-coefficient velocity / acceleration rows)
    if (lane < D) {
        const int k = lane;
        const double pos = (double) p.state[agent * 9 + k], vel = (double) p.state[agent * 9 + 3 + k],
                     acc = (double) p.state[agent * 9 + 6 + k];
        const double c0 = 0.0, c1 = vel * p.dt / 5.0, c2 = acc * p.dt * p.dt / 20.0 + 2.0 * c1 - c0;   // traj_optimizer.cpp:321-338
        s_org[k] = pos;
        s_x0[k * 3 + 0] = c0; s_x0[k * 3 + 1] = c1; s_x0[k * 3 + 2] = c2;
        s_vlim[k] = p.limits[agent * 8 + k] * p.dt / 5.0;                 // :448-453
        s_alim[k] = p.limits[agent * 8 + 3 + k] * p.dt * p.dt / 20.0;     // :462-471
        s_goal[k] = (double) p.goal[agent * 3 + k] - pos;
        for (int m = 0; m < M; m++) {
            double lo = p.world_min[k], hi = p.world_max[k];              // :252-253
            if (p.rsfc && k == 2 && m == 0) { lo = -100.0; hi = 100.0; }  // :255-258
            if (p.use_sfc) {                                               // :372-397
                lo = fmax(lo, (double) p.sfc[((size_t) agent * M + m) * 6 + k]);
                hi = fmin(hi, (double) p.sfc[((size_t) agent * M + m) * 6 + 3 + k]);
            }
            s_lb[k * M + m] = lo - pos; s_ub[k * M + m] = hi - pos;
        }
    }
    if (lane == 8) {
        // getTerminalSegments_old, traj_optimizer.cpp:530-538 (float norm of the point3d difference)
        const float dx = p.goal[agent * 3 + 0] - p.state[agent * 9 + 0], dy = p.goal[agent * 3 + 1] - p.state[agent * 9 + 1],
                    dz = p.goal[agent * 3 + 2] - p.state[agent * 9 + 2];
        const float nsq = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        const double flight = sqrt((double) nsq) / p.limits[agent * 8 + 7];
        int ts = (int) ((M * p.dt - flight + 1e-9) / p.dt);
        if (ts < 1) ts = 1;
        if (ts > M) ts = M;
        for (int m = 0; m < M; m++) s_termw[m] = (m >= M - ts) ? 2.0 * p.w_t : 0.0;
        s_termw[M] = (double) ts;
    }
    for (int e = lane; e < A::KRAW; e += 32) s_keep[e] = (e < K && !p.presolve) ? 1 : 0;
    __syncwarp();
    const int ts = (int) s_termw[M];

    // ---- rows of this lane.  Slot s < NLSC: LSC row of control point lane + 32 (s / KPT), kept obstacle s % KPT;   // @phase rows_presolve
    // slot NLSC + 6 u + e: box row e of variable lane + 32 u.  rmask: the rows that exist.
    unsigned long long rmask = 0, amask = 0;                 // existing rows / rows in the active set
    unsigned bmask[VPT];
#pragma unroll
    for (int u = 0; u < VPT; u++) {
        const int v = lane + u * 32, m_v = (v % NCP) / 6, i_v = v % 6;
        const bool on = v < NV;
        const bool has_bnd = on && !(m_v == 0 && i_v < 3);              // :260-265
        const bool has_vel = on && i_v < 5 && !(m_v == 0 && i_v < 2);   // :444
        const bool has_acc = on && i_v < 4 && !(m_v == 0 && i_v < 1);   // :458
        bmask[u] = (has_bnd ? 3u : 0u) | (has_vel ? 12u : 0u) | (has_acc ? 48u : 0u);
        if (p.presolve && has_bnd) {
            // exact presolve of the bound rows (pdip_kernel.cuh): the reach box of the velocity rows lies strictly inside
            const int k_v = v / NCP;
            const double reach = (double) (5 * m_v + i_v - 2) * s_vlim[k_v];
            if (s_x0[k_v * 3 + 2] - reach - s_lb[k_v * M + m_v] > 1e-6 && s_ub[k_v * M + m_v] - (s_x0[k_v * 3 + 2] + reach) > 1e-6)
                bmask[u] &= ~3u;
        }
        rmask |= (unsigned long long) bmask[u] << (A::NLSC + 6 * u);
    }
    // exact presolve of the obstacles (pdip_kernel.cuh): dropped when no point the velocity rows allow can activate a row
    if (p.presolve) {
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            const int cp = lane + 32 * c, m_cp = cp / 6, i_cp = cp % 6;
            if (cp >= NCP || (m_cp == 0 && i_cp < 3)) continue;
            const double steps = (double) (5 * m_cp + i_cp - 2);
#pragma unroll 4
            for (int oi = 0; oi < K; oi++) {
                const double* g = p.normals + ((size_t) (obs0 + oi) * M + m_cp) * 3;
                const double nx = g[0], ny = g[1], nz = (D == 3) ? g[2] : 0.0;
                double b = p.rhs[((size_t) (obs0 + oi) * M + m_cp) * 6 + i_cp] - (nx * s_org[0] + ny * s_org[1]);
                if (D == 3) b -= nz * s_org[2];
                double lo = nx * s_x0[2] + ny * s_x0[5] - steps * (fabs(nx) * s_vlim[0] + fabs(ny) * s_vlim[1]);
                if (D == 3) lo += nz * s_x0[8] - steps * fabs(nz) * s_vlim[2];
                const bool zero_normal = (float) nx == 0.0f && (float) ny == 0.0f && (D == 2 || (float) nz == 0.0f);
                if (!(lo - b > 1e-6) && !zero_normal) s_keep[oi] = 1;      // benign race: every writer stores 1
            }
        }
    }
    __syncwarp();
    for (int e = lane; e < K; e += 32) {
        if (!s_keep[e]) continue;
        int pos = 0;
        for (int u = 0; u < e; u++) pos += s_keep[u];
        if (pos < KPT) s_act[pos] = e;
    }
    if (lane == 0) {
        int n = 0;
        for (int u = 0; u < K; u++) n += s_keep[u];
        s_keep[A::KRAW] = n;
    }
    __syncwarp();
    K = s_keep[A::KRAW];
    if (K > KPT) { defer(2); return; }
    for (int e = lane; e < K * M; e += 32) {
        // rows with a (float) normal shorter than SP_EPSILON_FLOAT are skipped by the reference (traj_optimizer.cpp:409-411)
        const double* g = p.normals + ((size_t) (obs0 + s_act[e / M]) * M + e % M) * 3;
        double nx = g[0], ny = g[1], nz = g[2];
        const float fx = (float) nx, fy = (float) ny, fz = (float) nz;
        const float nsq = __fadd_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)), __fmul_rn(fz, fz));
        if (sqrt((double) nsq) < 1e-5) { nx = 0.0; ny = 0.0; nz = 0.0; }
        s_nrm[e * 3] = nx; s_nrm[e * 3 + 1] = ny; s_nrm[e * 3 + 2] = nz;
    }
    __syncwarp();
    // row constants in local coordinates (n . c - rb >= 0): registers, or shared memory [j][c][lane] in the large instance
    double rb_reg[A::BIG ? 1 : CPL][A::BIG ? 1 : KPT];
    double* s_rb = sm + A::O_RB;
#define LSCQP_RB(c, j) (A::BIG ? s_rb[((j) * CPL + (c)) * 32 + lane] : rb_reg[A::BIG ? 0 : (c)][A::BIG ? 0 : (j)])
#pragma unroll
    for (int c = 0; c < CPL; c++) {
        const int cp = lane + 32 * c, m_cp = cp / 6, i_cp = cp % 6;
        const bool lsc = cp < NCP && !(m_cp == 0 && i_cp < 3);          // :404
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            if (A::BIG && j >= K) break;
            if (A::BIG) s_rb[(j * CPL + c) * 32 + lane] = 0.0; else rb_reg[A::BIG ? 0 : c][A::BIG ? 0 : j] = 0.0;
            if (!lsc || j >= K) continue;
            const double* n = s_nrm + (j * M + m_cp) * 3;
            if (n[0] == 0.0 && n[1] == 0.0 && n[2] == 0.0) continue;
            double b = p.rhs[((size_t) (obs0 + s_act[j]) * M + m_cp) * 6 + i_cp] - (n[0] * s_org[0] + n[1] * s_org[1]);
            if (D == 3) b -= n[2] * s_org[2];
            if (A::BIG) s_rb[(j * CPL + c) * 32 + lane] = b; else rb_reg[A::BIG ? 0 : c][A::BIG ? 0 : j] = b;
            rmask |= 1ull << (c * KPT + j);
        }
    }

    auto expand = [&]() {
#pragma unroll
        for (int u = 0; u < VPT; u++) {
            const int v = lane + u * 32;
            if (v < NV) s_c[v] = full_from_reduced<C>(s_y, s_x0, v / NCP, (v % NCP) / 6, v % 6);
        }
    };
    // full-space gradient of the objective at s_c:  (2 w_c Q) c + terminal terms (traj_optimizer.cpp:294-315)
    auto grad_full = [&](double* out) {
#pragma unroll
        for (int u = 0; u < VPT; u++) {
            const int v = lane + u * 32;
            if (v >= NV) continue;
            const int k_v = v / NCP, m_v = (v % NCP) / 6, a = v % 6, v0 = k_v * NCP + m_v * 6;
            double gr = 0.0;
#pragma unroll
            for (int b = 0; b < 6; b++) gr += sQ2[a * 6 + b] * s_c[v0 + b];
            if (a == 5) gr += s_termw[m_v] * (s_c[v0 + 5] - s_goal[k_v]);
            out[v] = gr;
        }
    };

    // ---- unconstrained minimiser y = -H^-1 g, H^-1 from the table of this agent's terminal-segment count   // @phase start_point
    const double* Hinv = p.das_tab + (size_t) (ts - 1) * 2 * N1 * N1;
    const int ck_slot = A::BIG ? (kl_in >> 4) : 0;                        // > 0: resume from the throughput instance's checkpoint
#pragma unroll
    for (int t = 0; t < RPL; t++) { const int r = lane + 32 * t; if (r < NR) s_y[r] = 0.0; }
    __syncwarp();
    expand();
    __syncwarp();
    grad_full(s_full);
    __syncwarp();
    double g0[RPL];
#pragma unroll
    for (int t = 0; t < RPL; t++) { const int r = lane + 32 * t; g0[t] = r < NR ? reduce_from_full<C>(s_full, r) : 0.0; }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < RPL; t++) { const int r = lane + 32 * t; if (r < NR) s_full[r] = g0[t]; }
    __syncwarp();
    int my_k[RPL], my_r1[RPL];                               // dimension / index within the dimension of this lane's variables
#pragma unroll
    for (int t = 0; t < RPL; t++) {
        const int r = lane + 32 * t;
        my_k[t] = -1; my_r1[t] = 0;
        if (r >= NR) continue;
        das_decode<C>(r, my_k[t], my_r1[t]);
        double y = 0.0;
        for (int c1 = 0; c1 < N1; c1++) y -= Hinv[c1 * N1 + my_r1[t]] * s_full[das_encode<C>(my_k[t], c1)];
        s_y[r] = y;
    }
    __syncwarp();

    // ---- row evaluation at s_c.  only < 0: every existing row outside the active set, result = the most violated one   // @phase sweep_general
    // (smallest slack in the reference's row scaling: velocity rows carry 5/dt, acceleration rows 20/dt^2) of this lane;
    // only >= 0: that slot alone; only == -2: nothing; only == -3: the rows of the active set.
    const double wv = 5.0 / p.dt, wa = 20.0 / (p.dt * p.dt);
    auto sweep = [&](int only, double& best, double& best_raw, int& best_slot) {
        best = INFINITY; best_raw = 0.0; best_slot = 0;
        const unsigned long long live = only == -1 ? (rmask & ~amask) : (only == -3 ? amask : (only >= 0 ? 1ull << only : 0ull));
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            const int cp = lane + 32 * c;
            if (cp >= NCP || !((live >> (c * KPT)) & ((1ull << KPT) - 1))) continue;
            const int m_cp = cp / 6;
            const double cx = s_c[cp], cy = s_c[NCP + cp], cz = (D == 3) ? s_c[2 * NCP + cp] : 0.0;
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                if (A::BIG && j >= K) break;
                if (!(live >> (c * KPT + j) & 1ull)) continue;
                const double* n = s_nrm + (j * M + m_cp) * 3;
                double q = n[0] * cx + n[1] * cy - LSCQP_RB(c, j);
                if (D == 3) q += n[2] * cz;
                if (q < best) { best = q; best_raw = q; best_slot = c * KPT + j; }
            }
        }
#pragma unroll
        for (int u = 0; u < VPT; u++) {
            const int v = lane + u * 32;
            const unsigned bits = (unsigned) (live >> (A::NLSC + 6 * u)) & 63u;
            if (v >= NV || !bits) continue;
            const int k_v = v / NCP, m_v = (v % NCP) / 6;
            const double* cc = s_c + v;
            const double c0 = cc[0];
            double dv = 0.0, da = 0.0;
            if (bmask[u] & 4u) dv = cc[1] - c0;
            if (bmask[u] & 16u) da = cc[2] - 2.0 * cc[1] + c0;
            double q[6];
            q[0] = c0 - s_lb[k_v * M + m_v]; q[1] = s_ub[k_v * M + m_v] - c0;
            q[2] = s_vlim[k_v] - dv; q[3] = s_vlim[k_v] + dv; q[4] = s_alim[k_v] - da; q[5] = s_alim[k_v] + da;
#pragma unroll
            for (int e = 0; e < 6; e++) {
                if (!(bits >> e & 1u)) continue;
                const double w = e < 2 ? 1.0 : (e < 4 ? wv : wa);
                if (q[e] * w < best) { best = q[e] * w; best_raw = q[e]; best_slot = A::NLSC + 6 * u + e; }
            }
        }
    };

    // ---- the same over every live row, cheaper: the two rows of a bound / velocity / acceleration pair share one evaluation   // @phase sweep_fast
    // (lim - |x| is the smaller of the two slacks; with one of them in the active set the other cannot be violated), no
    // per-row branches.  Returns the lane's smallest scaled slack and its slot; the raw slack is best * row_unscale(slot).
    int kv_[VPT];
#pragma unroll
    for (int u = 0; u < VPT; u++) kv_[u] = (lane + 32 * u) / NCP;
    const bool any_bnd = __any_sync(FULL, [&]() { unsigned b = 0;
#pragma unroll
        for (int u = 0; u < VPT; u++) b |= bmask[u] & 3u;
        return b != 0; }());
    auto sweep_fast = [&](double& best, int& best_slot) {
        best = INFINITY; best_slot = 0;
        const unsigned long long live = rmask & ~amask;
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            const int cp = lane + 32 * c;
            if (cp >= NCP) continue;
            const int m_cp = cp / 6;
            const double cx = s_c[cp], cy = s_c[NCP + cp], cz = (D == 3) ? s_c[2 * NCP + cp] : 0.0;
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                if (j >= K) break;                                        // (warp uniform)
                const double* n = s_nrm + (j * M + m_cp) * 3;
                double q = n[0] * cx + n[1] * cy - LSCQP_RB(c, j);
                if (D == 3) q += n[2] * cz;
                q = (live >> (c * KPT + j) & 1ull) ? q : INFINITY;
                if (q < best) { best = q; best_slot = c * KPT + j; }
            }
        }
        const unsigned long long both = live & (live >> 1);               // bit 2i: rows 2i and 2i + 1 both live (NLSC is even)
#pragma unroll
        for (int u = 0; u < VPT; u++) {
            const int v = lane + u * 32;
            if (v >= NV) continue;
            const unsigned bits = (unsigned) (both >> (A::NLSC + 6 * u));
            const double* cc = s_c + v;                                    // (cc[1], cc[2] may lie past the array: masked rows)
            const double c0 = cc[0], c1 = cc[1], c2 = cc[2];
            const double dv = c1 - c0, da = c2 - 2.0 * c1 + c0;
            double qv = (s_vlim[kv_[u]] - fabs(dv)) * wv, qa = (s_alim[kv_[u]] - fabs(da)) * wa;
            qv = (bits & 4u) ? qv : INFINITY; qa = (bits & 16u) ? qa : INFINITY;
            if (qv < best) { best = qv; best_slot = A::NLSC + 6 * u + (dv > 0.0 ? 2 : 3); }
            if (qa < best) { best = qa; best_slot = A::NLSC + 6 * u + (da > 0.0 ? 4 : 5); }
            if (any_bnd) {                                                 // (warp uniform)
                const int m_v = (v % NCP) / 6;
                const double ql = c0 - s_lb[kv_[u] * M + m_v], qu = s_ub[kv_[u] * M + m_v] - c0;
                const double qb = (bits & 1u) ? fmin(ql, qu) : INFINITY;
                if (qb < best) { best = qb; best_slot = A::NLSC + 6 * u + (ql <= qu ? 0 : 1); }
            }
        }
    };
    auto row_unscale = [&](int slot) {
        const int e = slot < A::NLSC ? 0 : (slot - A::NLSC) % 6;
        return e < 2 ? 1.0 : (e < 4 ? p.dt / 5.0 : p.dt * p.dt / 20.0);
    };

    // ---- drop the active row at position l.  With S = R^-1: rotate the columns (j, j+1), j = l .. q-2, of S so that row l   // @phase drop
    // of S becomes zero left of the last column; the new inverse factor is S without row l and without its last column,
    // and J1 follows with the same column rotations (its last column leaves the span).
    int q = 0;                                                // size of the active set; active row j is owned by lane j % 32, slot j / 32
    double u_own[QSL];                                        // multipliers of this lane's active rows
#pragma unroll
    for (int t = 0; t < QSL; t++) u_own[t] = 0.0;
    auto drop = [&](int l) {
        const int id = s_ids[l];
        if (lane == (id & 31)) amask &= ~(1ull << (id >> 5));
        for (int j = l; j < q - 1; j++) {
            double* cj = s_S + j * (j + 1) / 2;               // column j: rows 0 .. j (unshifted: row l still present)
            double* cn = s_S + (j + 1) * (j + 2) / 2;         // column j + 1: rows 0 .. j + 1
            const double a = cj[l], b = cn[l];
            const double h = sqrt(a * a + b * b);
            double cs = 1.0, sn = 0.0;
            if (h > 0.0) { cs = b / h; sn = a / h; }
            double xs[QSL], ys[QSL];
#pragma unroll
            for (int t = 0; t < QSL; t++) {
                const int i = lane + 32 * t;
                xs[t] = i <= j ? cj[i] : 0.0; ys[t] = i <= j + 1 ? cn[i] : 0.0;
            }
            __syncwarp();
#pragma unroll
            for (int t = 0; t < QSL; t++) {
                const int i = lane + 32 * t;
                if (i > j + 1) continue;
                cn[i] = sn * xs[t] + cs * ys[t];                                    // stays in place for the next rotation
                if (i != l) cj[i < l ? i : i - 1] = cs * xs[t] - sn * ys[t];        // final: row l (now zero) removed
            }
#pragma unroll
            for (int t = 0; t < RPL; t++) {
                const int r = lane + 32 * t;
                if (r >= NR) continue;
                const double x = s_J[r * LDJ + j], yv = s_J[r * LDJ + j + 1];
                s_J[r * LDJ + j] = cs * x - sn * yv; s_J[r * LDJ + j + 1] = sn * x + cs * yv;
            }
            __syncwarp();
        }
        // the multipliers and ids behind position l move up by one (through s_d: dead until the next d1)
#pragma unroll
        for (int t = 0; t < QSL; t++) { const int j = lane + 32 * t; if (j < q) s_d[j] = u_own[t]; }
        __syncwarp();
        int idn[QSL];
#pragma unroll
        for (int t = 0; t < QSL; t++) {
            const int j = lane + 32 * t;
            idn[t] = (j >= l && j < q - 1) ? s_ids[j + 1] : 0;
            if (j >= l && j < q - 1) u_own[t] = s_d[j + 1];
            if (j == q - 1) u_own[t] = 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < QSL; t++) { const int j = lane + 32 * t; if (j >= l && j < q - 1) s_ids[j] = idn[t]; }
        q--;
        __syncwarp();
    };

    // ---- main loop   // @phase select_row
    int it = 0, why = 5, q_top = 0;
    bool ck_ok = false;                                       // the state at the point of failure can be handed over
    if (A::BIG && ck_slot > 0) {
        // resume: the state the throughput instance left when it needed one more active row than it holds
        const double* ck = p.das_ckpt + (size_t) (ck_slot - 1) * A::CK_STRIDE;
        q = (int) ck[0]; it = (int) ck[1]; q_top = q;
#pragma unroll
        for (int t = 0; t < RPL; t++) { const int r = lane + 32 * t; if (r < NR) s_y[r] = ck[A::CK_Y + r]; }
#pragma unroll
        for (int t = 0; t < QSL; t++) {
            const int j = lane + 32 * t;
            u_own[t] = j < q ? ck[A::CK_U + j] : 0.0;
            if (j < q) {
                // row ids are (slot, lane) pairs and the slot numbering depends on the instance's obstacle capacity
                const int idf = (int) ck[A::CK_ID + j], lo = idf & 31, sf = idf >> 5;
                const int sb = sf < CPL * A::KF ? (sf / A::KF) * KPT + sf % A::KF : sf - CPL * A::KF + A::NLSC;
                s_ids[j] = sb * 32 + lo;
            }
        }
        for (int e = lane; e < NR * q; e += 32) { const int r = e / q, k = e % q; s_J[r * LDJ + k] = ck[A::CK_J + r * A::QF + k]; }
        for (int e = lane; e < q * (q + 1) / 2; e += 32) s_S[e] = ck[A::CK_S + e];
        __syncwarp();
        for (int j = 0; j < q; j++) { const int id = s_ids[j]; if (lane == (id & 31)) amask |= 1ull << (id >> 5); }
    }
    const int it_max = 4 * NR + 40;
    bool ok = false;
    double viol = 0.0;
    while (true) {
        expand();
        __syncwarp();
        double best; int best_slot;
        sweep_fast(best, best_slot);
        const int lp = warp_argmin(best);
        const int slot = __shfl_sync(FULL, best_slot, lp);
        best = __shfl_sync(FULL, best, lp);
        if (!(best == best)) break;                           // NaN: leave it to the other pass
        if (!(best < -1e-10)) { ok = true; viol = fmax(0.0, -best); break; }
        const int pid = slot * 32 + lp;
        double sp = best * row_unscale(slot);                 // slack of the row being added (negative)
        int pr[9], pk[9], pc[9]; double pv[9];
        das_row_pairs<C, KPT_>(lp, slot, s_nrm, pr, pv, pk, pc);
        // this lane's entries of n and of w = H^-1 n (H^-1 is block diagonal over the dimensions)
        double nr[RPL], w[RPL];
#pragma unroll
        for (int t = 0; t < RPL; t++) { nr[t] = 0.0; w[t] = 0.0; }
#pragma unroll
        for (int i = 0; i < 9; i++) {
            if (pv[i] == 0.0) continue;
#pragma unroll
            for (int t = 0; t < RPL; t++) {
                if (pr[i] == lane + 32 * t) nr[t] += pv[i];
                if (pk[i] == my_k[t]) w[t] += Hinv[pc[i] * N1 + my_r1[t]] * pv[i];
            }
        }
        double nwp = 0.0;
#pragma unroll
        for (int t = 0; t < RPL; t++) nwp += nr[t] * w[t];
        const double nw = warp_sum(nwp);                      // n' H^-1 n
        double u_new = 0.0;                                   // multiplier of row p
        bool fail = false;
        while (true) {                                         // until row p is added (or the model is found infeasible)
            if (++it > it_max) { fail = true; why = 3; break; }
            // d1 = J1'n (lane = active row), z = w - J1 d1, r = S d1   // @phase directions
#pragma unroll
            for (int t = 0; t < QSL; t++) {
                const int j = lane + 32 * t;
                double dk = 0.0;
                if (j < q) {
#pragma unroll
                    for (int i = 0; i < 9; i++) if (pv[i] != 0.0) dk += pv[i] * s_J[pr[i] * LDJ + j];
                }
                if (j < QMAX) s_d[j] = dk;
            }
            __syncwarp();
            // one pass over the active columns k: z = w - J1 d1 (this lane's RPL rows) and r_j = sum_{k >= j} S(j, k) d1_k
            double zz[RPL], rj[QSL];
            {
                double za[RPL], zb[RPL], ra[QSL], rb2[QSL];
#pragma unroll
                for (int t = 0; t < RPL; t++) { za[t] = 0.0; zb[t] = 0.0; }
#pragma unroll
                for (int t = 0; t < QSL; t++) { ra[t] = 0.0; rb2[t] = 0.0; }
                const double* col = s_S;                        // column k of S starts at k (k + 1) / 2
                int k = 0;
                for (; k + 1 < q; k += 2) {
                    const double d0 = s_d[k], d1 = s_d[k + 1];
#pragma unroll
                    for (int t = 0; t < RPL; t++) {
                        const int r = lane + 32 * t;
                        if (r < NR) { za[t] += s_J[r * LDJ + k] * d0; zb[t] += s_J[r * LDJ + k + 1] * d1; }
                    }
#pragma unroll
                    for (int t = 0; t < QSL; t++) { const int j = lane + 32 * t; if (j <= k) ra[t] += col[j] * d0; }
                    col += k + 1;
#pragma unroll
                    for (int t = 0; t < QSL; t++) { const int j = lane + 32 * t; if (j <= k + 1) rb2[t] += col[j] * d1; }
                    col += k + 2;
                }
                if (k < q) {
                    const double d0 = s_d[k];
#pragma unroll
                    for (int t = 0; t < RPL; t++) { const int r = lane + 32 * t; if (r < NR) za[t] += s_J[r * LDJ + k] * d0; }
#pragma unroll
                    for (int t = 0; t < QSL; t++) { const int j = lane + 32 * t; if (j <= k) ra[t] += col[j] * d0; }
                }
#pragma unroll
                for (int t = 0; t < RPL; t++) zz[t] = w[t] - (za[t] + zb[t]);
#pragma unroll
                for (int t = 0; t < QSL; t++) rj[t] = ra[t] + rb2[t];
            }
            double nzp = 0.0;
#pragma unroll
            for (int t = 0; t < RPL; t++) nzp += nr[t] * zz[t];
            const double nz = warp_sum(nzp);                   // n'z = n'(H^-1 - J1 J1')n >= 0: the curvature along z
            // step lengths: t1 keeps the multipliers non-negative, t2 makes row p feasible   // @phase step_lengths
            double t1 = INFINITY;
            int l = 0;
#pragma unroll
            for (int t = 0; t < QSL; t++) {
                const int j = lane + 32 * t;
                if (j < q && rj[t] > 0.0) { const double ra = u_own[t] / rj[t]; if (ra < t1) { t1 = ra; l = j; } }
            }
            {
                const int wl = warp_argmin(t1);
                t1 = __shfl_sync(FULL, t1, wl);
                l = __shfl_sync(FULL, l, wl);
            }
            const bool dependent = !(nz > 1e-24 * nw);         // n in the span of the active normals: no primal step
            const double t2 = dependent ? INFINITY : -sp / nz;
            const double tt = fmin(t1, t2);
            if (tt == INFINITY) { fail = true; why = 4; break; }         // infeasible (or numerically so): the other pass decides
            if (t2 <= t1 && q >= QMAX) {
                // a full step would add a row this instance cannot hold.  Nothing of row p has been applied yet when its
                // multiplier is still zero: (y, active set, u) is then a consistent pair the large instance can resume from.
                fail = true; why = 7; ck_ok = u_new == 0.0; it--;
                break;
            }
#pragma unroll
            for (int t = 0; t < RPL; t++) {
                const int j = lane + 32 * t;
                if (!dependent && j < NR) s_y[j] += tt * zz[t];
            }
#pragma unroll
            for (int t = 0; t < QSL; t++) { const int j = lane + 32 * t; if (j < q) u_own[t] -= tt * rj[t]; }
            u_new += tt;
            __syncwarp();
            if (t2 <= t1) {
                // full step: row p joins the active set   // @phase add_row
                const double rinv = 1.0 / sqrt(nz);
#pragma unroll
                for (int t = 0; t < RPL; t++) { const int r = lane + 32 * t; if (r < NR) s_J[r * LDJ + q] = zz[t] * rinv; }
                // R gains the column (d1; rho), rho = sqrt(n'z):  S gains (-S d1; 1) / rho = (-r; 1) / rho
                double* cq = s_S + q * (q + 1) / 2;
#pragma unroll
                for (int t = 0; t < QSL; t++) {
                    const int j = lane + 32 * t;
                    if (j < q) cq[j] = -rj[t] * rinv;
                    if (j == q) { cq[q] = rinv; u_own[t] = u_new; s_ids[q] = pid; }
                }
                if (lane == lp) amask |= 1ull << slot;
                q++;
                q_top = q > q_top ? q : q_top;
                __syncwarp();
                break;
            }
            // partial step: the multiplier of active row l reached zero before row p became feasible   // @phase partial_step
            drop(l);
            if (!dependent) {
                expand();
                __syncwarp();
                double b2, raw2; int s2;
                sweep(lane == lp ? slot : -2, b2, raw2, s2);     // (-2: the other lanes evaluate nothing)
                sp = __shfl_sync(FULL, raw2, lp);
                if (!(sp < 0.0)) sp = -1e-300;                   // (rounding: the step length was t1 < t2)
            }
        }
        if (fail) break;
    }
    if (!ok) {
#ifndef LSCQP_DAS_NO_CKPT
        if (!A::BIG && why == 7 && ck_ok && p.das_ckpt) {
            // hand-over to the large instance (SolveParams::das_ckpt): the state goes into a pool slot, the slot number into
            // klass[] (bits 4 and up) beside the reason
            int slot_ck = 0;
            if (lane == 0) slot_ck = atomicAdd(p.das_ckpt_count, 1);
            slot_ck = __shfl_sync(FULL, slot_ck, 0);
            if (slot_ck < p.das_ckpt_slots) {
                double* ck = p.das_ckpt + (size_t) slot_ck * A::CK_STRIDE;
                if (lane == 0) { ck[0] = (double) q; ck[1] = (double) it; }
                for (int r = lane; r < NR; r += 32) ck[A::CK_Y + r] = s_y[r];
#pragma unroll
                for (int t = 0; t < QSL; t++) { const int j = lane + 32 * t; if (j < q) { ck[A::CK_U + j] = u_own[t]; ck[A::CK_ID + j] = (double) s_ids[j]; } }
                for (int e = lane; e < NR * q; e += 32) { const int r = e / q, k = e % q; ck[A::CK_J + r * A::QF + k] = s_J[r * LDJ + k]; }
                for (int e = lane; e < q * (q + 1) / 2; e += 32) ck[A::CK_S + e] = s_S[e];
                why = 7 | ((slot_ck + 1) << 4);
            }
        }
#endif
        defer(why); return;
    }

    // ---- verification and outputs.  s_c holds the final point.  The active rows are zero only up to the rounding of the   // @phase verify_output
    // updates: re-evaluated like the others.  Stationarity Z'(grad f - sum u_j n_j) from scratch (J1 is dead: its storage is
    // the full-space scratch).
    {
        double b3, raw3; int s3;
        sweep(-3, b3, raw3, s3);
        viol = fmax(viol, -warp_min(b3));
    }
    __syncwarp();
    grad_full(s_full);
    __syncwarp();
    double gscale = 0.0;
#pragma unroll
    for (int t = 0; t < RPL; t++) { const int r = lane + 32 * t; if (r < NR) gscale = fmax(gscale, fabs(reduce_from_full<C>(s_full, r))); }
    for (int j = 0; j < q; j++) {                             // the active rows in order: <= 3 full-space entries each
        const int id = s_ids[j];
        int fk[3], fcp[3]; double fa[3];
        das_row_full<C, KPT_>(id & 31, id >> 5, s_nrm, fk, fcp, fa);
        double um = 0.0;
#pragma unroll
        for (int t = 0; t < QSL; t++) if ((j >> 5) == t) um = u_own[t];
        const double uj = __shfl_sync(FULL, um, j & 31);
        // lane k applies the entries of dimension k, in program order: no two lanes ever touch the same address
#pragma unroll
        for (int t = 0; t < 3; t++) if (lane == fk[t] && fa[t] != 0.0) s_full[fk[t] * NCP + fcp[t]] -= uj * fa[t];
    }
    __syncwarp();
    double rd = 0.0;
#pragma unroll
    for (int t = 0; t < RPL; t++) { const int r = lane + 32 * t; if (r < NR) rd = fmax(rd, fabs(reduce_from_full<C>(s_full, r))); }
    rd = warp_max(rd); gscale = warp_max(gscale);
    double um_ = 0.0;
#pragma unroll
    for (int t = 0; t < QSL; t++) if (lane + 32 * t < q) um_ = fmin(um_, u_own[t]);
    const double umin = warp_min(um_);
#ifdef LSCQP_CUDA_EMUL
    if (lane == 0 && getenv("LSCQP_DAS_DEBUG")) fprintf(stderr, "das agent %d: it %d q %d rd %.3e gscale %.3e umin %.3e viol %.3e\n", agent, it, q, rd, gscale, umin, viol);
#endif
    if (!(rd <= 1e-7 * fmax(1.0, gscale)) || !(umin >= -1e-9 * fmax(1.0, gscale)) || !(viol <= 1e-9)) { defer(6); return; }

    double cost = 0.0;
#pragma unroll
    for (int u = 0; u < VPT; u++) {
        const int v = lane + u * 32;
        if (v >= NV) continue;
        // objective x'Px + q'x + c0 with P = w_c Q (no 1/2, traj_optimizer.cpp:294) + terminal terms (:301-315)
        const int k_v = v / NCP, m_v = (v % NCP) / 6, i_v = v % 6, v0 = k_v * NCP + m_v * 6;
        double qc = 0.0;
#pragma unroll
        for (int b = 0; b < 6; b++) qc += sQ2[i_v * 6 + b] * s_c[v0 + b];
        cost += 0.5 * s_c[v] * qc;
        if (i_v == 5) { const double e = s_c[v] - s_goal[k_v]; cost += 0.5 * s_termw[m_v] * e * e; }
        p.ctrl_out[(size_t) agent * NV + v] = s_c[v] + s_org[k_v];
    }
    cost = warp_sum(cost);
    if (lane == 0) {
        p.cost_out[agent] = cost;
        p.status_out[agent] = ST_OK;
        p.klass[agent] = 0;
        if (p.iters_out) p.iters_out[agent] = it;
        if (p.kkt_out) {
            p.kkt_out[agent * 4 + 0] = rd; p.kkt_out[agent * 4 + 1] = viol;
            p.kkt_out[agent * 4 + 2] = (double) (q_top * 64 + q); p.kkt_out[agent * 4 + 3] = 0.0;   // (largest, final active set)
        }
    }
    if (p.dual_out) {
        // layout of pdip_kernel.cuh: [KRAW][M][6] LSC rows, then [NV][6] box rows in the reference's row scaling
        double* du = p.dual_out + (size_t) agent * p.dual_stride;
        for (int e = lane; e < p.dual_stride; e += 32) du[e] = 0.0;
        __syncwarp();
        const double sv = p.dt / 5.0, sa = p.dt * p.dt / 20.0;
#pragma unroll
        for (int t = 0; t < QSL; t++) {
            const int j = lane + 32 * t;
            if (j >= q) continue;
            const int id = s_ids[j], lo = id & 31, slot = id >> 5;
            if (slot < A::NLSC) {
                const int cp = lo + 32 * (slot / KPT);
                du[(s_act[slot % KPT] * M + cp / 6) * 6 + cp % 6] = u_own[t];
            } else {
                const int b = slot - A::NLSC, e = b % 6, v = lo + 32 * (b / 6);
                du[A::KRAW * M * 6 + v * 6 + e] = u_own[t] * (e < 2 ? 1.0 : (e < 4 ? sv : sa));
            }
        }
    }
}

#undef LSCQP_RB

}  // namespace lscqp
