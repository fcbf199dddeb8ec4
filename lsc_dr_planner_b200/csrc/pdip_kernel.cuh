// pdip_kernel.cuh -- batched primal-dual interior-point solve of the per-agent trajectory QP.
//
// One CTA per agent -- a single warp in the light instances, 128 or 256 threads in the full-capacity ones (Cfg below,
// host_common.hpp:Instance).  Replaces the CPLEX call of the reference
// (src/traj_optimizer.cpp:18-156, model built by populatebyrow :216-514) for the whole batch the
// serial loop src/multi_sync_simulator.cpp:354-362 walks.
//
// Formulation:
//   * The equalities (initial state, C0..C2 continuity, LSC-mode terminal stop; :319-368, :504-511)
//     are eliminated analytically.  Per dimension the free variables are the last three control
//     points of each segment (one collapsed point for the last segment in LSC mode); the first three
//     control points of segment m >= 1 follow from C0..C2 continuity,
//         c[m][0] = y5,  c[m][1] = 2 y5 - y4,  c[m][2] = y3 - 4 y4 + 4 y5   (y = c[m-1][3..5]),
//     and c[0][0..2] from the initial position / velocity / acceleration.  Reduced size
//     NR = D(3M-2) (LSC mode) or 3DM; ordered by segment so the reduced KKT matrix is banded.
//   * Inequalities are kept as rows q_r(c) >= 0 on the full control points: LSC half-spaces
//     n.c - b >= 0 (:400-437), box bounds (world box, intersected with the SFC box when given;
//     :238-270, :372-397), velocity / acceleration differences (:440-474, rescaled to unit
//     coefficients).  Slack s and multiplier lam of every row live in registers of the owning thread.
//     An exact presolve drops the obstacles and bound rows no point allowed by the velocity rows can activate.
//   * The QP is solved in agent-local coordinates: every control point is taken relative to the agent's current position
//     (the jerk cost and all difference rows are shift invariant; goal, bounds and row constants are shifted once at
//     start-up, the position is added back on output).  Iterates then stay within the ~1 m the horizon can reach, so
//     their floating-point grid is 10-60x finer than at world coordinates of up to +-64 m.
//   * Mehrotra predictor-corrector on (y, s, lam).  Per iteration the row weights are accumulated
//     into the structured full-space Hessian (6x6 blocks per (dim, segment) + 3x3 blocks per control
//     point), projected to the reduced space (host-built per-thread term streams), factorised by a stage-aware
//     banded block-LDL^T on a skewed band in shared memory (warp 0) and used for the two triangular solve pairs.
#pragma once
#include <math.h>

#ifdef LSCQP_CUDA_EMUL
#define LSCQP_DYN_SMEM(name) double* name = emu_dyn_smem
#else
#define LSCQP_DYN_SMEM(name) extern __shared__ double name[]
#endif

#ifndef LSCQP_FULL_MINCTAS
#define LSCQP_FULL_MINCTAS 4
#endif
#ifndef LSCQP_LIGHT_MINCTAS
#define LSCQP_LIGHT_MINCTAS 8
#endif

namespace lscqp {   // @phase helpers

// one term of the projection stream (host_common.hpp:build_projection): val += coef * src[src]; a term with
// dest >= 0 closes an entry of the reduced matrix: A[dest & PROJ_MASK] = val (PROJ_ONE: 1.0, the identity padding;
// PROJ_DIAG: also kept as the unregularised diagonal for the pivot guard)
struct alignas(16) ProjTerm { double coef; int src; int dest; };
enum { PROJ_MASK = 0xFFFFF, PROJ_ONE = 1 << 29, PROJ_DIAG = 1 << 30 };

struct SolveParams {
    int n_agents;
    int max_iter;
    double mu_tol, rp_tol;
    double dt, w_t;
    double world_min[3], world_max[3];
    int use_sfc;
    int presolve;               // drop obstacles whose rows are all proven inactive by velocity-bound propagation (exact)
    int rsfc;                   // PlannerMode::RECIPROCALRSFC: its SlackMode::COLLISIONCONSTRAINT gives every LSC row a free slack
                                // epsilon <= 0 without a cost term (traj_optimizer.cpp:272-283, :423-425; slack_collision_weight is
                                // never read), so no LSC row can bind: the x-part of the minimiser is that of the model without
                                // them; and the z bounds of segment 0 are relaxed to +-100 (:255-258)
    int max_obs;                // lscqp_config.max_obs: a longer obstacle list is reported as ST_CAPACITY, never truncated
    const float*  state;        // [n][9]   position, velocity, acceleration
    const float*  goal;         // [n][3]   current_goal_point
    const double* limits;       // [n][8]   vmax[3], amax[3], radius, nominal_velocity
    const float*  sfc;          // [n][M][6] box_min, box_max  (use_sfc)
    const float*  next_waypoint;   // [n][3]  (communication-range rows)
    double comm_range;
    const int*    obs_offsets;  // [n+1]
    const double* normals;      // [sumK][M][3]
    const double* rhs;          // [sumK][M][6]   b = n.p + d
    const float*  warm_traj;    // [n][M][6][3] initial_traj used as the starting point (may be null: cold start)
    double mu0, warm_delta, warm_reject;   // warm start: complementarity target, minimum slack, rejection threshold
    double* ctrl_out;           // [n][D][M][6]
    double* cost_out;           // [n]
    int*    status_out;         // [n]
    int*    iters_out;          // [n]     (may be null)
    double* kkt_out;            // [n][4]  (may be null) stationarity, primal, dual, gap
    double* dual_out;           // [n][dual_stride] (may be null)
    int     dual_stride;
    double  Q2[36];             // 2 * w_c * Q_base  (Hessian block of the jerk cost)
    // projection table (host-built, see host_common.hpp:build_projection): reduced-matrix entry e is
    //   A[dest_e] = sum_t coef_t * src[idx_t],   src = [6x6 blocks | per-control-point cross blocks]
    const ProjTerm* proj;       // [proj_len][NT] term streams of the full-capacity instance
    const ProjTerm* proj_light; // [proj_len_light][NT_light] of the light instance (klass_mode 1)
    int proj_len, proj_len_light;
    double  w_c;
    // two-pass dispatch (lscqp.cu): the light instance (few kept obstacles per thread, many resident QPs per SM)
    // runs first with klass_mode 1 and flags the agents whose kept obstacles exceed its capacity; the full-capacity
    // instance then runs with klass_mode 2 and solves only the flagged agents.  0: solve every agent.
    int*    klass;              // [n]
    int     klass_mode;
    // dual active-set first pass (das_kernel.cuh): per terminal-segment count ts = 1..M the inverse reduced Hessian of one
    // dimension and its inverse Cholesky factor, [M][2][N1][N1] (host_common.hpp:build_das_table)
    const double* das_tab;
    // hand-over from the throughput active-set instance to the large one: an agent that needs a 33rd active row writes its
    // state (iterate, multipliers, active rows, J1, S) into a slot of this pool and passes the slot number in klass[]
    // (bits 4 and up), so the large instance resumes instead of starting over.  Null / exhausted pool: it starts over.
    double* das_ckpt;
    int*    das_ckpt_count;
    int     das_ckpt_slots;
};

enum { ST_OK = 0, ST_MAX_ITER = 1, ST_INFEASIBLE = 2, ST_NUMERICAL = 3, ST_CAPACITY = 4 };

template <int M_, int D_, bool TERM_, int G_, int KPT_, bool COMM_ = false>
struct Cfg {
    static constexpr int M = M_, D = D_, G = G_, KPT = KPT_;
    static constexpr bool TERM = TERM_;
    static constexpr bool COMM = COMM_;                  // communication-range rows (traj_optimizer.cpp:477-500): dense reduced matrix
    static constexpr int NCP = 6 * M;                    // control points per dimension
    static constexpr int CPW = ((NCP + 31) / 32) * 32;   // threads per obstacle group
    static constexpr int NT = CPW * G;                   // threads per CTA
    static constexpr int NW = NT / 32;
    static constexpr int NV = D * NCP;                   // full-space variables
    static constexpr int NZS = 3 * D;                    // reduced variables per stage
    static constexpr int NR = TERM ? (M - 1) * NZS + D : M * NZS;
    static constexpr int NRP0 = ((NR + 2) / 3) * 3;
    static constexpr int BW = COMM ? NRP0 - 1 : 2 * NZS - 1;   // half bandwidth of the reduced matrix (dense with comm rows)
    static constexpr int NRP = ((NR + 2) / 3) * 3;       // padded to whole 3x3 panels (identity rows)
    static constexpr int NB = NRP / 3;                   // panels
    static constexpr int BWS = COMM ? NRP0 - 1 : BW + 2;  // stored / assembled half bandwidth (panel overhang)
    // leading dimension of the reduced matrix.  Banded instances store only the band: with LD >= BWS the address
    // i LD + j of an entry with 0 <= i - j <= BWS is unique (row i occupies i (LD + 1) - BWS .. i (LD + 1)), so
    // the dense indexing works unchanged on a skewed band of (NRP - 1)(LD + 1) + 1 doubles.
    static constexpr int LD = COMM ? (NRP | 1) : (BWS | 1);
    static constexpr int A_SIZE = COMM ? NRP * LD : (NRP - 1) * (LD + 1) + 1;
    static constexpr int NS = D * (D + 1) / 2;           // unique entries of a DxD symmetric block
    static constexpr int KMAX = G * KPT;
    static constexpr int NPAIR = COMM ? 0 : BW * (BW + 1) / 2;   // trailing-update pairs per panel (banded instances)
    static constexpr int PP = M * (M - 1) / 2 + M;       // comm pairs per dimension: end-point differences + end-point boxes
    static constexpr int PC = COMM ? D * PP : 0;         // comm (+/-) row pairs, one per thread t < PC
    static constexpr int NBX = COMM ? 8 : 6;             // box-type rows per variable thread
    static constexpr int NPR = NPAIR ? (NPAIR + 31) / 32 : 1;
    static constexpr int NRED = 4;
    static constexpr int VPT = (NV + NT - 1) / NT;       // variables (box-row owners) per thread
    static constexpr int KRAW = 40;                      // obstacle capacity of the ABI (lscqp_config.max_obs <= 40)
    // register budget: 65536 / (MIN_CTAS * NT) per thread (the light instances are shared-memory bound)
    static constexpr int MIN_CTAS = NT > 128 ? 2 : (NT == 128 ? LSCQP_FULL_MINCTAS : LSCQP_LIGHT_MINCTAS);
    static_assert(!COMM || VPT == 1, "communication-range instances keep one variable per thread");
    static_assert(KMAX <= KRAW, "KMAX beyond the ABI capacity");
    // shared memory layout (doubles).  Two regions are shared by buffers whose lifetimes do not overlap:
    //   * the 6x6 blocks (projection source, live from the block build to the end of the projection) start on top of
    //     the two full-space directions dca / dc (dead between sweep A and the back-substitution of the predictor);
    //   * the row weights wB / wV / wA (live from sweep A to the block build) sit in the reduced matrix A
    //     (written by the projection, dead again once the corrector is solved).
    static constexpr int O_Q2 = 0;
    static constexpr int O_C = O_Q2 + 36;
    static constexpr int O_Y = O_C + NV;
    static constexpr int O_DY = O_Y + NR;
    static constexpr int O_X0 = O_DY + NR;               // [D][3]
    static constexpr int O_VLIM = O_X0 + 3 * D;          // [D]
    static constexpr int O_ALIM = O_VLIM + D;            // [D]
    static constexpr int O_GOAL = O_ALIM + D;            // [D]
    static constexpr int O_ORG = O_GOAL + D;             // [D] origin of the local coordinates: the agent's current position
    static constexpr int O_LB = O_ORG + D;               // [D][M]
    static constexpr int O_UB = O_LB + D * M;            // [D][M]
    static constexpr int O_TERMW = O_UB + D * M;         // [M] terminal weights, then the distance to the goal
    static constexpr int O_NRM = O_TERMW + M + 1;        // [KMAX][M][3]
    static constexpr int O_SLABT = O_NRM + KMAX * M * 3; // [G][NCP][D]
    static constexpr int O_UB_ = O_SLABT + G * NCP * D;  // [NV] x3: uB uV uA
    static constexpr int O_UV = O_UB_ + NV;
    static constexpr int O_UA = O_UV + NV;
    static constexpr int O_DCA = O_UA + NV;              // [NV] affine direction, [NV] combined direction ...
    static constexpr int O_DC = O_DCA + NV;
    static constexpr int O_BLK = O_DCA;                  // ... under [D][M][36] projection sources: blocks, then slab 0 of S
    static constexpr int O_SLABS = O_BLK + D * M * 36;   // [G][NCP][NS]
    static constexpr int O_RFULL = O_SLABS + G * NCP * NS;   // [NV]
    static constexpr int O_A = O_RFULL + NV;             // reduced matrix (A_SIZE) ...
    static constexpr int O_WB = O_A;                     // ... over [NV] x3: wB wV wA
    static constexpr int O_WV = O_WB + NV;
    static constexpr int O_WA = O_WV + NV;
    static constexpr int O_RHS = O_A + A_SIZE;           // [NRP]
    static constexpr int O_DIAG0 = O_RHS + NRP;          // [NRP]
    static constexpr int O_INVD = O_DIAG0 + NRP;         // [NB][6] inverse diagonal blocks
    static constexpr int O_SBUF = O_INVD + 2 * NRP;      // [rows][3] scaled panel rows of the running factorisation
    static constexpr int O_WC = O_SBUF + (COMM ? 3 * NRP : 3 * BW);   // [PC] comm pair weights (projection source), then [PC] rhs multipliers
    static constexpr int O_RED = O_WC + 2 * PC;          // [2][NW][NRED]
    static constexpr int O_ACT = O_RED + 2 * NW * NRED;  // int[KRAW]: original obstacle index of each kept slot, int keep[KRAW], int n_act
    static constexpr int O_END = O_ACT + KRAW + 2;
    static_assert(A_SIZE >= 3 * NV, "row weights do not fit under the reduced matrix");
    static_assert(D * M * 36 >= 2 * NV, "directions do not fit under the projection blocks");
    static constexpr int SMEM_BYTES = O_END * 8;
    // dual_out layout: [KRAW][M][6] LSC rows, then [NV][6] box rows (lb, ub, vel+, vel-, acc+, acc-)
    static constexpr int DUAL_STRIDE = KRAW * M * 6 + NV * 6 + 2 * PC;   // (+ comm pairs: upper-side, lower-side multiplier)
};

// continuity map c[m][0..2] = T y[m-1][3..5]
__device__ __forceinline__ double tmap(int a, double y3, double y4, double y5) {
    return a == 0 ? y5 : (a == 1 ? 2.0 * y5 - y4 : y3 - 4.0 * y4 + 4.0 * y5);
}
__device__ __forceinline__ double tcoef(int a, int j) {   // T[a][j]
    // a=0: (0,0,1)  a=1: (0,-1,2)  a=2: (1,-4,4)
    return a == 0 ? (j == 2 ? 1.0 : 0.0) : (a == 1 ? (j == 0 ? 0.0 : (j == 1 ? -1.0 : 2.0))
                                                    : (j == 0 ? 1.0 : (j == 1 ? -4.0 : 4.0)));
}

template <class C>
__device__ __forceinline__ int ridx(int s, int k, int j) {
    if (C::TERM && s == C::M - 1) return (C::M - 1) * C::NZS + k;
    return s * C::NZS + k * 3 + j;
}

// full-space control point (k, m, i) from reduced vector yv and fixed initial points x0 (null for directions)
template <class C>
__device__ __forceinline__ double full_from_reduced(const double* yv, const double* x0, int k, int m, int i) {
    if (i >= 3) return yv[ridx<C>(m, k, i - 3)];
    if (m == 0) return x0 ? x0[k * 3 + i] : 0.0;
    const double y3 = yv[ridx<C>(m - 1, k, 0)], y4 = yv[ridx<C>(m - 1, k, 1)], y5 = yv[ridx<C>(m - 1, k, 2)];
    return tmap(i, y3, y4, y5);
}

// reduced component r of Z^T v for a full-space vector v [D][NCP]
template <class C>
__device__ __forceinline__ double reduce_from_full(const double* v, int r) {
    if (C::TERM && r >= (C::M - 1) * C::NZS) {
        const int k = r - (C::M - 1) * C::NZS;
        const double* p = v + k * C::NCP + (C::M - 1) * 6;
        return p[3] + p[4] + p[5];
    }
    const int s = r / C::NZS, k = (r % C::NZS) / 3, j = r % 3;
    const double* p = v + k * C::NCP + s * 6;
    double val = p[3 + j];
    if (s + 1 < C::M) {
        const double* q = p + 6;
        val += tcoef(0, j) * q[0] + tcoef(1, j) * q[1] + tcoef(2, j) * q[2];
    }
    return val;
}

// support of reduced variable r in the full space (same dimension k): up to 4 (cp, coef) pairs
template <class C>
__device__ __forceinline__ int support(int r, int& k, int* cp, double* co) {
    if (C::TERM && r >= (C::M - 1) * C::NZS) {
        k = r - (C::M - 1) * C::NZS;
        for (int a = 0; a < 3; a++) { cp[a] = (C::M - 1) * 6 + 3 + a; co[a] = 1.0; }
        return 3;
    }
    const int s = r / C::NZS, j = r % 3;
    k = (r % C::NZS) / 3;
    cp[0] = s * 6 + 3 + j; co[0] = 1.0;
    int n = 1;
    if (s + 1 < C::M) {
        for (int a = 0; a < 3; a++) {
            const double t = tcoef(a, j);
            if (t != 0.0) { cp[n] = (s + 1) * 6 + a; co[n] = t; n++; }
        }
    }
    return n;
}

__device__ __forceinline__ int symidx3(int a, int b) {   // (0,0)(0,1)(0,2)(1,1)(1,2)(2,2)
    if (a > b) { int t = a; a = b; b = t; }
    return a == 0 ? b : (a == 1 ? 2 + b : 5);
}
__device__ __forceinline__ int symidx2(int a, int b) {   // (0,0)(0,1)(1,1)
    return a + b;
}

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// CTA barrier; a one-warp CTA only needs the warp-level one
template <class C>
__device__ __forceinline__ void cta_sync() {
    if (C::NW == 1) __syncwarp(); else __syncthreads();
}

// CTA-wide all-reduce of 4 values: v[0], v[1] summed, v[2] min, v[3] max.  One barrier (ping-pong scratch).   // @phase reduce4
template <class C>
__device__ __forceinline__ void block_reduce4(double* v, double* red, int& phase) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double a = warp_sum(v[0]), b = warp_sum(v[1]), c = warp_min(v[2]), d = warp_max(v[3]);
    if (C::NW == 1) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; __syncwarp(); return; }
    double* buf = red + phase * C::NW * C::NRED;
    if (lane == 0) { buf[warp * 4 + 0] = a; buf[warp * 4 + 1] = b; buf[warp * 4 + 2] = c; buf[warp * 4 + 3] = d; }
    __syncthreads();
    a = 0; b = 0; c = INFINITY; d = -INFINITY;
    for (int w = 0; w < C::NW; w++) {
        a += buf[w * 4 + 0]; b += buf[w * 4 + 1]; c = fmin(c, buf[w * 4 + 2]); d = fmax(d, buf[w * 4 + 3]);
    }
    v[0] = a; v[1] = b; v[2] = c; v[3] = d;
    phase ^= 1;
}

// ---------------------------------------------------------------------------------------------
// Banded block-LDL^T factorisation and solves in shared memory, executed by warp 0 only (callers bracket
// with __syncthreads).  Phi = Lb Db Lb^T with 3x3 diagonal blocks Db_J (the running Schur complements P_J) and a
// unit-block-diagonal Lb.  Right-looking over 3-column panels:
//   * every lane factorises P_J = L D L^T (3x3, closed form, reciprocals by rcp + Newton) from broadcast loads;
//   * lane l finishes row i = c0 + 3 + l of the panel in registers: t = a L^-T, s = t D^-1, abar = s L^-1 (= a P_J^-1);
//   * the trailing window is updated with 3-term products  A_ik -= s_i . t_k ;
//   * abar replaces the panel entries (that is Lb), P_J^-1 is kept for the solves.
// The solves then need no per-panel triangular solve on their dependency chain:
//   forward  w_J = b_J - sum_{I<J} Lb_JI w_I ;  middle  v_J = P_J^-1 w_J ;  backward  x_J = v_J - sum_{I>J} Lb_IJ^T x_I.
// 1/x on the factorisation's critical path: hardware seed + two Newton steps (1e-12 relative; the interior-point
// iteration is self-correcting, the residuals never go through the factor)
__device__ __forceinline__ double pivot_rcp(double x) {
#ifdef LSCQP_CUDA_EMUL
    return 1.0 / x;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
#endif
}

// rows below panel J that can be non-zero (up to the end of the next stage)
template <class C>
__device__ __forceinline__ int panel_rows_below(int J) {
    const int last = (J * 3 / C::NZS + 2) * C::NZS;
    return (last < C::NRP ? last : C::NRP) - 3 * J - 3;
}

template <class C>
__device__ __forceinline__ int chol_banded(double* A, const double* diag0, double* pinv, double* sbuf,
                                           const int* pr_i, const int* pr_k) {   // @phase chol
    const int lane = threadIdx.x & 31;
    int bad = 0;
    for (int J = 0; J < C::NB; J++) {
        const int c0 = 3 * J;
        const double p11 = A[c0 * C::LD + c0], p21 = A[(c0 + 1) * C::LD + c0], p31 = A[(c0 + 2) * C::LD + c0];
        const double p22 = A[(c0 + 1) * C::LD + c0 + 1], p32 = A[(c0 + 2) * C::LD + c0 + 1], p33 = A[(c0 + 2) * C::LD + c0 + 2];
        const int i = c0 + 3 + lane;
        // the matrix is block tridiagonal over stages (NZS variables each): below the panel only the rest of its
        // stage and the next stage are non-zero, and block elimination keeps it that way
        const int R = panel_rows_below<C>(J);
        const bool row = lane < R;
        double a0 = 0, a1 = 0, a2 = 0;
        if (row) { a0 = A[i * C::LD + c0]; a1 = A[i * C::LD + c0 + 1]; a2 = A[i * C::LD + c0 + 2]; }
        // 3x3 LDL^T with pivot guard (a non-positive pivot freezes that direction)
        double d1 = p11;
        if (!(d1 > 1e-13 * diag0[c0])) { d1 = 1e300; bad++; }
        const double r1 = pivot_rcp(d1);
        const double l21 = p21 * r1, l31 = p31 * r1;
        double d2 = p22 - l21 * p21;
        if (!(d2 > 1e-13 * diag0[c0 + 1])) { d2 = 1e300; bad++; }
        const double r2 = pivot_rcp(d2);
        const double u32 = p32 - l31 * p21, l32 = u32 * r2;
        double d3 = p33 - l31 * p31 - l32 * u32;
        if (!(d3 > 1e-13 * diag0[c0 + 2])) { d3 = 1e300; bad++; }
        const double r3 = pivot_rcp(d3);
        // L^-1 = [1 0 0; m21 1 0; m31 m32 1]
        const double m21 = -l21, m32 = -l32, m31 = l21 * l32 - l31;
        // row of the panel: t = a L^-T, s = t D^-1, abar = s L^-1
        const double t0 = a0, t1 = a1 + m21 * a0, t2 = a2 + m31 * a0 + m32 * a1;
        const double s0 = t0 * r1, s1 = t1 * r2, s2 = t2 * r3;
        const double ab0 = s0 + m21 * s1 + m31 * s2, ab1 = s1 + m32 * s2, ab2 = s2;
        __syncwarp();                                           // every lane has read the panel
        if (row) {
            A[i * C::LD + c0] = t0; A[i * C::LD + c0 + 1] = t1; A[i * C::LD + c0 + 2] = t2;
            sbuf[lane * 3] = s0; sbuf[lane * 3 + 1] = s1; sbuf[lane * 3 + 2] = s2;
        }
        if (lane == 0) {
            // P^-1 = L^-T D^-1 L^-1 (symmetric): 11 12 13 22 23 33
            double* q = pinv + 6 * J;
            q[0] = r1 + m21 * m21 * r2 + m31 * m31 * r3; q[1] = m21 * r2 + m31 * m32 * r3; q[2] = m31 * r3;
            q[3] = r2 + m32 * m32 * r3; q[4] = m32 * r3; q[5] = r3;
        }
        __syncwarp();
        const int npair = R * (R + 1) / 2;                      // pairs (di >= dk) with di < R come first in the lane order
#pragma unroll
        for (int t = 0; t < C::NPR; t++) {
            if (32 * t >= npair) break;
            const int ii = c0 + 3 + pr_i[t], kk = c0 + 3 + pr_k[t];
            if (lane + 32 * t < npair) {
                const double* si = sbuf + pr_i[t] * 3;
                const double* tk = A + kk * C::LD + c0;
                A[ii * C::LD + kk] -= si[0] * tk[0] + si[1] * tk[1] + si[2] * tk[2];
            }
        }
        __syncwarp();
        if (row) { A[i * C::LD + c0] = ab0; A[i * C::LD + c0 + 1] = ab1; A[i * C::LD + c0 + 2] = ab2; }
        // (the next panel reads its diagonal block and rows from columns >= c0 + 3 only; the writes above are ordered
        //  before any later read of these columns by the barriers of the next iterations / the caller)
    }
    __syncwarp();
    return bad;
}

// solves Phi x = b in place (b in shared memory) with the factors left by chol_banded
template <class C>
__device__ __forceinline__ void chol_solve(const double* A, const double* pinv, double* b) {   // @phase trisolve
    const int lane = threadIdx.x & 31;
    for (int J = 0; J < C::NB; J++) {
        const int c0 = 3 * J;
        const double w0 = b[c0], w1 = b[c0 + 1], w2 = b[c0 + 2];
        const int i = c0 + 3 + lane;
        if (lane < panel_rows_below<C>(J)) {
            const double* ab = A + i * C::LD + c0;
            b[i] = (b[i] - ab[0] * w0) - (ab[1] * w1 + ab[2] * w2);
        }
        __syncwarp();
    }
    if (lane < C::NB) {
        const double* q = pinv + 6 * lane;
        const double w0 = b[3 * lane], w1 = b[3 * lane + 1], w2 = b[3 * lane + 2];
        b[3 * lane] = q[0] * w0 + q[1] * w1 + q[2] * w2;
        b[3 * lane + 1] = q[1] * w0 + q[3] * w1 + q[4] * w2;
        b[3 * lane + 2] = q[2] * w0 + q[4] * w1 + q[5] * w2;
    }
    __syncwarp();
    for (int J = C::NB - 1; J >= 0; J--) {
        const int c0 = 3 * J;
        const double x0 = b[c0], x1 = b[c0 + 1], x2 = b[c0 + 2];
        const int i = c0 - 1 - lane;                            // rows above the panel that couple to it: back to the
        const int first = (c0 / C::NZS - 1) * C::NZS;           // start of the previous stage
        if (i >= first && i >= 0)
            b[i] = (b[i] - A[c0 * C::LD + i] * x0) - (A[(c0 + 1) * C::LD + i] * x1 + A[(c0 + 2) * C::LD + i] * x2);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Dense variant for the instances with communication-range rows (their end-point couplings fill the band).
// Same block-LDL^T recurrences; the panel rows are looped in chunks of 32 by warp 0 and the trailing update is
// spread over the whole CTA (two CTA barriers per panel).  These instances serve the small reference missions
// (10 agents, launch/simulation.launch), not the throughput benchmark.
template <class C>
__device__ __forceinline__ int chol_dense_cta(double* A, const double* diag0, double* pinv, double* sbuf) {
    constexpr int CH = (C::NRP + 31) / 32;
    const int tid = threadIdx.x, lane = tid & 31;
    int bad = 0;
    for (int J = 0; J < C::NB; J++) {
        const int c0 = 3 * J;
        double ab[CH][3];
        if (tid < 32) {
            const double p11 = A[c0 * C::LD + c0], p21 = A[(c0 + 1) * C::LD + c0], p31 = A[(c0 + 2) * C::LD + c0];
            const double p22 = A[(c0 + 1) * C::LD + c0 + 1], p32 = A[(c0 + 2) * C::LD + c0 + 1], p33 = A[(c0 + 2) * C::LD + c0 + 2];
            double d1 = p11;
            if (!(d1 > 1e-13 * diag0[c0])) { d1 = 1e300; bad++; }
            const double r1 = pivot_rcp(d1);
            const double l21 = p21 * r1, l31 = p31 * r1;
            double d2 = p22 - l21 * p21;
            if (!(d2 > 1e-13 * diag0[c0 + 1])) { d2 = 1e300; bad++; }
            const double r2 = pivot_rcp(d2);
            const double u32 = p32 - l31 * p21, l32 = u32 * r2;
            double d3 = p33 - l31 * p31 - l32 * u32;
            if (!(d3 > 1e-13 * diag0[c0 + 2])) { d3 = 1e300; bad++; }
            const double r3 = pivot_rcp(d3);
            const double m21 = -l21, m32 = -l32, m31 = l21 * l32 - l31;
            double tt[CH][3];
#pragma unroll
            for (int ch = 0; ch < CH; ch++) {
                const int i = c0 + 3 + lane + 32 * ch;
                double a0 = 0, a1 = 0, a2 = 0;
                if (i < C::NRP) { a0 = A[i * C::LD + c0]; a1 = A[i * C::LD + c0 + 1]; a2 = A[i * C::LD + c0 + 2]; }
                tt[ch][0] = a0; tt[ch][1] = a1 + m21 * a0; tt[ch][2] = a2 + m31 * a0 + m32 * a1;
                const double s0 = tt[ch][0] * r1, s1 = tt[ch][1] * r2, s2 = tt[ch][2] * r3;
                ab[ch][0] = s0 + m21 * s1 + m31 * s2; ab[ch][1] = s1 + m32 * s2; ab[ch][2] = s2;
            }
            __syncwarp();
#pragma unroll
            for (int ch = 0; ch < CH; ch++) {
                const int i = c0 + 3 + lane + 32 * ch;
                if (i < C::NRP) {
                    A[i * C::LD + c0] = tt[ch][0]; A[i * C::LD + c0 + 1] = tt[ch][1]; A[i * C::LD + c0 + 2] = tt[ch][2];
                    double* sb = sbuf + (i - c0 - 3) * 3;
                    sb[0] = tt[ch][0] * r1; sb[1] = tt[ch][1] * r2; sb[2] = tt[ch][2] * r3;
                }
            }
            if (lane == 0) {
                double* q = pinv + 6 * J;
                q[0] = r1 + m21 * m21 * r2 + m31 * m31 * r3; q[1] = m21 * r2 + m31 * m32 * r3; q[2] = m31 * r3;
                q[3] = r2 + m32 * m32 * r3; q[4] = m32 * r3; q[5] = r3;
            }
        }
        __syncthreads();
        const int w = C::NRP - c0 - 3;
        for (int e = tid; e < w * (w + 1) / 2; e += C::NT) {
            int di = (int) ((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while ((di + 1) * (di + 2) / 2 <= e) di++;
            while (di * (di + 1) / 2 > e) di--;
            const int dk = e - di * (di + 1) / 2;
            const double* si = sbuf + di * 3;
            const double* tk = A + (c0 + 3 + dk) * C::LD + c0;
            A[(c0 + 3 + di) * C::LD + c0 + 3 + dk] -= si[0] * tk[0] + si[1] * tk[1] + si[2] * tk[2];
        }
        __syncthreads();
        if (tid < 32) {
#pragma unroll
            for (int ch = 0; ch < CH; ch++) {
                const int i = c0 + 3 + lane + 32 * ch;
                if (i < C::NRP) { A[i * C::LD + c0] = ab[ch][0]; A[i * C::LD + c0 + 1] = ab[ch][1]; A[i * C::LD + c0 + 2] = ab[ch][2]; }
            }
        }
    }
    __syncthreads();
    return bad;
}

template <class C>
__device__ __forceinline__ void solve_dense_warp(const double* A, const double* pinv, double* b) {
    const int lane = threadIdx.x & 31;
    for (int J = 0; J < C::NB; J++) {
        const int c0 = 3 * J;
        const double w0 = b[c0], w1 = b[c0 + 1], w2 = b[c0 + 2];
        for (int i = c0 + 3 + lane; i < C::NRP; i += 32) {
            const double* ab = A + i * C::LD + c0;
            b[i] = (b[i] - ab[0] * w0) - (ab[1] * w1 + ab[2] * w2);
        }
        __syncwarp();
    }
    if (lane < C::NB) {
        const double* q = pinv + 6 * lane;
        const double w0 = b[3 * lane], w1 = b[3 * lane + 1], w2 = b[3 * lane + 2];
        b[3 * lane] = q[0] * w0 + q[1] * w1 + q[2] * w2;
        b[3 * lane + 1] = q[1] * w0 + q[3] * w1 + q[4] * w2;
        b[3 * lane + 2] = q[2] * w0 + q[4] * w1 + q[5] * w2;
    }
    __syncwarp();
    for (int J = C::NB - 1; J >= 0; J--) {
        const int c0 = 3 * J;
        const double x0 = b[c0], x1 = b[c0 + 1], x2 = b[c0 + 2];
        for (int i = c0 - 1 - lane; i >= 0; i -= 32)
            b[i] = (b[i] - A[c0 * C::LD + i] * x0) - (A[(c0 + 1) * C::LD + i] * x1 + A[(c0 + 2) * C::LD + i] * x2);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Row algebra.  Row convention: q(c) >= 0, slack s > 0, multiplier lam > 0.  Every row starts with
// the same primal residual rp = s - q (the start-up shift) and every step scales it by (1 - alpha),
// so rp is one scalar for the whole QP and neither q nor the row constants are needed after start-up.
// 1/x for normal positive x: hardware seed (2^-20) + two Newton steps (1e-12 relative), no special-case branch.
// It only scales Newton directions and multiplier updates; residuals never pass through it.
__device__ __forceinline__ double fast_rcp(double x) {
#ifdef LSCQP_CUDA_EMUL
    return 1.0 / x;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
#endif
}

// Step-to-boundary ratio test min_r v_r / (-dv_r) over dv_r < 0.  Kept in fp32: the step is multiplied by 0.99
// afterwards, so the 1e-7 relative error of the float quotient cannot push an iterate through its bound, and the
// value only gates "full step or not" otherwise.
struct MinRatio {   // @phase minratio
    float best;
    __device__ __forceinline__ void init() { best = 3.0e38f; }
    __device__ __forceinline__ void add(double v, double dv) {
        const float r = __fdividef((float) v, -(float) dv);
        best = (dv < 0.0) ? fminf(best, r) : best;
    }
    __device__ __forceinline__ double value() const { return (double) best; }
};

// ---------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(C::NT, C::MIN_CTAS)
pdip_solve_kernel(const SolveParams p) {   // @phase setup
    constexpr int M = C::M, D = C::D, NCP = C::NCP, NV = C::NV, NR = C::NR, NS = C::NS;
    constexpr int G = C::G, KPT = C::KPT, NT = C::NT;
    LSCQP_DYN_SMEM(sm);
    const int tid = threadIdx.x, lane = tid & 31;
    const int agent = blockIdx.x;
    if (agent >= p.n_agents) return;
    if (p.klass_mode == 2 && p.klass[agent] == 0) return;                // solved by the first pass already

    double* sQ2 = sm + C::O_Q2;
    double* s_c = sm + C::O_C;
    double* s_dca = sm + C::O_DCA;
    double* s_dc = sm + C::O_DC;
    double* s_y = sm + C::O_Y;
    double* s_dy = sm + C::O_DY;
    double* s_x0 = sm + C::O_X0;
    double* s_vlim = sm + C::O_VLIM;
    double* s_alim = sm + C::O_ALIM;
    double* s_goal = sm + C::O_GOAL;
    double* s_org = sm + C::O_ORG;
    double* s_lb = sm + C::O_LB;
    double* s_ub = sm + C::O_UB;
    double* s_termw = sm + C::O_TERMW;
    double* s_nrm = sm + C::O_NRM;
    double* slabS = sm + C::O_SLABS;
    double* slabT = sm + C::O_SLABT;
    double* s_wB = sm + C::O_WB;
    double* s_wV = sm + C::O_WV;
    double* s_wA = sm + C::O_WA;
    double* s_uB = sm + C::O_UB_;
    double* s_uV = sm + C::O_UV;
    double* s_uA = sm + C::O_UA;
    double* s_blk = sm + C::O_BLK;
    double* s_rfull = sm + C::O_RFULL;
    double* s_A = sm + C::O_A;
    double* s_rhs = sm + C::O_RHS;
    double* s_diag0 = sm + C::O_DIAG0;
    double* s_invd = sm + C::O_INVD;
    double* s_red = sm + C::O_RED;
    double* s_wC = sm + C::O_WC;           // comm pair weights [PC], then rhs multipliers [PC]
    int red_phase = 0;

    // ---- thread roles
    const int grp = tid / C::CPW, cp = tid % C::CPW;
    const bool cp_valid = cp < NCP;
    const int m_cp = cp / 6, i_cp = cp % 6;
    const bool lsc_thread = cp_valid && !(m_cp == 0 && i_cp < 3);       // traj_optimizer.cpp:404
    // box rows: variable v = tid + u NT (u < VPT) is owned by this thread; rows 0 lb, 1 ub, 2 vel+, 3 vel-, 4 acc+, 5 acc-
    constexpr int VPT = C::VPT, NBX = C::NBX;
    // communication-range pair of this thread (traj_optimizer.cpp:477-500), on the segment end points E_a = c[a][5]:
    //   |E_a - E_b| <= range/2 - radius  for b < a  (c[mi][0] = E_{mi-1});  E_a within range/2 - radius of the
    //   current position (mi = 0) and within range/2 - 1e-5 of next_waypoint (merged into one interval)
    const bool has_comm = C::COMM && tid < C::PC;
    int ca_idx = 0, cb_idx = -1;
    double c_hp = 0.0, c_hm = 0.0;
    if (has_comm) {
        const int kc = tid / C::PP, r = tid % C::PP;
        const double r1 = 0.5 * p.comm_range - p.limits[agent * 8 + 6], r2 = 0.5 * p.comm_range - 1e-5;
        if (r < M) {
            // (local coordinates: the current position is the origin)
            const double wp = (double) p.next_waypoint[agent * 3 + kc] - (double) p.state[agent * 9 + kc];
            ca_idx = kc * NCP + r * 6 + 5;
            c_hp = fmin(r1, wp + r2); c_hm = fmax(-r1, wp - r2);
        } else {
            int e = r - M, a = 1;
            while (a * (a + 1) / 2 <= e) a++;                     // pair e -> (a, b), 0 <= b < a <= M-1
            const int b = e - a * (a - 1) / 2;
            ca_idx = kc * NCP + a * 6 + 5; cb_idx = kc * NCP + b * 6 + 5;
            c_hp = r1; c_hm = -r1;
        }
    }
    unsigned bmask[VPT];
#pragma unroll
    for (int u = 0; u < VPT; u++) {
        const int v = tid + u * NT, m_v = (v % NCP) / 6, i_v = v % 6;
        const bool on = v < NV;
        const bool has_bnd = on && !(m_v == 0 && i_v < 3);              // :260-265
        const bool has_vel = on && i_v < 5 && !(m_v == 0 && i_v < 2);   // :444
        const bool has_acc = on && i_v < 4 && !(m_v == 0 && i_v < 1);   // :458
        bmask[u] = (has_bnd ? 3u : 0u) | (has_vel ? 12u : 0u) | (has_acc ? 48u : 0u) | ((u == 0 && has_comm) ? 192u : 0u);
    }

    const int obs0 = p.obs_offsets[agent];
    int K = p.obs_offsets[agent + 1] - obs0;
    // The reference's model takes every obstacle it is handed (traj_optimizer.cpp:400-437 loops over getObsSize()):
    // a list this instance cannot hold is reported through status_out (LSCQP_CAPACITY), never silently shortened.
    bool cap_fail = K < 0 || K > C::KRAW || K > p.max_obs;
    if (cap_fail) {
        if (p.klass_mode == 1) { if (tid == 0) p.klass[agent] = 1; return; }   // the full-capacity pass reports it
        K = 0;
    }
    if (p.rsfc) K = 0;                                                    // (rows with a free, cost-less slack: see SolveParams::rsfc)

    // ---- Cholesky trailing-update pair assignment (fixed per lane)
    int pr_i[C::NPR], pr_k[C::NPR];
#pragma unroll
    for (int t = 0; t < C::NPR; t++) {
        const int e = lane + 32 * t;
        pr_i[t] = -1; pr_k[t] = 0;
        if (e < C::NPAIR) {
            int di = (int) ((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while ((di + 1) * (di + 2) / 2 <= e) di++;
            while (di * (di + 1) / 2 > e) di--;
            pr_i[t] = di; pr_k[t] = e - di * (di + 1) / 2;
        }
    }

    // ---- stage per-agent constants
    for (int e = tid; e < 36; e += NT) sQ2[e] = p.Q2[e];
    if (tid < D) {
        const int k = tid;
        const double pos = (double) p.state[agent * 9 + k], vel = (double) p.state[agent * 9 + 3 + k],
                     acc = (double) p.state[agent * 9 + 6 + k];
        // c0 = pos; 5/dt (c1 - c0) = vel; 20/dt^2 (c2 - 2 c1 + c0) = acc    (traj_optimizer.cpp:321-338), relative to pos
        const double c0 = 0.0, c1 = vel * p.dt / 5.0, c2 = acc * p.dt * p.dt / 20.0 + 2.0 * c1 - c0;
        s_org[k] = pos;
        s_x0[k * 3 + 0] = c0; s_x0[k * 3 + 1] = c1; s_x0[k * 3 + 2] = c2;
        s_vlim[k] = p.limits[agent * 8 + k] * p.dt / 5.0;                 // :448-453 scaled to unit coefficients
        s_alim[k] = p.limits[agent * 8 + 3 + k] * p.dt * p.dt / 20.0;     // :462-471
        s_goal[k] = (double) p.goal[agent * 3 + k] - pos;
        for (int m = 0; m < M; m++) {
            double lo = p.world_min[k], hi = p.world_max[k];              // :252-253
            if (p.rsfc && k == 2 && m == 0) { lo = -100.0; hi = 100.0; }  // :255-258
            if (p.use_sfc) {                                               // :372-397, Box::convertToLSCs
                lo = fmax(lo, (double) p.sfc[((size_t) agent * M + m) * 6 + k]);
                hi = fmin(hi, (double) p.sfc[((size_t) agent * M + m) * 6 + 3 + k]);
            }
            s_lb[k * M + m] = lo - pos; s_ub[k * M + m] = hi - pos;
        }
    }
    if (tid == (NT > 32 ? 32 : 8)) {
        // getTerminalSegments_old, traj_optimizer.cpp:530-538 (float norm of the point3d difference)
        const float dx = p.goal[agent * 3 + 0] - p.state[agent * 9 + 0], dy = p.goal[agent * 3 + 1] - p.state[agent * 9 + 1],
                    dz = p.goal[agent * 3 + 2] - p.state[agent * 9 + 2];
        const float nsq = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        const double flight = sqrt((double) nsq) / p.limits[agent * 8 + 7];
        int ts = (int) ((M * p.dt - flight + 1e-9) / p.dt);
        if (ts < 1) ts = 1;
        for (int m = 0; m < M; m++) s_termw[m] = (m >= M - ts) ? 2.0 * p.w_t : 0.0;
        s_termw[M] = sqrt((double) nsq);
    }
    // ---- presolve: an obstacle is dropped when every one of its rows is strictly satisfied by *any* point that
    // obeys the velocity rows (traj_optimizer.cpp:443-454): chaining |c[j+1] - c[j]| <= vmax dt/5 from the fixed
    // control point c[0][2] bounds control point (m, i) to a box of half-width (5m + i - 2) vmax_k dt / 5 around it.
    // Such rows are redundant for the feasible set, so the minimiser (and its multipliers: zero) is unchanged.
    int* s_act = reinterpret_cast<int*>(sm + C::O_ACT);
    int* s_keep = s_act + C::KRAW;
    for (int e = tid; e < C::KRAW; e += NT) s_keep[e] = (e < K && !p.presolve) ? 1 : 0;
    cta_sync<C>();
    if (p.presolve) {
        // the same reach box makes the bound rows (world box / SFC box, :238-270, :372-397) of a control point redundant
        // when it lies strictly inside them: dropped (multipliers zero), exact like the obstacle rows below
#pragma unroll
        for (int u = 0; u < VPT; u++) {
            const int v = tid + u * NT;
            if (!(bmask[u] & 3u)) continue;
            const int k_v = v / NCP, m_v = (v % NCP) / 6, i_v = v % 6;
            const double reach = (double) (5 * m_v + i_v - 2) * s_vlim[k_v];
            if (s_x0[k_v * 3 + 2] - reach - s_lb[k_v * M + m_v] > 1e-6 && s_ub[k_v * M + m_v] - (s_x0[k_v * 3 + 2] + reach) > 1e-6)
                bmask[u] &= ~3u;
        }
    }
    if (p.presolve && lsc_thread) {
        const double steps = (double) (5 * m_cp + i_cp - 2);
        for (int oi = grp; oi < K; oi += G) {
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
    cta_sync<C>();
    for (int e = tid; e < K; e += NT) {
        if (!s_keep[e]) continue;
        int pos = 0;
        for (int u = 0; u < e; u++) pos += s_keep[u];
        if (pos < C::KMAX) s_act[pos] = e;
    }
    if (tid == 0) {
        int n = 0;
        for (int u = 0; u < K; u++) n += s_keep[u];
        s_keep[C::KRAW] = n;
    }
    cta_sync<C>();
    K = s_keep[C::KRAW];                                               // kept obstacles, compacted into slots 0..K-1
    if (p.klass_mode == 1) {
        // light instance: agents with more kept obstacles than it holds go to the full-capacity instance
        const bool over = K > C::KMAX;
        if (tid == 0) p.klass[agent] = over ? 1 : 0;
        if (over) return;
    }
    if (K > C::KMAX) { cap_fail = true; K = 0; }                          // (only without presolve, or on the compact instance)
    for (int e = tid; e < K * M; e += NT) {
        // rows with a (float) normal shorter than SP_EPSILON_FLOAT are skipped by the reference
        // (traj_optimizer.cpp:409-411): here they become the constant row 0.c >= -1, which never binds
        const double* g = p.normals + ((size_t) (obs0 + s_act[e / M]) * M + e % M) * 3;
        double nx = g[0], ny = g[1], nz = g[2];
        const float fx = (float) nx, fy = (float) ny, fz = (float) nz;
        const float nsq = __fadd_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)), __fmul_rn(fz, fz));
        if (sqrt((double) nsq) < 1e-5) { nx = 0.0; ny = 0.0; nz = 0.0; }
        s_nrm[e * 3] = nx; s_nrm[e * 3 + 1] = ny; s_nrm[e * 3 + 2] = nz;
    }
    // full-space control points of the reduced vector yv (x0: the fixed initial points, null for directions)
    auto expand = [&](const double* yv, const double* x0, double* out) {
#pragma unroll
        for (int u = 0; u < VPT; u++) {
            const int v = tid + u * NT;
            if (v < NV) out[v] = full_from_reduced<C>(yv, x0, v / NCP, (v % NCP) / 6, v % 6);
        }
    };
    // starting point: the free control points of initial_traj when given (the reference hands it to
    // TrajOptimizer::solve, traj_optimizer.cpp:18-21), else every free control point at the current position
    bool warm = p.warm_traj != nullptr;
    auto set_start = [&](bool from_traj) {
        for (int r = tid; r < NR; r += NT) {
            int st, k, j;
            if (C::TERM && r >= (M - 1) * C::NZS) { st = M - 1; k = r - (M - 1) * C::NZS; j = 2; }
            else { st = r / C::NZS; k = (r % C::NZS) / 3; j = r % 3; }
            s_y[r] = from_traj ? (double) p.warm_traj[(((size_t) agent * M + st) * 6 + 3 + j) * 3 + k] - s_org[k] : 0.0;
        }
        cta_sync<C>();
        expand(s_y, s_x0, s_c);
        cta_sync<C>();
    };
    set_start(warm);
    if (cap_fail) {
        // finite outputs: the starting point (initial_traj when given), no multipliers
#pragma unroll
        for (int u = 0; u < VPT; u++) {
            const int v = tid + u * NT;
            if (v < NV) p.ctrl_out[(size_t) agent * NV + v] = s_c[v] + s_org[v / NCP];
        }
        if (tid == 0) {
            p.cost_out[agent] = 0.0; p.status_out[agent] = ST_CAPACITY;
            if (p.iters_out) p.iters_out[agent] = 0;
            if (p.kkt_out) for (int e = 0; e < 4; e++) p.kkt_out[agent * 4 + e] = 0.0;
        }
        if (p.dual_out) for (int e = tid; e < p.dual_stride; e += NT) p.dual_out[(size_t) agent * p.dual_stride + e] = 0.0;
        return;
    }

    // ---- per-thread row state (registers): slack and multiplier of every owned row
    double ls[KPT], ll[KPT];
    double bs[VPT][NBX], bl[VPT][NBX];
    // rows of this thread: obstacles grp, grp + G, ... < K on its control point (none for the fixed points)
    const int nrow = (lsc_thread && K > grp) ? (K - grp + G - 1) / G : 0;
#pragma unroll
    for (int j = 0; j < KPT; j++) { ls[j] = 1.0; ll[j] = 0.0; }
#pragma unroll
    for (int u = 0; u < VPT; u++)
#pragma unroll
        for (int e = 0; e < NBX; e++) { bs[u][e] = 1.0; bl[u][e] = 0.0; }

    double red[4];
    {
        int nb = 0, bnd = 0;
#pragma unroll
        for (int u = 0; u < VPT; u++) { nb += __popc(bmask[u]); bnd |= (int) (bmask[u] & 3u); }
        red[0] = (double) (nrow + nb); red[1] = 0; red[2] = 0; red[3] = (double) bnd;
        block_reduce4<C>(red, s_red, red_phase);       // (barrier: s_c is complete after this)
    }
    const double n_rows = red[0];
    // CTA-uniform: no bound row left anywhere -> the row loops of the several-variables-per-thread instances start at
    // the velocity rows (with one variable per thread the masked rows cost nothing worth a branch)
    const int e_first = (VPT > 1 && !(red[3] > 0.0)) ? 2 : 0;

    // helpers -----------------------------------------------------------------------------------
    // directional change of the six box rows of variable slot u for a full-space direction dv
    auto box_dq = [&](const double* dv, double* dq, int u) {   // @phase row_helpers
        const int v = tid + u * NT;
        const double* d = dv + (v < NV ? v : 0);
        const double d0 = d[0];
        double dvv = 0.0, daa = 0.0;
        if (bmask[u] & 4u) dvv = d[1] - d0;
        if (bmask[u] & 16u) daa = d[2] - 2.0 * d[1] + d0;
        dq[0] = d0; dq[1] = -d0; dq[2] = -dvv; dq[3] = dvv; dq[4] = -daa; dq[5] = daa;
        if (C::COMM) {
            double dvc = 0.0;
            if (has_comm) dvc = dv[ca_idx] - (cb_idx >= 0 ? dv[cb_idx] : 0.0);
            dq[NBX - 2] = -dvc; dq[NBX - 1] = dvc;
        }
    };
    auto load_cp = [&](const double* v, double& x, double& y, double& z) {
        x = v[cp]; y = v[NCP + cp]; z = (D == 3) ? v[2 * NCP + cp] : 0.0;
    };
    auto store_slab = [&](const double* S, const double* T, bool withS) {
        if (cp_valid) {
            if (withS) {
#pragma unroll
                for (int e = 0; e < NS; e++) slabS[(grp * NCP + cp) * NS + e] = S[e];
            }
#pragma unroll
            for (int e = 0; e < D; e++) slabT[(grp * NCP + cp) * D + e] = T[e];
        }
    };
    auto accum = [&](double* S, double* T, const double* n, double W, double u, bool withS) {
        if (withS) {
            if (D == 3) {
                const double wx = W * n[0], wy = W * n[1], wz = W * n[2];
                S[0] += wx * n[0]; S[1] += wx * n[1]; S[2] += wx * n[2];
                S[3] += wy * n[1]; S[4] += wy * n[2]; S[5] += wz * n[2];
            } else {
                const double wx = W * n[0], wy = W * n[1];
                S[0] += wx * n[0]; S[1] += wx * n[1]; S[2] += wy * n[1];
            }
        }
        T[0] += u * n[0]; T[1] += u * n[1];
        if (D == 3) T[2] += u * n[2];
    };
    auto store_box = [&](const double* W, const double* u, bool withS, int slot) {
        // gradients of the rows: lb +e, ub -e, vel+ -(d), vel- +(d), acc+ -(d), acc- +(d)
        const int v = tid + slot * NT;
        if (v < NV) {
            if (withS) { s_wB[v] = W[0] + W[1]; s_wV[v] = W[2] + W[3]; s_wA[v] = W[4] + W[5]; }
            s_uB[v] = u[0] - u[1]; s_uV[v] = u[3] - u[2]; s_uA[v] = u[5] - u[4];
        }
        if (C::COMM && has_comm) {
            // rows hp - v >= 0 (gradient -g) and v - hm >= 0 (gradient +g), g = e_a - e_b
            if (withS) s_wC[tid] = W[NBX - 2] + W[NBX - 1];
            s_wC[C::PC + tid] = u[NBX - 1] - u[NBX - 2];
        }
    };

    // sum the group slabs, build the 6x6 blocks (withS) and the full-space rhs  -grad f + A^T u, then project.
    auto assemble = [&](bool withS) {   // @phase assemble
        cta_sync<C>();
        if (G > 1) {
            for (int e = tid; e < NCP * D; e += NT) {
                double a = slabT[e];
#pragma unroll
                for (int g = 1; g < G; g++) a += slabT[g * NCP * D + e];
                slabT[e] = a;
            }
            if (withS) {
                for (int e = tid; e < NCP * NS; e += NT) {
                    double a = slabS[e];
#pragma unroll
                    for (int g = 1; g < G; g++) a += slabS[g * NCP * NS + e];
                    slabS[e] = a;
                }
            }
            cta_sync<C>();
        }
#pragma unroll
        for (int u = 0; u < VPT; u++) {
            const int v = tid + u * NT;
            if (v >= NV) continue;
            const int k_v = v / NCP, cp_v = v % NCP, m_v = cp_v / 6, a = v % 6;
            const int v0 = k_v * NCP + m_v * 6;
            if (withS) {
                // row a of the 6x6 block of (dimension k_v, segment m_v): jerk Gram + terminal + bounds on the diagonal,
                // tri-diagonal velocity stencils (1,-1), penta-diagonal acceleration stencils (1,-2,1), LSC diagonal block
                const double wv0 = a <= 4 ? s_wV[v0 + a] : 0.0, wvm = a >= 1 ? s_wV[v0 + a - 1] : 0.0;
                const double wa0 = a <= 3 ? s_wA[v0 + a] : 0.0, wa1 = (a >= 1 && a <= 4) ? s_wA[v0 + a - 1] : 0.0,
                             wa2 = a >= 2 ? s_wA[v0 + a - 2] : 0.0;
                const int dd = (D == 3) ? symidx3(k_v, k_v) : symidx2(k_v, k_v);
                double row[6];
#pragma unroll
                for (int b = 0; b < 6; b++) row[b] = sQ2[a * 6 + b];
                // diagonal
                double dg = s_wB[v] + slabS[cp_v * NS + dd] + wv0 + wvm + wa0 + 4.0 * wa1 + wa2;
                if (a == 5) dg += s_termw[m_v];
                // neighbours: (a, a+1): -wV[a] - 2 wA[a] - 2 wA[a-1];  (a, a+2): wA[a];  mirrored for a-1, a-2
                const double up1 = -wv0 - 2.0 * wa0 - 2.0 * wa1, dn1 = -wvm - 2.0 * wa1 - 2.0 * wa2;
#pragma unroll
                for (int b = 0; b < 6; b++) {
                    double add = 0.0;
                    if (b == a) add = dg;
                    else if (b == a + 1) add = up1;
                    else if (b == a - 1) add = dn1;
                    else if (b == a + 2) add = wa0;
                    else if (b == a - 2) add = wa2;
                    s_blk[(k_v * M + m_v) * 36 + a * 6 + b] = row[b] + add;
                }
            }
            double gr = 0.0;
#pragma unroll
            for (int b = 0; b < 6; b++) gr += sQ2[a * 6 + b] * s_c[v0 + b];
            if (a == 5) gr += s_termw[m_v] * (s_c[v0 + 5] - s_goal[k_v]);
            double au = s_uB[v] + slabT[cp_v * D + k_v];
            if (a >= 1) au += s_uV[v0 + a - 1];
            if (a <= 4) au -= s_uV[v0 + a];
            // acceleration stencils (1, -2, 1) anchored at i = a - 2, a - 1, a (those with 0 <= i <= 3): branch free in the
            // several-variables-per-thread instances, a short loop otherwise (each measured faster where it is used)
            if (VPT > 1) {
                au += (a >= 2 ? s_uA[v0 + a - 2] : 0.0) - 2.0 * ((a >= 1 && a <= 4) ? s_uA[v0 + a - 1] : 0.0) + (a <= 3 ? s_uA[v0 + a] : 0.0);
            } else {
                for (int i = (a - 2 > 0 ? a - 2 : 0); i <= (a < 3 ? a : 3); i++) au += s_uA[v0 + i] * ((a - i == 1) ? -2.0 : 1.0);
            }
            s_rfull[v] = au - gr;
        }
        cta_sync<C>();
        if (withS) {
            const ProjTerm* tab = (p.klass_mode == 1 ? p.proj_light : p.proj) + tid;
            const int len = p.klass_mode == 1 ? p.proj_len_light : p.proj_len;
            double val = 0.0;
#pragma unroll 4
            for (int i = 0; i < len; i++) {
                const ProjTerm t = tab[i * NT];
                val = fma(t.coef, s_blk[t.src], val);
                if (t.dest >= 0) {
                    const int d = t.dest & PROJ_MASK;
                    if (t.dest & PROJ_ONE) val = 1.0;                           // identity padding rows
                    s_A[d] = val;
                    if (t.dest & PROJ_DIAG) s_diag0[d / (C::LD + 1)] = val;
                    val = 0.0;
                }
            }
        }
        for (int r0 = tid; r0 < C::NRP; r0 += NT) {
            double r = r0 < NR ? reduce_from_full<C>(s_rfull, r0) : 0.0;
            if (C::COMM && r0 < NR) {
                // end-point variables also collect A^T u of the comm pairs they appear in
                int st = -1, k = 0;
                if (C::TERM && r0 >= (M - 1) * C::NZS) { st = M - 1; k = r0 - (M - 1) * C::NZS; }
                else if (r0 % 3 == 2) { st = r0 / C::NZS; k = (r0 % C::NZS) / 3; }
                if (st >= 0) {
                    const double* uc = s_wC + C::PC + k * C::PP;
                    r += uc[st];                                                      // box pair of E_st
                    for (int b = 0; b < st; b++) r += uc[M + st * (st - 1) / 2 + b];           // pairs (st, b): +g
                    for (int a = st + 1; a < M; a++) r -= uc[M + a * (a - 1) / 2 + st];        // pairs (a, st): -g
                }
            }
            s_rhs[r0] = r;
        }
        cta_sync<C>();
    };

    int bad_piv = 0;
    auto factor_solve = [&](bool factor) {   // @phase factor_solve_call
        if (C::COMM) {
            if (factor) bad_piv += chol_dense_cta<C>(s_A, s_diag0, s_invd, sm + C::O_SBUF);
            if (tid < 32) solve_dense_warp<C>(s_A, s_invd, s_rhs);
        } else if (tid < 32) {
            if (factor) bad_piv += chol_banded<C>(s_A, s_diag0, s_invd, sm + C::O_SBUF, pr_i, pr_k);
            chol_solve<C>(s_A, s_invd, s_rhs);
        }
        cta_sync<C>();
    };

    // ---------------------------------------------------------------- initial point   // @phase init_point
    // q of every row at the starting point is kept in ls[] / bs[] until s and lam are set.
    auto box_q = [&](double* q, int u) {
        const int v = tid + u * NT;
        if (v < NV) {
            const int k_v = v / NCP, m_v = (v % NCP) / 6;
            const double* cc = s_c + v;
            const double c0 = cc[0];
            double dv = 0.0, da = 0.0;
            if (bmask[u] & 4u) dv = cc[1] - c0;
            if (bmask[u] & 16u) da = cc[2] - 2.0 * cc[1] + c0;
            q[0] = c0 - s_lb[k_v * M + m_v]; q[1] = s_ub[k_v * M + m_v] - c0;
            q[2] = s_vlim[k_v] - dv; q[3] = s_vlim[k_v] + dv; q[4] = s_alim[k_v] - da; q[5] = s_alim[k_v] + da;
        }
        if (C::COMM) {
            const double vv = has_comm ? s_c[ca_idx] - (cb_idx >= 0 ? s_c[cb_idx] : 0.0) : 0.0;
            q[NBX - 2] = c_hp - vv; q[NBX - 1] = vv - c_hm;
        }
    };
    auto rows_q = [&]() {
        double cx, cy, cz;
        if (cp_valid) load_cp(s_c, cx, cy, cz); else { cx = cy = cz = 0; }
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            if (j >= nrow) break;
            const int oi = grp + G * j;
            const double* n = s_nrm + (oi * M + m_cp) * 3;
            double q = n[0] * (cx + s_org[0]) + n[1] * (cy + s_org[1]) - p.rhs[((size_t) (obs0 + s_act[oi]) * M + m_cp) * 6 + i_cp];
            if (D == 3) q += n[2] * (cz + s_org[2]);
            if (n[0] == 0.0 && n[1] == 0.0 && n[2] == 0.0) q = 1.0;
            ls[j] = q;
        }
#pragma unroll
        for (int u = 0; u < VPT; u++) box_q(bs[u], u);
    };
    rows_q();
    if (warm) {
        // an initial_traj that violates rows by more than warm_reject is not used (cold start instead)
        double qmin = INFINITY;
#pragma unroll
        for (int j = 0; j < KPT; j++) if (j < nrow) qmin = fmin(qmin, ls[j]);
#pragma unroll
        for (int u = 0; u < VPT; u++)
#pragma unroll
            for (int e = 0; e < NBX; e++) if (bmask[u] >> e & 1u) qmin = fmin(qmin, bs[u][e]);
        red[0] = 0; red[1] = 0; red[2] = qmin; red[3] = 0;
        block_reduce4<C>(red, s_red, red_phase);
        if (red[2] < -p.warm_reject) {
            warm = false;
            cta_sync<C>();
            set_start(false);
            rows_q();
        }
    }
    if (!warm) {
        // cold start: minimise f(y) + 1/2 |q(y)|^2 (W = 1, u = -q at the hover point); then s = q, lam = -q,
        // both shifted into the positive orthant (Mehrotra / CVXOPT start)
        double S[6] = {0, 0, 0, 0, 0, 0}, T[3] = {0, 0, 0};
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            if (j >= nrow) break;
            const double* n = s_nrm + ((grp + G * j) * M + m_cp) * 3;
            accum(S, T, n, 1.0, -ls[j], true);
        }
        store_slab(S, T, true);
#pragma unroll
        for (int u = 0; u < VPT; u++) {
            double W[NBX], uu[NBX];
#pragma unroll
            for (int e = 0; e < NBX; e++) { const bool on = bmask[u] >> e & 1u; W[e] = on ? 1.0 : 0.0; uu[e] = on ? -bs[u][e] : 0.0; }
            store_box(W, uu, true, u);
        }
    }
    if (!warm) {
        assemble(true);
        factor_solve(true);
        for (int r = tid; r < NR; r += NT) { s_dy[r] = s_rhs[r]; s_y[r] += s_rhs[r]; }
        cta_sync<C>();
        expand(s_dy, nullptr, s_dc);
        expand(s_y, s_x0, s_c);
        cta_sync<C>();
    }
    double rp;      // the common primal residual s - q
    {
        double dx, dy, dz;
        if (!warm && cp_valid) load_cp(s_dc, dx, dy, dz); else { dx = dy = dz = 0; }
        double qmin = INFINITY, qmax = -INFINITY;
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            if (j >= nrow) break;
            const double* n = s_nrm + ((grp + G * j) * M + m_cp) * 3;
            double dq = n[0] * dx + n[1] * dy;
            if (D == 3) dq += n[2] * dz;
            ls[j] += dq;
            qmin = fmin(qmin, ls[j]); qmax = fmax(qmax, ls[j]);
        }
#pragma unroll
        for (int u = 0; u < VPT; u++) {
            double dq[NBX] = {};
            if (!warm) box_dq(s_dc, dq, u);
#pragma unroll
            for (int e = 0; e < NBX; e++) if (bmask[u] >> e & 1u) { bs[u][e] += dq[e]; qmin = fmin(qmin, bs[u][e]); qmax = fmax(qmax, bs[u][e]); }
        }
        red[0] = 0; red[1] = 0; red[2] = qmin; red[3] = qmax;
        block_reduce4<C>(red, s_red, red_phase);
        if (warm) {
            // warm start: s = q + shift (shift = 0 when initial_traj is strictly inside by warm_delta), s lam = mu0;
            // the multiplier scale follows the terminal-cost gradient (~ distance to the goal)
            const double shift_s = fmax(0.0, p.warm_delta - red[2]);
            const double mu0 = p.mu0 * fmax(1.0, 0.1 * s_termw[M]);
#pragma unroll
            for (int j = 0; j < KPT; j++) if (j < nrow) { ls[j] += shift_s; ll[j] = mu0 / ls[j]; }
#pragma unroll
            for (int u = 0; u < VPT; u++)
#pragma unroll
                for (int e = 0; e < NBX; e++) if (bmask[u] >> e & 1u) { bs[u][e] += shift_s; bl[u][e] = mu0 / bs[u][e]; }
            rp = shift_s;
        } else {
            const double shift_s = (red[2] <= 0.0) ? 1.0 - red[2] : 0.0;        // alpha_p = -min(s) >= 0  -> s += 1 + alpha_p
            const double shift_l = (red[3] >= 0.0) ? 1.0 + red[3] : 0.0;        // lam = -s; alpha_d = max(s) >= 0 -> lam += 1 + alpha_d
#pragma unroll
            for (int j = 0; j < KPT; j++) if (j < nrow) { ll[j] = -ls[j] + shift_l; ls[j] += shift_s; }
#pragma unroll
            for (int u = 0; u < VPT; u++)
#pragma unroll
                for (int e = 0; e < NBX; e++) if (bmask[u] >> e & 1u) { bl[u][e] = -bs[u][e] + shift_l; bs[u][e] += shift_s; }
            rp = shift_s;
        }
    }

    // ---------------------------------------------------------------- main loop   // @phase sweepA
    int status = ST_MAX_ITER, it = 0;
    double mu = 0.0, sigmu = 0.0, alpha = 0.0, mu_first = 0.0;
    double res_scale = 1.0;       // prod (1 - alpha): what is left of the initial primal and dual residuals (both linear)
    bool have_step = false;       // a (dca, dc, sigmu, alpha) step is pending and is applied by the next sweep A
    for (it = 0; it <= p.max_iter; it++) {
        // ---- sweep A: apply the pending step, then predictor weights of the new point
        {
            double S[6] = {0, 0, 0, 0, 0, 0}, T[3] = {0, 0, 0};
            double ax, ay, az, dx, dy, dz;
            if (have_step && cp_valid) { load_cp(s_dca, ax, ay, az); load_cp(s_dc, dx, dy, dz); }
            else { ax = ay = az = dx = dy = dz = 0; }
            const double rp_new = have_step ? (1.0 - alpha) * rp : rp;
            if (have_step) res_scale *= (1.0 - alpha);
            double sl = 0.0;
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                if (j >= nrow) break;
                const double* n = s_nrm + ((grp + G * j) * M + m_cp) * 3;
                if (have_step) {
                    double dqa = n[0] * ax + n[1] * ay, dq = n[0] * dx + n[1] * dy;
                    if (D == 3) { dqa += n[2] * az; dq += n[2] * dz; }
                    const double rs = fast_rcp(ls[j]), W = ll[j] * rs;
                    const double dsa = dqa - rp, dla = -ll[j] - W * dsa;
                    const double rc = ls[j] * ll[j] + dsa * dla - sigmu;
                    const double ds = dq - rp, dl = -(rc + ll[j] * ds) * rs;
                    ls[j] += alpha * ds; ll[j] += alpha * dl;
                }
                const double W = ll[j] * fast_rcp(ls[j]);
                sl += ls[j] * ll[j];
                accum(S, T, n, W, W * rp_new, true);
            }
            store_slab(S, T, true);
#pragma unroll
            for (int u = 0; u < VPT; u++) {
                double W[NBX], uu[NBX], dqa[NBX], dq[NBX];
                if (have_step) { box_dq(s_dca, dqa, u); box_dq(s_dc, dq, u); }
#pragma unroll
                for (int e = 0; e < NBX; e++) {
                    W[e] = 0.0; uu[e] = 0.0;
                    if (e < e_first || !(bmask[u] >> e & 1u)) continue;
                    if (have_step) {
                        const double rs = fast_rcp(bs[u][e]), Wo = bl[u][e] * rs;
                        const double dsa = dqa[e] - rp, dla = -bl[u][e] - Wo * dsa;
                        const double rc = bs[u][e] * bl[u][e] + dsa * dla - sigmu;
                        const double ds = dq[e] - rp, dl = -(rc + bl[u][e] * ds) * rs;
                        bs[u][e] += alpha * ds; bl[u][e] += alpha * dl;
                    }
                    W[e] = bl[u][e] * fast_rcp(bs[u][e]); uu[e] = W[e] * rp_new;
                    sl += bs[u][e] * bl[u][e];
                }
                store_box(W, uu, true, u);
            }
            rp = rp_new;
            red[0] = sl; red[1] = 0; red[2] = 0; red[3] = 0;
            block_reduce4<C>(red, s_red, red_phase);
            mu = red[0] / n_rows;
        }
        if (!(mu == mu)) { status = ST_NUMERICAL; break; }
        // Stop: complementarity and primal residual small, and the initial *dual* residual -- which every step scales by
        // (1 - alpha) like the primal one -- reduced by at least 1e-6 (a strictly interior warm start has rp = 0 from the
        // first iteration, so the primal test alone would accept a run of short steps at a non-stationary point)
        if (mu < p.mu_tol && fabs(rp) < p.rp_tol && res_scale < 1e-6) { status = ST_OK; break; }
        // an infeasible model drives the multipliers (and with them mu) to infinity: stop long before the overflow
        if (it == 0) mu_first = mu;
        if (mu > 1e12 * fmax(mu_first, 1.0)) { status = ST_INFEASIBLE; break; }
        if (it == p.max_iter) break;

        // ---- predictor   // @phase predictor_glue
        assemble(true);
        factor_solve(true);
        for (int r = tid; r < NR; r += NT) s_dy[r] = s_rhs[r];
        cta_sync<C>();
        expand(s_dy, nullptr, s_dca);
        cta_sync<C>();
        // ---- sweep B: affine step length and centering parameter   // @phase sweepB
        {
            double ax, ay, az;
            if (cp_valid) load_cp(s_dca, ax, ay, az); else { ax = ay = az = 0; }
            double tmax = 0.0;                                   // max over rows of -dv / v (>= 1/2 for every row pair)
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                if (j >= nrow) break;
                const double* n = s_nrm + ((grp + G * j) * M + m_cp) * 3;
                double dqa = n[0] * ax + n[1] * ay;
                if (D == 3) dqa += n[2] * az;
                const double rs = fast_rcp(ls[j]), W = ll[j] * rs;
                const double dsa = dqa - rp, dla = -ll[j] - W * dsa;
                // inverse step to the boundary of both rows without a division: -dsa / s and -dla / lam = 1 + dsa / s
                const double z = dsa * rs;
                tmax = fmax(tmax, fmax(-z, 1.0 + z));
                s1 += ls[j] * dla + ll[j] * dsa; s2 += dsa * dla;
            }
#pragma unroll
            for (int u = 0; u < VPT; u++) {
                double dqa[NBX];
                box_dq(s_dca, dqa, u);
#pragma unroll
                for (int e = 0; e < NBX; e++) {
                    if (e < e_first || !(bmask[u] >> e & 1u)) continue;
                    const double rs = fast_rcp(bs[u][e]), W = bl[u][e] * rs;
                    const double dsa = dqa[e] - rp, dla = -bl[u][e] - W * dsa;
                    const double z = dsa * rs;
                    tmax = fmax(tmax, fmax(-z, 1.0 + z));
                    s1 += bs[u][e] * dla + bl[u][e] * dsa; s2 += dsa * dla;
                }
            }
            red[0] = s1; red[1] = s2; red[2] = 0; red[3] = tmax;
            block_reduce4<C>(red, s_red, red_phase);
            const double a = red[3] > 1.0 ? 1.0 / red[3] : 1.0;
            const double mu_aff = (mu * n_rows + a * red[0] + a * a * red[1]) / n_rows;
            double sg = mu_aff / mu;
            sg = sg * sg * sg;
            if (!(sg >= 0.0)) sg = 0.0;
            if (sg > 1.0) sg = 1.0;
            sigmu = sg * mu;
        }
        // ---- sweep C: corrector right-hand side   // @phase sweepC
        {
            double S[6] = {0, 0, 0, 0, 0, 0}, T[3] = {0, 0, 0};
            double ax, ay, az;
            if (cp_valid) load_cp(s_dca, ax, ay, az); else { ax = ay = az = 0; }
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                if (j >= nrow) break;
                const double* n = s_nrm + ((grp + G * j) * M + m_cp) * 3;
                double dqa = n[0] * ax + n[1] * ay;
                if (D == 3) dqa += n[2] * az;
                const double rs = fast_rcp(ls[j]), W = ll[j] * rs;
                const double dsa = dqa - rp, dla = -ll[j] - W * dsa;
                const double rc = ls[j] * ll[j] + dsa * dla - sigmu;
                accum(S, T, n, 0.0, ll[j] + (ll[j] * rp - rc) * rs, false);
            }
            store_slab(S, T, false);
#pragma unroll
            for (int u = 0; u < VPT; u++) {
                double W[NBX], uu[NBX], dqa[NBX];
                box_dq(s_dca, dqa, u);
#pragma unroll
                for (int e = 0; e < NBX; e++) {
                    W[e] = 0.0; uu[e] = 0.0;
                    if (e < e_first || !(bmask[u] >> e & 1u)) continue;
                    const double rs = fast_rcp(bs[u][e]), Wo = bl[u][e] * rs;
                    const double dsa = dqa[e] - rp, dla = -bl[u][e] - Wo * dsa;
                    const double rc = bs[u][e] * bl[u][e] + dsa * dla - sigmu;
                    uu[e] = bl[u][e] + (bl[u][e] * rp - rc) * rs;
                }
                store_box(W, uu, false, u);
            }
        }
        assemble(false);
        factor_solve(false);
        for (int r = tid; r < NR; r += NT) s_dy[r] = s_rhs[r];
        cta_sync<C>();
        expand(s_dy, nullptr, s_dc);
        cta_sync<C>();
        // ---- sweep D: step length of the combined direction   // @phase sweepD
        {
            double ax, ay, az, dx, dy, dz;
            if (cp_valid) { load_cp(s_dca, ax, ay, az); load_cp(s_dc, dx, dy, dz); } else { ax = ay = az = dx = dy = dz = 0; }
            MinRatio mr; mr.init();                              // multiplier rows (float quotient)
            double tmax = 0.0;                                   // slack rows: max of -ds / s through the known 1 / s
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                if (j >= nrow) break;
                const double* n = s_nrm + ((grp + G * j) * M + m_cp) * 3;
                double dqa = n[0] * ax + n[1] * ay, dq = n[0] * dx + n[1] * dy;
                if (D == 3) { dqa += n[2] * az; dq += n[2] * dz; }
                const double rs = fast_rcp(ls[j]), W = ll[j] * rs;
                const double dsa = dqa - rp, dla = -ll[j] - W * dsa;
                const double rc = ls[j] * ll[j] + dsa * dla - sigmu;
                const double ds = dq - rp, dl = -(rc + ll[j] * ds) * rs;
                tmax = fmax(tmax, -ds * rs); mr.add(ll[j], dl);
            }
#pragma unroll
            for (int u = 0; u < VPT; u++) {
                double dqa[NBX], dq[NBX];
                box_dq(s_dca, dqa, u); box_dq(s_dc, dq, u);
#pragma unroll
                for (int e = 0; e < NBX; e++) {
                    if (e < e_first || !(bmask[u] >> e & 1u)) continue;
                    const double rs = fast_rcp(bs[u][e]), Wo = bl[u][e] * rs;
                    const double dsa = dqa[e] - rp, dla = -bl[u][e] - Wo * dsa;
                    const double rc = bs[u][e] * bl[u][e] + dsa * dla - sigmu;
                    const double ds = dq[e] - rp, dl = -(rc + bl[u][e] * ds) * rs;
                    tmax = fmax(tmax, -ds * rs); mr.add(bl[u][e], dl);
                }
            }
            red[0] = 0; red[1] = 0; red[2] = mr.value(); red[3] = tmax;
            block_reduce4<C>(red, s_red, red_phase);
            // (the ratios carry ~1e-7 relative error: a full step needs a margin above 1, otherwise stay 1% inside)
            const double ratio = fmin(red[2], red[3] > 0.0 ? 1.0 / red[3] : 2.0);
            alpha = ratio >= 1.0001 ? 1.0 : 0.99 * fmin(ratio, 1.0);
            have_step = true;
        }
        // the reduced iterate moves now; the rows follow in the next sweep A (s_dca / s_dc stay valid until then)
        for (int r = tid; r < NR; r += NT) s_y[r] += alpha * s_dy[r];
        cta_sync<C>();
        expand(s_y, s_x0, s_c);
        // (assemble() starts with a barrier before anything reads s_c)
    }

    // ---------------------------------------------------------------- outputs   // @phase outputs
    // a numerical breakdown (NaN in the complementarity measure; seen on infeasible models whose pivots were all
    // guarded) must not leak: the returned point is then the (finite) starting point, its multipliers zero
    if (status == ST_NUMERICAL) {
        cta_sync<C>();
        set_start(p.warm_traj != nullptr);
#pragma unroll
        for (int j = 0; j < KPT; j++) { ls[j] = 1.0; ll[j] = 0.0; }
#pragma unroll
        for (int u = 0; u < VPT; u++)
#pragma unroll
            for (int e = 0; e < NBX; e++) { bs[u][e] = 1.0; bl[u][e] = 0.0; }
        mu = -1.0;
    }
    // true primal residual of the returned point, recomputed from the row constants
    double rp_true = 0.0;
    cta_sync<C>();
    {
        double cx, cy, cz;
        if (cp_valid) load_cp(s_c, cx, cy, cz); else { cx = cy = cz = 0; }
        double S[6] = {0, 0, 0, 0, 0, 0}, T[3] = {0, 0, 0};
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            if (j >= nrow) break;
            const int oi = grp + G * j;
            const double* n = s_nrm + (oi * M + m_cp) * 3;
            double q = n[0] * (cx + s_org[0]) + n[1] * (cy + s_org[1]) - p.rhs[((size_t) (obs0 + s_act[oi]) * M + m_cp) * 6 + i_cp];
            if (D == 3) q += n[2] * (cz + s_org[2]);
            if (n[0] == 0.0 && n[1] == 0.0 && n[2] == 0.0) q = 1.0;
            rp_true = fmax(rp_true, fabs(ls[j] - q));
            accum(S, T, n, 0.0, ll[j], false);
        }
        store_slab(S, T, false);
#pragma unroll
        for (int u = 0; u < VPT; u++) {
            double q[NBX] = {};
            box_q(q, u);
            double W[NBX] = {}, uu[NBX];
#pragma unroll
            for (int e = 0; e < NBX; e++) {
                const bool on = bmask[u] >> e & 1u;
                uu[e] = on ? bl[u][e] : 0.0;
                if (on) rp_true = fmax(rp_true, fabs(bs[u][e] - q[e]));
            }
            store_box(W, uu, false, u);
        }
        // stationarity || Z^T (grad f - A^T lam) ||_inf through the same projection (u = lam)
        assemble(false);
    }
    double rd_inf = 0.0, cost = 0.0;
    for (int r = tid; r < NR; r += NT) rd_inf = fmax(rd_inf, fabs(s_rhs[r]));
#pragma unroll
    for (int u = 0; u < VPT; u++) {
        const int v = tid + u * NT;
        if (v >= NV) continue;
        // objective x'Px + q'x + c0 with P = w_c Q (no 1/2, traj_optimizer.cpp:294) + terminal terms (:301-315)
        const int k_v = v / NCP, m_v = (v % NCP) / 6, i_v = v % 6;
        const int v0 = k_v * NCP + m_v * 6;
        double qc = 0.0;
#pragma unroll
        for (int b = 0; b < 6; b++) qc += sQ2[i_v * 6 + b] * s_c[v0 + b];
        cost += 0.5 * s_c[v] * qc;
        if (i_v == 5) { const double e = s_c[v] - s_goal[k_v]; cost += 0.5 * s_termw[m_v] * e * e; }
        p.ctrl_out[(size_t) agent * NV + v] = s_c[v] + s_org[k_v];
    }
    red[0] = cost; red[1] = 0; red[2] = -rp_true; red[3] = rd_inf;
    block_reduce4<C>(red, s_red, red_phase);
    rp_true = -red[2];
    if (status == ST_MAX_ITER && rp_true > 1e-6) status = ST_INFEASIBLE;
    if (tid == 0) {
        p.cost_out[agent] = red[0];
        p.status_out[agent] = status;
        if (p.iters_out) p.iters_out[agent] = it;
        if (p.kkt_out) {
            p.kkt_out[agent * 4 + 0] = red[3]; p.kkt_out[agent * 4 + 1] = rp_true;
            p.kkt_out[agent * 4 + 2] = (double) bad_piv; p.kkt_out[agent * 4 + 3] = mu;
        }
    }
    if (p.dual_out) {
        double* du = p.dual_out + (size_t) agent * p.dual_stride;
        for (int e = tid; e < C::KRAW * M * 6; e += NT) du[e] = 0.0;       // dropped / skipped rows: multiplier 0
        cta_sync<C>();
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            if (j >= nrow) break;
            const int slot = grp + G * j;
            const double* n = s_nrm + (slot * M + m_cp) * 3;
            if (!(n[0] == 0.0 && n[1] == 0.0 && n[2] == 0.0)) du[(s_act[slot] * M + m_cp) * 6 + i_cp] = ll[j];
        }
        // back to the reference's row scaling: vel rows carry 5/dt, acc rows 20/dt^2
        const double sv = p.dt / 5.0, sa = p.dt * p.dt / 20.0;
#pragma unroll
        for (int u = 0; u < VPT; u++) {
            const int v = tid + u * NT;
            if (v >= NV) continue;
            double* db = du + C::KRAW * M * 6 + v * 6;
            db[0] = (bmask[u] & 1u) ? bl[u][0] : 0.0; db[1] = (bmask[u] & 2u) ? bl[u][1] : 0.0;
            db[2] = (bmask[u] & 4u) ? bl[u][2] * sv : 0.0; db[3] = (bmask[u] & 8u) ? bl[u][3] * sv : 0.0;
            db[4] = (bmask[u] & 16u) ? bl[u][4] * sa : 0.0; db[5] = (bmask[u] & 32u) ? bl[u][5] * sa : 0.0;
        }
        if (C::COMM && has_comm) { double* dc = du + C::KRAW * M * 6 + NV * 6 + tid * 2; dc[0] = bl[0][NBX - 2]; dc[1] = bl[0][NBX - 1]; }
    }
}

}  // namespace lscqp
