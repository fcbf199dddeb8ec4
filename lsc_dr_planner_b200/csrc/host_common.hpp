// host_common.hpp -- host-side constants shared by the CUDA library and the CPU emulation harness.
#pragma once
#include <cmath>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <vector>
#include "../../include/lscqp.h"
#include "pdip_kernel.cuh"
#include "das_kernel.cuh"

namespace lscqp {

// Integrated squared phi-th derivative of a degree-n Bernstein segment of duration dt, as a
// quadratic form in the control points.  Same matrix as TrajOptimizer::buildQBase
// (src/traj_optimizer.cpp:163-178, with include/polynomial.hpp:90-100, 281-294):
//   Q = B Z B^T dt^(1-2 phi),  Z(i,j) = i!/(i-phi)! j!/(j-phi)! / (i+j-2phi+1),
//   B(i,j) = C(n,i) C(n-i,n-j) (-1)^(j-i)  (Bernstein -> monomial).
inline void jerk_gram(int n, int phi, double dt, double* Q /* (n+1)^2 */) {
    const int N = n + 1;
    auto binom = [](int a, int b) -> double {
        if (b < 0 || b > a) return 0.0;
        double r = 1.0;
        for (int i = 1; i <= b; i++) r = r * (a - b + i) / i;
        return std::round(r);
    };
    auto falling = [](int a, int k) -> double {
        if (a < k) return 0.0;
        double r = 1.0;
        for (int i = 0; i < k; i++) r *= (a - i);
        return r;
    };
    double B[16][16], Z[16][16], T[16][16];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            B[i][j] = (j >= i) ? binom(n, i) * binom(n - i, n - j) * (((j - i) & 1) ? -1.0 : 1.0) : 0.0;
            const int e = i + j - 2 * phi + 1;
            Z[i][j] = (e > 0) ? falling(i, phi) * falling(j, phi) / e : 0.0;
        }
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            double s = 0;
            for (int l = 0; l < N; l++) s += B[i][l] * Z[l][j];
            T[i][j] = s;
        }
    const double scale = std::pow(dt, 1 - 2 * phi);
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            double s = 0;
            for (int l = 0; l < N; l++) s += T[i][l] * B[j][l];
            Q[i * N + j] = s * scale;
        }
}

inline int validate_config(const lscqp_config& c) {
    if (c.n != 5 || c.phi != 3) return LSCQP_E_INVALID;          // traj_optimizer.cpp:198-201
    if (!(c.M == 5 || c.M == 10)) return LSCQP_E_INVALID;
    if (!(c.dim == 2 || c.dim == 3)) return LSCQP_E_INVALID;
    if (!(c.planner_mode == LSCQP_MODE_DLSC || c.planner_mode == LSCQP_MODE_LSC || c.planner_mode == LSCQP_MODE_BVC ||
          c.planner_mode == LSCQP_MODE_RECIPROCALRSFC))
        return LSCQP_E_INVALID;
    if (c.max_obs < 0 || c.max_obs > 40) return LSCQP_E_INVALID;
    if (!(c.dt > 0) || !(c.w_control > 0) || !(c.w_terminal >= 0)) return LSCQP_E_INVALID;
    return 0;
}

inline void fill_solve_params(const lscqp_config& c, SolveParams& p) {
    std::memset(&p, 0, sizeof(p));
    p.max_iter = c.max_iter > 0 ? c.max_iter : 60;
    p.mu_tol = c.tol > 0 ? c.tol : 1e-11;
    p.rp_tol = 1e-9;
    p.mu0 = 0.1 * (c.w_terminal > 0 ? c.w_terminal : 1.0);   // warm start: initial complementarity target
    p.warm_delta = 1e-3;                                      // warm start: minimum initial slack
    p.warm_reject = 0.02;                                     // warm start: largest row violation still accepted
    // (mu0 and warm_delta were swept on the 4096-agent workload: 0.003..1.0 x 1e-3..1e-2; this pair gives the fewest
    //  iterations -- 9.17 on average, 10.4 at mu0 = 0.01, 10.8 at warm_delta = 1e-2)
    p.dt = c.dt; p.w_t = c.w_terminal; p.w_c = c.w_control;
    for (int k = 0; k < 3; k++) { p.world_min[k] = c.world_min[k]; p.world_max[k] = c.world_max[k]; }
    p.use_sfc = c.use_sfc;
    p.presolve = c.presolve & 1;
    p.max_obs = c.max_obs;
    p.rsfc = c.planner_mode == LSCQP_MODE_RECIPROCALRSFC;
    p.klass = nullptr; p.klass_mode = 0;
    p.comm_range = c.comm_range;
    double Q[36];
    jerk_gram(c.n, c.phi, c.dt, Q);
    for (int e = 0; e < 36; e++) p.Q2[e] = 2.0 * c.w_control * Q[e];
}

// Projection table of one kernel instance: every entry (r1 >= r2, inside the stored band) of the reduced matrix
// Z^T Phi Z as a linear combination of the full-space sources the kernel keeps in shared memory
// (6x6 blocks per (dimension, segment), then the per-control-point cross-dimension blocks).  Z is the continuity
// map of pdip_kernel.cuh: reduced variable (stage s, dim k, j) is control point (s, 3+j) and, through
// T = [[0,0,1],[0,-1,2],[1,-4,4]], the first three control points of segment s+1; the collapsed terminal variable
// is the sum of the last three control points.
// The table is laid out as one term stream per thread (ProjTerm, pdip_kernel.cuh): the entries are dealt to the
// threads so that every thread walks the same number of terms (longest-processing-time first), term i of thread t
// sits at [i * NT + t] (coalesced, address independent of the data, so the loads pipeline).
struct ProjTable {
    std::vector<ProjTerm> term;   // [len][nt]
    int len = 0, nt = 0;
};

template <class C>
inline ProjTable build_projection(int nt = C::NT) {
    constexpr int M = C::M, D = C::D, NR = C::NR, NS = C::NS;
    static const double T[3][3] = {{0, 0, 1}, {0, -1, 2}, {1, -4, 4}};
    auto symidx = [](int a, int b) { if (a > b) std::swap(a, b); return D == 3 ? (a == 0 ? b : (a == 1 ? 2 + b : 5)) : a + b; };
    auto blk = [](int k, int m, int a, int b) { return (k * M + m) * 36 + a * 6 + b; };
    auto sidx = [&](int cp, int k1, int k2) { return D * M * 36 + cp * NS + symidx(k1, k2); };
    struct Var { int s, k, j; };
    auto decode = [](int r) {
        Var v;
        if (C::TERM && r >= (M - 1) * C::NZS) { v.s = M - 1; v.k = r - (M - 1) * C::NZS; v.j = -1; }
        else { v.s = r / C::NZS; v.k = (r % C::NZS) / 3; v.j = r % 3; }
        return v;
    };
    struct Ent { int dest, diag, count; std::vector<std::pair<double, int>> t; };
    std::vector<Ent> ents;
    for (int r1 = 0; r1 < C::NRP; r1++)
        for (int off = 0; off <= C::BWS; off++) {
            const int r2 = r1 - off;
            if (r2 < 0) continue;
            Ent e; e.dest = r1 * C::LD + r2; e.diag = (r1 == r2) ? r1 : -1; e.count = 0;
            if (r1 >= NR) { e.count = (r1 == r2) ? -1 : 0; ents.push_back(e); continue; }
            const Var a = decode(r1), b = decode(r2);
            auto add = [&](double c, int idx) { if (c != 0.0) e.t.push_back({c, idx}); };
            if (a.s == b.s) {
                const int st = a.s;
                if (a.k == b.k) {
                    if (a.j < 0) { for (int x = 3; x < 6; x++) for (int y = 3; y < 6; y++) add(1.0, blk(a.k, st, x, y)); }
                    else {
                        add(1.0, blk(a.k, st, 3 + a.j, 3 + b.j));
                        if (st + 1 < M) for (int x = 0; x < 3; x++) for (int y = 0; y < 3; y++) add(T[x][a.j] * T[y][b.j], blk(a.k, st + 1, x, y));
                    }
                } else {
                    if (a.j < 0) { for (int x = 3; x < 6; x++) add(1.0, sidx(st * 6 + x, a.k, b.k)); }
                    else {
                        if (a.j == b.j) add(1.0, sidx(st * 6 + 3 + a.j, a.k, b.k));
                        if (st + 1 < M) for (int x = 0; x < 3; x++) add(T[x][a.j] * T[x][b.j], sidx((st + 1) * 6 + x, a.k, b.k));
                    }
                }
            } else if (a.s == b.s + 1 && a.k == b.k) {
                if (a.j < 0) { for (int x = 3; x < 6; x++) for (int y = 0; y < 3; y++) add(T[y][b.j], blk(a.k, a.s, x, y)); }
                else for (int y = 0; y < 3; y++) add(T[y][b.j], blk(a.k, a.s, 3 + a.j, y));
            }
            if (C::COMM) {
                // communication-range pairs act on the end-point variables E(k, a) directly
                auto endpoint = [](int k, int a) { return (C::TERM && a == M - 1) ? (M - 1) * C::NZS + k : a * C::NZS + k * 3 + 2; };
                const int src0 = C::O_WC - C::O_BLK;
                for (int k = 0; k < D; k++)
                    for (int x = 0; x < M; x++) {
                        const int Ea = endpoint(k, x);
                        if (r1 == Ea && r2 == Ea) add(1.0, src0 + k * C::PP + x);                    // box pair of E_x
                        for (int y = 0; y < x; y++) {
                            const int Eb = endpoint(k, y), pe = src0 + k * C::PP + M + x * (x - 1) / 2 + y;
                            if ((r1 == Ea && r2 == Ea) || (r1 == Eb && r2 == Eb)) add(1.0, pe);
                            if (r1 == std::max(Ea, Eb) && r2 == std::min(Ea, Eb)) add(-1.0, pe);
                        }
                    }
            }
            e.count = (int) e.t.size();
            ents.push_back(e);
        }
    std::stable_sort(ents.begin(), ents.end(), [](const Ent& x, const Ent& y) { return x.count > y.count; });
    std::vector<std::vector<ProjTerm>> lane(nt);
    for (const Ent& e : ents) {
        int best = 0;
        for (int t = 1; t < nt; t++) if (lane[t].size() < lane[best].size()) best = t;
        int dest = e.dest;
        if (e.diag >= 0) dest |= PROJ_DIAG;
        if (e.count < 0) dest |= PROJ_ONE;
        if (e.count <= 0) { lane[best].push_back(ProjTerm{0.0, 0, dest}); continue; }
        for (int i = 0; i < e.count; i++)
            lane[best].push_back(ProjTerm{e.t[i].first, e.t[i].second, i + 1 == e.count ? dest : -1});
    }
    ProjTable out;
    out.nt = nt;
    for (int t = 0; t < nt; t++) out.len = std::max(out.len, (int) lane[t].size());
    out.term.assign((size_t) out.len * nt, ProjTerm{0.0, 0, -1});
    for (int t = 0; t < nt; t++)
        for (size_t i = 0; i < lane[t].size(); i++) out.term[i * nt + t] = lane[t][i];
    return out;
}

// Table of the dual active-set first pass (das_kernel.cuh).  The reduced Hessian of the objective is the same for every
// agent up to the number of terminal segments ts (traj_optimizer.cpp:301-315, :530-538) and block diagonal over the
// dimensions: H1(ts) = Z1' (blkdiag_m 2 w_c Q_base + 2 w_t e5 e5' for m >= M - ts) Z1 with Z1 the continuity map of one
// dimension.  For ts = 1..M: [N1][N1] H1^-1, then [N1][N1] J = L^-T (H1 = L L').  Empty when a factorisation fails.
template <class C>
inline std::vector<double> build_das_table(const double* Q2, double w_t) {
    constexpr int M = C::M, N1 = C::NR / C::D, NF = 6 * M;
    static const double T[3][3] = {{0, 0, 1}, {0, -1, 2}, {1, -4, 4}};
    std::vector<double> Z((size_t) NF * N1, 0.0);
    for (int m = 0; m < M; m++)
        for (int i = 0; i < 6; i++) {
            const int f = m * 6 + i;
            if (i >= 3) Z[(size_t) f * N1 + ((C::TERM && m == M - 1) ? 3 * (M - 1) : 3 * m + i - 3)] = 1.0;
            else if (m >= 1) for (int j = 0; j < 3; j++) Z[(size_t) f * N1 + 3 * (m - 1) + j] = T[i][j];
        }
    std::vector<double> out((size_t) M * 2 * N1 * N1, 0.0);
    for (int ts = 1; ts <= M; ts++) {
        std::vector<double> P((size_t) NF * NF, 0.0), PZ((size_t) NF * N1, 0.0), H((size_t) N1 * N1, 0.0);
        for (int m = 0; m < M; m++) {
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) P[(size_t) (m * 6 + a) * NF + m * 6 + b] = Q2[a * 6 + b];
            if (m >= M - ts) P[(size_t) (m * 6 + 5) * NF + m * 6 + 5] += 2.0 * w_t;
        }
        for (int f = 0; f < NF; f++) for (int r = 0; r < N1; r++) {
            double a = 0; for (int g = 0; g < NF; g++) a += P[(size_t) f * NF + g] * Z[(size_t) g * N1 + r];
            PZ[(size_t) f * N1 + r] = a;
        }
        for (int r = 0; r < N1; r++) for (int c = 0; c < N1; c++) {
            double a = 0; for (int f = 0; f < NF; f++) a += Z[(size_t) f * N1 + r] * PZ[(size_t) f * N1 + c];
            H[(size_t) r * N1 + c] = a;
        }
        std::vector<double> L((size_t) N1 * N1, 0.0), Li((size_t) N1 * N1, 0.0);
        for (int j = 0; j < N1; j++) {
            double d = 0.5 * (H[(size_t) j * N1 + j] + H[(size_t) j * N1 + j]);
            for (int k = 0; k < j; k++) d -= L[(size_t) j * N1 + k] * L[(size_t) j * N1 + k];
            if (!(d > 0.0)) return {};
            L[(size_t) j * N1 + j] = std::sqrt(d);
            for (int i = j + 1; i < N1; i++) {
                double a = 0.5 * (H[(size_t) i * N1 + j] + H[(size_t) j * N1 + i]);
                for (int k = 0; k < j; k++) a -= L[(size_t) i * N1 + k] * L[(size_t) j * N1 + k];
                L[(size_t) i * N1 + j] = a / L[(size_t) j * N1 + j];
            }
        }
        for (int c = 0; c < N1; c++)                       // L Li = I, column by column (forward substitution)
            for (int i = c; i < N1; i++) {
                double a = (i == c) ? 1.0 : 0.0;
                for (int k = c; k < i; k++) a -= L[(size_t) i * N1 + k] * Li[(size_t) k * N1 + c];
                Li[(size_t) i * N1 + c] = a / L[(size_t) i * N1 + i];
            }
        double* Hinv = out.data() + (size_t) (ts - 1) * 2 * N1 * N1;
        double* J = Hinv + (size_t) N1 * N1;
        for (int r = 0; r < N1; r++) for (int c = 0; c < N1; c++) {
            J[(size_t) r * N1 + c] = Li[(size_t) c * N1 + r];                       // J = Li'
            double a = 0; for (int k = 0; k < N1; k++) a += Li[(size_t) k * N1 + r] * Li[(size_t) k * N1 + c];
            Hinv[(size_t) r * N1 + c] = a;                                           // H^-1 = Li' Li
        }
    }
    return out;
}

// Light instances (two-pass dispatch, SolveParams::klass_mode): one obstacle group, LSCQP_LIGHT_KPT kept obstacles at
// most, a quarter of the threads -- so the serial factorisation of one QP overlaps the row sweeps of many others on the
// same SM.  Agents whose presolve keeps more obstacles fall through to the full-capacity instance.
#ifndef LSCQP_LIGHT_G
#define LSCQP_LIGHT_G 1
#endif
#ifndef LSCQP_LIGHT_KPT
#define LSCQP_LIGHT_KPT 8
#endif
template <int M_, int D_, bool T_, bool COMM_>
struct Instance {
    using Full = Cfg<M_, D_, T_, 4, 10, COMM_>;
    using Light = Cfg<M_, D_, T_, LSCQP_LIGHT_G, LSCQP_LIGHT_KPT, false>;
    static constexpr bool HAS_LIGHT = !COMM_;
    // dual active-set first pass (das_kernel.cuh): the M = 5 models without communication-range rows.  For M = 10 the
    // randomised sweep (scripts/gpu_fuzz.py) shows the method at a disadvantage on both counts: 90-170 iterations (60
    // variables, many drops) against 10-14 interior-point iterations, and a reduced Hessian with nearly flat directions,
    // where its 4e-8 stationarity residual leaves a few agents per thousand up to 6e-5 m from the optimum.
    static constexpr bool HAS_DAS = !COMM_ && M_ == 5;
    // ... and its large instance (up to 20 kept obstacles, row constants in shared memory, as many active rows as there are
    // reduced variables) for the agents whose presolve keeps more obstacles, or whose optimum has more active rows, than
    // the throughput instance holds -- the dense spots of a closed loop, near-vertex solutions.  Beyond 20 kept obstacles the interior point is the better method (measured on config 4's synthetic
    // planes, 30-40 kept: ~79 active-set iterations with many drops, 3.05 ms per 4096 QPs against 2.75 ms)
    static constexpr bool HAS_DAS_BIG = HAS_DAS;
    static constexpr int DAS_BIG_KPT = 20;
    // Communication-range configurations with few neighbours (the shipped 10-agent missions: K <= 9): half the threads
    // (2 obstacle groups x 5 rows), so twice as many CTAs share an SM while the dense factorisation of each runs.
    // Only where one variable per thread and one comm pair per thread still fit.
    static constexpr int COMPACT_NT = 2 * (((6 * M_ + 31) / 32) * 32);
    static constexpr bool HAS_COMPACT = COMM_ && D_ * 6 * M_ <= COMPACT_NT && D_ * (M_ * (M_ - 1) / 2 + M_) <= COMPACT_NT;
    using Compact = Cfg<M_, D_, T_, 2, 5, HAS_COMPACT>;      // (a banded dummy where it does not apply: never launched)
    static constexpr int COMPACT_KMAX = 10;
};

// kernel instances: (M, D, TERM, COMM) with 4 obstacle groups x 10 rows per thread (K <= 40)
#define LSCQP_FOR_EACH_INSTANCE(X) \
    X(5, 3, true, false) X(5, 3, false, false) X(5, 2, true, false) X(5, 2, false, false) \
    X(10, 3, true, false) X(10, 3, false, false) X(10, 2, true, false) X(10, 2, false, false) \
    X(5, 3, true, true) X(5, 2, true, true) X(10, 3, true, true) X(10, 2, true, true) \
    X(5, 3, false, true) X(5, 2, false, true) X(10, 3, false, true) X(10, 2, false, true)

}  // namespace lscqp
