// host_common.hpp -- host-side constants shared by the CUDA library and the CPU emulation harness.
#pragma once
#include <cmath>
#include <cstring>
#include "../../include/lscqp.h"
#include "pdip_kernel.cuh"

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
    if (!(c.planner_mode == LSCQP_MODE_DLSC || c.planner_mode == LSCQP_MODE_LSC || c.planner_mode == LSCQP_MODE_BVC))
        return LSCQP_E_INVALID;
    if (c.comm_range > 0) return LSCQP_E_INVALID;
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
    p.dt = c.dt; p.w_t = c.w_terminal; p.w_c = c.w_control;
    for (int k = 0; k < 3; k++) { p.world_min[k] = c.world_min[k]; p.world_max[k] = c.world_max[k]; }
    p.use_sfc = c.use_sfc;
    double Q[36];
    jerk_gram(c.n, c.phi, c.dt, Q);
    for (int e = 0; e < 36; e++) p.Q2[e] = 2.0 * c.w_control * Q[e];
}

// kernel instances: (M, D, TERM) with 4 obstacle groups x 10 rows per thread (K <= 40)
#define LSCQP_FOR_EACH_INSTANCE(X) \
    X(5, 3, true) X(5, 3, false) X(5, 2, true) X(5, 2, false) \
    X(10, 3, true) X(10, 3, false) X(10, 2, true) X(10, 2, false)

}  // namespace lscqp
