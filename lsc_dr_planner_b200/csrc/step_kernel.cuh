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

struct StepParams {
    int n_agents, dim;
    double dt, step, z_2d;
    const double* ctrl;      // [n][dim][M][6]
    float* traj_out;         // [n][M][6][3]
    float* state_out;        // [n][9]      (may be null)
    float* shifted_out;      // [n][M][6][3] (may be null)
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

// Trajectory::getPointAt on control points cps[m][i] (stride 6 per segment) of degree deg
template <int M>
__device__ __forceinline__ void point_at(const float (*cps)[6][3], int deg, double dt, double time, float* out) {
    float px = 0.f, py = 0.f, pz = 0.f;
    int m = -1;
    double t_norm = 0.0, seg_end = 0.0;
    if (!(time < 0)) {
        for (int idx = 0; idx < M; idx++) {
            seg_end += dt;
            if (time < seg_end) { m = idx; t_norm = 1 - (seg_end - time) / dt; break; }
        }
        if (m == -1 && time < seg_end + 1e-5) { m = M - 1; t_norm = 1.0; }     // trajectory.cpp:130-134
    }
    if (m >= 0) {
        for (int i = 0; i < deg + 1; i++) {
            const float b = (float) (binom_small(deg, i) * ipow(t_norm, i) * ipow(1 - t_norm, deg - i));   // polynomial.hpp:22-24
            px = __fadd_rn(px, __fmul_rn(cps[m][i][0], b));
            py = __fadd_rn(py, __fmul_rn(cps[m][i][1], b));
            pz = __fadd_rn(pz, __fmul_rn(cps[m][i][2], b));
        }
    }
    out[0] = px; out[1] = py; out[2] = pz;
}

template <int M>
__global__ void __launch_bounds__(128)
step_kernel(const StepParams p) {
    const int agent = blockIdx.x * blockDim.x + threadIdx.x;
    if (agent >= p.n_agents) return;
    float c[M][6][3], d1[M][6][3], d2[M][6][3];
    const double* x = p.ctrl + (size_t) agent * p.dim * M * 6;
    for (int m = 0; m < M; m++)
        for (int i = 0; i < 6; i++) {
            c[m][i][0] = (float) x[0 * M * 6 + m * 6 + i];
            c[m][i][1] = (float) x[1 * M * 6 + m * 6 + i];
            c[m][i][2] = p.dim == 3 ? (float) x[2 * M * 6 + m * 6 + i] : (float) p.z_2d;
        }
    float* to = p.traj_out + (size_t) agent * M * 18;
    for (int m = 0; m < M; m++)
        for (int i = 0; i < 6; i++)
            for (int k = 0; k < 3; k++) to[(m * 6 + i) * 3 + k] = c[m][i][k];
    if (p.state_out) {
        // Trajectory::derivative, trajectory.cpp:183-199: (c[i+1]-c[i]) * (float)(deg / segment_time)
        const float s1 = (float) (5 / p.dt), s2 = (float) (4 / p.dt);
        for (int m = 0; m < M; m++) {
            for (int i = 0; i < 5; i++)
                for (int k = 0; k < 3; k++) d1[m][i][k] = __fmul_rn(__fsub_rn(c[m][i + 1][k], c[m][i][k]), s1);
            for (int i = 0; i < 4; i++)
                for (int k = 0; k < 3; k++) d2[m][i][k] = __fmul_rn(__fsub_rn(d1[m][i + 1][k], d1[m][i][k]), s2);
        }
        float* so = p.state_out + (size_t) agent * 9;
        point_at<M>(c, 5, p.dt, p.step, so);
        point_at<M>(d1, 4, p.dt, p.step, so + 3);
        point_at<M>(d2, 3, p.dt, p.step, so + 6);
        if (p.dim == 2) so[2] = (float) p.z_2d;                                  // agent_manager.cpp:42-44
    }
    if (p.shifted_out) {
        float* sh = p.shifted_out + (size_t) agent * M * 18;
        for (int m = 0; m < M; m++)
            for (int i = 0; i < 6; i++)
                for (int k = 0; k < 3; k++)
                    sh[(m * 6 + i) * 3 + k] = (m == M - 1) ? c[M - 1][5][k] : c[m + 1][i][k];
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
