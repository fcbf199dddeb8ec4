// goal_kernel.cuh -- batched GoalOptimizer (src/goal_optimizer.cpp:7-165), the second CPLEX call of a replan
// (TrajPlanner::goalPlanningWithGridBasedPlanner, src/traj_planner.cpp:545-550).
//
// The reference's model is a one-variable LP:  min t,  0 <= t <= 1 + SP_EPSILON_FLOAT,
//     n_r . ((g - w) t + w - p_r) - d_r >= 0
// over the SFC faces of the last segment (world_use_octomap) and the LSC record (oi, M-1, n) of every obstacle
// with a float normal of at least SP_EPSILON_FLOAT; the new goal is (g - w) * t + w (g = current_goal_point,
// w = next_waypoint).  An LP in one variable is an interval intersection: every row a t + b >= 0 with a > 0 is a
// lower bound -b/a, with a < 0 an upper bound, with a = 0 a feasibility test.  One warp per agent: the lanes
// stride over the obstacles, warp shuffles reduce the largest lower bound, and a second pass checks every row at
// the optimum (infeasible = the reference's `throw PlanningReport::QPFAILED`, :94, :103).
#pragma once
#include <math.h>

namespace lscqp {

struct GoalParams {
    int n_agents, M, dim, use_sfc;
    double feas_tol;                // row violation tolerated at the optimum (CPLEX feasibility tolerance, 1e-6)
    const float*  goal;             // [n][3]  agent.current_goal_point (of the previous replan)
    const float*  waypoint;         // [n][3]  agent.next_waypoint
    const float*  sfc;              // [n][M][6] box_min, box_max (use_sfc)
    const int*    obs_offsets;      // [n+1]
    const double* normals;          // [sumK][M][3]
    const double* rhs;              // [sumK][M][6]   b = n.p + d
    float*  goal_out;               // [n][3]
    double* t_out;                  // [n]   (may be null)
    int*    status_out;             // [n]   0 ok | 2 infeasible
};

__global__ void __launch_bounds__(128) goal_lp_kernel(const GoalParams p) {
    const int lane = threadIdx.x & 31;
    const int agent = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (agent >= p.n_agents) return;
    const float gx = p.goal[agent * 3], gy = p.goal[agent * 3 + 1], gz = p.goal[agent * 3 + 2];
    const float wx = p.waypoint[agent * 3], wy = p.waypoint[agent * 3 + 1], wz = p.waypoint[agent * 3 + 2];
    // current_goal_point.distance(next_waypoint) < SP_EPSILON_FLOAT: keep the waypoint (:12-14); Vector3::distance
    // takes the differences and squares in double
    const double ddx = (double) gx - (double) wx, ddy = (double) gy - (double) wy, ddz = (double) gz - (double) wz;
    if (sqrt(ddx * ddx + ddy * ddy + ddz * ddz) < 1e-5) {
        if (lane == 0) {
            p.goal_out[agent * 3] = wx; p.goal_out[agent * 3 + 1] = wy; p.goal_out[agent * 3 + 2] = wz;
            if (p.t_out) p.t_out[agent] = 0.0;
            p.status_out[agent] = 0;
        }
        return;
    }
    // point3d difference: float subtraction, then widened where Concert multiplies it (:133-134, :152-153)
    const float fx = __fsub_rn(gx, wx), fy = __fsub_rn(gy, wy), fz = __fsub_rn(gz, wz);
    const double dx = (double) fx, dy = (double) fy, dz = (p.dim == 3) ? (double) fz : 0.0;
    const int obs0 = p.obs_offsets[agent], K = p.obs_offsets[agent + 1] - obs0;
    const int nsfc = p.use_sfc ? 2 * p.dim : 0;
    const float* box = p.use_sfc ? p.sfc + ((size_t) agent * p.M + (p.M - 1)) * 6 : nullptr;

    // row r of this agent as (a, b):  r < nsfc box faces of the last segment (Box::convertToLSCs), then obstacles
    auto row = [&](int r, double& a, double& b) {
        if (r < nsfc) {
            const int k = r >> 1;
            const double d = (k == 0) ? dx : (k == 1 ? dy : dz);
            const double w = (double) ((k == 0) ? wx : (k == 1 ? wy : wz));
            if (r & 1) { a = -d; b = (double) box[3 + k] - w; }
            else { a = d; b = w - (double) box[k]; }
            return;
        }
        const size_t o = (size_t) (obs0 + r - nsfc) * p.M + (p.M - 1);
        const double nx = p.normals[o * 3], ny = p.normals[o * 3 + 1], nz = p.normals[o * 3 + 2];
        const float qx = (float) nx, qy = (float) ny, qz = (float) nz;
        const float nsq = __fadd_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)), __fmul_rn(qz, qz));
        if (sqrt((double) nsq) < 1e-5) { a = 0.0; b = 1.0; return; }            // skipped row (:148-150)
        a = nx * dx + ny * dy; b = nx * (double) wx + ny * (double) wy;
        if (p.dim == 3) { a += nz * dz; b += nz * (double) wz; }
        b -= p.rhs[o * 6 + 5];
    };

    double t = 0.0;
    for (int r = lane; r < nsfc + K; r += 32) {
        double a, b;
        row(r, a, b);
        if (a > 0.0) t = fmax(t, -b / a);
    }
    for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
    t = fmin(t, 1.0 + 1e-5);
    int bad = 0;
    for (int r = lane; r < nsfc + K; r += 32) {
        double a, b;
        row(r, a, b);
        if (a * t + b < -p.feas_tol) bad = 1;
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        // goal = (current_goal_point - next_waypoint) * vals[0] + next_waypoint in point3d (float) arithmetic (:51)
        const float tf = (float) t;
        p.goal_out[agent * 3] = __fadd_rn(__fmul_rn(fx, tf), wx);
        p.goal_out[agent * 3 + 1] = __fadd_rn(__fmul_rn(fy, tf), wy);
        p.goal_out[agent * 3 + 2] = __fadd_rn(__fmul_rn(fz, tf), wz);
        if (p.t_out) p.t_out[agent] = t;
        p.status_out[agent] = bad ? 2 : 0;
    }
}

}  // namespace lscqp
