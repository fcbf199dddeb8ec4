/* oracle/lscqp_oracle.h -- CPU restatement of the reference's agent-QP hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under lsc_dr_planner_b200/ may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and there only as the checker.
 *
 * PARITY PIN STATUS
 *   - LSC assembly (orc_generate_*): PINNED.  The min-norm-point routine is checked
 *     against the reference's own openGJK compiled from /root/reference
 *     (oracle/_ref/libopengjk_ref.so, recipe: oracle/Makefile) and against the
 *     golden vectors that build produced (tests/golden/gjk_golden.npz).
 *   - QP model (orc_qp_build): constants pinned against the exact-rational
 *     known-answer tables; rows restate
 *     src/traj_optimizer.cpp:216-514 line by line.
 *   - QP *solution*: PARITY UNPINNED.  The reference solves with IBM CPLEX 20.1
 *     (proprietary, absent; CMakeLists.txt:37-51) and ships no tests, LP dumps or
 *     stored solutions.  The oracle solves the restated model with HiGHS 1.12
 *     (SciPy-bundled); the QP has a unique minimiser, so agreement is checked by
 *     KKT certificate + coefficient distance.
 */
#ifndef LSCQP_ORACLE_H
#define LSCQP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* PlannerMode, include/sp_const.hpp:19-26 (same integer values) */
enum { ORC_MODE_DLSC = 0, ORC_MODE_LSC = 1, ORC_MODE_BVC = 2, ORC_MODE_ORCA = 3,
       ORC_MODE_RECIPROCALRSFC = 4 };

typedef struct {
    int M, n, phi, phi_n, dim;
    double dt, w_control, w_terminal;
    int planner_mode;
    int use_sfc;                 /* param.world_use_octomap */
    double comm_range;           /* param.communication_range, <=0 disables */
    double world_min[3], world_max[3];
    double z_2d;
} orc_config;

typedef struct {
    float position[3], velocity[3], acceleration[3];   /* State, sp_const.hpp:140-144 */
    float current_goal_point[3];
    float next_waypoint[3];
    double max_vel[3], max_acc[3];
    double radius, nominal_velocity;
} orc_agent;

/* include/polynomial.hpp:9-20, 90-100, 281-294 */
int  orc_nchoosek(int n, int k);
int  orc_coef_derivative(int n, int phi);
void orc_bernstein_basis(int n, double *B /* (n+1)^2 row-major */);

/* src/traj_optimizer.cpp:163-178, 180-214, 530-538 */
void orc_build_qbase(int n, int phi, int phi_n, double dt, double *Q /* (n+1)^2 */);
int  orc_build_aeq_base(int M, int n, int phi, double dt, double *Aeq /* (M-2)*phi x M*(n+1) */);
int  orc_terminal_segments(const orc_config *cfg, const orc_agent *ag);

/* Model sizes for populatebyrow (src/traj_optimizer.cpp:216-514) without slack vars. */
void orc_qp_sizes(const orc_config *cfg, int K, const float *lsc_normal /* [K][M][n+1][3] */,
                  int *nv, int *ne, int *ni);

/* Dense restatement of populatebyrow.  Rows come out in the reference's order.
 *   objective  = x'Px + q'x + c0   (no 1/2: traj_optimizer.cpp:294)
 *   equalities   Aeq x = beq
 *   ranged rows  rlo <= G x <= rhi   (one side is +-1e30)
 *   bounds       lb <= x <= ub       (+-1e30 = free)
 * lsc_*: [K][M][n+1] records as stored by CollisionConstraints::setLSC
 * (collision_constraints.cpp:514-539); sfc: [M][6] box_min, box_max. */
int orc_qp_build(const orc_config *cfg, const orc_agent *ag, int K,
                 const float *lsc_point, const float *lsc_normal, const double *lsc_d,
                 const float *sfc,
                 double *P, double *q, double *c0,
                 double *Aeq, double *beq,
                 double *G, double *rlo, double *rhi,
                 double *lb, double *ub);

/* Closest point of conv{pts} to the origin (what gjk() returns in v,
 * src/openGJK/openGJK.cpp:674-780), by exhaustive enumeration of the faces. */
double orc_min_norm_hull(const double *pts /* [npts][3] */, int npts, double *v /* [3] */);

/* include/geometry.hpp:67-102, 129-264 in float arithmetic */
void orc_closest_points_segments(const float *l1s, const float *l1e, const float *l2s, const float *l2e,
                                 float *cp1, float *cp2, double *dist);

/* LSC generators (src/traj_planner.cpp:611-657 generateLSC, :659-706 generateCLSC,
 * :708-736 generateBVC) for agent-type obstacles.
 *   own_traj [M][n+1][3], obs_traj [K][M][n+1][3], obs_radius/obs_downwash [K],
 *   obs_goal [K][3] (CLSC), obs_position [K][3] (LSC zero-normal fallback)
 * out: lsc_point/lsc_normal [K][M][n+1][3] float, lsc_d [K][M][n+1] double */
enum { ORC_GEN_LSC = 0, ORC_GEN_CLSC = 1, ORC_GEN_BVC = 2, ORC_GEN_RSFC = 3 /* generateReciprocalRSFC :581-609 */ };
void orc_obstacle_sizes(int M, int n, double dt, double obs_radius, double obs_max_acc, double uncertainty_horizon,
                        double velocity_guard, double *size /* [M][n+1] */);
extern const double *orc_rsfc_sizes;   /* [K][M][n+1] sizes used by ORC_GEN_RSFC (NULL: the obstacle radius) */
void orc_generate_lsc(const orc_config *cfg, int generator, const orc_agent *ag, double agent_downwash,
                      const float *own_traj, int K, const float *obs_traj,
                      const float *obs_radius, const float *obs_downwash,
                      const float *obs_goal, const float *obs_position,
                      float *lsc_point, float *lsc_normal, double *lsc_d);

/* Closed-loop glue: Trajectory::getStateAt (src/trajectory.cpp:111-199),
 * previous-solution shift (src/traj_planner.cpp:287-297, 402-411),
 * planConstVelTraj (src/trajectory.cpp:77-89). */
void orc_get_state_at(int M, int n, double dt, const float *traj, double time, float *state9);
void orc_shift_traj(int M, int n, const float *prev, float *out);
void orc_const_vel_traj(int M, int n, double dt, const float *pos, const float *vel, float *out);

/* TrajPlanner::isSolValid, src/traj_planner.cpp:990-1045 (1 valid | 0 not) */
int orc_is_sol_valid(const orc_config *cfg, const orc_agent *ag, const float *traj, const float *state9, const float *sfc);

/* GoalOptimizer, src/goal_optimizer.cpp:7-165: rows a t + b >= 0 of the one-variable LP (returns the row count,
 * a / b sized 2 dim + K) and its closed-form optimum (0 ok | 2 infeasible = the reference's QPFAILED throw). */
int orc_goal_rows(const orc_config *cfg, const float *goal, const float *waypoint, int K,
                  const float *lsc_point, const float *lsc_normal, const double *lsc_d,
                  const float *sfc_last, double *a, double *b);
int orc_goal_solve(const orc_config *cfg, const float *goal, const float *waypoint, int nrows,
                   const double *a, const double *b, double feas_tol, float *goal_out, double *t_out);

#ifdef __cplusplus
}
#endif
#endif
