/* include/lscqp.h -- C ABI of the B200 batched agent-QP path (liblscqp.so).
 *
 * Drop-in boundary for the reference's per-agent planning hot path.  The reference has no FFI
 * layer; the boundary there is two C++ classes owned by TrajPlanner
 * (/root/reference/include/traj_planner.hpp:104,110):
 *     TrajOptimizer::solve            include/traj_optimizer.hpp:25-30, src/traj_optimizer.cpp:18-156
 *     CollisionConstraints (LSC/SFC)  include/collision_constraints.hpp:98-201
 * and the LSC generators TrajPlanner::generateLSC / generateCLSC / generateBVC
 * (src/traj_planner.cpp:611-736).  include/lscqp_shim.hpp re-creates those classes on top of
 * this ABI; INTEGRATION.md shows the patch a maintainer applies.
 *
 * Conventions
 *   - POD only, caller-owned buffers, no exceptions across the boundary.
 *   - Every function returns 0 on success or a negative LSCQP_E_* code (nothing written).
 *     Per-agent solver outcomes are reported through status_out only.
 *   - Buffers are DEVICE pointers for the *_batch entry points and HOST pointers for the *_host
 *     entry points (which stage through pinned memory owned by the handle and include the
 *     host<->device copies).  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - One handle per (device, host thread); calls on a handle are stream-ordered, not re-entrant.
 *   - Obstacle lists are never truncated.  The *_host entry points validate them (every list within max_obs, every
 *     neighbour id inside the batch) before anything is launched and return LSCQP_E_CAPACITY / LSCQP_E_INVALID.  The
 *     device entry points cannot inspect device memory without a synchronisation: an agent whose list is longer than
 *     max_obs is reported through status_out = LSCQP_CAPACITY and left unsolved (ctrl_out = its starting point).
 *   - No CPU fallback exists: without a CUDA device lscqp_create fails with LSCQP_E_NODEVICE.
 */
#ifndef LSCQP_H
#define LSCQP_H

#ifdef __cplusplus
extern "C" {
#endif

#define LSCQP_E_INVALID     (-1)   /* bad argument / unsupported configuration            */
#define LSCQP_E_NODEVICE    (-2)   /* no usable CUDA device                                */
#define LSCQP_E_CUDA        (-3)   /* CUDA runtime error (see lscqp_last_error)            */
#define LSCQP_E_CAPACITY    (-4)   /* n_agents / obstacle count above the handle's capacity */

/* per-agent status_out values */
#define LSCQP_OK            0
#define LSCQP_MAX_ITER      1
#define LSCQP_INFEASIBLE    2
#define LSCQP_NUMERICAL     3
#define LSCQP_CAPACITY      4      /* obstacle list longer than max_obs (or, without presolve, than the kernel instance
                                      holds): nothing was dropped, the agent was not solved; ctrl_out = starting point */

/* PlannerMode, include/sp_const.hpp:19-26 (same integer values) */
#define LSCQP_MODE_DLSC 0
#define LSCQP_MODE_LSC  1
#define LSCQP_MODE_BVC  2
#define LSCQP_MODE_RECIPROCALRSFC 4   /* SlackMode::COLLISIONCONSTRAINT (src/param.cpp:157-161): the LSC rows carry free,
                                         cost-less slacks (traj_optimizer.cpp:272-283, 423-425) and cannot bind; z bounds of
                                         segment 0 relaxed to +-100 (:255-258) */

/* LSC generator selected by TrajPlanner::constructLSC, src/traj_planner.cpp:552-569 */
#define LSCQP_GEN_LSC   0          /* generateLSC  :611-657 */
#define LSCQP_GEN_CLSC  1          /* generateCLSC :659-706 (mode lsc + grid_based_planner) */
#define LSCQP_GEN_BVC   2          /* generateBVC  :708-736 */
#define LSCQP_GEN_RSFC  3          /* generateReciprocalRSFC :581-609 (obstacle sizes: lscqp_set_obstacle_sizes) */

/* The fields of Param / Mission the QP reads (src/param.cpp:5-173). */
typedef struct lscqp_config {
    int M, n, phi, dim;            /* segments (5|10), degree (5), phi (3), world_dimension (2|3) */
    double dt, w_control, w_terminal;   /* param.dt, control_input_weight, terminal_weight    */
    int planner_mode;              /* LSCQP_MODE_*; LSC adds the terminal-stop equalities      */
    int use_sfc;                   /* param.world_use_octomap: box constraints from `sfc`      */
    double comm_range;             /* param.communication_range; > 0 adds the rows of
                                      traj_optimizer.cpp:477-500 (LSC mode only), <= 0 disables  */
    double world_min[3], world_max[3], z_2d;   /* mission.world_min/max, param.world_z_2d      */
    int max_obs;                   /* capacity: obstacles per agent (<= 40)                    */
    int max_agents;                /* capacity of the staging buffers for the *_host calls     */
    int max_iter;                  /* interior-point iteration cap (0 = default 60)            */
    double tol;                    /* complementarity / primal tolerance (0 = default 1e-11)   */
    int presolve;                  /* bit 0: drop obstacles whose rows are all proven inactive by
                                      bound propagation through the velocity rows (exact);
                                      bit 1: keep every agent on the full-capacity kernel
                                      instance (no light-instance first pass);
                                      bit 2: light first pass at any batch size (default: only
                                      from 1536 agents, below that a batch is latency bound);
                                      bit 3: no dual active-set first pass (das_kernel.cuh): the
                                      interior-point instances alone                            */
} lscqp_config;

typedef struct lscqp_handle lscqp_handle;

/* Packed LSC half-spaces: row (oi, m, i) reads  normal[oi][m] . c[m][i] >= rhs[oi][m][i]
 * (rhs = n . obs_control_point + d, the constant of traj_optimizer.cpp:413-429). */
typedef struct lscqp_planes {
    const int*    obs_offsets;     /* [n_agents+1] CSR offsets into the obstacle axis          */
    double*       normals;         /* [sum K][M][3]                                            */
    double*       rhs;             /* [sum K][M][6]                                            */
} lscqp_planes;

const char* lscqp_version(void);
const char* lscqp_last_error(void);

int lscqp_create(const lscqp_config* cfg, int device, lscqp_handle** out);
int lscqp_destroy(lscqp_handle* h);

/* Number of doubles per agent in dual_out (layout: [max_obs_padded][M][6] LSC rows, then
 * [dim*M*6][6] box rows = lb, ub, vel+, vel-, acc+, acc- of the variable's stencil, then with
 * comm_range > 0 [dim][M + M(M-1)/2][2] communication pairs = upper-side, lower-side). */
int lscqp_dual_stride(const lscqp_handle* h);
/* Number of CUDA kernels this handle has launched so far (bench.py reports it as gpu_launches). */
unsigned long long lscqp_launch_count(const lscqp_handle* h);
int lscqp_max_obs_padded(const lscqp_handle* h);

/* LSC assembly for agent-type obstacles: generateLSC / generateCLSC / generateBVC
 * (src/traj_planner.cpp:611-736; normals by GJK as normalVectorBetweenPolys :1179-1205).
 * All pointers are device pointers. */
int lscqp_assemble_lsc_batch(lscqp_handle* h, int generator, int n_agents,
        const float* own_traj,      /* [n_agents][M][6][3]  initial_traj                        */
        const double* agent_meta,   /* [n_agents][2]        radius, downwash (doubles in Agent) */
        const float* agent_goal,    /* [n_agents][3]        current_goal_point (CLSC, LSC fallback) */
        const int*   obs_offsets,   /* [n_agents+1]                                              */
        const float* obs_traj,      /* [sum K][M][6][3]     obs_pred_trajs                      */
        const float* obs_meta,      /* [sum K][4]           radius, downwash, -, -              */
        const float* obs_goal,      /* [sum K][3]           obstacle goal_point (CLSC)          */
        const float* obs_position,  /* [sum K][3]           obstacle position (LSC fallback)    */
        double* normals_out,        /* [sum K][M][3]                                             */
        double* rhs_out,            /* [sum K][M][6]                                             */
        void* stream);

/* Predicted obstacle sizes for LSCQP_GEN_RSFC (obs_pred_sizes of obstacleSizePredictionWithConstAcc,
 * src/traj_planner.cpp:321-358): DEVICE array [sum K][M][6], or NULL = every obstacle's radius (obs/size_prediction off).
 * Kept by the handle until set again. */
int lscqp_set_obstacle_sizes(lscqp_handle* h, const double* obs_size);

/* Fused variant for obstacles that are agents of the same population (what MultiSyncSimulator::broadcastMsgs hands
 * to every planner, src/multi_sync_simulator.cpp:305-352): the obstacles' trajectories / radii / goals / positions are
 * read in place through obs_index from the population arrays (all_*: [n_total] rows; no gathered copies), and with
 * prune != 0 an (obstacle, segment) pair whose rows provably cannot bind at any point the velocity rows allow is
 * written as a zero normal -- a row the QP drops (traj_optimizer.cpp:409-411) -- without running the hull enumeration.
 * The minimiser of the QP is unchanged; the planes are no longer the reference's for the dropped pairs, so
 * GoalOptimizer (which reads the last control point's plane without the velocity rows) must use prune = 0. */
int lscqp_assemble_lsc_fused(lscqp_handle* h, int generator, int prune, int n_agents,
        const float* own_traj, const double* agent_meta, const float* agent_goal,
        const float* state,          /* [n_agents][9]  (prune)                                   */
        const double* limits,        /* [n_agents][8]  (prune)                                   */
        const int* obs_offsets, const int* obs_index,
        const float* all_traj,       /* [n_total][M][6][3]                                        */
        const double* all_meta,      /* [n_total][2]   radius, downwash                           */
        const float* all_goal,       /* [n_total][3]                                              */
        const float* all_state,      /* [n_total][9]                                              */
        double* normals_out, double* rhs_out, void* stream);

/* Batched QP solve (TrajOptimizer::solve for every agent of the batch).  Device pointers. */
int lscqp_solve_batch(lscqp_handle* h, int n_agents,
        const float*  state,        /* [n_agents][9]  position, velocity, acceleration          */
        const float*  goal,         /* [n_agents][3]  current_goal_point                        */
        const double* limits,       /* [n_agents][8]  max_vel[3], max_acc[3], radius, nominal_velocity */
        const float*  sfc,          /* [n_agents][M][6] box_min, box_max (use_sfc) or NULL      */
        const float*  next_waypoint,/* [n_agents][3] agent.next_waypoint (comm_range > 0) or NULL */
        const int*    obs_offsets,  /* [n_agents+1]                                              */
        const double* normals,      /* [sum K][M][3]                                             */
        const double* rhs,          /* [sum K][M][6]                                             */
        const float*  initial_traj, /* [n_agents][M][6][3] optional (NULL = cold start): the initial_traj the
                                       reference passes to TrajOptimizer::solve; used only as the solver's
                                       starting point, never in the model (traj_optimizer.cpp:216-218)  */
        double* ctrl_out,           /* [n_agents][dim][M][6]  index order of traj_optimizer.cpp:241 */
        double* cost_out,           /* [n_agents]  objective incl. constant (cplex.getObjValue) */
        int*    status_out,         /* [n_agents]  LSCQP_OK | MAX_ITER | INFEASIBLE | NUMERICAL */
        int*    iters_out,          /* [n_agents]  optional                                      */
        double* kkt_out,            /* [n_agents][4] optional: stationarity, primal, guarded pivots (interior point) or
                                       64 * largest + final active-set size (active set), gap (0 for the active set) */
        double* dual_out,           /* [n_agents][lscqp_dual_stride] optional                   */
        void* stream);

/* Diagnostics: which pass solved each agent in the last lscqp_solve_batch call on this handle (0 = the first pass: the
 * dual active-set kernel, or the light one-warp interior-point instance when that pass is switched off; otherwise the
 * full-capacity interior-point instance -- after an active-set first pass the value is its reason for deferring,
 * das_kernel.cuh); klass_out is a HOST array [n_agents]; synchronises `stream`.
 * LSCQP_E_INVALID when that call ran the full-capacity instance alone. */
int lscqp_last_instances(lscqp_handle* h, int n_agents, int* klass_out, void* stream);

/* Same as lscqp_solve_batch with HOST buffers: copies inputs host->device, solves, copies
 * ctrl/cost/status (and the optional outputs) back, and synchronises the stream. */
int lscqp_solve_host(lscqp_handle* h, int n_agents,
        const float* state, const float* goal, const double* limits, const float* sfc, const float* next_waypoint,
        const int* obs_offsets, const double* normals, const double* rhs, const float* initial_traj,
        double* ctrl_out, double* cost_out, int* status_out, int* iters_out, double* kkt_out,
        double* dual_out);

/* Fused replan step with HOST buffers: copy trajectories in, assemble LSCs on the device,
 * solve, copy the solutions back.  Obstacles are given as indices into the batch's own agents
 * (what MultiSyncSimulator::broadcastMsgs builds, src/multi_sync_simulator.cpp:305-352). */
int lscqp_replan_host(lscqp_handle* h, int generator, int n_agents,
        const float* state, const float* goal, const double* limits, const float* sfc, const float* next_waypoint,
        const float* own_traj, const double* agent_meta,
        const int* obs_offsets, const int* obs_index,      /* [sum K] neighbour agent ids */
        double* ctrl_out, double* cost_out, int* status_out, int* iters_out);

/* Batched GoalOptimizer::solve (src/goal_optimizer.cpp:7-165; called from
 * TrajPlanner::goalPlanningWithGridBasedPlanner, src/traj_planner.cpp:545-550): the one-variable LP
 *     min t,  0 <= t <= 1 + 1e-5,  n.((g - w) t + w - p) - d >= 0
 * over the SFC faces of the last segment (use_sfc) and the LSC record (oi, M-1, n) of every obstacle, solved in
 * closed form; goal_out = (g - w) * t + w in float arithmetic.  status_out: LSCQP_OK or LSCQP_INFEASIBLE (where the
 * reference throws PlanningReport::QPFAILED).  Uses the packed planes of lscqp_assemble_lsc_batch.  Device pointers. */
int lscqp_goal_batch(lscqp_handle* h, int n_agents,
        const float*  goal,          /* [n_agents][3]  agent.current_goal_point (previous replan)  */
        const float*  next_waypoint, /* [n_agents][3]  agent.next_waypoint                          */
        const float*  sfc,           /* [n_agents][M][6] or NULL (use_sfc)                          */
        const int*    obs_offsets, const double* normals, const double* rhs,
        float*  goal_out,            /* [n_agents][3]  new current_goal_point                       */
        double* t_out,               /* [n_agents]     optional: the LP optimum                     */
        int*    status_out,          /* [n_agents]                                                  */
        void* stream);
/* Same with HOST buffers (copies in, solves, copies back, synchronises). */
int lscqp_goal_host(lscqp_handle* h, int n_agents, const float* goal, const float* next_waypoint, const float* sfc,
        const int* obs_offsets, const double* normals, const double* rhs,
        float* goal_out, double* t_out, int* status_out);

/* Device-side gather used by lscqp_replan_*: obs_traj[j] = own_traj[obs_index[j]] etc. */
int lscqp_gather_obstacles(lscqp_handle* h, int n_obs, const int* obs_index,
        const float* own_traj, const double* agent_meta, const float* agent_goal, const float* state,
        float* obs_traj, float* obs_meta, float* obs_goal, float* obs_position, void* stream);

/* Neighbour selection on the device (MultiSyncSimulator::broadcastMsgs, src/multi_sync_simulator.cpp:305-352): for the
 * agents [lo, lo + n_local) of a population of n_total, the ids of the other agents whose position is within the
 * Chebyshev communication range (:319-328; comm_range <= 0: every other agent), in ascending id order, as a ragged CSR
 * list -- exactly the reference's obstacle set.  K (<= max_obs) is the capacity per agent: an agent with more than K in
 * range keeps its K nearest and overflow_out[a] holds its in-range count (0 = the list is complete); nothing is padded.
 * state: [n_total][9] (position first); obs_offsets_out: [n_local + 1]; obs_index_out: capacity n_local * K;
 * overflow_out: [n_local] or NULL.  Device pointers. */
int lscqp_select_neighbours(lscqp_handle* h, int n_total, int lo, int n_local, int K, double comm_range,
        const float* state, int* obs_offsets_out, int* obs_index_out, int* overflow_out, void* stream);

/* Closed-loop glue on the device (AgentManager::doStep src/agent_manager.cpp:29-50 via
 * Trajectory::getStateAt src/trajectory.cpp:156-170, and the previous-solution shift
 * src/traj_planner.cpp:287-297, 402-411).  ctrl: [n][dim][M][6] doubles from the solve;
 * writes the float trajectory [n][M][6][3] the reference stores (traj_optimizer.cpp:71-83),
 * the next state [n][9] at time `step`, and the shifted trajectory for the next replan. */
int lscqp_step_batch(lscqp_handle* h, int n_agents, const double* ctrl, double step,
        float* traj_out, float* state_out, float* shifted_traj_out, void* stream);

/* ---- Static map and Safe Flight Corridors (TrajPlanner::generateSFC, src/traj_planner.cpp:738-753).
 * lscqp_map_set builds, once per world, what MapManager::updateOctreeFromCSV (src/map_manager.cpp:262-305) and
 * DynamicEDTOctomap(maxdist) (src/map_manager.cpp:59-80) hold in the reference: the occupied cells of the world's boxes
 * (boxes: HOST array [n_boxes][6] = centre xyz, size xyz -- one row of a world CSV; param.world_resolution; the world
 * box is lscqp_config.world_min/max) and the Euclidean-nearest occupied cell of every cell within max_dist
 * (the reference passes 1.0).  lscqp_map_get copies them back (HOST arrays; occ [nx][ny][nz] bytes, closest
 * [nx][ny][nz] packed x | y << 10 | z << 20 or -1; either may be NULL).
 * lscqp_sfc_batch (device pointers) fills / advances the corridors sfc [n_agents][M][6] (box_min, box_max per segment,
 * the array lscqp_solve_batch takes) for every agent, one CTA each:
 *   LSCQP_SFC_INIT       CollisionConstraints::initializeSFC (src/collision_constraints.cpp:366-383): point = current
 *                        position; all M corridors = the box grown around its grid cell; status 1, or 0 where the
 *                        reference throws "Invalid initial SFC" (corridors untouched)
 *   LSCQP_SFC_FROM_POINT constructSFCFromPoint (:396-411, :669-694): corridors shift by one segment, the last one is grown
 *                        from point = initial_traj.lastPoint() towards goal = agent.current_goal_point (setAxisCand
 *                        :1134-1170); status 1, or 0 = no valid box, previous corridor reused
 *   LSCQP_SFC_FROM_HULL  constructSFCFromConvexHull (:413-436, :696-777; goal mode grid_based_planner): hull = {point,
 *                        goal} first together with next_waypoint (status 2), else alone inside the previous corridor
 *                        (status 1), else the previous corridor again (status 0)
 * limits: [n_agents][8] as for lscqp_solve_batch (the agent radius at [6] is the margin). */
#define LSCQP_SFC_INIT        0
#define LSCQP_SFC_FROM_POINT  1
#define LSCQP_SFC_FROM_HULL   2
int lscqp_map_set(lscqp_handle* h, const double* boxes, int n_boxes, double resolution, double max_dist);
int lscqp_map_get(lscqp_handle* h, int* n3, unsigned char* occ_out, int* closest_out);
int lscqp_sfc_batch(lscqp_handle* h, int mode, int n_agents, const float* point, const float* goal,
        const float* next_waypoint, const double* limits, float* sfc, int* status_out, void* stream);
/* the same with HOST buffers (copies in, runs, copies sfc / status back, synchronises) */
int lscqp_sfc_host(lscqp_handle* h, int mode, int n_agents, const float* point, const float* goal,
        const float* next_waypoint, const double* limits, float* sfc, int* status_out);

/* ---- Peer exchange for the sharded closed loop (BASELINE config 5: agents sharded over GPUs, one exchange of the solved
 * trajectories per replan).  Stands in for MultiSyncSimulator::broadcastMsgs (src/multi_sync_simulator.cpp:305-352), which
 * copies every agent's state and prev_traj (AgentManager::getAgent, src/agent_manager.cpp:184-199) to every planner.
 * Each rank owns one exchange block in its HBM, mapped into every peer through CUDA IPC; lscqp_step_exchange fuses the
 * failsafe + doStep + shift with the all-gather: the new rows are stored straight into every rank's block over NVLink and
 * a per-rank sequence flag is raised; lscqp_exchange_begin (first launch of the next step) waits for all flags and copies
 * the rows into the rank's replicated arrays.  No host involvement per step: both launches can be captured in a CUDA graph.
 *   create : allocates the local block for a population of n_total agents; ipc_handle_out receives 64 bytes
 *            (a cudaIpcMemHandle_t) to be exchanged between the ranks by the caller (e.g. torch.distributed.all_gather)
 *   connect: all_ipc_handles = world x 64 bytes in rank order (the own entry is ignored)
 *   begin  : traj [n_total][M][6][3], state [n_total][9] (device): overwritten with the rows published for this step;
 *            a no-op before the first publish.  A peer that does not arrive within ~2 s raises the time-out counter
 *            (lscqp_exchange_status) instead of hanging the device.
 *   step_exchange: ctrl [n_local][dim][M][6] of the agents [lo, lo + n_local); status / fallback_traj (both or neither):
 *            agents with status != 0 keep fallback_traj (TrajPlanner::trajOptimization's catch branch,
 *            src/traj_planner.cpp:767-797); traj_out (optional) receives the float trajectories [n_local][M][6][3].
 *            Every rank must call it once per step with n_local >= 1.
 *   status : out3 = { steps published, wait time-outs, failsafe uses } of this rank (synchronises the stream). */
int lscqp_exchange_create(lscqp_handle* h, int n_total, int world, int rank, void* ipc_handle_out);
int lscqp_exchange_connect(lscqp_handle* h, const void* all_ipc_handles);
int lscqp_exchange_begin(lscqp_handle* h, float* traj, float* state, void* stream);
int lscqp_step_exchange(lscqp_handle* h, int lo, int n_local, const double* ctrl, const int* status,
        const float* fallback_traj, double step, float* traj_out, void* stream);
int lscqp_exchange_status(lscqp_handle* h, unsigned long long* out3, void* stream);
int lscqp_exchange_destroy(lscqp_handle* h);
/* single-process wiring (several handles, one per device, in one process): the peers' blocks as plain device pointers */
void* lscqp_exchange_local_base(lscqp_handle* h);
int lscqp_exchange_connect_ptrs(lscqp_handle* h, void* const* bases);

/* TrajPlanner::isSolValid (src/traj_planner.cpp:990-1045) for every agent: SFC containment of the float control points
 * (use_sfc; segment 0 from point phi on) and velocity / acceleration of the state at the replanning period within 1 %
 * of the limits.  traj / state_at_step are the outputs of lscqp_step_batch (step = multisim time step).
 * valid_out[a] = 1 | 0; in DLSC mode the reference re-runs the solver when the check fails (:763-766). */
int lscqp_validate_batch(lscqp_handle* h, int n_agents, const float* traj, const float* state_at_step,
        const double* limits, const float* sfc, int* valid_out, void* stream);

/* Measurement aid (bench.py): sustained FP64 FMA rate of the device in GFLOP/s from a register-resident DFMA
 * microbenchmark (8 independent chains per thread), the denominator of the solve kernel's FP64 view
 * (MEASURED_PEAKS.json holds only the HBM and bf16 peaks). */
int lscqp_measure_fp64_peak(lscqp_handle* h, double* gflops_out);

#ifdef __cplusplus
}
#endif
#endif /* LSCQP_H */
