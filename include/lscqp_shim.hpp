// include/lscqp_shim.hpp -- the reference's class surfaces on top of the C ABI (include/lscqp.h).
//
// Header-only, ROS/octomap/Eigen/CPLEX-free re-creation of the two classes TrajPlanner owns
// (/root/reference/include/traj_planner.hpp:104,110):
//     DynamicPlanning::CollisionConstraints   include/collision_constraints.hpp:98-201 (LSC/SFC part)
//     DynamicPlanning::TrajOptimizer          include/traj_optimizer.hpp:18-55
//     DynamicPlanning::GoalOptimizer          include/goal_optimizer.hpp:18-38
// with the same method names, argument meaning and error behaviour, so traj_planner.cpp compiles
// against them unchanged apart from the include.  Plus BatchTrajOptimizer, the
// single dispatch the serial loop of MultiSyncSimulator::plan (src/multi_sync_simulator.cpp:354-362)
// collapses into.  All arithmetic is done by liblscqp.so on the GPU; nothing here solves anything.
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "lscqp.h"

namespace DynamicPlanning {

// ---- minimal stand-ins for the reference's value types (include/sp_const.hpp) -----------------
#ifndef LSCQP_SHIM_EXTERNAL_TYPES
struct point3d {                               // octomap::point3d: three floats
    float v[3] = {0, 0, 0};
    point3d() = default;
    point3d(float x, float y, float z) { v[0] = x; v[1] = y; v[2] = z; }
    float x() const { return v[0]; } float y() const { return v[1]; } float z() const { return v[2]; }
    float& x() { return v[0]; } float& y() { return v[1]; } float& z() { return v[2]; }
    float operator()(int i) const { return v[i]; } float& operator()(int i) { return v[i]; }
    bool operator==(const point3d& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
};
typedef point3d vector3d;
typedef std::vector<point3d> points_t;

enum class PlannerMode { DLSC, LSC, BVC, ORCA, RECIPROCALRSFC, CIRCLETEST };          // sp_const.hpp:19-26
enum class SlackMode { NONE, CONTINUITY, COLLISIONCONSTRAINT };                        // sp_const.hpp:43-47
enum PlanningReport { Initialized, INITTRAJGENERATIONFAILED, CONSTRAINTGENERATIONFAILED, QPFAILED, WAITFORROSMSG, SUCCESS };

struct State { point3d position, velocity, acceleration; };                            // sp_const.hpp:140-144
struct Agent {                                                                         // sp_const.hpp:146-160
    int id = 0, cid = 0;
    State current_state;
    point3d start_point, desired_goal_point, current_goal_point, next_waypoint;
    std::vector<double> max_vel{1, 1, 1}, max_acc{2, 2, 2};
    double radius = 0.15, downwash = 2.0, nominal_velocity = 1.0;
    bool collision_alert = false;
};
struct Param {                                 // the fields the QP reads (src/param.cpp:5-173)
    int world_dimension = 3, M = 5, n = 5, phi = 3, phi_n = 1;
    double dt = 0.2, control_input_weight = 1, terminal_weight = 1, communication_range = 3.0, world_z_2d = 1.0;
    PlannerMode planner_mode = PlannerMode::LSC;
    SlackMode slack_mode = SlackMode::NONE;
    bool world_use_octomap = false, log_solver = false;
    double world_resolution = 0.1;
    std::string package_path;
};
struct Mission { point3d world_min{-5, -5, 0}, world_max{5, 5, 2.5}; };               // include/mission.hpp (point3d there too)

template <typename T> struct Segment {                                                 // include/trajectory.hpp:9-28
    std::vector<T> control_points; double segment_time = 0;
    T operator[](int i) const { return control_points[i]; } T& operator[](int i) { return control_points[i]; }
};
template <typename T> class Trajectory {                                               // include/trajectory.hpp:31-65
public:
    Trajectory() = default;
    Trajectory(size_t M, size_t n, double dt) : segments(M) { for (auto& s : segments) { s.control_points.resize(n + 1); s.segment_time = dt; } }
    int size() const { return (int) segments.size(); }
    bool empty() const { return segments.empty(); }
    Segment<T> operator[](int i) const { return segments[i]; } Segment<T>& operator[](int i) { return segments[i]; }
private:
    std::vector<Segment<T>> segments;
};
typedef Trajectory<point3d> traj_t;
#endif

// ---- LSC / Box / CollisionConstraints (include/collision_constraints.hpp:19-96, 98-201) --------
class LSC {
public:
    LSC() = default;
    LSC(const point3d& p, const point3d& n, double d_) : obs_control_point(p), normal_vector(n), d(d_) {}
    point3d obs_control_point, normal_vector;
    double d = 0;
};
typedef std::vector<LSC> LSCs;

class Box {
public:
    point3d box_min, box_max;
    Box() = default;
    Box(const point3d& mn, const point3d& mx) : box_min(mn), box_max(mx) {}
    LSCs convertToLSCs(int dim) const {                                                 // collision_constraints.cpp:37-59
        LSCs out(2 * dim);
        for (int i = 0; i < dim; i++) {
            point3d nmin, nmax; nmin(i) = 1; nmax(i) = -1;
            out[2 * i] = LSC(point3d(), nmin, box_min(i)); out[2 * i + 1] = LSC(point3d(), nmax, -box_max(i));
        }
        return out;
    }
};
typedef std::vector<std::vector<LSCs>> RSFCs;
typedef std::vector<Box> SFCs;

// The static map behind the corridors: what MapManager holds as octree + DynamicEDTOctomap (src/map_manager.cpp:59-80,
// 262-305), here the occupancy grid and nearest-obstacle field liblscqp keeps in HBM.  boxes: the rows of a world CSV.
class StaticMap {
public:
    StaticMap(const Param& p, const Mission& m, const std::vector<std::array<double, 6>>& boxes, double max_dist = 1.0) {
        lscqp_config c{};
        c.M = p.M; c.n = p.n; c.phi = p.phi; c.dim = p.world_dimension; c.dt = p.dt; c.w_control = p.control_input_weight;
        c.w_terminal = p.terminal_weight; c.planner_mode = (int) p.planner_mode <= 2 ? (int) p.planner_mode : 0;
        for (int k = 0; k < 3; k++) { c.world_min[k] = (double) m.world_min(k); c.world_max[k] = (double) m.world_max(k); }
        c.z_2d = p.world_z_2d; c.max_obs = 40; c.presolve = 1;
        if (lscqp_create(&c, 0, &handle) != 0) throw std::runtime_error(std::string("[StaticMap] ") + lscqp_last_error());
        std::vector<double> flat;
        for (const auto& b : boxes) flat.insert(flat.end(), b.begin(), b.end());
        if (lscqp_map_set(handle, flat.data(), (int) boxes.size(), p.world_resolution, max_dist) != 0) {
            const std::string msg = lscqp_last_error();
            lscqp_destroy(handle);
            throw std::runtime_error("[StaticMap] " + msg);
        }
    }
    ~StaticMap() { lscqp_destroy(handle); }
    StaticMap(const StaticMap&) = delete;
    StaticMap& operator=(const StaticMap&) = delete;
    lscqp_handle* handle = nullptr;
};

class CollisionConstraints {
public:
    CollisionConstraints(const Param& p, const Mission&) : param(p) { sfcs.resize(p.M); }
    void setDistmap(std::shared_ptr<StaticMap> map) { distmap_ptr = std::move(map); }    // collision_constraints.hpp:136

    // ---- Safe Flight Corridors (src/collision_constraints.cpp:366-436), grown by lscqp_sfc_host against the static map
    void initializeSFC(const point3d& agent_position, double agent_radius) {            // :366-383
        if (sfc_call(LSCQP_SFC_INIT, agent_position, point3d(), point3d(), agent_radius) == 0)
            throw std::invalid_argument("[CollisionConstraints] Invalid initial SFC");
    }
    void constructSFCFromPoint(const point3d& point, const point3d& goal_point, double agent_radius) {      // :396-411
        sfc_call(LSCQP_SFC_FROM_POINT, point, goal_point, point3d(), agent_radius);
    }
    void constructSFCFromConvexHull(const points_t& convex_hull, const point3d& next_waypoint, double agent_radius) {   // :413-436
        if (convex_hull.size() != 2) throw std::invalid_argument("[lscqp] the convex hull of generateSFC holds two points (traj_planner.cpp:745-747)");
        sfc_call(LSCQP_SFC_FROM_HULL, convex_hull[0], convex_hull[1], next_waypoint, agent_radius);
    }
    void initializeLSC(size_t N_obs) {                                                  // collision_constraints.cpp:385-394
        lscs.assign(N_obs, std::vector<LSCs>(param.M, LSCs(param.n + 1)));
    }
    LSC getLSC(int oi, int m, int i) const { return lscs[oi][m][i]; }                   // :482-484
    Box getSFC(int m) const { return sfcs[m]; }                                         // :486-488
    size_t getObsSize() const { return lscs.size(); }                                   // :490-492
    bool isDynamicObstacle(int oi) const { return dynamic_obstacle_indices.count(oi) != 0; }   // :498-500
    void setLSC(int oi, int m, const points_t& p, const vector3d& n, const std::vector<double>& ds) {   // :514-521
        for (int i = 0; i < param.n + 1; i++) lscs[oi][m][i] = LSC(p[i], n, ds[i]);
    }
    void setLSC(int oi, int m, const points_t& p, const vector3d& n, double d) {        // :523-530
        for (int i = 0; i < param.n + 1; i++) lscs[oi][m][i] = LSC(p[i], n, d);
    }
    void setLSC(int oi, int m, const point3d& p, const vector3d& n, double d) {         // :532-539
        for (int i = 0; i < param.n + 1; i++) lscs[oi][m][i] = LSC(p, n, d);
    }
    void setSFC(int m, const Box& b) { sfcs[m] = b; }                                   // :541-543
    int last_sfc_status = 0;                   // lscqp_sfc_batch status of the last corridor update (0: previous corridor reused)

    // Packed planes for lscqp_solve_*: normal[oi][m] and rhs = n.p + d (the constant of traj_optimizer.cpp:413-429).
    // The reference's generators write one normal per (obstacle, segment); anything else is rejected.
    void pack(int dim, std::vector<double>& normals, std::vector<double>& rhs) const {
        const int M = param.M, N = param.n + 1;
        normals.assign(lscs.size() * M * 3, 0.0); rhs.assign(lscs.size() * M * N, 0.0);
        for (size_t oi = 0; oi < lscs.size(); oi++)
            for (int m = 0; m < M; m++) {
                const LSC& f = lscs[oi][m][0];
                for (int k = 0; k < 3; k++) normals[(oi * M + m) * 3 + k] = (double) f.normal_vector(k);
                for (int i = 0; i < N; i++) {
                    const LSC& l = lscs[oi][m][i];
                    if (!(l.normal_vector == f.normal_vector))
                        throw std::invalid_argument("[lscqp] LSC normals differ inside one (obstacle, segment)");
                    double b = l.d;
                    for (int k = 0; k < dim; k++) b += (double) l.normal_vector(k) * (double) l.obs_control_point(k);
                    rhs[(oi * M + m) * N + i] = b;
                }
            }
    }
private:
    int sfc_call(int mode, const point3d& point, const point3d& goal, const point3d& wp, double radius) {
        if (!distmap_ptr) throw std::runtime_error("[CollisionConstraints] setDistmap has not been called");
        const int M = param.M;
        std::vector<float> boxes((size_t) M * 6);
        for (int m = 0; m < M; m++)
            for (int k = 0; k < 3; k++) { boxes[m * 6 + k] = sfcs[m].box_min(k); boxes[m * 6 + 3 + k] = sfcs[m].box_max(k); }
        const float pt[3] = {point(0), point(1), point(2)}, g[3] = {goal(0), goal(1), goal(2)}, w[3] = {wp(0), wp(1), wp(2)};
        const double limits[8] = {0, 0, 0, 0, 0, 0, radius, 0};
        int status = 0;
        if (lscqp_sfc_host(distmap_ptr->handle, mode, 1, pt, g, w, limits, boxes.data(), &status) != 0)
            throw std::runtime_error(std::string("[CollisionConstraints] ") + lscqp_last_error());
        for (int m = 0; m < M; m++)
            for (int k = 0; k < 3; k++) { sfcs[m].box_min(k) = boxes[m * 6 + k]; sfcs[m].box_max(k) = boxes[m * 6 + 3 + k]; }
        last_sfc_status = status;
        return status;
    }
    std::shared_ptr<StaticMap> distmap_ptr;
    Param param;
    RSFCs lscs;
    SFCs sfcs;
    std::set<int> dynamic_obstacle_indices;     // never populated by the reference
};

// ---- TrajOptimizer (include/traj_optimizer.hpp:18-55) -----------------------------------------
struct TrajOptResult { traj_t desired_traj; double total_qp_cost = 0; };                // traj_optimizer.hpp:18-21

inline lscqp_config make_lscqp_config(const Param& p, const Mission& m, int max_obs = 40) {
    lscqp_config c{};
    c.M = p.M; c.n = p.n; c.phi = p.phi; c.dim = p.world_dimension;
    c.dt = p.dt; c.w_control = p.control_input_weight; c.w_terminal = p.terminal_weight;
    c.planner_mode = (int) p.planner_mode; c.use_sfc = p.world_use_octomap ? 1 : 0;
    c.comm_range = p.communication_range;      // rows of traj_optimizer.cpp:477-500, every planner mode
    // SlackMode (sp_const.hpp:43-47): CONTINUITY adds neither variables nor rows in populatebyrow (it only moves the column
    // offset of slacks that do not exist, traj_optimizer.cpp:227-231); COLLISIONCONSTRAINT is what mode reciprocal_rsfc
    // selects (param.cpp:157-161) and is handled there; any other combination is not a model the reference builds
    if (p.slack_mode == SlackMode::COLLISIONCONSTRAINT && p.planner_mode != PlannerMode::RECIPROCALRSFC)
        throw std::invalid_argument("[lscqp] SlackMode::COLLISIONCONSTRAINT is only built for PlannerMode::RECIPROCALRSFC");
    for (int k = 0; k < 3; k++) { c.world_min[k] = (double) m.world_min(k); c.world_max[k] = (double) m.world_max(k); }
    c.z_2d = p.world_z_2d; c.max_obs = max_obs; c.max_agents = 1; c.max_iter = 0; c.tol = 0; c.presolve = 1;
    return c;
}

class TrajOptimizer {
public:
    // third argument: the Bernstein basis B the reference passes (traj_planner.cpp:27); the constants are rebuilt on the device side
    template <class Matrix>
    TrajOptimizer(const Param& p, const Mission& m, const Matrix&) : param(p), mission(m) { open(); }
    TrajOptimizer(const Param& p, const Mission& m) : param(p), mission(m) { open(); }
    ~TrajOptimizer() { lscqp_destroy(handle); }
    TrajOptimizer(const TrajOptimizer&) = delete;
    TrajOptimizer& operator=(const TrajOptimizer&) = delete;

    // TrajOptimizer::solve, src/traj_optimizer.cpp:18-156.  Failure = throw PlanningReport::QPFAILED (an enum by value,
    // :143,152), which TrajPlanner::trajOptimization catches with catch(...) and answers with initial_traj (:767-797).
    TrajOptResult solve(const Agent& agent, const CollisionConstraints& constraints, const traj_t& initial_traj,
                        bool /*use_primal_algorithm*/) {
        const int M = param.M, N = param.n + 1, D = param.world_dimension;
        float state[9], goal[3], wp[3];
        for (int k = 0; k < 3; k++) {
            state[k] = agent.current_state.position(k); state[3 + k] = agent.current_state.velocity(k);
            state[6 + k] = agent.current_state.acceleration(k); goal[k] = agent.current_goal_point(k);
            wp[k] = agent.next_waypoint(k);
        }
        double limits[8] = {agent.max_vel[0], agent.max_vel[1], agent.max_vel[2], agent.max_acc[0], agent.max_acc[1],
                            agent.max_acc[2], agent.radius, agent.nominal_velocity};
        std::vector<double> normals, rhs;
        constraints.pack(D, normals, rhs);
        std::vector<float> sfc(M * 6), warm(M * N * 3);
        for (int m = 0; m < M; m++)
            for (int k = 0; k < 3; k++) { sfc[m * 6 + k] = constraints.getSFC(m).box_min(k); sfc[m * 6 + 3 + k] = constraints.getSFC(m).box_max(k); }
        const bool have_warm = initial_traj.size() == M;
        if (have_warm)
            for (int m = 0; m < M; m++) for (int i = 0; i < N; i++) for (int k = 0; k < 3; k++) warm[(m * N + i) * 3 + k] = initial_traj[m][i](k);
        int offsets[2] = {0, (int) constraints.getObsSize()};
        std::vector<double> ctrl(D * M * N);
        double cost = 0; int status = 0;
        int rc = lscqp_solve_host(handle, 1, state, goal, limits, param.world_use_octomap ? sfc.data() : nullptr, wp, offsets,
                                  normals.data(), rhs.data(), have_warm ? warm.data() : nullptr, ctrl.data(), &cost, &status,
                                  nullptr, nullptr, nullptr);
        if (rc != 0 || status != LSCQP_OK) throw PlanningReport::QPFAILED;
        TrajOptResult result;
        result.desired_traj = traj_t(M, param.n, param.dt);
        for (int m = 0; m < M; m++)
            for (int i = 0; i < N; i++)                                                 // :71-83: narrowed to float, z := world_z_2d in 2-D
                result.desired_traj[m][i] = point3d((float) ctrl[0 * M * N + m * N + i], (float) ctrl[1 * M * N + m * N + i],
                                                    D == 3 ? (float) ctrl[2 * M * N + m * N + i] : (float) param.world_z_2d);
        result.total_qp_cost = cost;                                                    // :100
        return result;
    }

    void updateParam(const Param& p) { param = p; lscqp_destroy(handle); open(); }      // traj_optimizer.cpp:158-160

private:
    void open() {
        if (!(param.n == 5 && param.phi == 3))                                          // traj_optimizer.cpp:198-201
            throw std::invalid_argument("[TrajOptimizer] Currently, only n=5, phi=3 is available");
        lscqp_config c = make_lscqp_config(param, mission);
        if (lscqp_create(&c, 0, &handle) != 0) throw std::runtime_error(std::string("[TrajOptimizer] ") + lscqp_last_error());
    }
    Param param;
    Mission mission;
    lscqp_handle* handle = nullptr;
};

// ---- GoalOptimizer (include/goal_optimizer.hpp:18-38, src/goal_optimizer.cpp:7-165) -------------
// the second CPLEX call of a replan (TrajPlanner::goalPlanningWithGridBasedPlanner, traj_planner.cpp:545-550)
class GoalOptimizer {
public:
    GoalOptimizer(const Param& p, const Mission& m) : param(p), mission(m) {
        lscqp_config c = make_lscqp_config(param, mission);
        c.comm_range = 0.0;                                                             // the goal LP has no comm rows
        if (lscqp_create(&c, 0, &handle) != 0) throw std::runtime_error(std::string("[GoalOptimizer] ") + lscqp_last_error());
    }
    ~GoalOptimizer() { lscqp_destroy(handle); }
    GoalOptimizer(const GoalOptimizer&) = delete;
    GoalOptimizer& operator=(const GoalOptimizer&) = delete;

    // returns the new current_goal_point; infeasible LP = throw PlanningReport::QPFAILED (goal_optimizer.cpp:94, 103)
    point3d solve(const Agent& /*agent*/, const CollisionConstraints& constraints, const point3d& current_goal_point,
                  const point3d& next_waypoint) {
        const int M = param.M, D = param.world_dimension;
        float goal[3], wp[3], out[3] = {0, 0, 0};
        for (int k = 0; k < 3; k++) { goal[k] = current_goal_point(k); wp[k] = next_waypoint(k); }
        std::vector<double> normals, rhs;
        constraints.pack(D, normals, rhs);
        std::vector<float> sfc(M * 6);
        for (int m = 0; m < M; m++)
            for (int k = 0; k < 3; k++) { sfc[m * 6 + k] = constraints.getSFC(m).box_min(k); sfc[m * 6 + 3 + k] = constraints.getSFC(m).box_max(k); }
        int offsets[2] = {0, (int) constraints.getObsSize()};
        int status = 0;
        int rc = lscqp_goal_host(handle, 1, goal, wp, param.world_use_octomap ? sfc.data() : nullptr, offsets, normals.data(),
                                 rhs.data(), out, nullptr, &status);
        if (rc != 0 || status != LSCQP_OK) throw PlanningReport::QPFAILED;
        return point3d(out[0], out[1], out[2]);
    }

private:
    Param param;
    Mission mission;
    lscqp_handle* handle = nullptr;
};

// ---- the batched dispatch that replaces `for (qi) agents[qi]->plan()` --------------------------
// (src/multi_sync_simulator.cpp:354-362).  Inputs are the per-agent quantities broadcastMsgs() / getAgent()
// already hold (src/agent_manager.cpp:184-199); obstacles are indices into the same agent arrays.
class BatchTrajOptimizer {
public:
    BatchTrajOptimizer(const Param& p, const Mission& m, int max_obs = 40) : param(p) {
        lscqp_config c = make_lscqp_config(p, m, max_obs);
        if (lscqp_create(&c, 0, &handle) != 0) throw std::runtime_error(std::string("[BatchTrajOptimizer] ") + lscqp_last_error());
    }
    ~BatchTrajOptimizer() { lscqp_destroy(handle); }
    // one call = constructLSC + trajOptimization for every agent; status[a] != 0 -> caller keeps initial_traj[a]
    void plan(int generator, const std::vector<Agent>& agents, const std::vector<traj_t>& initial_trajs,
              const std::vector<std::vector<int>>& neighbours, std::vector<traj_t>& desired, std::vector<int>& status) {
        const int n = (int) agents.size(), M = param.M, N = param.n + 1, D = param.world_dimension;
        std::vector<float> state(n * 9), goal(n * 3), wp(n * 3), own(n * M * N * 3);
        std::vector<double> limits(n * 8), meta(n * 2), ctrl((size_t) n * D * M * N), cost(n);
        std::vector<int> off(n + 1, 0), index, iters(n);
        for (int a = 0; a < n; a++) {
            const Agent& g = agents[a];
            for (int k = 0; k < 3; k++) {
                state[a * 9 + k] = g.current_state.position(k); state[a * 9 + 3 + k] = g.current_state.velocity(k);
                state[a * 9 + 6 + k] = g.current_state.acceleration(k); goal[a * 3 + k] = g.current_goal_point(k);
                wp[a * 3 + k] = g.next_waypoint(k);
                limits[a * 8 + k] = g.max_vel[k]; limits[a * 8 + 3 + k] = g.max_acc[k];
            }
            limits[a * 8 + 6] = g.radius; limits[a * 8 + 7] = g.nominal_velocity; meta[a * 2] = g.radius; meta[a * 2 + 1] = g.downwash;
            for (int m = 0; m < M; m++) for (int i = 0; i < N; i++) for (int k = 0; k < 3; k++) own[((a * M + m) * N + i) * 3 + k] = initial_trajs[a][m][i](k);
            index.insert(index.end(), neighbours[a].begin(), neighbours[a].end());
            off[a + 1] = (int) index.size();
        }
        status.assign(n, 0);
        int rc = lscqp_replan_host(handle, generator, n, state.data(), goal.data(), limits.data(), nullptr, wp.data(), own.data(), meta.data(),
                                   off.data(), index.data(), ctrl.data(), cost.data(), status.data(), iters.data());
        if (rc != 0) throw std::runtime_error(std::string("[BatchTrajOptimizer] ") + lscqp_last_error());
        desired.assign(n, traj_t(M, param.n, param.dt));
        for (int a = 0; a < n; a++)
            for (int m = 0; m < M; m++)
                for (int i = 0; i < N; i++) {
                    const double* x = ctrl.data() + (size_t) a * D * M * N;
                    desired[a][m][i] = status[a] == 0
                        ? point3d((float) x[m * N + i], (float) x[M * N + m * N + i], D == 3 ? (float) x[2 * M * N + m * N + i] : (float) param.world_z_2d)
                        : initial_trajs[a][m][i];                                       // failsafe, traj_planner.cpp:795-797
                }
    }
private:
    Param param;
    lscqp_handle* handle = nullptr;
};

}  // namespace DynamicPlanning
