// external_types_stub.hpp -- test-only stand-ins shaped like the types a lsc_dr_planner build provides (octomap::point3d,
// sp_const.hpp's enums / State / Agent, param.hpp's Param, mission.hpp's Mission, trajectory.hpp's Trajectory), so that
// include/lscqp_shim.hpp can be compiled in its LSCQP_SHIM_EXTERNAL_TYPES mode -- the mode a maintainer uses inside the
// reference tree, where the shim must not define these names itself.  Written for this test; no reference source is
// included or copied (the member lists follow include/param.hpp:8-90, include/sp_const.hpp:19-160).
#pragma once
#include <string>
#include <vector>

namespace octomap_stub {
class point3d {                                   // octomath::Vector3: three floats, x() / y() / z() / operator()(i)
public:
    point3d() : d{0, 0, 0} {}
    point3d(float x, float y, float z) : d{x, y, z} {}
    float& x() { return d[0]; } float& y() { return d[1]; } float& z() { return d[2]; }
    const float& x() const { return d[0]; } const float& y() const { return d[1]; } const float& z() const { return d[2]; }
    float& operator()(unsigned i) { return d[i]; }
    const float& operator()(unsigned i) const { return d[i]; }
    bool operator==(const point3d& o) const { return d[0] == o.d[0] && d[1] == o.d[1] && d[2] == o.d[2]; }
private:
    float d[3];
};
}  // namespace octomap_stub

namespace DynamicPlanning {
typedef octomap_stub::point3d point3d;
typedef octomap_stub::point3d vector3d;
typedef std::vector<point3d> points_t;

enum class PlannerMode { DLSC, LSC, BVC, ORCA, RECIPROCALRSFC, CIRCLETEST };
enum class PredictionMode { POSITION, VELOCITY, PREVIOUSSOLUTION, ORCA };
enum class InitialTrajMode { GREEDY, PREVIOUSSOLUTION, SKIP, ORCA };
enum class SlackMode { NONE, CONTINUITY, COLLISIONCONSTRAINT };
enum class GoalMode { STATIC, ORCA, RIGHTHAND, PRIORBASED, GRIDBASEDPLANNER };
enum class MAPFMode { PIBT, ECBS };
enum PlanningReport { Initialized, INITTRAJGENERATIONFAILED, CONSTRAINTGENERATIONFAILED, QPFAILED, WAITFORROSMSG, SUCCESS };

struct State { point3d position, velocity, acceleration; };
struct Agent {
    int id, cid;
    State current_state;
    point3d start_point, desired_goal_point, current_goal_point, next_waypoint;
    std::vector<double> max_vel, max_acc;
    double radius, downwash, nominal_velocity;
    bool collision_alert;
};

class Param {
public:
    bool log_solver = false, log_vis = false;
    std::string package_path, world_frame_id;
    int world_dimension = 3;
    bool world_use_octomap = false;
    double world_resolution = 0.1, world_z_2d = 1.0;
    bool world_use_global_map = true;
    double world_max_dist = 1.0;
    bool multisim_patrol = false;
    int multisim_qn = 0;
    double multisim_time_step = 0.2;
    int multisim_planning_rate = -1;
    PlannerMode planner_mode = PlannerMode::DLSC;
    PredictionMode prediction_mode = PredictionMode::PREVIOUSSOLUTION;
    InitialTrajMode initial_traj_mode = InitialTrajMode::PREVIOUSSOLUTION;
    SlackMode slack_mode = SlackMode::NONE;
    GoalMode goal_mode = GoalMode::GRIDBASEDPLANNER;
    MAPFMode mapf_mode = MAPFMode::PIBT;
    double dt = 0.2;
    int M = 5, n = 5, phi = 3, phi_n = 1;
    double control_input_weight = 0.01, terminal_weight = 1.0, slack_collision_weight = 1.0, slack_dynamic_weight = 1.0;
    double communication_range = 3.0;
    double grid_resolution = 0.5, grid_margin = 0.1, goal_threshold = 0.1;
};

class Mission {
public:
    size_t qn = 0, on = 0;
    std::vector<Agent> agents;
    point3d world_min{-5, -5, 0}, world_max{5, 5, 2.5};
};

template <typename T> struct Segment {
    std::vector<T> control_points;
    double segment_time = 0;
    T operator[](int i) const { return control_points[i]; }
    T& operator[](int i) { return control_points[i]; }
};
template <typename T> class Trajectory {
public:
    Trajectory() = default;
    Trajectory(size_t M, size_t n, double dt) : segments(M) { for (auto& s : segments) { s.control_points.resize(n + 1); s.segment_time = dt; } }
    int size() const { return (int) segments.size(); }
    bool empty() const { return segments.empty(); }
    Segment<T> operator[](int i) const { return segments[i]; }
    Segment<T>& operator[](int i) { return segments[i]; }
private:
    std::vector<Segment<T>> segments;
};
typedef Trajectory<point3d> traj_t;
}  // namespace DynamicPlanning
