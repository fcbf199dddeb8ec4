// shim_external.cpp -- include/lscqp_shim.hpp in LSCQP_SHIM_EXTERNAL_TYPES mode, against project-provided types (here: the
// stubs of external_types_stub.hpp).  Compiled with -fsyntax-only by tests/test_capi_load.py: every class surface of the
// shim must be valid C++ when Param / Mission / Agent / point3d / Trajectory come from the host project.
#include "external_types_stub.hpp"
#define LSCQP_SHIM_EXTERNAL_TYPES
#include "../../include/lscqp_shim.hpp"

using namespace DynamicPlanning;

int shim_external_surfaces() {
    Param param;
    Mission mission;
    param.planner_mode = PlannerMode::LSC;
    Agent agent{};
    agent.max_vel = {1, 1, 1}; agent.max_acc = {2, 2, 2}; agent.radius = 0.15; agent.downwash = 2.0; agent.nominal_velocity = 1.0;
    CollisionConstraints constraints(param, mission);
    constraints.initializeLSC(1);
    traj_t initial(param.M, param.n, param.dt);
    TrajOptimizer optimizer(param, mission);
    GoalOptimizer goal_optimizer(param, mission);
    BatchTrajOptimizer batch(param, mission);
    (void) optimizer; (void) goal_optimizer; (void) batch; (void) initial; (void) agent;
    return constraints.getObsSize();
}
