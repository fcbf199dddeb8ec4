// tests/shim/shim_smoke.cpp -- exercises include/lscqp_shim.hpp the way traj_planner.cpp uses the reference classes:
// CollisionConstraints::initializeLSC / setLSC, TrajOptimizer::solve, the QPFAILED throw, GoalOptimizer::solve,
// BatchTrajOptimizer::plan.
// Prints the solution so the pytest wrapper can compare it with the Python / oracle path.
#include <cstdio>
#include "../../include/lscqp_shim.hpp"

using namespace DynamicPlanning;

int main(int argc, char** argv) {
    Param param; Mission mission;
    param.control_input_weight = 0.01; param.terminal_weight = 1.0; param.planner_mode = PlannerMode::LSC;
    param.communication_range = 0.0;     // (the comm-range variant is covered by the Python parity tests)
    Agent agent;
    agent.current_state.position = point3d(0, 0, 1);
    agent.current_goal_point = point3d(3, 0, 1);
    traj_t initial(param.M, param.n, param.dt);
    for (int m = 0; m < param.M; m++) for (int i = 0; i <= param.n; i++) initial[m][i] = point3d(0, 0, 1);

    CollisionConstraints constraints(param, mission);
    constraints.initializeLSC(1);
    for (int m = 0; m < param.M; m++) {
        points_t obs(param.n + 1, point3d(1.0f, 0.0f, 1.0f));
        constraints.setLSC(0, m, obs, point3d(-1, 0, 0), std::vector<double>(param.n + 1, 0.4));   // x <= 0.6
    }
    TrajOptimizer opt(param, mission, 0);
    TrajOptResult r = opt.solve(agent, constraints, initial, true);
    std::printf("cost %.12g\n", r.total_qp_cost);
    for (int m = 0; m < param.M; m++)
        for (int i = 0; i <= param.n; i++) std::printf("cp %d %d %.9g %.9g %.9g\n", m, i, r.desired_traj[m][i].x(), r.desired_traj[m][i].y(), r.desired_traj[m][i].z());

    // infeasible model -> PlanningReport::QPFAILED thrown by value, caught with catch(...) by the caller
    constraints.initializeLSC(2);
    for (int m = 0; m < param.M; m++) {
        points_t obs(param.n + 1, point3d(0, 0, 1));
        constraints.setLSC(0, m, obs, point3d(1, 0, 0), 100.0);
        constraints.setLSC(1, m, obs, point3d(-1, 0, 0), 100.0);
    }
    bool thrown = false;
    try { opt.solve(agent, constraints, initial, true); } catch (PlanningReport rep) { thrown = rep == PlanningReport::QPFAILED; }
    std::printf("qpfailed %d\n", (int) thrown);

    // GoalOptimizer::solve as goalPlanningWithGridBasedPlanner calls it (traj_planner.cpp:545-550): one obstacle plane
    // x <= 0.6 on the last control point, previous goal inside (x = 0.2), next waypoint outside (x = 1.0) -> t* = 0.5
    constraints.initializeLSC(1);
    for (int m = 0; m < param.M; m++) {
        points_t obs(param.n + 1, point3d(1.0f, 0.0f, 1.0f));
        constraints.setLSC(0, m, obs, point3d(-1, 0, 0), std::vector<double>(param.n + 1, 0.4));
    }
    GoalOptimizer gopt(param, mission);
    point3d ng = gopt.solve(agent, constraints, point3d(0.2f, 0.5f, 1.0f), point3d(1.0f, 0.0f, 1.0f));
    std::printf("goal %.9g %.9g %.9g\n", ng.x(), ng.y(), ng.z());
    bool gthrown = false;      // previous goal outside as well: no t in [0, 1] satisfies the row
    try { gopt.solve(agent, constraints, point3d(0.8f, 0.5f, 1.0f), point3d(1.0f, 0.0f, 1.0f)); } catch (PlanningReport rep) { gthrown = rep == PlanningReport::QPFAILED; }
    std::printf("goalfailed %d\n", (int) gthrown);

    // batched dispatch: two agents swapping, each the other's neighbour
    std::vector<Agent> agents(2, agent);
    agents[1].current_state.position = point3d(3, 0.2f, 1); agents[1].current_goal_point = point3d(0, 0.2f, 1);
    std::vector<traj_t> inits(2, initial), desired;
    for (int m = 0; m < param.M; m++) for (int i = 0; i <= param.n; i++) inits[1][m][i] = point3d(3, 0.2f, 1);
    std::vector<std::vector<int>> nb = {{1}, {0}};
    std::vector<int> status;
    BatchTrajOptimizer batch(param, mission);
    batch.plan(LSCQP_GEN_LSC, agents, inits, nb, desired, status);
    std::printf("batch %d %d %.9g %.9g\n", status[0], status[1], desired[0][param.M - 1][param.n].x(), desired[1][param.M - 1][param.n].x());
    return 0;
}
