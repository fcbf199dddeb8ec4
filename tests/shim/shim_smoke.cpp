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

    // Safe Flight Corridors as TrajPlanner::generateSFC drives them (traj_planner.cpp:738-753): one pillar at (2, 0), the
    // agent at the origin side; then a solve with world_use_octomap so that the corridor bounds the QP
    {
        Param ps = param; ps.world_use_octomap = true;
        auto map = std::make_shared<StaticMap>(ps, mission, std::vector<std::array<double, 6>>{{2.0, 0.0, 1.25, 0.5, 0.5, 2.5}});
        CollisionConstraints cs(ps, mission);
        cs.setDistmap(map);
        cs.initializeSFC(point3d(1.0f, 0.3f, 1.0f), agent.radius);
        Box b0 = cs.getSFC(0), b4 = cs.getSFC(ps.M - 1);
        std::printf("sfc_init %.9g %.9g %.9g %.9g %.9g %.9g same %d\n", b0.box_min.x(), b0.box_min.y(), b0.box_min.z(), b0.box_max.x(),
                    b0.box_max.y(), b0.box_max.z(), (int) (b0.box_min == b4.box_min && b0.box_max == b4.box_max));
        cs.constructSFCFromPoint(point3d(1.2f, 0.35f, 1.0f), point3d(3.0f, 2.0f, 1.0f), agent.radius);
        Box b = cs.getSFC(ps.M - 1);
        std::printf("sfc_point %d %.9g %.9g %.9g %.9g %.9g %.9g\n", cs.last_sfc_status, b.box_min.x(), b.box_min.y(), b.box_min.z(), b.box_max.x(), b.box_max.y(), b.box_max.z());
        cs.constructSFCFromConvexHull(points_t{point3d(1.2f, 0.35f, 1.0f), point3d(1.3f, 0.5f, 1.0f)}, point3d(1.4f, 1.0f, 1.0f), agent.radius);
        b = cs.getSFC(ps.M - 1);
        std::printf("sfc_hull %d %.9g %.9g %.9g %.9g %.9g %.9g\n", cs.last_sfc_status, b.box_min.x(), b.box_min.y(), b.box_min.z(), b.box_max.x(), b.box_max.y(), b.box_max.z());
        bool sthrown = false;                           // a start cell inside the pillar: the reference throws std::invalid_argument
        try { CollisionConstraints bad(ps, mission); bad.setDistmap(map); bad.initializeSFC(point3d(2.0f, 0.0f, 1.0f), agent.radius); }
        catch (const std::invalid_argument&) { sthrown = true; }
        std::printf("sfc_invalid %d\n", (int) sthrown);
        // the corridor of the start box bounds the QP: goal behind the pillar, x stays below the corridor's x_max
        CollisionConstraints cq(ps, mission);
        cq.setDistmap(map);
        cq.initializeSFC(point3d(1.3f, 0.0f, 1.0f), agent.radius);
        cq.initializeLSC(0);
        Agent as = agent; as.current_state.position = point3d(1.3f, 0.0f, 1.0f); as.current_goal_point = point3d(4.0f, 0.0f, 1.0f);
        traj_t init_s(ps.M, ps.n, ps.dt);
        for (int m = 0; m < ps.M; m++) for (int i = 0; i <= ps.n; i++) init_s[m][i] = point3d(1.3f, 0.0f, 1.0f);
        TrajOptimizer opt_s(ps, mission, 0);
        TrajOptResult rs = opt_s.solve(as, cq, init_s, true);
        std::printf("sfc_qp %.9g %.9g\n", rs.desired_traj[ps.M - 1][ps.n].x(), cq.getSFC(0).box_max.x());
    }

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
