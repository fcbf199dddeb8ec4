#include <cstdio>
// tests/cuda_emul/emul_pdip.cpp -- runs the unmodified PDIP kernel source on the CPU emulator.
// TEST INFRASTRUCTURE ONLY (see cuda_emul.h).  Built by tests/cuda_emul/build.sh with g++.
#define CUDA_EMUL_IMPL
#define LSCQP_CUDA_EMUL
#include "cuda_emul.h"
#include "../../lsc_dr_planner_b200/csrc/host_common.hpp"
#include "../../lsc_dr_planner_b200/csrc/lsc_assemble.cuh"
#include "../../lsc_dr_planner_b200/csrc/step_kernel.cuh"
#include "../../lsc_dr_planner_b200/csrc/goal_kernel.cuh"
#include "../../lsc_dr_planner_b200/csrc/knn_kernel.cuh"
#include "../../lsc_dr_planner_b200/csrc/sfc_kernel.cuh"

using namespace lscqp;

extern "C" int emul_dual_stride(int M, int D, int comm) { return 40 * M * 6 + D * M * 6 * 6 + (comm ? 2 * D * (M * (M - 1) / 2 + M) : 0); }

static int g_emul_das_solved = -1;
// dual active-set first pass on the emulator (the dispatch of solve_instances.hpp:das_launch); false: not available
template <class I>
static bool emul_das(const lscqp_config* cfg, SolveParams& p, int n_agents, std::vector<int>& klass) {
    if constexpr (I::HAS_DAS) {
        using C = typename I::Full;
        static std::vector<double> dtab;
        dtab = build_das_table<C>(p.Q2, cfg->w_terminal);
        if (dtab.empty()) return false;
        p.das_tab = dtab.data();
        static std::vector<double> ckpt; static int ckpt_count;
        ckpt.assign((size_t) 8 * Das<C, LSCQP_DAS_KPT>::CK_STRIDE, 0.0); ckpt_count = 0;
        if (I::HAS_DAS_BIG && !std::getenv("LSCQP_DAS_NO_CKPT")) { p.das_ckpt = ckpt.data(); p.das_ckpt_count = &ckpt_count; p.das_ckpt_slots = 8; }
        p.klass = klass.data(); p.klass_mode = 1;
        emu::launch(n_agents, 32, Das<C, LSCQP_DAS_KPT>::SMEM_BYTES, [&]() { das_solve_kernel<C, LSCQP_DAS_KPT>(p); });
        if (std::getenv("LSCQP_DAS_DEBUG")) for (int a = 0; a < n_agents; a++) if (klass[a]) fprintf(stderr, "das: agent %d deferred, reason %d\n", a, klass[a]);
        if constexpr (I::HAS_DAS_BIG) {
            p.klass_mode = 3;
            emu::launch(n_agents, 32, Das<C, I::DAS_BIG_KPT>::SMEM_BYTES, [&]() { das_solve_kernel<C, I::DAS_BIG_KPT>(p); });
        }
        g_emul_das_solved = 0;
        for (int a = 0; a < n_agents; a++) g_emul_das_solved += klass[a] == 0;
        p.klass_mode = 2;
        return true;
    }
    return false;
}      // agents the dual active-set pass solved in the last emul_solve_batch (-1: it did not run)
extern "C" int emul_das_solved() { return g_emul_das_solved; }
extern "C" int emul_solve_batch(const lscqp_config* cfg, int n_agents,
                                const float* state, const float* goal, const double* limits, const float* sfc,
                                const float* next_waypoint, const int* obs_offsets, const double* normals, const double* rhs,
                                const float* initial_traj, double* ctrl_out, double* cost_out, int* status_out, int* iters_out,
                                double* kkt_out, double* dual_out) {
    int rc = validate_config(*cfg);
    if (rc) return rc;
    SolveParams p;
    fill_solve_params(*cfg, p);
    p.n_agents = n_agents;
    p.state = state; p.goal = goal; p.limits = limits; p.sfc = sfc; p.next_waypoint = next_waypoint;
    p.obs_offsets = obs_offsets; p.normals = normals; p.rhs = rhs; p.warm_traj = initial_traj;
    p.ctrl_out = ctrl_out; p.cost_out = cost_out; p.status_out = status_out; p.iters_out = iters_out;
    p.kkt_out = kkt_out; p.dual_out = dual_out;
    g_emul_das_solved = -1;
    const bool term = cfg->planner_mode == LSCQP_MODE_LSC;
    const bool comm = cfg->comm_range > 0;
#define X(M_, D_, T_, C_)                                                                   \
    if (cfg->M == M_ && cfg->dim == D_ && term == T_ && comm == C_) {                       \
        using I = Instance<M_, D_, T_, C_>;                                                 \
        using C = typename I::Full;                                                         \
        p.dual_stride = C::DUAL_STRIDE;                                                     \
        static const ProjTable tab = build_projection<C>();                                 \
        p.proj = tab.term.data(); p.proj_len = tab.len;                                     \
        static const ProjTable tabl = build_projection<C>(I::Light::NT);                    \
        p.proj_light = tabl.term.data(); p.proj_len_light = tabl.len;                       \
        if (I::HAS_COMPACT && cfg->max_obs <= I::COMPACT_KMAX) {                            \
            using K = typename I::Compact;                                                  \
            static const ProjTable tabk = build_projection<K>();                            \
            p.proj = tabk.term.data(); p.proj_len = tabk.len;                               \
            emu::launch(n_agents, K::NT, K::SMEM_BYTES, [&]() { pdip_solve_kernel<K>(p); }); \
            return 0;                                                                       \
        }                                                                                   \
        std::vector<int> klass(n_agents + 1, 0);                                            \
        if (!(cfg->presolve & 8) && !std::getenv("LSCQP_DAS_OFF") && emul_das<I>(cfg, p, n_agents, klass)) {   \
        } else if (I::HAS_LIGHT && (cfg->presolve & 1) && !(cfg->presolve & 2)) {           \
            using L = typename I::Light;                                                    \
            p.klass = klass.data(); p.klass_mode = 1;                                       \
            emu::launch(n_agents, L::NT, L::SMEM_BYTES, [&]() { pdip_solve_kernel<L>(p); }); \
            p.klass_mode = 2;                                                               \
        }                                                                                   \
        emu::launch(n_agents, C::NT, C::SMEM_BYTES, [&]() { pdip_solve_kernel<C>(p); });    \
        return 0;                                                                           \
    }
    LSCQP_FOR_EACH_INSTANCE(X)
#undef X
    return LSCQP_E_INVALID;
}

static const double* g_emul_obs_size = nullptr;
extern "C" void emul_set_obstacle_sizes(const double* s) { g_emul_obs_size = s; }
extern "C" int emul_assemble_lsc_batch(const lscqp_config* cfg, int generator, int n_agents, const float* own_traj,
                                       const double* agent_meta, const float* agent_goal, const int* obs_offsets,
                                       const float* obs_traj, const float* obs_meta, const float* obs_goal,
                                       const float* obs_position, double* normals_out, double* rhs_out) {
    AssembleParams p{};
    p.obs_size = g_emul_obs_size;
    p.n_agents = n_agents; p.generator = generator; p.dim = cfg->dim;
    p.own_traj = own_traj; p.agent_meta = agent_meta; p.agent_goal = agent_goal; p.obs_offsets = obs_offsets;
    p.obs_traj = obs_traj; p.obs_meta = obs_meta; p.obs_goal = obs_goal; p.obs_position = obs_position;
    p.normals = normals_out; p.rhs = rhs_out;
    if (cfg->M == 5) emu::launch(n_agents, 128, 4096, [&]() { lsc_assemble_kernel<5>(p); });
    else if (cfg->M == 10) emu::launch(n_agents, 128, 4096, [&]() { lsc_assemble_kernel<10>(p); });
    else return LSCQP_E_INVALID;
    return 0;
}

extern "C" int emul_assemble_lsc_fused(const lscqp_config* cfg, int generator, int prune, int n_agents, const float* own_traj,
                                       const double* agent_meta, const float* agent_goal, const float* state, const double* limits,
                                       const int* obs_offsets, const int* obs_index, const float* all_traj, const double* all_meta,
                                       const float* all_goal, const float* all_state, double* normals_out, double* rhs_out) {
    AssembleParams p{};
    p.n_agents = n_agents; p.generator = generator; p.dim = cfg->dim;
    p.own_traj = own_traj; p.agent_meta = agent_meta; p.agent_goal = agent_goal; p.obs_offsets = obs_offsets;
    p.obs_index = obs_index; p.all_traj = all_traj; p.all_meta = all_meta; p.all_goal = all_goal; p.all_state = all_state;
    p.prune = prune; p.state = state; p.limits = limits; p.dt = cfg->dt;
    p.normals = normals_out; p.rhs = rhs_out;
    // prune > 1: the split dispatch (prune kernel + global work list + one thread per surviving pair)
    std::vector<int2> work((size_t) obs_offsets[n_agents] * cfg->M + 1);
    int count = 0;
    const bool split = prune > 1 && cfg->dim == 3 && generator < 2;      // (the condition of lscqp_assemble_lsc_fused)
    if (prune > 1) p.prune = 1;
    if (split) { p.work_list = work.data(); p.work_count = &count; }
    if (cfg->M != 5 && cfg->M != 10) return LSCQP_E_INVALID;
    if (split) {
        if (cfg->M == 5) emu::launch(n_agents, 128, 64, [&]() { lsc_prune_kernel<5>(p); });
        else emu::launch(n_agents, 128, 64, [&]() { lsc_prune_kernel<10>(p); });
    } else {
        if (cfg->M == 5) emu::launch(n_agents, 128, 4096, [&]() { lsc_assemble_kernel<5>(p); });
        else emu::launch(n_agents, 128, 4096, [&]() { lsc_assemble_kernel<10>(p); });
    }
    if (split) {
        if (cfg->M == 5) emu::launch(3, 128, 64, [&]() { lsc_pairs_kernel<5>(p); });
        else emu::launch(3, 128, 64, [&]() { lsc_pairs_kernel<10>(p); });
    }
    return 0;
}

static void emul_launch_step(const lscqp_config* cfg, const StepParams& p) {
    const int blocks = (p.n_agents + STEP_WARPS - 1) / STEP_WARPS;
    if (cfg->M == 5) emu::launch(blocks, STEP_WARPS * 32, STEP_WARPS * 5 * 18 * 4 + 64, [&]() { step_kernel<5>(p); });
    else emu::launch(blocks, STEP_WARPS * 32, STEP_WARPS * 10 * 18 * 4 + 64, [&]() { step_kernel<10>(p); });
}

extern "C" int emul_step_batch(const lscqp_config* cfg, int n_agents, const double* ctrl, double step, float* traj_out,
                               float* state_out, float* shifted_out, const int* status, const float* fallback) {
    if (cfg->M != 5 && cfg->M != 10) return LSCQP_E_INVALID;
    StepParams p;
    p.n_agents = n_agents; p.dim = cfg->dim; p.dt = cfg->dt; p.step = step; p.z_2d = cfg->z_2d;
    p.ctrl = ctrl; p.traj_out = traj_out; p.state_out = state_out; p.shifted_out = shifted_out;
    p.status = status; p.fallback = fallback; p.peers = nullptr; p.lo = 0;
    emul_launch_step(cfg, p);
    return 0;
}

// The peer exchange on the emulator: `world` ranks simulated in one process (their blocks are plain host arrays), one
// closed-loop exchange step = every rank publishes its shard with step_kernel, then every rank runs exchange_begin.
// blocks: world pointers to zero-initialised buffers of emul_exchange_bytes(); ctrl / status / fallback: [n_total] rows;
// traj / state: [world][n_total] replicated arrays (outputs).
extern "C" long emul_exchange_bytes(int n_total, int M) { return 64 + 8 * EXCHANGE_MAX_WORLD + (long) 2 * n_total * (M * 18 + 9) * 4; }
extern "C" int emul_exchange_step(const lscqp_config* cfg, int world, int n_total, void* const* blocks, const double* ctrl,
                                  const int* status, const float* fallback, double step, float* traj, float* state,
                                  int skip_rank) {
    const int M = cfg->M, row = M * 18 + 9;
    std::vector<ExchangePeers> peers(world);
    for (int r = 0; r < world; r++) {
        peers[r].world = world; peers[r].rank = r; peers[r].n_total = n_total; peers[r].row = row;
        for (int q = 0; q < world; q++) {
            char* c = static_cast<char*>(blocks[q]);
            peers[r].blk[q].ctl = reinterpret_cast<unsigned long long*>(c);
            peers[r].blk[q].flags = reinterpret_cast<unsigned long long*>(c + 64);
            peers[r].blk[q].inbox = reinterpret_cast<float*>(c + 64 + 8 * EXCHANGE_MAX_WORLD);
        }
    }
    const int per = (n_total + world - 1) / world;
    for (int r = 0; r < world; r++) {
        if (r == skip_rank) continue;                       // (a rank that never publishes: the others must time out)
        const int lo = std::min(n_total, r * per), hi = std::min(n_total, lo + per);
        StepParams p;
        p.n_agents = hi - lo; p.dim = cfg->dim; p.dt = cfg->dt; p.step = step; p.z_2d = cfg->z_2d;
        p.ctrl = ctrl + (size_t) lo * cfg->dim * M * 6; p.traj_out = nullptr; p.state_out = nullptr; p.shifted_out = nullptr;
        p.status = status ? status + lo : nullptr; p.fallback = fallback ? fallback + (size_t) lo * M * 18 : nullptr;
        p.peers = &peers[r]; p.lo = lo;
        emul_launch_step(cfg, p);
    }
    for (int r = 0; r < world; r++) {
        if (r == skip_rank) continue;
        ExchangeBeginParams b;
        b.peers = &peers[r]; b.traj = traj + (size_t) r * n_total * M * 18; b.state = state + (size_t) r * n_total * 9;
        b.timeout_cycles = 0;
        emu::launch(3, 256, 64, [&]() { exchange_begin_kernel(b); });
    }
    return 0;
}

extern "C" int emul_goal_batch(const lscqp_config* cfg, int n_agents, const float* goal, const float* waypoint, const float* sfc,
                               const int* obs_offsets, const double* normals, const double* rhs, float* goal_out,
                               double* t_out, int* status_out) {
    GoalParams p;
    p.n_agents = n_agents; p.M = cfg->M; p.dim = cfg->dim; p.use_sfc = cfg->use_sfc && sfc; p.feas_tol = 1e-6;
    p.goal = goal; p.waypoint = waypoint; p.sfc = sfc; p.obs_offsets = obs_offsets; p.normals = normals; p.rhs = rhs;
    p.goal_out = goal_out; p.t_out = t_out; p.status_out = status_out;
    emu::launch((n_agents + 3) / 4, 128, 64, [&]() { goal_lp_kernel(p); });
    return 0;
}

extern "C" int emul_select_neighbours(int n_total, int lo, int n_local, int K, double comm_range, const float* state,
                                      int* offsets_out, int* index_out, int* overflow_out) {
    std::vector<int> rows((size_t) n_local * (K > 0 ? K : 1), -1), count(n_local, 0);
    KnnParams p;
    p.n_total = n_total; p.lo = lo; p.n_local = n_local; p.K = K; p.comm_range = comm_range; p.state = state;
    p.obs_index = rows.data(); p.count = count.data(); p.overflow = overflow_out;
    emu::launch(n_local, KNN_THREADS, knn_smem_bytes(n_total) + 64, [&]() { knn_select_kernel(p); });
    KnnCsrParams c;
    c.n_local = n_local; c.K = K; c.rows = rows.data(); c.count = count.data(); c.obs_offsets = offsets_out; c.obs_index = index_out;
    emu::launch(1, KNN_CSR_THREADS, 1024, [&]() { knn_csr_kernel(c); });
    return 0;
}

extern "C" int emul_validate_batch(const lscqp_config* cfg, int n_agents, const float* traj, const float* state, const double* limits,
                                   const float* sfc, int* valid_out) {
    ValidateParams p;
    p.n_agents = n_agents; p.M = cfg->M; p.dim = cfg->dim; p.use_sfc = cfg->use_sfc && sfc;
    p.traj = traj; p.state = state; p.limits = limits; p.sfc = sfc; p.valid_out = valid_out;
    emu::launch((n_agents + 127) / 128, 128, 64, [&]() { validate_kernel(p); });
    return 0;
}

extern "C" void emul_jerk_gram(int n, int phi, double dt, double* Q) { jerk_gram(n, phi, dt, Q); }

// Safe Flight Corridors on the emulator: the map kernels + sfc_kernel on host arrays.
struct EmulMap { MapView m; std::vector<unsigned char> occ; std::vector<int> closest; };
extern "C" void* emul_map_build(const lscqp_config* cfg, const double* boxes, int n_boxes, double res, double max_dist) {
    EmulMap* e = new EmulMap();
    MapView& m = e->m;
    m.res = res; m.inv_res = 1.0 / res;
    size_t cells = 1;
    for (int k = 0; k < 3; k++) {
        m.world_min[k] = (float) cfg->world_min[k]; m.world_max[k] = (float) cfg->world_max[k];
        m.key0[k] = (int) std::floor(m.inv_res * (double) m.world_min[k]);
        m.n[k] = (int) std::floor(m.inv_res * (double) m.world_max[k]) - m.key0[k] + 1;
        cells *= (size_t) m.n[k];
    }
    m.maxd2 = (int) std::pow(max_dist / res, 2);
    e->occ.assign(cells, 0); e->closest.assign(cells, -1);
    m.occ = e->occ.data(); m.closest = e->closest.data();
    if (n_boxes > 0) {
        OccParams op{m, e->occ.data(), boxes, n_boxes};
        emu::launch(n_boxes, 256, 64, [&]() { occupancy_kernel(op); });
    }
    return e;
}
// (the brute-force window scan of every cell is too slow on the fiber emulator: only the listed cells are computed)
extern "C" int emul_map_edt_cells(void* map, const long* cells, int n) {
    EmulMap* e = static_cast<EmulMap*>(map);
    // run edt_closest_kernel's body for single cells: one emulated CTA of 128 threads covers 128 consecutive cells
    for (int i = 0; i < n; i++) {
        const long blk = cells[i] / 128;
        EdtParams ep{e->m, e->closest.data()};
        emu::State& s = emu::st(); (void) s;
        // launch a 1-CTA grid whose blockIdx is forced through an offset copy of the params: emulate by temporary remap
        struct Local { static void run(const EdtParams& p, long blk) {
            blockIdx.x = (unsigned) blk; edt_closest_kernel(p); } };
        emu::launch(1, 128, 64, [&]() { Local::run(ep, blk); });
    }
    return 0;
}
extern "C" const unsigned char* emul_map_occ(void* map, int* n3) {
    EmulMap* e = static_cast<EmulMap*>(map);
    for (int k = 0; k < 3; k++) n3[k] = e->m.n[k];
    return e->occ.data();
}
extern "C" int* emul_map_closest(void* map) { return static_cast<EmulMap*>(map)->closest.data(); }
extern "C" void emul_map_free(void* map) { delete static_cast<EmulMap*>(map); }
extern "C" int emul_sfc_batch(void* map, int mode, int M, int n_agents, const float* point, const float* goal, const float* waypoint,
                              const double* limits, float* sfc, int* status) {
    EmulMap* e = static_cast<EmulMap*>(map);
    SfcParams p;
    p.map = e->m; p.mode = mode; p.n_agents = n_agents; p.M = M;
    p.point = point; p.goal = goal; p.waypoint = waypoint; p.limits = limits; p.sfc = sfc; p.status = status;
    emu::launch(n_agents, SFC_THREADS, 64, [&]() { sfc_kernel(p); });
    return 0;
}
