// lscqp.cu -- liblscqp.so: the C ABI of include/lscqp.h on top of the sm_100a kernels.
// No CPU path exists here: every entry point launches CUDA kernels or fails with an error code.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "host_common.hpp"
#include "solve_instances.hpp"
#include "lsc_assemble.cuh"
#include "step_kernel.cuh"
#include "goal_kernel.cuh"
#include "knn_kernel.cuh"
#include "sfc_kernel.cuh"

using namespace lscqp;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(LSCQP_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int reserve(size_t n) {
        if (n <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        if (cudaMalloc(&p, n) != cudaSuccess) return -1;
        bytes = n;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

struct Exchange;
struct lscqp_handle {
    lscqp_config cfg;
    int device;
    SolveParams base;
    int dual_stride, kmax, nv;
    cudaStream_t stream;
    // device staging for the *_host entry points
    DevBuf d_state, d_goal, d_limits, d_sfc, d_off, d_normals, d_rhs, d_ctrl, d_cost, d_status, d_iters, d_kkt, d_dual;
    DevBuf d_own, d_ameta, d_index;
    DevBuf d_proj_ent, d_proj_term, d_wp, d_klass, d_gout, d_knn;
    bool two_pass = false, last_two_pass = false;
    bool das = false;                   // dual active-set first pass available and enabled (das_kernel.cuh)
    int das_big_slots = 0;              // batch size up to which the large active-set instance alone is the first pass
    DevBuf d_das, d_ckpt;               // active-set table; checkpoint pool of the instance hand-over (+ its counter)
    size_t knn_smem = 0;
    int two_pass_min = 1536;       // batch size from which the light first pass is used (LSCQP_TWO_PASS_MIN overrides)
    unsigned long long launches = 0;
    DevBuf d_work;                      // global work list of the split LSC assembly
    int asm_split_min = 256;            // batch size from which lscqp_assemble_lsc_fused prunes and enumerates in two kernels
    DevBuf d_occ, d_closest, d_boxes;   // static map (lscqp_map_set)
    MapView map{};
    bool has_map = false;
    const double* obs_size = nullptr;   // lscqp_set_obstacle_sizes (generateReciprocalRSFC)
    Exchange* xchg = nullptr;          // peer exchange of the sharded closed loop (lscqp_exchange_*)
};

extern "C" const char* lscqp_version(void) { return "lscqp-b200 0.2 (sm_100a)"; }
extern "C" const char* lscqp_last_error(void) { return g_err.c_str(); }

extern "C" int lscqp_create(const lscqp_config* cfg, int device, lscqp_handle** out) {
    if (!cfg || !out) return fail(LSCQP_E_INVALID, "null argument");
    int rc = validate_config(*cfg);
    if (rc) return fail(rc, "unsupported configuration (need n=5, phi=3, M in {5,10}, dim in {2,3}, "
                            "mode in {DLSC,LSC,BVC}, max_obs<=40)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(LSCQP_E_NODEVICE, "no CUDA device: liblscqp has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(LSCQP_E_NODEVICE, "device index out of range");
    CK(cudaSetDevice(device));
    lscqp_handle* h = new lscqp_handle();
    h->cfg = *cfg;
    h->device = device;
    fill_solve_params(*cfg, h->base);
    InstanceInfo info;
    int found = inst_query_0(*cfg, &info);
    if (!found) found = inst_query_1(*cfg, &info);
    if (!found) found = inst_query_2(*cfg, &info);
    if (!found) found = inst_query_3(*cfg, &info);
    if (!found) found = inst_query_4(*cfg, &info);
    if (found < 0) { delete h; return fail(LSCQP_E_CUDA, "cudaFuncSetAttribute failed"); }
    if (!found) { delete h; return fail(LSCQP_E_INVALID, "no kernel instance"); }
    h->dual_stride = info.dual_stride; h->kmax = info.kmax; h->nv = info.nv;
    h->two_pass = info.has_light && (cfg->presolve & 1) && !(cfg->presolve & 2);
    if (cfg->presolve & 4) h->two_pass_min = 0;            // light first pass at any batch size
    if (const char* e = std::getenv("LSCQP_TWO_PASS_MIN")) h->two_pass_min = std::atoi(e);
    if (const char* e = std::getenv("LSCQP_ASM_SPLIT_MIN")) h->asm_split_min = std::atoi(e);
    // dual active-set first pass: every banded configuration it is instantiated for, unless switched off (presolve bit 8,
    // LSCQP_DAS=0: the interior-point instances alone, as before)
    h->das = info.has_das && !(cfg->presolve & 8);
    if (const char* e = std::getenv("LSCQP_DAS")) h->das = h->das && std::atoi(e) != 0;
    if (h->das) {
        if (h->d_das.reserve(info.das_tab.size() * sizeof(double)) ||
            cudaMemcpy(h->d_das.p, info.das_tab.data(), info.das_tab.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
            delete h; return fail(LSCQP_E_CUDA, "active-set table upload failed");
        }
        h->base.das_tab = h->d_das.as<double>();
        h->das_big_slots = info.das_big_slots;
        if (info.das_ckpt_stride > 0) {
            const int slots = 64;
            if (h->d_ckpt.reserve(16 + (size_t) slots * info.das_ckpt_stride * sizeof(double))) { delete h; return fail(LSCQP_E_CUDA, "cudaMalloc failed"); }
            h->base.das_ckpt_count = h->d_ckpt.as<int>();
            h->base.das_ckpt = reinterpret_cast<double*>(h->d_ckpt.as<char>() + 16);
            h->base.das_ckpt_slots = slots;
        }
    }
    const ProjTable& tab = info.tab;
    const ProjTable& tabl = info.tab_light;
    if (h->d_proj_ent.reserve(tab.term.size() * sizeof(ProjTerm)) || h->d_proj_term.reserve((tabl.term.size() + 1) * sizeof(ProjTerm)) ||
        cudaMemcpy(h->d_proj_ent.p, tab.term.data(), tab.term.size() * sizeof(ProjTerm), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->d_proj_term.p, tabl.term.data(), tabl.term.size() * sizeof(ProjTerm), cudaMemcpyHostToDevice) != cudaSuccess) {
        delete h; return fail(LSCQP_E_CUDA, "projection table upload failed");
    }
    h->base.proj = h->d_proj_ent.as<ProjTerm>(); h->base.proj_len = tab.len;
    h->base.proj_light = h->d_proj_term.as<ProjTerm>(); h->base.proj_len_light = tabl.len;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h; return fail(LSCQP_E_CUDA, "cudaStreamCreate failed");
    }
    *out = h;
    return 0;
}

extern "C" int lscqp_exchange_destroy(lscqp_handle* h);
extern "C" int lscqp_destroy(lscqp_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    lscqp_exchange_destroy(h);
    DevBuf* bufs[] = {&h->d_state, &h->d_goal, &h->d_limits, &h->d_sfc, &h->d_off, &h->d_normals, &h->d_rhs, &h->d_ctrl,
                      &h->d_cost, &h->d_status, &h->d_iters, &h->d_kkt, &h->d_dual, &h->d_own, &h->d_ameta, &h->d_index,
                      &h->d_proj_ent, &h->d_proj_term, &h->d_wp,
                      &h->d_klass, &h->d_gout, &h->d_knn, &h->d_occ, &h->d_closest, &h->d_boxes, &h->d_work, &h->d_das, &h->d_ckpt};
    for (DevBuf* b : bufs) b->release();
    cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

extern "C" int lscqp_dual_stride(const lscqp_handle* h) { return h ? h->dual_stride : LSCQP_E_INVALID; }
extern "C" int lscqp_max_obs_padded(const lscqp_handle* h) { return h ? h->kmax : LSCQP_E_INVALID; }
extern "C" unsigned long long lscqp_launch_count(const lscqp_handle* h) { return h ? h->launches : 0; }

extern "C" int lscqp_solve_batch(lscqp_handle* h, int n_agents, const float* state, const float* goal,
                                 const double* limits, const float* sfc, const float* next_waypoint, const int* obs_offsets,
                                 const double* normals, const double* rhs, const float* initial_traj, double* ctrl_out,
                                 double* cost_out, int* status_out, int* iters_out, double* kkt_out, double* dual_out,
                                 void* stream) {
    if (!h || n_agents < 0 || !state || !goal || !limits || !obs_offsets || !ctrl_out || !cost_out || !status_out)
        return fail(LSCQP_E_INVALID, "null argument");
    if (h->cfg.use_sfc && !sfc) return fail(LSCQP_E_INVALID, "use_sfc set but sfc is null");
    if (h->cfg.comm_range > 0 && !next_waypoint) return fail(LSCQP_E_INVALID, "comm_range set but next_waypoint is null");
    if (n_agents == 0) return 0;
    SolveParams p = h->base;
    p.n_agents = n_agents;
    p.state = state; p.goal = goal; p.limits = limits; p.sfc = sfc; p.next_waypoint = next_waypoint;
    p.obs_offsets = obs_offsets; p.normals = normals; p.rhs = rhs; p.warm_traj = initial_traj;
    p.ctrl_out = ctrl_out; p.cost_out = cost_out; p.status_out = status_out; p.iters_out = iters_out;
    p.kkt_out = kkt_out; p.dual_out = dual_out; p.dual_stride = h->dual_stride;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // The light first pass pays off in the throughput regime (several waves of one-warp CTAs); a small batch is
    // latency bound and finishes sooner on the 128-thread instance alone.
    // first pass: the dual active-set kernel when available (any batch size: its per-QP latency is below that of the
    // 128-thread interior-point instance), else the light interior-point instance in the throughput regime
    const int first_pass = h->das ? (n_agents <= h->das_big_slots ? 3 : 2) : ((h->two_pass && n_agents >= h->two_pass_min) ? 1 : 0);
    const bool two_pass = first_pass != 0;
    h->last_two_pass = two_pass;
    if (two_pass) {
        if (h->d_klass.reserve((size_t) n_agents * sizeof(int))) return fail(LSCQP_E_CUDA, "cudaMalloc failed");
        p.klass = h->d_klass.as<int>();
        // (routing the remainder of the light pass's last round to the full-capacity pass was measured: 1.65 vs 1.52 ms
        //  per 4096 QPs -- the QPs' durations spread over 8..11 iterations, so the rounds do not end together)
    }
    if (first_pass == 2 && p.das_ckpt) CK(cudaMemsetAsync(p.das_ckpt_count, 0, sizeof(int), st));
    int launched = inst_launch_0(h->cfg, p, n_agents, first_pass, st);
    if (!launched) launched = inst_launch_1(h->cfg, p, n_agents, first_pass, st);
    if (!launched) launched = inst_launch_2(h->cfg, p, n_agents, first_pass, st);
    if (!launched) launched = inst_launch_3(h->cfg, p, n_agents, first_pass, st);
    if (!launched) launched = inst_launch_4(h->cfg, p, n_agents, first_pass, st);
    h->launches += launched;
    CK(cudaGetLastError());
    return 0;
}

// which kernel instance solved each agent in the last two-pass lscqp_solve_batch call (0 = light one-warp instance,
// 1 = full-capacity instance); synchronises the stream.  Returns LSCQP_E_INVALID when the last call was one-pass.
extern "C" int lscqp_last_instances(lscqp_handle* h, int n_agents, int* klass_out_host, void* stream) {
    if (!h || !klass_out_host || n_agents < 0) return fail(LSCQP_E_INVALID, "null argument");
    if (!h->last_two_pass || h->d_klass.bytes < (size_t) n_agents * sizeof(int)) return fail(LSCQP_E_INVALID, "last solve was one-pass");
    CK(cudaMemcpyAsync(klass_out_host, h->d_klass.p, (size_t) n_agents * sizeof(int), cudaMemcpyDeviceToHost, reinterpret_cast<cudaStream_t>(stream)));
    CK(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
    return 0;
}

extern "C" int lscqp_set_obstacle_sizes(lscqp_handle* h, const double* obs_size) {
    if (!h) return fail(LSCQP_E_INVALID, "null handle");
    h->obs_size = obs_size;
    return 0;
}

extern "C" int lscqp_assemble_lsc_batch(lscqp_handle* h, int generator, int n_agents, const float* own_traj,
                                        const double* agent_meta, const float* agent_goal, const int* obs_offsets,
                                        const float* obs_traj, const float* obs_meta, const float* obs_goal,
                                        const float* obs_position, double* normals_out, double* rhs_out, void* stream) {
    if (!h || n_agents < 0 || !own_traj || !agent_meta || !obs_offsets || !obs_traj || !obs_meta || !normals_out || !rhs_out)
        return fail(LSCQP_E_INVALID, "null argument");
    if (generator < 0 || generator > 3) return fail(LSCQP_E_INVALID, "unknown generator");
    if ((generator == LSCQP_GEN_CLSC && (!obs_goal || !agent_goal)) || (generator == LSCQP_GEN_LSC && (!obs_position || !agent_goal)))
        return fail(LSCQP_E_INVALID, "generator needs goal / position arrays");
    if (n_agents == 0) return 0;
    AssembleParams p{};
    p.n_agents = n_agents; p.generator = generator; p.dim = h->cfg.dim;
    p.own_traj = own_traj; p.agent_meta = agent_meta; p.agent_goal = agent_goal; p.obs_offsets = obs_offsets;
    p.obs_traj = obs_traj; p.obs_meta = obs_meta; p.obs_goal = obs_goal; p.obs_position = obs_position;
    p.normals = normals_out; p.rhs = rhs_out; p.obs_size = h->obs_size;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (h->cfg.M == 5) lsc_assemble_kernel<5><<<n_agents, 128, 0, st>>>(p);
    else lsc_assemble_kernel<10><<<n_agents, 128, 0, st>>>(p);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int lscqp_assemble_lsc_fused(lscqp_handle* h, int generator, int prune, int n_agents, const float* own_traj,
                                        const double* agent_meta, const float* agent_goal, const float* state,
                                        const double* limits, const int* obs_offsets, const int* obs_index,
                                        const float* all_traj, const double* all_meta, const float* all_goal,
                                        const float* all_state, double* normals_out, double* rhs_out, void* stream) {
    if (!h || n_agents < 0 || !own_traj || !agent_meta || !agent_goal || !obs_offsets || !obs_index || !all_traj || !all_meta ||
        !all_goal || !all_state || !normals_out || !rhs_out)
        return fail(LSCQP_E_INVALID, "null argument");
    if (generator < 0 || generator > 3) return fail(LSCQP_E_INVALID, "unknown generator");
    if (prune && (!state || !limits)) return fail(LSCQP_E_INVALID, "prune needs state and limits");
    if (n_agents == 0) return 0;
    AssembleParams p{};
    p.n_agents = n_agents; p.generator = generator; p.dim = h->cfg.dim;
    p.own_traj = own_traj; p.agent_meta = agent_meta; p.agent_goal = agent_goal; p.obs_offsets = obs_offsets;
    p.obs_index = obs_index; p.all_traj = all_traj; p.all_meta = all_meta; p.all_goal = all_goal; p.all_state = all_state;
    p.prune = prune ? 1 : 0; p.state = state; p.limits = limits; p.dt = h->cfg.dt;
    p.normals = normals_out; p.rhs = rhs_out; p.obs_size = h->obs_size;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // Pruned path at throughput batch sizes: prune + global work list, then one thread per surviving pair (split dispatch,
    // lsc_assemble.cuh).  The list is sized for max_obs pairs per agent; it must exist before a stream capture starts.
    const bool split = p.prune && h->cfg.dim == 3 && generator < 2 && n_agents >= h->asm_split_min;
    if (split) {
        const size_t cap = (size_t) n_agents * (size_t) h->cfg.max_obs * h->cfg.M;
        if (h->d_work.reserve(cap * sizeof(int2) + 16)) return fail(LSCQP_E_CUDA, "cudaMalloc failed");
        p.work_count = h->d_work.as<int>();
        p.work_list = reinterpret_cast<int2*>(h->d_work.as<char>() + 16);
        CK(cudaMemsetAsync(p.work_count, 0, sizeof(int), st));
    }
    if (split) {
        if (h->cfg.M == 5) lsc_prune_kernel<5><<<n_agents, 128, 0, st>>>(p);
        else lsc_prune_kernel<10><<<n_agents, 128, 0, st>>>(p);
    } else {
        if (h->cfg.M == 5) lsc_assemble_kernel<5><<<n_agents, 128, 0, st>>>(p);
        else lsc_assemble_kernel<10><<<n_agents, 128, 0, st>>>(p);
    }
    h->launches++;
    if (split) {
        const int blocks = 148 * 4;
        if (h->cfg.M == 5) lsc_pairs_kernel<5><<<blocks, 128, 0, st>>>(p);
        else lsc_pairs_kernel<10><<<blocks, 128, 0, st>>>(p);
        h->launches++;
    }
    CK(cudaGetLastError());
    return 0;
}

extern "C" int lscqp_gather_obstacles(lscqp_handle* h, int n_obs, const int* obs_index, const float* own_traj,
                                      const double* agent_meta, const float* agent_goal, const float* state,
                                      float* obs_traj, float* obs_meta, float* obs_goal, float* obs_position, void* stream) {
    if (!h || n_obs < 0 || !obs_index || !own_traj || !agent_meta || !agent_goal || !state || !obs_traj || !obs_meta ||
        !obs_goal || !obs_position)
        return fail(LSCQP_E_INVALID, "null argument");
    if (n_obs == 0) return 0;
    GatherParams p;
    p.n_obs = n_obs; p.M = h->cfg.M; p.obs_index = obs_index; p.own_traj = own_traj; p.agent_meta = agent_meta;
    p.agent_goal = agent_goal; p.state = state; p.obs_traj = obs_traj; p.obs_meta = obs_meta; p.obs_goal = obs_goal;
    p.obs_position = obs_position;
    const size_t total = (size_t) n_obs * h->cfg.M * 18;
    int blocks = (int) ((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    gather_obstacles_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int launch_step(lscqp_handle* h, int n_agents, const double* ctrl, double step, float* traj_out, float* state_out,
                       float* shifted_out, const int* status, const float* fallback, const ExchangePeers* peers, int lo,
                       cudaStream_t st) {
    StepParams p;
    p.n_agents = n_agents; p.dim = h->cfg.dim; p.dt = h->cfg.dt; p.step = step; p.z_2d = h->cfg.z_2d;
    p.ctrl = ctrl; p.traj_out = traj_out; p.state_out = state_out; p.shifted_out = shifted_out;
    p.status = status; p.fallback = fallback; p.peers = peers; p.lo = lo;
    const int blocks = (n_agents + STEP_WARPS - 1) / STEP_WARPS;
    if (h->cfg.M == 5) step_kernel<5><<<blocks, STEP_WARPS * 32, 0, st>>>(p);
    else step_kernel<10><<<blocks, STEP_WARPS * 32, 0, st>>>(p);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int lscqp_step_batch(lscqp_handle* h, int n_agents, const double* ctrl, double step, float* traj_out,
                                float* state_out, float* shifted_traj_out, void* stream) {
    if (!h || n_agents < 0 || !ctrl || !traj_out) return fail(LSCQP_E_INVALID, "null argument");
    if (n_agents == 0) return 0;
    return launch_step(h, n_agents, ctrl, step, traj_out, state_out, shifted_traj_out, nullptr, nullptr, nullptr, 0,
                       reinterpret_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------------------------------------
// Peer exchange of the sharded closed loop (step_kernel.cuh: ExchangeBlock): one block per rank, mapped into every
// peer with CUDA IPC; the step kernel publishes into all of them over NVLink, exchange_begin waits and copies.
struct Exchange {
    ExchangePeers host{};                 // pointers as seen from this process
    ExchangePeers* dev = nullptr;         // device copy read by the kernels
    void* local = nullptr;                // base of the local block (cudaMalloc)
    void* mapped[EXCHANGE_MAX_WORLD] = {};   // IPC mappings of the peers' blocks
    size_t bytes = 0;
};

static void exchange_layout(void* base, int world, int n_total, int row, ExchangeBlock& b) {
    char* c = static_cast<char*>(base);
    b.ctl = reinterpret_cast<unsigned long long*>(c);                       // [4] (+ padding to 64 bytes)
    b.flags = reinterpret_cast<unsigned long long*>(c + 64);                // [EXCHANGE_MAX_WORLD]
    b.inbox = reinterpret_cast<float*>(c + 64 + 8 * EXCHANGE_MAX_WORLD);
    (void) world; (void) n_total; (void) row;
}
static size_t exchange_bytes(int n_total, int row) { return 64 + 8 * EXCHANGE_MAX_WORLD + (size_t) 2 * n_total * row * sizeof(float); }

extern "C" int lscqp_exchange_create(lscqp_handle* h, int n_total, int world, int rank, void* ipc_handle_out) {
    if (!h || n_total <= 0 || world < 1 || world > EXCHANGE_MAX_WORLD || rank < 0 || rank >= world || !ipc_handle_out)
        return fail(LSCQP_E_INVALID, "bad argument");
    if (h->xchg) return fail(LSCQP_E_INVALID, "exchange already created on this handle");
    CK(cudaSetDevice(h->device));
    Exchange* x = new Exchange();
    const int row = h->cfg.M * 18 + 9;
    x->bytes = exchange_bytes(n_total, row);
    if (cudaMalloc(&x->local, x->bytes) != cudaSuccess) { delete x; return fail(LSCQP_E_CUDA, "cudaMalloc failed"); }
    cudaMemset(x->local, 0, x->bytes);
    x->host.world = world; x->host.rank = rank; x->host.n_total = n_total; x->host.row = row;
    exchange_layout(x->local, world, n_total, row, x->host.blk[rank]);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "lscqp.h documents a 64-byte handle");
    cudaIpcMemHandle_t ih;
    std::memset(&ih, 0, sizeof(ih));
    if (world > 1 && cudaIpcGetMemHandle(&ih, x->local) != cudaSuccess) {
        const std::string msg = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(cudaGetLastError());
        cudaFree(x->local); delete x;
        return fail(LSCQP_E_CUDA, msg);
    }
    std::memcpy(ipc_handle_out, &ih, sizeof(ih));
    if (cudaMalloc(&x->dev, sizeof(ExchangePeers)) != cudaSuccess) { cudaFree(x->local); delete x; return fail(LSCQP_E_CUDA, "cudaMalloc failed"); }
    h->xchg = x;
    if (world == 1) CK(cudaMemcpy(x->dev, &x->host, sizeof(ExchangePeers), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int lscqp_exchange_connect(lscqp_handle* h, const void* all_ipc_handles) {
    if (!h || !h->xchg || !all_ipc_handles) return fail(LSCQP_E_INVALID, "exchange not created / null handles");
    Exchange* x = h->xchg;
    CK(cudaSetDevice(h->device));
    const char* hs = static_cast<const char*>(all_ipc_handles);
    for (int r = 0; r < x->host.world; r++) {
        if (r == x->host.rank) continue;
        cudaIpcMemHandle_t ih;
        std::memcpy(&ih, hs + (size_t) r * sizeof(ih), sizeof(ih));
        void* base = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&base, ih, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(LSCQP_E_CUDA, std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
        x->mapped[r] = base;
        exchange_layout(base, x->host.world, x->host.n_total, x->host.row, x->host.blk[r]);
    }
    CK(cudaMemcpy(x->dev, &x->host, sizeof(ExchangePeers), cudaMemcpyHostToDevice));
    return 0;
}

// In-process wiring for tests and single-process multi-GPU drivers: the peers' blocks are given as plain device
// pointers (lscqp_exchange_local_base of the other handles; peer access must be enabled by the caller).
extern "C" void* lscqp_exchange_local_base(lscqp_handle* h) { return (h && h->xchg) ? h->xchg->local : nullptr; }
extern "C" int lscqp_exchange_connect_ptrs(lscqp_handle* h, void* const* bases) {
    if (!h || !h->xchg || !bases) return fail(LSCQP_E_INVALID, "exchange not created / null bases");
    Exchange* x = h->xchg;
    CK(cudaSetDevice(h->device));
    for (int r = 0; r < x->host.world; r++)
        if (r != x->host.rank) exchange_layout(bases[r], x->host.world, x->host.n_total, x->host.row, x->host.blk[r]);
    CK(cudaMemcpy(x->dev, &x->host, sizeof(ExchangePeers), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int lscqp_exchange_begin(lscqp_handle* h, float* traj, float* state, void* stream) {
    if (!h || !h->xchg || !traj || !state) return fail(LSCQP_E_INVALID, "exchange not created / null argument");
    ExchangeBeginParams p;
    p.peers = h->xchg->dev; p.traj = traj; p.state = state;
    p.timeout_cycles = 4000000000ll;                          // ~2 s at 1.9 GHz: a missing peer must not hang the device
    const size_t total = (size_t) h->xchg->host.n_total * h->xchg->host.row;
    int blocks = (int) ((total + 256 * 8 - 1) / (256 * 8));
    if (blocks > 148) blocks = 148;
    if (blocks < 1) blocks = 1;
    exchange_begin_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int lscqp_step_exchange(lscqp_handle* h, int lo, int n_local, const double* ctrl, const int* status,
                                   const float* fallback_traj, double step, float* traj_out, void* stream) {
    if (!h || !h->xchg || n_local <= 0 || lo < 0 || lo + n_local > h->xchg->host.n_total || !ctrl)
        return fail(LSCQP_E_INVALID, "exchange not created / bad argument (every rank must publish at least one agent)");
    if ((status == nullptr) != (fallback_traj == nullptr)) return fail(LSCQP_E_INVALID, "status and fallback_traj go together");
    return launch_step(h, n_local, ctrl, step, traj_out, nullptr, nullptr, status, fallback_traj, h->xchg->dev, lo,
                       reinterpret_cast<cudaStream_t>(stream));
}

// counters of the local block: out[0] = steps published, out[1] = wait time-outs seen, out[2] = failsafe uses; synchronises
extern "C" int lscqp_exchange_status(lscqp_handle* h, unsigned long long* out3, void* stream) {
    if (!h || !h->xchg || !out3) return fail(LSCQP_E_INVALID, "exchange not created / null argument");
    unsigned long long ctl[4];
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CK(cudaMemcpyAsync(ctl, h->xchg->host.blk[h->xchg->host.rank].ctl, sizeof(ctl), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    out3[0] = ctl[0]; out3[1] = ctl[2]; out3[2] = ctl[3];
    return 0;
}

extern "C" int lscqp_exchange_destroy(lscqp_handle* h) {
    if (!h || !h->xchg) return 0;
    Exchange* x = h->xchg;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < EXCHANGE_MAX_WORLD; r++) if (x->mapped[r]) cudaIpcCloseMemHandle(x->mapped[r]);
    if (x->dev) cudaFree(x->dev);
    if (x->local) cudaFree(x->local);
    delete x;
    h->xchg = nullptr;
    return 0;
}

extern "C" int lscqp_select_neighbours(lscqp_handle* h, int n_total, int lo, int n_local, int K, double comm_range,
                                       const float* state, int* obs_offsets_out, int* obs_index_out, int* overflow_out,
                                       void* stream) {
    if (!h || !state || !obs_offsets_out || !obs_index_out || n_total < 0 || n_local < 0 || lo < 0 || lo + n_local > n_total)
        return fail(LSCQP_E_INVALID, "bad argument");
    if (K < 0 || K > h->cfg.max_obs) return fail(LSCQP_E_CAPACITY, "K above max_obs");
    if (n_local == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = knn_smem_bytes(n_total);
    if (smem > 200 * 1024) return fail(LSCQP_E_CAPACITY, "n_total above the shared-memory capacity of the selection kernel");
    if (smem > 48 * 1024 && smem > h->knn_smem) {
        CK(cudaFuncSetAttribute(knn_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        h->knn_smem = smem;
    }
    // fixed-stride rows + counts (scratch owned by the handle; allocate before any stream capture: the first call does)
    if (h->d_knn.reserve(((size_t) n_local * (K > 0 ? K : 1) + n_local) * sizeof(int))) return fail(LSCQP_E_CUDA, "cudaMalloc failed");
    KnnParams p;
    p.n_total = n_total; p.lo = lo; p.n_local = n_local; p.K = K; p.comm_range = comm_range;
    p.state = state; p.obs_index = h->d_knn.as<int>(); p.count = p.obs_index + (size_t) n_local * (K > 0 ? K : 1);
    p.overflow = overflow_out;
    knn_select_kernel<<<n_local, KNN_THREADS, smem, st>>>(p);
    KnnCsrParams c;
    c.n_local = n_local; c.K = K; c.rows = p.obs_index; c.count = p.count; c.obs_offsets = obs_offsets_out; c.obs_index = obs_index_out;
    knn_csr_kernel<<<1, KNN_CSR_THREADS, 0, st>>>(c);
    h->launches += 2;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int lscqp_validate_batch(lscqp_handle* h, int n_agents, const float* traj, const float* state_at_step,
                                    const double* limits, const float* sfc, int* valid_out, void* stream) {
    if (!h || n_agents < 0 || !traj || !state_at_step || !limits || !valid_out) return fail(LSCQP_E_INVALID, "null argument");
    if (h->cfg.use_sfc && !sfc) return fail(LSCQP_E_INVALID, "use_sfc set but sfc is null");
    if (n_agents == 0) return 0;
    ValidateParams p;
    p.n_agents = n_agents; p.M = h->cfg.M; p.dim = h->cfg.dim; p.use_sfc = h->cfg.use_sfc;
    p.traj = traj; p.state = state_at_step; p.limits = limits; p.sfc = sfc; p.valid_out = valid_out;
    validate_kernel<<<(n_agents + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int lscqp_goal_batch(lscqp_handle* h, int n_agents, const float* goal, const float* next_waypoint,
                                const float* sfc, const int* obs_offsets, const double* normals, const double* rhs,
                                float* goal_out, double* t_out, int* status_out, void* stream) {
    if (!h || n_agents < 0 || !goal || !next_waypoint || !obs_offsets || !goal_out || !status_out)
        return fail(LSCQP_E_INVALID, "null argument");
    if (h->cfg.use_sfc && !sfc) return fail(LSCQP_E_INVALID, "use_sfc set but sfc is null");
    if (n_agents == 0) return 0;
    GoalParams p;
    p.n_agents = n_agents; p.M = h->cfg.M; p.dim = h->cfg.dim; p.use_sfc = h->cfg.use_sfc; p.feas_tol = 1e-6;
    p.goal = goal; p.waypoint = next_waypoint; p.sfc = sfc; p.obs_offsets = obs_offsets; p.normals = normals; p.rhs = rhs;
    p.goal_out = goal_out; p.t_out = t_out; p.status_out = status_out;
    goal_lp_kernel<<<(n_agents + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Static map + Safe Flight Corridors (sfc_kernel.cuh)
extern "C" int lscqp_map_set(lscqp_handle* h, const double* boxes, int n_boxes, double resolution, double max_dist) {
    if (!h || n_boxes < 0 || (n_boxes > 0 && !boxes) || !(resolution > 0) || !(max_dist > 0)) return fail(LSCQP_E_INVALID, "bad argument");
    CK(cudaSetDevice(h->device));
    MapView m{};
    m.res = resolution; m.inv_res = 1.0 / resolution;
    size_t cells = 1;
    for (int k = 0; k < 3; k++) {
        m.world_min[k] = (float) h->cfg.world_min[k]; m.world_max[k] = (float) h->cfg.world_max[k];     // Mission::world_min/max are point3d
        m.key0[k] = (int) std::floor(m.inv_res * (double) m.world_min[k]);
        m.n[k] = (int) std::floor(m.inv_res * (double) m.world_max[k]) - m.key0[k] + 1;
        if (m.n[k] < 1 || m.n[k] > 1023) return fail(LSCQP_E_CAPACITY, "map needs 1..1023 cells per axis");
        cells *= (size_t) m.n[k];
    }
    m.maxd2 = (int) std::pow(max_dist / resolution, 2);
    if (h->d_occ.reserve(cells) || h->d_closest.reserve(cells * sizeof(int)) || h->d_boxes.reserve((size_t) (n_boxes > 0 ? n_boxes : 1) * 6 * sizeof(double)))
        return fail(LSCQP_E_CUDA, "cudaMalloc failed");
    m.occ = h->d_occ.as<unsigned char>(); m.closest = h->d_closest.as<int>();
    cudaStream_t st = h->stream;
    CK(cudaMemsetAsync(h->d_occ.p, 0, cells, st));
    if (n_boxes > 0) {
        CK(cudaMemcpyAsync(h->d_boxes.p, boxes, (size_t) n_boxes * 6 * sizeof(double), cudaMemcpyHostToDevice, st));
        OccParams op{m, h->d_occ.as<unsigned char>(), h->d_boxes.as<double>(), n_boxes};
        occupancy_kernel<<<n_boxes, 256, 0, st>>>(op);
    }
    EdtParams ep{m, h->d_closest.as<int>()};
    edt_closest_kernel<<<(unsigned) ((cells + 127) / 128), 128, 0, st>>>(ep);
    h->launches += n_boxes > 0 ? 2 : 1;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    h->map = m; h->has_map = true;
    return 0;
}

extern "C" int lscqp_map_get(lscqp_handle* h, int* n3, unsigned char* occ_out_host, int* closest_out_host) {
    if (!h || !h->has_map || !n3) return fail(LSCQP_E_INVALID, "no map set / null argument");
    CK(cudaSetDevice(h->device));
    const size_t cells = (size_t) h->map.n[0] * h->map.n[1] * h->map.n[2];
    for (int k = 0; k < 3; k++) n3[k] = h->map.n[k];
    if (occ_out_host) CK(cudaMemcpy(occ_out_host, h->d_occ.p, cells, cudaMemcpyDeviceToHost));
    if (closest_out_host) CK(cudaMemcpy(closest_out_host, h->d_closest.p, cells * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int lscqp_sfc_batch(lscqp_handle* h, int mode, int n_agents, const float* point, const float* goal,
                               const float* next_waypoint, const double* limits, float* sfc, int* status_out, void* stream) {
    if (!h || n_agents < 0 || !point || !limits || !sfc || !status_out) return fail(LSCQP_E_INVALID, "null argument");
    if (!h->has_map) return fail(LSCQP_E_INVALID, "lscqp_map_set has not been called on this handle");
    if (mode < SFC_INIT || mode > SFC_FROM_HULL) return fail(LSCQP_E_INVALID, "unknown mode");
    if (mode != SFC_INIT && !goal) return fail(LSCQP_E_INVALID, "goal is null");
    if (mode == SFC_FROM_HULL && !next_waypoint) return fail(LSCQP_E_INVALID, "next_waypoint is null");
    if (n_agents == 0) return 0;
    SfcParams p;
    p.map = h->map; p.mode = mode; p.n_agents = n_agents; p.M = h->cfg.M;
    p.point = point; p.goal = goal; p.waypoint = next_waypoint; p.limits = limits; p.sfc = sfc; p.status = status_out;
    sfc_kernel<<<n_agents, SFC_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

// host-buffer variant (what the CollisionConstraints shim calls for one agent): stages, launches, copies back, synchronises
extern "C" int lscqp_sfc_host(lscqp_handle* h, int mode, int n_agents, const float* point, const float* goal,
                              const float* next_waypoint, const double* limits, float* sfc, int* status_out) {
    if (!h || n_agents < 0 || !point || !limits || !sfc || !status_out) return fail(LSCQP_E_INVALID, "null argument");
    if (mode != SFC_INIT && !goal) return fail(LSCQP_E_INVALID, "goal is null");
    if (mode == SFC_FROM_HULL && !next_waypoint) return fail(LSCQP_E_INVALID, "next_waypoint is null");
    if (n_agents == 0) return 0;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const size_t n = (size_t) n_agents, M = (size_t) h->cfg.M;
    if (h->d_state.reserve(n * 9 * sizeof(float)) || h->d_goal.reserve(n * 3 * sizeof(float)) || h->d_wp.reserve(n * 3 * sizeof(float)) ||
        h->d_limits.reserve(n * 8 * sizeof(double)) || h->d_sfc.reserve(n * M * 6 * sizeof(float)) || h->d_status.reserve(n * sizeof(int)))
        return fail(LSCQP_E_CUDA, "cudaMalloc failed");
    CK(cudaMemcpyAsync(h->d_state.p, point, n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    if (goal) CK(cudaMemcpyAsync(h->d_goal.p, goal, n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    if (next_waypoint) CK(cudaMemcpyAsync(h->d_wp.p, next_waypoint, n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_limits.p, limits, n * 8 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_sfc.p, sfc, n * M * 6 * sizeof(float), cudaMemcpyHostToDevice, st));
    int rc = lscqp_sfc_batch(h, mode, n_agents, h->d_state.as<float>(), goal ? h->d_goal.as<float>() : nullptr,
                             next_waypoint ? h->d_wp.as<float>() : nullptr, h->d_limits.as<double>(), h->d_sfc.as<float>(),
                             h->d_status.as<int>(), st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(sfc, h->d_sfc.p, n * M * 6 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(status_out, h->d_status.p, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// FP64 roofline denominator: SURVEY 8(d) asks for the achieved FP64 rate of the solve kernel against an on-box
// FMA microbenchmark (MEASURED_PEAKS.json holds only HBM and bf16).  8 independent DFMA chains per thread.
__global__ void __launch_bounds__(256) fp64_fma_peak_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0 - 1e-9, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678) out[0] = r;                      // never true: keeps the chains alive
}

extern "C" int lscqp_measure_fp64_peak(lscqp_handle* h, double* gflops_out) {
    if (!h || !gflops_out) return fail(LSCQP_E_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    if (h->d_cost.reserve(64)) return fail(LSCQP_E_CUDA, "cudaMalloc failed");
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    const int iters = 1 << 15, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(e0, h->stream));
        fp64_fma_peak_kernel<<<blocks, threads, 0, h->stream>>>(h->d_cost.as<double>(), iters, 1.0);
        CK(cudaEventRecord(e1, h->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double gf = 2.0 * 8.0 * iters * (double) blocks * threads / (ms * 1e-3) / 1e9;
        if (rep > 0 && gf > best) best = gf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    h->launches += 4;
    *gflops_out = best;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// HOST-buffer entry points: what a TrajOptimizer / simulator running on the CPU calls.
#define RESERVE(buf, n) do { if ((buf).reserve(n)) return fail(LSCQP_E_CUDA, "cudaMalloc failed"); } while (0)

// every obstacle list of a host CSR must be non-negative and within the handle's capacity: the reference's model takes
// every obstacle it is given (traj_optimizer.cpp:400-437), so a list this library cannot hold is an error, not a cut
static int check_host_lists(const lscqp_handle* h, int n_agents, const int* obs_offsets) {
    if (obs_offsets[0] != 0) return fail(LSCQP_E_INVALID, "obs_offsets[0] must be 0");
    for (int a = 0; a < n_agents; a++) {
        if (obs_offsets[a + 1] < obs_offsets[a]) return fail(LSCQP_E_INVALID, "obs_offsets must be non-decreasing");
        if (obs_offsets[a + 1] - obs_offsets[a] > h->cfg.max_obs)
            return fail(LSCQP_E_CAPACITY, "obstacle list of an agent is longer than max_obs");
    }
    return 0;
}

extern "C" int lscqp_solve_host(lscqp_handle* h, int n_agents, const float* state, const float* goal,
                                const double* limits, const float* sfc, const float* next_waypoint, const int* obs_offsets,
                                const double* normals, const double* rhs, const float* initial_traj, double* ctrl_out,
                                double* cost_out, int* status_out, int* iters_out, double* kkt_out, double* dual_out) {
    if (!h || n_agents < 0 || !state || !goal || !limits || !obs_offsets || !ctrl_out || !cost_out || !status_out)
        return fail(LSCQP_E_INVALID, "null argument");
    if (n_agents == 0) return 0;
    if (int rc = check_host_lists(h, n_agents, obs_offsets)) return rc;
    CK(cudaSetDevice(h->device));
    const int M = h->cfg.M;
    const size_t sumK = (size_t) obs_offsets[n_agents];
    if (sumK > 0 && (!normals || !rhs)) return fail(LSCQP_E_INVALID, "null planes");
    cudaStream_t st = h->stream;
    RESERVE(h->d_state, n_agents * 9 * sizeof(float)); RESERVE(h->d_goal, n_agents * 3 * sizeof(float));
    RESERVE(h->d_limits, n_agents * 8 * sizeof(double)); RESERVE(h->d_off, (n_agents + 1) * sizeof(int));
    RESERVE(h->d_normals, (sumK * M * 3 + 1) * sizeof(double)); RESERVE(h->d_rhs, (sumK * M * 6 + 1) * sizeof(double));
    RESERVE(h->d_ctrl, (size_t) n_agents * h->nv * sizeof(double)); RESERVE(h->d_cost, n_agents * sizeof(double));
    RESERVE(h->d_status, n_agents * sizeof(int)); RESERVE(h->d_iters, n_agents * sizeof(int));
    RESERVE(h->d_kkt, n_agents * 4 * sizeof(double));
    if (dual_out) RESERVE(h->d_dual, (size_t) n_agents * h->dual_stride * sizeof(double));
    if (h->cfg.use_sfc) {
        if (!sfc) return fail(LSCQP_E_INVALID, "use_sfc set but sfc is null");
        RESERVE(h->d_sfc, (size_t) n_agents * M * 6 * sizeof(float));
        CK(cudaMemcpyAsync(h->d_sfc.p, sfc, (size_t) n_agents * M * 6 * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    CK(cudaMemcpyAsync(h->d_state.p, state, n_agents * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_goal.p, goal, n_agents * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_limits.p, limits, n_agents * 8 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_off.p, obs_offsets, (n_agents + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    if (sumK) {
        CK(cudaMemcpyAsync(h->d_normals.p, normals, sumK * M * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->d_rhs.p, rhs, sumK * M * 6 * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    if (h->cfg.comm_range > 0) {
        if (!next_waypoint) return fail(LSCQP_E_INVALID, "comm_range set but next_waypoint is null");
        RESERVE(h->d_wp, n_agents * 3 * sizeof(float));
        CK(cudaMemcpyAsync(h->d_wp.p, next_waypoint, n_agents * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    if (initial_traj) {
        RESERVE(h->d_own, (size_t) n_agents * M * 18 * sizeof(float));
        CK(cudaMemcpyAsync(h->d_own.p, initial_traj, (size_t) n_agents * M * 18 * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    int rc = lscqp_solve_batch(h, n_agents, h->d_state.as<float>(), h->d_goal.as<float>(), h->d_limits.as<double>(),
                               h->cfg.use_sfc ? h->d_sfc.as<float>() : nullptr,
                               h->cfg.comm_range > 0 ? h->d_wp.as<float>() : nullptr, h->d_off.as<int>(),
                               h->d_normals.as<double>(), h->d_rhs.as<double>(), initial_traj ? h->d_own.as<float>() : nullptr,
                               h->d_ctrl.as<double>(),
                               h->d_cost.as<double>(), h->d_status.as<int>(), h->d_iters.as<int>(), h->d_kkt.as<double>(),
                               dual_out ? h->d_dual.as<double>() : nullptr, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctrl_out, h->d_ctrl.p, (size_t) n_agents * h->nv * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(cost_out, h->d_cost.p, n_agents * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(status_out, h->d_status.p, n_agents * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (iters_out) CK(cudaMemcpyAsync(iters_out, h->d_iters.p, n_agents * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (kkt_out) CK(cudaMemcpyAsync(kkt_out, h->d_kkt.p, n_agents * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (dual_out) CK(cudaMemcpyAsync(dual_out, h->d_dual.p, (size_t) n_agents * h->dual_stride * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int lscqp_goal_host(lscqp_handle* h, int n_agents, const float* goal, const float* next_waypoint,
                               const float* sfc, const int* obs_offsets, const double* normals, const double* rhs,
                               float* goal_out, double* t_out, int* status_out) {
    if (!h || n_agents < 0 || !goal || !next_waypoint || !obs_offsets || !goal_out || !status_out)
        return fail(LSCQP_E_INVALID, "null argument");
    if (n_agents == 0) return 0;
    if (int rc = check_host_lists(h, n_agents, obs_offsets)) return rc;
    CK(cudaSetDevice(h->device));
    const int M = h->cfg.M;
    const size_t sumK = (size_t) obs_offsets[n_agents];
    if (sumK > 0 && (!normals || !rhs)) return fail(LSCQP_E_INVALID, "null planes");
    cudaStream_t st = h->stream;
    RESERVE(h->d_goal, n_agents * 3 * sizeof(float)); RESERVE(h->d_wp, n_agents * 3 * sizeof(float));
    RESERVE(h->d_off, (n_agents + 1) * sizeof(int));
    RESERVE(h->d_normals, (sumK * M * 3 + 1) * sizeof(double)); RESERVE(h->d_rhs, (sumK * M * 6 + 1) * sizeof(double));
    RESERVE(h->d_gout, n_agents * 3 * sizeof(float)); RESERVE(h->d_cost, n_agents * sizeof(double));
    RESERVE(h->d_status, n_agents * sizeof(int));
    if (h->cfg.use_sfc) {
        if (!sfc) return fail(LSCQP_E_INVALID, "use_sfc set but sfc is null");
        RESERVE(h->d_sfc, (size_t) n_agents * M * 6 * sizeof(float));
        CK(cudaMemcpyAsync(h->d_sfc.p, sfc, (size_t) n_agents * M * 6 * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    CK(cudaMemcpyAsync(h->d_goal.p, goal, n_agents * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_wp.p, next_waypoint, n_agents * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_off.p, obs_offsets, (n_agents + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    if (sumK) {
        CK(cudaMemcpyAsync(h->d_normals.p, normals, sumK * M * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->d_rhs.p, rhs, sumK * M * 6 * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    int rc = lscqp_goal_batch(h, n_agents, h->d_goal.as<float>(), h->d_wp.as<float>(), h->cfg.use_sfc ? h->d_sfc.as<float>() : nullptr,
                              h->d_off.as<int>(), h->d_normals.as<double>(), h->d_rhs.as<double>(), h->d_gout.as<float>(),
                              h->d_cost.as<double>(), h->d_status.as<int>(), st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(goal_out, h->d_gout.p, n_agents * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (t_out) CK(cudaMemcpyAsync(t_out, h->d_cost.p, n_agents * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(status_out, h->d_status.p, n_agents * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

// device-visible alias of a host pointer when it lies in pinned (page-locked, UVA-mapped) memory, else null
static void* mapped_alias(const void* host_ptr) {
    if (!host_ptr) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host_ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return at.devicePointer;
}

extern "C" int lscqp_replan_host(lscqp_handle* h, int generator, int n_agents, const float* state, const float* goal,
                                 const double* limits, const float* sfc, const float* next_waypoint, const float* own_traj,
                                 const double* agent_meta,
                                 const int* obs_offsets, const int* obs_index, double* ctrl_out, double* cost_out,
                                 int* status_out, int* iters_out) {
    if (!h || n_agents < 0 || !state || !goal || !limits || !own_traj || !agent_meta || !obs_offsets || !ctrl_out ||
        !cost_out || !status_out)
        return fail(LSCQP_E_INVALID, "null argument");
    if (n_agents == 0) return 0;
    if (h->cfg.use_sfc && !sfc) return fail(LSCQP_E_INVALID, "use_sfc set but sfc is null");
    if (h->cfg.comm_range > 0 && !next_waypoint) return fail(LSCQP_E_INVALID, "comm_range set but next_waypoint is null");
    if (int rc = check_host_lists(h, n_agents, obs_offsets)) return rc;
    CK(cudaSetDevice(h->device));
    const int M = h->cfg.M;
    const size_t sumK = (size_t) obs_offsets[n_agents];
    if (sumK > 0 && !obs_index) return fail(LSCQP_E_INVALID, "null obs_index");
    cudaStream_t st = h->stream;
    RESERVE(h->d_state, n_agents * 9 * sizeof(float)); RESERVE(h->d_goal, n_agents * 3 * sizeof(float));
    RESERVE(h->d_limits, n_agents * 8 * sizeof(double)); RESERVE(h->d_off, (n_agents + 1) * sizeof(int));
    RESERVE(h->d_own, (size_t) n_agents * M * 18 * sizeof(float)); RESERVE(h->d_ameta, n_agents * 2 * sizeof(double));
    RESERVE(h->d_index, (sumK + 1) * sizeof(int));
    // (no gathered obstacle copies: the assembly reads the neighbours in place through d_index)
    RESERVE(h->d_normals, (sumK * M * 3 + 1) * sizeof(double)); RESERVE(h->d_rhs, (sumK * M * 6 + 1) * sizeof(double));
    // Outputs in pinned host memory are written by the solve kernel itself through their device alias (every QP stores
    // its 720 bytes when it finishes, so the device->host transfer overlaps the rest of the batch); pageable outputs
    // are staged in device buffers and copied back after the kernel.
    double* k_ctrl = static_cast<double*>(mapped_alias(ctrl_out));
    double* k_cost = static_cast<double*>(mapped_alias(cost_out));
    int* k_status = static_cast<int*>(mapped_alias(status_out));
    int* k_iters = iters_out ? static_cast<int*>(mapped_alias(iters_out)) : nullptr;
    const bool direct = k_ctrl && k_cost && k_status && (!iters_out || k_iters);
    if (!direct) {
        RESERVE(h->d_ctrl, (size_t) n_agents * h->nv * sizeof(double)); RESERVE(h->d_cost, n_agents * sizeof(double));
        RESERVE(h->d_status, n_agents * sizeof(int)); RESERVE(h->d_iters, n_agents * sizeof(int));
        k_ctrl = h->d_ctrl.as<double>(); k_cost = h->d_cost.as<double>(); k_status = h->d_status.as<int>(); k_iters = h->d_iters.as<int>();
    }
    // host -> device: queue every copy first (largest first), then validate the neighbour ids while they are in flight
    CK(cudaMemcpyAsync(h->d_own.p, own_traj, (size_t) n_agents * M * 18 * sizeof(float), cudaMemcpyHostToDevice, st));
    if (sumK) CK(cudaMemcpyAsync(h->d_index.p, obs_index, sumK * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_limits.p, limits, n_agents * 8 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_state.p, state, n_agents * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_ameta.p, agent_meta, n_agents * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_goal.p, goal, n_agents * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_off.p, obs_offsets, (n_agents + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    if (h->cfg.use_sfc) {
        RESERVE(h->d_sfc, (size_t) n_agents * M * 6 * sizeof(float));
        CK(cudaMemcpyAsync(h->d_sfc.p, sfc, (size_t) n_agents * M * 6 * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    if (h->cfg.comm_range > 0) {
        RESERVE(h->d_wp, n_agents * 3 * sizeof(float));
        CK(cudaMemcpyAsync(h->d_wp.p, next_waypoint, n_agents * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    {
        unsigned bad = 0;
        const unsigned lim = (unsigned) n_agents;
        for (size_t j = 0; j < sumK; j++) bad |= (unsigned) ((unsigned) obs_index[j] >= lim);   // (negative ids wrap above lim)
        if (bad) {
            cudaStreamSynchronize(st);                           // nothing is launched, no output is written
            return fail(LSCQP_E_INVALID, "obs_index outside [0, n_agents)");
        }
    }
    // obstacles are the batch's own agents: the assembly reads them in place through the index list (no gathered
    // copies) and drops the (obstacle, segment) pairs that provably cannot bind (exact; presolve bit 0)
    int rc = lscqp_assemble_lsc_fused(h, generator, h->cfg.presolve & 1, n_agents, h->d_own.as<float>(), h->d_ameta.as<double>(),
                                      h->d_goal.as<float>(), h->d_state.as<float>(), h->d_limits.as<double>(), h->d_off.as<int>(),
                                      h->d_index.as<int>(), h->d_own.as<float>(), h->d_ameta.as<double>(), h->d_goal.as<float>(),
                                      h->d_state.as<float>(), h->d_normals.as<double>(), h->d_rhs.as<double>(), st);
    if (rc) return rc;
    rc = lscqp_solve_batch(h, n_agents, h->d_state.as<float>(), h->d_goal.as<float>(), h->d_limits.as<double>(),
                           h->cfg.use_sfc ? h->d_sfc.as<float>() : nullptr, h->cfg.comm_range > 0 ? h->d_wp.as<float>() : nullptr,
                           h->d_off.as<int>(), h->d_normals.as<double>(),
                           h->d_rhs.as<double>(), h->d_own.as<float>(), k_ctrl, k_cost, k_status, k_iters, nullptr, nullptr, st);
    if (rc) return rc;
    if (!direct) {
        CK(cudaMemcpyAsync(ctrl_out, h->d_ctrl.p, (size_t) n_agents * h->nv * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(cost_out, h->d_cost.p, n_agents * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(status_out, h->d_status.p, n_agents * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (iters_out) CK(cudaMemcpyAsync(iters_out, h->d_iters.p, n_agents * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}
