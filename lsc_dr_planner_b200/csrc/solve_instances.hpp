// solve_instances.hpp -- the PDIP kernel instances, split over translation units so that nvcc compiles them in
// parallel.  Each inst_<n>.cu defines LSCQP_TU (0..4) and includes this file; lscqp.cu sees only the declarations.
#pragma once
#include <cuda_runtime.h>
#include "host_common.hpp"

namespace lscqp {

struct InstanceInfo {
    int dual_stride = 0, kmax = 0, nv = 0;
    bool has_light = false, has_das = false;
    int das_big_slots = 0;         // CTAs of the large active-set instance the device holds at once (0: no such instance)
    int das_ckpt_stride = 0;       // doubles per checkpoint slot of the hand-over between the two instances (0: none)
    std::vector<double> das_tab;   // host_common.hpp:build_das_table (empty: no dual active-set pass)
    ProjTable tab, tab_light;      // projection term streams of the full-capacity / light instance
};

// Looks the (M, dim, mode, comm) instance up in translation unit N: fills `info`, raises the dynamic shared-memory
// limit of its kernels.  Returns 1 when found, 0 when the instance lives elsewhere, -1 on a CUDA error.
int inst_query_0(const lscqp_config& cfg, InstanceInfo* info);
int inst_query_1(const lscqp_config& cfg, InstanceInfo* info);
int inst_query_2(const lscqp_config& cfg, InstanceInfo* info);
int inst_query_3(const lscqp_config& cfg, InstanceInfo* info);
int inst_query_4(const lscqp_config& cfg, InstanceInfo* info);
// Launches the instance; first_pass 0: the full-capacity interior-point instance alone, 1: the light interior-point
// instance first, 2: the dual active-set kernels first (throughput instance, then the large one over the agents it could
// not hold), 3: the large active-set instance alone first (batches that fit the device in one wave); the full-capacity
// instance then solves only the agents the first pass flagged.  Returns the number of kernels launched, 0 when the instance lives elsewhere.
int inst_launch_0(const lscqp_config& cfg, SolveParams& p, int n_agents, int first_pass, cudaStream_t st);
int inst_launch_1(const lscqp_config& cfg, SolveParams& p, int n_agents, int first_pass, cudaStream_t st);
int inst_launch_2(const lscqp_config& cfg, SolveParams& p, int n_agents, int first_pass, cudaStream_t st);
int inst_launch_3(const lscqp_config& cfg, SolveParams& p, int n_agents, int first_pass, cudaStream_t st);
int inst_launch_4(const lscqp_config& cfg, SolveParams& p, int n_agents, int first_pass, cudaStream_t st);

#ifdef LSCQP_TU
#if LSCQP_TU == 0
#define LSCQP_TU_INSTANCES(X) X(5, 3, true, false) X(5, 3, false, false) X(5, 2, true, false) X(5, 2, false, false)
#define LSCQP_TU_NAME(f) f##_0
#elif LSCQP_TU == 1
#define LSCQP_TU_INSTANCES(X) X(10, 2, true, false) X(10, 2, false, false)
#define LSCQP_TU_NAME(f) f##_1
#elif LSCQP_TU == 2
#define LSCQP_TU_INSTANCES(X) X(10, 3, true, false) X(10, 3, false, false)
#define LSCQP_TU_NAME(f) f##_2
#elif LSCQP_TU == 3
#define LSCQP_TU_INSTANCES(X) X(5, 3, true, true) X(5, 2, true, true) X(10, 3, true, true) X(10, 2, true, true)
#define LSCQP_TU_NAME(f) f##_3
#else   // communication-range rows without the terminal stop: DLSC / BVC modes (the reference's defaults, param.cpp:117,129)
#define LSCQP_TU_INSTANCES(X) X(5, 3, false, true) X(5, 2, false, true) X(10, 3, false, true) X(10, 2, false, true)
#define LSCQP_TU_NAME(f) f##_4
#endif

template <class C>
static int set_smem_attr() {
    return cudaFuncSetAttribute(pdip_solve_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) == cudaSuccess ? 0 : -1;
}
template <class C, int KPT_>
static int set_smem_attr_das() {
    return cudaFuncSetAttribute(das_solve_kernel<C, KPT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, Das<C, KPT_>::SMEM_BYTES) == cudaSuccess ? 0 : -1;
}

// dual active-set first pass of instance family I (das_kernel.cuh): table + shared-memory limits; launch of the throughput
// instance over every agent and, where it exists, of the large instance over the agents flagged "too many kept obstacles"
template <class I>
static int das_prepare(const lscqp_config& cfg, InstanceInfo* info) {
    if constexpr (I::HAS_DAS) {
        using C = typename I::Full;
        SolveParams sp;
        fill_solve_params(cfg, sp);
        info->das_tab = build_das_table<C>(sp.Q2, cfg.w_terminal);
        info->has_das = !info->das_tab.empty();
        if (info->has_das && set_smem_attr_das<C, LSCQP_DAS_KPT>()) return -1;
        if constexpr (I::HAS_DAS_BIG) {
            if (info->has_das) {
                if (set_smem_attr_das<C, I::DAS_BIG_KPT>()) return -1;
                int per_sm = 0, dev = 0, sms = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, das_solve_kernel<C, I::DAS_BIG_KPT>, 32,
                                                                  Das<C, I::DAS_BIG_KPT>::SMEM_BYTES) != cudaSuccess) return -1;
                if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
                info->das_big_slots = per_sm * sms;
                info->das_ckpt_stride = Das<C, LSCQP_DAS_KPT>::CK_STRIDE;
            }
        }
    }
    return 0;
}
// big_only: the batch fits the device in one wave of the large instance (latency regime: every pass costs one QP's
// latency, so a single pass over everyone beats throughput instance + large instance for the flagged agents)
template <class I>
static int das_launch(SolveParams& p, int n_agents, cudaStream_t st, bool big_only) {
    int launched = 0;
    if constexpr (I::HAS_DAS) {
        using C = typename I::Full;
        p.klass_mode = 1;
        if constexpr (I::HAS_DAS_BIG) {
            if (big_only) {
                das_solve_kernel<C, I::DAS_BIG_KPT><<<n_agents, 32, Das<C, I::DAS_BIG_KPT>::SMEM_BYTES, st>>>(p);
                p.klass_mode = 2;
                return 1;
            }
        }
        das_solve_kernel<C, LSCQP_DAS_KPT><<<n_agents, 32, Das<C, LSCQP_DAS_KPT>::SMEM_BYTES, st>>>(p);
        launched = 1;
        if constexpr (I::HAS_DAS_BIG) {
            p.klass_mode = 3;
            das_solve_kernel<C, I::DAS_BIG_KPT><<<n_agents, 32, Das<C, I::DAS_BIG_KPT>::SMEM_BYTES, st>>>(p);
            launched = 2;
        }
        p.klass_mode = 2;
    }
    return launched;
}

int LSCQP_TU_NAME(inst_query)(const lscqp_config& cfg, InstanceInfo* info) {
    const bool term = cfg.planner_mode == LSCQP_MODE_LSC, comm = cfg.comm_range > 0;
#define X(M_, D_, T_, C_)                                                           \
    if (cfg.M == M_ && cfg.dim == D_ && term == T_ && comm == C_) {                 \
        using I = Instance<M_, D_, T_, C_>;                                         \
        using C = typename I::Full;                                                 \
        if (set_smem_attr<C>()) return -1;                                          \
        if (I::HAS_LIGHT && set_smem_attr<typename I::Light>()) return -1;          \
        if (I::HAS_COMPACT && set_smem_attr<typename I::Compact>()) return -1;      \
        info->dual_stride = C::DUAL_STRIDE; info->kmax = C::KMAX; info->nv = C::NV; \
        info->has_light = I::HAS_LIGHT;                                             \
        info->tab = build_projection<C>();                                          \
        if (I::HAS_COMPACT && cfg.max_obs <= I::COMPACT_KMAX)                       \
            info->tab = build_projection<typename I::Compact>();                    \
        if (I::HAS_LIGHT) info->tab_light = build_projection<C>(I::Light::NT);      \
        if (das_prepare<I>(cfg, info)) return -1;                                   \
        return 1;                                                                   \
    }
    LSCQP_TU_INSTANCES(X)
#undef X
    return 0;
}

int LSCQP_TU_NAME(inst_launch)(const lscqp_config& cfg, SolveParams& p, int n_agents, int first_pass, cudaStream_t st) {
    const bool term = cfg.planner_mode == LSCQP_MODE_LSC, comm = cfg.comm_range > 0;
#define X(M_, D_, T_, C_)                                                           \
    if (cfg.M == M_ && cfg.dim == D_ && term == T_ && comm == C_) {                 \
        using I = Instance<M_, D_, T_, C_>;                                         \
        using C = typename I::Full;                                                 \
        int launched = 1;                                                           \
        p.klass_mode = 0;                                                           \
        if (I::HAS_COMPACT && cfg.max_obs <= I::COMPACT_KMAX) {                     \
            using K = typename I::Compact;                                          \
            pdip_solve_kernel<K><<<n_agents, K::NT, K::SMEM_BYTES, st>>>(p);        \
            return launched;                                                        \
        }                                                                           \
        if (first_pass >= 2) launched += das_launch<I>(p, n_agents, st, first_pass == 3); \
        if (I::HAS_LIGHT && first_pass == 1) {                                      \
            using L = typename I::Light;                                            \
            p.klass_mode = 1;                                                       \
            pdip_solve_kernel<L><<<n_agents, L::NT, L::SMEM_BYTES, st>>>(p);        \
            p.klass_mode = 2;                                                       \
            launched = 2;                                                           \
        }                                                                           \
        pdip_solve_kernel<C><<<n_agents, C::NT, C::SMEM_BYTES, st>>>(p);            \
        return launched;                                                            \
    }
    LSCQP_TU_INSTANCES(X)
#undef X
    return 0;
}
#endif  // LSCQP_TU

}  // namespace lscqp
