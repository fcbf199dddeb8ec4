// sfc_kernel.cuh -- Safe Flight Corridor construction on the device.
//
// Replaces, for the whole batch, what TrajPlanner::generateSFC (src/traj_planner.cpp:738-753) does per agent through
//   CollisionConstraints::initializeSFC              src/collision_constraints.cpp:366-383
//   CollisionConstraints::constructSFCFromPoint      :396-411  (+ expandSFCFromPoint :669-694)
//   CollisionConstraints::constructSFCFromConvexHull :413-436  (+ expandSFCFromConvexHull :696-733, :735-777)
//   CollisionConstraints::expandSFC                  :820-881 / :883-946, setAxisCand :1134-1170
//   CollisionConstraints::isObstacleInSFC            :779-809
// and, once per world, the static-map pipeline behind them: MapManager::updateOctreeFromCSV (src/map_manager.cpp:262-305,
// world CSV boxes -> occupied 0.1 m cells) and the nearest-obstacle field DynamicEDTOctomap(maxdist = 1.0) serves
// (src/map_manager.cpp:59-80).  octomap / dynamicEDT3D are not used: the occupancy grid is a byte array in HBM and the
// nearest occupied cell of every cell is found by a brute-force window scan (edt_closest_kernel; one launch per world).
//
// Arithmetic follows the reference's types exactly (box corners and points are float, the resolution, the margin and
// every product with them double, narrowed on assignment); the _rn intrinsics keep nvcc from contracting what the
// reference's host compiler evaluates as separate operations.  Ties between equally near occupied cells go to the lowest
// (x, y, z) index (the reference's brushfire order is not specified).
#pragma once
#include <math.h>

namespace lscqp {

struct MapView {
    int n[3];                     // cells per axis
    int key0[3];                  // octree key floor(coord * (1 / res)) of cell 0 (= key of world_min)
    double res, inv_res;
    int maxd2;                    // (maxdist / res)^2: nearest-obstacle data is valid below it
    const unsigned char* occ;     // [nx][ny][nz]
    const int* closest;           // [nx][ny][nz] packed nearest occupied cell x | y << 10 | z << 20, or -1
    float world_min[3], world_max[3];
};

// ---------------------------------------------------------------------------------------------
// occupancy from the world boxes: one thread per (box, cell of its index range)
struct OccParams {
    MapView map;
    unsigned char* occ;
    const double* boxes;          // [n_boxes][6] centre xyz, size xyz
    int n_boxes;
};

__device__ __forceinline__ int map_key(const MapView& m, float coord) { return (int) floor(__dmul_rn(m.inv_res, (double) coord)); }

__global__ void __launch_bounds__(256) occupancy_kernel(const OccParams p) {
    const int b = blockIdx.x;
    if (b >= p.n_boxes) return;
    int lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
        const float com = (float) p.boxes[b * 6 + k], size = (float) p.boxes[b * 6 + 3 + k];
        lo[k] = (int) round(__ddiv_rn(__dsub_rn((double) com, __dmul_rn(0.5, (double) size)), p.map.res));
        hi[k] = (int) round(__ddiv_rn(__dadd_rn((double) com, __dmul_rn(0.5, (double) size)), p.map.res));
    }
    const int ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    if (ex <= 0 || ey <= 0 || ez <= 0) return;
    const long total = (long) ex * ey * ez;
    for (long e = threadIdx.x; e < total; e += blockDim.x) {
        const int idx[3] = {lo[0] + (int) (e / ((long) ey * ez)), lo[1] + (int) ((e / ez) % ey), lo[2] + (int) (e % ez)};
        int c[3];
        bool in = true;
        for (int a = 0; a < 3; a++) {
            const float pt = (float) __dmul_rn(__dadd_rn((double) idx[a], 0.5), p.map.res);   // the inserted cell-centre point
            c[a] = map_key(p.map, pt) - p.map.key0[a];
            in = in && c[a] >= 0 && c[a] < p.map.n[a];
        }
        if (in) p.occ[((size_t) c[0] * p.map.n[1] + c[1]) * p.map.n[2] + c[2]] = 1;
    }
}

// nearest occupied cell of every cell (Euclidean, cell units, valid below maxd2; ties: lowest x, then y, then z)
struct EdtParams {
    MapView map;
    int* closest;
};

__global__ void __launch_bounds__(128) edt_closest_kernel(const EdtParams p) {
    const MapView& m = p.map;
    const long cells = (long) m.n[0] * m.n[1] * m.n[2];
    const long cell = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= cells) return;
    const int z = (int) (cell % m.n[2]), y = (int) ((cell / m.n[2]) % m.n[1]), x = (int) (cell / ((long) m.n[1] * m.n[2]));
    const int R = (int) ceil(sqrt((double) m.maxd2));
    int best = m.maxd2, packed = -1;
    for (int i = (x - R < 0 ? 0 : x - R); i <= x + R && i < m.n[0]; i++) {
        const int dx2 = (i - x) * (i - x);
        if (dx2 >= best) continue;
        for (int j = (y - R < 0 ? 0 : y - R); j <= y + R && j < m.n[1]; j++) {
            const int dxy2 = dx2 + (j - y) * (j - y);
            if (dxy2 >= best) continue;
            const unsigned char* col = m.occ + ((size_t) i * m.n[1] + j) * m.n[2];
            for (int k = (z - R < 0 ? 0 : z - R); k <= z + R && k < m.n[2]; k++) {
                if (!col[k]) continue;
                const int d2 = dxy2 + (k - z) * (k - z);
                if (d2 < best) { best = d2; packed = i | (j << 10) | (k << 20); }
            }
        }
    }
    p.closest[cell] = packed;
}

// ---------------------------------------------------------------------------------------------
enum { SFC_INIT = 0, SFC_FROM_POINT = 1, SFC_FROM_HULL = 2 };

struct SfcParams {
    MapView map;
    int mode, n_agents, M;
    const float*  point;          // [n][3] INIT: current position; otherwise initial_traj.lastPoint()
    const float*  goal;           // [n][3] agent.current_goal_point (FROM_POINT: growth order; FROM_HULL: second hull point)
    const float*  waypoint;       // [n][3] agent.next_waypoint (FROM_HULL)
    const double* limits;         // [n][8] radius at [6]
    float* sfc;                   // [n][M][6] box_min, box_max per segment (in / out)
    int*   status;                // [n] INIT: 1 ok, 0 invalid start box (the reference throws); FROM_POINT: 1 / 0 (previous
                                  //     corridor reused); FROM_HULL: 2 hull + waypoint, 1 hull clipped to the previous, 0 reused
};

constexpr int SFC_THREADS = 128;
#define SFC_EPS 1e-5               // SP_EPSILON_FLOAT

// CollisionConstraints::isObstacleInSFC (:779-809), the CTA's threads striding over the grid points of the box
__device__ __forceinline__ bool sfc_obstacle_in_box(const MapView& m, const float* box, double margin) {
    const float delta = (float) __dmul_rn(0.5, m.res);
    int size[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float ext = __fsub_rn(box[3 + i], box[i]);
        size[i] = (int) floor(__ddiv_rn(__dadd_rn((double) ext, SFC_EPS), m.res)) + 1;
    }
    int found = 0;
    if (size[0] > 0 && size[1] > 0 && size[2] > 0) {
        const int total = size[0] * size[1] * size[2];
        for (int e0 = 0; e0 < total; e0 += SFC_THREADS) {
            const int e = e0 + (int) threadIdx.x;
            if (e < total) {
                const int it[3] = {e / (size[1] * size[2]), (e / size[2]) % size[1], e % size[2]};
                float sp[3], cl[3] = {0.f, 0.f, 0.f};                               // default-constructed point3d
                int c[3];
                bool in = true;
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    sp[i] = (float) __dadd_rn((double) box[i], __dmul_rn((double) it[i], m.res));
                    c[i] = map_key(m, sp[i]) - m.key0[i];
                    in = in && c[i] >= 0 && c[i] < m.n[i];
                }
                if (in) {
                    const int o = m.closest[((size_t) c[0] * m.n[1] + c[1]) * m.n[2] + c[2]];
                    if (o >= 0) {
                        const int oc[3] = {o & 1023, (o >> 10) & 1023, (o >> 20) & 1023};
#pragma unroll
                        for (int i = 0; i < 3; i++) cl[i] = (float) __dmul_rn(__dadd_rn((double) (oc[i] + m.key0[i]), 0.5), m.res);
                    }
                }
                double dist = 0.0;
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    const float lo = __fsub_rn(cl[i], delta), hi = __fadd_rn(cl[i], delta);
                    float cp = sp[i];
                    if (sp[i] < lo) cp = lo; else if (sp[i] > hi) cp = hi;
                    const double ad = fabs((double) __fsub_rn(cp, sp[i]));
                    if (dist < ad) dist = ad;
                }
                if (dist < __dadd_rn(margin, SFC_EPS)) found = 1;
            }
            if (__syncthreads_or(found)) return true;
        }
    }
    return false;
}

__device__ __forceinline__ bool sfc_in_boundary(const MapView& m, const float* box) {   // isSFCInBoundary(box, 0), :811-818
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (!((double) box[k] > (double) m.world_min[k] - SFC_EPS)) return false;
        if (!((double) box[3 + k] < (double) m.world_max[k] + SFC_EPS)) return false;
    }
    return true;
}

__device__ __forceinline__ void sfc_axis_order(const float* box, const float* goal, int* axis_cand) {   // setAxisCand :1134-1170
    int offsets[3], order[3], n = 0;
    double values[3];
    for (int k = 0; k < 3; k++) {
        const float mid = __fmul_rn(__fadd_rn(box[k], box[3 + k]), 0.5f);
        const float d = __fsub_rn(goal[k], mid);
        offsets[k] = d > 0 ? 3 : 0;
        values[k] = fabs((double) d);
    }
    double max_value = -1, min_value = 1e9;
    for (int i = 0; i < 3; i++) {
        int pos;
        if (values[i] > max_value) { pos = 0; max_value = values[i]; }
        else if (values[i] < min_value) { pos = n; min_value = values[i]; }
        else pos = 1;
        for (int j = n; j > pos; j--) order[j] = order[j - 1];
        order[pos] = i; n++;
    }
    for (int i = 0; i < 3; i++) {
        axis_cand[i] = order[i] + offsets[order[i]];
        axis_cand[5 - i] = order[i] + (3 - offsets[order[i]]);
    }
}

// CollisionConstraints::expandSFC (:820-881; goal-directed axis order :883-946 when goal != nullptr); CTA-uniform control flow
__device__ __forceinline__ bool sfc_expand(const MapView& m, const float* initial, const float* goal, double margin, float* out) {
    if (sfc_obstacle_in_box(m, initial, margin)) return false;
    int axis_cand[6] = {0, 1, 2, 3, 4, 5}, n_cand = 6;
    if (goal) sfc_axis_order(initial, goal, axis_cand);
    float sfc[6], cand[6], upd[6];
#pragma unroll
    for (int e = 0; e < 6; e++) sfc[e] = initial[e];
    int i = -1;
    while (n_cand > 0) {
#pragma unroll
        for (int e = 0; e < 6; e++) { cand[e] = sfc[e]; upd[e] = sfc[e]; }
        while (sfc_in_boundary(m, upd) && !sfc_obstacle_in_box(m, upd, margin)) {
            i++;
            if (i >= n_cand) i = 0;
            const int axis = axis_cand[i];
#pragma unroll
            for (int e = 0; e < 6; e++) { sfc[e] = cand[e]; upd[e] = cand[e]; }
            if (axis < 3) {
                upd[3 + axis] = cand[axis];
                cand[axis] = (float) __dsub_rn((double) cand[axis], m.res);
                upd[axis] = cand[axis];
            } else {
                upd[axis - 3] = cand[axis];
                cand[axis] = (float) __dadd_rn((double) cand[axis], m.res);
                upd[axis] = cand[axis];
            }
        }
        if (i < 0) i = 0;
        for (int j = i; j + 1 < n_cand; j++) axis_cand[j] = axis_cand[j + 1];
        n_cand--;
        if (i > 0) i--; else i = n_cand - 1;
    }
    const double delta = __dsub_rn(margin, __dmul_rn((double) (int) __ddiv_rn(margin, m.res), m.res));   // margin compensation :868-877
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if ((double) sfc[k] > (double) m.world_min[k] + SFC_EPS) sfc[k] = (float) __dsub_rn((double) sfc[k], delta);
        if ((double) sfc[3 + k] < (double) m.world_max[k] - SFC_EPS) sfc[3 + k] = (float) __dadd_rn((double) sfc[3 + k], delta);
    }
#pragma unroll
    for (int e = 0; e < 6; e++) out[e] = sfc[e];
    return true;
}

__device__ __forceinline__ bool sfc_point_in_box(const float* box, const float* p) {   // Box::isPointInBox :81-88
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (!((double) p[k] > (double) box[k] - SFC_EPS && (double) p[k] < (double) box[3 + k] + SFC_EPS)) return false;
    return true;
}

// a start box that leaves the previous corridor is intersected with it and re-aligned inwards (:680-688, :764-771)
__device__ __forceinline__ void sfc_clip_to_prev(const MapView& m, const float* prev, float* init) {
    if (sfc_point_in_box(prev, init) && sfc_point_in_box(prev, init + 3)) return;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        init[k] = init[k] > prev[k] ? init[k] : prev[k];
        init[3 + k] = init[3 + k] < prev[3 + k] ? init[3 + k] : prev[3 + k];
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        init[k] = (float) __dmul_rn(ceil(__ddiv_rn(__dsub_rn((double) init[k], SFC_EPS), m.res)), m.res);
        init[3 + k] = (float) __dmul_rn(floor(__ddiv_rn(__dadd_rn((double) init[3 + k], SFC_EPS), m.res)), m.res);
    }
}

__device__ __forceinline__ bool sfc_superset(const float* box, const float (*pts)[3], int n) {   // isSuperSetOfConvexHull :135-150
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float lo = pts[0][i], hi = pts[0][i];
        for (int q = 1; q < n; q++) { if (pts[q][i] < lo) lo = pts[q][i]; if (pts[q][i] > hi) hi = pts[q][i]; }
        if ((double) lo < (double) box[i] - SFC_EPS || (double) hi > (double) box[3 + i] + SFC_EPS) return false;
    }
    return true;
}

// one CTA per agent
__global__ void __launch_bounds__(SFC_THREADS) sfc_kernel(const SfcParams p) {
    const int agent = blockIdx.x;
    if (agent >= p.n_agents) return;
    const MapView& m = p.map;
    const double radius = p.limits[(size_t) agent * 8 + 6];
    float* boxes = p.sfc + (size_t) agent * p.M * 6;
    float pt[3], init[6], out[6];
#pragma unroll
    for (int k = 0; k < 3; k++) pt[k] = p.point[(size_t) agent * 3 + k];
    int status = 0;
    if (p.mode == SFC_INIT) {                                                     // initializeSFC :366-383
#pragma unroll
        for (int k = 0; k < 3; k++) {
            init[k] = (float) __dmul_rn(floor(__ddiv_rn((double) pt[k], m.res)), m.res);
            init[3 + k] = (float) __dmul_rn(ceil(__ddiv_rn((double) pt[k], m.res)), m.res);
        }
        status = sfc_expand(m, init, nullptr, radius, out) ? 1 : 0;
        if (status)
            for (int e = threadIdx.x; e < p.M * 6; e += blockDim.x) boxes[e] = out[e % 6];
        if (threadIdx.x == 0) p.status[agent] = status;
        return;
    }
    float prev[6], goal[3];
#pragma unroll
    for (int e = 0; e < 6; e++) prev[e] = boxes[(p.M - 1) * 6 + e];
#pragma unroll
    for (int k = 0; k < 3; k++) goal[k] = p.goal[(size_t) agent * 3 + k];
    bool ok = false;
    if (p.mode == SFC_FROM_POINT) {                                               // constructSFCFromPoint :396-411, :669-694
#pragma unroll
        for (int k = 0; k < 3; k++) {
            init[k] = (float) __dmul_rn(floor(__ddiv_rn((double) pt[k], m.res)), m.res);
            init[3 + k] = (float) __dmul_rn(ceil(__ddiv_rn((double) pt[k], m.res)), m.res);
        }
        sfc_clip_to_prev(m, prev, init);
        ok = sfc_expand(m, init, goal, radius, out);
        status = ok ? 1 : 0;
    } else {                                                                      // constructSFCFromConvexHull :413-436
        float pts[3][3];
#pragma unroll
        for (int k = 0; k < 3; k++) { pts[0][k] = pt[k]; pts[1][k] = goal[k]; pts[2][k] = p.waypoint[(size_t) agent * 3 + k]; }
        // (i) hull + next waypoint, corners rounded to the grid (:696-733)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float lo = fminf(fminf(pts[0][k], pts[1][k]), pts[2][k]), hi = fmaxf(fmaxf(pts[0][k], pts[1][k]), pts[2][k]);
            init[k] = (float) __dmul_rn(round(__ddiv_rn((double) lo, m.res)), m.res);
            init[3 + k] = (float) __dmul_rn(round(__ddiv_rn((double) hi, m.res)), m.res);
        }
        ok = sfc_expand(m, init, nullptr, radius, out) && sfc_superset(out, pts, 3);
        status = ok ? 2 : 0;
        if (!ok) {                                                                // (ii) hull alone, inside the previous corridor (:735-777)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float lo = fminf(pts[0][k], pts[1][k]), hi = fmaxf(pts[0][k], pts[1][k]);
                init[k] = (float) __dmul_rn(floor(__ddiv_rn((double) lo, m.res)), m.res);
                init[3 + k] = (float) __dmul_rn(ceil(__ddiv_rn((double) hi, m.res)), m.res);
            }
            sfc_clip_to_prev(m, prev, init);
            ok = sfc_expand(m, init, nullptr, radius, out);
            status = ok ? 1 : 0;
        }
    }
    // shift the corridors by one segment; the last one is the new box, or the previous one again (:398-410, :415-435)
    const int e = threadIdx.x;                                                    // (M * 6 <= SFC_THREADS)
    float v = 0.f;
    if (e < p.M * 6) v = (e / 6 < p.M - 1) ? boxes[e + 6] : (ok ? out[e % 6] : prev[e % 6]);
    __syncthreads();
    if (e < p.M * 6) boxes[e] = v;
    if (threadIdx.x == 0) p.status[agent] = status;
}

}  // namespace lscqp
