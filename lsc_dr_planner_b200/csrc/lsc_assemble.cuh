// lsc_assemble.cuh -- device assembly of the Linear Safe Corridor half-spaces.
//
// Replaces TrajPlanner::generateLSC / generateCLSC / generateBVC
// (src/traj_planner.cpp:611-657, :659-706, :708-736) and CollisionConstraints::setLSC
// (src/collision_constraints.cpp:514-539) for agent-type obstacles, for every (agent, obstacle,
// segment) of the batch in one launch.  The output is the packed form the solve kernel reads:
//   normal[oi][m] (3 doubles, the float normal widened) and rhs[oi][m][i] = n . p_i + d_i,
// so that row (oi, m, i) of the QP is  normal . c[m][i] >= rhs  (src/traj_optimizer.cpp:413-429).
//
// Arithmetic follows the reference's types: control points and normals are float
// (octomap::point3d); the closest point of the relative hull is found in double
// (normalVectorBetweenPolys :1179-1205 -> openGJK) and narrowed to float before normalisation
// (include/geometry.hpp:292).  Float expressions use the _rn intrinsics so nvcc cannot contract
// them into FMAs the host compiler of the reference would not emit.
#pragma once
#include <math.h>

namespace lscqp {

struct AssembleParams {
    int n_agents, generator, dim;
    const float*  own_traj;      // [n][M][6][3]
    const double* agent_meta;    // [n][2] radius, downwash (doubles in the reference's Agent)
    const float*  agent_goal;    // [n][3]
    const int*    obs_offsets;   // [n+1]
    const float*  obs_traj;      // [sumK][M][6][3]
    const float*  obs_meta;      // [sumK][4] radius, downwash (floats in the reference's Obstacle)
    const float*  obs_goal;      // [sumK][3]
    const float*  obs_position;  // [sumK][3]
    double* normals;             // [sumK][M][3]
    double* rhs;                 // [sumK][M][6]
    // fused replan path (lscqp_assemble_lsc_fused): obstacles are read in place through obs_index from the
    // population's own arrays (no gathered copies), and with `prune` the (obstacle, segment) pairs whose rows provably
    // cannot bind at any point the velocity rows allow are written as zero normals (= rows the QP drops,
    // traj_optimizer.cpp:409-411) without running the hull enumeration
    const int*    obs_index;     // [sumK] ids into the all_* arrays, or null (use obs_*)
    const float*  all_traj;      // [n_total][M][6][3]
    const double* all_meta;      // [n_total][2]
    const float*  all_goal;      // [n_total][3]
    const float*  all_state;     // [n_total][9]
    // split dispatch of the pruned path (lscqp_assemble_lsc_fused at throughput batch sizes): lsc_assemble_kernel only
    // prunes and appends the surviving pairs to a global work list, lsc_pairs_kernel then computes their planes with
    // every thread of the grid busy (in one kernel the ~15 % surviving pairs leave most threads of a CTA idle while a few
    // run the long fp64 hull enumeration)
    int2* work_list;             // [sum K * M] (agent, pair row j * M + m); null: planes computed in the same kernel
    int*  work_count;            // [1] zeroed before the launch
    const double* obs_size;      // [sumK][M][6] predicted obstacle sizes (generator 3; null: the obstacle radius)
    int prune;
    const float*  state;         // [n][9]   (prune)
    const double* limits;        // [n][8]   (prune)
    double dt;
};

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 f3_make(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 f3_sub(f3 a, f3 b) { return f3_make(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ f3 f3_add(f3 a, f3 b) { return f3_make(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
__device__ __forceinline__ f3 f3_neg(f3 a) { return f3_make(-a.x, -a.y, -a.z); }
// Vector3::operator*(float): a double factor is narrowed first
__device__ __forceinline__ f3 f3_scale(f3 a, double s) {
    const float f = (float) s;
    return f3_make(__fmul_rn(a.x, f), __fmul_rn(a.y, f), __fmul_rn(a.z, f));
}
// Vector3::dot / norm: float arithmetic, widened on return
__device__ __forceinline__ double f3_dot(f3 a, f3 b) {
    return (double) __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ double f3_norm(f3 a) { return sqrt(f3_dot(a, a)); }
__device__ __forceinline__ double f3_distance(f3 a, f3 b) {
    const double dx = (double) a.x - (double) b.x, dy = (double) a.y - (double) b.y, dz = (double) a.z - (double) b.z;
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}
__device__ __forceinline__ f3 f3_normalized(f3 a) {
    const double len = f3_norm(a);
    if (len > 0) { const float f = (float) len; return f3_make(__fdiv_rn(a.x, f), __fdiv_rn(a.y, f), __fdiv_rn(a.z, f)); }
    return a;
}
__device__ __forceinline__ f3 f3_cross(f3 a, f3 b) {
    return f3_make(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)),
                   __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
                   __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ bool f3_eq(f3 a, f3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

// ---------------------------------------------------------------------------------------------
// Closest point of conv{p_0..p_5} to the origin (the witness vector gjk() returns,
// src/openGJK/openGJK.cpp:674-780).  Every vertex, edge and triangle whose affine projection of
// the origin has positive barycentrics is a candidate; the optimality condition
// p_i . v >= v . v then tells whether the origin is inside the hull (v = 0).
__device__ __forceinline__ void min_norm_hull6(const double (*p)[3], double* v) {
    double best = INFINITY, bx = 0, by = 0, bz = 0;
#pragma unroll
    for (int a = 0; a < 6; a++) {
        const double nn = p[a][0] * p[a][0] + p[a][1] * p[a][1] + p[a][2] * p[a][2];
        if (nn < best) { best = nn; bx = p[a][0]; by = p[a][1]; bz = p[a][2]; }
    }
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = a + 1; b < 6; b++) {
            const double dx = p[b][0] - p[a][0], dy = p[b][1] - p[a][1], dz = p[b][2] - p[a][2];
            const double dd = dx * dx + dy * dy + dz * dz;
            const double pd = p[a][0] * dx + p[a][1] * dy + p[a][2] * dz;
            // t = -pd/dd in (0,1)
            if (dd > 0.0 && pd < 0.0 && -pd < dd) {
                const double t = -pd / dd;
                const double cx = p[a][0] + t * dx, cy = p[a][1] + t * dy, cz = p[a][2] + t * dz;
                const double nn = cx * cx + cy * cy + cz * cz;
                if (nn < best) { best = nn; bx = cx; by = cy; bz = cz; }
            }
        }
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = a + 1; b < 6; b++)
#pragma unroll
            for (int c = b + 1; c < 6; c++) {
                const double d1x = p[b][0] - p[a][0], d1y = p[b][1] - p[a][1], d1z = p[b][2] - p[a][2];
                const double d2x = p[c][0] - p[a][0], d2y = p[c][1] - p[a][1], d2z = p[c][2] - p[a][2];
                const double g11 = d1x * d1x + d1y * d1y + d1z * d1z, g12 = d1x * d2x + d1y * d2y + d1z * d2z,
                             g22 = d2x * d2x + d2y * d2y + d2z * d2z;
                const double det = g11 * g22 - g12 * g12;
                if (!(det > 1e-14 * g11 * g22)) continue;          // degenerate: the edges cover it
                const double r1 = -(p[a][0] * d1x + p[a][1] * d1y + p[a][2] * d1z),
                             r2 = -(p[a][0] * d2x + p[a][1] * d2y + p[a][2] * d2z);
                const double n1 = r1 * g22 - r2 * g12, n2 = g11 * r2 - g12 * r1;   // m = n / det
                if (!(n1 > 0.0 && n2 > 0.0 && n1 + n2 < det)) continue;
                const double m1 = n1 / det, m2 = n2 / det;
                const double cx = p[a][0] + m1 * d1x + m2 * d2x, cy = p[a][1] + m1 * d1y + m2 * d2y,
                             cz = p[a][2] + m1 * d1z + m2 * d2z;
                const double nn = cx * cx + cy * cy + cz * cz;
                if (nn < best) { best = nn; bx = cx; by = cy; bz = cz; }
            }
    // optimality: v is the min-norm point iff p_i . v >= v . v for every vertex
    double worst = INFINITY;
#pragma unroll
    for (int a = 0; a < 6; a++) worst = fmin(worst, p[a][0] * bx + p[a][1] * by + p[a][2] * bz);
    if (worst < best * (1.0 - 1e-9)) { bx = 0; by = 0; bz = 0; }
    v[0] = bx; v[1] = by; v[2] = bz;
}

// ---------------------------------------------------------------------------------------------
// include/geometry.hpp:67-102 (point vs segment), :129-172 (lines), :174-264 (segments), in float.
struct ClosestPts { double dist; f3 cp1, cp2; };

__device__ __forceinline__ ClosestPts closest_point_segment(f3 point, f3 ls, f3 le) {
    const f3 a = f3_sub(ls, point), b = f3_sub(le, point);
    f3 rel = a;
    double dist_min = f3_norm(a);
    if (!f3_eq(a, b)) {
        double dist = f3_norm(b);
        if (dist_min > dist) { dist_min = dist; rel = b; }
        const f3 n_line = f3_normalized(f3_sub(b, a));
        const f3 c = f3_sub(a, f3_scale(n_line, f3_dot(a, n_line)));
        dist = f3_norm(c);
        if (f3_dot(f3_sub(c, a), f3_sub(c, b)) < 0 && dist_min > dist) { dist_min = dist; rel = c; }
    }
    ClosestPts r; r.dist = dist_min; r.cp1 = point; r.cp2 = f3_add(rel, point);
    return r;
}

__device__ __forceinline__ ClosestPts closest_points_lines(f3 l1s, f3 l1e, f3 l2s, f3 l2e) {
    ClosestPts r;
    const f3 n1 = f3_normalized(f3_sub(l1e, l1s)), n2 = f3_normalized(f3_sub(l2e, l2s));
    if (f3_distance(n1, n2) < 1e-5 || f3_distance(n1, f3_neg(n2)) < 1e-5) {
        f3 delta = f3_sub(l2s, l1s);
        delta = f3_sub(delta, f3_scale(n1, f3_dot(delta, n1)));
        r.dist = f3_norm(delta); r.cp1 = l1s; r.cp2 = f3_add(l1s, delta);
    } else {
        const f3 delta = f3_sub(l2s, l1s);
        const f3 n3 = f3_normalized(f3_cross(n2, n1));
        // alphas = A^-1 delta, A = [n1 | -n2 | n3]; 3x3 inverse by cofactors of column 0 in float (Eigen::Matrix3f)
        const float A00 = n1.x, A01 = -n2.x, A02 = n3.x, A10 = n1.y, A11 = -n2.y, A12 = n3.y, A20 = n1.z, A21 = -n2.z, A22 = n3.z;
#define FSUBMUL(a, b, c, d) __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d))
        const float c00 = FSUBMUL(A11, A22, A12, A21), c10 = FSUBMUL(A21, A02, A22, A01), c20 = FSUBMUL(A01, A12, A02, A11);
        const float det = __fadd_rn(__fadd_rn(__fmul_rn(c00, A00), __fmul_rn(c10, A10)), __fmul_rn(c20, A20));
        const float id = __fdiv_rn(1.0f, det);
        const float i00 = __fmul_rn(c00, id), i01 = __fmul_rn(c10, id), i02 = __fmul_rn(c20, id);
        const float i10 = __fmul_rn(FSUBMUL(A12, A20, A10, A22), id), i11 = __fmul_rn(FSUBMUL(A22, A00, A20, A02), id),
                    i12 = __fmul_rn(FSUBMUL(A02, A10, A00, A12), id);
        const float i20 = __fmul_rn(FSUBMUL(A10, A21, A11, A20), id), i21 = __fmul_rn(FSUBMUL(A20, A01, A21, A00), id),
                    i22 = __fmul_rn(FSUBMUL(A00, A11, A01, A10), id);
#undef FSUBMUL
        const float al0 = __fadd_rn(__fadd_rn(__fmul_rn(i00, delta.x), __fmul_rn(i01, delta.y)), __fmul_rn(i02, delta.z));
        const float al1 = __fadd_rn(__fadd_rn(__fmul_rn(i10, delta.x), __fmul_rn(i11, delta.y)), __fmul_rn(i12, delta.z));
        const float al2 = __fadd_rn(__fadd_rn(__fmul_rn(i20, delta.x), __fmul_rn(i21, delta.y)), __fmul_rn(i22, delta.z));
        r.dist = fabs((double) al2);
        r.cp1 = f3_add(l1s, f3_scale(n1, al0));
        r.cp2 = f3_add(l2s, f3_scale(n2, al1));
    }
    return r;
}

__device__ __forceinline__ ClosestPts closest_points_segments(f3 l1s, f3 l1e, f3 l2s, f3 l2e) {
    ClosestPts r;
    if (f3_distance(l1s, l1e) < 1e-5) {
        r = closest_point_segment(l1s, l2s, l2e);
    } else if (f3_distance(l2s, l2e) < 1e-5) {
        r = closest_point_segment(l2s, l1s, l1e);
        const f3 t = r.cp1; r.cp1 = r.cp2; r.cp2 = t;
    } else {
        const f3 v1 = f3_sub(l1e, l1s), v2 = f3_sub(l2e, l2s);
        const double l1 = f3_norm(v1), l2 = f3_norm(v2);
        const f3 n1 = f3_scale(v1, 1 / l1), n2 = f3_scale(v2, 1 / l2);
        if (f3_norm(f3_cross(n1, n2)) < 1e-5) {
            double bound_min = f3_dot(f3_sub(l2s, l1s), n1), bound_max = f3_dot(f3_sub(l2e, l1s), n1);
            f3 p2_min = l2s, p2_max = l2e;
            if (bound_max < bound_min) {
                const double t = bound_min; bound_min = bound_max; bound_max = t;
                const f3 tp = p2_min; p2_min = p2_max; p2_max = tp;
            }
            f3 delta = f3_sub(l2s, l1s);
            delta = f3_sub(delta, f3_scale(n1, f3_dot(delta, n1)));
            if (l1 < bound_min) { r.cp1 = l1e; r.cp2 = p2_min; }
            else if (bound_max < 0) { r.cp1 = l1s; r.cp2 = p2_max; }
            else if (bound_min < 0) { r.cp1 = l1s; r.cp2 = f3_add(l1s, delta); }
            else { r.cp1 = f3_sub(p2_min, delta); r.cp2 = p2_min; }
            r.dist = f3_distance(r.cp1, r.cp2);
        } else {
            r = closest_points_lines(l1s, l1e, l2s, l2e);
            const double alpha1 = f3_dot(f3_sub(r.cp1, l1s), n1) / l1;
            const double alpha2 = f3_dot(f3_sub(r.cp2, l2s), n2) / l2;
            if (alpha1 < 0) r.cp1 = l1s; else if (alpha1 > 1) r.cp1 = l1e;
            if (alpha2 < 0) r.cp2 = l2s; else if (alpha2 > 1) r.cp2 = l2e;
            if (alpha1 < 0 || alpha1 > 1) {
                double dot = f3_dot(n2, f3_sub(r.cp1, l2s));
                if (dot < 0) dot = 0; else if (dot > l2) dot = l2;
                r.cp2 = f3_add(l2s, f3_scale(n2, dot));
            }
            if (alpha2 < 0 || alpha2 > 1) {
                double dot = f3_dot(n1, f3_sub(r.cp2, l1s));
                if (dot < 0) dot = 0; else if (dot > l1) dot = l1;
                r.cp1 = f3_add(l1s, f3_scale(n1, dot));
            }
            r.dist = f3_distance(r.cp1, r.cp2);
        }
    }
    return r;
}

// one segment of a trajectory -- 6 control points x 3 floats = 72 bytes, 8-byte aligned -- as nine 64-bit read-only loads
__device__ __forceinline__ void load_record(const float* rec, float* out /* [18] */) {
    const float2* r2 = reinterpret_cast<const float2*>(rec);
#pragma unroll
    for (int i = 0; i < 9; i++) { const float2 v = __ldg(r2 + i); out[2 * i] = v.x; out[2 * i + 1] = v.y; }
}
// six row constants (48 bytes, 16-byte aligned) as three 128-bit stores
__device__ __forceinline__ void store_rhs(double* ro, const double* v /* [6] */) {
    double2* r2 = reinterpret_cast<double2*>(ro);
#pragma unroll
    for (int i = 0; i < 3; i++) { double2 t; t.x = v[2 * i]; t.y = v[2 * i + 1]; r2[i] = t; }
}

// ---------------------------------------------------------------------------------------------
// One (obstacle, segment) pair: the plane of the selected generator, packed (normal, rhs_i = n . p_i + d_i).
// own_traj_: the agent's initial trajectory [M][6][3] (shared or global memory); j: row of the obstacle in the CSR arrays.
template <int M>
__device__ __forceinline__ void assemble_pair(const AssembleParams& p, int agent, size_t j, int m, const float* own_traj_,
                                              double a_r, double a_dw) {
    double o_r, o_dw;
    size_t src = j;
if (p.obs_index) {
    src = (size_t) p.obs_index[j];
    o_r = (double) (float) p.all_meta[src * 2 + 0]; o_dw = (double) (float) p.all_meta[src * 2 + 1];   // agent_manager.cpp:184-199
} else { o_r = (double) p.obs_meta[j * 4 + 0]; o_dw = (double) p.obs_meta[j * 4 + 1]; }
    const double collision_dist = o_r + a_r;                                    // traj_planner.cpp:642, :661
    const double downwash = (a_dw * a_r + o_dw * o_r) / (a_r + o_r);            // downwashBetween :1229-1240
    const float dwf = (float) downwash;
    const bool transform = !(p.generator == 1 && p.dim == 2);                   // :666-672
    const float* obase = p.obs_index ? p.all_traj : p.obs_traj;
    const float* ot = obase + (src * M + m) * 18;
    const float* ogoal = p.obs_index ? p.all_goal + src * 3 : p.obs_goal + j * 3;
    const float* opos = p.obs_index ? p.all_state + src * 9 : p.obs_position + j * 3;

    f3 own[6], obs[6], own_t[6], obs_t[6];
    float orec[18];
    load_record(ot, orec);
#pragma unroll
    for (int i = 0; i < 6; i++) {
        own[i] = f3_make(own_traj_[m * 18 + i * 3], own_traj_[m * 18 + i * 3 + 1], own_traj_[m * 18 + i * 3 + 2]);
        obs[i] = f3_make(orec[i * 3], orec[i * 3 + 1], orec[i * 3 + 2]);
        own_t[i] = own[i]; obs_t[i] = obs[i];
        if (transform) {                                                        // trajectory.cpp:207-219
            own_t[i].z = __fdiv_rn(own[i].z, dwf); obs_t[i].z = __fdiv_rn(obs[i].z, dwf);
        }
    }
    f3 normal;
    double d[6];
    f3 pt[6];
    if (p.generator == 3) {                                                     // generateReciprocalRSFC :581-609
        // normalVectorBetweenLines (:1157-1177) on closestPointsBetweenLinePaths (geometry.hpp:104-127): the paths
        // first -> last control point of the obstacle and of the agent, in the original coordinates
        const f3 rs = f3_sub(own[0], obs[0]), re = f3_sub(own[5], obs[5]);
        const ClosestPts rc = closest_point_segment(f3_make(0.f, 0.f, 0.f), rs, re);
        const double len = f3_distance(rs, re);
        double alpha = 0.0;
        if (len > 0) alpha = f3_norm(f3_sub(rc.cp2, rs)) / len;
        const f3 cp1 = f3_add(obs[0], f3_scale(f3_sub(obs[5], obs[0]), alpha));
        const f3 cp2 = f3_add(own[0], f3_scale(f3_sub(own[5], own[0]), alpha));
        normal = f3_normalized(f3_sub(cp2, cp1));
        if (f3_norm(normal) == 0) {
            if (f3_norm(rs) == 0 && f3_norm(re) == 0) normal = f3_make(1.f, 0.f, 0.f);
            else normal = f3_cross(f3_sub(re, rs), f3_make(0.f, 0.f, 1.f));
        }
        normal.z = (float) ((double) normal.z / (downwash * downwash));         // :603-604
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const double size = p.obs_size ? p.obs_size[(j * M + m) * 6 + i] : o_r;
            d[i] = rc.dist < size + a_r ? 0.5 * (size + a_r + rc.dist) : size + a_r;     // :593-600
            pt[i] = obs[i];
        }
        const double nx3 = (double) normal.x, ny3 = (double) normal.y, nz3 = (double) normal.z;
        double* no3 = p.normals + (j * M + m) * 3;
        no3[0] = nx3; no3[1] = ny3; no3[2] = nz3;
        double rv3[6];
#pragma unroll
        for (int i = 0; i < 6; i++) {
            double b = nx3 * (double) pt[i].x + ny3 * (double) pt[i].y;
            if (p.dim == 3) b += nz3 * (double) pt[i].z;
            rv3[i] = b + d[i];
        }
        store_rhs(p.rhs + (j * M + m) * 6, rv3);
        return;
    }
    if (p.generator == 2) {                                                     // generateBVC :708-736
        f3 a0 = f3_make(own_traj_[0], own_traj_[1], __fdiv_rn(own_traj_[2], dwf));
        const float* o0p = obase + src * M * 18;
        f3 o0 = f3_make(o0p[0], o0p[1], __fdiv_rn(o0p[2], dwf));
        const f3 diff = f3_sub(a0, o0);
        normal = f3_normalized(diff);
        const double dd = 0.5 * (collision_dist + f3_dot(diff, normal));
#pragma unroll
        for (int i = 0; i < 6; i++) { d[i] = dd; pt[i] = obs[i]; }
    } else if (p.generator == 1 && m == M - 1) {                                // generateCLSC :691-703
        const f3 og = f3_make(ogoal[0], ogoal[1], ogoal[2]);
        const f3 ag = f3_make(p.agent_goal[agent * 3], p.agent_goal[agent * 3 + 1], p.agent_goal[agent * 3 + 2]);
        const ClosestPts cp = closest_points_segments(obs_t[5], og, own_t[5], ag);
        normal = f3_normalized(f3_sub(cp.cp2, cp.cp1));
        const double dd = 0.5 * (collision_dist + cp.dist);
#pragma unroll
        for (int i = 0; i < 6; i++) { d[i] = dd; pt[i] = cp.cp1; }
    } else {                                                                    // :625 / :678
        double rel[6][3], v[3];
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const f3 r = f3_sub(own_t[i], obs_t[i]);                            // :1186
            rel[i][0] = (double) r.x; rel[i][1] = (double) r.y; rel[i][2] = (double) r.z;
        }
        min_norm_hull6(rel, v);
        normal = f3_normalized(f3_make((float) v[0], (float) v[1], (float) v[2]));   // geometry.hpp:292, :1196
        if (p.generator == 0 && f3_norm(normal) < 1e-5) {                       // :626-634
            f3 vec = f3_sub(f3_make(p.agent_goal[agent * 3], p.agent_goal[agent * 3 + 1], p.agent_goal[agent * 3 + 2]),
                            f3_make(opos[0], opos[1], opos[2]));
            vec.z = (float) ((double) vec.z / downwash);                        // coordinateTransform :1262-1266
            normal = f3_normalized(vec);
        }
#pragma unroll
        for (int i = 0; i < 6; i++) {                                           // :640-645 / :683-686
            d[i] = 0.5 * (collision_dist + f3_dot(f3_sub(own_t[i], obs_t[i]), normal));
            pt[i] = obs[i];
        }
    }
    normal.z = (float) ((double) normal.z / downwash);                          // :653 / :689 / :701 / :730
    const double nx = (double) normal.x, ny = (double) normal.y, nz = (double) normal.z;
    double* no = p.normals + (j * M + m) * 3;
    no[0] = nx; no[1] = ny; no[2] = nz;
    double rv[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double b = nx * (double) pt[i].x + ny * (double) pt[i].y;               // traj_optimizer.cpp:414-421
        if (p.dim == 3) b += nz * (double) pt[i].z;
        rv[i] = b + d[i];
    }
    store_rhs(p.rhs + (j * M + m) * 6, rv);
}

// ---------------------------------------------------------------------------------------------
// one CTA per agent; threads stride over that agent's (obstacle, segment) pairs
#ifdef LSCQP_CUDA_EMUL
#define LSCQP_ASM_SMEM(name, count) float* name = reinterpret_cast<float*>(emu_dyn_smem)
#else
#define LSCQP_ASM_SMEM(name, count) __shared__ float name[count]
#endif

// (ptxas takes 255 registers when left alone; capped at 80 it spills ~300 bytes and three times as many CTAs fit: 0.20 -> 0.115 ms, measured --
//  the kernel is latency bound on L2 reads of the neighbours' control points)
#ifndef LSCQP_ASM_MINBLOCKS
#define LSCQP_ASM_MINBLOCKS 6
#endif
template <int M>
__global__ void __launch_bounds__(128, LSCQP_ASM_MINBLOCKS)
lsc_assemble_kernel(const AssembleParams p) {
    LSCQP_ASM_SMEM(s_own, M * 18);
#ifdef LSCQP_CUDA_EMUL
    unsigned short* s_list = reinterpret_cast<unsigned short*>(s_own + M * 18 + 32);
    int* s_cnt = reinterpret_cast<int*>(s_own + M * 18);
    double* s_x0 = reinterpret_cast<double*>(s_own + M * 18 + 2);
#else
    __shared__ unsigned short s_list[40 * M];          // (obstacle, segment) pairs that need the hull enumeration
    __shared__ int s_cnt[2];
    __shared__ double s_x0[6];                         // fixed control point c[0][2], then vmax dt / 5
#endif
    const int agent = blockIdx.x;
    if (agent >= p.n_agents) return;
    for (int e = threadIdx.x; e < M * 18; e += blockDim.x) s_own[e] = p.own_traj[(size_t) agent * M * 18 + e];
    const int obs0 = p.obs_offsets[agent], K = p.obs_offsets[agent + 1] - obs0;
    const double a_r = p.agent_meta[agent * 2 + 0], a_dw = p.agent_meta[agent * 2 + 1];
    const bool prune = p.prune && p.dim == 3 && p.generator < 2 && K * M <= 40 * M;
    if (prune && threadIdx.x < 3) {
        // the same fixed third control point and velocity step the solve kernel's presolve uses
        const int k = threadIdx.x;
        const double pos = (double) p.state[agent * 9 + k], vel = (double) p.state[agent * 9 + 3 + k],
                     acc = (double) p.state[agent * 9 + 6 + k];
        const double c1 = pos + vel * p.dt / 5.0;
        s_x0[k] = acc * p.dt * p.dt / 20.0 + 2.0 * c1 - pos;
        s_x0[3 + k] = p.limits[agent * 8 + k] * p.dt / 5.0;
    }
    if (threadIdx.x == 0) s_cnt[0] = 0;
    __syncthreads();

    auto obstacle = [&](size_t j, double& o_r, double& o_dw) -> size_t {
        if (p.obs_index) {
            const size_t src = (size_t) p.obs_index[j];
            o_r = (double) (float) p.all_meta[src * 2 + 0]; o_dw = (double) (float) p.all_meta[src * 2 + 1];   // agent_manager.cpp:184-199
            return src;
        }
        o_r = (double) p.obs_meta[j * 4 + 0]; o_dw = (double) p.obs_meta[j * 4 + 1];
        return j;
    };

    int n_work = K * M;
    if (prune) {
        // Phase 1: a pair (oi, m) is dropped when  min_i u . r_i  >  R + 2 e_max  for the unit vector u along the
        // centroid of the relative control points r_i (any unit u bounds the hull's distance to the origin from below):
        // with n the hull's unit normal, every row reads  1/2 n . r_i + n . e >= 1/2 R  where e = T(c - c0_i) is the
        // move of the control point away from the initial trajectory, and |e| <= |T(x0 - c0_i)| + steps |T(vmax dt/5)|
        // for every c the velocity rows allow (T: z / downwash).  A margin of 1e-3 covers the float roundings.
        for (int e = threadIdx.x; e < K * M; e += blockDim.x) {
            const int oi = e / M, m = e % M;
            const size_t j = (size_t) obs0 + oi;
            double o_r, o_dw;
            const size_t src = obstacle(j, o_r, o_dw);
            bool drop = false;
            if (!(p.generator == 1 && m == M - 1)) {
                const double downwash = (a_dw * a_r + o_dw * o_r) / (a_r + o_r);
                const double iz = 1.0 / downwash;
                const float* ot = (p.obs_index ? p.all_traj : p.obs_traj) + (src * M + m) * 18;
                float orec[18];
                load_record(ot, orec);
                double r[6][3], cx = 0, cy = 0, cz = 0, emax = 0;
                // (square roots in float: their 6e-8 relative error on lengths of a few metres is far inside the 1e-3 margin;
                //  one division for the whole pair: min_i (r_i . c) / |c|)
                const double vstep = (double) sqrtf((float) (s_x0[3] * s_x0[3] + s_x0[4] * s_x0[4] + s_x0[5] * s_x0[5] * iz * iz));
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    const double ox = (double) s_own[m * 18 + i * 3], oy = (double) s_own[m * 18 + i * 3 + 1], oz = (double) s_own[m * 18 + i * 3 + 2];
                    r[i][0] = ox - (double) orec[i * 3]; r[i][1] = oy - (double) orec[i * 3 + 1]; r[i][2] = (oz - (double) orec[i * 3 + 2]) * iz;
                    cx += r[i][0]; cy += r[i][1]; cz += r[i][2];
                    if (m == 0 && i < 3) continue;                                      // no such rows (traj_optimizer.cpp:404)
                    const double ex = s_x0[0] - ox, ey = s_x0[1] - oy, ez = (s_x0[2] - oz) * iz;
                    emax = fmax(emax, (double) sqrtf((float) (ex * ex + ey * ey + ez * ez)) + (double) (5 * m + i - 2) * vstep);
                }
                const double cn = (double) sqrtf((float) (cx * cx + cy * cy + cz * cz));
                if (cn > 0.0) {
                    double lb = 1e300;
#pragma unroll
                    for (int i = 0; i < 6; i++) lb = fmin(lb, r[i][0] * cx + r[i][1] * cy + r[i][2] * cz);
                    lb /= cn;
                    drop = lb > (o_r + a_r) + 2.0 * emax + 1e-3;
                }
            }
            if (drop) {
                double* no = p.normals + (j * M + m) * 3;
                no[0] = 0.0; no[1] = 0.0; no[2] = 0.0;
                const double zero6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                store_rhs(p.rhs + (j * M + m) * 6, zero6);
            } else {
                s_list[atomicAdd(&s_cnt[0], 1)] = (unsigned short) e;      // (order does not matter: pairs are independent)
            }
        }
        __syncthreads();
        n_work = s_cnt[0];
    }

    for (int w = threadIdx.x; w < n_work; w += blockDim.x) {
        const int e = prune ? (int) s_list[w] : w;
        assemble_pair<M>(p, agent, (size_t) obs0 + e / M, e % M, s_own, a_r, a_dw);
    }
}

// first half of the split dispatch (lscqp_assemble_lsc_fused at throughput batch sizes): the pruning test alone, one CTA per
// agent, no shared memory and no CTA barrier -- every warp runs on its own: the agent's constants and trajectory come
// through L1 (broadcast loads), a dropped pair is zeroed at once, the surviving pairs of a warp are appended to the global
// work list with one atomicAdd per warp (ballot + prefix).  Without the hull enumeration in the same kernel the register
// budget allows twice the resident warps, which is what hides the cold-cache latency of the neighbours' records.
template <int M>
__global__ void __launch_bounds__(128, 8)
lsc_prune_kernel(const AssembleParams p) {
    const int agent = blockIdx.x;
    if (agent >= p.n_agents) return;
    const int lane = threadIdx.x & 31;
    const int obs0 = p.obs_offsets[agent], K = p.obs_offsets[agent + 1] - obs0;
    const double a_r = p.agent_meta[agent * 2 + 0], a_dw = p.agent_meta[agent * 2 + 1];
    // the same fixed third control point and velocity step the solve kernels' presolve uses
    double x0[3], vl[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double pos = (double) p.state[agent * 9 + k], vel = (double) p.state[agent * 9 + 3 + k], acc = (double) p.state[agent * 9 + 6 + k];
        const double c1 = pos + vel * p.dt / 5.0;
        x0[k] = acc * p.dt * p.dt / 20.0 + 2.0 * c1 - pos;
        vl[k] = p.limits[agent * 8 + k] * p.dt / 5.0;
    }
    const float* own_traj = p.own_traj + (size_t) agent * M * 18;
    for (int e0 = 0; e0 < K * M; e0 += blockDim.x) {
        const int e = e0 + threadIdx.x;
        const bool valid = e < K * M;
        bool drop = false;
        size_t j = 0;
        int m = 0;
        if (valid) {
            const int oi = e / M;
            m = e % M;
            j = (size_t) obs0 + oi;
            double o_r, o_dw;
            size_t src = j;
            if (p.obs_index) {
                src = (size_t) p.obs_index[j];
                o_r = (double) (float) p.all_meta[src * 2 + 0]; o_dw = (double) (float) p.all_meta[src * 2 + 1];   // agent_manager.cpp:184-199
            } else { o_r = (double) p.obs_meta[j * 4 + 0]; o_dw = (double) p.obs_meta[j * 4 + 1]; }
            if (!(p.generator == 1 && m == M - 1)) {
                // the sufficient condition of lsc_assemble_kernel's phase 1 (see there)
                const double downwash = (a_dw * a_r + o_dw * o_r) / (a_r + o_r);
                const double iz = 1.0 / downwash;
                float orec[18], arec[18];
                load_record((p.obs_index ? p.all_traj : p.obs_traj) + (src * M + m) * 18, orec);
                load_record(own_traj + m * 18, arec);
                double r[6][3], cx = 0, cy = 0, cz = 0, emax = 0;
                const double vstep = (double) sqrtf((float) (vl[0] * vl[0] + vl[1] * vl[1] + vl[2] * vl[2] * iz * iz));
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    const double ox = (double) arec[i * 3], oy = (double) arec[i * 3 + 1], oz = (double) arec[i * 3 + 2];
                    r[i][0] = ox - (double) orec[i * 3]; r[i][1] = oy - (double) orec[i * 3 + 1]; r[i][2] = (oz - (double) orec[i * 3 + 2]) * iz;
                    cx += r[i][0]; cy += r[i][1]; cz += r[i][2];
                    if (m == 0 && i < 3) continue;                                      // no such rows (traj_optimizer.cpp:404)
                    const double ex = x0[0] - ox, ey = x0[1] - oy, ez = (x0[2] - oz) * iz;
                    emax = fmax(emax, (double) sqrtf((float) (ex * ex + ey * ey + ez * ez)) + (double) (5 * m + i - 2) * vstep);
                }
                const double cn = (double) sqrtf((float) (cx * cx + cy * cy + cz * cz));
                if (cn > 0.0) {
                    double lb = 1e300;
#pragma unroll
                    for (int i = 0; i < 6; i++) lb = fmin(lb, r[i][0] * cx + r[i][1] * cy + r[i][2] * cz);
                    lb /= cn;
                    drop = lb > (o_r + a_r) + 2.0 * emax + 1e-3;
                }
            }
            if (drop) {
                double* no = p.normals + (j * M + m) * 3;
                no[0] = 0.0; no[1] = 0.0; no[2] = 0.0;
                const double zero6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                store_rhs(p.rhs + (j * M + m) * 6, zero6);
            }
        }
        const unsigned keep = __ballot_sync(0xffffffffu, valid && !drop);
        if (keep) {
            const int leader = __ffs(keep) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(p.work_count, __popc(keep));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (valid && !drop) {
                int2 item; item.x = agent; item.y = (int) j * M + m;
                p.work_list[base + __popc(keep & ((1u << lane) - 1u))] = item;      // (order does not matter: pairs are independent)
            }
        }
    }
}

// second half of the split dispatch: one thread per surviving (obstacle, segment) pair of the whole batch
template <int M>
__global__ void __launch_bounds__(128, 4)
lsc_pairs_kernel(const AssembleParams p) {
    const int n = *p.work_count;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n; w += gridDim.x * blockDim.x) {
        const int2 item = p.work_list[w];
        const int agent = item.x;
        assemble_pair<M>(p, agent, (size_t) (item.y / M), item.y % M, p.own_traj + (size_t) agent * M * 18,
                         p.agent_meta[agent * 2 + 0], p.agent_meta[agent * 2 + 1]);
    }
}

// ---------------------------------------------------------------------------------------------
// obstacle gather for batches whose obstacles are the batch's own agents
// (MultiSyncSimulator::broadcastMsgs, src/multi_sync_simulator.cpp:305-352; AgentManager::getAgent
// src/agent_manager.cpp:184-199 narrows radius / downwash to float)
struct GatherParams {
    int n_obs, M;
    const int* obs_index;
    const float* own_traj; const double* agent_meta; const float* agent_goal; const float* state;
    float* obs_traj; float* obs_meta; float* obs_goal; float* obs_position;
};

__global__ void __launch_bounds__(256)
gather_obstacles_kernel(const GatherParams p) {
    const int per = p.M * 18;
    const size_t total = (size_t) p.n_obs * per;
    for (size_t e = (size_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t) gridDim.x * blockDim.x) {
        const int j = (int) (e / per), r = (int) (e % per);
        const int src = p.obs_index[j];
        p.obs_traj[e] = p.own_traj[(size_t) src * per + r];
        if (r < 4) p.obs_meta[(size_t) j * 4 + r] = r < 2 ? (float) p.agent_meta[src * 2 + r] : (r == 2 ? 1.0f : 0.0f);
        else if (r < 7) p.obs_goal[(size_t) j * 3 + (r - 4)] = p.agent_goal[src * 3 + (r - 4)];
        else if (r < 10) p.obs_position[(size_t) j * 3 + (r - 7)] = p.state[src * 9 + (r - 7)];
    }
}

}  // namespace lscqp
