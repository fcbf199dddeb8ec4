/* oracle/lscqp_oracle.c -- CPU restatement of the reference's agent-QP hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see lscqp_oracle.h for the pin status).  Every
 * function cites the reference file:line it follows.  Arithmetic on point3d
 * values is done in float exactly where the reference holds octomap::point3d
 * (octomath::Vector3, octomap 1.9 math/Vector3.h -- a third-party dependency that
 * is not vendored under /root/reference; its published semantics are restated in
 * the v3_* helpers below).
 */
#include "lscqp_oracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

#define SP_EPSILON       1e-9      /* include/sp_const.hpp:3 */
#define SP_EPSILON_FLOAT 1e-5      /* include/sp_const.hpp:4 */
#define ORC_INF          1e30

/* ---------- octomath::Vector3 semantics (float storage) -------------------- */
typedef struct { float x, y, z; } v3;
static v3 v3_make(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static v3 v3_load(const float *p) { return v3_make(p[0], p[1], p[2]); }
static void v3_store(float *p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
static v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
/* Vector3::operator*(float): a double argument is narrowed to float first */
static v3 v3_scale(v3 a, double s) { float f = (float) s; return v3_make(a.x * f, a.y * f, a.z * f); }
/* Vector3::dot / norm_sq: float arithmetic, widened on return */
static double v3_dot(v3 a, v3 b) { float r = a.x * b.x + a.y * b.y + a.z * b.z; return (double) r; }
static double v3_norm(v3 a) { float r = a.x * a.x + a.y * a.y + a.z * a.z; return sqrt((double) r); }
/* Vector3::distance: differences and squares in double */
static double v3_distance(v3 a, v3 b) {
    double dx = (double) a.x - (double) b.x, dy = (double) a.y - (double) b.y, dz = (double) a.z - (double) b.z;
    return sqrt(dx * dx + dy * dy + dz * dz);
}
/* Vector3::normalized: len = norm(); if (len > 0) *this /= (float) len */
static v3 v3_normalized(v3 a) {
    double len = v3_norm(a);
    if (len > 0) { float f = (float) len; return v3_make(a.x / f, a.y / f, a.z / f); }
    return a;
}
static v3 v3_cross(v3 a, v3 b) {
    return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static int v3_eq(v3 a, v3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

/* ---------- include/polynomial.hpp ----------------------------------------- */
int orc_nchoosek(int n, int k) {               /* polynomial.hpp:9-20 */
    if (k > n) return 0;
    if (k * 2 > n) k = n - k;
    if (k == 0) return 1;
    int result = n;
    for (int i = 2; i <= k; i++) { result *= (n - i + 1); result /= i; }
    return result;
}

int orc_coef_derivative(int n, int phi) {      /* polynomial.hpp:90-100 */
    if (n < phi) return 0;
    int coef = 1;
    for (int i = 0; i < phi; i++) coef *= n - i;
    return coef;
}

void orc_bernstein_basis(int n, double *B) {   /* polynomial.hpp:281-294 */
    int N = n + 1;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
            B[i * N + j] = (j >= i) ? orc_nchoosek(n, i) * orc_nchoosek(n - i, n - j) * pow(-1, j - i) : 0.0;
}

/* ---------- src/traj_optimizer.cpp constants -------------------------------- */
void orc_build_qbase(int n, int phi, int phi_n, double dt, double *Q) {   /* :163-178 */
    int N = n + 1;
    double *B = (double *) malloc(sizeof(double) * N * N);
    double *Z = (double *) malloc(sizeof(double) * N * N);
    double *T = (double *) malloc(sizeof(double) * N * N);
    orc_bernstein_basis(n, B);
    memset(Q, 0, sizeof(double) * N * N);
    for (int k = phi; k > phi - phi_n; k--) {
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) {
                Z[i * N + j] = 0;
                if (i + j - 2 * k + 1 > 0)
                    Z[i * N + j] = (double) orc_coef_derivative(i, k) * orc_coef_derivative(j, k) / (i + j - 2 * k + 1);
            }
        /* Z = B * Z * B^T */
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) {
                double s = 0;
                for (int l = 0; l < N; l++) s += B[i * N + l] * Z[l * N + j];
                T[i * N + j] = s;
            }
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) {
                double s = 0;
                for (int l = 0; l < N; l++) s += T[i * N + l] * B[j * N + l];
                Q[i * N + j] += s * pow(dt, -2 * k + 1);
            }
    }
    free(B); free(Z); free(T);
}

static const double A0_n5[6][6] = {            /* :185-190 */
    { 1,  0,   0,  0,  0, 0}, {-1,  1,   0,  0,  0, 0}, { 1, -2,   1,  0,  0, 0},
    {-1,  3,  -3,  1,  0, 0}, { 1, -4,   6, -4,  1, 0}, {-1,  5, -10, 10, -5, 1}};
static const double AT_n5[6][6] = {            /* :192-197 */
    { 0,  0,   0,  0,  0, 1}, { 0,  0,   0,  0, -1, 1}, { 0,  0,   0,  1, -2, 1},
    { 0,  0,  -1,  3, -3, 1}, { 0,  1,  -4,  6, -4, 1}, {-1,  5, -10, 10, -5, 1}};

int orc_build_aeq_base(int M, int n, int phi, double dt, double *Aeq) {   /* :180-214 */
    if (!(n == 5 && phi == 3)) return -1;      /* :198-201 throws */
    int cols = M * (n + 1);
    int rows = (M - 2) * phi;
    if (rows > 0) memset(Aeq, 0, sizeof(double) * rows * cols);
    for (int m = 2; m < M; m++) {
        int nn = 1;
        for (int j = 0; j < phi; j++) {
            for (int c = 0; c < n + 1; c++) {
                Aeq[(phi * (m - 2) + j) * cols + (n + 1) * (m - 1) + c] = pow(dt, -j) * nn * AT_n5[j][c];
                Aeq[(phi * (m - 2) + j) * cols + (n + 1) * m + c] = -pow(dt, -j) * nn * A0_n5[j][c];
            }
            nn = nn * (n - j);
        }
    }
    return 0;
}

int orc_terminal_segments(const orc_config *cfg, const orc_agent *ag) {   /* :530-538 */
    v3 g = v3_load(ag->current_goal_point), p = v3_load(ag->position);
    double ideal_flight_time = v3_norm(v3_sub(g, p)) / ag->nominal_velocity;
    int ts = (int) ((cfg->M * cfg->dt - ideal_flight_time + SP_EPSILON) / cfg->dt);
    return ts > 1 ? ts : 1;
}

/* ---------- populatebyrow, src/traj_optimizer.cpp:216-514 ------------------- */
static int lsc_row_active(const float *nrm) {  /* :409-411, Vector3::norm() */
    return !(v3_norm(v3_load(nrm)) < SP_EPSILON_FLOAT);
}

void orc_qp_sizes(const orc_config *cfg, int K, const float *lsc_normal, int *nv, int *ne, int *ni) {
    int M = cfg->M, n = cfg->n, phi = cfg->phi, dim = cfg->dim;
    *nv = dim * M * (n + 1);
    /* slack variables, :272-283: one per (obstacle, segment) in SlackMode::COLLISIONCONSTRAINT, which is what
     * mode reciprocal_rsfc selects (src/param.cpp:157-161); dynamic_obstacle_indices is never populated */
    if (cfg->planner_mode == ORC_MODE_RECIPROCALRSFC) *nv += K * M;
    int e = dim * 6 + dim * (M - 2) * phi;                       /* :319-368 */
    if (cfg->planner_mode == ORC_MODE_LSC) e += dim * (phi - 1); /* :504-511 */
    *ne = e;
    int r = 0;
    if (cfg->use_sfc) r += 2 * dim * (M * (n + 1) - phi);        /* :372-397 */
    for (int oi = 0; oi < K; oi++)                               /* :401-432 */
        for (int m = 0; m < M; m++)
            for (int i = 0; i < n + 1; i++) {
                if (m == 0 && i < phi) continue;
                if (!lsc_normal || lsc_row_active(lsc_normal + (((size_t) oi * M + m) * (n + 1) + i) * 3)) r++;
            }
    r += 2 * dim * (M * n - 2);                                  /* :443-454 */
    r += 2 * dim * (M * (n - 1) - 1);                            /* :457-472 */
    if (cfg->comm_range > 0) r += dim * M * (M + 1) + 2 * dim * M;   /* :478-500 */
    *ni = r;
}

int orc_qp_build(const orc_config *cfg, const orc_agent *ag, int K,
                 const float *lsc_point, const float *lsc_normal, const double *lsc_d,
                 const float *sfc,
                 double *P, double *q, double *c0,
                 double *Aeq, double *beq,
                 double *G, double *rlo, double *rhi,
                 double *lb, double *ub) {
    const int M = cfg->M, n = cfg->n, phi = cfg->phi, dim = cfg->dim;
    const double dt = cfg->dt;
    if (!(n == 5 && phi == 3) || M < 2 || dim < 2 || dim > 3) return -1;
    int nv, ne, ni;
    orc_qp_sizes(cfg, K, lsc_normal, &nv, &ne, &ni);
    const int offset_seg = n + 1, offset_dim = M * (n + 1);      /* :220-221 */
    const int N = n + 1;

    memset(P, 0, sizeof(double) * nv * nv);
    memset(q, 0, sizeof(double) * nv);
    memset(Aeq, 0, sizeof(double) * ne * nv);
    memset(beq, 0, sizeof(double) * ne);
    memset(G, 0, sizeof(double) * (size_t) ni * nv);
    *c0 = 0;

    /* variables and bounds, :238-270 */
    for (int k = 0; k < dim; k++)
        for (int m = 0; m < M; m++)
            for (int i = 0; i < N; i++) {
                int row = k * offset_dim + m * offset_seg + i;
                double lower = cfg->world_min[k], upper = cfg->world_max[k];
                if (k == 2 && m == 0 && cfg->planner_mode == ORC_MODE_RECIPROCALRSFC) { lower = -100; upper = 100; }
                if (m == 0 && i < 3) { lb[row] = -ORC_INF; ub[row] = ORC_INF; }
                else { lb[row] = lower; ub[row] = upper; }
            }

    /* epsilon_slack_col_oi_m in (-inf, 0], no cost term anywhere (:272-283; slack_collision_weight is never read) */
    if (cfg->planner_mode == ORC_MODE_RECIPROCALRSFC)
        for (int j = dim * M * N; j < nv; j++) { lb[j] = -ORC_INF; ub[j] = 0.0; }

    /* cost 1: jerk, :286-299 */
    double Q[36];
    orc_build_qbase(n, phi, cfg->phi_n, dt, Q);
    for (int k = 0; k < dim; k++)
        for (int m = 0; m < M; m++)
            for (int i = 0; i < N; i++)
                for (int j = 0; j < N; j++) {
                    int row = k * offset_dim + m * offset_seg + i, col = k * offset_dim + m * offset_seg + j;
                    if (Q[i * N + j] != 0 && cfg->w_control != 0) P[row * nv + col] += cfg->w_control * Q[i * N + j];
                }
    /* cost 2: error to goal, :301-315 */
    int ts = orc_terminal_segments(cfg, ag);
    for (int m = M - ts; m < M; m++)
        for (int k = 0; k < dim; k++) {
            int v = k * offset_dim + m * offset_seg + n;
            double g = (double) ag->current_goal_point[k];
            P[v * nv + v] += cfg->w_terminal;
            q[v] += -2.0 * cfg->w_terminal * g;
            *c0 += cfg->w_terminal * g * g;
        }

    /* equalities, :319-353 */
    int r = 0;
    for (int k = 0; k < dim; k++) {
        int b0 = k * offset_dim;
        Aeq[r * nv + b0 + 0] = 1; beq[r] = (double) ag->position[k]; r++;                 /* :321 */
        if (M > 1) { Aeq[r * nv + b0 + n] = 1; Aeq[r * nv + b0 + offset_seg] = -1; r++; } /* :324-326 */
        Aeq[r * nv + b0 + 1] = pow(dt, -1) * n; Aeq[r * nv + b0 + 0] = -pow(dt, -1) * n;  /* :330-332 */
        beq[r] = (double) ag->velocity[k]; r++;
        {   double c = pow(dt, -2) * n * (n - 1);                                         /* :335-338 */
            Aeq[r * nv + b0 + 2] = c; Aeq[r * nv + b0 + 1] = -2 * c; Aeq[r * nv + b0 + 0] = c;
            beq[r] = (double) ag->acceleration[k]; r++; }
        Aeq[r * nv + b0 + offset_seg + 1] += 1; Aeq[r * nv + b0 + offset_seg + 0] += -1;  /* :341-344 */
        Aeq[r * nv + b0 + n] += -1; Aeq[r * nv + b0 + n - 1] += 1; r++;
        Aeq[r * nv + b0 + offset_seg + 2] += 1; Aeq[r * nv + b0 + offset_seg + 1] += -2;  /* :347-352 */
        Aeq[r * nv + b0 + offset_seg + 0] += 1;
        Aeq[r * nv + b0 + n] += -1; Aeq[r * nv + b0 + n - 1] += 2; Aeq[r * nv + b0 + n - 2] += -1; r++;
    }
    /* :357-368 */
    if (M > 2) {
        double *base = (double *) malloc(sizeof(double) * (M - 2) * phi * offset_dim);
        orc_build_aeq_base(M, n, phi, dt, base);
        for (int k = 0; k < dim; k++)
            for (int i = 0; i < (M - 2) * phi; i++) {
                for (int j = 0; j < offset_dim; j++)
                    if (base[i * offset_dim + j] != 0) Aeq[r * nv + k * offset_dim + j] = base[i * offset_dim + j];
                r++;
            }
        free(base);
    }
    int r_eq_tail = r;   /* LSC-mode terminal rows are appended to c last (:504-511) */

    /* inequalities */
    int g = 0;
    /* SFC, :372-397 with Box::convertToLSCs (collision_constraints.cpp:37-59) */
    if (cfg->use_sfc) {
        for (int m = 0; m < M; m++) {
            const float *bmin = sfc + m * 6, *bmax = sfc + m * 6 + 3;
            for (int f = 0; f < 2 * dim; f++) {
                int axis = f / 2;
                double nrm = (f % 2 == 0) ? 1.0 : -1.0;
                double d = (f % 2 == 0) ? (double) bmin[axis] : -(double) bmax[axis];
                for (int j = 0; j < N; j++) {
                    if (m == 0 && j < phi) continue;
                    G[(size_t) g * nv + axis * offset_dim + m * offset_seg + j] = nrm;
                    rlo[g] = d; rhi[g] = ORC_INF;   /* n.(x - 0) - d >= 0 */
                    g++;
                }
            }
        }
    }
    /* LSC, :400-437 */
    for (int oi = 0; oi < K; oi++)
        for (int m = 0; m < M; m++)
            for (int i = 0; i < N; i++) {
                if (m == 0 && i < phi) continue;
                size_t rec = ((size_t) oi * M + m) * N + i;
                const float *nr = lsc_normal + rec * 3, *pt = lsc_point + rec * 3;
                if (!lsc_row_active(nr)) continue;
                double cst = 0;
                for (int k = 0; k < dim; k++) {
                    G[(size_t) g * nv + k * offset_dim + m * offset_seg + i] = (double) nr[k];
                    cst += (double) nr[k] * (double) pt[k];
                }
                if (cfg->planner_mode == ORC_MODE_RECIPROCALRSFC)          /* expr += -(lsc.d + x[offset_slack_col + M oi + m]), :423-425 */
                    G[(size_t) g * nv + dim * M * N + M * oi + m] = -1.0;
                rlo[g] = cst + lsc_d[rec]; rhi[g] = ORC_INF;
                g++;
            }
    /* dynamic feasibility, :440-474 */
    for (int k = 0; k < dim; k++)
        for (int m = 0; m < M; m++) {
            for (int i = 0; i < n; i++) {
                if (m == 0 && (i == 0 || i == 1)) continue;
                double c = pow(dt, -1) * n;
                int a = k * offset_dim + m * offset_seg + i;
                G[(size_t) g * nv + a + 1] = c; G[(size_t) g * nv + a] = -c;
                rlo[g] = -ORC_INF; rhi[g] = ag->max_vel[k]; g++;
                G[(size_t) g * nv + a + 1] = -c; G[(size_t) g * nv + a] = c;
                rlo[g] = -ORC_INF; rhi[g] = ag->max_vel[k]; g++;
            }
            for (int i = 0; i < n - 1; i++) {
                if (m == 0 && i == 0) continue;
                double c = pow(dt, -2) * n * (n - 1);
                int a = k * offset_dim + m * offset_seg + i;
                G[(size_t) g * nv + a + 2] = c; G[(size_t) g * nv + a + 1] = -2 * c; G[(size_t) g * nv + a] = c;
                rlo[g] = -ORC_INF; rhi[g] = ag->max_acc[k]; g++;
                G[(size_t) g * nv + a + 2] = -c; G[(size_t) g * nv + a + 1] = 2 * c; G[(size_t) g * nv + a] = -c;
                rlo[g] = -ORC_INF; rhi[g] = ag->max_acc[k]; g++;
            }
        }
    /* communication range, :477-500 */
    if (cfg->comm_range > 0) {
        for (int k = 0; k < dim; k++)
            for (int mi = 0; mi < M; mi++)
                for (int m = mi; m < M; m++) {
                    int a = k * offset_dim + m * offset_seg + n, b = k * offset_dim + mi * offset_seg + 0;
                    double h = 0.5 * cfg->comm_range - ag->radius;
                    G[(size_t) g * nv + a] += 1; G[(size_t) g * nv + b] += -1; rlo[g] = -ORC_INF; rhi[g] = h; g++;
                    G[(size_t) g * nv + a] += -1; G[(size_t) g * nv + b] += 1; rlo[g] = -ORC_INF; rhi[g] = h; g++;
                }
        for (int k = 0; k < dim; k++)
            for (int m = 0; m < M; m++) {
                int a = k * offset_dim + m * offset_seg + n;
                double h = 0.5 * cfg->comm_range - SP_EPSILON_FLOAT, w = (double) ag->next_waypoint[k];
                G[(size_t) g * nv + a] = 1;  rlo[g] = -ORC_INF; rhi[g] = h + w; g++;
                G[(size_t) g * nv + a] = -1; rlo[g] = -ORC_INF; rhi[g] = h - w; g++;
            }
    }
    /* stop at the end of the horizon, :504-511 */
    r = r_eq_tail;
    if (cfg->planner_mode == ORC_MODE_LSC) {
        for (int k = 0; k < dim; k++) {
            int m = M - 1;
            for (int i = 1; i < phi; i++) {
                Aeq[r * nv + k * offset_dim + m * offset_seg + n] = 1;
                Aeq[r * nv + k * offset_dim + m * offset_seg + n - i] = -1;
                r++;
            }
        }
    }
    return (r == ne && g == ni) ? 0 : -2;
}

/* ---------- min-norm point of a small convex hull --------------------------- */
static double d3_dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* Same quantity gjk() leaves in v (src/openGJK/openGJK.cpp:674-780): the point of
 * conv{pts} nearest the origin.  Exhaustive: every vertex, edge and triangle whose
 * affine projection of the origin has non-negative barycentrics is a candidate;
 * a non-degenerate tetrahedron containing the origin gives v = 0. */
double orc_min_norm_hull(const double *pts, int npts, double *v) {
    double best = INFINITY, bv[3] = {0, 0, 0};
    for (int a = 0; a < npts; a++) {
        const double *pa = pts + 3 * a;
        double nn = d3_dot(pa, pa);
        if (nn < best) { best = nn; memcpy(bv, pa, sizeof bv); }
    }
    for (int a = 0; a < npts; a++)
        for (int b = a + 1; b < npts; b++) {
            const double *pa = pts + 3 * a, *pb = pts + 3 * b;
            double d[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
            double dd = d3_dot(d, d);
            if (dd <= 0) continue;
            double t = -d3_dot(pa, d) / dd;
            if (t <= 0 || t >= 1) continue;
            double c[3] = {pa[0] + t * d[0], pa[1] + t * d[1], pa[2] + t * d[2]};
            double nn = d3_dot(c, c);
            if (nn < best) { best = nn; memcpy(bv, c, sizeof bv); }
        }
    for (int a = 0; a < npts; a++)
        for (int b = a + 1; b < npts; b++)
            for (int c = b + 1; c < npts; c++) {
                const double *pa = pts + 3 * a, *pb = pts + 3 * b, *pc = pts + 3 * c;
                double d1[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
                double d2[3] = {pc[0] - pa[0], pc[1] - pa[1], pc[2] - pa[2]};
                double g11 = d3_dot(d1, d1), g12 = d3_dot(d1, d2), g22 = d3_dot(d2, d2);
                double det = g11 * g22 - g12 * g12;
                if (!(det > 1e-14 * g11 * g22)) continue;      /* degenerate: edges cover it */
                double r1 = -d3_dot(pa, d1), r2 = -d3_dot(pa, d2);
                double m1 = (r1 * g22 - r2 * g12) / det, m2 = (g11 * r2 - g12 * r1) / det;
                if (m1 <= 0 || m2 <= 0 || m1 + m2 >= 1) continue;
                double p[3];
                for (int t = 0; t < 3; t++) p[t] = pa[t] + m1 * d1[t] + m2 * d2[t];
                double nn = d3_dot(p, p);
                if (nn < best) { best = nn; memcpy(bv, p, sizeof bv); }
            }
    for (int a = 0; a < npts && best > 0; a++)
        for (int b = a + 1; b < npts && best > 0; b++)
            for (int c = b + 1; c < npts && best > 0; c++)
                for (int e = c + 1; e < npts && best > 0; e++) {
                    const double *pa = pts + 3 * a, *pb = pts + 3 * b, *pc = pts + 3 * c, *pe = pts + 3 * e;
                    double m[3][3], rhs[3];
                    for (int t = 0; t < 3; t++) {
                        m[t][0] = pb[t] - pa[t]; m[t][1] = pc[t] - pa[t]; m[t][2] = pe[t] - pa[t];
                        rhs[t] = -pa[t];
                    }
                    double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1])
                               - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0])
                               + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
                    double scale = sqrt(d3_dot(m[0], m[0]) + d3_dot(m[1], m[1]) + d3_dot(m[2], m[2]));
                    if (!(fabs(det) > 1e-12 * scale * scale * scale)) continue;
                    double u0 = (rhs[0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1])
                               - m[0][1] * (rhs[1] * m[2][2] - m[1][2] * rhs[2])
                               + m[0][2] * (rhs[1] * m[2][1] - m[1][1] * rhs[2])) / det;
                    double u1 = (m[0][0] * (rhs[1] * m[2][2] - m[1][2] * rhs[2])
                               - rhs[0] * (m[1][0] * m[2][2] - m[1][2] * m[2][0])
                               + m[0][2] * (m[1][0] * rhs[2] - rhs[1] * m[2][0])) / det;
                    double u2 = (m[0][0] * (m[1][1] * rhs[2] - rhs[1] * m[2][1])
                               - m[0][1] * (m[1][0] * rhs[2] - rhs[1] * m[2][0])
                               + rhs[0] * (m[1][0] * m[2][1] - m[1][1] * m[2][0])) / det;
                    if (u0 >= 0 && u1 >= 0 && u2 >= 0 && u0 + u1 + u2 <= 1) { best = 0; bv[0] = bv[1] = bv[2] = 0; }
                }
    memcpy(v, bv, sizeof bv);
    return sqrt(best);
}

/* ---------- include/geometry.hpp -------------------------------------------- */
typedef struct { double dist; v3 cp1, cp2; } closest_t;

/* geometry.hpp:67-102 */
static closest_t closest_point_segment(v3 point, v3 ls, v3 le) {
    v3 a = v3_sub(ls, point), b = v3_sub(le, point), rel;
    double dist_min;
    if (v3_eq(a, b)) { dist_min = v3_norm(a); rel = a; }
    else {
        dist_min = v3_norm(a); rel = a;
        double dist = v3_norm(b);
        if (dist_min > dist) { dist_min = dist; rel = b; }
        v3 n_line = v3_normalized(v3_sub(b, a));
        v3 c = v3_sub(a, v3_scale(n_line, v3_dot(a, n_line)));
        dist = v3_norm(c);
        if (v3_dot(v3_sub(c, a), v3_sub(c, b)) < 0 && dist_min > dist) { dist_min = dist; rel = c; }
    }
    closest_t r; r.dist = dist_min; r.cp1 = point; r.cp2 = v3_add(rel, point);
    return r;
}

/* geometry.hpp:129-172 (Eigen::Matrix3f inverse = cofactors / determinant, in float) */
static closest_t closest_points_lines(v3 l1s, v3 l1e, v3 l2s, v3 l2e) {
    closest_t r;
    v3 n1 = v3_normalized(v3_sub(l1e, l1s)), n2 = v3_normalized(v3_sub(l2e, l2s)), delta;
    if (v3_distance(n1, n2) < SP_EPSILON_FLOAT || v3_distance(n1, v3_neg(n2)) < SP_EPSILON_FLOAT) {
        delta = v3_sub(l2s, l1s);
        delta = v3_sub(delta, v3_scale(n1, v3_dot(delta, n1)));
        r.dist = v3_norm(delta); r.cp1 = l1s; r.cp2 = v3_add(l1s, delta);
    } else {
        delta = v3_sub(l2s, l1s);
        v3 n3 = v3_normalized(v3_cross(n2, n1));
        float A[3][3] = {{n1.x, -n2.x, n3.x}, {n1.y, -n2.y, n3.y}, {n1.z, -n2.z, n3.z}};
        float b[3] = {delta.x, delta.y, delta.z};
        /* Eigen compute_inverse_size3: cofactors of column 0, det by column expansion,
         * inverse(r,c) = cofactor(c,r) / det */
        float c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
        float c10 = A[2][1] * A[0][2] - A[2][2] * A[0][1];
        float c20 = A[0][1] * A[1][2] - A[0][2] * A[1][1];
        float det = (c00 * A[0][0] + c10 * A[1][0]) + c20 * A[2][0];
        float id = 1.0f / det;
        float inv[3][3];
        inv[0][0] = c00 * id; inv[0][1] = c10 * id; inv[0][2] = c20 * id;
        inv[1][0] = (A[1][2] * A[2][0] - A[1][0] * A[2][2]) * id;
        inv[1][1] = (A[2][2] * A[0][0] - A[2][0] * A[0][2]) * id;
        inv[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
        inv[2][0] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) * id;
        inv[2][1] = (A[2][0] * A[0][1] - A[2][1] * A[0][0]) * id;
        inv[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
        float al[3];
        for (int i = 0; i < 3; i++) al[i] = inv[i][0] * b[0] + inv[i][1] * b[1] + inv[i][2] * b[2];
        r.dist = fabs((double) al[2]);
        r.cp1 = v3_add(l1s, v3_scale(n1, al[0]));
        r.cp2 = v3_add(l2s, v3_scale(n2, al[1]));
    }
    return r;
}

/* geometry.hpp:174-264 */
static closest_t closest_points_segments(v3 l1s, v3 l1e, v3 l2s, v3 l2e) {
    closest_t r;
    if (v3_distance(l1s, l1e) < SP_EPSILON_FLOAT) {
        r = closest_point_segment(l1s, l2s, l2e);
    } else if (v3_distance(l2s, l2e) < SP_EPSILON_FLOAT) {
        r = closest_point_segment(l2s, l1s, l1e);
        v3 t = r.cp1; r.cp1 = r.cp2; r.cp2 = t;
    } else {
        v3 v1 = v3_sub(l1e, l1s), v2 = v3_sub(l2e, l2s);
        double l1 = v3_norm(v1), l2 = v3_norm(v2);
        v3 n1 = v3_scale(v1, 1 / l1), n2 = v3_scale(v2, 1 / l2);
        if (v3_norm(v3_cross(n1, n2)) < SP_EPSILON_FLOAT) {
            double bound_min = v3_dot(v3_sub(l2s, l1s), n1), bound_max = v3_dot(v3_sub(l2e, l1s), n1);
            v3 p2_min = l2s, p2_max = l2e;
            if (bound_max < bound_min) {
                double t = bound_min; bound_min = bound_max; bound_max = t;
                v3 tp = p2_min; p2_min = p2_max; p2_max = tp;
            }
            v3 delta = v3_sub(l2s, l1s);
            delta = v3_sub(delta, v3_scale(n1, v3_dot(delta, n1)));
            if (l1 < bound_min) { r.cp1 = l1e; r.cp2 = p2_min; }
            else if (bound_max < 0) { r.cp1 = l1s; r.cp2 = p2_max; }
            else if (bound_min < 0) { r.cp1 = l1s; r.cp2 = v3_add(l1s, delta); }
            else { r.cp1 = v3_sub(p2_min, delta); r.cp2 = p2_min; }
            r.dist = v3_distance(r.cp1, r.cp2);
        } else {
            r = closest_points_lines(l1s, l1e, l2s, l2e);
            double alpha1 = v3_dot(v3_sub(r.cp1, l1s), n1) / l1;
            double alpha2 = v3_dot(v3_sub(r.cp2, l2s), n2) / l2;
            if (alpha1 < 0) r.cp1 = l1s; else if (alpha1 > 1) r.cp1 = l1e;
            if (alpha2 < 0) r.cp2 = l2s; else if (alpha2 > 1) r.cp2 = l2e;
            if (alpha1 < 0 || alpha1 > 1) {
                double dot = v3_dot(n2, v3_sub(r.cp1, l2s));
                if (dot < 0) dot = 0; else if (dot > l2) dot = l2;
                r.cp2 = v3_add(l2s, v3_scale(n2, dot));
            }
            if (alpha2 < 0 || alpha2 > 1) {
                double dot = v3_dot(n1, v3_sub(r.cp2, l1s));
                if (dot < 0) dot = 0; else if (dot > l1) dot = l1;
                r.cp1 = v3_add(l1s, v3_scale(n1, dot));
            }
            r.dist = v3_distance(r.cp1, r.cp2);
        }
    }
    return r;
}

void orc_closest_points_segments(const float *l1s, const float *l1e, const float *l2s, const float *l2e,
                                 float *cp1, float *cp2, double *dist) {
    closest_t r = closest_points_segments(v3_load(l1s), v3_load(l1e), v3_load(l2s), v3_load(l2e));
    v3_store(cp1, r.cp1); v3_store(cp2, r.cp2); *dist = r.dist;
}

/* ---------- LSC generators, src/traj_planner.cpp ---------------------------- */
/* normalVectorBetweenPolys :1179-1205 via closestPointsBetweenPointAndConvexHull
 * (geometry.hpp:266-296): relative control points in float, widened to double for
 * the hull query (util.hpp:104-116), v narrowed to float (geometry.hpp:292), then
 * normalized() in float. */
static v3 normal_between_polys(const v3 *own, const v3 *obs, int N, double *dist_out) {
    double pts[16 * 3], v[3];
    for (int i = 0; i < N; i++) {
        v3 r = v3_sub(own[i], obs[i]);
        pts[3 * i] = r.x; pts[3 * i + 1] = r.y; pts[3 * i + 2] = r.z;
    }
    double dist = orc_min_norm_hull(pts, N, v);
    if (dist_out) *dist_out = dist;
    v3 cp2 = v3_add(v3_make(0, 0, 0), v3_make((float) v[0], (float) v[1], (float) v[2]));
    return v3_normalized(cp2);
}

/* obstacleSizePredictionWithConstAcc, src/traj_planner.cpp:321-358, for one obstacle in mode reciprocal_rsfc with
 * obs/size_prediction on: size[m][i] = radius + velocity_guard + Bernstein control points of 1/2 a (m dt + tau dt)^2 for
 * the first M_u = (int)((uncertainty_horizon + 1e-9) / dt) segments, the value at M_u dt afterwards.
 * velocity_guard = ratio |v_agent|^2 / max_acc_agent[0] (0 when use_velocity_guard is off). */
void orc_obstacle_sizes(int M, int n, double dt, double obs_radius, double obs_max_acc, double uncertainty_horizon,
                        double velocity_guard, double *size /* [M][n+1] */) {
    const int N = n + 1;
    double B[36], Binv[36];
    orc_bernstein_basis(n, B);
    /* B_inv: monomial -> Bernstein (polynomial.hpp:281-294 builds B and its inverse); B is upper triangular */
    for (int c = 0; c < N; c++) {                      /* solve X B = I row by row: Binv = B^-1 */
        for (int r = 0; r < N; r++) Binv[r * N + c] = 0;
    }
    for (int r = 0; r < N; r++)
        for (int c = 0; c < N; c++) {
            double v = (r == c) ? 1.0 : 0.0;
            for (int k = 0; k < c; k++) v -= Binv[r * N + k] * B[k * N + c];
            Binv[r * N + c] = v / B[c * N + c];
        }
    const int Mu = (int) ((uncertainty_horizon + 1e-9) / dt);
    for (int m = 0; m < M; m++)
        for (int i = 0; i < N; i++) {
            if (m < Mu) {
                double coef[3] = {0.5 * obs_max_acc * pow(m * dt, 2), obs_max_acc * m * dt * dt, 0.5 * obs_max_acc * pow(dt, 2)};
                double cp = 0;
                for (int k = 0; k < 3; k++) cp += coef[k] * Binv[k * N + i];
                size[m * N + i] = obs_radius + velocity_guard + cp;
            } else {
                size[m * N + i] = obs_radius + velocity_guard + 0.5 * obs_max_acc * pow(Mu * dt, 2);
            }
        }
}

/* optional [K][M][n+1] obstacle sizes for ORC_GEN_RSFC (set before the call, reset to NULL after; test infrastructure) */
const double *orc_rsfc_sizes = 0;

void orc_generate_lsc(const orc_config *cfg, int generator, const orc_agent *ag, double agent_downwash,
                      const float *own_traj, int K, const float *obs_traj,
                      const float *obs_radius, const float *obs_downwash,
                      const float *obs_goal, const float *obs_position,
                      float *lsc_point, float *lsc_normal, double *lsc_d) {
    const int M = cfg->M, N = cfg->n + 1;
    v3 own[16], obs[16], own_t[16], obs_t[16];
    for (int oi = 0; oi < K; oi++) {
        /* downwashBetween :1229-1240, agent-type obstacle; Obstacle::radius/downwash are float */
        double o_r = (double) obs_radius[oi], o_dw = (double) obs_downwash[oi];
        double collision_dist = o_r + ag->radius;                                  /* :661 / :642 */
        double downwash = (agent_downwash * ag->radius + o_dw * o_r) / (ag->radius + o_r);
        int transform = !(generator == ORC_GEN_CLSC && cfg->dim == 2);             /* :666-672 */
        const float *ot = obs_traj + (size_t) oi * M * N * 3;

        if (generator == ORC_GEN_RSFC) {                                           /* generateReciprocalRSFC :581-609 */
            for (int m = 0; m < M; m++) {
                for (int i = 0; i < N; i++) {
                    own[i] = v3_load(own_traj + ((size_t) m * N + i) * 3);
                    obs[i] = v3_load(ot + ((size_t) m * N + i) * 3);
                }
                /* normalVectorBetweenLines :1157-1177 on closestPointsBetweenLinePaths (geometry.hpp:104-127) */
                v3 rs = v3_sub(own[0], obs[0]), re = v3_sub(own[N - 1], obs[N - 1]);          /* rel_path = line2 - line1 */
                closest_t rc = closest_point_segment(v3_make(0, 0, 0), rs, re);
                double len = v3_distance(rs, re), alpha = 0;
                if (len > 0) alpha = v3_norm(v3_sub(rc.cp2, rs)) / len;
                v3 cp1 = v3_add(obs[0], v3_scale(v3_sub(obs[N - 1], obs[0]), alpha));
                v3 cp2 = v3_add(own[0], v3_scale(v3_sub(own[N - 1], own[0]), alpha));
                double closest_dist = rc.dist;
                v3 normal = v3_normalized(v3_sub(cp2, cp1));
                if (v3_norm(normal) == 0) {
                    if (v3_norm(rs) == 0 && v3_norm(re) == 0) normal = v3_make(1, 0, 0);
                    else normal = v3_cross(v3_sub(re, rs), v3_make(0, 0, 1));
                }
                normal.z = (float) ((double) normal.z / (downwash * downwash));   /* :603-604 */
                size_t rec0 = ((size_t) oi * M + m) * N;
                for (int i = 0; i < N; i++) {
                    /* obs_pred_sizes (orc_obstacle_sizes); without it the obstacle's radius (obs/size_prediction off) */
                    const double size = orc_rsfc_sizes ? orc_rsfc_sizes[rec0 + i] : o_r;
                    lsc_d[rec0 + i] = closest_dist < size + ag->radius ? 0.5 * (size + ag->radius + closest_dist) : size + ag->radius;
                    v3_store(lsc_point + (rec0 + i) * 3, obs[i]);
                    v3_store(lsc_normal + (rec0 + i) * 3, normal);
                }
            }
            continue;
        }
        v3 bvc_normal = v3_make(0, 0, 0); double bvc_d = 0;
        if (generator == ORC_GEN_BVC) {                                            /* :708-736 */
            v3 a0 = v3_load(own_traj), o0 = v3_load(ot);
            a0.z /= (float) downwash; o0.z /= (float) downwash;                    /* trajectory.cpp:207-219 */
            v3 diff = v3_sub(a0, o0);
            bvc_normal = v3_normalized(diff);
            bvc_d = 0.5 * (collision_dist + v3_dot(diff, bvc_normal));
            bvc_normal.z = (float) ((double) bvc_normal.z / downwash);
        }
        for (int m = 0; m < M; m++) {
            for (int i = 0; i < N; i++) {
                own[i] = v3_load(own_traj + ((size_t) m * N + i) * 3);
                obs[i] = v3_load(ot + ((size_t) m * N + i) * 3);
                own_t[i] = own[i]; obs_t[i] = obs[i];
                if (transform) { own_t[i].z /= (float) downwash; obs_t[i].z /= (float) downwash; }
            }
            size_t rec0 = ((size_t) oi * M + m) * N;
            if (generator == ORC_GEN_BVC) {
                for (int i = 0; i < N; i++) {
                    v3_store(lsc_point + (rec0 + i) * 3, obs[i]);
                    v3_store(lsc_normal + (rec0 + i) * 3, bvc_normal);
                    lsc_d[rec0 + i] = bvc_d;
                }
            } else if (generator == ORC_GEN_CLSC && m == M - 1) {                  /* :691-703 */
                v3 own_last, obs_last;
                {   /* lastPoint() of the transformed trajectories */
                    own_last = own_t[N - 1]; obs_last = obs_t[N - 1];
                }
                closest_t cp = closest_points_segments(obs_last, v3_load(obs_goal + 3 * oi),
                                                       own_last, v3_load(ag->current_goal_point));
                v3 normal = v3_normalized(v3_sub(cp.cp2, cp.cp1));
                double d = 0.5 * (collision_dist + cp.dist);
                normal.z = (float) ((double) normal.z / downwash);
                for (int i = 0; i < N; i++) {                                      /* setLSC(point) :532-539 */
                    v3_store(lsc_point + (rec0 + i) * 3, cp.cp1);
                    v3_store(lsc_normal + (rec0 + i) * 3, normal);
                    lsc_d[rec0 + i] = d;
                }
            } else {
                v3 normal = normal_between_polys(own_t, obs_t, N, 0);              /* :625 / :678 */
                if (generator == ORC_GEN_LSC && v3_norm(normal) < SP_EPSILON_FLOAT) {   /* :626-634 */
                    v3 vec = v3_sub(v3_load(ag->current_goal_point), v3_load(obs_position + 3 * oi));
                    vec.z = (float) ((double) vec.z / downwash);                   /* coordinateTransform :1262-1266 */
                    normal = v3_normalized(vec);
                }
                for (int i = 0; i < N; i++) {                                      /* :640-645 / :683-686 */
                    lsc_d[rec0 + i] = 0.5 * (collision_dist + v3_dot(v3_sub(own_t[i], obs_t[i]), normal));
                }
                normal.z = (float) ((double) normal.z / downwash);                 /* :653 / :689 */
                for (int i = 0; i < N; i++) {                                      /* setLSC :514-521 */
                    v3_store(lsc_point + (rec0 + i) * 3, obs[i]);
                    v3_store(lsc_normal + (rec0 + i) * 3, normal);
                }
            }
        }
    }
}

/* ---------- closed-loop glue ------------------------------------------------ */
/* Trajectory::getPointAt src/trajectory.cpp:111-148 on a trajectory of degree deg */
static void traj_point_at(int M, int deg, double dt, const float *cps /* [M][deg+1][3] */, int stride,
                          double time, float *out) {
    v3 point = v3_make(0, 0, 0);
    int m = -1; double t_norm = 0, seg_end = 0;
    if (time < 0) { v3_store(out, point); return; }
    for (int idx = 0; idx < M; idx++) {
        seg_end += dt;
        if (time < seg_end) { m = idx; t_norm = 1 - (seg_end - time) / dt; break; }
    }
    if (m == -1) {
        if (time < seg_end + SP_EPSILON_FLOAT) { m = M - 1; t_norm = 1.0; }
        else { v3_store(out, point); return; }
    }
    for (int i = 0; i < deg + 1; i++) {
        double b = orc_nchoosek(deg, i) * pow(t_norm, i) * pow(1 - t_norm, deg - i);   /* polynomial.hpp:22-24 */
        point = v3_add(point, v3_scale(v3_load(cps + ((size_t) m * stride + i) * 3), b));
    }
    v3_store(out, point);
}

/* Trajectory::derivative src/trajectory.cpp:183-199 (float points, scale n/segment_time narrowed to float) */
static void traj_derivative(int M, int deg, double dt, const float *in, int stride, float *out) {
    for (int m = 0; m < M; m++)
        for (int i = 0; i < deg; i++) {
            v3 d = v3_sub(v3_load(in + ((size_t) m * stride + i + 1) * 3), v3_load(in + ((size_t) m * stride + i) * 3));
            v3_store(out + ((size_t) m * stride + i) * 3, v3_scale(d, deg / dt));
        }
}

void orc_get_state_at(int M, int n, double dt, const float *traj, double time, float *state9) {   /* :156-170 */
    int stride = n + 1;
    float *d1 = (float *) calloc((size_t) M * stride * 3, sizeof(float));
    float *d2 = (float *) calloc((size_t) M * stride * 3, sizeof(float));
    traj_point_at(M, n, dt, traj, stride, time, state9);
    traj_derivative(M, n, dt, traj, stride, d1);
    traj_point_at(M, n - 1, dt, d1, stride, time, state9 + 3);
    traj_derivative(M, n - 1, dt, d1, stride, d2);
    traj_point_at(M, n - 2, dt, d2, stride, time, state9 + 6);
    free(d1); free(d2);
}

void orc_shift_traj(int M, int n, const float *prev, float *out) {   /* traj_planner.cpp:287-297, 402-411 */
    int N = n + 1;
    for (int m = 0; m < M; m++)
        for (int i = 0; i < N; i++) {
            const float *src = (m == M - 1) ? prev + ((size_t) m * N + n) * 3 : prev + ((size_t) (m + 1) * N + i) * 3;
            memcpy(out + ((size_t) m * N + i) * 3, src, 3 * sizeof(float));
        }
}

void orc_const_vel_traj(int M, int n, double dt, const float *pos, const float *vel, float *out) {   /* trajectory.cpp:77-89 */
    double time = 0;
    v3 p = v3_load(pos), v = v3_load(vel);
    for (int m = 0; m < M; m++)
        for (int i = 0; i < n + 1; i++) {
            v3_store(out + ((size_t) m * (n + 1) + i) * 3, v3_add(p, v3_scale(v, time)));
            time += dt / n;
        }
}

/* ------------------------------------------------------------------------------------------------
 * GoalOptimizer (src/goal_optimizer.cpp:7-107 solve, :109-165 populatebyrow): the one-variable LP
 *     min t   s.t.  0 <= t <= 1 + SP_EPSILON_FLOAT,
 *                   n_r . ((g - w) t + w - p_r) - d_r >= 0      for every row r
 * with g = current_goal_point, w = next_waypoint (point3d: the difference g - w is a float subtraction,
 * :133-134, :152-153), rows = the 2 dim faces of the last segment's SFC box (Box::convertToLSCs,
 * collision_constraints.cpp:37-59; only with world_use_octomap, :128-143) and the LSC record
 * (oi, M-1, n) of every obstacle whose float normal is not shorter than SP_EPSILON_FLOAT (:146-162).
 * Row r is written a_r t + b_r >= 0.  Rows come out in the reference's order. */
int orc_goal_rows(const orc_config *cfg, const float *goal, const float *waypoint, int K,
                  const float *lsc_point /* [K][M][n+1][3] */, const float *lsc_normal, const double *lsc_d,
                  const float *sfc_last /* [6] box_min, box_max or NULL */, double *a, double *b) {
    int N = cfg->n + 1, r = 0;
    if (cfg->use_sfc && sfc_last) {
        /* convertToLSCs: for each k: (p=0, n=+e_k, d=box_min_k) then (p=0, n=-e_k, d=-box_max_k) */
        for (int k = 0; k < cfg->dim; k++) {
            float diff = goal[k] - waypoint[k];
            a[r] = 1.0 * (double) diff;  b[r] = 1.0 * ((double) waypoint[k] - 0.0) - (double) sfc_last[k];      r++;
            a[r] = -1.0 * (double) diff; b[r] = -1.0 * ((double) waypoint[k] - 0.0) + (double) sfc_last[3 + k]; r++;
        }
    }
    for (int oi = 0; oi < K; oi++) {
        size_t rec = ((size_t) oi * cfg->M + (cfg->M - 1)) * N + cfg->n;
        const float *nv = lsc_normal + rec * 3, *pv = lsc_point + rec * 3;
        if (v3_norm(v3_load(nv)) < SP_EPSILON_FLOAT) continue;                    /* :148-150 */
        double ar = 0, br = 0;
        for (int k = 0; k < cfg->dim; k++) {
            float diff = goal[k] - waypoint[k];
            ar += (double) nv[k] * (double) diff;
            br += (double) nv[k] * ((double) waypoint[k] - (double) pv[k]);
        }
        a[r] = ar; b[r] = br - lsc_d[rec]; r++;
    }
    return r;
}

/* Closed-form optimum of the LP above.  t* = max(0, max_{a_r > 0} -b_r / a_r); rows with a_r < 0 bound t from
 * above, rows with a_r = 0 are pure feasibility tests.  Infeasible (the reference throws QPFAILED, :94, :103) when
 * a row is violated at t* by more than feas_tol (CPLEX's default feasibility tolerance is 1e-6).
 * Early-out g ~ w (:12-14).  Output: (g - w) * (float) t + w in float arithmetic (:51).  Returns 0 ok, 2 infeasible. */
int orc_goal_solve(const orc_config *cfg, const float *goal, const float *waypoint, int nrows,
                   const double *a, const double *b, double feas_tol, float *goal_out, double *t_out) {
    v3 g = v3_load(goal), w = v3_load(waypoint);
    (void) cfg;
    if (v3_distance(g, w) < SP_EPSILON_FLOAT) { v3_store(goal_out, w); if (t_out) *t_out = 0.0; return 0; }
    double t = 0.0;
    for (int r = 0; r < nrows; r++) if (a[r] > 0 && -b[r] / a[r] > t) t = -b[r] / a[r];
    if (t > 1.0 + SP_EPSILON_FLOAT) t = 1.0 + SP_EPSILON_FLOAT;
    int status = 0;
    for (int r = 0; r < nrows; r++) if (a[r] * t + b[r] < -feas_tol) status = 2;
    if (t_out) *t_out = t;
    v3_store(goal_out, v3_add(v3_scale(v3_sub(g, w), (float) t), w));
    return status;
}

/* TrajPlanner::isSolValid src/traj_planner.cpp:990-1045 (Box::isPointInBox / isSegmentInBox
 * src/collision_constraints.cpp:81-111); the LSC part is commented out in the reference (:1012-1027).
 * state9 = desired_traj.getStateAt(multisim_time_step).  abs() there is the floating-point overload. */
int orc_is_sol_valid(const orc_config *cfg, const orc_agent *ag, const float *traj /* [M][n+1][3] */,
                     const float *state9, const float *sfc /* [M][6] or NULL */) {
    int N = cfg->n + 1;
    if (cfg->use_sfc && sfc) {
        for (int m = 0; m < cfg->M; m++)
            for (int i = (m == 0 ? cfg->phi : 0); i < N; i++) {
                const float *c = traj + ((size_t) m * N + i) * 3, *b = sfc + (size_t) m * 6;
                for (int k = 0; k < 3; k++)
                    if (!((double) c[k] > (double) b[k] - SP_EPSILON_FLOAT && (double) c[k] < (double) b[3 + k] + SP_EPSILON_FLOAT))
                        return 0;
            }
    }
    double dyn_err_tol_ratio = 0.01;
    for (int k = 0; k < cfg->dim; k++) {
        if (fabs((double) state9[3 + k]) > ag->max_vel[k] * (1 + dyn_err_tol_ratio)) return 0;
        if (fabs((double) state9[6 + k]) > ag->max_acc[k] * (1 + dyn_err_tol_ratio)) return 0;
    }
    return 1;
}
