/* oracle/pdip_cpu.c -- the same ALGORITHM CLASS as the CUDA kernel, on the host cores (SURVEY.md 8(d) baseline (ii)).
 *
 * TEST INFRASTRUCTURE ONLY (see lscqp_oracle.h): a third, independent checker solver for the restated model and the
 * "honest" CPU baseline of bench.py -- a Mehrotra predictor-corrector interior-point method in plain C, compiled with
 * -O3 -march=native, OpenMP over the agents of a batch.  It shares nothing with the kernel's formulation: the model is
 * taken exactly as populatebyrow states it (src/traj_optimizer.cpp:216-514; all variables, equalities KEPT as
 * equalities), the Newton system is the quasi-definite KKT matrix
 *        [ H + G' W G + dI    A' ] [dx]   [r1]
 *        [ A                 -dI ] [dy] = [r2]
 * (d = 1e-9) factorised by a dense LU with partial pivoting, with iterative refinement against the unregularised system.
 * The inequality rows are kept sparse (every row of the model touches at most three variables).
 *
 * orc_replan_batch_pdip runs the reference's whole per-agent path for a batch -- LSC generation (orc_generate_lsc),
 * model build (orc_qp_build), solve -- the work MultiSyncSimulator::plan does serially per agent
 * (src/multi_sync_simulator.cpp:354-362), here one agent per OpenMP task.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "lscqp_oracle.h"

#define ORC_BIG 1e29

typedef struct { int n; int idx[8]; double val[8]; double h; } srow;     /* sparse row: sum val x[idx] <= h */

/* in-place LU with partial pivoting of the (symmetric, indefinite) KKT matrix K (n x n, full storage); near the end of
 * the iteration multipliers / slacks of 1e13 sit next to the jerk Gram's 1e5, and an unpivoted LDL' of the unreduced
 * system loses pivots to cancellation -- row pivoting is what keeps this formulation alive */
static int lu_factor(double *K, int n, int *piv) {
    for (int j = 0; j < n; j++) {
        int p = j; double best = fabs(K[j * n + j]);
        for (int i = j + 1; i < n; i++) { const double v = fabs(K[i * n + j]); if (v > best) { best = v; p = i; } }
        piv[j] = p;
        if (!(best > 0.0)) return 1;
        if (p != j) for (int k = 0; k < n; k++) { const double t = K[j * n + k]; K[j * n + k] = K[p * n + k]; K[p * n + k] = t; }
        const double inv = 1.0 / K[j * n + j];
        for (int i = j + 1; i < n; i++) {
            const double l = K[i * n + j] * inv;
            if (l == 0.0) continue;
            K[i * n + j] = l;
            double *ri = K + i * n; const double *rj = K + j * n;
            for (int k = j + 1; k < n; k++) ri[k] -= l * rj[k];
        }
    }
    return 0;
}
static void lu_solve(const double *LU, const int *piv, int n, double *b) {
    for (int j = 0; j < n; j++) { const int p = piv[j]; if (p != j) { const double t = b[j]; b[j] = b[p]; b[p] = t; } }
    for (int i = 0; i < n; i++) { double v = b[i]; const double *ri = LU + i * n; for (int k = 0; k < i; k++) v -= ri[k] * b[k]; b[i] = v; }
    for (int i = n - 1; i >= 0; i--) { double v = b[i]; const double *ri = LU + i * n; for (int k = i + 1; k < n; k++) v -= ri[k] * b[k]; b[i] = v / ri[i]; }
}

/* min x'Px + q'x  s.t.  Aeq x = beq,  rlo <= G x <= rhi,  lb <= x <= ub   (dense inputs as orc_qp_build writes them).
 * info[0..3] = iterations, mean complementarity, primal residual, dual residual.  Returns 0 optimal | 1 iteration
 * limit | 3 numerical. */
int orc_pdip_solve(int nv, int ne, int ng, const double *P, const double *q, const double *Aeq, const double *beq,
                   const double *G, const double *rlo, const double *rhi, const double *lb, const double *ub,
                   double tol, int max_iter, double *x, double *info) {
    /* inequality rows g'x <= h, sparse */
    int m = 0, cap = 2 * ng + 2 * nv;
    srow *R = (srow *) malloc((size_t) cap * sizeof(srow));
    for (int i = 0; i < ng; i++) {
        srow r; r.n = 0;
        for (int j = 0; j < nv && r.n < 8; j++) if (G[(size_t) i * nv + j] != 0.0) { r.idx[r.n] = j; r.val[r.n] = G[(size_t) i * nv + j]; r.n++; }
        if (rhi[i] < ORC_BIG) { R[m] = r; R[m].h = rhi[i]; m++; }
        if (rlo[i] > -ORC_BIG) { R[m] = r; for (int k = 0; k < r.n; k++) R[m].val[k] = -r.val[k]; R[m].h = -rlo[i]; m++; }
    }
    for (int j = 0; j < nv; j++) {
        if (ub[j] < ORC_BIG) { R[m].n = 1; R[m].idx[0] = j; R[m].val[0] = 1.0; R[m].h = ub[j]; m++; }
        if (lb[j] > -ORC_BIG) { R[m].n = 1; R[m].idx[0] = j; R[m].val[0] = -1.0; R[m].h = -lb[j]; m++; }
    }
    const int N = nv + ne;
    double *K = (double *) malloc((size_t) N * N * sizeof(double));
    int *piv = (int *) malloc((size_t) N * sizeof(int));
    double *buf = (double *) calloc((size_t) (10 * m + 12 * N + nv), sizeof(double));
    double *s = buf, *z = s + m, *rg = z + m, *dsa = rg + m, *dza = dsa + m, *ds = dza + m, *dz = ds + m, *w = dz + m, *rc = w + m, *gx = rc + m;
    double *y = gx + m, *rd = y + N, *rp = rd + N, *rhs = rp + N, *sol = rhs + N, *res = sol + N, *dxa = res + N, *dx = dxa + N, *dya = dx + N, *dy = dya + N, *tmp = dy + N, *tmp2 = tmp + N;
    const double delta = 1e-12;
    memset(x, 0, (size_t) nv * sizeof(double));
    for (int r = 0; r < m; r++) { s[r] = R[r].h > 1.0 ? R[r].h : 1.0; z[r] = 1.0; }
    int status = 1, it;
    double mu = 0, rpn = 0, rdn = 0;
    for (it = 0; it < max_iter; it++) {
        /* residuals */
        for (int i = 0; i < nv; i++) { double v = q[i]; const double *pi = P + (size_t) i * nv; for (int j = 0; j < nv; j++) v += 2.0 * pi[j] * x[j]; rd[i] = v; }
        for (int e = 0; e < ne; e++) {
            const double *a = Aeq + (size_t) e * nv; double v = -beq[e];
            for (int j = 0; j < nv; j++) { v += a[j] * x[j]; rd[j] += a[j] * y[e]; }
            rp[e] = v;
        }
        double scale = 1.0;
        for (int i = 0; i < nv; i++) { const double a = fabs(q[i]); if (a > scale) scale = a; }
        mu = 0;
        double rgn = 0;
        for (int r = 0; r < m; r++) {
            double v = 0;
            for (int k = 0; k < R[r].n; k++) { v += R[r].val[k] * x[R[r].idx[k]]; rd[R[r].idx[k]] += R[r].val[k] * z[r]; }
            gx[r] = v; rg[r] = v + s[r] - R[r].h; mu += s[r] * z[r];
            if (fabs(rg[r]) > rgn) rgn = fabs(rg[r]);
        }
        mu /= m;
        rdn = 0; rpn = rgn;
        for (int i = 0; i < nv; i++) if (fabs(rd[i]) > rdn) rdn = fabs(rd[i]);
        for (int e = 0; e < ne; e++) if (fabs(rp[e]) > rpn) rpn = fabs(rp[e]);
        if (!(mu == mu) || !(rdn == rdn)) { status = 3; break; }
        if (rdn < 1e2 * tol * scale && rpn < 1e2 * tol && mu < tol) { status = 0; break; }
        /* KKT matrix */
        memset(K, 0, (size_t) N * N * sizeof(double));
        for (int i = 0; i < nv; i++) { for (int j = 0; j < nv; j++) K[(size_t) i * N + j] = 2.0 * P[(size_t) i * nv + j]; K[(size_t) i * N + i] += delta; }
        for (int r = 0; r < m; r++) {
            w[r] = z[r] / s[r];
            for (int a = 0; a < R[r].n; a++)
                for (int b = 0; b < R[r].n; b++) K[(size_t) R[r].idx[a] * N + R[r].idx[b]] += w[r] * R[r].val[a] * R[r].val[b];
        }
        for (int e = 0; e < ne; e++) {
            for (int j = 0; j < nv; j++) { K[(size_t) (nv + e) * N + j] = Aeq[(size_t) e * nv + j]; K[(size_t) j * N + nv + e] = Aeq[(size_t) e * nv + j]; }
            K[(size_t) (nv + e) * N + nv + e] = -delta;
        }
        if (lu_factor(K, N, piv)) { status = 3; break; }
        /* two directions with the same factor: rc = s z (predictor), then s z + dsa dza - sigma mu (corrector) */
        double sigma_mu = 0.0;
        for (int pass = 0; pass < 2; pass++) {
            double *ddx = pass ? dx : dxa, *ddy = pass ? dy : dya, *dds = pass ? ds : dsa, *ddz = pass ? dz : dza;
            for (int i = 0; i < nv; i++) rhs[i] = -rd[i];
            for (int e = 0; e < ne; e++) rhs[nv + e] = -rp[e];
            for (int r = 0; r < m; r++) {
                rc[r] = s[r] * z[r] + (pass ? dsa[r] * dza[r] - sigma_mu : 0.0);
                const double t = rc[r] / s[r] - w[r] * rg[r];
                for (int k = 0; k < R[r].n; k++) rhs[R[r].idx[k]] += R[r].val[k] * t;
            }
            /* solve with two steps of iterative refinement against the unregularised KKT matrix */
            memcpy(sol, rhs, (size_t) N * sizeof(double));
            lu_solve(K, piv, N, sol);
            for (int ref = 0; ref < 2; ref++) {
                /* res = rhs - Kfull sol, Kfull = [H + G'WG, A'; A, 0] */
                for (int i = 0; i < nv; i++) { double v = rhs[i]; const double *pi = P + (size_t) i * nv; for (int j = 0; j < nv; j++) v -= 2.0 * pi[j] * sol[j]; res[i] = v; }
                for (int r = 0; r < m; r++) {
                    double v = 0; for (int k = 0; k < R[r].n; k++) v += R[r].val[k] * sol[R[r].idx[k]];
                    v *= w[r];
                    for (int k = 0; k < R[r].n; k++) res[R[r].idx[k]] -= R[r].val[k] * v;
                }
                for (int e = 0; e < ne; e++) {
                    const double *a = Aeq + (size_t) e * nv; double v = rhs[nv + e];
                    for (int j = 0; j < nv; j++) { v -= a[j] * sol[j]; res[j] -= a[j] * sol[nv + e]; }
                    res[nv + e] = v;
                }
                lu_solve(K, piv, N, res);
                for (int i = 0; i < N; i++) sol[i] += res[i];
            }
            memcpy(ddx, sol, (size_t) nv * sizeof(double)); memcpy(ddy, sol + nv, (size_t) ne * sizeof(double));
            double alpha = 1.0;
            for (int r = 0; r < m; r++) {
                double v = 0; for (int k = 0; k < R[r].n; k++) v += R[r].val[k] * ddx[R[r].idx[k]];
                dds[r] = -rg[r] - v;
                ddz[r] = -(rc[r] + z[r] * dds[r]) / s[r];
                if (dds[r] < 0) { const double a = -s[r] / dds[r]; if (a < alpha) alpha = a; }
                if (ddz[r] < 0) { const double a = -z[r] / ddz[r]; if (a < alpha) alpha = a; }
            }
            if (pass == 0) {
                double mu_aff = 0;
                for (int r = 0; r < m; r++) mu_aff += (s[r] + alpha * dds[r]) * (z[r] + alpha * ddz[r]);
                mu_aff /= m;
                const double sg = mu > 0 ? mu_aff / mu : 0.0;
                sigma_mu = sg * sg * sg * mu;
            } else {
                const double a = alpha < 1.0 ? 0.995 * alpha : 1.0;
                for (int i = 0; i < nv; i++) x[i] += a * ddx[i];
                for (int e = 0; e < ne; e++) y[e] += a * ddy[e];
                for (int r = 0; r < m; r++) { s[r] += a * dds[r]; z[r] += a * ddz[r]; }
            }
        }
    }
    (void) tmp; (void) tmp2;
    if (info) { info[0] = it; info[1] = mu; info[2] = rpn; info[3] = rdn; }
    free(R); free(K); free(piv); free(buf);
    return status;
}

/* The reference's per-agent path for a whole batch (obstacles = agents of the batch, as broadcastMsgs hands them over):
 * generate the LSCs, build the model, solve it; one agent per OpenMP task on `threads` threads.  ctrl_out [n][nv],
 * status_out [n] (0 optimal).  Returns the wall time of the parallel region in seconds. */
double orc_replan_batch_pdip(const orc_config *cfg, int generator, int n_agents, const orc_agent *agents,
                             const double *agent_downwash, const float *own_traj, const int *obs_offsets, const int *obs_index,
                             const float *agent_radius_f, const float *agent_downwash_f, const float *agent_goal,
                             const float *agent_position, double *ctrl_out, int *status_out, int *iters_out, int threads) {
    const int M = cfg->M, N1 = cfg->n + 1, per = M * N1 * 3;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
    const double t0 = omp_get_wtime();
#else
    const double t0 = 0.0;
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int a = 0; a < n_agents; a++) {
        const int K = obs_offsets[a + 1] - obs_offsets[a];
        float *obs_traj = (float *) malloc((size_t) (K > 0 ? K : 1) * per * sizeof(float));
        float *orad = (float *) malloc((size_t) (K + 1) * sizeof(float)), *odw = (float *) malloc((size_t) (K + 1) * sizeof(float));
        float *ogoal = (float *) malloc((size_t) (K + 1) * 3 * sizeof(float)), *opos = (float *) malloc((size_t) (K + 1) * 3 * sizeof(float));
        for (int j = 0; j < K; j++) {
            const int o = obs_index[obs_offsets[a] + j];
            memcpy(obs_traj + (size_t) j * per, own_traj + (size_t) o * per, (size_t) per * sizeof(float));
            orad[j] = agent_radius_f[o]; odw[j] = agent_downwash_f[o];
            memcpy(ogoal + j * 3, agent_goal + o * 3, 3 * sizeof(float)); memcpy(opos + j * 3, agent_position + o * 3, 3 * sizeof(float));
        }
        float *pt = (float *) malloc((size_t) (K > 0 ? K : 1) * per * sizeof(float)), *nr = (float *) malloc((size_t) (K > 0 ? K : 1) * per * sizeof(float));
        double *dd = (double *) malloc((size_t) (K > 0 ? K : 1) * M * N1 * sizeof(double));
        orc_generate_lsc(cfg, generator, &agents[a], agent_downwash[a], own_traj + (size_t) a * per, K, obs_traj, orad, odw, ogoal, opos, pt, nr, dd);
        int nv, ne, ni;
        orc_qp_sizes(cfg, K, nr, &nv, &ne, &ni);
        double *P = (double *) calloc((size_t) nv * nv, sizeof(double)), *q = (double *) calloc(nv, sizeof(double)), c0 = 0;
        double *Aeq = (double *) calloc((size_t) ne * nv, sizeof(double)), *beq = (double *) calloc(ne, sizeof(double));
        double *G = (double *) calloc((size_t) ni * nv, sizeof(double)), *rlo = (double *) calloc(ni, sizeof(double)), *rhi = (double *) calloc(ni, sizeof(double));
        double *lb = (double *) calloc(nv, sizeof(double)), *ub = (double *) calloc(nv, sizeof(double));
        orc_qp_build(cfg, &agents[a], K, pt, nr, dd, NULL, P, q, &c0, Aeq, beq, G, rlo, rhi, lb, ub);
        double info[4];
        status_out[a] = orc_pdip_solve(nv, ne, ni, P, q, Aeq, beq, G, rlo, rhi, lb, ub, 1e-10, 100, ctrl_out + (size_t) a * nv, info);
        if (iters_out) iters_out[a] = (int) info[0];
        free(obs_traj); free(orad); free(odw); free(ogoal); free(opos); free(pt); free(nr); free(dd);
        free(P); free(q); free(Aeq); free(beq); free(G); free(rlo); free(rhi); free(lb); free(ub);
    }
#ifdef _OPENMP
    return omp_get_wtime() - t0;
#else
    return t0;
#endif
}
