/*
 * ismpc_oracle.c -- plain-C CPU restatement of the reference ISMPC hot path.
 * TEST INFRASTRUCTURE ONLY (see ismpc_oracle.h).  Deliberately literal: it
 * follows the reference's own loops (including the O(N^2) Phi recomputation and
 * the repeated-multiplication matrixPower) so that it can serve as the checker
 * for the restructured O(N) CUDA kernels.  Citations: paths under /root/reference/.
 */
#include "ismpc_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================= */
/*  Portable dense dual active-set QP solver (Goldfarb-Idnani, Schur form)  */
/* ======================================================================= */

static int chol_lower(int n, double* M) /* in-place, row-major, lower */
{
    for (int j = 0; j < n; ++j) {
        double d = M[j * n + j];
        for (int k = 0; k < j; ++k) d -= M[j * n + k] * M[j * n + k];
        if (!(d > 0.0)) return -1;
        d = sqrt(d);
        M[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = M[i * n + j];
            for (int k = 0; k < j; ++k) s -= M[i * n + k] * M[j * n + k];
            M[i * n + j] = s / d;
        }
    }
    return 0;
}

static void chol_solve(int n, const double* L, const double* b, double* x)
{
    for (int i = 0; i < n; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[i * n + k] * x[k];
        x[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = x[i];
        for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * x[k];
        x[i] = s / L[i * n + i];
    }
}

static double dotn(int n, const double* a, const double* b)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

int oracle_qp_dual_active_set(int nV, int nC, const double* H, const double* g,
                              const double* A, const double* lbA, const double* ubA,
                              double* x, double* y, int* ws, int* nwsr)
{
    const double INF = 1e300;
    int ret = 0, iters = 0;
    int maxq = nV < nC ? nV : nC;
    double* L = (double*)malloc(sizeof(double) * nV * nV);
    double* d = (double*)malloc(sizeof(double) * nV);
    double* z = (double*)malloc(sizeof(double) * nV);
    double* np = (double*)malloc(sizeof(double) * nV);
    double* D = (double*)malloc(sizeof(double) * (size_t)(maxq + 1) * nV); /* D[k] = H^-1 n_k */
    double* Sm = (double*)malloc(sizeof(double) * (size_t)(maxq + 1) * (maxq + 1));
    double* Sc = (double*)malloc(sizeof(double) * (size_t)(maxq + 1) * (maxq + 1));
    double* r = (double*)malloc(sizeof(double) * (maxq + 1));
    double* rhs = (double*)malloc(sizeof(double) * (maxq + 1));
    double* u = (double*)malloc(sizeof(double) * (maxq + 1));
    int* widx = (int*)malloc(sizeof(int) * (maxq + 1));
    int* wsgn = (int*)malloc(sizeof(int) * (maxq + 1));
    int* weq = (int*)malloc(sizeof(int) * (maxq + 1));
    int* state = (int*)calloc(nC > 0 ? nC : 1, sizeof(int)); /* 0 free, -1 lower in W, +1 upper in W, 2 skipped */
    int q = 0;

    memcpy(L, H, sizeof(double) * nV * nV);
    if (chol_lower(nV, L) != 0) { ret = -1; goto done; }
    for (int i = 0; i < nV; ++i) z[i] = -g[i];
    chol_solve(nV, L, z, x);

    int eq_cursor = 0;
    const int itmax = 20 * (nV + nC) + 50;
    for (;;) {
        /* ---- choose entering constraint p (row ip, sign sp) ---- */
        int ip = -1, sp = 0, p_is_eq = 0;
        double viol = 0.0;
        while (eq_cursor < nC) { /* equalities first, in row order (enableEqualities) */
            int i = eq_cursor++;
            double tol = 1e-12 * fmax(1.0, fabs(lbA[i]));
            if (fabs(ubA[i] - lbA[i]) <= tol && state[i] == 0) {
                double ax = dotn(nV, A + (size_t)i * nV, x);
                ip = i; p_is_eq = 1;
                sp = (ax - lbA[i] <= 0.0) ? +1 : -1;
                viol = (sp > 0) ? (ax - lbA[i]) : (ubA[i] - ax); /* <= 0 */
                break;
            }
        }
        if (ip < 0) {
            double worst = 0.0;
            for (int i = 0; i < nC; ++i) {
                if (state[i] != 0) continue;
                double ax = dotn(nV, A + (size_t)i * nV, x);
                double sl = ax - lbA[i], su = ubA[i] - ax;
                double tl = 1e-11 * fmax(1.0, fabs(lbA[i])), tu = 1e-11 * fmax(1.0, fabs(ubA[i]));
                if (lbA[i] > -1e19 && sl < -tl && sl < worst) { worst = sl; ip = i; sp = +1; }
                if (ubA[i] < 1e19 && su < -tu && su < worst) { worst = su; ip = i; sp = -1; }
            }
            viol = worst;
            if (ip < 0) break; /* primal feasible: optimal */
        }
        for (int k = 0; k < nV; ++k) np[k] = sp * A[(size_t)ip * nV + k];
        double beta = (sp > 0) ? lbA[ip] : -ubA[ip];
        double up = 0.0;
        /* ---- inner loop: dual/primal steps until p is added or problem declared infeasible ---- */
        for (;;) {
            if (++iters > itmax) { ret = -3; goto done; }
            chol_solve(nV, L, np, d);
            double npd = dotn(nV, np, d);
            double zn = npd;
            if (q > 0) {
                for (int a = 0; a < q; ++a) {
                    const double* na = A + (size_t)widx[a] * nV;
                    rhs[a] = wsgn[a] * dotn(nV, na, d);
                    for (int b = 0; b < q; ++b) Sc[a * q + b] = Sm[a * (maxq + 1) + b];
                }
                if (chol_lower(q, Sc) != 0) { ret = -4; goto done; }
                chol_solve(q, Sc, rhs, r);
                for (int k = 0; k < nV; ++k) {
                    double s = d[k];
                    for (int a = 0; a < q; ++a) s -= D[(size_t)a * nV + k] * r[a];
                    z[k] = s;
                }
                zn = dotn(nV, np, z);
            } else {
                memcpy(z, d, sizeof(double) * nV);
            }
            double t1 = INF; int l = -1;
            for (int a = 0; a < q; ++a)
                if (!weq[a] && r[a] > 1e-14) {
                    double t = u[a] / r[a];
                    if (t < t1) { t1 = t; l = a; }
                }
            /* nV independent rows span the space: whatever rounding leaves in zn, the next row depends on them */
            int dependent = q >= maxq || !(zn > 1e-13 * fmax(npd, 1e-300));
            double sviol = dotn(nV, np, x) - beta; /* <0 when violated */
            double t2 = dependent ? INF : (-sviol / zn);
            if (!dependent && t2 < 0.0) t2 = 0.0;
            if (dependent && sviol >= -1e-10 * fmax(1.0, fabs(beta))) {
                /* linearly dependent on W and already satisfied: nothing to do (e.g. zero rows) */
                state[ip] = 2;
                break;
            }
            double t = t1 < t2 ? t1 : t2;
            if (t >= INF) { ret = -2; goto done; } /* infeasible */
            if (!dependent && t > 0.0) {
                for (int k = 0; k < nV; ++k) x[k] += t * z[k];
                /* rows set aside as dependent-and-satisfied were judged at the old x: look at them again */
                for (int i = 0; i < nC; ++i) if (state[i] == 2) state[i] = 0;
            }
            for (int a = 0; a < q; ++a) u[a] -= t * r[a];
            up += t;
            if (!dependent && t2 <= t1) {
                /* full step: add p */
                widx[q] = ip; wsgn[q] = sp; weq[q] = p_is_eq; u[q] = up;
                memcpy(D + (size_t)q * nV, d, sizeof(double) * nV);
                for (int a = 0; a <= q; ++a) {
                    const double* na = A + (size_t)widx[a] * nV;
                    double v = wsgn[a] * dotn(nV, na, d);
                    Sm[a * (maxq + 1) + q] = v;
                    Sm[q * (maxq + 1) + a] = v;
                }
                state[ip] = sp > 0 ? -1 : +1;
                ++q;
                break;
            }
            /* partial step: drop l */
            state[widx[l]] = 0;
            for (int a = l; a < q - 1; ++a) {
                widx[a] = widx[a + 1]; wsgn[a] = wsgn[a + 1]; weq[a] = weq[a + 1]; u[a] = u[a + 1];
                memcpy(D + (size_t)a * nV, D + (size_t)(a + 1) * nV, sizeof(double) * nV);
            }
            for (int a = 0; a < q; ++a)
                for (int b = l; b < q - 1; ++b) Sm[a * (maxq + 1) + b] = Sm[a * (maxq + 1) + b + 1];
            for (int a = l; a < q - 1; ++a)
                for (int b = 0; b < q - 1; ++b) Sm[a * (maxq + 1) + b] = Sm[(a + 1) * (maxq + 1) + b];
            --q;
        }
        (void)viol;
    }
done:
    /* A return code of 0 must mean "this point is feasible": on infeasible problems the working set fills up to nV rows,
     * the Schur complement turns singular and rounding can carry the iteration to a point that violates rows it holds
     * active (qpOASES reports RET_INIT_FAILED_* on the same problems).  Checked on every row, equalities included. */
    if (ret == 0) {
        for (int i = 0; i < nC; ++i) {
            double ax = dotn(nV, A + (size_t)i * nV, x);
            double sc = fmax(1.0, fmax(fabs(lbA[i]) < 1e19 ? fabs(lbA[i]) : 0.0, fabs(ubA[i]) < 1e19 ? fabs(ubA[i]) : 0.0));
            if ((lbA[i] > -1e19 && lbA[i] - ax > 1e-7 * sc) || (ubA[i] < 1e19 && ax - ubA[i] > 1e-7 * sc)) { ret = -2; break; }
        }
    }
    if (ws) for (int i = 0; i < nC; ++i) ws[i] = 0;
    if (y) for (int i = 0; i < nC; ++i) y[i] = 0.0;
    for (int a = 0; a < q; ++a) {
        if (ws) ws[widx[a]] = weq[a] ? -1 : (wsgn[a] > 0 ? -1 : +1);
        if (y) y[widx[a]] = wsgn[a] * u[a];
    }
    if (nwsr) *nwsr = iters;
    free(L); free(d); free(z); free(np); free(D); free(Sm); free(Sc); free(r); free(rhs); free(u);
    free(widx); free(wsgn); free(weq); free(state);
    return ret;
}

/* ======================================================================= */
/*  Formulation C                                                          */
/* ======================================================================= */

void oracle_formc_default_params(oracle_formc_params* p)
{
    /* AMR_code_DART/parameters.cpp:9-45 */
    p->dt = 0.01; p->dtc = 0.01; p->h = 0.69; p->mass = 50.0; p->g = 9.81;
    p->box_w = 0.09; p->box_w_init = 2.0;
    p->q_p = 1005000.0; p->q_v = 100.0; p->q_u = 0.01; /* MPCSolver.cpp:253-255 */
    p->fz_max = 10000.0;                                /* MPCSolver.cpp:159 */
    p->N = 100; p->S = 35; p->F = 10;
}

/* MPCSolver.cpp:167-180 */
void oracle_formc_midpoint(const double* plan, int n_steps, int S, int F, double* mid)
{
    int per = S + F;
    memset(mid, 0, sizeof(double) * (size_t)n_steps * per * 3);
    for (int i = 0; i < n_steps - 1; ++i) {
        for (int c = 0; c < 3; ++c) {
            double a = plan[i * 4 + c], b = plan[(i + 1) * 4 + c];
            for (int r = 0; r < S; ++r) mid[((size_t)i * per + r) * 3 + c] = a * 1.0;
            for (int k = 0; k < F; ++k) {
                double tr = (double)k / (double)F;            /* :170 */
                mid[((size_t)i * per + S + k) * 3 + c] = a * 1.0 + (b - a) * tr; /* :177-179 */
            }
        }
    }
}

/* utils.cpp:73-81 applied to A_z = [1 dt; 0 1] (MPCSolver.cpp:125): repeated multiplication. */
static void az_power(double dt, int e, double M[4])
{
    M[0] = 1; M[1] = 0; M[2] = 0; M[3] = 1;
    for (int i = 0; i < e; ++i) {
        double r0 = M[0] * 1.0 + M[1] * 0.0, r1 = M[0] * dt + M[1] * 1.0;
        double r2 = M[2] * 1.0 + M[3] * 0.0, r3 = M[2] * dt + M[3] * 1.0;
        M[0] = r0; M[1] = r1; M[2] = r2; M[3] = r3;
    }
}

/* MPCSolver.cpp:124-156 */
void oracle_formc_vertical_matrices(const oracle_formc_params* p, double* Sz, double* Szv,
                                    double* Tz, double* Tzv, double* Tg, double* Tgv)
{
    int N = p->N;
    double Bz[2] = {0.0, p->dt / p->mass};  /* :126 */
    double Bg[2] = {0.0, -p->dt};           /* :127 */
    double* Sgz = (double*)calloc((size_t)N * N, sizeof(double));
    double* Sgzv = (double*)calloc((size_t)N * N, sizeof(double));
    memset(Sz, 0, sizeof(double) * N * N);
    memset(Szv, 0, sizeof(double) * N * N);
    double M[4];
    for (int k = 0; k < N; ++k) {
        az_power(p->dt, k + 1, M);
        Tz[k * 2 + 0] = M[0]; Tz[k * 2 + 1] = M[1];       /* C_z = [1 0]  :146 */
        Tzv[k * 2 + 0] = M[2]; Tzv[k * 2 + 1] = M[3];     /* C_v = [0 1]  :147 */
        for (int j = 0; j < k; ++j) {
            az_power(p->dt, k - j, M);
            Sz[k * N + j] = M[0] * Bz[0] + M[1] * Bz[1];   /* :149 */
            Szv[k * N + j] = M[2] * Bz[0] + M[3] * Bz[1];  /* :150 */
            Sgz[k * N + j] = M[0] * Bg[0] + M[1] * Bg[1];  /* :151 */
            Sgzv[k * N + j] = M[2] * Bg[0] + M[3] * Bg[1]; /* :152 */
        }
    }
    for (int k = 0; k < N; ++k) {  /* T_bar_g = S_bar_g * p * g  :155-156 */
        double s = 0, sv = 0;
        for (int j = 0; j < N; ++j) { s += Sgz[k * N + j] * 1.0; sv += Sgzv[k * N + j] * 1.0; }
        Tg[k] = s * p->g; Tgv[k] = sv * p->g;
    }
    free(Sgz); free(Sgzv);
}

/* The reference builds these matrices ONCE, in the constructor (MPCSolver.cpp:144-156), and every solve() reuses the
 * members.  The per-tick functions below therefore take them from a per-model cache (keyed by what they depend on: N,
 * dt, mass, g) instead of rebuilding them each tick -- what stays per tick is what solve() itself recomputes, notably
 * H_z (MPCSolver.cpp:258).  Entries are immutable once published; the lock covers look-up and construction. */
#include <pthread.h>
typedef struct { int N; double dt, mass, g; double *Sz, *Szv, *Tz, *Tzv, *Tg, *Tgv; } formc_vm_entry;
#define FORMC_VM_SLOTS 16
static formc_vm_entry g_vm[FORMC_VM_SLOTS];
static int g_vm_n = 0;
static pthread_mutex_t g_vm_lock = PTHREAD_MUTEX_INITIALIZER;

static const formc_vm_entry* formc_vertical_cached(const oracle_formc_params* p)
{
    const formc_vm_entry* hit = NULL;
    pthread_mutex_lock(&g_vm_lock);
    for (int i = 0; i < g_vm_n; ++i)
        if (g_vm[i].N == p->N && g_vm[i].dt == p->dt && g_vm[i].mass == p->mass && g_vm[i].g == p->g) { hit = &g_vm[i]; break; }
    if (!hit) {
        formc_vm_entry* e = &g_vm[g_vm_n < FORMC_VM_SLOTS ? g_vm_n : FORMC_VM_SLOTS - 1];
        if (g_vm_n >= FORMC_VM_SLOTS) { free(e->Sz); free(e->Szv); free(e->Tz); free(e->Tzv); free(e->Tg); free(e->Tgv); }
        else ++g_vm_n;
        int N = p->N;
        e->N = N; e->dt = p->dt; e->mass = p->mass; e->g = p->g;
        e->Sz = (double*)malloc(sizeof(double) * N * N); e->Szv = (double*)malloc(sizeof(double) * N * N);
        e->Tz = (double*)malloc(sizeof(double) * N * 2); e->Tzv = (double*)malloc(sizeof(double) * N * 2);
        e->Tg = (double*)malloc(sizeof(double) * N); e->Tgv = (double*)malloc(sizeof(double) * N);
        oracle_formc_vertical_matrices(p, e->Sz, e->Szv, e->Tz, e->Tzv, e->Tg, e->Tgv);
        hit = e;
    }
    pthread_mutex_unlock(&g_vm_lock);
    return hit;
}

/* number of flight-phase equality rows, MPCSolver.cpp:223-229 */
static int formc_ne(const oracle_formc_params* p, int mpc_iter)
{
    int ne = (mpc_iter < p->S) ? p->F : (p->S + p->F - mpc_iter);
    return ne < 0 ? 0 : ne;
}

/* MPCSolver.cpp:220-269 */
int oracle_formc_vertical_qp(const oracle_formc_params* p, const double z0[2], const double* mid_z,
                             int mpc_iter, int footstep_counter,
                             double* H, double* gq, double* A, double* lbA, double* ubA, int* ne_out)
{
    int N = p->N;
    const formc_vm_entry* vm = formc_vertical_cached(p);          /* members built by the constructor, :144-156 */
    const double *Sz = vm->Sz, *Szv = vm->Szv, *Tz = vm->Tz, *Tzv = vm->Tzv, *Tg = vm->Tg, *Tgv = vm->Tgv;
    double* v = (double*)malloc(sizeof(double) * N);
    double* w = (double*)malloc(sizeof(double) * N);
    /* H_z, :258 (rebuilt every tick by the reference, although constant) */
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s1 = 0, s2 = 0;
            for (int k = 0; k < N; ++k) { s1 += Sz[k * N + i] * Sz[k * N + j]; s2 += Szv[k * N + i] * Szv[k * N + j]; }
            H[i * N + j] = p->q_p * s1 + p->q_v * s2 + (i == j ? p->q_u : 0.0);
        }
    /* F_z, :259 */
    for (int k = 0; k < N; ++k) {
        v[k] = Tz[k * 2] * z0[0] + Tz[k * 2 + 1] * z0[1] + Tg[k] - 1.0 * p->h - mid_z[k];
        w[k] = Tzv[k * 2] * z0[0] + Tzv[k * 2 + 1] * z0[1] + Tgv[k];
    }
    for (int j = 0; j < N; ++j) {
        double s1 = 0, s2 = 0;
        for (int k = 0; k < N; ++k) { s1 += Sz[k * N + j] * v[k]; s2 += Szv[k * N + j] * w[k]; }
        gq[j] = p->q_p * s1 + p->q_v * s2 + p->q_u * (-1.0 * p->mass * p->g);
    }
    /* equality rows, :223-243, scaled by is_running (:262-269): when not running the block is all-zero
     * with zero rhs -> dropped from the stack (qpOASES demotes zero equality rows). */
    int is_running = footstep_counter > 1;
    int ne = is_running ? formc_ne(p, mpc_iter) : 0;
    memset(A, 0, sizeof(double) * (size_t)(ne + N) * N);
    for (int i = 0; i < N; ++i) {
        if (!is_running) break;
        if (mpc_iter < p->S) {
            if (i >= p->S && i < p->S + p->F) {
                int r = i - p->S, c = i - mpc_iter;           /* :235 (mixed index shift, as written) */
                if (r < ne && c >= 0 && c < N) A[r * N + c] = 1.0;
            }
        } else {
            if (i < p->S + p->F - mpc_iter) A[i * N + i] = 1.0; /* :240 */
        }
    }
    for (int r = 0; r < ne; ++r) { lbA[r] = 0.0; ubA[r] = 0.0; }
    for (int k = 0; k < N; ++k) {                              /* Aineq_z = S_bar_z, 0..1e4  :158-160 */
        memcpy(A + (size_t)(ne + k) * N, Sz + (size_t)k * N, sizeof(double) * N);
        lbA[ne + k] = 0.0; ubA[ne + k] = p->fz_max;
    }
    if (ne_out) *ne_out = ne;
    free(v); free(w);
    return ne + N;
}

/* MPCSolver.cpp:296-309 */
void oracle_formc_lambda(const oracle_formc_params* p, const double z0[2], const double* f,
                         double* lambda, double* zpos)
{
    int N = p->N;
    const formc_vm_entry* vm = formc_vertical_cached(p);          /* members built by the constructor, :144-156 */
    const double *Sz = vm->Sz, *Tz = vm->Tz, *Tg = vm->Tg;
    for (int j = 0; j < N; ++j) {
        double zacc = (1.0 / p->mass) * f[j] - 1.0 * p->g;      /* :296 */
        double zp = 0;
        for (int k = 0; k < N; ++k) zp += Sz[j * N + k] * f[k];
        zp += Tz[j * 2] * z0[0] + Tz[j * 2 + 1] * z0[1] + Tg[j]; /* :297 */
        lambda[j] = (p->g + zacc) / zp;                          /* :306 */
        if (zpos) zpos[j] = zp;
    }
}

static void mat2_mul(const double A[4], const double B[4], double C[4])
{
    double c0 = A[0] * B[0] + A[1] * B[2], c1 = A[0] * B[1] + A[1] * B[3];
    double c2 = A[2] * B[0] + A[3] * B[2], c3 = A[2] * B[1] + A[3] * B[3];
    C[0] = c0; C[1] = c1; C[2] = c2; C[3] = c3;
}

/* MPCSolver.cpp:325-389 */
void oracle_formc_horizontal_qp(const oracle_formc_params* p, const double* lambda, const double cs[2],
                                const double* mid_q, int footstep_counter,
                                double* a, double* b, double* lo, double* hi, double* gq,
                                double* phi_state_out)
{
    int N = p->N;
    double dt = p->dt;
    double eta = sqrt(p->g / p->h);                       /* parameters.cpp:41 */
    double half = (footstep_counter > 1) ? p->box_w / 2 : p->box_w_init / 2; /* :328-338 */
    for (int i = 0; i < N; ++i) { lo[i] = mid_q[i] - 1.0 * half; hi[i] = mid_q[i] + 1.0 * half; }
    double phi_state[4] = {1, 0, 0, 1};                   /* :349 */
    double* phi_in = (double*)malloc(sizeof(double) * 2 * N);
    for (int i = 0; i < N; ++i) {                          /* :351-373 */
        double Axy[4], Bxy[2];
        if (lambda[i] < 2.0) {
            Axy[0] = 1.0; Axy[1] = dt; Axy[2] = 0.0; Axy[3] = 1.0; Bxy[0] = 0.0; Bxy[1] = 0.0;
        } else {
            double s = sqrt(lambda[i]);
            double ch = cosh(s * dt), sh = sinh(s * dt);
            Axy[0] = ch; Axy[1] = sh / s; Axy[2] = s * sh; Axy[3] = ch;
            Bxy[0] = 1 - ch; Bxy[1] = -s * sh;
        }
        mat2_mul(Axy, phi_state, phi_state);              /* :362 */
        double v0 = Bxy[0], v1 = Bxy[1];                  /* :363 */
        for (int j = i + 1; j < N; ++j) {                 /* :365-371 */
            double Aj[4];
            if (lambda[j] < 2.0) { Aj[0] = 1; Aj[1] = dt; Aj[2] = 0; Aj[3] = 1; }
            else {
                double s = sqrt(lambda[j]);
                double ch = cosh(s * dt), sh = sinh(s * dt);
                Aj[0] = ch; Aj[1] = sh / s; Aj[2] = s * sh; Aj[3] = ch;
            }
            double n0 = Aj[0] * v0 + Aj[1] * v1, n1 = Aj[2] * v0 + Aj[3] * v1;
            v0 = n0; v1 = n1;
        }
        phi_in[i] = v0; phi_in[N + i] = v1;
    }
    double Csc[2] = {1.0, 1.0 / eta};                     /* :375-377 (eta_sc forced to nominal eta) */
    for (int i = 0; i < N; ++i) a[i] = Csc[0] * phi_in[i] + Csc[1] * phi_in[N + i]; /* :378 */
    double cps0 = Csc[0] * phi_state[0] + Csc[1] * phi_state[2];
    double cps1 = Csc[0] * phi_state[1] + Csc[1] * phi_state[3];
    double tail = 0.0;
    for (int i = 0; i < N; ++i) tail += exp(-dt * eta * i) * mid_q[N + i]; /* deltas :183-184, tail :381 */
    *b = -(cps0 * cs[0] + cps1 * cs[1]) + eta * dt * tail;
    for (int i = 0; i < N; ++i) gq[i] = -mid_q[i];        /* :388-389 */
    if (phi_state_out) memcpy(phi_state_out, phi_state, sizeof(phi_state));
    free(phi_in);
}

void oracle_formc_stack_horizontal(int N, const double* a, double b, const double* lo, const double* hi,
                                   double* H, double* A, double* lbA, double* ubA)
{
    memset(H, 0, sizeof(double) * N * N);
    memset(A, 0, sizeof(double) * (size_t)(N + 1) * N);
    for (int i = 0; i < N; ++i) H[i * N + i] = 1.0;       /* costFunctionH_xy = I  MPCSolver.cpp:187 */
    for (int i = 0; i < N; ++i) A[i] = a[i];
    lbA[0] = b; ubA[0] = b;
    for (int i = 0; i < N; ++i) { A[(size_t)(i + 1) * N + i] = 1.0; lbA[i + 1] = lo[i]; ubA[i + 1] = hi[i]; }
}

int oracle_formc_tick(const oracle_formc_params* p, oracle_qp_fn solver,
                      const double com_pos[3], const double com_vel[3],
                      double sim_time, int mpc_iter, int control_iter, int footstep_counter,
                      const double* plan, int n_steps, oracle_formc_out* out,
                      double* f_out, double* ux_out, double* uy_out,
                      int* ws_z, int* ws_x, int* ws_y, double* y_z, double* y_x, double* y_y)
{
    (void)control_iter; /* gate controlIter % (int)(100*dt) == 0 is always true for dt=0.01, MPCSolver.cpp:214 */
    int N = p->N, per = p->S + p->F;
    int rows = n_steps * per;
    int k0 = (int)(sim_time / (p->dt / p->dtc));          /* :259,329 */
    memset(out, 0, sizeof(*out));
    for (int c = 0; c < 3; ++c) { out->com_pos[c] = com_pos[c]; out->com_vel[c] = com_vel[c]; }
    if (k0 < 0 || k0 + 2 * N > rows) return -10;
    double* mid = (double*)malloc(sizeof(double) * (size_t)rows * 3);
    oracle_formc_midpoint(plan, n_steps, p->S, p->F, mid);
    int nCz_max = p->F + p->S + N + 2;
    size_t szA = (size_t)(nCz_max > N + 1 ? nCz_max : N + 1) * N;
    double* H = (double*)malloc(sizeof(double) * N * N);
    double* A = (double*)malloc(sizeof(double) * szA);
    double* gq = (double*)malloc(sizeof(double) * N);
    double* lbA = (double*)malloc(sizeof(double) * (nCz_max + N));
    double* ubA = (double*)malloc(sizeof(double) * (nCz_max + N));
    double* yv = (double*)malloc(sizeof(double) * (nCz_max + N));
    int* wsv = (int*)malloc(sizeof(int) * (nCz_max + N));
    double* f = (double*)malloc(sizeof(double) * N);
    double* midq = (double*)malloc(sizeof(double) * 2 * N);
    double* lambda = (double*)malloc(sizeof(double) * N);
    double* a = (double*)malloc(sizeof(double) * N);
    double* lo = (double*)malloc(sizeof(double) * N);
    double* hi = (double*)malloc(sizeof(double) * N);
    double* u[2];
    u[0] = (double*)calloc(N, sizeof(double));
    u[1] = (double*)calloc(N, sizeof(double));

    /* STAGE 1 */
    double z0[2] = {com_pos[2], com_vel[2]};
    for (int k = 0; k < N; ++k) midq[k] = mid[(size_t)(k0 + k) * 3 + 2];
    int ne = 0;
    int nCz = oracle_formc_vertical_qp(p, z0, midq, mpc_iter, footstep_counter, H, gq, A, lbA, ubA, &ne);
    out->ne_z = ne;
    out->ret[0] = solver(N, nCz, H, gq, A, lbA, ubA, f, yv, wsv, &out->nwsr[0]);
    for (int k = 0; k < N; ++k) {
        if (f_out) f_out[k] = f[k];
        if (ws_z) ws_z[k] = wsv[ne + k];
        if (y_z) y_z[k] = yv[ne + k];
    }
    out->fz0 = f[0];
    /* :274-278 */
    double nz0 = 1.0 * z0[0] + p->dt * z0[1] + 0.0 * f[0] + 0.0 * p->g;
    double nz1 = 0.0 * z0[0] + 1.0 * z0[1] + (p->dt / p->mass) * f[0] + (-p->dt) * p->g;
    out->com_pos[2] = isnan(nz0) ? p->h : nz0;
    out->com_vel[2] = isnan(nz1) ? 0.0 : nz1;
    /* STAGE 2 */
    oracle_formc_lambda(p, z0, f, lambda, NULL);
    out->lambda0 = lambda[0];
    /* STAGE 3 */
    double st[2][2] = {{com_pos[0], com_vel[0]}, {com_pos[1], com_vel[1]}};
    if (lambda[0] > 2.0) {                                 /* :322 */
        for (int ax = 0; ax < 2; ++ax) {
            double b;
            for (int k = 0; k < 2 * N; ++k) midq[k] = mid[(size_t)(k0 + k) * 3 + ax];
            oracle_formc_horizontal_qp(p, lambda, st[ax], midq, footstep_counter, a, &b, lo, hi, gq, NULL);
            oracle_formc_stack_horizontal(N, a, b, lo, hi, H, A, lbA, ubA);
            out->ret[1 + ax] = solver(N, N + 1, H, gq, A, lbA, ubA, u[ax], yv, wsv, &out->nwsr[1 + ax]);
            int* wso = ax == 0 ? ws_x : ws_y;
            double* yo = ax == 0 ? y_x : y_y;
            for (int k = 0; k < N; ++k) { if (wso) wso[k] = wsv[1 + k]; if (yo) yo[k] = yv[1 + k]; }
        }
    } else {
        for (int k = 0; k < N; ++k) { if (ws_x) ws_x[k] = 0; if (ws_y) ws_y[k] = 0; if (y_x) y_x[k] = 0; if (y_y) y_y[k] = 0; }
    }
    for (int k = 0; k < N; ++k) { if (ux_out) ux_out[k] = u[0][k]; if (uy_out) uy_out[k] = u[1][k]; }
    out->zmp_in[0] = u[0][0]; out->zmp_in[1] = u[1][0];   /* :402-403 */
    /* integrate :406-422 */
    double Axy[4], Bxy[2];
    if (lambda[0] < 2.0) { Axy[0] = 1; Axy[1] = p->dt; Axy[2] = 0; Axy[3] = 1; Bxy[0] = 0; Bxy[1] = 0; }
    else {
        double s = sqrt(lambda[0]);
        double ch = cosh(s * p->dt), sh = sinh(s * p->dt);
        Axy[0] = ch; Axy[1] = sh / s; Axy[2] = s * sh; Axy[3] = ch; Bxy[0] = 1.0 - ch; Bxy[1] = -s * sh;
    }
    for (int ax = 0; ax < 2; ++ax) {
        double c = st[ax][0], cd = st[ax][1], uin = u[ax][0];
        out->com_pos[ax] = Axy[0] * c + Axy[1] * cd + Bxy[0] * uin;
        out->com_vel[ax] = Axy[2] * c + Axy[3] * cd + Bxy[1] * uin;
    }
    free(mid); free(H); free(A); free(gq); free(lbA); free(ubA); free(yv); free(wsv); free(f); free(midq);
    free(lambda); free(a); free(lo); free(hi); free(u[0]); free(u[1]);
    return (out->ret[0] || out->ret[1] || out->ret[2]) ? 1 : 0;
}

/* ======================================================================= */
/*  Formulation A                                                          */
/* ======================================================================= */

/* quad_as_bip_bang.m:74-84 (initial) / :547-555 (rebuilt).  t is the 1-based MATLAB index. */
double oracle_forma_centerline(const double* fs_plan, int n_fs, int axis, int step, int ds,
                               int first_ramp, int t)
{
    int seg = (t - 1) / step;       /* 0-based segment: MATLAB footstep i = seg+1 */
    int r = (t - 1) % step;
    if (seg > n_fs - 2) { seg = n_fs - 2; r = step - 1; } /* beyond the built centerline: clamp */
    double a = fs_plan[seg * 2 + axis], b = fs_plan[(seg + 1) * 2 + axis];
    if (seg == 0 && !first_ramp) return a;                /* :547-548 */
    if (r < step - ds) return a;
    int k = r - (step - ds);                              /* linspace(a,b,ds)(k+1) */
    if (ds == 1) return b;
    if (k == ds - 1) return b;
    return a + (double)k * ((b - a) / (double)(ds - 1));
}

void oracle_forma_build(const oracle_forma_params* p, const double st[6], const double cur_fs[2],
                        const double fs_store[2], int j, int fs_counter,
                        const int* fs_timing, int n_timing, int ds,
                        const double* fs_plan, int n_fs, int cl_first_ramp,
                        double* Hdiag, double* gq, double* A, double* lbA, double* ubA)
{
    int C = p->C, P = p->P, F = p->F;
    int nV = 2 * (C + F), nC = nV + 2;
    double dt = p->dt, eta = p->eta;
    int step = fs_timing[1] - fs_timing[0];
    memset(A, 0, sizeof(double) * (size_t)nC * nV);
    /* mapping, bang.m:126-140.  fs_timing(k) (1-based) == fs_timing[k-1]. */
    int mcols = F + 3;
    double* map = (double*)calloc((size_t)C * mcols, sizeof(double));
    int pf = 0;
    for (int i = 1; i <= C; ++i) {
        int idx = fs_counter + pf + 1;                       /* 1-based */
        if (idx <= n_timing && j + i >= fs_timing[idx - 1]) pf = pf + 1;
        idx = fs_counter + pf + 1;
        int rem = fs_timing[(idx <= n_timing ? idx : n_timing) - 1] - (j + i);
        if (rem > ds) map[(i - 1) * mcols + pf] = 1.0;
        else {
            map[(i - 1) * mcols + pf] = (double)rem / (double)ds;
            map[(i - 1) * mcols + pf + 1] = 1.0 - (double)rem / (double)ds;
        }
    }
    /* cost, bang.m:239-245 */
    for (int ax = 0; ax < 2; ++ax) {
        int o = ax * (C + F);
        for (int k = 0; k < C; ++k) { Hdiag[o + k] = p->Qzdot; gq[o + k] = 0.0; }
        for (int f = 0; f < F; ++f) {
            int row = fs_counter + 1 + f;                    /* fs_plan(fsCounter+1 : fsCounter+F) 1-based */
            if (row > n_fs) row = n_fs;
            Hdiag[o + C + f] = p->Qfoot;
            gq[o + C + f] = -p->Qfoot * fs_plan[(row - 1) * 2 + ax];
        }
    }
    /* stability rows, bang.m:195-210 */
    double lam = exp(-eta * dt);
    double lamC = pow(lam, (double)C);
    for (int ax = 0; ax < 2; ++ax) {
        int o = ax * (C + F);
        double* row = A + (size_t)ax * nV;
        for (int i = 0; i < C; ++i)
            row[o + i] = (1.0 / eta) * (1.0 - lam) / (1.0 - lamC) * exp(-eta * dt * i) - dt * exp(-eta * dt * C);
        double ant = 0.0;
        for (int i = C + 1; i <= P; ++i)
            ant += exp(-eta * dt * i) * (1.0 - exp(-eta * dt)) *
                   (oracle_forma_centerline(fs_plan, n_fs, ax, step, ds, cl_first_ramp, j + i) - fs_store[ax]);
        /* cl(P): ABSOLUTE index P, not j+P (bang.m:196,198) -- copied as written */
        ant += exp(-eta * dt * P) * (oracle_forma_centerline(fs_plan, n_fs, ax, step, ds, cl_first_ramp, P) - fs_store[ax]);
        double rhs = st[ax * 3 + 0] + st[ax * 3 + 1] / eta - st[ax * 3 + 2] - ant;
        lbA[ax] = rhs; ubA[ax] = rhs;
    }
    /* ZMP rows, bang.m:142-150 (two-sided) */
    for (int ax = 0; ax < 2; ++ax) {
        int o = ax * (C + F);
        double w = ax == 0 ? p->wx : p->wy;
        double zq = st[ax * 3 + 2];
        for (int i = 0; i < C; ++i) {
            double* row = A + (size_t)(2 + ax * C + i) * nV;
            for (int k = 0; k <= i; ++k) row[o + k] = 1.0 * dt;          /* Pzmp = tril(ones)*dt :69 */
            for (int f = 0; f < F; ++f) row[o + C + f] = -map[i * mcols + 1 + f];
            lbA[2 + ax * C + i] = 1.0 * (-zq - w / 2) + map[i * mcols] * cur_fs[ax];
            ubA[2 + ax * C + i] = 1.0 * (-zq + w / 2) + map[i * mcols] * cur_fs[ax];
        }
    }
    /* kinematic rows, bang.m:156-190 */
    for (int ax = 0; ax < 2; ++ax) {
        int o = ax * (C + F);
        for (int f = 0; f < F; ++f) {
            int r = 2 + 2 * C + ax * F + f;
            double* row = A + (size_t)r * nV;
            row[o + C + f] = 1.0;
            if (f > 0) row[o + C + f - 1] = -1.0;
            double bnd;
            if (ax == 0) bnd = (fs_counter == 1 && f == 0) ? p->disp_forw_dummy : p->disp_forw;
            else bnd = p->disp_L / 2 + p->disp_L / 2;
            double c0 = (f == 0) ? cur_fs[ax] : 0.0;
            lbA[r] = -bnd + c0; ubA[r] = bnd + c0;
        }
    }
    free(map);
}

/* bang.m:55-58, 265-290 */
void oracle_forma_integrate(const oracle_forma_params* p, double st[6], double zdx0, double zdy0)
{
    double eta = p->eta, dt = p->dt;
    double ch = cosh(eta * dt), sh = sinh(eta * dt);
    double Au[9] = {ch, sh / eta, 1 - ch, eta * sh, ch, -eta * sh, 0, 0, 1};
    double Bu[3] = {dt - sh / eta, 1 - ch, dt};
    double in[2] = {zdx0, zdy0};
    for (int ax = 0; ax < 2; ++ax) {
        double* s = st + ax * 3;
        double n0 = Au[0] * s[0] + Au[1] * s[1] + Au[2] * s[2] + Bu[0] * in[ax];
        double n1 = Au[3] * s[0] + Au[4] * s[1] + Au[5] * s[2] + Bu[1] * in[ax];
        double n2 = Au[6] * s[0] + Au[7] * s[1] + Au[8] * s[2] + Bu[2] * in[ax];
        s[0] = n0; s[1] = n1; s[2] = n2;
    }
}

int oracle_forma_tick(const oracle_forma_params* p, oracle_qp_fn solver,
                      const double st[6], const double cur_fs[2], const double fs_store[2],
                      int j, int fs_counter, const int* fs_timing, int n_timing, int ds,
                      const double* fs_plan, int n_fs, int cl_first_ramp,
                      oracle_forma_out* out, double* v_out, int* ws_out, double* yd_out)
{
    int C = p->C, F = p->F, nV = 2 * (C + F), nC = nV + 2;
    double* Hd = (double*)malloc(sizeof(double) * nV);
    double* H = (double*)calloc((size_t)nV * nV, sizeof(double));
    double* gq = (double*)malloc(sizeof(double) * nV);
    double* A = (double*)malloc(sizeof(double) * (size_t)nC * nV);
    double* lbA = (double*)malloc(sizeof(double) * nC);
    double* ubA = (double*)malloc(sizeof(double) * nC);
    double* v = (double*)malloc(sizeof(double) * nV);
    double* yv = (double*)malloc(sizeof(double) * nC);
    int* wsv = (int*)malloc(sizeof(int) * nC);
    oracle_forma_build(p, st, cur_fs, fs_store, j, fs_counter, fs_timing, n_timing, ds, fs_plan, n_fs,
                       cl_first_ramp, Hd, gq, A, lbA, ubA);
    for (int i = 0; i < nV; ++i) H[(size_t)i * nV + i] = Hd[i];
    out->ret = solver(nV, nC, H, gq, A, lbA, ubA, v, yv, wsv, &out->nwsr);
    memcpy(out->st, st, sizeof(double) * 6);
    oracle_forma_integrate(p, out->st, v[0], v[C + F]);
    out->pred_fs[0] = v[C]; out->pred_fs[1] = v[C + F + C];
    if (v_out) memcpy(v_out, v, sizeof(double) * nV);
    for (int i = 0; i < nC - 2; ++i) { if (ws_out) ws_out[i] = wsv[2 + i]; if (yd_out) yd_out[i] = yv[2 + i]; }
    free(Hd); free(H); free(gq); free(A); free(lbA); free(ubA); free(v); free(yv); free(wsv);
    return out->ret;
}

/* bang.m:99-563 (QP-1 loop only) */
int oracle_forma_closed_loop2(const oracle_forma_params* p, oracle_qp_fn solver,
                              double st[6], double* fs_plan, int n_fs,
                              const int* fs_timing, int n_timing, int ds, int n_ticks,
                              int push_fs, int push_ct0, int push_ct1, double push_ax, double push_ay,
                              double* traj, int* nwsr_total, double* pred_traj, int* fsc_traj)
{
    int fails = 0, fs_counter = 1, ct = 0, first_ramp = 1, wsr = 0;
    double cur[2] = {fs_plan[0], fs_plan[1]};             /* bang.m:50-51 */
    for (int j = 1; j <= n_ticks; ++j) {
        if (fs_counter == push_fs && ct >= push_ct0 && ct < push_ct1) { /* :104-114 */
            st[1] += p->dt * push_ax; st[4] += p->dt * push_ay;
        }
        oracle_forma_out o;
        int rc = oracle_forma_tick(p, solver, st, cur, cur, j, fs_counter, fs_timing, n_timing, ds,
                                   fs_plan, n_fs, first_ramp, &o, NULL, NULL, NULL);
        if (rc) ++fails;
        wsr += o.nwsr;
        memcpy(st, o.st, sizeof(double) * 6);
        if (traj) {
            double* t = traj + (size_t)(j - 1) * 6;
            t[0] = st[0]; t[1] = st[3]; t[2] = st[1]; t[3] = st[4]; t[4] = st[2]; t[5] = st[5];
        }
        if (pred_traj) { pred_traj[(size_t)(j - 1) * 2] = o.pred_fs[0]; pred_traj[(size_t)(j - 1) * 2 + 1] = o.pred_fs[1]; }
        if (fsc_traj) fsc_traj[j - 1] = fs_counter;
        ct = ct + 1;
        if (fs_counter + 1 <= n_timing && j + 1 >= fs_timing[fs_counter]) { /* :529  fs_timing(fsCounter+1) */
            fs_counter += 1;
            cur[0] = o.pred_fs[0]; cur[1] = o.pred_fs[1];
            if (fs_counter >= 2 && fs_counter <= n_fs) {  /* :539-556 */
                double dx = o.pred_fs[0] - fs_plan[(fs_counter - 1) * 2 + 0];
                double dy = o.pred_fs[1] - fs_plan[(fs_counter - 1) * 2 + 1];
                for (int i = 0; i < n_fs; ++i) { fs_plan[i * 2] += dx; fs_plan[i * 2 + 1] += dy; }
                first_ramp = 0;
            }
            ct = 0;
        }
    }
    if (nwsr_total) *nwsr_total = wsr;
    return fails;
}

int oracle_forma_closed_loop(const oracle_forma_params* p, oracle_qp_fn solver,
                             double st[6], double* fs_plan, int n_fs,
                             const int* fs_timing, int n_timing, int ds, int n_ticks,
                             int push_fs, int push_ct0, int push_ct1, double push_ax, double push_ay,
                             double* traj, int* nwsr_total)
{
    return oracle_forma_closed_loop2(p, solver, st, fs_plan, n_fs, fs_timing, n_timing, ds, n_ticks, push_fs,
                                     push_ct0, push_ct1, push_ax, push_ay, traj, nwsr_total, NULL, NULL);
}

/* ======================================================================================================
 * Real-foot placement stage ("SECOND QUAD_PROG") and trajectory export
 * ====================================================================================================== */
#define FP(r, c) foot_plan[(size_t)((r) - 1) * 8 + ((c) - 1)]   /* MATLAB 1-based foot_plan(r,c) */

/* compute_two_feet1.m:5-16 == compute_one_feet_walk.m:99-113: line through the two fixed feet (polyfit of degree 1
 * through two points), line of opposite slope through the ZMP, their intersection, and the offset of the ZMP from
 * it.  The scripts use the symbolic toolbox (`solve`); this is the same 2x2 intersection in closed form.
 * A horizontal diagonal (slope 0) has no intersection in the scripts (solve returns empty and they error out):
 * reported as ok = 0 and treated as "nothing changes". */
static int diag_offset(const double fixed[4], const double zmp[2], double* slope, double* dist_x, double* dist_y)
{
    const double m = (fixed[3] - fixed[1]) / (fixed[2] - fixed[0]);
    const double q = fixed[1] - m * fixed[0];
    *slope = m;
    if (!(fabs(m) > 0.0) || !isfinite(m)) { *dist_x = 0.0; *dist_y = 0.0; return 0; }
    const double xs = (zmp[1] + m * zmp[0] - q) / (2.0 * m);   /* m x + q = zmp_y - m (x - zmp_x) */
    const double ys = m * xs + q;
    *dist_x = zmp[0] - xs; *dist_y = zmp[1] - ys;
    return 1;
}

static double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

int oracle_feet_trot_tick(const oracle_feet_params* p, int fs_counter, const double pred[2], double phi,
                          double* foot_plan, int rows)
{
    const int fs = fs_counter;
    if (fs < 1 || fs + 1 > rows) return 0;
    const int odd = (fs % 2) == 1;
    /* odd: rr and fl stay, rl and fr move (quad_as_bip_no_plots.m:339,364); even: rl and fr stay, rr and fl move (:389) */
    const int fc1 = odd ? 3 : 1, fc2 = odd ? 7 : 5;     /* fixed feet columns (x) in row fs */
    const int mc1 = odd ? 1 : 3, mc2 = odd ? 5 : 7;     /* free feet columns (x) in row fs+1 */
    const double fixed[4] = {FP(fs, fc1), FP(fs, fc1 + 1), FP(fs, fc2), FP(fs, fc2 + 1)};
    const double free_[4] = {FP(fs + 1, mc1), FP(fs + 1, mc1 + 1), FP(fs + 1, mc2), FP(fs + 1, mc2 + 1)};
    double m, dx, dy;
    int changed = 0;
    if (diag_offset(fixed, pred, &m, &dx, &dy)) {
        /* compute_two_feet1.m:18-38: the free feet slide along the heading phi onto the line of slope -m through the ZMP */
        double x1, y1, x2, y2;
        if (phi == 3.14159265358979323846 / 2) {        /* MATLAB: phi==pi/2 */
            x1 = free_[0]; x2 = free_[2];
            y1 = pred[1] - m * (x1 - pred[0]); y2 = pred[1] - m * (x2 - pred[0]);
        } else {
            const double t = tan(phi);
            x1 = (pred[1] + m * pred[0] + t * free_[0] - free_[1]) / (t + m);
            y1 = t * (x1 - free_[0]) + free_[1];
            x2 = (pred[1] + m * pred[0] + t * free_[2] - free_[3]) / (t + m);
            y2 = t * (x2 - free_[2]) + free_[3];
        }
        changed = (dy != 0.0 || dx != 0.0);
        if (changed) {                                   /* foot_plan(fsCounter+1,:) = quattro_piedi */
            FP(fs + 1, mc1) = x1; FP(fs + 1, mc1 + 1) = y1; FP(fs + 1, mc2) = x2; FP(fs + 1, mc2 + 1) = y2;
            FP(fs + 1, fc1) = fixed[0]; FP(fs + 1, fc1 + 1) = fixed[1]; FP(fs + 1, fc2) = fixed[2]; FP(fs + 1, fc2 + 1) = fixed[3];
        }
    }
    /* 4-variable QP, H = I, six one-sided rows each on a single variable (:344-410) -> clip per variable.
     * Variable order: odd  X = (rl.x, rl.y, fr.x, fr.y); even X = (fl.x, fl.y, rr.x, rr.y). */
    const int dummy = (odd && fs == 1);
    const double d_o = dummy ? p->disp_o_dummy : p->disp_o, d_i = dummy ? p->disp_i_dummy : p->disp_i;
    const double d_f = dummy ? p->disp_forw_dummy : p->disp_forw;
    const int a = odd ? 1 : 7, b = odd ? 5 : 3;         /* first / second foot of X (x column) */
    const double X1 = fmin(FP(fs + 1, a), FP(fs, a) + d_f);
    const double X2 = clipd(FP(fs + 1, a + 1), FP(fs, a + 1) - d_i, FP(fs, a + 1) + d_o);
    const double X3 = fmin(FP(fs + 1, b), FP(fs, b) + d_f);
    const double X4 = clipd(FP(fs + 1, b + 1), FP(fs, b + 1) - d_o, FP(fs, b + 1) + d_i);
    FP(fs + 1, a) = X1; FP(fs + 1, a + 1) = X2; FP(fs + 1, b) = X3; FP(fs + 1, b + 1) = X4;
    return changed;
}

int oracle_feet_walk_tick(const oracle_feet_params* p, int counter, int fs_counter, const double pred[2],
                          double* foot_plan, int rows)
{
    const int fs = fs_counter;
    if (counter != 2 && counter != 4 && counter != 6 && counter != 8) return 0;
    if (fs < 1 || fs + 8 > rows) return 0;               /* MATLAB would grow the array; callers size it */
    /* diagonal of the two feet that stay: rl-fr for phases 2 and 4, rr-fl for 6 and 8 (quad_walk_no_plots.m:342,379,421,438) */
    const int d1 = (counter <= 4) ? 1 : 3, d2 = (counter <= 4) ? 5 : 7;
    const int mc = counter == 2 ? 7 : counter == 4 ? 3 : counter == 6 ? 5 : 1;   /* the foot that moves */
    const double fixed[4] = {FP(fs, d1), FP(fs, d1 + 1), FP(fs, d2), FP(fs, d2 + 1)};
    double m, dx, dy;
    int changed = 0;
    if (diag_offset(fixed, pred, &m, &dx, &dy)) {
        const double xf = FP(fs + 1, mc) + dx, yf = FP(fs + 1, mc + 1) + dy;   /* compute_one_feet_walk.m:115-116 */
        changed = (dy != 0.0 || dx != 0.0);
        if (changed) for (int l = 1; l <= 8; ++l) { FP(fs + l, mc) = xf; FP(fs + l, mc + 1) = yf; }
    }
    const int dummy = (counter <= 4) && fs <= 4;
    const double d_o = dummy ? p->disp_o_dummy : p->disp_o, d_i = dummy ? p->disp_i_dummy : p->disp_i;
    const double d_f = dummy ? p->disp_forw_dummy : p->disp_forw;
    /* left feet (phases 2: fl, 8: rl) may move outwards by disp_o / inwards by disp_i; right feet (4: rr, 6: fr) mirrored */
    const int left = (counter == 2 || counter == 8);
    const double up = left ? d_o : d_i, dn = left ? d_i : d_o;
    const double X1 = fmin(FP(fs + 1, mc), FP(fs, mc) + d_f);
    const double X2 = clipd(FP(fs + 1, mc + 1), FP(fs, mc + 1) - dn, FP(fs, mc + 1) + up);
    for (int l = 1; l <= 8; ++l) {
        FP(fs + l, mc) = X1;
        if (counter != 8 || l == 1) FP(fs + l, mc + 1) = X2;   /* :500-503 writes foot_plan(fsCounter+1,2) only: copied */
    }
    return changed;
}

static void put3(double* dst, size_t k, double x, double y, double z) { dst[3 * k] = x; dst[3 * k + 1] = y; dst[3 * k + 2] = z; }

void oracle_feet_export_trot(const double* foot_plan_c, int rows, int n_steps, int fixed, int swing,
                             double* fl, double* fr, double* rl, double* rr)
{
    const double* foot_plan = foot_plan_c;
    size_t k = 0;
    for (int i = 1; i <= n_steps && i + 1 <= rows; ++i) {
        for (int s = 1; s <= fixed; ++s, ++k) {
            put3(fl, k, FP(i, 7), FP(i, 8), 0.0); put3(rr, k, FP(i, 3), FP(i, 4), 0.0);
            put3(fr, k, FP(i, 5), FP(i, 6), 0.0); put3(rl, k, FP(i, 1), FP(i, 2), 0.0);
        }
        for (int j = 1; j <= swing; ++j, ++k) {
            const double z = -0.000032 * j * j + 0.0016 * j;
            if (i % 2 == 1) {
                put3(fl, k, FP(i, 7), FP(i, 8), 0.0); put3(rr, k, FP(i, 3), FP(i, 4), 0.0);
                put3(rl, k, FP(i, 1) + (FP(i + 1, 1) - FP(i, 1)) / swing * j, FP(i, 2) + (FP(i + 1, 2) - FP(i, 2)) / swing * j, z);
                put3(fr, k, FP(i, 5) + (FP(i + 1, 5) - FP(i, 5)) / swing * j, FP(i, 6) + (FP(i + 1, 6) - FP(i, 6)) / swing * j, z);
            } else {
                put3(rl, k, FP(i, 1), FP(i, 2), 0.0); put3(fr, k, FP(i, 5), FP(i, 6), 0.0);
                put3(fl, k, FP(i, 7) + (FP(i + 1, 7) - FP(i, 7)) / swing * j, FP(i, 8) + (FP(i + 1, 8) - FP(i, 8)) / swing * j, z);
                put3(rr, k, FP(i, 3) + (FP(i + 1, 3) - FP(i, 3)) / swing * j, FP(i, 4) + (FP(i + 1, 4) - FP(i, 4)) / swing * j, z);
            }
        }
    }
}

void oracle_feet_export_walk(const double* foot_plan_c, int rows, int n_steps, int step_duration,
                             double* fl, double* fr, double* rl, double* rr)
{
    const double* foot_plan = foot_plan_c;
    size_t k = 0;
    int conteggio = 1;
    for (int i = 1; i <= n_steps && i + 1 <= rows; ++i) {
        for (int s = 1; s <= step_duration; ++s, ++k) {
            const double z = -0.000032 * s * s + 0.0016 * s;
            const int mv = (conteggio == 2) ? 7 : (conteggio == 4) ? 3 : (conteggio == 6) ? 5 : (conteggio == 8) ? 1 : 0;
            double* dst[4] = {rl, rr, fr, fl};
            for (int f = 0; f < 4; ++f) {
                const int c = 2 * f + 1;
                if (c == mv)
                    put3(dst[f], k, FP(i, c) + (FP(i + 1, c) - FP(i, c)) / step_duration * s,
                         FP(i, c + 1) + (FP(i + 1, c + 1) - FP(i, c + 1)) / step_duration * s, z);
                else
                    put3(dst[f], k, FP(i, c), FP(i, c + 1), 0.0);
            }
        }
        conteggio = (conteggio == 8) ? 1 : conteggio + 1;
    }
}
#undef FP

/* ======================================================================================================
 * LIP Kalman filter -- AMR_code_DART/StateFiltering.cpp restated with generic small dense helpers (float).
 * ====================================================================================================== */
static void kf_mm(const float* a, int ar, int ac, const float* b, int bc, float* out)      /* out = a (ar x ac) * b (ac x bc) */
{
    for (int i = 0; i < ar; ++i)
        for (int j = 0; j < bc; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < ac; ++k) acc += a[i * ac + k] * b[k * bc + j];
            out[i * bc + j] = acc;
        }
}
static void kf_tr(const float* a, int ar, int ac, float* out)                               /* out = a' */
{
    for (int i = 0; i < ar; ++i) for (int j = 0; j < ac; ++j) out[j * ar + i] = a[i * ac + j];
}
static void kf_inv3x3(const float* m, float* r)
{
    const float a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const float A = e * i - f * h, B = f * g - d * i, C = d * h - e * g;
    const float det = a * A + b * B + c * C;
    r[0] = A / det; r[1] = (c * h - b * i) / det; r[2] = (b * f - c * e) / det;
    r[3] = B / det; r[4] = (a * i - c * g) / det; r[5] = (c * d - a * f) / det;
    r[6] = C / det; r[7] = (b * g - a * h) / det; r[8] = (a * e - b * d) / det;
}
/* predict_z / predict_xy (StateFiltering.cpp:97-103,115-124) */
static void kf_pred(const float* A, const float* B, const float* q, float u, float* st, float* sg)
{
    float t5[5], in2[2] = {u, 0.0f}, bu[5], At[25], AS[25], ASA[25], Bt[10], BQ[10], BQB[25];
    kf_mm(A, 5, 5, st, 1, t5); kf_mm(B, 5, 2, in2, 1, bu);
    for (int i = 0; i < 5; ++i) st[i] = t5[i] + bu[i];
    kf_tr(A, 5, 5, At); kf_mm(A, 5, 5, sg, 5, AS); kf_mm(AS, 5, 5, At, 5, ASA);
    kf_tr(B, 5, 2, Bt); kf_mm(B, 5, 2, q, 2, BQ); kf_mm(BQ, 5, 2, Bt, 5, BQB);
    for (int i = 0; i < 25; ++i) sg[i] = ASA[i] + BQB[i];
}
/* update_z / update_xy (StateFiltering.cpp:104-112,125-133) */
static void kf_upd(const float* Cm, const float* R, const float* z, const float* off, float* st, float* sg)
{
    float Ct[15], SC[15], CSC[9], S[9], Si[9], K[15], Cs[3], inn[3], Ki[5], KC[25], KCS[25];
    kf_tr(Cm, 3, 5, Ct); kf_mm(sg, 5, 5, Ct, 3, SC); kf_mm(Cm, 3, 5, SC, 3, CSC);
    for (int i = 0; i < 9; ++i) S[i] = R[i] + CSC[i];
    kf_inv3x3(S, Si); kf_mm(SC, 5, 3, Si, 3, K);
    kf_mm(Cm, 3, 5, st, 1, Cs);
    for (int i = 0; i < 3; ++i) inn[i] = z[i] - (Cs[i] + off[i]);
    kf_mm(K, 5, 3, inn, 1, Ki);
    for (int i = 0; i < 5; ++i) st[i] += Ki[i];
    kf_mm(K, 5, 3, Cm, 5, KC); kf_mm(KC, 5, 5, sg, 5, KCS);
    for (int i = 0; i < 25; ++i) sg[i] -= KCS[i];
}
void oracle_kf_filter(const oracle_kf_model* m, oracle_kf_state* s, const oracle_kf_sample* samples, int n_steps, float* zmp)
{
    const float T = m->sampling_time;
    const float A[25] = {1.0f, T, T * T / 2, 0, 0,  0, 1.0f, T, T, 0,  0, 0, 1.0f, 0, 0,  0, 0, 0, 1.0f, T,  0, 0, 0, 0, 1.0f};   /* :36-40 */
    const float B[10] = {T * T * T / 6, 0,  T * T / 2, 0,  T, 0,  0, T * T / 2,  0, T};                                          /* :42-46 */
    const float Cz[15] = {1.0f, 0, 0, 0, 0,  0, 0, 1.0f, 0, 0,  0, 0, -m->mass, 1.0f, 0};                                         /* :48-50 */
    float Cxy[15] = {1.0f, 0, 0, 0, 0,  0, 0, 1.0f, 0, 0,  1.0f, 0, 0, 0, 0};                                                     /* :52-54 */
    const float offz[3] = {0.0f, 0.0f, -m->g * m->mass}, off0[3] = {0.0f, 0.0f, 0.0f};
    for (int t = 0; t < n_steps; ++t) {
        const oracle_kf_sample* u = samples + t;
        kf_pred(A, B, m->q_process[2], u->input[2], s->state[2], s->sigma[2]);             /* predict_z */
        kf_upd(Cz, m->q_measurement[2], u->meas[2], offz, s->state[2], s->sigma[2]);       /* update_z  */
        kf_pred(A, B, m->q_process[0], u->input[0], s->state[0], s->sigma[0]);             /* predict_xy */
        kf_pred(A, B, m->q_process[1], u->input[1], s->state[1], s->sigma[1]);
        const float f_n = -m->mass * m->g - m->mass * s->state[2][2] + s->state[2][3];     /* update_xy :127-129 */
        Cxy[2 * 5 + 2] = m->mass * s->state[2][0] / f_n;
        Cxy[2 * 5 + 3] = -s->state[2][0] / f_n;
        kf_upd(Cxy, m->q_measurement[0], u->meas[0], off0, s->state[0], s->sigma[0]);
        kf_upd(Cxy, m->q_measurement[1], u->meas[1], off0, s->state[1], s->sigma[1]);
        if (zmp) {
            for (int ax = 0; ax < 2; ++ax) {
                float acc = 0.0f;
                for (int l = 0; l < 5; ++l) acc += Cxy[2 * 5 + l] * s->state[ax][l];
                zmp[t * 2 + ax] = acc;
            }
        }
    }
}
