/*
 * ismpc_oracle.h -- CPU restatement ("oracle") of the reference's ISMPC hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the shipped product path (the CUDA
 * library under quadruped_gait_generation_ismpc_b200/) may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker /
 * the CPU arm being timed.
 *
 * Parity pinning: the QP *solver* half of the oracle is the reference's own
 * vendored qpOASES 3.2 compiled unmodified from /root/reference (see
 * oracle/Makefile -> oracle/_ref/libismpc_oracle_ref.so) and driven with the
 * exact call form of AMR_code_DART/utils.cpp:89-139.  The QP *builder* half
 * is the plain-C restatement below (MPCSolver.cpp cannot be compiled here:
 * Eigen/HPIPM/BLASFEO/DART are absent), pinned against (i) the MATLAB
 * closed-loop CoM fixtures shipped in the reference
 * (AMR_code_DART/MATLAB_trajectories) for formulation A and (ii) the literal
 * O(N^2) loops of MPCSolver.cpp for formulation C.  Parity against the
 * HPIPM calls that MPCSolver::solve makes as shipped is UNPINNED (HPIPM and
 * BLASFEO are un-vendored, un-versioned dependencies: CMakeLists.txt:10-14).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/).
 */
#ifndef ISMPC_ORACLE_H
#define ISMPC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- generic dense QP solver callback ---------------------------------
 * min 1/2 x'Hx + g'x  s.t.  lbA <= A x <= ubA      (H nV x nV, A nC x nV, row-major)
 * Same argument meaning as solveQP(H,f,A,lbA,ubA), AMR_code_DART/utils.cpp:89-90.
 * x[nV] primal, y[nC] constraint duals (qpOASES sign: >0 lower active, <0 upper),
 * ws[nC] working set (-1 lower / 0 inactive / +1 upper, as
 * QProblem::getWorkingSetConstraints), *nwsr = iterations used.
 * Returns 0 on success, non-zero solver code otherwise. */
typedef int (*oracle_qp_fn)(int nV, int nC, const double* H, const double* g,
                            const double* A, const double* lbA, const double* ubA,
                            double* x, double* y, int* ws, int* nwsr);

/* Portable textbook dual active-set (Goldfarb-Idnani, Schur-complement form,
 * dense, re-factorised every iteration).  Used where the compiled reference
 * qpOASES (oracle/_ref) is not available; pinned against it in tests. */
int oracle_qp_dual_active_set(int nV, int nC, const double* H, const double* g,
                              const double* A, const double* lbA, const double* ubA,
                              double* x, double* y, int* ws, int* nwsr);

/* ---- formulation C : what MPCSolver::solve builds (MPCSolver.cpp:204-501) ---- */
typedef struct {
    double dt;          /* mpcTimeStep            parameters.cpp:9  */
    double dtc;         /* controlTimeStep        parameters.cpp:10 */
    double h;           /* comTargetHeight        parameters.cpp:17 */
    double mass;        /* mass_hrp4              parameters.cpp:39 */
    double g;           /* g                      parameters.cpp:40 */
    double box_w;       /* footConstraintSquareWidth parameters.cpp:22 */
    double box_w_init;  /* half-width "1" used while footstepCounter<=1, MPCSolver.cpp:334-337 (full width = 2) */
    double q_p, q_v, q_u; /* MPCSolver.cpp:253-255 */
    double fz_max;      /* 10000.0, MPCSolver.cpp:159 */
    int N, S, F;        /* parameters.cpp:42-44 */
} oracle_formc_params;

void oracle_formc_default_params(oracle_formc_params* p);

/* MPCSolver.cpp:167-180.  mid: n_steps*(S+F) rows x 3 cols row-major (last step's rows stay 0). */
void oracle_formc_midpoint(const double* plan_xyzt, int n_steps, int S, int F, double* mid);

/* MPCSolver.cpp:124-156 via literal matrixPower (utils.cpp:73-81).
 * Sz,Szv,Sgz,Sgzv: N x N row-major; Tz,Tzv: N x 2; Tg,Tgv: N. */
void oracle_formc_vertical_matrices(const oracle_formc_params* p, double* Sz, double* Szv,
                                    double* Tz, double* Tzv, double* Tg, double* Tgv);

/* Stage 1, MPCSolver.cpp:220-269, stacked for solveQP: A=[Aeq (ne rows, only when running); S_bar_z],
 * lbA=[0;0], ubA=[0;fz_max].  Returns nC = ne + N; *ne_out = ne.
 * H: N*N, gq: N, A: (F+N)*N max, lbA/ubA: F+N max.  mid_z: N values mid[k0..k0+N). */
int oracle_formc_vertical_qp(const oracle_formc_params* p, const double z0[2], const double* mid_z,
                             int mpc_iter, int footstep_counter,
                             double* H, double* gq, double* A, double* lbA, double* ubA, int* ne_out);

/* Stage 2, MPCSolver.cpp:296-309.  f: N forces -> lambda[N]. zpos optional (N) */
void oracle_formc_lambda(const oracle_formc_params* p, const double z0[2], const double* f,
                         double* lambda, double* zpos);

/* Stage 3, MPCSolver.cpp:325-389 (literal O(N^2) loops), one axis.
 * mid_q: 2N values mid_q[k0..k0+2N).  cs[2] = (c, cdot).
 * Outputs: a[N] stability row, *b rhs, lo/hi[N], gq[N] cost vector, phi_state[4] (row-major 2x2). */
void oracle_formc_horizontal_qp(const oracle_formc_params* p, const double* lambda, const double cs[2],
                                const double* mid_q, int footstep_counter,
                                double* a, double* b, double* lo, double* hi, double* gq,
                                double* phi_state);

/* Stack one horizontal axis for solveQP: H=I, A=[a';I], lbA=[b;lo], ubA=[b;hi]. */
void oracle_formc_stack_horizontal(int N, const double* a, double b, const double* lo, const double* hi,
                                   double* H, double* A, double* lbA, double* ubA);

typedef struct {
    double com_pos[3], com_vel[3];
    double zmp_in[2];   /* decisionVariables_x(0), _y(0): MPCSolver.cpp:402-403 */
    double fz0;         /* decisionVariables_z(0) */
    double lambda0;
    int ret[3];         /* solver return codes z,x,y (the reference drops them: utils.cpp:128) */
    int nwsr[3];
    int ne_z;           /* number of (non-zero) equality rows of the vertical QP */
} oracle_formc_out;

/* Full tick of MPCSolver::solve (MPCSolver.cpp:204-430) for one instance.
 * plan_xyzt: n_steps x 4 row-major.  Optional outputs (may be NULL):
 * f[N], ux[N], uy[N] primal; ws_z[N] (inequality rows only), ws_x[N], ws_y[N] (box rows only);
 * y_z[N], y_x[N], y_y[N] duals of those rows. */
int oracle_formc_tick(const oracle_formc_params* p, oracle_qp_fn solver,
                      const double com_pos[3], const double com_vel[3],
                      double sim_time, int mpc_iter, int control_iter, int footstep_counter,
                      const double* plan_xyzt, int n_steps,
                      oracle_formc_out* out,
                      double* f, double* ux, double* uy,
                      int* ws_z, int* ws_x, int* ws_y,
                      double* y_z, double* y_x, double* y_y);

/* ---- formulation A : canonical ISMPC with footsteps (MATLAB scripts) ---- */
typedef struct {
    double dt;         /* mpcTimeStep        quad_as_bip_bang.m:28 */
    double eta;        /* sqrt(9.8/height)   quad_as_bip_bang.m:31 */
    double wx, wy;     /* ZMP box            quad_as_bip_bang.m:36-37 */
    double disp_forw;  /* init_quadruped.m:35 */
    double disp_forw_dummy; /* init_quadruped.m:36 */
    double disp_L;     /* quad_as_bip_bang.m:11 */
    double Qzdot, Qfoot; /* quad_as_bip_bang.m:239-240 */
    int C, P, F;       /* quad_as_bip_bang.m:25-27 */
} oracle_forma_params;

/* Centerline sample cl(t), t 1-based absolute tick (quad_as_bip_bang.m:74-84 initial,
 * :547-555 rebuilt).  first_ramp=1: initial centerline (segment 1 has the ds ramp);
 * first_ramp=0: rebuilt centerline (segment 1 constant).  fs_plan: n_fs x 2 row-major. */
double oracle_forma_centerline(const double* fs_plan, int n_fs, int axis, int step, int ds,
                               int first_ramp, int t);

/* One tick QP build (quad_as_bip_bang.m:121-257 == quad_as_bip_no_plots.m:138-274 ==
 * walking/quad_walk_no_plots.m:151-288) in two-sided stacked form:
 * vars v=[zdx(C); xf(F); zdy(C); yf(F)], rows=[stab_x; stab_y; ZMPx(C); ZMPy(C); kinx(F); kiny(F)].
 * Hdiag[nV], gq[nV], A[nC*nV] row-major, lbA/ubA[nC].  nV=2(C+F), nC=nV+2.
 * st = (x,xd,xz,y,yd,yz); cur_fs[2]; fs_store[2] = (xfs_store(fsCounter), yfs_store(fsCounter));
 * j (1-based tick), fs_counter (1-based), fs_timing[n_timing] (0-based C array of the MATLAB vector),
 * fs_plan n_fs x 2, cl_first_ramp as above, step = fs_timing(2)-fs_timing(1). */
void oracle_forma_build(const oracle_forma_params* p, const double st[6], const double cur_fs[2],
                        const double fs_store[2], int j, int fs_counter,
                        const int* fs_timing, int n_timing, int ds,
                        const double* fs_plan, int n_fs, int cl_first_ramp,
                        double* Hdiag, double* gq, double* A, double* lbA, double* ubA);

/* LIP 3-state update, quad_as_bip_bang.m:55-58,265-290. st (x,xd,xz,y,yd,yz) in/out. */
void oracle_forma_integrate(const oracle_forma_params* p, double st[6], double zdx0, double zdy0);

typedef struct {
    double st[6];       /* next (x,xd,xz,y,yd,yz) */
    double pred_fs[2];  /* predicted_xfs(1), predicted_yfs(1) */
    int ret, nwsr;
} oracle_forma_out;

/* Build + solve + integrate one tick. Optional: v[nV] primal, ws[nC-2] (rows after the 2 stability rows), yd[nC-2]. */
int oracle_forma_tick(const oracle_forma_params* p, oracle_qp_fn solver,
                      const double st[6], const double cur_fs[2], const double fs_store[2],
                      int j, int fs_counter, const int* fs_timing, int n_timing, int ds,
                      const double* fs_plan, int n_fs, int cl_first_ramp,
                      oracle_forma_out* out, double* v, int* ws, double* yd);

/* Closed loop of the MATLAB scripts (QP-1 only; the 2nd QP never feeds back into the CoM loop):
 * quad_as_bip_bang.m:99-563.  fs_plan (n_fs x 2) is modified in place (plan shift at step switches).
 * push_fs/push_ct0/push_ct1/push_ax/push_ay: impulsive disturbance
 * "if fsCounter==push_fs && ct>=push_ct0 && ct<push_ct1: xd+=dt*push_ax; yd+=dt*push_ay" (:104-114).
 * traj: n_ticks x 6 (x,y,xd,yd,xz,yz) AFTER each tick's update. Returns number of failed solves. */
int oracle_forma_closed_loop(const oracle_forma_params* p, oracle_qp_fn solver,
                             double st[6], double* fs_plan, int n_fs,
                             const int* fs_timing, int n_timing, int ds, int n_ticks,
                             int push_fs, int push_ct0, int push_ct1, double push_ax, double push_ay,
                             double* traj, int* nwsr_total);

/* Same loop; additionally records, per tick, the predicted footstep handed to the second QP
 * (pred_traj: n_ticks x 2, nullable) and the fsCounter the tick ran with (fsc_traj: n_ticks, nullable). */
int oracle_forma_closed_loop2(const oracle_forma_params* p, oracle_qp_fn solver,
                              double st[6], double* fs_plan, int n_fs,
                              const int* fs_timing, int n_timing, int ds, int n_ticks,
                              int push_fs, int push_ct0, int push_ct1, double push_ax, double push_ay,
                              double* traj, int* nwsr_total, double* pred_traj, int* fsc_traj);

/* ---- real-foot placement stage (the scripts' "SECOND QUAD_PROG") and trajectory export ---- */
typedef struct {
    double disp_forw, disp_i, disp_o;                   /* init_quadruped.m:31-35 */
    double disp_forw_dummy, disp_i_dummy, disp_o_dummy; /* init_quadruped.m:33-36 */
} oracle_feet_params;

/* foot_plan: rows x 8 row-major, columns (1-based in MATLAB) 1,2 = rear-left x,y; 3,4 = rear-right;
 * 5,6 = front-right; 7,8 = front-left (init_quadruped.m:151-152).
 * Trot, one tick: trotting/quad_as_bip_no_plots.m:332-426 + trotting/compute_two_feet1.m.
 * pred = (predicted_xfs(1), predicted_yfs(1)); fs_counter 1-based.  Returns `changed`. */
int oracle_feet_trot_tick(const oracle_feet_params* p, int fs_counter, const double pred[2], double phi,
                          double* foot_plan, int rows);
/* Walk, one tick: walking/quad_walk_no_plots.m:334-504 + walking/compute_one_feet_walk.m.
 * counter: the 8-phase counter (acts for 2, 4, 6, 8 only).  Returns `changed` (0 when the phase does nothing). */
int oracle_feet_walk_tick(const oracle_feet_params* p, int counter, int fs_counter, const double pred[2],
                          double* foot_plan, int rows);
/* Foot trajectories as written to foot_{fl,fr,rl,rr}_*.txt: n_steps*(fixed+swing) samples x 3 each.
 * Trot (quad_as_bip_no_plots.m:482-509): per step `fixed` samples on the ground, then `swing` samples with the
 * diagonal pair rl/fr (odd step) or fl/rr (even step) interpolated and z = -3.2e-5 k^2 + 1.6e-3 k. */
void oracle_feet_export_trot(const double* foot_plan, int rows, int n_steps, int fixed, int swing,
                             double* fl, double* fr, double* rl, double* rr);
/* Walk (quad_walk_no_plots.m:563-613): step_duration samples per step; phases 2/4/6/8 swing fl/rr/fr/rl. */
void oracle_feet_export_walk(const double* foot_plan, int rows, int n_steps, int step_duration,
                             double* fl, double* fr, double* rl, double* rr);

/* ---- LIP Kalman filter (AMR_code_DART/StateFiltering.cpp), single precision like the reference ---- */
typedef struct {
    float h_com, mass, sampling_time, g;
    float q_process[3][4];      /* x, y, z : 2x2 row-major */
    float q_measurement[3][9];  /* x, y, z : 3x3 row-major */
} oracle_kf_model;
typedef struct { float state[3][5]; float sigma[3][25]; } oracle_kf_state;
typedef struct { float meas[3][3]; float input[3]; } oracle_kf_sample;
/* n_steps calls of StateFiltering::FilterWithKalman (StateFiltering.cpp:77-133) on one filter; zmp (nullable):
 * n_steps x 2, GetZMP() (StateFiltering.cpp:180-186) after each call. */
void oracle_kf_filter(const oracle_kf_model* m, oracle_kf_state* s, const oracle_kf_sample* samples, int n_steps, float* zmp);

#ifdef __cplusplus
}
#endif
#endif
