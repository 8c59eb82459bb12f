// formc_pair.cuh -- formulation C tick, TWO warps (one 64-thread CTA) per instance: the latency build of the
// warp-per-instance tick (formc_warp.cuh), used while every instance of the batch gets a resident CTA.
//
// A single warp runs the tick as one dependent instruction stream (~4 cycles per instruction, profiles/r1m_*), and at
// 1,024 instances a B200 has issue slots to spare.  Two stages of the tick do not depend on each other until the
// horizontal QPs are solved, so they run side by side:
//   warp 0:  midpoint window mid_x, mid_y (MPCSolver.cpp:167-180) and the anticipative tails (:381-383)
//   warp 1:  stage 1 (vertical QP, feedback law staged by TMA), stage 2 (lambda, LIP matrices), stability row
//   -- CTA barrier --
//   warp 0:  horizontal QP of the x axis          warp 1:  horizontal QP of the y axis      (MPCSolver.cpp:322-398)
//   -- CTA barrier --    warp 1 integrates and writes the record.
// Same arithmetic as formc_tick_warp stage by stage (the single-axis Newton pass is the two-axis pass with one axis
// removed), so results agree to rounding; tests/test_formc_gpu.py holds the two builds to 1e-12 of each other.
#pragma once
#include "formc_warp.cuh"

namespace ismpc {

constexpr int FORMC_PAIR_RED = 48;      // doubles exchanged between the two warps ([16..32): the rollout's state hand-over / the tick's result staging, [32..48): the tick's packed input record)

// CTA barrier of the two warps (a named barrier: the two warps reach it from different code paths)
__device__ __forceinline__ void pair_barrier() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

__host__ __device__ inline size_t formc_pair_smem_bytes(int N) { return formc_warp_smem_bytes(N) + FORMC_PAIR_RED * sizeof(double); }

// One horizontal QP (one axis) by one warp; stability row in sm.av, box centres in `mid` (shared-memory vector).
//   b = -(ps0 pos + ps1 vel) + eta dt tail   (MPCSolver.cpp:381-384).  See formc_tick_warp for the method.
__device__ __forceinline__ void formc_knapsack_axis(const FormCWarpShared& sm, const double* mid, int N, int E, int lane,
                                                    double rho, double bq, double* prim_ax, signed char* act_ax,
                                                    double& u0_out, int& nsat_out, int& fail_out, double& resid_out)
{
    double am = 0.0, t1 = 0.0, t2 = 0.0, mx = 0.0;
#pragma unroll 1
    for (int e = 0; e < E; ++e) {
        const int x = e * 32 + lane;
        const double a = sm.av[x], ai = fabs(a);
        am += a * mid[x];
        t1 += ai; t2 += ai * ai; mx = fmax(mx, ai);
    }
    double aa = t2;
    const double amax = warp_max_nonneg(mx);
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        am += __shfl_xor_sync(ISMPC_FULL_MASK, am, o); aa += __shfl_xor_sync(ISMPC_FULL_MASK, aa, o);
    }
    const double rr = bq - am;
    const double sg = rr >= 0.0 ? 1.0 : -1.0, ra = fabs(rr);
    double tq = 0.0;
    int fail = 0;
    if (__any_sync(ISMPC_FULL_MASK, aa > 0.0)) {
        tq = ra * fast_rcp(aa);                                                       // no row saturated
        if (__any_sync(ISMPC_FULL_MASK, tq * amax > rho)) {
            // prefix candidates: P1(k) = sum_{i<k} |a_i|,  S2(k) = sum_{i>=k} a_i^2  -> lower bound max_k t_k
            double p1 = t1, sf = t2;
#pragma unroll 1
            for (int o = 1; o < 32; o <<= 1) {
                const double a1 = __shfl_up_sync(ISMPC_FULL_MASK, p1, o), a2 = __shfl_down_sync(ISMPC_FULL_MASK, sf, o);
                if (lane >= o) p1 += a1;
                if (lane + o < 32) sf += a2;
            }
            double P1 = p1 - t1, S2 = sf, cb = 0.0;
#pragma unroll 1
            for (int e = 0; e < E; ++e) {
                const double ak = fabs(sm.av[e * 32 + lane]);
                // (S2 is a running difference: rounding residue after the last non-zero row, and the padding beyond the
                //  horizon, are not candidates)
                if (lane * E + e < N && S2 > 1e-12 * aa) cb = fmax(cb, (ra - rho * P1) * fast_rcp(S2));
                P1 += ak; S2 -= ak * ak;
            }
            tq = fmax(tq, warp_max_nonneg(cb));
            int prev = -1;
#pragma unroll 1
            for (int it = 0; it < N + 3; ++it) {
                double q1 = 0.0, q2 = 0.0;
                int cnt = 0;
#pragma unroll 1
                for (int e = 0; e < E; ++e) {
                    const double ai = fabs(sm.av[e * 32 + lane]);
                    const bool s = tq * ai > rho;
                    q1 += s ? ai : 0.0; q2 += s ? 0.0 : ai * ai;
                    cnt += s ? 1 : 0;
                }
#pragma unroll 1
                for (int o = 16; o > 0; o >>= 1) {
                    q1 += __shfl_xor_sync(ISMPC_FULL_MASK, q1, o); q2 += __shfl_xor_sync(ISMPC_FULL_MASK, q2, o);
                }
                cnt = __reduce_add_sync(ISMPC_FULL_MASK, cnt);
                bool done = cnt == prev;                                              // same set as the one tq was solved for
                if (!done) {
                    prev = cnt;
                    const double rem = ra - rho * q1;
                    if (!(q2 > 0.0)) { if (rem > 1e-12 * fmax(1.0, ra)) fail = 1; done = true; }
                    else {
                        const double tn = rem * fast_rcp(q2);
                        done = fabs(tn - tq) <= 1e-13 * tq;                           // the guess was this set's solution
                        tq = (it == 0 || tn > tq) ? tn : tq;                          // (a guess may sit a rounding above t*)
                    }
                }
                if (__all_sync(ISMPC_FULL_MASK, done)) break;
                ISMPC_WCOUNT(29);
            }
        }
    } else if (ra > 1e-12) fail = 1;
    const double nu = sg * tq;
    double au = 0.0, u0 = 0.0;
    int ns = 0;
#pragma unroll 1
    for (int e = E - 1; e >= 0; --e) {
        const int i = lane * E + e, x = e * 32 + lane;
        const double a = sm.av[x];
        const double d = nu * a;
        const double ue = mid[x] + fmin(fmax(d, -rho), rho);
        au += a * ue;
        ns += (tq * fabs(a) > rho);
        u0 = ue;                                                                      // e = 0 is written last
        if (i < N) {
            if (prim_ax) prim_ax[i] = ue;
            if (act_ax) act_ax[i] = (signed char)((d > rho) ? 1 : ((d < -rho) ? -1 : 0));
        }
    }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) au += __shfl_xor_sync(ISMPC_FULL_MASK, au, o);
    ns = __reduce_add_sync(ISMPC_FULL_MASK, ns);
    u0_out = __shfl_sync(ISMPC_FULL_MASK, u0, 0);
    nsat_out = ns; fail_out = fail; resid_out = fabs(au - bq);
}

// One tick of one instance by the two warps of a 64-thread CTA.  red: FORMC_PAIR_RED doubles of shared memory.
// Only warp 1 returns a meaningful record in r (it writes it out); both warps execute the same CTA barriers.
__device__ __forceinline__ void formc_tick_pair(const FormCWarpShared& sm, double* red, const ismpc_formc_model_t& mdl,
                                                const FormCTables& T, const FormCRiccati& R, const ismpc_state_t& st,
                                                const ismpc_walk_t& wk, const ismpc_formc_inst_t& in,
                                                const double* __restrict__ plan_all, double* ws, ismpc_formc_out_t& r,
                                                double* prim, signed char* act, uint32_t& bar_parity)
{
    const int lane = lane_id();
    const int role = __shfl_sync(ISMPC_FULL_MASK, (int)(threadIdx.x >> 5), 0);       // 0: midpoints + x axis, 1: vertical + y axis
    const int N = mdl.N, E = formc_warp_epl(N);
    const double dt = mdl.dt, mass = mdl.mass, g = mdl.g;
    const double h = in.com_height;
    const double eta = sqrt(g / h);                       // parameters.cpp:41
    const int S = in.S, F = in.F_ds, per = S + F;
    const int k0 = (int)(wk.sim_time / (dt / mdl.dtc));   // MPCSolver.cpp:259,329
    int status = 0;

    r.next = st; r.zmp_in[0] = r.zmp_in[1] = 0.0; r.fz0 = 0.0; r.lambda0 = 0.0; r.kkt_res = 0.0;
    r.status = 0; r.iters[0] = r.iters[1] = r.iters[2] = 0;
    if (__any_sync(ISMPC_FULL_MASK, k0 < 0 || per <= 0 || (long long)k0 + 2 * N > (long long)in.n_steps * per)) {
        r.status = ISMPC_ST_WINDOW;                        // (both warps take this exit: the records are CTA-uniform)
        return;
    }
    pair_barrier();                                       // the previous instance's shared-memory traffic is done

    const double z0 = st.com_pos[2], zd0 = st.com_vel[2];
    const double c1 = dt * dt / mass;
    const bool running = wk.footstep_counter > 1;
    int ne = 0, c_lo = 0;
    if (running) formc_flight_range(N, S, F, wk.mpc_iter, c_lo, ne);
    const double* tab = nullptr;
    const double* law = nullptr;
    if (ne == 0) { tab = R.none; law = R.law_none; }
    else if (R.gait != nullptr && S == R.gS && F == R.gF && wk.mpc_iter >= 0 && wk.mpc_iter < per) {
        tab = R.gait + (size_t)wk.mpc_iter * N * FORMC_RIC_W;
        law = R.law_gait + (size_t)wk.mpc_iter * formc_law_pattern_doubles(N);
    }
    // warp 1 starts the bulk copy of its pattern's law table before anything else that waits on memory
    const bool staged = __any_sync(ISMPC_FULL_MASK, law != nullptr);
    if (staged && role == 1) formc_law_stage(sm, law, N, lane);
    const double* rows = plan_all + (size_t)in.plan_first_row * 4;
    const float rcp_per = 1.0f / (float)per;
    int q0, r0;                                               // window start: step index, in-step sample
    fast_divmod(k0, per, rcp_per, q0, r0);
    // Flat reference decided from the plan rows (both warps, same answer): every row the first N samples can touch
    // has the same z, and the "last step stays 0" rule (:167) is not in play unless that z is 0.
    int qn, rn;
    fast_divmod(r0 + N - 1, per, rcp_per, qn, rn);
    const int step_last = q0 + qn;                                                    // step of sample N-1
    int row_last = step_last + 1;
    if (row_last > in.n_steps - 1) row_last = in.n_steps - 1;
    double zr = 0.0;
    if (q0 + lane <= row_last) zr = __ldg(rows + 4 * (q0 + lane) + 2);
    const double zc = __shfl_sync(ISMPC_FULL_MASK, zr, 0);
    const bool rows_fit = row_last - q0 < 32;
    const bool flat = __all_sync(ISMPC_FULL_MASK, rows_fit && (q0 + lane > row_last || zr == zc)) &&
                      (step_last < in.n_steps - 1 || zc == 0.0);
    const double rq0 = -mdl.q_p * (h + zc * 1.0);
    const bool use_law = __all_sync(ISMPC_FULL_MASK, flat && law != nullptr);

    double lam0 = 0.0, fz0 = 0.0, viol = 0.0;
    int it_z = 0;
    ISMPC_WPHASE_BEGIN;
    if (role == 0) {
        // ---------------- warp 0: midpoint window and anticipative tails ----------------
        double tx = 0.0, ty = 0.0;
        const double invF = fast_rcp((double)F);
        const double qd = exp(-dt * eta);
        double dl = exp(-dt * eta * (double)(lane * E));                             // deltas (:183-184), dl_i = qd^i
        int qi, ri, qt, rt;
        fast_divmod(r0 + lane * E, per, rcp_per, qi, ri);
        fast_divmod(r0 + N + lane * E, per, rcp_per, qt, rt);
        qi += q0; qt += q0;
        int cq = -1, ct = -1;
        double ax = 0.0, ay = 0.0, az = 0.0, bx = 0.0, by = 0.0, bz = 0.0, cx_ = 0.0, cy_ = 0.0, dx_ = 0.0, dy_ = 0.0;
#pragma unroll 1
        for (int e = 0; e < E; ++e) {
            const int i = lane * E + e, x = e * 32 + lane;
            double mxv = 0.0, myv = 0.0, mzv = 0.0;
            if (i < N) {
                if (qi != cq) {
                    cq = qi;
                    ax = ay = az = bx = by = bz = 0.0;                               // last step's rows stay 0 (:167)
                    if (qi < in.n_steps - 1) {
                        ax = __ldg(rows + 4 * qi); ay = __ldg(rows + 4 * qi + 1); az = __ldg(rows + 4 * qi + 2);
                        bx = __ldg(rows + 4 * qi + 4); by = __ldg(rows + 4 * qi + 5); bz = __ldg(rows + 4 * qi + 6);
                    }
                }
                if (qt != ct) {
                    ct = qt;
                    cx_ = cy_ = dx_ = dy_ = 0.0;
                    if (qt < in.n_steps - 1) {
                        cx_ = __ldg(rows + 4 * qt); cy_ = __ldg(rows + 4 * qt + 1);
                        dx_ = __ldg(rows + 4 * qt + 4); dy_ = __ldg(rows + 4 * qt + 5);
                    }
                }
                const double w = ri < S ? 0.0 : (double)(ri - S) * invF;
                mxv = ax * 1.0 + (bx - ax) * w; myv = ay * 1.0 + (by - ay) * w; mzv = az * 1.0 + (bz - az) * w;
                const double wt = rt < S ? 0.0 : (double)(rt - S) * invF;
                tx += dl * (cx_ * 1.0 + (dx_ - cx_) * wt); ty += dl * (cy_ * 1.0 + (dy_ - cy_) * wt);
            }
            if (++ri == per) { ri = 0; ++qi; }
            if (++rt == per) { rt = 0; ++qt; }
            sm.mx[x] = mxv; sm.my[x] = myv; sm.rq[x] = -mdl.q_p * (h + mzv);
            dl *= qd;
        }
#pragma unroll 1
        for (int o = 16; o > 0; o >>= 1) {
            tx += __shfl_xor_sync(ISMPC_FULL_MASK, tx, o); ty += __shfl_xor_sync(ISMPC_FULL_MASK, ty, o);
        }
        if (lane == 0) { red[0] = tx; red[1] = ty; }
        if (!use_law) pair_barrier();                     // hand sm.rq to warp 1 (the Riccati scans need it)
        ISMPC_WPHASE(8);
    } else {
        // ---------------- warp 1: stage 1, stage 2, stability row ----------------
        bool bad = false;
        if (staged) { mbar_wait(sm.bar, bar_parity); bar_parity ^= 1u; }
        if (use_law) formc_law_apply(sm, N, E, lane, rq0, z0, zd0, dt, g, mdl.fz_max, bad, viol);
        else {
            pair_barrier();                               // sm.rq from warp 0
            if (__any_sync(ISMPC_FULL_MASK, tab != nullptr)) riccati_load(sm, tab, N, E, lane);
            else {
                RicP P{0.0, 0.0, 0.0};
                const double rho_u = mdl.q_u * mass * mass;
                int ol = (N - 1) / E, oe = (N - 1) - ol * E;
#pragma unroll 1
                for (int k = N - 1; k >= 0; --k) {
                    double a, b, c, d;
                    riccati_step(P, k >= c_lo && k < c_lo + ne, dt, rho_u, mdl.q_p, mdl.q_v, g, a, b, c, d);
                    if (lane == ol) { const int x = oe * 32 + lane; sm.ta[x] = a; sm.tb[x] = b; sm.tc[x] = c; sm.td[x] = d; }
                    if (--oe < 0) { oe = E - 1; --ol; }
                }
            }
            riccati_solve(sm, N, E, lane, dt, mass, g, z0 + dt * zd0, zd0);
#pragma unroll 1
            for (int e = 0; e < E; ++e) {
                const int k = lane * E + e;
                if (k < N) {
                    const double base = 1.0 * z0 + ((double)(k + 1) * dt) * zd0 - g * (dt * dt) * (0.5 * (double)k * (double)(k + 1));
                    const double v = sm.p[e * 32 + lane] - base;
                    bad = bad || (fmin(v + 1e-10, (mdl.fz_max - v) + 1e-10 * (1.0 + fabs(mdl.fz_max))) < 0.0);
                    viol = fmax(viol, fmax(-v, v - mdl.fz_max));
                }
            }
        }
        signed char* zstate = nullptr;
        if (__any_sync(ISMPC_FULL_MASK, bad)) {
            // general path: unconstrained minimiser -> workspace -> dual active set -> back to shared memory
            if (ne > 0) {
                bool b2 = false; double v2 = 0.0;
                if (flat) {
                    __syncwarp();
                    formc_law_stage(sm, R.law_none, N, lane);
                    mbar_wait(sm.bar, bar_parity); bar_parity ^= 1u;
                    formc_law_apply(sm, N, E, lane, rq0, z0, zd0, dt, g, mdl.fz_max, b2, v2);
                } else {
                    riccati_load(sm, R.none, N, E, lane);
                    riccati_solve(sm, N, E, lane, dt, mass, g, z0 + dt * zd0, zd0);
                }
            }
            pdl_dependency_wait();         // the workspace slice is shared with the previous launch on this handle (no-op unless launched as its programmatic dependent)
            for (int e = 0; e < E; ++e) if (lane * E + e < N) ws[lane * E + e] = sm.f[e * 32 + lane];
            __syncwarp();
            status |= formc_vertical_general(N, T, c1, mdl.fz_max, ws, c_lo, ne, &it_z, &zstate);
            viol = 0.0;
            for (int e = 0; e < E; ++e) {
                const int k = lane * E + e;
                if (k < N) {
                    const double base = 1.0 * z0 + ((double)(k + 1) * dt) * zd0 - g * (dt * dt) * (0.5 * (double)k * (double)(k + 1));
                    const double v = ws[N + k];
                    sm.f[e * 32 + lane] = ws[k];
                    sm.p[e * 32 + lane] = v + base;
                    viol = fmax(viol, fmax(-v, v - mdl.fz_max));
                }
            }
            __syncwarp();
        }
        if (prim) for (int e = 0; e < E; ++e) if (lane * E + e < N) prim[lane * E + e] = sm.f[e * 32 + lane];
        if (act) for (int e = 0; e < E; ++e) if (lane * E + e < N) act[lane * E + e] = zstate ? zstate[lane * E + e] : (signed char)0;

        // stage 2 + chunk products of the stability row (see formc_tick_warp)
        double p00 = 1.0, p01 = 0.0, p10 = 0.0, p11 = 1.0;
        double lam_first = 0.0, f_first = 0.0;
#pragma unroll 1
        for (int e = E - 1; e >= 0; --e) {
            const int x = e * 32 + lane;
            double ch = 1.0, shs = 0.0, ssh = 0.0;
            if (lane * E + e < N) {
                const double fe = sm.f[x];
                const double zacc = (1.0 / mass) * fe - g;
                const double l = (g + zacc) * fast_rcp(sm.p[x]);
                lip_matrices(l, dt, ch, shs, ssh);
                lam_first = l; f_first = fe;
                const double n00 = p00 * ch + p01 * ssh, n01 = p00 * shs + p01 * ch;
                const double n10 = p10 * ch + p11 * ssh, n11 = p10 * shs + p11 * ch;
                p00 = n00; p01 = n01; p10 = n10; p11 = n11;
            }
            sm.tc[x] = ch; sm.td[x] = shs; sm.om[x] = ssh;
        }
        fz0 = __shfl_sync(ISMPC_FULL_MASK, f_first, 0);
        lam0 = __shfl_sync(ISMPC_FULL_MASK, lam_first, 0);
        double ps0 = 0.0, ps1 = 0.0;
        if (__any_sync(ISMPC_FULL_MASK, lam0 > 2.0)) {
#pragma unroll 1
            for (int o = 1; o < 32; o <<= 1) {
                const double q00 = __shfl_down_sync(ISMPC_FULL_MASK, p00, o), q01 = __shfl_down_sync(ISMPC_FULL_MASK, p01, o);
                const double q10 = __shfl_down_sync(ISMPC_FULL_MASK, p10, o), q11 = __shfl_down_sync(ISMPC_FULL_MASK, p11, o);
                if (lane + o < 32) {
                    const double n00 = q00 * p00 + q01 * p10, n01 = q00 * p01 + q01 * p11;
                    const double n10 = q10 * p00 + q11 * p10, n11 = q10 * p01 + q11 * p11;
                    p00 = n00; p01 = n01; p10 = n10; p11 = n11;
                }
            }
            double t00 = __shfl_down_sync(ISMPC_FULL_MASK, p00, 1), t01 = __shfl_down_sync(ISMPC_FULL_MASK, p01, 1);
            double t10 = __shfl_down_sync(ISMPC_FULL_MASK, p10, 1), t11 = __shfl_down_sync(ISMPC_FULL_MASK, p11, 1);
            if (lane == 31) { t00 = 1; t01 = 0; t10 = 0; t11 = 1; }
            const double cs0 = 1.0, cs1 = 1.0 / eta;                                 // C_sc (:375-377), nominal eta
            double c0 = cs0 * t00 + cs1 * t10, c1r = cs0 * t01 + cs1 * t11;
#pragma unroll 1
            for (int e = E - 1; e >= 0; --e) {
                const int x = e * 32 + lane;
                const double ch = sm.tc[x], shs = sm.td[x], ssh = sm.om[x];
                sm.av[x] = c0 * (1.0 - ch) + c1r * (-ssh);
                const double n0 = c0 * ch + c1r * ssh, n1 = c0 * shs + c1r * ch;
                c0 = n0; c1r = n1;
            }
            ps0 = __shfl_sync(ISMPC_FULL_MASK, c0, 0); ps1 = __shfl_sync(ISMPC_FULL_MASK, c1r, 0);
        }
        if (lane == 0) { red[2] = ps0; red[3] = ps1; red[4] = lam0; }
        ISMPC_WPHASE(9);
    }
    pair_barrier();
    ISMPC_WPHASE(10);                                       // ---- midpoints, tails, stability row are in shared memory ----

    // ================= STAGE 3: one horizontal QP per warp (MPCSolver.cpp:322-398) =================
    lam0 = red[4];
    const bool horizontal = __any_sync(ISMPC_FULL_MASK, lam0 > 2.0);
    double kkt = fmax(0.0, viol);
    if (horizontal) {
        const double rho = running ? in.box_w / 2 : in.box_w_init / 2;                // (:328-338)
        const double ps0 = red[2], ps1 = red[3];
        const double posq = role == 0 ? st.com_pos[0] : st.com_pos[1], velq = role == 0 ? st.com_vel[0] : st.com_vel[1];
        const double bq = -(ps0 * posq + ps1 * velq) + eta * dt * (role == 0 ? red[0] : red[1]);       // (:381-384)
        double u0, resid; int ns, fail;
        formc_knapsack_axis(sm, role == 0 ? sm.mx : sm.my, N, E, lane, rho, bq,
                            prim ? prim + (1 + role) * N : nullptr, act ? act + (1 + role) * N : nullptr, u0, ns, fail, resid);
        if (role == 0 && lane == 0) { red[8] = u0; red[9] = (double)ns; red[10] = (double)fail; red[11] = resid; }
        ISMPC_WPHASE(11);
        pair_barrier();
        if (role == 1) {
            const double ux0 = red[8], uy0 = u0;
            if ((int)red[10]) status |= ISMPC_ST_X_FAIL;
            if (fail) status |= ISMPC_ST_Y_FAIL;
            kkt = fmax(warp_max_nonneg(kkt), fmax(resid, red[11]));
            r.zmp_in[0] = ux0; r.zmp_in[1] = uy0;
            r.iters[1] = (int)red[9]; r.iters[2] = ns;
        }
    } else {
        status |= ISMPC_ST_XY_SKIPPED;
        kkt = warp_max_nonneg(kkt);
        if (prim) for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) prim[N + i] = 0.0;
        if (act) for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) act[N + i] = 0;
        pair_barrier();
    }
    if (role == 1) {
        // ================= integrate (MPCSolver.cpp:402-422, 274-278) =================
        double nz0 = 1.0 * z0 + dt * zd0, nz1 = zd0 + (dt / mass) * fz0 - dt * g;
        if (isnan(nz0)) { nz0 = h; status |= ISMPC_ST_NAN_GUARD; }
        if (isnan(nz1)) { nz1 = 0.0; status |= ISMPC_ST_NAN_GUARD; }
        double a00, a01, a10, a11, b0, b1;
        if (lam0 < 2.0) { a00 = 1.0; a01 = dt; a10 = 0.0; a11 = 1.0; b0 = 0.0; b1 = 0.0; }
        else {
            double ch, shs, ssh;
            lip_matrices(lam0, dt, ch, shs, ssh);
            a00 = ch; a01 = shs; a10 = ssh; a11 = ch; b0 = 1.0 - ch; b1 = -ssh;
        }
        const double ux0 = r.zmp_in[0], uy0 = r.zmp_in[1];
        r.next.com_pos[0] = a00 * st.com_pos[0] + a01 * st.com_vel[0] + b0 * ux0;
        r.next.com_vel[0] = a10 * st.com_pos[0] + a11 * st.com_vel[0] + b1 * ux0;
        r.next.com_pos[1] = a00 * st.com_pos[1] + a01 * st.com_vel[1] + b0 * uy0;
        r.next.com_vel[1] = a10 * st.com_pos[1] + a11 * st.com_vel[1] + b1 * uy0;
        r.next.com_pos[2] = nz0; r.next.com_vel[2] = nz1;
        r.fz0 = fz0; r.lambda0 = lam0; r.kkt_res = kkt;
        r.status = status; r.iters[0] = it_z;
    }
}

}  // namespace ismpc
