// feet_kernels.cu -- the stage after the hot path: real-foot placement (the MATLAB scripts' "SECOND QUAD_PROG")
// and the foot-trajectory export in the layout the reference's Controller reads.
//
// Reference map:
//   trotting/compute_two_feet1.m:5-16, walking/compute_one_feet_walk.m:99-116  -> diag_offset()
//   trotting/quad_as_bip_no_plots.m:332-426                                    -> feet_trot_tick()
//   walking/quad_walk_no_plots.m:334-504                                       -> feet_walk_tick()
//   trotting/quad_as_bip_no_plots.m:482-509, walking/quad_walk_no_plots.m:563-613 -> feet_export_kernel
// The work per tick is a handful of scalar operations with a data-dependent update of the instance's foot plan, and
// ticks are sequential: one thread per instance walks its ticks; instances are independent.  The export is one thread
// per (instance, sample).
#include "common.cuh"
#include "launch.h"

namespace ismpc {

#define FP(r, c) fp[(size_t)((r) - 1) * 8 + ((c) - 1)]   // MATLAB 1-based foot_plan(r,c)

// Line through the two fixed feet, line of opposite slope through the ZMP, offset of the ZMP from their intersection.
__device__ __forceinline__ bool diag_offset(const double fixed[4], double zx, double zy, double& m, double& dx, double& dy)
{
    m = (fixed[3] - fixed[1]) / (fixed[2] - fixed[0]);
    const double q = fixed[1] - m * fixed[0];
    if (!(fabs(m) > 0.0) || !isfinite(m)) { dx = 0.0; dy = 0.0; return false; }   // no intersection: nothing changes
    const double xs = (zy + m * zx - q) / (2.0 * m);
    const double ys = m * xs + q;
    dx = zx - xs; dy = zy - ys;
    return true;
}

__device__ __forceinline__ double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ inline void feet_trot_tick(const ismpc_feet_model_t& p, int fs, double zx, double zy, double phi, double* fp, int rows)
{
    if (fs < 1 || fs + 1 > rows) return;
    const bool odd = (fs % 2) == 1;
    const int fc1 = odd ? 3 : 1, fc2 = odd ? 7 : 5;     // feet that stay (x columns, row fs)
    const int mc1 = odd ? 1 : 3, mc2 = odd ? 5 : 7;     // feet that move (x columns, row fs+1)
    const double fixed[4] = {FP(fs, fc1), FP(fs, fc1 + 1), FP(fs, fc2), FP(fs, fc2 + 1)};
    const double fr0 = FP(fs + 1, mc1), fr1 = FP(fs + 1, mc1 + 1), fr2 = FP(fs + 1, mc2), fr3 = FP(fs + 1, mc2 + 1);
    double m, dx, dy;
    if (diag_offset(fixed, zx, zy, m, dx, dy)) {
        double x1, y1, x2, y2;
        if (phi == 3.14159265358979323846 / 2) {          // compute_two_feet1.m:19-25
            x1 = fr0; x2 = fr2;
            y1 = zy - m * (x1 - zx); y2 = zy - m * (x2 - zx);
        } else {                                          // :26-37
            const double t = tan(phi);
            x1 = (zy + m * zx + t * fr0 - fr1) / (t + m); y1 = t * (x1 - fr0) + fr1;
            x2 = (zy + m * zx + t * fr2 - fr3) / (t + m); y2 = t * (x2 - fr2) + fr3;
        }
        if (dy != 0.0 || dx != 0.0) {                     // foot_plan(fsCounter+1,:) = quattro_piedi
            FP(fs + 1, mc1) = x1; FP(fs + 1, mc1 + 1) = y1; FP(fs + 1, mc2) = x2; FP(fs + 1, mc2 + 1) = y2;
            FP(fs + 1, fc1) = fixed[0]; FP(fs + 1, fc1 + 1) = fixed[1]; FP(fs + 1, fc2) = fixed[2]; FP(fs + 1, fc2 + 1) = fixed[3];
        }
    }
    const bool dummy = odd && fs == 1;
    const double d_o = dummy ? p.disp_o_dummy : p.disp_o, d_i = dummy ? p.disp_i_dummy : p.disp_i;
    const double d_f = dummy ? p.disp_forw_dummy : p.disp_forw;
    const int a = odd ? 1 : 7, b = odd ? 5 : 3;
    const double X1 = fmin(FP(fs + 1, a), FP(fs, a) + d_f);
    const double X2 = clipd(FP(fs + 1, a + 1), FP(fs, a + 1) - d_i, FP(fs, a + 1) + d_o);
    const double X3 = fmin(FP(fs + 1, b), FP(fs, b) + d_f);
    const double X4 = clipd(FP(fs + 1, b + 1), FP(fs, b + 1) - d_o, FP(fs, b + 1) + d_i);
    FP(fs + 1, a) = X1; FP(fs + 1, a + 1) = X2; FP(fs + 1, b) = X3; FP(fs + 1, b + 1) = X4;
}

__device__ inline void feet_walk_tick(const ismpc_feet_model_t& p, int counter, int fs, double zx, double zy, double* fp, int rows)
{
    if (counter != 2 && counter != 4 && counter != 6 && counter != 8) return;
    if (fs < 1 || fs + 8 > rows) return;
    const int d1 = (counter <= 4) ? 1 : 3, d2 = (counter <= 4) ? 5 : 7;
    const int mc = counter == 2 ? 7 : counter == 4 ? 3 : counter == 6 ? 5 : 1;
    const double fixed[4] = {FP(fs, d1), FP(fs, d1 + 1), FP(fs, d2), FP(fs, d2 + 1)};
    double m, dx, dy;
    if (diag_offset(fixed, zx, zy, m, dx, dy)) {
        const double xf = FP(fs + 1, mc) + dx, yf = FP(fs + 1, mc + 1) + dy;
        if (dy != 0.0 || dx != 0.0) for (int l = 1; l <= 8; ++l) { FP(fs + l, mc) = xf; FP(fs + l, mc + 1) = yf; }
    }
    const bool dummy = (counter <= 4) && fs <= 4;
    const double d_o = dummy ? p.disp_o_dummy : p.disp_o, d_i = dummy ? p.disp_i_dummy : p.disp_i;
    const double d_f = dummy ? p.disp_forw_dummy : p.disp_forw;
    const bool left = (counter == 2 || counter == 8);
    const double up = left ? d_o : d_i, dn = left ? d_i : d_o;
    const double X1 = fmin(FP(fs + 1, mc), FP(fs, mc) + d_f);
    const double X2 = clipd(FP(fs + 1, mc + 1), FP(fs, mc + 1) - dn, FP(fs, mc + 1) + up);
    for (int l = 1; l <= 8; ++l) {
        FP(fs + l, mc) = X1;
        if (counter != 8 || l == 1) FP(fs + l, mc + 1) = X2;     // quad_walk_no_plots.m:500-503, copied as written
    }
}

// A record is used only if its rows lie inside the tables handed to the call (the caller's records are data, not
// trusted indices): plan rows [plan_first_row, +plan_rows) within foot_plan_rows, timing entries [timing_first,
// +n_timing) within timing_len, fs_counter >= 1 (1-based).  Anything else is skipped: the placement leaves the plan
// untouched, the export writes zeros.
__device__ __forceinline__ bool feet_inst_in_range(const ismpc_feet_inst_t& in, int foot_plan_rows, int timing_len)
{
    return in.plan_first_row >= 0 && in.plan_rows >= 0 && (long long)in.plan_first_row + in.plan_rows <= foot_plan_rows &&
           in.timing_first >= 0 && in.n_timing >= 0 && (long long)in.timing_first + in.n_timing <= timing_len &&
           in.fs_counter >= 1;
}

__global__ void feet_place_kernel(int n, int n_ticks, ismpc_feet_model_t mdl, const ismpc_feet_inst_t* inst,
                                  const int32_t* fs_timing, int timing_len, const double* pred_traj, double* foot_plan,
                                  int foot_plan_rows)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ismpc_feet_inst_t in = inst[i];
    if (!feet_inst_in_range(in, foot_plan_rows, timing_len)) return;
    const int32_t* ft = fs_timing + in.timing_first;
    double* fp = foot_plan + (size_t)in.plan_first_row * 8;
    int j = in.j, fsc = in.fs_counter;
    int counter = fsc;                                   // both start at 1 and advance together (quad_walk_no_plots.m:526-527)
    if (mdl.wrap_counter) counter = (fsc - 1) % 8 + 1;
    for (int t = 0; t < n_ticks; ++t) {
        const double zx = pred_traj[((size_t)i * n_ticks + t) * 2], zy = pred_traj[((size_t)i * n_ticks + t) * 2 + 1];
        if (mdl.gait == ISMPC_GAIT_TROT) feet_trot_tick(mdl, fsc, zx, zy, in.phi, fp, in.plan_rows);
        else feet_walk_tick(mdl, counter, fsc, zx, zy, fp, in.plan_rows);
        if (fsc + 1 <= in.n_timing && j + 1 >= ft[fsc]) {
            fsc += 1;
            counter = mdl.wrap_counter ? (counter == 8 ? 1 : counter + 1) : counter + 1;
        }
        j += 1;
    }
}

__global__ void feet_export_kernel(int n, ismpc_feet_model_t mdl, const ismpc_feet_inst_t* inst, const double* foot_plan,
                                   int foot_plan_rows, int n_steps, int fixed, int swing, double* fl, double* fr, double* rl,
                                   double* rr)
{
    const int per = fixed + swing;
    const long long total = (long long)n * n_steps * per;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int i = (int)(gid / ((long long)n_steps * per));
    const int k = (int)(gid - (long long)i * n_steps * per);
    const int step = k / per + 1, s = k - (step - 1) * per + 1;        // 1-based step and sample within the step
    const ismpc_feet_inst_t in = inst[i];
    // the export reads plan rows only (no timing table)
    const bool ok = in.plan_first_row >= 0 && in.plan_rows >= 0 && (long long)in.plan_first_row + in.plan_rows <= foot_plan_rows;
    const double* fp = foot_plan + (size_t)(ok ? in.plan_first_row : 0) * 8;
    double* dst[4] = {rl, rr, fr, fl};
    int mvA = 0, mvB = 0, kk = 0;                                       // x columns of the swinging feet, swing sample index
    const bool have = ok && step + 1 <= in.plan_rows;
    if (have) {
        if (mdl.gait == ISMPC_GAIT_TROT) {
            if (s > fixed) { kk = s - fixed; if (step % 2 == 1) { mvA = 1; mvB = 5; } else { mvA = 7; mvB = 3; } }
        } else {
            const int c = (step - 1) % 8 + 1;                           // conteggio
            kk = s;
            mvA = c == 2 ? 7 : c == 4 ? 3 : c == 6 ? 5 : c == 8 ? 1 : 0;
        }
    }
    const double z = -0.000032 * kk * kk + 0.0016 * kk;
    for (int f = 0; f < 4; ++f) {
        const int c = 2 * f + 1;
        double x = 0.0, y = 0.0, zz = 0.0;
        if (have) {
            x = FP(step, c); y = FP(step, c + 1);
            if (c == mvA || c == mvB) {
                x = FP(step, c) + (FP(step + 1, c) - FP(step, c)) / swing * kk;
                y = FP(step, c + 1) + (FP(step + 1, c + 1) - FP(step, c + 1)) / swing * kk;
                zz = z;
            }
        }
        double* o = dst[f] + (size_t)gid * 3;
        o[0] = x; o[1] = y; o[2] = zz;
    }
}
#undef FP

// ---------------------------------------------------------------------------------------------------------
// Plan generators: trotting/init_quadruped.m:54-184, walking/init_quadruped2.m:54-300.  One thread per instance;
// rows depend on their predecessors, instances are independent.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void clip_step(double& x, double& y, double forw, double vert, double phi)
{
    if (y > vert || x > forw) {                                   // init_quadruped.m:62-102
        if (phi > atan(vert / forw)) { y = vert; x = vert * cos(phi) / sin(phi); }
        else { x = forw; y = forw * sin(phi) / cos(phi); }
    }
}

__device__ __forceinline__ void diag_center(const double* row, double* c)   // init_quadruped.m:171-183
{
    const double m1 = (row[5] - row[1]) / (row[4] - row[0]), q1 = row[1] - m1 * row[0];   // rear-left -- front-right
    const double m2 = (row[7] - row[3]) / (row[6] - row[2]), q2 = row[3] - m2 * row[2];   // rear-right -- front-left
    const double x = (q2 - q1) / (m1 - m2);
    c[0] = x; c[1] = m1 * x + q1;
}

__global__ void plan_generate_kernel(int n, ismpc_plan_model_t m, const ismpc_plan_req_t* req, double* foot_plan, double* center, int rows)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const double phi = req[idx].phi, dA = req[idx].disp_A;
    const double B = m.disp_B, C = m.disp_C;
    const double vert = fmin(m.disp_i, m.disp_o);
    double xp = dA * cos(phi), yp = dA * sin(phi);
    double xd = xp / 2, yd = yp / 2;
    clip_step(xd, yd, m.disp_forw / 2, vert / 2, phi);
    clip_step(xp, yp, m.disp_forw, vert, phi);
    double* fp = foot_plan + (size_t)idx * rows * 8;
    double* ce = center + (size_t)idx * rows * 2;
    const int N = m.N_gait;
#define BL(r, c) fp[(size_t)(r) * 8 + 0 + (c)]
#define BR(r, c) fp[(size_t)(r) * 8 + 2 + (c)]
#define FR(r, c) fp[(size_t)(r) * 8 + 4 + (c)]
#define FL(r, c) fp[(size_t)(r) * 8 + 6 + (c)]
    for (int r = 0; r < rows; ++r) {
        BL(r, 0) = 0.0; BL(r, 1) = B; BR(r, 0) = 0.0; BR(r, 1) = -B; FL(r, 0) = C; FL(r, 1) = B; FR(r, 0) = C; FR(r, 1) = -B;
        ce[r * 2] = 0.0; ce[r * 2 + 1] = 0.0;
    }
    const double st[2] = {xp, yp};
    if (m.gait == ISMPC_GAIT_TROT) {
        if (N > 1) { BL(1, 0) = xd; FR(1, 0) = C + xd; BL(1, 1) = B + yd; FR(1, 1) = -B + yd; }     // init_quadruped.m:113-117
        for (int j = 3; j <= N; ++j) {                                                              // :120-149
            const int i = j - 1;
            for (int c = 0; c < 2; ++c) {
                if (j % 2 == 0) { BL(i, c) = BL(i - 1, c) + st[c]; FR(i, c) = FR(i - 1, c) + st[c]; BR(i, c) = BR(i - 1, c); FL(i, c) = FL(i - 1, c); }
                else { BR(i, c) = BR(i - 1, c) + st[c]; FL(i, c) = FL(i - 1, c) + st[c]; BL(i, c) = BL(i - 1, c); FR(i, c) = FR(i - 1, c); }
            }
        }
        ce[0] = C / 2;
        for (int k = 1; k < N; ++k) diag_center(fp + (size_t)k * 8, ce + (size_t)k * 2);
    } else {
        // init_quadruped2.m:114-138 (dummy first half-cycle)
        const double dm[2] = {xd, yd};
        const double base[2] = {C, B};
        for (int c = 0; c < 2; ++c) {
            FL(2, c) = base[c] + dm[c]; FL(3, c) = FL(2, c); FL(4, c) = FL(2, c);
            BR(1, c) = BR(0, c); BR(2, c) = BR(0, c); BR(3, c) = BR(2, c); BR(4, c) = BR(3, c) + dm[c];
        }
        int grown = N;
        for (int j = 6; j <= N; j += 8) {                                                           // :141-219
            const int i = j - 1;
            for (int c = 0; c < 2; ++c) {
                FR(i, c) = FR(i - 1, c); FR(i + 1, c) = FR(i, c) + st[c];
                for (int k = 2; k < 8; ++k) FR(i + k, c) = FR(i + 1, c);
                BL(i, c) = BL(i - 1, c); BL(i + 1, c) = BL(i, c); BL(i + 2, c) = BL(i, c); BL(i + 3, c) = BL(i + 2, c) + st[c];
                for (int k = 4; k < 8; ++k) BL(i + k, c) = BL(i + 3, c);
                FL(i, c) = FL(i - 1, c);
                for (int k = 1; k < 5; ++k) FL(i + k, c) = FL(i, c);
                FL(i + 5, c) = FL(i + 4, c) + st[c]; FL(i + 6, c) = FL(i + 5, c); FL(i + 7, c) = FL(i + 5, c);
                BR(i, c) = BR(i - 1, c);
                for (int k = 1; k < 7; ++k) BR(i + k, c) = BR(i, c);
                BR(i + 7, c) = BR(i + 6, c) + st[c];
            }
            if (j + 7 > grown) grown = j + 7;
        }
        ce[0] = C / 2;
        const int rc = grown > N ? grown : N;
        for (int j = 1; j <= N - 4; j += 8) {                                                       // :242-284
            for (int k = 0; k < 8; k += 2) if (j + k - 1 < grown) diag_center(fp + (size_t)(j + k - 1) * 8, ce + (size_t)(j + k - 1) * 2);
            for (int k = 1; k < 8; k += 2)
                if (j + k - 1 < rc) { ce[(size_t)(j + k - 1) * 2] = ce[(size_t)(j + k - 2) * 2]; ce[(size_t)(j + k - 1) * 2 + 1] = ce[(size_t)(j + k - 2) * 2 + 1]; }
        }
    }
#undef BL
#undef BR
#undef FR
#undef FL
}

int plan_generate_launch(int n, const ismpc_plan_model_t& m, const ismpc_plan_req_t* req, double* foot_plan, double* center,
                         int rows, cudaStream_t st)
{
    plan_generate_kernel<<<(n + 63) / 64, 64, 0, st>>>(n, m, req, foot_plan, center, rows);
    return (int)cudaGetLastError();
}

int feet_place_launch(int n, int n_ticks, const ismpc_feet_model_t& m, const ismpc_feet_inst_t* inst, const int32_t* fs_timing,
                      int timing_len, const double* pred_traj, double* foot_plan, int foot_plan_rows, cudaStream_t st)
{
    feet_place_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, n_ticks, m, inst, fs_timing, timing_len, pred_traj, foot_plan,
                                                       foot_plan_rows);
    return (int)cudaGetLastError();
}

int feet_export_launch(int n, const ismpc_feet_model_t& m, const ismpc_feet_inst_t* inst, const double* foot_plan,
                       int foot_plan_rows, int n_steps, int fixed, int swing, double* fl, double* fr, double* rl, double* rr,
                       cudaStream_t st)
{
    const long long total = (long long)n * n_steps * (fixed + swing);
    if (total <= 0) return 0;
    feet_export_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(n, m, inst, foot_plan, foot_plan_rows, n_steps, fixed, swing, fl, fr, rl, rr);
    return (int)cudaGetLastError();
}

}  // namespace ismpc
