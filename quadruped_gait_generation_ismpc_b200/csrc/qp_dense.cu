// qp_dense.cu -- generic dense QP batch: the solveQP seam (AMR_code_DART/utils.cpp:89-139),
//     min 1/2 x'Hx + g'x   s.t.  lbA <= A x <= ubA,      H symmetric positive definite,
// for n independent problems of one shape.  Same dual active-set engine as the structured paths (das.cuh);
// here H^-1 is not cheap, so a set-up pass builds per-problem tables once --
//     H = L L',  Hinv = L^-T L^-1,  D = A Hinv (rows = H^-1 a_i),  Sfull = A Hinv A',  x0 = -Hinv g,  rv0 = A x0
// (the "condensing" GEMMs of this path; FP64 on CUDA cores, tiled through shared memory) -- after which
// every Schur-complement entry, step direction and row-value update of the active-set loop is a coalesced
// table lookup.  One warp per QP; the first R rows of the packed inverse factor of the working-set Schur
// complement (das.cuh) live in shared memory, the rest in the problem's global workspace slice.
#include "common.cuh"
#include "das.cuh"
#include "launch.h"

namespace ismpc {

struct DenseLayout {          // offsets (in doubles) inside one problem's workspace slice
    size_t Lh, Linv, Hinv, D, S, x0, rv, az, z, Lw, mu, r, y, ints, total;
    int qmax;
};

__host__ __device__ inline DenseLayout dense_layout(int nV, int nC)
{
    DenseLayout l;
    const size_t vv = (size_t)nV * nV, cv = (size_t)nC * nV, cc = (size_t)nC * nC;
    l.qmax = (nV + 1 < nC ? nV + 1 : nC); if (l.qmax < 1) l.qmax = 1;
    size_t o = 0;
    l.Lh = o; o += vv; l.Linv = o; o += vv; l.Hinv = o; o += vv;
    l.D = o; o += cv; l.S = o; o += cc;
    l.x0 = o; o += nV; l.rv = o; o += nC; l.az = o; o += nC; l.z = o; o += nV;
    l.Lw = o; o += (size_t)l.qmax * (l.qmax + 1) / 2;
    l.mu = o; o += l.qmax; l.r = o; o += l.qmax; l.y = o; o += l.qmax;
    l.ints = o; o += (size_t)(l.qmax * sizeof(int) + l.qmax + nC + 15) / 8 + 2;
    l.total = (o + 1) & ~(size_t)1;
    return l;
}

size_t qp_dense_work_doubles(int n, int nV, int nC) { return dense_layout(nV, nC).total * (size_t)n + 16; }

// ---- set-up kernels (batched over problems with blockIdx.z / blockIdx.y) --------------------------------
__global__ void dense_copy_H(int nV, const double* H, double* work, size_t stride, size_t offL)
{
    const size_t p = blockIdx.y;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < (size_t)nV * nV) work[p * stride + offL + e] = H[p * nV * nV + e];
}

__global__ void dense_cholesky(int N, double* work, size_t stride, size_t offL, int32_t* status)
{
    double* A = work + (size_t)blockIdx.x * stride + offL;
    __shared__ double piv;
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int k = 0; k < N; ++k) {
        if (threadIdx.x == 0) {
            double d = A[(size_t)k * N + k];
            if (!(d > 0.0)) { bad = 1; d = 1.0; }
            piv = sqrt(d);
            A[(size_t)k * N + k] = piv;
        }
        __syncthreads();
        const double p = piv;
        for (int i = k + 1 + threadIdx.x; i < N; i += blockDim.x) A[(size_t)i * N + k] /= p;
        __syncthreads();
        const int rem = N - k - 1;
        for (int e = threadIdx.x; e < rem * rem; e += blockDim.x) {
            int i = k + 1 + e / rem, j = k + 1 + e % rem;
            if (j <= i) A[(size_t)i * N + j] -= A[(size_t)i * N + k] * A[(size_t)j * N + k];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) status[blockIdx.x] = bad ? ISMPC_ST_QP_FAIL : 0;
}

__global__ void dense_tri_inverse(int N, double* work, size_t stride, size_t offL, size_t offLinv)
{
    const double* L = work + (size_t)blockIdx.y * stride + offL;
    double* Linv = work + (size_t)blockIdx.y * stride + offLinv;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    for (int i = 0; i < N; ++i) {
        if (i < c) { Linv[(size_t)i * N + c] = 0.0; continue; }
        double s = (i == c) ? 1.0 : 0.0;
        for (int k = c; k < i; ++k) s -= L[(size_t)i * N + k] * Linv[(size_t)k * N + c];
        Linv[(size_t)i * N + c] = s / L[(size_t)i * N + i];
    }
}

// C (M x Nc) = X (M x K) * Y (Nc x K)' on the FP64 tensor cores (DMMA: mma.sync.aligned.m8n8k4.row.col.f64), batched over
// blockIdx.z.  This is the one GEMM-shaped piece of the path -- the "condensing" products of the dense solveQP seam:
// Hinv = Linv' Linv, D = A Hinv, S = A D' -- and BASELINE.json's tensor-core clause applies to it: FP64 in, FP64
// accumulate, so the 1e-6 parity bound is not touched (the same products on the CUDA cores agree to rounding).
// Element (r, k) of X sits at X[r * sxr + k * sxk] (likewise Y), which covers the row-major operands and the transposed
// read of Linv.  CTA = 4 warps, 64 x 64 tile of C, each warp 32 x 32 = 4 x 4 DMMA tiles (32 accumulators per thread);
// K in chunks of 16 through shared memory, rows padded to 20 doubles so that the 8 x 4 / 4 x 8 fragment loads of a
// half-warp fall into 16 distinct 8-byte banks.
constexpr int DG_T = 64, DG_KC = 16, DG_LD = 20;

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128)
dense_gemm_nt(int M, int Nc, int K, const double* Xb, size_t sx, int sxr, int sxk, const double* Yb, size_t sy, int syr, int syk,
              double* Cb, size_t sc)
{
    __shared__ double xs[DG_T * DG_LD], ys[DG_T * DG_LD];
    const double* X = Xb + (size_t)blockIdx.z * sx;
    const double* Y = Yb + (size_t)blockIdx.z * sy;
    double* C = Cb + (size_t)blockIdx.z * sc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.y * DG_T, col0 = blockIdx.x * DG_T;
    const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;          // this warp's 32 x 32 corner inside the tile
    const int g = lane >> 2, t4 = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    for (int k0 = 0; k0 < K; k0 += DG_KC) {
        // stage the two 64 x 16 operand slabs; the index that is contiguous in global memory runs fastest over the threads
        for (int e = tid; e < DG_T * DG_KC; e += 128) {
            int r, k;
            if (sxk == 1) { r = e / DG_KC; k = e - r * DG_KC; } else { k = e / DG_T; r = e - k * DG_T; }
            const int gr = row0 + r, gk = k0 + k;
            xs[r * DG_LD + k] = (gr < M && gk < K) ? X[(size_t)gr * sxr + (size_t)gk * sxk] : 0.0;
            if (syk == 1) { r = e / DG_KC; k = e - r * DG_KC; } else { k = e / DG_T; r = e - k * DG_T; }
            const int gc = col0 + r, gk2 = k0 + k;
            ys[r * DG_LD + k] = (gc < Nc && gk2 < K) ? Y[(size_t)gc * syr + (size_t)gk2 * syk] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < DG_KC; kk += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = xs[(wr + 8 * i + g) * DG_LD + kk + t4];     // A fragment: row g, column t4
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = ys[(wc + 8 * j + g) * DG_LD + kk + t4];     // B fragment: k = t4, column g
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }
    // C fragment: row g, columns 2 t4 and 2 t4 + 1 of every 8 x 8 tile
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = row0 + wr + 8 * i + g, c = col0 + wc + 8 * j + 2 * t4;
            if (r < M) {
                if (c < Nc) C[(size_t)r * Nc + c] = acc[i][j][0];
                if (c + 1 < Nc) C[(size_t)r * Nc + c + 1] = acc[i][j][1];
            }
        }
}

// The same product on the CUDA cores (16 x 16 shared-memory tiles): kept as the cross-check of the tensor-core kernel
// (ismpc_set_option "dense_dmma" = 0) and for the record of what DMMA buys (3.8 ms -> 1.6 ms per product of the
// nV = 206 / nC = 208 shape over 1,024 problems).
__global__ void dense_gemm_nt_cc(int M, int Nc, int K, const double* Xb, size_t sx, int sxr, int sxk, const double* Yb, size_t sy,
                                 int syr, int syk, double* Cb, size_t sc)
{
    __shared__ double xs[16][17], ys[16][17];
    const double* X = Xb + (size_t)blockIdx.z * sx;
    const double* Y = Yb + (size_t)blockIdx.z * sy;
    double* C = Cb + (size_t)blockIdx.z * sc;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
    double acc = 0.0;
    for (int k0 = 0; k0 < K; k0 += 16) {
        int xr = blockIdx.y * 16 + ty, xk = k0 + tx;
        xs[ty][tx] = (xr < M && xk < K) ? X[(size_t)xr * sxr + (size_t)xk * sxk] : 0.0;
        int yr = blockIdx.x * 16 + ty, yk = k0 + tx;
        ys[ty][tx] = (yr < Nc && yk < K) ? Y[(size_t)yr * syr + (size_t)yk * syk] : 0.0;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) acc += xs[ty][kk] * ys[tx][kk];
        __syncthreads();
    }
    if (row < M && col < Nc) C[(size_t)row * Nc + col] = acc;
}

static void dense_gemm_launch(int use_dmma, int n, int M, int Nc, int K, const double* X, size_t sx, int sxr, int sxk,
                              const double* Y, size_t sy, int syr, int syk, double* C, size_t sc, cudaStream_t st)
{
    if (use_dmma)
        dense_gemm_nt<<<dim3((Nc + DG_T - 1) / DG_T, (M + DG_T - 1) / DG_T, n), 128, 0, st>>>(M, Nc, K, X, sx, sxr, sxk, Y, sy, syr, syk, C, sc);
    else
        dense_gemm_nt_cc<<<dim3((Nc + 15) / 16, (M + 15) / 16, n), dim3(16, 16), 0, st>>>(M, Nc, K, X, sx, sxr, sxk, Y, sy, syr, syk, C, sc);
}

// x0 = -Hinv g ; rv0 = A x0 ; one CTA per problem
__global__ void dense_x0_rv(int nV, int nC, const double* g, const double* A, double* work, size_t stride,
                            DenseLayout l, double* x_out)
{
    const size_t p = blockIdx.x;
    double* w = work + p * stride;
    const double* Hinv = w + l.Hinv;
    const double* gp = g + p * nV;
    for (int i = threadIdx.x; i < nV; i += blockDim.x) {
        double s = 0.0;
        for (int j = 0; j < nV; ++j) s += Hinv[(size_t)j * nV + i] * gp[j];
        w[l.x0 + i] = -s;
        x_out[p * nV + i] = -s;
    }
    __syncthreads();
    const double* Ap = A + p * (size_t)nC * nV;
    for (int c = threadIdx.x; c < nC; c += blockDim.x) {
        double s = 0.0;
        for (int j = 0; j < nV; ++j) s += Ap[(size_t)c * nV + j] * w[l.x0 + j];
        w[l.rv + c] = s;
    }
}

// ---- the active-set policy over the tables ---------------------------------------------------------------
struct DenseProb {
    int nV, nC;
    const double *D, *S, *lb, *ub;
    double *rv, *az;
    __device__ int m() const { return nC; }
    __device__ int nvar() const { return nV; }
    __device__ double lo(int i) const { return lb[i]; }
    __device__ double hi(int i) const { return ub[i]; }
    __device__ void eval(const double*, double*) const { __syncwarp(); }      // rv is maintained by on_step
    __device__ double schur(int a, int b) const { return S[(size_t)a * nC + b]; }
    __device__ void step_dir(int idp, int sgp, const int* wid, const signed char* wsg, const double* r, int q,
                             double* z) const
    {
        const int lane = lane_id();
        for (int i = lane; i < nV; i += 32) {
            double acc = (double)sgp * D[(size_t)idp * nV + i];
            for (int k = 0; k < q; ++k) acc -= (double)wsg[k] * r[k] * D[(size_t)wid[k] * nV + i];
            z[i] = acc;
        }
        for (int i = lane; i < nC; i += 32) {       // A z through the symmetric Schur table rows
            double acc = (double)sgp * S[(size_t)idp * nC + i];
            for (int k = 0; k < q; ++k) acc -= (double)wsg[k] * r[k] * S[(size_t)wid[k] * nC + i];
            az[i] = acc;
        }
        __syncwarp();
    }
    __device__ void on_step(double t) const
    {
        const int lane = lane_id();
        for (int i = lane; i < nC; i += 32) rv[i] += t * az[i];
        __syncwarp();
    }
};

constexpr int DENSE_R = 64;      // rows of the inverse factor kept in shared memory (16.6 KB per warp)
constexpr int DENSE_WARPS = 4;   // QPs per CTA

__global__ void dense_das_kernel(int n, int nV, int nC, const double* lbA, const double* ubA, double* work,
                                 size_t stride, DenseLayout l, int R, double* x, double* y, signed char* ws,
                                 int32_t* status, int32_t* iters)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int p = blockIdx.x * DENSE_WARPS + warp_id();
    if (p >= n) return;
    const int lane = lane_id();
    double* w = work + (size_t)p * stride;
    DasWork dw;
    dw.Js = reinterpret_cast<double*>(smem_raw) + (size_t)warp_id() * tri(R, 0);
    dw.Jg = w + l.Lw; dw.R = R;
    dw.mu = w + l.mu; dw.r = w + l.r; dw.y = w + l.y;
    dw.wid = reinterpret_cast<int*>(w + l.ints);
    dw.wsg = reinterpret_cast<signed char*>(dw.wid + l.qmax);
    dw.state = dw.wsg + l.qmax;
    dw.q = 0; dw.neq = 0; dw.qmax = l.qmax;
    const double* lb = lbA + (size_t)p * nC;
    const double* ub = ubA + (size_t)p * nC;
    double* xp = x + (size_t)p * nV;
    int st = status[p];
    int it = 0;
    if (st == 0 && nC > 0) {
        for (int i = lane; i < nC; i += 32) dw.state[i] = 0;
        __syncwarp();
        DenseProb pb{nV, nC, w + l.D, w + l.S, lb, ub, w + l.rv, w + l.az};
        // equalities first, in row order; never dropped (enableEqualities, qpOASES/Options.cpp:191-218)
        for (int c = 0; c < nC; ++c) {
            const double lo = lb[c], hi = ub[c];
            if (fabs(hi - lo) <= 1e-12 * fmax(1.0, fabs(lo))) {
                const double val = pb.rv[c];
                int rc = das_add_equality(pb, dw, xp, w + l.z, c, val, lo);
                if (rc == 0) {
                    // das_add_equality moved x by t*z: replay the row-value update (t = last multiplier)
                    pb.on_step(dw.mu[dw.q - 1]);
                    if (lane == 0) dw.state[c] = -1;
                } else if (rc < 0) st |= ISMPC_ST_QP_FAIL;
                __syncwarp();
            }
        }
        dw.neq = dw.q;
        int rc = das_solve(pb, dw, xp, pb.rv, w + l.z, 20 * (nV + nC) + 50, &it);
        if (rc != 0) st |= ISMPC_ST_QP_FAIL;
        if (y) for (int i = lane; i < nC; i += 32) y[(size_t)p * nC + i] = 0.0;
        if (ws) for (int i = lane; i < nC; i += 32) ws[(size_t)p * nC + i] = dw.state[i];
        __syncwarp();
        if (y) for (int k = lane; k < dw.q; k += 32) y[(size_t)p * nC + dw.wid[k]] = (double)dw.wsg[k] * dw.mu[k];
    }
    if (lane == 0) { status[p] = st; if (iters) iters[p] = it; }
}

__global__ void dense_hinv(int N, double* work, size_t stride, size_t offLinv, size_t offHinv)
{
    const double* Linv = work + (size_t)blockIdx.z * stride + offLinv;
    double* Hinv = work + (size_t)blockIdx.z * stride + offHinv;
    const int i = blockIdx.y * blockDim.y + threadIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || j >= N) return;
    const int k0 = i > j ? i : j;
    double s = 0.0;
    for (int k = k0; k < N; ++k) s += Linv[(size_t)k * N + i] * Linv[(size_t)k * N + j];
    Hinv[(size_t)i * N + j] = s;
}

void dense_hinv_launch(int n, int nV, double* work, size_t stride, size_t offLinv, size_t offHinv, cudaStream_t st)
{
    dim3 b(16, 16), gr((nV + 15) / 16, (nV + 15) / 16, n);
    dense_hinv<<<gr, b, 0, st>>>(nV, work, stride, offLinv, offHinv);
}

int qp_dense_launch(int n, int nV, int nC, const double* H, const double* g, const double* A, const double* lbA,
                    const double* ubA, double* x, double* y, signed char* ws, int32_t* status, int32_t* iters,
                    double* work, int use_dmma, cudaStream_t st)
{
    const DenseLayout l = dense_layout(nV, nC);
    const size_t stride = l.total;
    const size_t vv = (size_t)nV * nV;
    dense_copy_H<<<dim3((unsigned)((vv + 255) / 256), n), 256, 0, st>>>(nV, H, work, stride, l.Lh);
    dense_cholesky<<<n, 256, 0, st>>>(nV, work, stride, l.Lh, status);
    dense_tri_inverse<<<dim3((nV + 63) / 64, n), 64, 0, st>>>(nV, work, stride, l.Lh, l.Linv);
    // Hinv = Linv' Linv: element (i, k) of the left operand is Linv[k][i] (strides 1, nV); the triangular zeros are
    // multiplied through (2x the flops of the triangular sum, at several times its speed)
    dense_gemm_launch(use_dmma, n, nV, nV, nV, work + l.Linv, stride, 1, nV, work + l.Linv, stride, 1, nV, work + l.Hinv, stride, st);
    if (nC > 0) {
        // D (nC x nV) = A (nC x nV) * Hinv' (Hinv symmetric, row-major over k)
        dense_gemm_launch(use_dmma, n, nC, nV, nV, A, (size_t)nC * nV, nV, 1, work + l.Hinv, stride, nV, 1, work + l.D, stride, st);
        // S (nC x nC) = A * D'
        dense_gemm_launch(use_dmma, n, nC, nC, nV, A, (size_t)nC * nV, nV, 1, work + l.D, stride, nV, 1, work + l.S, stride, st);
    }
    dense_x0_rv<<<n, 128, 0, st>>>(nV, nC, g, A, work, stride, l, x);
    const int R = l.qmax < DENSE_R ? l.qmax : DENSE_R;
    const size_t sbytes = (size_t)tri(R, 0) * sizeof(double) * DENSE_WARPS;
    cudaFuncSetAttribute(dense_das_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbytes);
    dense_das_kernel<<<(n + DENSE_WARPS - 1) / DENSE_WARPS, 32 * DENSE_WARPS, sbytes, st>>>(
        n, nV, nC, lbA, ubA, work, stride, l, R, x, y, ws, status, iters);
    return (int)cudaGetLastError();
}

}  // namespace ismpc
