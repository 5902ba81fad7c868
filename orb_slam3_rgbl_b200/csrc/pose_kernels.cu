// Optimizer::PoseOptimization (src/Optimizer.cc:814-1114) as ONE persistent CTA per frame: the whole
// 4 x (<= 10 Levenberg-Marquardt iterations x <= 10 trials) schedule of g2o runs on the device in FP64
// with no host round trip.  Threads stride over the edges; J^T W J (21 unique entries), J^T W r (6) and
// the robust chi2 are reduced with warp shuffles + a fixed 8-way shared-memory tree (deterministic);
// thread 0 does the 6x6 pivoted LDL^T solve, the SE3 exponential update and the LM bookkeeping.
// g2o semantics reproduced (Thirdparty/g2o/g2o/: core/optimization_algorithm_levenberg.cpp:61-185,
// core/base_unary_edge.hpp:43-72, core/robust_kernel_impl.cpp:65-91, types/se3quat.h:104-110,214-254,280-285,
// types/types_six_dof_expmap.cpp:339-404; src/OptimizableTypes.cpp:49-63):
//   * update T <- exp(delta) * T, delta = (omega, upsilon); quaternion normalised with w >= 0;
//   * H += rho' J^T Omega J, b -= rho' J^T Omega e (no second-order Huber term); Huber dsqr is a float;
//   * stereo edge: float invz in cam_project, double Jacobian; mono edge: -projectJac * SE3deriv;
//   * lambda0 = 1e-5 * max diag(H) at iteration 0 of every round; rho test; nu doubling; <= 10 trials;
//     stop when (ini - cur) * 1e3 < ini three times in a row;
//   * every round restarts from the frame's initial pose; inlier edges are classified with the error of
//     the LAST evaluated trial (g2o does not recompute it), outlier edges are recomputed.
#include <cfloat>

#include "rgbl_device.cuh"
#include "rgbl_kernels.h"

namespace rgbl {

namespace {

struct Se3d { double qx, qy, qz, qw, tx, ty, tz; };

__device__ __forceinline__ void normalize_rotation(Se3d& T) {
    if (T.qw < 0) { T.qx *= -1; T.qy *= -1; T.qz *= -1; T.qw *= -1; }
    const double n = sqrt(T.qx * T.qx + T.qy * T.qy + T.qz * T.qz + T.qw * T.qw);
    T.qx /= n; T.qy /= n; T.qz /= n; T.qw /= n;
}

__device__ __forceinline__ void quat_rotate(const Se3d& q, const double v[3], double out[3]) {
    double uv[3] = {q.qy * v[2] - q.qz * v[1], q.qz * v[0] - q.qx * v[2], q.qx * v[1] - q.qy * v[0]};
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    const double c[3] = {q.qy * uv[2] - q.qz * uv[1], q.qz * uv[0] - q.qx * uv[2], q.qx * uv[1] - q.qy * uv[0]};
    out[0] = v[0] + q.qw * uv[0] + c[0]; out[1] = v[1] + q.qw * uv[1] + c[1]; out[2] = v[2] + q.qw * uv[2] + c[2];
}

__device__ __forceinline__ void se3_map(const Se3d& T, const double p[3], double out[3]) {
    quat_rotate(T, p, out);
    out[0] += T.tx; out[1] += T.ty; out[2] += T.tz;
}

template <int I>
__device__ __forceinline__ void quat_from_matrix_case(const double R[3][3], Se3d& q) {
    constexpr int J = (I + 1) % 3, K = (J + 1) % 3;
    double t = sqrt(R[I][I] - R[J][J] - R[K][K] + 1.0);
    double v[3];
    v[I] = 0.5 * t;
    t = 0.5 / t;
    q.qw = (R[K][J] - R[J][K]) * t;
    v[J] = (R[J][I] + R[I][J]) * t;
    v[K] = (R[K][I] + R[I][K]) * t;
    q.qx = v[0]; q.qy = v[1]; q.qz = v[2];
}

// Eigen quaternion-from-rotation-matrix; every array index is a compile-time constant (registers, no local memory)
__device__ __forceinline__ void quat_from_matrix(const double R[3][3], Se3d& q) {
    double t = R[0][0] + R[1][1] + R[2][2];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q.qw = 0.5 * t;
        t = 0.5 / t;
        q.qx = (R[2][1] - R[1][2]) * t; q.qy = (R[0][2] - R[2][0]) * t; q.qz = (R[1][0] - R[0][1]) * t;
    } else {
        int i = 0;
        if (R[1][1] > R[0][0]) i = 1;
        if (i == 0) { if (R[2][2] > R[0][0]) i = 2; } else { if (R[2][2] > R[1][1]) i = 2; }
        if (i == 0) quat_from_matrix_case<0>(R, q);
        else if (i == 1) quat_from_matrix_case<1>(R, q);
        else quat_from_matrix_case<2>(R, q);
    }
}

__device__ __noinline__ void se3_exp(const double* u, Se3d& T) {
    const double om[3] = {u[0], u[1], u[2]}, up[3] = {u[3], u[4], u[5]};
    const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    const double O[3][3] = {{0, -om[2], om[1]}, {om[2], 0, -om[0]}, {-om[1], om[0], 0}};
    double O2[3][3], R[3][3], V[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) O2[i][j] = O[i][0] * O[0][j] + O[i][1] * O[1][j] + O[i][2] * O[2][j];
    if (theta < 0.00001) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) { R[i][j] = (i == j ? 1.0 : 0.0) + O[i][j] + O2[i][j]; V[i][j] = R[i][j]; }
    } else {
        const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / (theta * theta * theta);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                R[i][j] = (i == j ? 1.0 : 0.0) + a * O[i][j] + b * O2[i][j];
                V[i][j] = (i == j ? 1.0 : 0.0) + b * O[i][j] + c * O2[i][j];
            }
    }
    quat_from_matrix(R, T);
    T.tx = V[0][0] * up[0] + V[0][1] * up[1] + V[0][2] * up[2];
    T.ty = V[1][0] * up[0] + V[1][1] * up[1] + V[1][2] * up[2];
    T.tz = V[2][0] * up[0] + V[2][1] * up[1] + V[2][2] * up[2];
    normalize_rotation(T);
}

__device__ __noinline__ void se3_mul(const Se3d& a, const Se3d& b, Se3d& r) {
    const double bt[3] = {b.tx, b.ty, b.tz};
    double rt[3];
    quat_rotate(a, bt, rt);
    r.tx = a.tx + rt[0]; r.ty = a.ty + rt[1]; r.tz = a.tz + rt[2];
    r.qw = a.qw * b.qw - a.qx * b.qx - a.qy * b.qy - a.qz * b.qz;
    r.qx = a.qw * b.qx + a.qx * b.qw + a.qy * b.qz - a.qz * b.qy;
    r.qy = a.qw * b.qy + a.qy * b.qw + a.qz * b.qx - a.qx * b.qz;
    r.qz = a.qw * b.qz + a.qz * b.qw + a.qx * b.qy - a.qy * b.qx;
    normalize_rotation(r);
}

// 6x6 LDL^T with diagonal pivoting (what Eigen::LDLT does), positive-semidefinite check.  Unrolled with branch-free
// (select-based) row/column swaps so that every array index is a compile-time constant and the matrix lives in
// registers; __noinline__ keeps one copy of it in the kernel (an earlier inlined, branchy version made ptxas emit
// 180k instructions and the single-CTA kernel became instruction-fetch bound).
__device__ __noinline__ bool solve6(const double* Hsym /*21 upper*/, double lambda, const double* b, double* x) {
    double A[6][6], y[6];
    int pivs[6];
    {
        int t = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) { A[i][j] = Hsym[t]; A[j][i] = Hsym[t]; ++t; }
#pragma unroll
        for (int i = 0; i < 6; ++i) { A[i][i] += lambda; y[i] = b[i]; }
    }
    bool positive = true, stop = false;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        int piv = k;
        double best = fabs(A[k][k]);
#pragma unroll
        for (int i = k + 1; i < 6; ++i) { const bool g = fabs(A[i][i]) > best; best = g ? fabs(A[i][i]) : best; piv = g ? i : piv; }
        piv = stop ? k : piv;
        pivs[k] = piv;
#pragma unroll
        for (int p = k + 1; p < 6; ++p) {
            const bool sw = (piv == p);
#pragma unroll
            for (int j = 0; j < 6; ++j) { const double a = A[k][j], c = A[p][j]; A[k][j] = sw ? c : a; A[p][j] = sw ? a : c; }
#pragma unroll
            for (int i = 0; i < 6; ++i) { const double a = A[i][k], c = A[i][p]; A[i][k] = sw ? c : a; A[i][p] = sw ? a : c; }
            { const double a = y[k], c = y[p]; y[k] = sw ? c : a; y[p] = sw ? a : c; }
        }
        const double d = A[k][k];
        positive = positive && (stop || !(d < 0));
        stop = stop || (d == 0);
        const double dd = stop ? 1.0 : d;            // after an exact zero pivot Eigen leaves the trailing block untouched
#pragma unroll
        for (int i = k + 1; i < 6; ++i) A[i][k] = stop ? A[i][k] : A[i][k] / dd;
#pragma unroll
        for (int i = k + 1; i < 6; ++i)
#pragma unroll
            for (int j = k + 1; j <= i; ++j) {
                const double nv = A[i][j] - A[i][k] * dd * A[j][k];
                A[i][j] = stop ? A[i][j] : nv;
                A[j][i] = A[i][j];
            }
    }
    if (!positive) return false;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) y[i] -= A[i][j] * y[j];
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = (A[i][i] != 0) ? y[i] / A[i][i] : 0.0;
#pragma unroll
    for (int i = 5; i >= 0; --i)
#pragma unroll
        for (int j = i + 1; j < 6; ++j) y[i] -= A[j][i] * y[j];
    // x = P^T y: undo the swaps in reverse order
#pragma unroll
    for (int k = 5; k >= 0; --k) {
#pragma unroll
        for (int p = k + 1; p < 6; ++p) {
            const bool sw = (pivs[k] == p);
            const double a = y[k], c = y[p];
            y[k] = sw ? c : a; y[p] = sw ? a : c;
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = y[i];
    return true;
}

__device__ __forceinline__ void huber(double e2, double delta, float dsqr, double& rho0, double& rho1) {
    if (e2 <= (double)dsqr) { rho0 = e2; rho1 = 1.0; }
    else { const double sq = sqrt(e2); rho0 = 2 * sq * delta - (double)dsqr; rho1 = delta / sq; }
}

constexpr int kAcc = 28;      // 21 H + 6 b + 1 chi
constexpr int kPoseThreads = 512, kPoseWarps = kPoseThreads / 32;

// block-wide sum of kAcc doubles (fixed tree: lane shuffles, then warps 0..7 in order); result broadcast in `out`
__device__ __forceinline__ void block_reduce(double* v, double* smem /* 8*kAcc */, double* out /* kAcc, shared */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kAcc; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) smem[warp * kAcc + k] = x;
    }
    __syncthreads();
    if (threadIdx.x < kAcc) {
        double s = 0;
        for (int w = 0; w < kPoseWarps; ++w) s += smem[w * kAcc + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

}  // namespace

__global__ void __launch_bounds__(kPoseThreads) pose_optimize_kernel(PoseProblemDev p, double* __restrict__ work, uint8_t* __restrict__ level,
                                                            uint8_t* __restrict__ outlier, float* __restrict__ pose_out,
                                                            int* __restrict__ n_inliers) {
    __shared__ double red[kPoseWarps * kAcc];
    __shared__ double acc[kAcc];
    __shared__ Se3d s_est, s_init, s_backup;
    __shared__ double s_x[6], s_lambda, s_ni, s_rho, s_current, s_ini, s_temp;
    __shared__ int s_nbad_lm, s_qmax, s_ok, s_ok2, s_continue;
    const int tid = threadIdx.x, n = p.n_dev ? *p.n_dev : p.n;
    const float* pose_in = p.pose_in_dev ? p.pose_in_dev : p.pose_in;

    if (n < 3) {                      // src/Optimizer.cc:996
        if (tid < 7) pose_out[tid] = pose_in[tid];
        if (tid == 0) *n_inliers = 0;
        return;
    }
    const float delta_mono_f = sqrtf(5.991f), delta_stereo_f = sqrtf(7.815f);           // float deltaMono = sqrt(5.991)
    const double dm = (double)(float)sqrt(5.991), ds = (double)(float)sqrt(7.815);
    (void)delta_mono_f; (void)delta_stereo_f;
    const float dsqr_m = (float)(dm * dm), dsqr_s = (float)(ds * ds);
    const double fx = p.fx, fy = p.fy, cx = p.cx, cy = p.cy, bf = p.bf;

    if (tid == 0) {
        Se3d T = {pose_in[0], pose_in[1], pose_in[2], pose_in[3], pose_in[4], pose_in[5], pose_in[6]};
        normalize_rotation(T);
        s_init = T;
    }
    for (int k = tid; k < n; k += kPoseThreads) { level[k] = 0; outlier[k] = 0; }
    __syncthreads();
    bool robust = true;
    int n_bad = 0;

    // evaluates errors of the active edges at s_est, stores them, returns this thread's partial robust chi2
    auto eval_errors = [&](double& chi_part) {
        const Se3d T = s_est;
        chi_part = 0;
        for (int k = tid; k < n; k += kPoseThreads) {
            if (level[k] != 0) continue;
            const double xw[3] = {p.xw[3 * k], p.xw[3 * k + 1], p.xw[3 * k + 2]};
            double pc[3];
            se3_map(T, xw, pc);
            const bool st = p.stereo[k] != 0;
            const double info = (double)p.inv_sigma2[k];
            double e0, e1, e2 = 0;
            if (st) {
                const float invz = (float)(1.0 / pc[2]);
                const double u = pc[0] * invz * fx + cx, v = pc[1] * invz * fy + cy;
                e0 = (double)p.obs[3 * k] - u; e1 = (double)p.obs[3 * k + 1] - v; e2 = (double)p.obs[3 * k + 2] - (u - bf * invz);
            } else {
                e0 = (double)p.obs[3 * k] - ((double)p.fx * pc[0] / pc[2] + (double)p.cx);
                e1 = (double)p.obs[3 * k + 1] - ((double)p.fy * pc[1] / pc[2] + (double)p.cy);
            }
            work[3 * (size_t)k] = e0; work[3 * (size_t)k + 1] = e1; work[3 * (size_t)k + 2] = e2;
            double chi = e0 * (info * e0) + e1 * (info * e1);
            if (st) chi += e2 * (info * e2);
            if (robust) { double r0, r1; huber(chi, st ? ds : dm, st ? dsqr_s : dsqr_m, r0, r1); chi = r0; }
            chi_part += chi;
        }
    };

    for (int it = 0; it < 4; ++it) {
        if (tid == 0) { s_est = s_init; s_ok = 1; }
        __syncthreads();
        for (int iter = 0; iter < 10; ++iter) {
            if (!s_ok) break;
            // ---- computeActiveErrors + activeRobustChi2 + buildSystem at the current estimate ----
            double v[kAcc];
#pragma unroll
            for (int k = 0; k < kAcc; ++k) v[k] = 0;
            {
                const Se3d T = s_est;
                for (int k = tid; k < n; k += kPoseThreads) {
                    if (level[k] != 0) continue;
                    const double xw[3] = {p.xw[3 * k], p.xw[3 * k + 1], p.xw[3 * k + 2]};
                    double pc[3];
                    se3_map(T, xw, pc);
                    const bool st = p.stereo[k] != 0;
                    const double info = (double)p.inv_sigma2[k];
                    const double x = pc[0], y = pc[1];
                    double e[3] = {0, 0, 0}, J[3][6];
                    if (st) {
                        const float invzf = (float)(1.0 / pc[2]);
                        const double u = pc[0] * invzf * fx + cx, vv = pc[1] * invzf * fy + cy;
                        e[0] = (double)p.obs[3 * k] - u; e[1] = (double)p.obs[3 * k + 1] - vv; e[2] = (double)p.obs[3 * k + 2] - (u - bf * invzf);
                        const double invz = 1.0 / pc[2], invz_2 = invz * invz;
                        J[0][0] = x * y * invz_2 * fx;  J[0][1] = -(1 + (x * x * invz_2)) * fx; J[0][2] = y * invz * fx;
                        J[0][3] = -invz * fx;           J[0][4] = 0;                           J[0][5] = x * invz_2 * fx;
                        J[1][0] = (1 + y * y * invz_2) * fy; J[1][1] = -x * y * invz_2 * fy;   J[1][2] = -x * invz * fy;
                        J[1][3] = 0;                    J[1][4] = -invz * fy;                  J[1][5] = y * invz_2 * fy;
                        J[2][0] = J[0][0] - bf * y * invz_2; J[2][1] = J[0][1] + bf * x * invz_2; J[2][2] = J[0][2];
                        J[2][3] = J[0][3];              J[2][4] = 0;                           J[2][5] = J[0][5] - bf * invz_2;
                    } else {
                        const double z = pc[2];
                        e[0] = (double)p.obs[3 * k] - ((double)p.fx * pc[0] / pc[2] + (double)p.cx);
                        e[1] = (double)p.obs[3 * k + 1] - ((double)p.fy * pc[1] / pc[2] + (double)p.cy);
                        const double pj[2][3] = {{(double)p.fx / z, 0.0, -(double)p.fx * x / (z * z)}, {0.0, (double)p.fy / z, -(double)p.fy * y / (z * z)}};
                        const double D[3][6] = {{0, z, -y, 1, 0, 0}, {-z, 0, x, 0, 1, 0}, {y, -x, 0, 0, 0, 1}};
#pragma unroll
                        for (int r = 0; r < 2; ++r)
#pragma unroll
                            for (int c = 0; c < 6; ++c) J[r][c] = (-pj[r][0]) * D[0][c] + (-pj[r][1]) * D[1][c] + (-pj[r][2]) * D[2][c];
#pragma unroll
                        for (int c = 0; c < 6; ++c) J[2][c] = 0;
                    }
                    work[3 * (size_t)k] = e[0]; work[3 * (size_t)k + 1] = e[1]; work[3 * (size_t)k + 2] = e[2];
                    // mono edges carry a zero third row (J[2][*] = 0, e[2] = 0): identical sums, static indices -> registers
                    double chi = 0;
#pragma unroll
                    for (int r = 0; r < 3; ++r) chi += e[r] * (info * e[r]);
                    double w = 1.0, r0 = chi;
                    if (robust) huber(chi, st ? ds : dm, st ? dsqr_s : dsqr_m, r0, w);
                    v[27] += r0;
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        double sb = 0;
#pragma unroll
                        for (int r = 0; r < 3; ++r) sb += J[r][i] * (info * e[r]);
                        v[21 + i] -= w * sb;
#pragma unroll
                        for (int j = i; j < 6; ++j) {
                            double h = 0;
#pragma unroll
                            for (int r = 0; r < 3; ++r) h += J[r][i] * (w * info) * J[r][j];
                            v[i * 6 - (i * (i - 1)) / 2 + (j - i)] += h;
                        }
                    }
                }
            }
            block_reduce(v, red, acc);
            if (tid == 0) {
                s_current = acc[27]; s_ini = acc[27]; s_temp = acc[27];
                if (iter == 0) {
                    double mx = 0; int t = 0;
                    for (int i = 0; i < 6; ++i) { mx = fmax(fabs(acc[t]), mx); t += 6 - i; }
                    s_lambda = 1e-5 * mx; s_ni = 2; s_nbad_lm = 0;
                }
                s_rho = 0; s_qmax = 0;
            }
            __syncthreads();
            // ---- LM trials ----
            for (;;) {
                if (tid == 0) {
                    s_backup = s_est;
                    double x[6] = {0, 0, 0, 0, 0, 0};
                    s_ok2 = solve6(acc, s_lambda, acc + 21, x) ? 1 : 0;
                    for (int i = 0; i < 6; ++i) s_x[i] = x[i];
                    Se3d E, Tn;
                    se3_exp(x, E);
                    se3_mul(E, s_est, Tn);
                    s_est = Tn;
                }
                __syncthreads();
                double part;
                eval_errors(part);
                __syncthreads();
                {
                    // reduce only the chi2 slot (slot 27) but reuse the fixed tree
                    const int lane = tid & 31, warp = tid >> 5;
                    double x = part;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
                    if (lane == 0) red[warp] = x;
                    __syncthreads();
                    if (tid == 0) {
                        double s = 0;
                        for (int w = 0; w < kPoseWarps; ++w) s += red[w];
                        double temp = s;
                        if (!s_ok2) temp = DBL_MAX;
                        double rho = s_current - temp;
                        double scale = 0;
                        for (int j = 0; j < 6; ++j) scale += s_x[j] * (s_lambda * s_x[j] + acc[21 + j]);
                        scale += 1e-3;
                        rho /= scale;
                        if (rho > 0 && isfinite(temp)) {
                            const double r21 = 2 * rho - 1;
                            double alpha = 1. - r21 * r21 * r21;
                            alpha = fmin(alpha, 2. / 3.);
                            const double sf = fmax(1. / 3., alpha);
                            s_lambda *= sf; s_ni = 2; s_current = temp;
                        } else {
                            s_lambda *= s_ni; s_ni *= 2; s_est = s_backup;
                        }
                        s_rho = rho;
                        s_qmax += 1;
                        s_continue = (rho < 0 && s_qmax < 10) ? 1 : 0;
                    }
                    __syncthreads();
                }
                if (!s_continue) break;
            }
            if (tid == 0) {
                int ok = 1;
                if (s_qmax == 10 || s_rho == 0) ok = 0;
                else {
                    if ((s_ini - s_current) * 1e3 < s_ini) s_nbad_lm += 1; else s_nbad_lm = 0;
                    if (s_nbad_lm >= 3) ok = 0;
                }
                s_ok = ok;
            }
            __syncthreads();
        }
        // ---- classification (src/Optimizer.cc:1014-1100) ----
        {
            const Se3d T = s_est;
            int bad = 0;
            for (int k = tid; k < n; k += kPoseThreads) {
                const bool st = p.stereo[k] != 0;
                const double info = (double)p.inv_sigma2[k];
                double e0, e1, e2 = 0;
                if (outlier[k]) {
                    const double xw[3] = {p.xw[3 * k], p.xw[3 * k + 1], p.xw[3 * k + 2]};
                    double pc[3];
                    se3_map(T, xw, pc);
                    if (st) {
                        const float invz = (float)(1.0 / pc[2]);
                        const double u = pc[0] * invz * fx + cx, v = pc[1] * invz * fy + cy;
                        e0 = (double)p.obs[3 * k] - u; e1 = (double)p.obs[3 * k + 1] - v; e2 = (double)p.obs[3 * k + 2] - (u - bf * invz);
                    } else {
                        e0 = (double)p.obs[3 * k] - ((double)p.fx * pc[0] / pc[2] + (double)p.cx);
                        e1 = (double)p.obs[3 * k + 1] - ((double)p.fy * pc[1] / pc[2] + (double)p.cy);
                    }
                    work[3 * (size_t)k] = e0; work[3 * (size_t)k + 1] = e1; work[3 * (size_t)k + 2] = e2;
                } else {
                    e0 = work[3 * (size_t)k]; e1 = work[3 * (size_t)k + 1]; e2 = work[3 * (size_t)k + 2];
                }
                double chi = e0 * (info * e0) + e1 * (info * e1);
                if (st) chi += e2 * (info * e2);
                const float chi2 = (float)chi;
                if (chi2 > (st ? 7.815f : 5.991f)) { outlier[k] = 1; level[k] = 1; ++bad; }
                else { outlier[k] = 0; level[k] = 0; }
            }
            __syncthreads();
            // block sum of `bad`
            __shared__ int s_bad;
            if (tid == 0) s_bad = 0;
            __syncthreads();
            if (bad) atomicAdd(&s_bad, bad);
            __syncthreads();
            n_bad = s_bad;
            __syncthreads();
        }
        if (it == 2) robust = false;
        if (n < 10) break;
    }
    if (tid == 0) {
        const Se3d T = s_est;
        pose_out[0] = (float)T.qx; pose_out[1] = (float)T.qy; pose_out[2] = (float)T.qz; pose_out[3] = (float)T.qw;
        pose_out[4] = (float)T.tx; pose_out[5] = (float)T.ty; pose_out[6] = (float)T.tz;
        *n_inliers = n - n_bad;
    }
}

void launch_pose_optimize(cudaStream_t st, const PoseProblemDev& p, double* work, uint8_t* level, uint8_t* outlier,
                          float* pose_out, int* n_inliers) {
    pose_optimize_kernel<<<1, kPoseThreads, 0, st>>>(p, work, level, outlier, pose_out, n_inliers);
}

}  // namespace rgbl
