// Numerical study (CPU, test tooling): how far does PoseOptimization move when the per-edge Jacobian / normal-equation products
// are formed in float32 (sums in float64), and when the 6x6 solve runs in float32 with iterative refinement?  The chi2 /
// error evaluation that decides LM acceptance and the outlier flags stays in float64 in every variant.
//   g++ -O2 -std=c++17 -ffp-contract=off -o /tmp/pose_study tests/tools/pose_precision_study.cpp && /tmp/pose_study [problems]
// Reference arithmetic = oracle/pose_oracle.cpp (restatement of g2o's LM, src/Optimizer.cc:814-1114).
#define ORC_POSE_STUDY 1
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../oracle/pose_oracle.cpp"

namespace {

int g_build_mode = 0;      // 0 double (reference), 1 float products + double sums, 2 float products + float sums
int g_solve_mode = 0;      // 0 reference pivoted LDLT in double, k > 0: unpivoted float LDLT + (k - 1) refinement steps

void study_build_system(LM& lm, double H[6][6], double b[6]) {
    if (g_build_mode == 0) { lm.build_system(H, b); return; }
    double Hd[6][6] = {}, bd[6] = {};
    float Hf[6][6] = {}, bf_[6] = {};
    for (const Edge& e : lm.edges) {
        if (e.level != 0) continue;
        // camera-frame point and errors come from the float64 evaluation (they exist already when the system is built)
        double p[3];
        se3_map(lm.est, e.xw, p);
        const float x = (float)p[0], y = (float)p[1], z = (float)p[2];
        const float fx = lm.cam.fx, fy = lm.cam.fy, bf = lm.cam.bf;
        const float invz = 1.0f / z, invz_2 = invz * invz;
        float J[3][6];
        J[0][0] = x * y * invz_2 * fx;  J[0][1] = -(1 + (x * x * invz_2)) * fx; J[0][2] = y * invz * fx;
        J[0][3] = -invz * fx;           J[0][4] = 0;                           J[0][5] = x * invz_2 * fx;
        J[1][0] = (1 + y * y * invz_2) * fy; J[1][1] = -x * y * invz_2 * fy;   J[1][2] = -x * invz * fy;
        J[1][3] = 0;                    J[1][4] = -invz * fy;                  J[1][5] = y * invz_2 * fy;
        J[2][0] = J[0][0] - bf * y * invz_2; J[2][1] = J[0][1] + bf * x * invz_2; J[2][2] = J[0][2];
        J[2][3] = J[0][3];              J[2][4] = 0;                           J[2][5] = J[0][5] - bf * invz_2;
        const int d = e.stereo ? 3 : 2;
        double w = 1.0;
        if (e.robust) { double rho[3]; huber(edge_chi2(e), e.stereo ? lm.delta_stereo : lm.delta_mono, e.stereo ? lm.dsqr_stereo : lm.dsqr_mono, rho); w = rho[1]; }
        const float wi = (float)(w * e.info);
        const float er[3] = {(float)e.err[0], (float)e.err[1], (float)e.err[2]};
        for (int i = 0; i < 6; ++i) {
            float s = 0;
            for (int r = 0; r < d; ++r) s += J[r][i] * er[r];
            const float bi = wi * s;
            if (g_build_mode == 1) bd[i] -= (double)bi; else bf_[i] -= bi;
            for (int j = 0; j < 6; ++j) {
                float h = 0;
                for (int r = 0; r < d; ++r) h += J[r][i] * wi * J[r][j];
                if (g_build_mode == 1) Hd[i][j] += (double)h; else Hf[i][j] += h;
            }
        }
    }
    for (int i = 0; i < 6; ++i) { b[i] = g_build_mode == 1 ? bd[i] : (double)bf_[i]; for (int j = 0; j < 6; ++j) H[i][j] = g_build_mode == 1 ? Hd[i][j] : (double)Hf[i][j]; }
}

bool study_solve6(const double H[6][6], const double b[6], double x[6]) {
    if (g_solve_mode == 0) return solve6(H, b, x);
    float L[6][6], D[6];
    for (int j = 0; j < 6; ++j) {                       // unpivoted LDL^T in float
        float d = (float)H[j][j];
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k] * D[k];
        if (!(d > 0)) return false;
        D[j] = d;
        for (int i = j + 1; i < 6; ++i) {
            float v = (float)H[i][j];
            for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k] * D[k];
            L[i][j] = v / d;
        }
    }
    auto solve_f = [&](const float r[6], float out[6]) {
        float y[6];
        for (int i = 0; i < 6; ++i) { y[i] = r[i]; for (int k = 0; k < i; ++k) y[i] -= L[i][k] * y[k]; }
        for (int i = 0; i < 6; ++i) y[i] /= D[i];
        for (int i = 5; i >= 0; --i) { for (int k = i + 1; k < 6; ++k) y[i] -= L[k][i] * y[k]; out[i] = y[i]; }
    };
    float r[6], dx[6];
    for (int i = 0; i < 6; ++i) { r[i] = (float)b[i]; x[i] = 0; }
    solve_f(r, dx);
    for (int i = 0; i < 6; ++i) x[i] = dx[i];
    for (int it = 1; it < g_solve_mode; ++it) {         // refinement: residual in double, correction in float
        for (int i = 0; i < 6; ++i) { double s = b[i]; for (int j = 0; j < 6; ++j) s -= H[i][j] * x[j]; r[i] = (float)s; }
        solve_f(r, dx);
        for (int i = 0; i < 6; ++i) x[i] += dx[i];
    }
    return true;
}

struct Problem { std::vector<float> xw, obs, info; std::vector<uint8_t> stereo; float pose0[7]; };

Problem make_problem(std::mt19937& g, int n) {
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::normal_distribution<float> N(0.f, 1.f);
    const float fx = 718.856f, fy = 718.856f, cx = 607.19f, cy = 185.21f, bf = 386.14f;
    Problem p; p.xw.resize(3 * n); p.obs.resize(3 * n); p.info.resize(n); p.stereo.resize(n);
    // true pose: small rotation + ~1 m forward motion (world -> camera)
    const float ax = 0.02f * N(g), ay = 0.03f * N(g), az = 0.01f * N(g);
    const float tt[3] = {0.05f * N(g), 0.03f * N(g), -0.8f - 0.4f * U(g)};
    const float th = std::sqrt(ax * ax + ay * ay + az * az) + 1e-12f, sh = std::sin(th / 2) / th;
    const float q[4] = {ax * sh, ay * sh, az * sh, std::cos(th / 2)};
    for (int i = 0; i < n; ++i) {
        const float z = 4.f + 60.f * U(g) * U(g), u = 1241 * U(g), v = 376 * U(g);
        const float pc[3] = {(u - cx) * z / fx, (v - cy) * z / fy, z};
        // world = R^T (pc - t)
        const float d[3] = {pc[0] - tt[0], pc[1] - tt[1], pc[2] - tt[2]};
        const float qc[4] = {-q[0], -q[1], -q[2], q[3]};
        const float uvx = qc[1] * d[2] - qc[2] * d[1], uvy = qc[2] * d[0] - qc[0] * d[2], uvz = qc[0] * d[1] - qc[1] * d[0];
        const float t2[3] = {2 * uvx, 2 * uvy, 2 * uvz};
        p.xw[3 * i] = d[0] + qc[3] * t2[0] + (qc[1] * t2[2] - qc[2] * t2[1]);
        p.xw[3 * i + 1] = d[1] + qc[3] * t2[1] + (qc[2] * t2[0] - qc[0] * t2[2]);
        p.xw[3 * i + 2] = d[2] + qc[3] * t2[2] + (qc[0] * t2[1] - qc[1] * t2[0]);
        const int level = (int)(8 * U(g) * U(g));
        const float s = std::pow(1.2f, (float)level);
        p.info[i] = 1.f / (s * s);
        const bool outl = U(g) < 0.08f;
        const float noise = outl ? 30.f : 0.7f * s;
        const float uo = u + noise * N(g), vo = v + noise * N(g);
        p.stereo[i] = U(g) < 0.7f;
        p.obs[3 * i] = uo; p.obs[3 * i + 1] = vo; p.obs[3 * i + 2] = p.stereo[i] ? uo - bf / z + 0.5f * s * N(g) : -1.f;
    }
    // initial pose: the truth perturbed like a constant-velocity prediction error
    const float e[3] = {0.004f * N(g), 0.004f * N(g), 0.002f * N(g)};
    p.pose0[0] = q[0] + e[0]; p.pose0[1] = q[1] + e[1]; p.pose0[2] = q[2] + e[2]; p.pose0[3] = q[3];
    p.pose0[4] = tt[0] + 0.05f * N(g); p.pose0[5] = tt[1] + 0.03f * N(g); p.pose0[6] = tt[2] + 0.1f * N(g);
    return p;
}

}  // namespace

int main(int argc, char** argv) {
    const int n_prob = argc > 1 ? atoi(argv[1]) : 300, n = 520;
    const float fx = 718.856f, fy = 718.856f, cx = 607.19f, cy = 185.21f, bf = 386.14f;
    struct Var { const char* name; int build, solve; } vars[] = {
        {"float products, double sums, double solve", 1, 0}, {"float products, float sums, double solve", 2, 0},
        {"double build, float LDLT (no refinement)", 0, 1},  {"double build, float LDLT + 1 refinement", 0, 2},
        {"double build, float LDLT + 2 refinements", 0, 3},  {"float products + float LDLT + 2 refinements", 1, 3}};
    for (const Var& v : vars) {
        std::mt19937 g(12345);
        double max_dq = 0, max_dt = 0; int flag_diff = 0, inl_diff = 0, exact = 0;
        for (int k = 0; k < n_prob; ++k) {
            const Problem p = make_problem(g, n);
            float ref[7], out[7]; std::vector<uint8_t> o_ref(n), o_var(n);
            g_build_mode = 0; g_solve_mode = 0;
            const int i_ref = orc_pose_optimize(p.pose0, n, p.xw.data(), p.obs.data(), p.info.data(), p.stereo.data(), fx, fy, cx, cy, bf, ref, o_ref.data());
            g_build_mode = v.build; g_solve_mode = v.solve;
            const int i_var = orc_pose_optimize(p.pose0, n, p.xw.data(), p.obs.data(), p.info.data(), p.stereo.data(), fx, fy, cx, cy, bf, out, o_var.data());
            bool same = true;
            for (int i = 0; i < 4; ++i) { max_dq = std::max(max_dq, (double)std::fabs(out[i] - ref[i])); same &= out[i] == ref[i]; }
            for (int i = 4; i < 7; ++i) { max_dt = std::max(max_dt, (double)std::fabs(out[i] - ref[i])); same &= out[i] == ref[i]; }
            exact += same;
            for (int i = 0; i < n; ++i) flag_diff += o_ref[i] != o_var[i];
            inl_diff += i_ref != i_var;
        }
        printf("%-48s max|dq| %.2e  max|dt| %.2e m  bit-identical float poses %d/%d  outlier flags differing %d (of %d)  inlier counts differing %d\n",
               v.name, max_dq, max_dt, exact, n_prob, flag_diff, n_prob * n, inl_diff);
    }
    return 0;
}
