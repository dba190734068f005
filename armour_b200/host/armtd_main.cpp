// armtd_main — drop-in for the reference's ARMTD comparison planner executable (kinova_planner_realtime_armtd_comparison/
// armtd_main.cu, "KPA"): same input file (buffer/armtd.in: q0, qd0, q_des, then per joint the six offline-JRS arrays and k_range,
// then the obstacles; armtd_main.cu:54-103), same four output files (:4-8, 209-268), same exit behaviour (0 = ran, even if no
// feasible plan was found; a single -1 in armtd.out on a failed stage).  Section II of the reference's main() is ONE call into
// libarmour_b200.so (armour_armtd_build), the NLP below mirrors KPA's armtd_NLP member by member over the armour_armtd_* calls,
// the optimiser is the built-in local solver (Ipopt is not in this image; with it the class derives from Ipopt::TNLP unchanged).
//
// usage: armtd_main [buffer_dir]      buffer_dir defaults to $ARMTD_BUFFER_PATH or ./buffer/
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/armour_b200.h"
#include "local_solver.h"
#include "tnlp_min.h"

using namespace Ipopt;

namespace {
constexpr int NF = ARMOUR_NF, T = 100;   // NUM_FACTORS, NUM_TIME_STEPS (KPA/Parameters.h:17)

// KPA/NLPclass.h:11-176 — same callback names and public members
class armtd_NLP : public TNLP {
public:
    bool set_parameters(const double* q_des_input, armour_ctx* ctx_input) {
        for (int i = 0; i < NF; i++) q_des[i] = q_des_input[i];
        ctx = ctx_input;
        constraint_number = armour_armtd_num_constraints(ctx);  // NLPclass.cu:43-44
        NJ = armour_num_joints(ctx);
        g_copy.assign(constraint_number > 0 ? constraint_number : 0, 0.0);
        link_sliced_center.assign(size_t(T) * NJ * 3, 0.0);
        return constraint_number > 0;
    }
    bool get_nlp_info(Index& n, Index& m, Index& nnz_jac_g, Index& nnz_h_lag, IndexStyleEnum& index_style) override {
        n = NF;
        m = constraint_number;
        nnz_jac_g = m * n;
        nnz_h_lag = 0;
        index_style = TNLP::C_STYLE;
        return true;
    }
    bool get_bounds_info(Index n, Number* x_l, Number* x_u, Index, Number* g_l, Number* g_u) override {
        for (Index i = 0; i < n; i++) {
            x_l[i] = -1.0;
            x_u[i] = 1.0;
        }
        return armour_armtd_get_bounds(ctx, g_l, g_u) == ARMOUR_OK;
    }
    bool get_starting_point(Index n, bool, Number* x, bool, Number*, Number*, Index, bool, Number*) override {
        for (Index i = 0; i < n; i++) x[i] = 0.0;  // NLPclass.cu:166-172
        return true;
    }
    bool eval_f(Index, const Number* x, bool, Number& obj_value) override {
        return armour_armtd_cost(ctx, q_des, x, &obj_value, nullptr) == ARMOUR_OK;
    }
    bool eval_grad_f(Index, const Number* x, bool, Number* grad_f) override {
        return armour_armtd_cost(ctx, q_des, x, nullptr, grad_f) == ARMOUR_OK;
    }
    bool eval_g(Index, const Number* x, bool, Index, Number* g) override { return armour_armtd_eval(ctx, x, g, nullptr) == ARMOUR_OK; }
    bool eval_jac_g(Index n, const Number* x, bool, Index m, Index, Index* iRow, Index* jCol, Number* values) override {
        if (values == nullptr) {  // dense structure, NLPclass.cu:300-306
            for (Index i = 0; i < m; i++)
                for (Index j = 0; j < n; j++) {
                    iRow[i * n + j] = i;
                    jCol[i * n + j] = j;
                }
            return true;
        }
        return armour_armtd_eval(ctx, x, nullptr, values) == ARMOUR_OK;
    }
    bool eval_h(Index, const Number*, bool, Number, Index, const Number*, bool, Index, Index*, Index*, Number*) override { return false; }
    void finalize_solution(SolverReturn, Index n, const Number* x, const Number*, const Number*, Index m, const Number* g,
                           const Number*, Number obj_value, const IpoptData*, IpoptCalculatedQuantities*) override {
        for (Index i = 0; i < n; i++) solution[i] = x[i];
        std::cout << "        CUDA & C++: final cost function value: " << obj_value / 10.0 << std::endl;
        std::memcpy(g_copy.data(), g, size_t(m) * sizeof(Number));
        int ok = 0;
        armour_armtd_verdict(ctx, g, &ok, &first_violation);  // NLPclass.cu:395-455
        feasible = ok != 0;
        // the sliced link centres at the solution, as the reference's last eval_g leaves them (NLPclass.h:146)
        std::vector<double> gtmp(static_cast<size_t>(m), 0.0);
        armour_armtd_eval(ctx, solution, gtmp.data(), nullptr);
        armour_armtd_get_link_sliced_center(ctx, link_sliced_center.data());
    }

    double solution[NF] = {0};
    bool feasible = false;
    int constraint_number = 0, NJ = 0, first_violation = -1;
    std::vector<Number> g_copy;
    std::vector<double> link_sliced_center;

private:
    double q_des[NF] = {0};
    armour_ctx* ctx = nullptr;
};

struct Input {
    double q0[NF], qd0[NF], q_des[NF], k_range[NF];
    std::vector<double> jrs;        // [6][NF][T]
    int num_obstacles = 0;
    std::vector<double> obstacles;  // [num_obstacles][12]
};

// armtd_main.cu:54-103; returns 0, -1 (cannot read), -2 (too many obstacles)
int parse_input(const std::string& path, int max_obstacles, Input* in) {
    std::ifstream s(path);
    if (!s.is_open()) return -1;
    for (int i = 0; i < NF; i++) s >> in->q0[i];
    for (int i = 0; i < NF; i++) s >> in->qd0[i];
    for (int i = 0; i < NF; i++) s >> in->q_des[i];
    in->jrs.assign(size_t(6) * NF * T, 0.0);
    for (int i = 0; i < NF; i++) {
        for (int a = 0; a < 6; a++)  // c_cos, g_cos, r_cos, c_sin, g_sin, r_sin of joint i
            for (int j = 0; j < T; j++) s >> in->jrs[(size_t(a) * NF + i) * T + j];
        s >> in->k_range[i];
    }
    s >> in->num_obstacles;
    if (!s) return -1;
    if (in->num_obstacles > max_obstacles || in->num_obstacles < 0) return -2;
    in->obstacles.assign(size_t(in->num_obstacles) * 12, 0.0);
    for (double& v : in->obstacles) s >> v;
    return s ? 0 : -1;
}
}  // namespace

int main(int argc, char** argv) {
    std::string dir = argc > 1 ? argv[1] : (std::getenv("ARMTD_BUFFER_PATH") ? std::getenv("ARMTD_BUFFER_PATH") : "./buffer/");
    if (!dir.empty() && dir.back() != '/') dir += '/';
    std::ofstream out1(dir + "armtd.out");  // first, so that there is always a new output (armtd_main.cu:36)
    auto fail_early = [&](const char* msg) {
        std::fprintf(stderr, "        CUDA & C++: %s\n", msg);
        out1 << -1;
        out1.close();
        return -1;
    };
    armour_config cfg;
    armour_config_default(&cfg);
    Input in;
    const int prc = parse_input(dir + "armtd.in", cfg.max_obstacles, &in);
    if (prc == -1) return fail_early("Error reading input files !");
    if (prc == -2) return fail_early("Number of obstacles larger than MAX_OBSTACLE_NUM !");
    armour_ctx* ctx = nullptr;
    if (armour_armtd_ctx_create(&cfg, &ctx) != ARMOUR_OK) return fail_early("cannot create the CUDA context (a GPU is required; there is no CPU path)");

    const auto start1 = std::chrono::high_resolution_clock::now();
    if (armour_armtd_build(ctx, in.q0, in.qd0, in.jrs.data(), in.k_range, in.obstacles.data(), in.num_obstacles) != ARMOUR_OK) {
        std::fprintf(stderr, "        CUDA & C++: %s\n", armour_last_error(ctx));
        armour_ctx_destroy(ctx);
        return fail_early("Error computing link PZs! Check previous error message!");
    }
    const int NJ = armour_num_joints(ctx);
    std::vector<double> gens(size_t(T) * NJ * 18);
    armour_armtd_get_link_independent_generators(ctx, gens.data());
    const auto stop1 = std::chrono::high_resolution_clock::now();
    const auto ms1 = std::chrono::duration_cast<std::chrono::milliseconds>(stop1 - start1).count();
    std::cout << "        CUDA & C++: Time taken by generating trajectory & forward kinematics: " << ms1 << " milliseconds ("
              << std::chrono::duration<double, std::micro>(stop1 - start1).count() << " us)" << std::endl;

    const auto start2 = std::chrono::high_resolution_clock::now();
    armtd_NLP nlp;
    if (!nlp.set_parameters(in.q_des, ctx)) {
        armour_ctx_destroy(ctx);
        return fail_early("Error initializing the NLP!");
    }
    LocalSolverOptions opt;
    opt.tol = 1e-7;             // IPOPT_OPTIMIZATION_TOLERANCE, KPA/Parameters.h:43
    opt.max_wall_time = 0.4;    // IPOPT_MAX_WALL_TIME, :45
    LocalSolverStats st;
    const SolverReturn status = local_solve(nlp, opt, &st);
    const auto stop2 = std::chrono::high_resolution_clock::now();
    const auto ms2 = std::chrono::duration_cast<std::chrono::milliseconds>(stop2 - start2).count();
    if (status == CPUTIME_EXCEEDED) std::cout << "        CUDA & C++: optimiser wall time exceeded!\n";
    std::cout << "        CUDA & C++: Time taken by the optimiser: " << ms2 << " milliseconds (" << st.iterations << " iterations)\n";

    // section IV, armtd_main.cu:209-268
    out1 << std::setprecision(10);
    if (nlp.feasible) {
        for (int i = 0; i < NF; i++) out1 << nlp.solution[i] << '\n';
    } else {
        out1 << -1 << '\n';
    }
    out1 << ms1 + ms2;
    out1.close();
    std::ofstream out2(dir + "armtd_joint_position_center.out");
    out2 << std::setprecision(10);
    for (int i = 0; i < T; i++)
        for (int j = 0; j < NJ; j++) {
            for (int l = 0; l < 3; l++) out2 << nlp.link_sliced_center[(size_t(i) * NJ + j) * 3 + l] << ' ';
            out2 << '\n';
        }
    out2.close();
    std::ofstream out3(dir + "armtd_joint_position_radius.out");
    out3 << std::setprecision(10);
    for (int i = 0; i < T; i++)
        for (int j = 0; j < NJ; j++)
            for (int k = 0; k < 3; k++) {
                for (int l = 0; l < 6; l++) out3 << gens[(size_t(i) * NJ + j) * 18 + l * 3 + k] << ' ';
                out3 << '\n';
            }
    out3.close();
    std::ofstream out4(dir + "armtd_constraints.out");
    out4 << std::setprecision(6);
    for (int i = 0; i < nlp.constraint_number; i++) out4 << nlp.g_copy[i] << '\n';
    out4.close();
    armour_ctx_destroy(ctx);
    return 0;
}
