// armour_main — drop-in for the reference planner executable (KPR/armour_main.cu): same input file
// (buffer/armour.in written by uarmtd_planner.m:158-185), same five output files (KPR/armour_main.cu:4-9,
// 312-372), same exit behaviour (0 = ran, even if infeasible; -1 + "-1" in armour.out on a failed stage).
// Sections II.A-II.D of the reference's main() are ONE call into libarmour_b200.so; the NLP is the
// armtd_NLP twin of armtd_nlp.h; the optimiser is Ipopt when compiled with -DARMOUR_HAVE_IPOPT (not in this
// image) and the built-in local solver otherwise.
//
// usage: armour_main [buffer_dir]      buffer_dir defaults to $ARMOUR_BUFFER_PATH or ./buffer/
//        armour_main --selftest        host-logic self test (parser / writers / local solver), no GPU needed
//        armour_main --serve           persistent server: one buffer directory per stdin line (see main)
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

#include "armtd_nlp.h"
#include "local_solver.h"

namespace {
constexpr int NF = ARMOUR_NF;
constexpr double DURATION = 1.0, IPOPT_TIME_BUFFER = 0.05;  // KPR/Parameters.h:14,52

struct PlannerInput {
    double q0[NF], qd0[NF], qdd0[NF], q_des[NF];
    int num_obstacles = 0;
    std::vector<double> obstacles;  // num_obstacles x 12: centre, g1, g2, g3
};

// KPR/armour_main.cu:46-78.  Returns 0, or -1 (unreadable) / -2 (obstacle count out of range)
int parse_input(const std::string& path, int max_obstacles, PlannerInput* in) {
    std::ifstream s(path);
    if (!s.is_open()) return -1;
    for (int i = 0; i < NF; i++) s >> in->q0[i];
    for (int i = 0; i < NF; i++) s >> in->qd0[i];
    for (int i = 0; i < NF; i++) s >> in->qdd0[i];
    for (int i = 0; i < NF; i++) s >> in->q_des[i];
    s >> in->num_obstacles;
    if (!s || in->num_obstacles > max_obstacles || in->num_obstacles < 0) return -2;
    in->obstacles.assign(size_t(in->num_obstacles) * 12, 0.0);
    for (double& v : in->obstacles) s >> v;
    return s.fail() ? -1 : 0;
}

void write_outputs(const std::string& dir, const armtd_NLP& nlp, int T, int NJ, const std::vector<double>& gens,
                   const std::vector<double>& torque_radius, long long total_ms) {  // KPR/armour_main.cu:312-372
    {
        std::ofstream o(dir + "armour.out");
        o << std::setprecision(10);
        if (nlp.feasible) {
            for (int i = 0; i < NF; i++) o << nlp.solution[i] << '\n';
        } else {
            o << -1 << '\n';
        }
        o << total_ms;
    }
    {
        std::ofstream o(dir + "armour_joint_position_center.out");
        o << std::setprecision(10);
        for (int i = 0; i < T; i++)
            for (int j = 0; j < NJ; j++) {
                for (int l = 0; l < 3; l++) o << nlp.link_sliced_center[(size_t(i) * NJ + j) * 3 + l] << ' ';
                o << '\n';
            }
    }
    {
        std::ofstream o(dir + "armour_joint_position_radius.out");
        o << std::setprecision(10);
        for (int i = 0; i < T; i++)
            for (int j = 0; j < NJ; j++)
                for (int k = 0; k < 3; k++) {
                    for (int l = 0; l < 6; l++) o << gens[(size_t(i) * NJ + j) * 18 + k + l * 3] << ' ';  // (k, l) of a column-major 3x6
                    o << '\n';
                }
    }
    {
        std::ofstream o(dir + "armour_control_input_radius.out");
        o << std::setprecision(10);
        for (int i = 0; i < T; i++) {
            for (int j = 0; j < NF; j++) o << torque_radius[size_t(j) * T + i] << ' ';
            o << '\n';
        }
    }
    {
        std::ofstream o(dir + "armour_constraints.out");
        o << std::setprecision(6);
        for (int i = 0; i < nlp.constraint_number; i++) o << nlp.g_copy[i] << '\n';
    }
}

// ---- host-logic self test: a toy TNLP with a known constrained optimum ---------------------------------
class ToyNLP : public TNLP {  // min 10*|x - c|^2  s.t.  x0 + x1 <= 0.5,  -1 <= x <= 1  (n = 7)
public:
    double c[NF] = {0.9, 0.8, -0.3, 0.2, 0, 0, 0.1};
    double sol[NF];
    bool get_nlp_info(Index& n, Index& m, Index& nj, Index& nh, IndexStyleEnum& st) override {
        n = NF; m = 2; nj = 2 * NF; nh = 0; st = C_STYLE;
        return true;
    }
    bool get_bounds_info(Index n, Number* xl, Number* xu, Index, Number* gl, Number* gu) override {
        for (int i = 0; i < n; i++) { xl[i] = -1; xu[i] = 1; }
        gl[0] = -1e19; gu[0] = 0.5; gl[1] = -0.7; gu[1] = 0.7;
        return true;
    }
    bool get_starting_point(Index n, bool, Number* x, bool, Number*, Number*, Index, bool, Number*) override {
        for (int i = 0; i < n; i++) x[i] = 0;
        return true;
    }
    bool eval_f(Index n, const Number* x, bool, Number& f) override {
        f = 0;
        for (int i = 0; i < n; i++) f += 10 * (x[i] - c[i]) * (x[i] - c[i]);
        return true;
    }
    bool eval_grad_f(Index n, const Number* x, bool, Number* g) override {
        for (int i = 0; i < n; i++) g[i] = 20 * (x[i] - c[i]);
        return true;
    }
    bool eval_g(Index, const Number* x, bool, Index, Number* g) override {
        g[0] = x[0] + x[1];
        g[1] = x[2] * x[2] + x[3];
        return true;
    }
    bool eval_jac_g(Index n, const Number* x, bool, Index, Index, Index*, Index*, Number* v) override {
        if (!v) return true;
        for (int i = 0; i < 2 * n; i++) v[i] = 0;
        v[0] = 1; v[1] = 1; v[n + 2] = 2 * x[2]; v[n + 3] = 1;
        return true;
    }
    void finalize_solution(SolverReturn, Index n, const Number* x, const Number*, const Number*, Index, const Number*,
                           const Number*, Number, const IpoptData*, IpoptCalculatedQuantities*) override {
        for (int i = 0; i < n; i++) sol[i] = x[i];
    }
};

int selftest() {
    int bad = 0;
    // 1. local solver on the toy problem: optimum x0 = 0.3, x1 = 0.2 (projection of (0.9, 0.8) on x0 + x1 = 0.5)
    ToyNLP toy;
    LocalSolverOptions opt;
    opt.max_wall_time = 5;
    LocalSolverStats st;
    local_solve(toy, opt, &st);
    const double expect[NF] = {0.3, 0.2, -0.3, 0.2, 0, 0, 0.1};
    for (int i = 0; i < NF; i++)
        if (std::fabs(toy.sol[i] - expect[i]) > 2e-3) {
            std::printf("selftest: solver x[%d] = %.6f, expected %.6f\n", i, toy.sol[i], expect[i]);
            bad++;
        }
    // 2. armour.in parser round trip
    const char* tmp = std::getenv("TMPDIR") ? std::getenv("TMPDIR") : "/tmp";
    const std::string path = std::string(tmp) + "/armour_selftest.in";
    {
        std::ofstream o(path);
        o << std::fixed << std::setprecision(10);
        for (int r = 0; r < 4; r++) {
            for (int i = 0; i < NF; i++) o << (r + 1) * 0.1 + i * 0.01 << ' ';
            o << '\n';
        }
        o << 2 << '\n';
        for (int r = 0; r < 2; r++) {
            for (int i = 0; i < 12; i++) o << r + i * 0.5 << ' ';
            o << '\n';
        }
    }
    PlannerInput in;
    if (parse_input(path, 40, &in) != 0 || in.num_obstacles != 2 || std::fabs(in.q_des[6] - 0.46) > 1e-12 ||
        std::fabs(in.obstacles[23] - 6.5) > 1e-12) {
        std::printf("selftest: parser round trip failed\n");
        bad++;
    }
    if (parse_input(path, 1, &in) != -2) {
        std::printf("selftest: obstacle limit not enforced\n");
        bad++;
    }
    if (parse_input(path + ".missing", 40, &in) != -1) bad++;
    std::remove(path.c_str());
    std::printf("selftest: %s (solver iterations %d, evals %d)\n", bad ? "FAILED" : "ok", st.iterations, st.evals);
    return bad ? 1 : 0;
}



// One planning iteration on a live context: armour.in -> reach sets -> NLP -> the five output files of the reference
// (KPR/armour_main.cu:36-78, 86-372).  Returns the process exit code the reference would give (0 = ran, even if no
// feasible plan was found; -1 = error, with a single -1 in armour.out).
int plan_once(armour_ctx* ctx, const armour_config& cfg, std::string dir) {
    if (!dir.empty() && dir.back() != '/') dir += '/';
    std::ofstream out1(dir + "armour.out");  // declared first so that there is always a new output (reference :36)
    auto fail_early = [&](const char* msg) {
        std::fprintf(stderr, "        CUDA & C++: %s\n", msg);
        out1 << -1;
        out1.close();
        return -1;
    };
    PlannerInput in;
    const int prc = parse_input(dir + "armour.in", cfg.max_obstacles, &in);
    if (prc == -1) return fail_early("Error reading input files !");
    if (prc == -2) return fail_early("Number of obstacles larger than MAX_OBSTACLE_NUM !");
    if (!ctx) return fail_early("cannot create the CUDA context (a GPU is required; there is no CPU path)");
    const double t_plan = 0.5;  // reference :80

    armour_ctx_reserve(ctx, 1, in.num_obstacles);  // device buffers, like the Obstacles constructor before the timer (:86-88)

    const auto start1 = std::chrono::high_resolution_clock::now();
    int rc = armour_reachsets_build(ctx, in.q0, in.qd0, in.qdd0, in.obstacles.data(), in.num_obstacles);  // sections II.A-II.D
    if (rc != ARMOUR_OK) {
        std::fprintf(stderr, "        CUDA & C++: %s\n", armour_last_error(ctx));
        return fail_early("Error computing link PZs and nominal torque PZs!");
    }
    const int T = armour_num_time_steps(ctx), NJ = armour_num_joints(ctx);
    std::vector<double> torque_radius(size_t(NF) * T), gens(size_t(T) * NJ * 18);
    armour_get_torque_radius(ctx, torque_radius.data());
    armour_get_link_independent_generators(ctx, gens.data());
    const auto stop1 = std::chrono::high_resolution_clock::now();
    const auto ms1 = std::chrono::duration_cast<std::chrono::milliseconds>(stop1 - start1).count();
    const double us1 = std::chrono::duration<double, std::micro>(stop1 - start1).count();
    std::cout << "        CUDA & C++: Time taken by generating reachable sets: " << ms1 << " milliseconds (" << us1 << " us)\n";
    double time_for_optimization = std::max(DURATION * 0.5 - ms1 / 1000.0 - IPOPT_TIME_BUFFER, 0.0);  // reference :227-229
    std::cout << "        CUDA & C++: Time allocated for the optimiser: " << time_for_optimization * 1000.0 << " milliseconds\n";

    const auto start2 = std::chrono::high_resolution_clock::now();
    const long long launches0 = armour_kernel_launches(ctx);
    armtd_NLP nlp;
    if (!nlp.set_parameters(in.q_des, t_plan, ctx, in.num_obstacles)) return fail_early("Error initializing the NLP!");
    LocalSolverOptions opt;
    opt.tol = 1e-4;  // IPOPT_OPTIMIZATION_TOLERANCE
    opt.max_wall_time = time_for_optimization;
    LocalSolverStats st;
    const SolverReturn status = local_solve(nlp, opt, &st);
    const auto stop2 = std::chrono::high_resolution_clock::now();
    const auto ms2 = std::chrono::duration_cast<std::chrono::milliseconds>(stop2 - start2).count();
    if (status == CPUTIME_EXCEEDED) std::cout << "        CUDA & C++: optimiser wall time exceeded!\n";
    std::cout << (nlp.feasible ? "        CUDA & C++: Found a feasible solution!\n" : "        CUDA & C++: Did not find a feasible solution!\n");
    std::cout << "        CUDA & C++: Time taken by the optimiser: " << ms2 << " milliseconds (" << st.iterations
              << " iterations, " << st.evals << " constraint evaluations, " << (armour_kernel_launches(ctx) - launches0)
              << " kernel launches)\n";

    out1.close();
    write_outputs(dir, nlp, T, NJ, gens, torque_radius, ms1 + ms2);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc > 1 && std::strcmp(argv[1], "--selftest") == 0) return selftest();
    const bool serve = argc > 1 && std::strcmp(argv[1], "--serve") == 0;
    const char* env = std::getenv("ARMOUR_BUFFER_PATH");
    const std::string default_dir = env ? env : "buffer/";

    armour_config cfg;
    armour_config_default(&cfg);  // NUM_TIME_STEPS 128, SIMPLIFY_THRESHOLD 5e-4, k_range pi/48, MAX_OBSTACLE_NUM 40
    armour_ctx* ctx = nullptr;
    if (armour_ctx_create(&cfg, &ctx) != ARMOUR_OK) ctx = nullptr;  // reported per request, with -1 in armour.out

    if (!serve) {
        const int rc = plan_once(ctx, cfg, argc > 1 ? argv[1] : default_dir);
        if (ctx) armour_ctx_destroy(ctx);
        return rc;
    }
    // Server mode: the process - CUDA context, device buffers, loaded kernels - stays alive between
    // replans; the reference pays process start + context + 10 cudaMallocs on every planning iteration
    // (uarmtd_planner.m:187-208 runs the executable once per replan).  Protocol on stdin / stdout, one request per
    // line: a buffer directory (empty line = the default directory) -> the five output files are written there and
    // the line "done <exit code>" is printed; "quit" or end of input ends the server.
    if (!ctx) {
        std::fprintf(stderr, "        CUDA & C++: cannot create the CUDA context (a GPU is required; there is no CPU path)\n");
        return -1;
    }
    armour_ctx_reserve(ctx, 1, cfg.max_obstacles);
    std::cout << "ready" << std::endl;
    std::string line;
    while (std::getline(std::cin, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
        if (line == "quit") break;
        const int rc = plan_once(ctx, cfg, line.empty() ? default_dir : line);
        std::cout << "done " << rc << std::endl;
    }
    armour_ctx_destroy(ctx);
    return 0;
}
