// armtd_NLP — the planner's NLP behind the Ipopt TNLP interface, with every body that touches reach sets or
// constraints replaced by a call into the C ABI of libarmour_b200.so (include/armour_b200.h).
//
// Mirrors the reference class (KPR/NLPclass.h:11-184): same callback names, argument meaning and return
// values (always true, eval_h false -> limited-memory Hessian), same public members read by main() after
// the solve (solution, feasible, constraint_number, g_copy, link_sliced_center; KPR/NLPclass.h:138-151,
// KPR/armour_main.cu:289-371).  set_parameters takes the library context instead of the raw pointers to
// traj / kd / torque_radius / obstacles (KPR/NLPclass.h:173-181).
#pragma once
#include <vector>

#include "../../include/armour_b200.h"
#include "tnlp_min.h"

using namespace Ipopt;

class armtd_NLP : public TNLP {
public:
    armtd_NLP() {}
    ~armtd_NLP() override;

    bool set_parameters(const double* q_des_input, double t_plan_input, armour_ctx* ctx_input, int num_obstacles_input);

    bool get_nlp_info(Index& n, Index& m, Index& nnz_jac_g, Index& nnz_h_lag, IndexStyleEnum& index_style) override;
    bool get_bounds_info(Index n, Number* x_l, Number* x_u, Index m, Number* g_l, Number* g_u) override;
    bool get_starting_point(Index n, bool init_x, Number* x, bool init_z, Number* z_L, Number* z_U, Index m,
                            bool init_lambda, Number* lambda) override;
    bool eval_f(Index n, const Number* x, bool new_x, Number& obj_value) override;
    bool eval_grad_f(Index n, const Number* x, bool new_x, Number* grad_f) override;
    bool eval_g(Index n, const Number* x, bool new_x, Index m, Number* g) override;
    bool eval_jac_g(Index n, const Number* x, bool new_x, Index m, Index nele_jac, Index* iRow, Index* jCol,
                    Number* values) override;
    bool eval_h(Index n, const Number* x, bool new_x, Number obj_factor, Index m, const Number* lambda, bool new_lambda,
                Index nele_hess, Index* iRow, Index* jCol, Number* values) override;
    void finalize_solution(SolverReturn status, Index n, const Number* x, const Number* z_L, const Number* z_U, Index m,
                           const Number* g, const Number* lambda, Number obj_value, const IpoptData* ip_data,
                           IpoptCalculatedQuantities* ip_cq) override;

    // same names as the reference's public members
    double solution[ARMOUR_NF] = {0};
    bool feasible = false;
    int constraint_number = 0;
    Number* g_copy = nullptr;
    std::vector<double> link_sliced_center;  // [(t*NJ + l)*3 + e]  (reference: Eigen::Vector3d[T*NJ])
    int first_violation = -1;                // row index of the first violated constraint (-1: none)
    bool quiet = false;

private:
    armtd_NLP(const armtd_NLP&);
    armtd_NLP& operator=(const armtd_NLP&);
    double q_des[ARMOUR_NF] = {0};
    double t_plan = 0;
    armour_ctx* ctx = nullptr;
    int num_obstacles = 0;
};
