// See armtd_nlp.h.  Reference lines cited per member are in KPR/NLPclass.cu.
#include "armtd_nlp.h"

#include <cstdio>
#include <cstring>

#define WARNING_PRINT(...) std::fprintf(stderr, __VA_ARGS__)

armtd_NLP::~armtd_NLP() { delete[] g_copy; }

bool armtd_NLP::set_parameters(const double* q_des_input, double t_plan_input, armour_ctx* ctx_input,
                               int num_obstacles_input) {  // :30-60
    for (int i = 0; i < ARMOUR_NF; i++) q_des[i] = q_des_input[i];
    t_plan = t_plan_input;
    ctx = ctx_input;
    num_obstacles = num_obstacles_input;
    constraint_number = armour_num_constraints(ctx, num_obstacles);
    if (constraint_number <= 0) return false;
    delete[] g_copy;
    g_copy = new Number[constraint_number];
    link_sliced_center.assign(size_t(armour_num_time_steps(ctx)) * armour_num_joints(ctx) * 3, 0.0);
    return true;
}

bool armtd_NLP::get_nlp_info(Index& n, Index& m, Index& nnz_jac_g, Index& nnz_h_lag, IndexStyleEnum& index_style) {  // :62-84
    n = ARMOUR_NF;
    m = constraint_number;
    nnz_jac_g = m * n;  // dense
    nnz_h_lag = n * (n + 1) / 2;
    index_style = TNLP::C_STYLE;
    return true;
}

bool armtd_NLP::get_bounds_info(Index n, Number* x_l, Number* x_u, Index m, Number* g_l, Number* g_u) {  // :86-165
    if (n != ARMOUR_NF) WARNING_PRINT("*** Error wrong value of n in get_bounds_info!");
    if (m != constraint_number) WARNING_PRINT("*** Error wrong value of m in get_bounds_info!");
    for (Index i = 0; i < n; i++) {
        x_l[i] = -1.0;
        x_u[i] = 1.0;
    }
    return armour_get_bounds(ctx, g_l, g_u) == ARMOUR_OK;
}

bool armtd_NLP::get_starting_point(Index n, bool init_x, Number* x, bool init_z, Number*, Number*, Index, bool init_lambda,
                                   Number*) {  // :167-201
    if (init_x == false || init_z == true || init_lambda == true)
        WARNING_PRINT("*** Error wrong value of init in get_starting_point!");
    if (n != ARMOUR_NF) WARNING_PRINT("*** Error wrong value of n in get_starting_point!");
    for (Index i = 0; i < n; i++) x[i] = 0.0;
    return true;
}

bool armtd_NLP::eval_f(Index n, const Number* x, bool, Number& obj_value) {  // :207-236
    if (n != ARMOUR_NF) WARNING_PRINT("*** Error wrong value of n in eval_f!");
    return armour_cost(ctx, q_des, x, &obj_value, nullptr) == ARMOUR_OK;
}

bool armtd_NLP::eval_grad_f(Index n, const Number* x, bool, Number* grad_f) {  // :241-268
    if (n != ARMOUR_NF) WARNING_PRINT("*** Error wrong value of n in eval_grad_f!");
    return armour_cost(ctx, q_des, x, nullptr, grad_f) == ARMOUR_OK;
}

bool armtd_NLP::eval_g(Index n, const Number* x, bool, Index m, Number* g) {  // :272-324
    if (n != ARMOUR_NF) WARNING_PRINT("*** Error wrong value of n in eval_g!");
    if (m != constraint_number) WARNING_PRINT("*** Error wrong value of m in eval_g!");
    return armour_eval_g(ctx, x, g) == ARMOUR_OK;
}

bool armtd_NLP::eval_jac_g(Index n, const Number* x, bool, Index m, Index, Index* iRow, Index* jCol, Number* values) {  // :330-396
    if (n != ARMOUR_NF) WARNING_PRINT("*** Error wrong value of n in eval_jac_g!");
    if (m != constraint_number) WARNING_PRINT("*** Error wrong value of m in eval_jac_g!");
    if (values == nullptr) {  // structure of the dense Jacobian (:348-357)
        for (Index i = 0; i < m; i++)
            for (Index j = 0; j < n; j++) {
                iRow[i * n + j] = i;
                jCol[i * n + j] = j;
            }
        return true;
    }
    return armour_eval_jac_g(ctx, x, values) == ARMOUR_OK;
}

bool armtd_NLP::eval_h(Index, const Number*, bool, Number, Index, const Number*, bool, Index, Index*, Index*, Number*) {
    return false;  // :398-416, limited-memory Hessian approximation
}

void armtd_NLP::finalize_solution(SolverReturn, Index n, const Number* x, const Number*, const Number*, Index m,
                                  const Number* g, const Number*, Number obj_value, const IpoptData*,
                                  IpoptCalculatedQuantities*) {  // :422-538
    for (Index i = 0; i < n; i++) solution[i] = double(x[i]);
    if (!quiet) std::printf("        CUDA & C++: final cost function value: %g\n", obj_value / 10.0);
    std::memcpy(g_copy, g, size_t(m) * sizeof(Number));
    // the sliced link centres of THIS x, as the reference leaves them after its last eval_g (NLPclass.h:150)
    std::vector<double> g_tmp(static_cast<size_t>(m), 0.0);
    if (armour_eval_g(ctx, x, g_tmp.data()) == ARMOUR_OK) armour_get_link_sliced_center(ctx, link_sliced_center.data());
    int ok = 0, first = -1;
    armour_verdict(ctx, g_copy, &ok, &first);  // the tolerance checks of :449-537, same order
    feasible = ok != 0;
    first_violation = first;
    if (!feasible && !quiet) std::printf("        CUDA & C++: constraint row %d violated\n", first);
}
