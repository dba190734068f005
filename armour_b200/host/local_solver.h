// Built-in local NLP solver that drives a TNLP through its callbacks when Ipopt is not linked.
//
// The reference hands armtd_NLP to IpoptApplication::OptimizeTNLP (KPR/armour_main.cu:237-278: tol 1e-4,
// max_wall_time, ma97, L-BFGS Hessian).  Ipopt and HSL are not in this image, so the CLI needs a caller of
// eval_f / eval_grad_f / eval_g / eval_jac_g of its own.  The planner's cost is 10 * sum_j (q_des_j - q_j(t_plan; k))^2
// with q_j affine in k_j (KPR/NLPclass.cu:207-268), i.e. an exactly spherical quadratic, so a trust-region SQP
// step with a scaled-identity Hessian and the constraints linearised at the iterate is the exact Newton/SQP
// step.  Each iteration solves   min 1/2 h |d|^2 + grad_f . d   s.t.  g_l <= g + J d <= g_u,  x_l <= x + d <= x_u,
// |d|_inf <= Delta   exactly, by a dual active-set method (active_set_qp.h) over the rows that can become active inside
// the trust region, then accepts / rejects on (violation, cost).  Deterministic; no claim of matching Ipopt's
// iterates — the contract is the TNLP one: finalize_solution() receives the best point found.
#pragma once
#include "tnlp_min.h"

struct LocalSolverOptions {
    double tol = 1e-4;          // IPOPT_OPTIMIZATION_TOLERANCE (KPR/Parameters.h:51): step-size stopping test
    double max_wall_time = 0.45;  // seconds
    int max_iter = 60;
    int qp_sweeps = 200;        // cap on the active-set iterations of one QP (a guard: the method is finite)
    int qp_update_budget = 16384;  // unused since the QP is solved exactly (kept for the layout of the options)
    double torque_tol = 1e-2, collision_tol = 1e-4;  // acceptance tolerances = the verdict's (KPR/Parameters.h:40-43)
};
struct LocalSolverStats {
    int iterations = 0, evals = 0, qp_iterations = 0;
    double final_violation = 0, final_cost = 0, seconds = 0;
};

Ipopt::SolverReturn local_solve(Ipopt::TNLP& nlp, const LocalSolverOptions& opt, LocalSolverStats* stats);

// The QP of one step on its own (tests): min 1/2 h |d|^2 + c . d  s.t.  A d <= b  with A row-major [nrows][n], n <= 16.
// Returns the number of active-set iterations, -1 for bad arguments.
int local_qp(int n, double h, const double* c, int nrows, const double* A, const double* b, double* d, int max_outer);
