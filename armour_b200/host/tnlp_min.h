// Minimal stand-in for the part of Ipopt's TNLP interface that the planner's NLP class implements
// (IpTNLP.hpp: Number, Index, IndexStyleEnum, SolverReturn, the virtual callbacks).  Ipopt is not installed in
// this image; with -DARMOUR_HAVE_IPOPT the real header is used instead and armtd_NLP plugs into
// IpoptApplication::OptimizeTNLP exactly like the reference's class (KPR/NLPclass.h:11, KPR/armour_main.cu:237-273).
#pragma once
#ifdef ARMOUR_HAVE_IPOPT
#include <IpTNLP.hpp>
#else
namespace Ipopt {
typedef double Number;
typedef int Index;
class IpoptData;
class IpoptCalculatedQuantities;
enum SolverReturn { SUCCESS, MAXITER_EXCEEDED, CPUTIME_EXCEEDED, STOP_AT_TINY_STEP, LOCAL_INFEASIBILITY, INTERNAL_ERROR };
class TNLP {
public:
    enum IndexStyleEnum { C_STYLE = 0, FORTRAN_STYLE = 1 };
    virtual ~TNLP() {}
    virtual bool get_nlp_info(Index& n, Index& m, Index& nnz_jac_g, Index& nnz_h_lag, IndexStyleEnum& index_style) = 0;
    virtual bool get_bounds_info(Index n, Number* x_l, Number* x_u, Index m, Number* g_l, Number* g_u) = 0;
    virtual bool get_starting_point(Index n, bool init_x, Number* x, bool init_z, Number* z_L, Number* z_U, Index m,
                                    bool init_lambda, Number* lambda) = 0;
    virtual bool eval_f(Index n, const Number* x, bool new_x, Number& obj_value) = 0;
    virtual bool eval_grad_f(Index n, const Number* x, bool new_x, Number* grad_f) = 0;
    virtual bool eval_g(Index n, const Number* x, bool new_x, Index m, Number* g) = 0;
    virtual bool eval_jac_g(Index n, const Number* x, bool new_x, Index m, Index nele_jac, Index* iRow, Index* jCol,
                            Number* values) = 0;
    virtual bool eval_h(Index n, const Number* x, bool new_x, Number obj_factor, Index m, const Number* lambda,
                        bool new_lambda, Index nele_hess, Index* iRow, Index* jCol, Number* values) {
        return false;
    }
    virtual void finalize_solution(SolverReturn status, Index n, const Number* x, const Number* z_L, const Number* z_U,
                                   Index m, const Number* g, const Number* lambda, Number obj_value,
                                   const IpoptData* ip_data, IpoptCalculatedQuantities* ip_cq) = 0;
};
}  // namespace Ipopt
#endif
