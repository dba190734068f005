// ORACLE — TEST INFRASTRUCTURE ONLY.
// Stand-in for the few Ipopt declarations the reference's armtd_NLP (KPR/NLPclass.h, KPR/NLPclass.cu) needs to
// compile: the TNLP base class with its enums and the Index / Number typedefs (Ipopt is not installed in this
// image and is not part of /root/reference).  No solver: oracle/ref_driver_nlp.cu calls the armtd_NLP members
// (get_bounds_info, eval_f, eval_grad_f, eval_g, eval_jac_g, finalize_solution) directly, as Ipopt would.
#pragma once

namespace Ipopt {

typedef int Index;
typedef double Number;

enum SolverReturn { SUCCESS, MAXITER_EXCEEDED, CPUTIME_EXCEEDED, STOP_AT_TINY_STEP, STOP_AT_ACCEPTABLE_POINT,
                    LOCAL_INFEASIBILITY, USER_REQUESTED_STOP, FEASIBLE_POINT_FOUND, DIVERGING_ITERATES,
                    RESTORATION_FAILURE, ERROR_IN_STEP_COMPUTATION, INVALID_NUMBER_DETECTED, TOO_FEW_DEGREES_OF_FREEDOM,
                    INVALID_OPTION, OUT_OF_MEMORY, INTERNAL_ERROR, UNASSIGNED };

class IpoptData;
class IpoptCalculatedQuantities;

class TNLP {
public:
    enum IndexStyleEnum { C_STYLE = 0, FORTRAN_STYLE = 1 };
    TNLP() {}
    virtual ~TNLP() {}
};

}  // namespace Ipopt
