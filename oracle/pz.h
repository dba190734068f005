// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned: bit-identical to the reference's own sources (oracle/_ref, tests/test_oracle_pinned.py, tests/golden/ref).
//
// CPU restatement of the reference's sparse polynomial zonotope (KPR/PZsparse.h:50-183,
// KPR/PZsparse.cu).  A PZ is  centre + sum_i coeff_i * prod_j x_j^{d_ij} + [-indep, +indep]
// with matrix-valued coefficients of shape 1x1, 3x1 or 3x3, the 42 exponents of a monomial
// packed into one 63-bit integer (KPR/PZsparse.h:23-40), plain round-to-nearest doubles.
// Matrices are stored column-major like Eigen's MatrixXd so that linear indexing and the
// Frobenius-norm traversal follow the reference.
//
// Every operation mirrors the reference's sequence of element insertions and its calls to
// simplify() (std::sort with the same comparator => the same permutation of equal keys =>
// the same floating-point summation order).
#pragma once
#include <cstdint>
#include <vector>

#include "interval.h"
#include "robot_model.h"

namespace orc {

// bit width / offset of each of the 42 variables inside the degree hash (KPR/PZsparse.h:23-35)
inline int var_bits(int v) { return (v < NF || v >= 4 * NF) ? 2 : 1; }
inline int var_offset(int v) {
    int off = 0;
    for (int i = 0; i < v; i++) off += var_bits(i);
    return off;
}
inline uint64_t var_hash(int v) { return uint64_t(1) << var_offset(v); }  // degree 1 in variable v

constexpr uint64_t kMaxHashKOnly = uint64_t(1) << (2 * NF);      // KPR/PZsparse.h:37
constexpr uint64_t kMaxHashKLinksOnly = uint64_t(1) << (5 * NF);  // KPR/PZsparse.h:39
constexpr uint64_t kKMask = kMaxHashKOnly - 1;                    // KPR/PZsparse.h:40

struct Mono {
    uint64_t degree = 0;
    double c[9] = {0};
};

// work counters (F_build counts coefficient products only)
struct Stats {
    uint64_t n_simplify = 0;     // simplify() calls
    uint64_t n_mul = 0;          // PZ*PZ products
    uint64_t n_pairs = 0;        // monomial pair products inside PZ*PZ
    uint64_t flops = 0;          // 2*r*c*p*(n1*n2+n1+n2+1) per matrix product, size*(...) per scalar one
    uint64_t terms_sorted = 0;   // total elements passed to std::sort
    uint32_t max_terms = 0;      // largest pre-merge term list
    uint32_t max_monos = 0;      // largest post-merge monomial list
    uint64_t near_threshold = 0; // prune decisions with |norm - thr| <= 1e-12*thr
    void add(const Stats& o);
};
Stats& tls_stats();
// optional per-thread op trace (design statistics / debugging): records
// {kind, n1, n2, terms_before_merge, monomials_after, coeff_size, unique_keys} per simplify();
// kind 0 = other, 1 = product
std::vector<uint32_t>*& tls_trace();
double& tls_threshold();  // SIMPLIFY_THRESHOLD for the calling thread

struct PZ {
    int nr = 0, nc = 0;
    double center[9] = {0};
    double indep[9] = {0};
    std::vector<Mono> poly;

    PZ() = default;
    PZ(int r, int c) : nr(r), nc(c) {}
    static PZ scalar(double c);                                 // KPR/PZsparse.cu:66-72
    static PZ matrix(int r, int c, const double* colmajor);     // :75-80
    static PZ matrix_uncertain(int r, int c, const double* colmajor, double pct);  // :93-98
    // 1x1 with monomial list, simplified                        // :120-136
    static PZ scalar_poly(double c, const double* coeff, const uint64_t* hash, int n);
    static PZ rpy(double roll, double pitch, double yaw);        // :160-176
    // 3x3 rotation about `axis` from cos/sin polynomials        // :179-205
    static PZ rotation(double cc, const double* ccoef, const uint64_t* chash, int cn,
                       double sc, const double* scoef, const uint64_t* shash, int sn, int axis);

    int size() const { return nr * nc; }
    double& at(int r, int c) { return center[r + c * nr]; }

    void simplify();                                             // :284-350
    void reduce();                                               // :352-368
    void reduce_link_PZ(double out_3x6_colmajor[18]);            // :370-402
    void slice(const double* k, double* out_center, double* out_radius) const;   // :404-435
    void slice_gradient(const double* k, double* grad /* [NF][size] */) const;   // :437-555
    void to_interval(Interval* out) const;                       // :557-576

    PZ elem(int r, int c) const;                                 // operator()(r,c) :678-697
    PZ transpose() const;                                        // :1050-1066
    void add_one_dim(const PZ& a, int r, int c);                 // :1068-1085
};

PZ operator+(const PZ& a, const PZ& b);       // :743-764 (and += :794-811, same insertion order)
PZ operator-(const PZ& a, const PZ& b);       // :813-834
PZ operator*(const PZ& a, const PZ& b);       // :864-994
PZ scale(double s, const PZ& b);              // double*PZ and PZ*double, :996-1030 (no simplify)
PZ stack3(const PZ& a0, const PZ& a1, const PZ& a2);   // :1087-1116
PZ cross_mat_pz(const double a[3], const PZ& b);       // :1118-1132
PZ cross_pz_pz(const PZ& a, const PZ& b);              // :1134-1151
PZ cross_pz_mat(const PZ& a, const double b[3]);       // :1153-1167

void hash_to_degree(uint64_t h, int deg[NVAR]);        // :578-585

}  // namespace orc
