// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned: bit-identical to the reference's own sources (oracle/_ref, tests/test_oracle_pinned.py, tests/golden/ref).
// See pz.h.  Line references are to the reference file KPR/PZsparse.cu.
#include "pz.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>

namespace orc {

Stats& tls_stats() {
    static thread_local Stats s;
    return s;
}
std::vector<uint32_t>*& tls_trace() {
    static thread_local std::vector<uint32_t>* t = nullptr;
    return t;
}
static thread_local uint32_t g_pending_mul[3] = {0, 0, 0};
double& tls_threshold() {
    static thread_local double thr = 5e-4;
    return thr;
}
void Stats::add(const Stats& o) {
    n_simplify += o.n_simplify;
    n_mul += o.n_mul;
    n_pairs += o.n_pairs;
    flops += o.flops;
    terms_sorted += o.terms_sorted;
    max_terms = std::max(max_terms, o.max_terms);
    max_monos = std::max(max_monos, o.max_monos);
    near_threshold += o.near_threshold;
}

void hash_to_degree(uint64_t h, int deg[NVAR]) {  // :578-585
    for (int i = 0; i < NVAR; i++) {
        const int b = var_bits(i);
        deg[i] = int(h & ((uint64_t(1) << b) - 1));
        h >>= b;
    }
}

// ---- small dense helpers (column-major, Eigen-like evaluation order) -------------------
static inline double frob_norm(const double* c, int n) {
    double s = 0;
    for (int i = 0; i < n; i++) s += c[i] * c[i];
    return std::sqrt(s);
}
// C(r x p) = A(r x c) * B(c x p); inner index summed in increasing order
static inline void matmul(const double* A, int r, int c, const double* B, int p, double* C) {
    for (int j = 0; j < p; j++)
        for (int i = 0; i < r; i++) {
            double acc = A[i] * B[j * c];
            for (int k = 1; k < c; k++) acc += A[i + k * r] * B[k + j * c];
            C[i + j * r] = acc;
        }
}

// ---- constructors ----------------------------------------------------------------------
PZ PZ::scalar(double c) {
    PZ p(1, 1);
    p.center[0] = c;
    return p;
}
PZ PZ::matrix(int r, int c, const double* cm) {
    PZ p(r, c);
    std::memcpy(p.center, cm, sizeof(double) * r * c);
    return p;
}
PZ PZ::matrix_uncertain(int r, int c, const double* cm, double pct) {  // :93-98
    PZ p = matrix(r, c, cm);
    for (int i = 0; i < r * c; i++) p.indep[i] = pct * std::fabs(cm[i]);
    return p;
}
PZ PZ::scalar_poly(double c, const double* coeff, const uint64_t* hash, int n) {  // :120-136
    PZ p(1, 1);
    p.center[0] = c;
    p.poly.reserve(n);
    for (int i = 0; i < n; i++) {
        Mono m;
        m.degree = hash[i];
        m.c[0] = coeff[i];
        p.poly.push_back(m);
    }
    p.simplify();
    return p;
}
PZ PZ::rpy(double roll, double pitch, double yaw) {  // :160-176
    PZ p(3, 3);
    using std::cos;
    using std::sin;
    p.at(0, 0) = cos(pitch) * cos(yaw);
    p.at(0, 1) = -cos(pitch) * sin(yaw);
    p.at(0, 2) = sin(pitch);
    p.at(1, 0) = cos(roll) * sin(yaw) + cos(yaw) * sin(pitch) * sin(roll);
    p.at(1, 1) = cos(roll) * cos(yaw) - sin(pitch) * sin(roll) * sin(yaw);
    p.at(1, 2) = -cos(pitch) * sin(roll);
    p.at(2, 0) = sin(roll) * sin(yaw) - cos(roll) * cos(yaw) * sin(pitch);
    p.at(2, 1) = cos(yaw) * sin(roll) + cos(roll) * sin(pitch) * sin(yaw);
    p.at(2, 2) = cos(pitch) * cos(roll);
    return p;
}
// makeRotationMatrix :211-250 (column-major 3x3)
static void make_rotation(double* R, double c, double s, int axis, bool from_zero) {
    for (int i = 0; i < 9; i++) R[i] = 0;
    if (!from_zero) R[0] = R[4] = R[8] = 1.0;
    const double ns = -1.0 * s;
    auto set = [&](int r, int col, double v) { R[r + col * 3] = v; };
    switch (axis) {
        case 0: return;
        case 1: set(1, 1, c); set(1, 2, ns); set(2, 1, s); set(2, 2, c); break;
        case 2: set(0, 0, c); set(0, 2, s); set(2, 0, ns); set(2, 2, c); break;
        case 3: set(0, 0, c); set(0, 1, ns); set(1, 0, s); set(1, 1, c); break;
        default: throw -1;
    }
}
PZ PZ::rotation(double cc, const double* ccoef, const uint64_t* chash, int cn, double sc, const double* scoef,
                const uint64_t* shash, int sn, int axis) {  // :179-205
    PZ p(3, 3);
    make_rotation(p.center, cc, sc, axis, false);
    p.poly.reserve(cn + sn);
    for (int i = 0; i < cn; i++) {
        Mono m;
        m.degree = chash[i];
        make_rotation(m.c, ccoef[i], 0, axis, true);
        p.poly.push_back(m);
    }
    for (int i = 0; i < sn; i++) {
        Mono m;
        m.degree = shash[i];
        make_rotation(m.c, 0, scoef[i], axis, true);
        p.poly.push_back(m);
    }
    p.simplify();
    return p;
}

// ---- simplify / reduce -----------------------------------------------------------------
void PZ::simplify() {  // :284-350
    Stats& st = tls_stats();
    const double thr = tls_threshold();
    const int sz = size();
    st.n_simplify++;
    const uint32_t pre_terms = uint32_t(poly.size());
    st.terms_sorted += poly.size();
    st.max_terms = std::max<uint32_t>(st.max_terms, uint32_t(poly.size()));

    std::sort(poly.begin(), poly.end(), [](const Mono& l, const Mono& r) { return l.degree < r.degree; });

    double reduce_amount[9] = {0};
    uint32_t n_unique = 0;
    std::vector<Mono> out;
    out.reserve(poly.size());
    size_t i = 0;
    while (i < poly.size()) {
        size_t j;
        const uint64_t d = poly[i].degree;
        for (j = i + 1; j < poly.size(); j++) {
            if (poly[j].degree != d) break;
            for (int e = 0; e < sz; e++) poly[i].c[e] += poly[j].c[e];
        }
        n_unique++;
        const double nrm = frob_norm(poly[i].c, sz);
        if (std::fabs(nrm - thr) <= 1e-12 * thr) st.near_threshold++;
        if (nrm <= thr) {
            for (int e = 0; e < sz; e++) reduce_amount[e] += std::fabs(poly[i].c[e]);
        } else {
            out.push_back(poly[i]);
        }
        i = j;
    }
    poly.swap(out);
    if (frob_norm(reduce_amount, sz) != 0) {
        for (int e = 0; e < sz; e++) indep[e] = indep[e] + reduce_amount[e];
    }
    st.max_monos = std::max<uint32_t>(st.max_monos, uint32_t(poly.size()));
    if (std::vector<uint32_t>* tr = tls_trace()) {
        tr->push_back(g_pending_mul[0]);
        tr->push_back(g_pending_mul[1]);
        tr->push_back(g_pending_mul[2]);
        tr->push_back(pre_terms);
        tr->push_back(uint32_t(poly.size()));
        tr->push_back(uint32_t(sz));
        tr->push_back(n_unique);
    }
    g_pending_mul[0] = g_pending_mul[1] = g_pending_mul[2] = 0;
}

void PZ::reduce() {  // :352-368
    std::vector<Mono> out;
    out.reserve(poly.size());
    const int sz = size();
    for (const Mono& m : poly) {
        if (m.degree < kMaxHashKOnly) {
            out.push_back(m);
        } else {
            for (int e = 0; e < sz; e++) indep[e] += std::fabs(m.c[e]);
        }
    }
    poly.swap(out);
}

void PZ::reduce_link_PZ(double G[18]) {  // :370-402, G is 3x6 column-major
    assert(nr == 3 && nc == 1);
    for (int i = 0; i < 18; i++) G[i] = 0;
    std::vector<Mono> out;
    out.reserve(poly.size());
    int j = 0;
    for (const Mono& m : poly) {
        if (m.degree < kMaxHashKOnly) {
            out.push_back(m);
        } else if (m.degree < kMaxHashKLinksOnly && (m.degree & kKMask) == 0) {
            assert(j < 3);
            for (int e = 0; e < 3; e++) G[e + j * 3] = m.c[e];
            j++;
        } else {
            for (int e = 0; e < 3; e++) indep[e] += std::fabs(m.c[e]);
        }
    }
    poly.swap(out);
    G[0 + 3 * 3] = indep[0];
    G[1 + 4 * 3] = indep[1];
    G[2 + 5 * 3] = indep[2];
}

// ---- slice -----------------------------------------------------------------------------
void PZ::slice(const double* k, double* lo, double* hi) const {  // :404-435
    const int sz = size();
    double c[9], r[9];
    for (int e = 0; e < sz; e++) {
        c[e] = center[e];
        r[e] = indep[e];
    }
    int deg[NVAR];
    for (const Mono& m : poly) {
        if (m.degree < kMaxHashKOnly) {
            hash_to_degree(m.degree, deg);
            double t[9];
            for (int e = 0; e < sz; e++) t[e] = m.c[e];
            for (int j = 0; j < NF; j++) {
                const double f = std::pow(k[j], double(deg[j]));
                for (int e = 0; e < sz; e++) t[e] *= f;
            }
            for (int e = 0; e < sz; e++) c[e] += t[e];
        } else {
            for (int e = 0; e < sz; e++) r[e] += std::fabs(m.c[e]);
        }
    }
    for (int e = 0; e < sz; e++) {
        lo[e] = c[e] - r[e];
        hi[e] = c[e] + r[e];
    }
}

void PZ::slice_gradient(const double* k, double* grad) const {  // :437-555; grad[v*size + e]
    const int sz = size();
    for (int i = 0; i < NF * sz; i++) grad[i] = 0;
    int deg[NVAR];
    for (const Mono& m : poly) {
        if (m.degree <= kMaxHashKOnly) {  // sic: '<=' in the gradient overloads
            hash_to_degree(m.degree, deg);
            for (int v = 0; v < NF; v++) {
                double t[9];
                for (int e = 0; e < sz; e++) t[e] = m.c[e];
                for (int j = 0; j < NF; j++) {
                    if (j == v) {
                        if (deg[j] == 0) {
                            for (int e = 0; e < sz; e++) t[e] = 0;
                        } else {
                            const double f = double(deg[j]) * std::pow(k[j], double(deg[j] - 1));
                            for (int e = 0; e < sz; e++) t[e] *= f;
                        }
                    } else {
                        const double f = std::pow(k[j], double(deg[j]));
                        for (int e = 0; e < sz; e++) t[e] *= f;
                    }
                }
                for (int e = 0; e < sz; e++) grad[v * sz + e] += t[e];
            }
        }
    }
}

void PZ::to_interval(Interval* out) const {  // :557-576
    const int sz = size();
    double r[9];
    for (int e = 0; e < sz; e++) r[e] = indep[e];
    for (const Mono& m : poly)
        for (int e = 0; e < sz; e++) r[e] += std::fabs(m.c[e]);
    for (int e = 0; e < sz; e++) out[e] = Interval(center[e] - r[e], center[e] + r[e]);
}

// ---- element access / transpose --------------------------------------------------------
PZ PZ::elem(int r, int c) const {  // :678-697 — keeps every monomial, no simplify
    PZ res(1, 1);
    const int idx = r + c * nr;
    res.center[0] = center[idx];
    res.poly.reserve(poly.size());
    for (const Mono& m : poly) {
        Mono o;
        o.degree = m.degree;
        o.c[0] = m.c[idx];
        res.poly.push_back(o);
    }
    res.indep[0] = indep[idx];
    return res;
}

PZ PZ::transpose() const {  // :1050-1066
    PZ res(nc, nr);
    auto tr = [&](const double* src, double* dst) {
        for (int r = 0; r < nr; r++)
            for (int c = 0; c < nc; c++) dst[c + r * nc] = src[r + c * nr];
    };
    tr(center, res.center);
    tr(indep, res.indep);
    res.poly.reserve(poly.size());
    for (const Mono& m : poly) {
        Mono o;
        o.degree = m.degree;
        tr(m.c, o.c);
        res.poly.push_back(o);
    }
    return res;
}

void PZ::add_one_dim(const PZ& a, int r, int c) {  // :1068-1085
    assert(a.nr == 1 && a.nc == 1);
    const int idx = r + c * nr;
    center[idx] += a.center[0];
    for (const Mono& m : a.poly) {
        Mono o;
        o.degree = m.degree;
        o.c[idx] = m.c[0];
        poly.push_back(o);
    }
    indep[idx] += a.indep[0];
    simplify();
}

// ---- arithmetic ------------------------------------------------------------------------
PZ operator+(const PZ& a, const PZ& b) {  // :743-764
    PZ res(a.nr, a.nc);
    const int sz = a.size();
    for (int e = 0; e < sz; e++) res.center[e] = a.center[e] + b.center[e];
    res.poly.reserve(a.poly.size() + b.poly.size());
    res.poly.insert(res.poly.end(), a.poly.begin(), a.poly.end());
    res.poly.insert(res.poly.end(), b.poly.begin(), b.poly.end());
    for (int e = 0; e < sz; e++) res.indep[e] = a.indep[e] + b.indep[e];
    res.simplify();
    return res;
}

PZ operator-(const PZ& a, const PZ& b) {  // :813-834
    PZ res(a.nr, a.nc);
    const int sz = a.size();
    for (int e = 0; e < sz; e++) res.center[e] = a.center[e] - b.center[e];
    res.poly.reserve(a.poly.size() + b.poly.size());
    res.poly.insert(res.poly.end(), a.poly.begin(), a.poly.end());
    for (const Mono& m : b.poly) {
        Mono o;
        o.degree = m.degree;
        for (int e = 0; e < sz; e++) o.c[e] = -m.c[e];
        res.poly.push_back(o);
    }
    for (int e = 0; e < sz; e++) res.indep[e] = a.indep[e] + b.indep[e];
    res.simplify();
    return res;
}

PZ scale(double s, const PZ& b) {  // :996-1030
    PZ res(b.nr, b.nc);
    const int sz = b.size();
    for (int e = 0; e < sz; e++) res.center[e] = b.center[e] * s;
    res.poly.reserve(b.poly.size());
    for (const Mono& m : b.poly) {
        Mono o;
        o.degree = m.degree;
        for (int e = 0; e < sz; e++) o.c[e] = s * m.c[e];
        res.poly.push_back(o);
    }
    for (int e = 0; e < sz; e++) res.indep[e] = b.indep[e] * std::fabs(s);
    return res;
}

PZ operator*(const PZ& L, const PZ& R) {  // :864-994
    const bool ls = (L.nr == 1 && L.nc == 1);
    const bool rs = (R.nr == 1 && R.nc == 1);
    PZ res;
    if (ls) {
        res.nr = R.nr;
        res.nc = R.nc;
    } else if (rs) {
        res.nr = L.nr;
        res.nc = L.nc;
    } else {
        assert(L.nc == R.nr);
        res.nr = L.nr;
        res.nc = R.nc;
    }
    const int sz = res.size();
    const int lsz = L.size(), rsz = R.size();
    // generic product of two coefficient blocks following the three shape cases
    auto prod = [&](const double* a, const double* b, double* out) {
        if (ls) {
            for (int e = 0; e < sz; e++) out[e] = a[0] * b[e];
        } else if (rs) {
            for (int e = 0; e < sz; e++) out[e] = a[e] * b[0];
        } else {
            matmul(a, L.nr, L.nc, b, R.nc, out);
        }
    };
    prod(L.center, R.center, res.center);

    const size_t n1 = L.poly.size(), n2 = R.poly.size();
    res.poly.reserve(n1 + n2 + n1 * n2);
    for (const Mono& m : L.poly) {  // polynomial * a.center
        Mono o;
        o.degree = m.degree;
        prod(m.c, R.center, o.c);
        res.poly.push_back(o);
    }
    for (const Mono& m : R.poly) {  // center * a.polynomial
        Mono o;
        o.degree = m.degree;
        prod(L.center, m.c, o.c);
        res.poly.push_back(o);
    }
    for (const Mono& m1 : L.poly)
        for (const Mono& m2 : R.poly) {
            Mono o;
            o.degree = m1.degree + m2.degree;  // no carry by construction (:938-940)
            prod(m1.c, m2.c, o.c);
            res.poly.push_back(o);
        }

    // radius: r = rL*rR + (|cL| + sum|gL|)*rR + rL*(|cR| + sum|gR|)        (:944-989)
    double absL[9], absR[9];
    for (int e = 0; e < lsz; e++) absL[e] = std::fabs(L.center[e]);
    for (const Mono& m : L.poly)
        for (int e = 0; e < lsz; e++) absL[e] += std::fabs(m.c[e]);
    for (int e = 0; e < rsz; e++) absR[e] = std::fabs(R.center[e]);
    for (const Mono& m : R.poly)
        for (int e = 0; e < rsz; e++) absR[e] += std::fabs(m.c[e]);
    double ra2[9], ra3[9], rr[9];
    prod(absL, R.indep, ra2);
    prod(L.indep, absR, ra3);
    prod(L.indep, R.indep, rr);
    for (int e = 0; e < sz; e++) res.indep[e] = rr[e] + (ra2[e] + ra3[e]);

    Stats& st = tls_stats();
    st.n_mul++;
    st.n_pairs += n1 * n2;
    const uint64_t blocks = n1 * n2 + n1 + n2 + 1;
    st.flops += (ls || rs) ? uint64_t(sz) * blocks : uint64_t(2) * L.nr * L.nc * R.nc * blocks;

    g_pending_mul[0] = 1;
    g_pending_mul[1] = uint32_t(n1);
    g_pending_mul[2] = uint32_t(n2);
    res.simplify();
    return res;
}

PZ stack3(const PZ& a0, const PZ& a1, const PZ& a2) {  // :1087-1116
    const PZ* a[3] = {&a0, &a1, &a2};
    PZ res(3, 1);
    for (int i = 0; i < 3; i++) res.center[i] = a[i]->center[0];
    res.poly.reserve(3 * a0.poly.size());
    for (int i = 0; i < 3; i++)
        for (const Mono& m : a[i]->poly) {
            Mono o;
            o.degree = m.degree;
            o.c[i] = m.c[0];
            res.poly.push_back(o);
        }
    for (int i = 0; i < 3; i++) res.indep[i] = a[i]->indep[0];
    res.simplify();
    return res;
}

PZ cross_mat_pz(const double a[3], const PZ& b) {  // :1118-1132
    const PZ b0 = b.elem(0, 0), b1 = b.elem(1, 0), b2 = b.elem(2, 0);
    PZ r0 = scale(a[1], b2) - scale(a[2], b1);
    PZ r1 = scale(a[2], b0) - scale(a[0], b2);
    PZ r2 = scale(a[0], b1) - scale(a[1], b0);
    return stack3(r0, r1, r2);
}

PZ cross_pz_pz(const PZ& a, const PZ& b) {  // :1134-1151
    const PZ a0 = a.elem(0, 0), a1 = a.elem(1, 0), a2 = a.elem(2, 0);
    const PZ b0 = b.elem(0, 0), b1 = b.elem(1, 0), b2 = b.elem(2, 0);
    PZ r0 = a1 * b2 - a2 * b1;
    PZ r1 = a2 * b0 - a0 * b2;
    PZ r2 = a0 * b1 - a1 * b0;
    return stack3(r0, r1, r2);
}

PZ cross_pz_mat(const PZ& a, const double b[3]) {  // :1153-1167
    const PZ a0 = a.elem(0, 0), a1 = a.elem(1, 0), a2 = a.elem(2, 0);
    PZ r0 = scale(b[2], a1) - scale(b[1], a2);
    PZ r1 = scale(b[0], a2) - scale(b[2], a0);
    PZ r2 = scale(b[1], a0) - scale(b[0], a1);
    return stack3(r0, r1, r2);
}

}  // namespace orc
