// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned against the reference's own sources compiled where they lie
// (oracle/Makefile.ref target `mex` -> oracle/_ref/libarmour_ref_controller.so; tests/test_controller_oracle.py):
// interval outputs bit for bit, the nominal (double) torque to 1e-12 — the stand-in Eigen fixes an order for the double
// products that real Eigen's SSE2 paths may not share (oracle/ref_shim_mex/eigen3/Eigen/Dense, header).
//
// CPU restatement of the robust controller's hot path:
//   Model::Model(file)              MEX/robot_models.cpp:20-156   -> load_model()
//   IntModel::IntModel(model, eps)  MEX/robot_models.cpp:175-237  -> inside load_model()
//   passRNEA / passRNEA_Int         MEX/rnea.cpp:6-94 / 96-187    -> newton_euler<S>()
//   Transform / Twist / ... algebra MEX/spatial.cpp, MEX/spatial_interval.cpp -> the small functions below
//   RobustController::update        MEX/robust_controller.cpp:67-181 (ARMOUR method) -> orcctl_update()
// Written over a scalar S that is double or orc::Interval (oracle/interval.h: Boost.Interval's outward rounding).  Matrix
// coefficients follow Eigen 3.3's evaluation order for fixed sizes: coefficient (i, j) of a product is the sum() of three
// products, unrolled by halves: p0 + (p1 + p2).
#include <array>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "interval.h"

namespace {
using orc::Interval;

template <class S> using Vec = std::array<S, 3>;
template <class S> using Mat = std::array<std::array<S, 3>, 3>;  // [row][col]

inline double scale_by(double x, double s) { return x * s; }
inline Interval scale_by(const Interval& x, double s) { return s * x; }  // boost: T * interval
inline double plus_double(double x, double d) { return x + d; }
inline Interval plus_double(const Interval& x, double d) { return x + d; }  // interval += T

template <class S> S three(const S& p0, const S& p1, const S& p2) { return p0 + (p1 + p2); }
template <class S> Mat<S> mul(const Mat<S>& A, const Mat<S>& B) {
    Mat<S> C;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i][j] = three<S>(A[i][0] * B[0][j], A[i][1] * B[1][j], A[i][2] * B[2][j]);
    return C;
}
template <class S> Vec<S> mul(const Mat<S>& A, const Vec<S>& b) {
    Vec<S> c;
    for (int i = 0; i < 3; i++) c[i] = three<S>(A[i][0] * b[0], A[i][1] * b[1], A[i][2] * b[2]);
    return c;
}
template <class S> Mat<S> tr(const Mat<S>& A) {
    Mat<S> T;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) T[i][j] = A[j][i];
    return T;
}
template <class S> Mat<S> neg(const Mat<S>& A) {
    Mat<S> T;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) T[i][j] = -A[i][j];
    return T;
}
template <class S> Vec<S> add(const Vec<S>& a, const Vec<S>& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
template <class S> Vec<S> sub(const Vec<S>& a, const Vec<S>& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
template <class S> Vec<S> cross(const Vec<S>& a, const Vec<S>& b) {
    return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
template <class S> S inner(const Vec<S>& a, const Vec<S>& b) { return three<S>(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }
template <class S> Mat<S> skew(const Vec<S>& w) {
    const S z(0.0);
    Mat<S> H;
    H[0] = {z, -w[2], w[1]};
    H[1] = {w[2], z, -w[0]};
    H[2] = {-w[1], w[0], z};
    return H;
}
template <class S> Mat<S> identity() {
    Mat<S> I;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) I[i][j] = S(i == j ? 1.0 : 0.0);
    return I;
}

template <class S> struct Twist {
    Vec<S> w, v;
};
template <class S> struct Wrench {
    Vec<S> tau, f;
};
template <class S> struct Transform {
    Mat<S> R;
    Vec<S> p;
};
template <class S> Twist<S> scaled(const Twist<S>& z, double s) {
    Twist<S> r;
    for (int k = 0; k < 3; k++) {
        r.w[k] = scale_by(z.w[k], s);
        r.v[k] = scale_by(z.v[k], s);
    }
    return r;
}
template <class S> Twist<S> add(const Twist<S>& a, const Twist<S>& b) { return {add(a.w, b.w), add(a.v, b.v)}; }
template <class S> Twist<S> cross(const Twist<S>& a, const Twist<S>& z2) {
    const Mat<S> H = skew(a.w);
    return {mul(H, z2.w), add(mul(H, z2.v), cross(a.v, z2.w))};
}
template <class S> Transform<S> compose(const Transform<S>& X, const Transform<S>& x2) {
    return {mul(X.R, x2.R), add(x2.p, mul(tr(x2.R), X.p))};
}
template <class S> Transform<S> inverse(const Transform<S>& X) { return {tr(X.R), mul(neg(X.R), X.p)}; }
template <class S> Twist<S> apply(const Transform<S>& X, const Twist<S>& z) {
    return {mul(X.R, z.w), mul(X.R, sub(z.v, cross(X.p, z.w)))};
}
template <class S> Twist<S> invapply(const Transform<S>& X, const Twist<S>& z) {
    const Vec<S> w = mul(tr(X.R), z.w);
    return {w, add(mul(tr(X.R), z.v), cross(X.p, w))};
}
template <class S> Wrench<S> invapply(const Transform<S>& X, const Wrench<S>& w) {
    const Vec<S> f = mul(tr(X.R), w.f);
    return {add(mul(tr(X.R), w.tau), cross(X.p, f)), f};
}
// Rodrigues' formula around the twist axis (spatial_interval.cpp:147-157)
template <class S> Transform<S> joint_transform(const Twist<S>& zeta, double theta) {
    const Mat<S> H = skew(zeta.w), I = identity<S>();
    const double s = std::sin(theta), c1 = 1 - std::cos(theta);
    Mat<S> Hs, Hc;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            Hs[i][j] = scale_by(H[i][j], s);
            Hc[i][j] = scale_by(H[i][j], c1);
        }
    const Mat<S> HcH = mul(Hc, H);
    Transform<S> X;
    Mat<S> ImR;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            X.R[i][j] = (I[i][j] + Hs[i][j]) + HcH[i][j];
            ImR[i][j] = I[i][j] - X.R[i][j];
        }
    X.p = mul(neg(tr(X.R)), mul(mul(ImR, H), zeta.v));
    return X;
}

template <class S> struct Joint {
    Twist<S> S_;
    Transform<S> X;
    S m;
    Mat<S> Ibar, mch;
    S transI;
};
struct Robot {
    int nj = 0;
    std::vector<int> parent;
    std::vector<Joint<double>> nom;
    std::vector<Joint<Interval>> iv;
    std::array<double, 3> gravity{};
    std::vector<double> friction, damping;
};
template <class S> const std::vector<Joint<S>>& joints_of(const Robot& R);
template <> const std::vector<Joint<double>>& joints_of<double>(const Robot& R) { return R.nom; }
template <> const std::vector<Joint<Interval>>& joints_of<Interval>(const Robot& R) { return R.iv; }

template <class S> Vec<S> lift(const Vec<double>& v) { return {S(v[0]), S(v[1]), S(v[2])}; }
template <class S> Mat<S> lift(const Mat<double>& M) {
    Mat<S> r;
    for (int i = 0; i < 3; i++) r[i] = lift<S>(M[i]);
    return r;
}

bool load_model(const char* path, double eps, Robot& R) {
    std::ifstream in(path);
    if (!in.is_open()) return false;
    struct Raw {
        Twist<double> S;
        Transform<double> X;
        double m = 0;
        Mat<double> Ibar{}, mch{};
        Vec<double> com{};
    };
    std::vector<Raw> raw;
    std::vector<double> transI;
    std::string row;
    while (std::getline(in, row)) {
        // the reference's scanner: letters and '_' make the field, digits before '<' the index, the text between '<' and
        // '>' the values (robot_models.cpp:43-69)
        std::string field, index, values;
        bool inside = false;
        for (char ch : row) {
            if (ch == '<') { inside = true; continue; }
            if (ch == '>') break;
            if (inside) values += ch;
            else if (std::isalpha(static_cast<unsigned char>(ch)) || ch == '_') field += ch;
            else if (std::isdigit(static_cast<unsigned char>(ch))) index += ch;
        }
        std::vector<double> v;
        std::stringstream ss(values);
        std::string tok;
        while (std::getline(ss, tok, ' '))
            if (!tok.empty()) v.push_back(std::stod(tok));
        const int ind = index.empty() ? -1 : std::stoi(index);
        if (field == "numJoints") {
            R.nj = int(v.at(0));
            raw.assign(R.nj, Raw());
            for (auto& r : raw) r.X = {identity<double>(), {0, 0, 0}};
            R.parent.assign(R.nj, -1);
            transI.assign(R.nj, 0.0);
            R.friction.assign(R.nj, 0.0);
            R.damping.assign(R.nj, 0.0);
        } else if (field == "twist") {
            raw.at(ind).S = {{v.at(0), v.at(1), v.at(2)}, {v.at(3), v.at(4), v.at(5)}};
        } else if (field == "gravity") {
            R.gravity = {v.at(0), v.at(1), v.at(2)};
        } else if (field == "inertia") {
            raw.at(ind).m = v.at(0);
            for (int k = 0; k < 9; k++) {
                raw[ind].Ibar[k / 3][k % 3] = v.at(1 + k);
                raw[ind].mch[k / 3][k % 3] = v.at(10 + k);
            }
        } else if (field == "Xtree") {
            for (int k = 0; k < 9; k++) raw.at(ind).X.R[k / 3][k % 3] = v.at(k);
            raw[ind].X.p = {v.at(9), v.at(10), v.at(11)};
        } else if (field == "parent") {
            for (int j = 0; j < R.nj; j++) R.parent[j] = int(v.at(j));
        } else if (field == "CoM") {
            raw.at(ind).com = {v.at(0), v.at(1), v.at(2)};
        } else if (field == "transI") {
            for (int j = 0; j < R.nj; j++) transI[j] = v.at(j);
        } else if (field == "friction") {
            for (int j = 0; j < R.nj; j++) R.friction[j] = v.at(j);
        } else if (field == "damping") {
            for (int j = 0; j < R.nj; j++) R.damping[j] = v.at(j);
        }
    }
    if (R.nj == 0) return false;
    R.nom.resize(R.nj);
    R.iv.resize(R.nj);
    const double lowP = 1 - eps, highP = 1 + eps;
    for (int i = 0; i < R.nj; i++) {
        // robot_models.cpp:135-153
        Transform<double> Xwj = raw[i].X;
        for (int p = R.parent[i]; p > -1; p = R.parent[p]) Xwj = compose(Xwj, raw[p].X);
        Joint<double>& N = R.nom[i];
        N.S_ = invapply(Xwj, raw[i].S);
        // CoM[i].apply(I[i]) (spatial.cpp:221-237) with the CoM transform (R = identity, p = com)
        const Transform<double> C{identity<double>(), raw[i].com};
        const Mat<double> ph = skew(C.p);
        Mat<double> mR, two_mch;
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) {
                mR[a][b] = raw[i].m * C.R[a][b];
                two_mch[a][b] = 2.0 * raw[i].mch[a][b];
            }
        const Mat<double> mRp = mul(mR, ph), Rt = tr(C.R);
        const Mat<double> t1 = mul(mul(C.R, raw[i].mch), Rt), t2 = mul(mul(mR, ph), Rt);
        const Mat<double> inner_sum_prod = mul(two_mch, ph);
        Mat<double> inner_sum, t3, t4 = mul(mRp, ph);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) inner_sum[a][b] = raw[i].Ibar[a][b] + inner_sum_prod[a][b];
        t3 = mul(C.R, inner_sum);
        Mat<double> diff;
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) {
                N.mch[a][b] = t1[a][b] - t2[a][b];
                diff[a][b] = t3[a][b] - t4[a][b];
            }
        N.Ibar = mul(diff, Rt);
        N.m = raw[i].m;
        Transform<double> prev{identity<double>(), {0, 0, 0}};
        if (R.parent[i] != -1) prev.p = raw[R.parent[i]].com;
        const Transform<double> next{identity<double>(), raw[i].com};
        N.X = compose(prev, compose(inverse(raw[i].X), inverse(next)));
        N.transI = transI[i];
        // IntModel (robot_models.cpp:188-232)
        Joint<Interval>& I = R.iv[i];
        I.S_ = {lift<Interval>(N.S_.w), lift<Interval>(N.S_.v)};
        I.X = {lift<Interval>(N.X.R), lift<Interval>(N.X.p)};
        I.mch = lift<Interval>(N.mch);
        I.transI = Interval(N.transI);
        I.m = Interval(N.m * lowP, N.m * highP);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) {
                const double val = N.Ibar[a][b];
                I.Ibar[a][b] = val >= 0 ? Interval(val * lowP, val * highP) : Interval(val * highP, val * lowP);
            }
    }
    return true;
}

template <class S>
void newton_euler(const Robot& R, const double* q, const double* qd, const double* qda, const double* qdd, bool friction,
                  bool gravity, S* tau) {
    const std::vector<Joint<S>>& J = joints_of<S>(R);
    const int n = R.nj;
    std::vector<Twist<S>> v(n), va(n), a(n), Sb(n);
    std::vector<Wrench<S>> f(n);
    std::vector<Transform<S>> Xbw(n), Xli(n);
    Twist<S> neg_g{{S(0.0), S(0.0), S(0.0)}, {S(0.0), S(0.0), S(0.0)}};
    if (gravity)
        for (int k = 0; k < 3; k++) neg_g.v[k] = -S(R.gravity[k]);
    for (int i = 0; i < n; i++) {
        const int li = R.parent[i];
        Xbw[i] = li != -1 ? compose(Xbw[li], J[i].X) : J[i].X;
        Sb[i] = invapply(Xbw[i], J[i].S_);
        Xli[i] = compose(joint_transform(Sb[i], -q[i]), inverse(J[i].X));
        const Twist<S> sa = scaled(Sb[i], qda[i]);
        if (li == -1) {
            v[i] = scaled(Sb[i], qd[i]);
            va[i] = sa;
            a[i] = add(add(apply(Xli[i], neg_g), scaled(Sb[i], qdd[i])), cross(v[i], va[i]));
        } else {
            v[i] = add(apply(Xli[i], v[li]), scaled(Sb[i], qd[i]));
            va[i] = add(apply(Xli[i], va[li]), sa);
            a[i] = add(add(apply(Xli[i], a[li]), scaled(Sb[i], qdd[i])), cross(v[i], sa));
        }
        Vec<S> vt = cross(va[i].w, mul(J[i].Ibar, v[i].w));
        vt = add(vt, mul(J[i].Ibar, cross(va[i].w, v[i].w)));
        const Vec<S> cf = cross(va[i].w, v[i].v);
        const Vec<S> vf = {J[i].m * cf[0], J[i].m * cf[1], J[i].m * cf[2]};
        const Vec<S> it = add(mul(J[i].Ibar, a[i].w), mul(J[i].mch, a[i].v));
        const Vec<S> mv = {J[i].m * a[i].v[0], J[i].m * a[i].v[1], J[i].m * a[i].v[2]};
        const Vec<S> itf = sub(mv, mul(J[i].mch, a[i].w));
        f[i] = {add(it, vt), add(itf, vf)};
    }
    for (int i = n - 1; i >= 0; i--) {
        S t = inner(Sb[i].w, f[i].tau) + inner(Sb[i].v, f[i].f);
        t = t + scale_by(J[i].transI, qdd[i]);
        t = plus_double(t, R.damping[i] * qd[i]);
        if (friction) t = plus_double(t, R.friction[i] * double((qd[i] > 0) - (qd[i] < 0)));
        tau[i] = t;
        if (R.parent[i] != -1) {
            const Wrench<S> up = invapply(Xli[i], f[i]);
            f[R.parent[i]] = {add(f[R.parent[i]].tau, up.tau), add(f[R.parent[i]].f, up.f)};
        }
    }
}

// Euclidean norm of a dynamic vector of doubles as Eigen's SSE2 reduction adds it (Redux.h, two packet accumulators)
double norm_packets(const std::vector<double>& x) {
    const int n = int(x.size()), a2 = (n / 4) * 4, a1 = (n / 2) * 2;
    std::vector<double> s(n);
    for (int i = 0; i < n; i++) s[i] = x[i] * x[i];
    if (a1 == 0) return std::sqrt(s[0]);
    double p0 = s[0], p1 = s[1];
    if (a1 > 2) {
        double r0 = s[2], r1 = s[3];
        for (int i = 4; i < a2; i += 4) {
            p0 += s[i];
            p1 += s[i + 1];
            r0 += s[i + 2];
            r1 += s[i + 3];
        }
        p0 += r0;
        p1 += r1;
        if (a1 > a2) {
            p0 += s[a2];
            p1 += s[a2 + 1];
        }
    }
    double res = p0 + p1;
    for (int i = a1; i < n; i++) res += s[i];
    return std::sqrt(res);
}
double wrap(double a) {
    while (a >= M_PI) a -= 6.283185307179586476925286766559;
    while (a < -M_PI) a += 6.283185307179586476925286766559;
    return a;
}
}  // namespace

extern "C" {
void* orcctl_create(const char* model_file, double eps) {
    Robot* R = new Robot;
    if (!load_model(model_file, eps, *R)) {
        delete R;
        return nullptr;
    }
    return R;
}
void orcctl_destroy(void* h) { delete static_cast<Robot*>(h); }
int orcctl_num_joints(void* h) { return static_cast<Robot*>(h)->nj; }
void orcctl_rnea(void* h, const double* q, const double* qd, const double* qda, const double* qdd, int friction, int gravity,
                 double* tau) {
    newton_euler<double>(*static_cast<Robot*>(h), q, qd, qda, qdd, friction != 0, gravity != 0, tau);
}
void orcctl_rnea_int(void* h, const double* q, const double* qd, const double* qda, const double* qdd, int friction, int gravity,
                     double* lo, double* hi) {
    const Robot& R = *static_cast<Robot*>(h);
    std::vector<Interval> t(R.nj);
    newton_euler<Interval>(R, q, qd, qda, qdd, friction != 0, gravity != 0, t.data());
    for (int i = 0; i < R.nj; i++) {
        lo[i] = t[i].lo;
        hi[i] = t[i].hi;
    }
}
int orcctl_update(void* h, const double* Kr, double alpha, double V_max, double r_norm_threshold, int friction, const double* q,
                  const double* qd, const double* q_des, const double* qd_des, const double* qdd_des, double* u, double* u_nominal,
                  double* v) {
    const Robot& R = *static_cast<Robot*>(h);
    const int n = R.nj;
    std::vector<double> qa_d(n), qa_dd(n), r(n), zero(n, 0.0), un(n), bound(n);
    for (int i = 0; i < n; i++) {
        const double qdiff = wrap(q_des[i] - q[i]);
        qa_d[i] = qd_des[i] + Kr[i] * qdiff;
        qa_dd[i] = qdd_des[i] + Kr[i] * (qd_des[i] - qd[i]);
        r[i] = (qd_des[i] - qd[i]) + Kr[i] * qdiff;
    }
    std::vector<Interval> ui(n), Mr(n);
    newton_euler<double>(R, q, qd, qa_d.data(), qa_dd.data(), friction != 0, true, un.data());
    newton_euler<Interval>(R, q, qd, qa_d.data(), qa_dd.data(), friction != 0, true, ui.data());
    int status = 0;
    for (int i = 0; i < n; i++) {
        if (un[i] > ui[i].hi || un[i] < ui[i].lo) status = 1;
        const Interval phi = ui[i] - Interval(un[i]);
        bound[i] = std::max(std::fabs(phi.lo), std::fabs(phi.hi));
    }
    std::vector<double> vv(n, 0.0);
    const double r_norm = norm_packets(r);
    if (r_norm > r_norm_threshold) {
        newton_euler<Interval>(R, q, zero.data(), zero.data(), r.data(), false, false, Mr.data());
        Interval V(0.0);
        for (int i = 0; i < n; i++) V = V + (0.5 * r[i]) * Mr[i];
        const double hh = -V.hi + V_max;
        const double lambda = std::max(0.0, -alpha * hh / r_norm + norm_packets(bound));
        for (int i = 0; i < n; i++) vv[i] = -lambda * r[i] / r_norm;
    }
    for (int i = 0; i < n; i++) {
        u[i] = un[i] - vv[i];
        u_nominal[i] = un[i];
        v[i] = vv[i];
    }
    return status;
}
void orcctl_int_model(void* h, double* out) {
    const Robot& R = *static_cast<Robot*>(h);
    int k = 0;
    auto put = [&](const Interval& x) {
        out[k++] = x.lo;
        out[k++] = x.hi;
    };
    for (int i = 0; i < R.nj; i++) {
        const Joint<Interval>& J = R.iv[i];
        for (int a = 0; a < 3; a++) put(J.S_.w[a]);
        for (int a = 0; a < 3; a++) put(J.S_.v[a]);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) put(J.X.R[a][b]);
        for (int a = 0; a < 3; a++) put(J.X.p[a]);
        put(J.m);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) put(J.Ibar[a][b]);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) put(J.mch[a][b]);
    }
}
}
