// rbp_planner.hpp -- host-side mirror of SwarmPlanning::RBPPlanner
//   /root/reference/swarm_planner/include/rbp_planner.hpp L18-L84
// Same class name, namespace, constructor and `bool update(bool log, PlanResult*)`; same ownership (Mission / Param are
// copied, PlanResult is the caller's and is mutated in place: reads T, initTraj, SFC, RSFC; writes msgs_traj_info,
// msgs_traj_coef, and rescales T / SFC / RSFC times when plan/time_scale applies, L249-L264); never throws; returns
// false where the reference does (QP infeasible / solver failure, L62-L69 and L158-L161).
//
// What differs is the inside: buildConstMtx + solveQP (constraint assembly, one CPLEX call per batch, Bernstein ->
// monomial conversion, `dummy` propagation) are one call into the B200 engine through the C ABI of include/rbpe.h.
// timeScale, generateROSMsg and generateCoefCSV stay on the host (they are O(N M) arithmetic on the result).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

#include <mission.hpp>
#include <param.hpp>
#include <sp_const.hpp>

#include "rbpe.h"

namespace SwarmPlanning {
class RBPPlanner {
public:
    RBPPlanner(Mission _mission, Param _param) : mission(std::move(_mission)), param(std::move(_param)) {
        n = param.n;       // degree of polynomial
        phi = param.phi;   // desired derivatives
        N = mission.qn;    // the number of agents
        outdim = 3;        // the number of outputs (x,y,z)
    }
    ~RBPPlanner() {
        if (engine) rbpe_destroy(engine);
    }
    RBPPlanner(const RBPPlanner &) = delete;
    RBPPlanner &operator=(const RBPPlanner &) = delete;

    bool update(bool log, SwarmPlanning::PlanResult *_planResult_ptr) {
        planResult_ptr = _planResult_ptr;
        M = (int)planResult_ptr->T.size() - 1;  // the number of segments
        offset_quad = M * (n + 1);
        offset_seg = n + 1;
        if (n != 5 || phi != 3) {  // the reference's hard-coded matrices exist for these values only (L328, L361)
            std::fprintf(stderr, "RBPPlanner: n=5, phi=3 only\n");
            return false;
        }
        if (!engine) {
            rbpe_config cfg = {};
            cfg.device = device;
            if (rbpe_create(&cfg, &engine) != RBPE_OK) {
                std::fprintf(stderr, "RBPPlanner: %s\n", rbpe_last_error(nullptr));
                return false;
            }
        }
        if (!solveQP(log)) return false;
        if (param.time_scale) timeScale();
        generateROSMsg();
        if (param.log) generateCoefCSV();
        return true;
    }

    // results of the last update (coef[qi] is the reference's Eigen matrix M(n+1) x 3, stored column-major)
    const std::vector<std::vector<double>> &coefficients() const { return coef; }
    const std::vector<double> &batch_cost() const { return qp_obj; }
    const std::vector<int> &batch_iterations() const { return qp_iters; }
    double last_time_scale() const { return time_scale_used; }
    int device = 0;

    // Host-only entry to the timeScale stage (L209-L266) for given monomial coefficients (coef_in[qi] column-major
    // M(n+1) x 3, highest power first per segment): rescales coef_in and *pr exactly as update() would and returns the scale.
    // Used by tests/test_time_scale.py (no device needed).
    double time_scale_only(std::vector<std::vector<double>> &coef_in, SwarmPlanning::PlanResult *pr) {
        planResult_ptr = pr;
        M = (int)pr->T.size() - 1;
        offset_quad = M * (n + 1);
        offset_seg = n + 1;
        coef = coef_in;
        timeScale();
        coef_in = coef;
        return time_scale_used;
    }

private:
    Mission mission;
    Param param;
    SwarmPlanning::PlanResult *planResult_ptr = nullptr;
    rbpe_handle *engine = nullptr;
    int M = 0, n = 5, phi = 3, N = 0, outdim = 3, offset_quad = 0, offset_seg = 0;
    std::vector<std::vector<double>> coef;  // [N][3 * 6M] column-major
    std::vector<double> qp_obj;
    std::vector<int> qp_iters;
    double time_scale_used = 1;

    double &C(int qi, int row, int k) { return coef[qi][(size_t)k * offset_quad + row]; }

    bool solveQP(bool log) {
        PlanResult &pr = *planResult_ptr;
        const size_t P = (size_t)N * (N - 1) / 2;
        std::vector<double> start((size_t)N * 9), goal((size_t)N * 9), radius(N), sfc_box, sfc_t, rsfc_t(P * M);
        std::vector<int> sfc_offs(N + 1, 0);
        std::vector<float> rsfc_n(P * M * 3), init_traj((size_t)N * (M + 1) * 3, 0.f);
        for (int qi = 0; qi < N; qi++) {
            for (int i = 0; i < 9; i++) {
                start[(size_t)qi * 9 + i] = mission.startState[qi][i];
                goal[(size_t)qi * 9 + i] = mission.goalState[qi][i];
            }
            radius[qi] = mission.quad_size[qi];
            for (auto &b : pr.SFC[qi]) {
                if (b.first.size() != 6) return false;
                sfc_box.insert(sfc_box.end(), b.first.begin(), b.first.end());
                sfc_t.push_back(b.second);
            }
            sfc_offs[qi + 1] = (int)sfc_t.size();
            if (param.sequential) {
                // a path shorter than M + 1 points holds its last point (build_dummy L519-L528: `idx >= path_size - 1`)
                const int npts = (int)pr.initTraj[qi].size();
                if (npts < 1) return false;
                for (int j = 0; j <= M; j++)
                    for (int k = 0; k < 3; k++)
                        init_traj[((size_t)qi * (M + 1) + j) * 3 + k] = pr.initTraj[qi][j < npts ? j : npts - 1](k);
            }
        }
        size_t it = 0;
        for (int qi = 0; qi < N; qi++)
            for (int qj = qi + 1; qj < N; qj++, it++) {
                if ((int)pr.RSFC[qi][qj].size() < M) return false;
                for (int ri = 0; ri < M; ri++) {
                    for (int k = 0; k < 3; k++) rsfc_n[(it * M + ri) * 3 + k] = pr.RSFC[qi][qj][ri].first(k);
                    rsfc_t[it * M + ri] = pr.RSFC[qi][qj][ri].second;
                }
            }
        rbpe_problem p = {};
        p.N = N; p.M = M; p.sequential = param.sequential; p.batch_size = param.batch_size;
        p.batch_iter = param.batch_iter; p.iteration = param.iteration;
        p.T = pr.T.data(); p.start = start.data(); p.goal = goal.data(); p.radius = radius.data();
        p.sfc_offs = sfc_offs.data(); p.sfc_base = nullptr; p.sfc_box = sfc_box.data(); p.sfc_t = sfc_t.data();
        p.rsfc_n = rsfc_n.data(); p.rsfc_t = rsfc_t.data(); p.init_traj = init_traj.data();

        int ebs = 0, ebi = 0;
        rbpe_set_batch(N, param.sequential, param.batch_size, param.batch_iter, &ebs, &ebi);
        size_t nrec = (size_t)param.iteration * ebi;
        std::vector<double> flat((size_t)N * 3 * offset_quad);
        qp_obj.assign(nrec ? nrec : 1, 0.0);
        qp_iters.assign(nrec ? nrec : 1, 0);
        std::vector<int> qp_status(nrec ? nrec : 1, 0);
        int status = 0;
        rbpe_result r = {};
        r.coef = flat.data(); r.qp_obj = qp_obj.data(); r.qp_iters = qp_iters.data(); r.qp_status = qp_status.data();
        r.status = &status;
        int rc = rbpe_solve(engine, &p, &r);
        if (rc != RBPE_OK) {
            std::fprintf(stderr, "RBPPlanner: Failed to optimize QP (%d): %s\n", rc, rbpe_last_error(engine));
            return false;
        }
        coef.assign(N, std::vector<double>((size_t)3 * offset_quad));
        for (int qi = 0; qi < N; qi++)
            std::copy(flat.begin() + (size_t)qi * 3 * offset_quad, flat.begin() + (size_t)(qi + 1) * 3 * offset_quad,
                      coef[qi].begin());
        if (log) {
            double total = 0;
            for (size_t l = 0; l < nrec; l++) {
                if (param.sequential)
                    std::fprintf(stderr, "RBPPlanner: QP cost of batch %zu=%g (%d interior-point iterations)\n", l % ebi,
                                 qp_obj[l], qp_iters[l]);
                total += qp_obj[l];
            }
            std::fprintf(stderr, "RBPPlanner: QP total cost=%g\n", total);
        }
        return true;
    }

    // ---- timeScale (L209-L266) with its helpers (L708-L847) -------------------------------------------------------
    // coef_der(i, .) = coefficients (highest power first) of the i-th derivative of segment (qi, k, m)
    void derivative_segment(int qi, int k, int m, double cd[4][6]) {
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 6; j++) {
                double f = 1;
                for (int t = 0; t < i; t++) f *= (j - t);          // coef_derivative(i, j) = j (j-1) ... (j-i+1)
                cd[i][5 - j] = (i <= j) ? f * C(qi, m * offset_seg + 5 - j, k) : 0.0;
            }
    }
    // Roots of a x^3 + b x^2 + c x + d.  The reference takes the eigenvalues of the companion matrix from
    // Eigen::EigenSolver and keeps those of THE FIRST TWO that are real (`for j < i`, i = 2: L741-L753), so with three real
    // roots one extremum candidate is silently dropped -- which one depends on the order in which Eigen's real Schur
    // iteration deflates the 3 x 3 companion matrix (Eigen is absent here; that order cannot be reproduced bit for bit).
    // Two selectable rules (Param::time_scale_roots):
    //   0 (default)  every real root: the velocity bound is checked at all true extrema;
    //   2            the reference rule with a DOCUMENTED order: eigenvalues by decreasing modulus (complex pairs by their
    //                modulus, the member with positive imaginary part first), the first two inspected, real ones kept.
    //                The shifted QR iteration converges its bottom diagonal entry to the eigenvalue of smallest modulus
    //                first, which leaves the larger ones in the leading positions -- the closest statement of Eigen's
    //                order that does not need Eigen.
    // tests/test_time_scale.py measures how often, and by how much, the two rules differ.
    static std::vector<double> real_roots_cubic(double a, double b, double c, double d) {
        std::vector<double> r;
        const double PI = 3.14159265358979323846;
        if (a == 0) {
            if (b == 0) { if (c != 0) r.push_back(-d / c); return r; }
            double D = c * c - 4 * b * d;
            if (D >= 0) { r.push_back((-c + std::sqrt(D)) / (2 * b)); r.push_back((-c - std::sqrt(D)) / (2 * b)); }
            return r;
        }
        double p = b / a, q = c / a, s = d / a;
        double A = q - p * p / 3, B = 2 * p * p * p / 27 - p * q / 3 + s, disc = B * B / 4 + A * A * A / 27;
        if (disc > 0) {
            double u = std::cbrt(-B / 2 + std::sqrt(disc)), v = std::cbrt(-B / 2 - std::sqrt(disc));
            r.push_back(u + v - p / 3);
        } else if (A == 0) {
            r.push_back(-p / 3);
        } else {
            double rr = 2 * std::sqrt(-A / 3), arg = 3 * B / (A * rr);
            arg = arg < -1 ? -1 : (arg > 1 ? 1 : arg);
            double th = std::acos(arg) / 3;
            for (int k = 0; k < 3; k++) r.push_back(rr * std::cos(th - 2 * PI * k / 3) - p / 3);
        }
        return r;
    }
    // extremum candidates of the velocity under the selected rule (see above)
    std::vector<double> velocity_extrema(double a, double b, double c, double d) const {
        std::vector<double> all = real_roots_cubic(a, b, c, d);
        if (param.time_scale_roots != 2) return all;
        int deg = a != 0 ? 3 : (b != 0 ? 2 : (c != 0 ? 1 : 0));
        struct Ev { double re, im, mod; };
        std::vector<Ev> ev;
        for (double r : all) ev.push_back({r, 0.0, std::fabs(r)});
        if ((int)all.size() < deg) {   // the missing eigenvalues form one complex pair: from the sum and product of the roots
            double lead = deg == 3 ? a : b, s1 = deg == 3 ? -b / a : -c / b, prod = deg == 3 ? -d / a : d / b;
            (void)lead;
            double re = (deg == 3) ? (s1 - all[0]) / 2 : s1 / 2;
            double mod2 = (deg == 3) ? (all[0] != 0 ? prod / all[0] : 0.0) : prod;
            if (deg == 3 && all[0] == 0) mod2 = c / a;               // x (a x^2 + b x + c): pair of the quadratic factor
            double mod = std::sqrt(std::fabs(mod2)), im2 = mod2 - re * re;
            double im = im2 > 0 ? std::sqrt(im2) : 0.0;
            ev.push_back({re, im, mod});
            ev.push_back({re, -im, mod});
        }
        std::stable_sort(ev.begin(), ev.end(), [](const Ev &x, const Ev &y) { return x.mod > y.mod; });
        std::vector<double> out;
        for (size_t j = 0; j < ev.size() && j < 2; j++)
            if (ev[j].im == 0) out.push_back(ev[j].re);
        return out;
    }
    double scale_to_max_vel(int qi, int k, int m, const double cd[4][6]) {
        const double dt = planResult_ptr->T[m + 1] - planResult_ptr->T[m], rate = 1.1;
        double vel_max = 0, t_max = 0;
        // extrema of the velocity = roots of the acceleration (cubic: cd[2][0..3])
        int lead = 0;
        while (lead < 3 && cd[2][lead] == 0) lead++;
        std::vector<double> ts;
        if (lead < 3) {
            double co[4] = {0, 0, 0, 0};
            for (int j = lead; j < 4; j++) co[j] = cd[2][j];
            ts = velocity_extrema(co[0], co[1], co[2], co[3]);
        }
        ts.push_back(0);
        ts.push_back(dt);
        for (double t : ts) {
            if (t < 0 || t > dt) continue;
            double vel = 0;
            for (int i = 0; i <= 4; i++) vel += cd[1][i] * std::pow(t, 4 - i);
            vel = std::fabs(vel);
            if (vel_max < vel) { vel_max = vel; t_max = t; }
        }
        double ts_ = 1;
        while (vel_max > mission.max_vel[qi][k]) {   // L781-L790: rescaled polynomial evaluated at the unscaled t_max
            ts_ *= rate;
            double vel = 0;
            for (int i = 0; i <= 4; i++) vel += cd[1][i] * std::pow(1 / ts_, 5 - i) * std::pow(t_max, 4 - i);
            vel_max = std::fabs(vel);
        }
        return ts_;
    }
    double scale_to_max_acc(int qi, int k, int m, const double cd[4][6]) {
        const double dt = planResult_ptr->T[m + 1] - planResult_ptr->T[m], rate = 1.1;
        double a = cd[3][0], b = cd[3][1], c = cd[3][2], D = b * b - 4 * a * c, acc_max = 0, t_max = 0;
        std::vector<double> ts{0, dt};
        if (D >= 0 && a != 0) { ts.push_back((-b + std::sqrt(D)) / (2 * a)); ts.push_back((-b - std::sqrt(D)) / (2 * a)); }
        else if (a == 0 && b != 0) ts.push_back(-c / b);
        for (double t : ts) {
            if (t < 0 || t > dt) continue;
            double acc = 0;
            for (int i = 0; i < 4; i++) acc += cd[2][i] * std::pow(t, 3 - i);
            acc = std::fabs(acc);
            if (acc_max < acc) { acc_max = acc; t_max = t; }
        }
        double ts_ = 1;
        while (acc_max > mission.max_acc[qi][k]) {   // L835-L844
            ts_ *= rate;
            double acc = 0;
            for (int i = 0; i < 4; i++) acc += cd[2][i] * std::pow(1 / ts_, 5 - i) * std::pow(t_max, 3 - i);
            acc_max = std::fabs(acc);
        }
        return ts_;
    }
    void timeScale() {
        double time_scale = 1;
        for (int qi = 0; qi < N; qi++)
            for (int k = 0; k < outdim; k++)
                for (int m = 0; m < M; m++) {
                    double cd[4][6];
                    derivative_segment(qi, k, m, cd);
                    double t1 = scale_to_max_vel(qi, k, m, cd);
                    if (time_scale < t1) time_scale = t1;
                    double t2 = scale_to_max_acc(qi, k, m, cd);
                    if (time_scale < t2) time_scale = t2;
                }
        time_scale_used = time_scale;
        if (time_scale != 1) {
            for (int qi = 0; qi < N; qi++) {
                for (int k = 0; k < outdim; k++)
                    for (int m = 0; m < M; m++)
                        for (int i = 0; i <= n; i++) C(qi, m * offset_seg + i, k) *= std::pow(1.0 / time_scale, n - i);
                for (auto &b : planResult_ptr->SFC[qi]) b.second *= time_scale;
                for (int qj = qi + 1; qj < N; qj++)
                    for (auto &r : planResult_ptr->RSFC[qi][qj]) r.second *= time_scale;
            }
            for (int m = 0; m < M + 1; m++) planResult_ptr->T[m] *= time_scale;
        }
    }

    // generateROSMsg (L269-L293)
    void generateROSMsg() {
        std::vector<double> traj_info;
        traj_info.emplace_back(N);
        traj_info.emplace_back(n);
        traj_info.insert(traj_info.end(), planResult_ptr->T.begin(), planResult_ptr->T.end());
        planResult_ptr->msgs_traj_info.data = traj_info;
        planResult_ptr->msgs_traj_coef.resize(N);
        for (int qi = 0; qi < N; qi++) {
            std_msgs::MultiArrayDimension rows, cols;
            rows.size = M * (n + 1);
            cols.size = outdim;
            planResult_ptr->msgs_traj_coef[qi].layout.dim.emplace_back(rows);
            planResult_ptr->msgs_traj_coef[qi].layout.dim.emplace_back(cols);
            planResult_ptr->msgs_traj_coef[qi].data.insert(planResult_ptr->msgs_traj_coef[qi].data.end(), coef[qi].begin(),
                                                           coef[qi].end());
        }
    }

    // generateCoefCSV (L295-L324): Crazyswarm trajectory format, lowest power first, padded to 8 per axis + 8 yaw zeros
    void generateCoefCSV() {
        for (int qi = 0; qi < N; qi++) {
            std::ofstream f(param.package_path + "/log/coef" + std::to_string(qi + 1) + ".csv");
            f << "duration,x^0,x^1,x^2,x^3,x^4,x^5,x^6,x^7,y^0,y^1,y^2,y^3,y^4,y^5,y^6,y^7,z^0,z^1,z^2,z^3,z^4,z^5,z^6,z^7,"
                 "yaw^0,yaw^1,yaw^2,yaw^3,yaw^4,yaw^5,yaw^6,yaw^7\n";
            for (int m = 0; m < M; m++) {
                f << planResult_ptr->T[m + 1] - planResult_ptr->T[m] << ",";
                for (int k = 0; k < outdim; k++) {
                    for (int i = n; i >= 0; i--) f << C(qi, m * offset_seg + i, k) << ",";
                    for (int i = 0; i < 7 - n; i++) f << "0,";
                }
                for (int i = 0; i < 8; i++) f << "0,";
                f << "\n";
            }
        }
    }
};
}  // namespace SwarmPlanning
