// time_scale_cli.cpp -- host-only driver of RBPPlanner's timeScale stage (rbp_planner.hpp L209-L266 of the reference) for
// tests/test_time_scale.py.  stdin: "N M rule vmax amax", T[0..M], then per agent 3 * 6M coefficients (axis-major, segment,
// highest power first).  stdout: "scale=<s>" and the rescaled coefficients.
#include <cstdio>
#include <iostream>
#include <map>

#include <rbp_planner.hpp>

using namespace SwarmPlanning;

int main() {
    int N, M, rule;
    double vmax, amax;
    if (!(std::cin >> N >> M >> rule >> vmax >> amax)) return 2;
    PlanResult pr;
    pr.T.resize(M + 1);
    for (auto &t : pr.T) std::cin >> t;
    Mission mission;
    mission.qn = N;
    mission.max_vel.assign(N, {vmax, vmax, vmax});
    mission.max_acc.assign(N, {amax, amax, amax});
    mission.quad_size.assign(N, 0.15);
    std::vector<std::vector<double>> coef(N, std::vector<double>((size_t)18 * M));
    for (auto &c : coef)
        for (auto &v : c) std::cin >> v;
    pr.SFC.assign(N, {});
    pr.RSFC.assign(N, {});
    for (auto &r : pr.RSFC) r.resize(N);
    Param param;
    std::map<std::string, std::string> kv{{"plan/time_scale_roots", std::to_string(rule)}};
    param.setParam(kv);
    RBPPlanner planner(mission, param);
    double s = planner.time_scale_only(coef, &pr);
    std::printf("scale=%.17g\n", s);
    for (auto &c : coef) {
        for (auto &v : c) std::printf("%.17g ", v);
        std::printf("\n");
    }
    return 0;
}
