// planner_cli.cpp -- minimal stand-in for the planner node's stage 3 (src/swarm_traj_planner_rbp.cpp L106-L116):
// reads a PlanResult{T, initTraj, SFC, RSFC} + Mission dump, runs SwarmPlanning::RBPPlanner::update() and writes the
// coefficient CSVs.  Dump format (text, written by tests / swarm_simulator_b200/synth.py::dump_text):
//   N M
//   T[0..M]
//   per agent: start[9] goal[9] radius max_vel[3] max_acc[3]
//   per agent: M+1 lines of initTraj (x y z)
//   per agent: nbox, then nbox lines "xmin ymin zmin xmax ymax zmax t_end"
//   per pair qi<qj: M lines "nx ny nz t_end"
// usage: planner_cli dump.txt out_dir [key=value ...]   (keys as in param.hpp, e.g. plan/sequential=true)
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>

#include <rbp_planner.hpp>

using namespace SwarmPlanning;

int main(int argc, char **argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s dump.txt out_dir [key=value ...]\n", argv[0]);
        return 2;
    }
    std::ifstream in(argv[1]);
    if (!in) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    std::map<std::string, std::string> kv;
    for (int i = 3; i < argc; i++) {
        std::string a(argv[i]);
        size_t e = a.find('=');
        if (e != std::string::npos) kv[a.substr(0, e)] = a.substr(e + 1);
    }
    kv["package_path"] = argv[2];
    Param param;
    param.setParam(kv);
    Mission mission;
    PlanResult pr;
    int N, M;
    in >> N >> M;
    pr.T.resize(M + 1);
    for (auto &t : pr.T) in >> t;
    mission.qn = N;
    mission.startState.assign(N, std::vector<double>(9));
    mission.goalState.assign(N, std::vector<double>(9));
    mission.max_vel.assign(N, std::vector<double>(3));
    mission.max_acc.assign(N, std::vector<double>(3));
    mission.quad_size.assign(N, 0);
    mission.quad_speed.assign(N, 1);
    for (int qi = 0; qi < N; qi++) {
        for (auto &v : mission.startState[qi]) in >> v;
        for (auto &v : mission.goalState[qi]) in >> v;
        in >> mission.quad_size[qi];
        for (auto &v : mission.max_vel[qi]) in >> v;
        for (auto &v : mission.max_acc[qi]) in >> v;
    }
    pr.initTraj.resize(N);
    for (int qi = 0; qi < N; qi++)
        for (int j = 0; j <= M; j++) {
            float x, y, z;
            in >> x >> y >> z;
            pr.initTraj[qi].emplace_back(x, y, z);
        }
    pr.SFC.resize(N);
    for (int qi = 0; qi < N; qi++) {
        int nb;
        in >> nb;
        for (int b = 0; b < nb; b++) {
            std::vector<double> box(6);
            double t;
            for (auto &v : box) in >> v;
            in >> t;
            pr.SFC[qi].emplace_back(box, t);
        }
    }
    pr.RSFC.resize(N);
    for (int qi = 0; qi < N; qi++) {
        pr.RSFC[qi].resize(N);
        for (int qj = qi + 1; qj < N; qj++)
            for (int ri = 0; ri < M; ri++) {
                float x, y, z;
                double t;
                in >> x >> y >> z >> t;
                pr.RSFC[qi][qj].emplace_back(octomap::point3d(x, y, z), t);
            }
    }
    if (!in) { std::fprintf(stderr, "truncated dump\n"); return 2; }
    param.log = true;  // always write the CSVs (the reference writes them when `log`)
    RBPPlanner planner(mission, param);
    bool ok = planner.update(false, &pr);
    if (!ok) { std::printf("update=false\n"); return 1; }
    std::printf("update=true time_scale=%.17g M=%d\n", planner.last_time_scale(), M);
    // full-precision dump of msgs_traj_coef for the tests
    std::ofstream out(std::string(argv[2]) + "/traj_coef.txt");
    out << std::setprecision(17);
    for (double v : pr.msgs_traj_info.data) out << v << " ";
    out << "\n";
    for (auto &msg : pr.msgs_traj_coef) {
        out << msg.layout.dim[0].size << " " << msg.layout.dim[1].size;
        for (double v : msg.data) out << " " << v;
        out << "\n";
    }
    return 0;
}
