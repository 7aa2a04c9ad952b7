// corridor_cli.cpp -- stand-in for stage 2 of the planner node (src/swarm_traj_planner_rbp.cpp L96-L104): reads a world
// (distance grid), Mission, T and initTraj from a text dump, runs SwarmPlanning::Corridor::update() and prints SFC / RSFC
// with full precision.  Dump: "res kx0 ky0 kz0 nx ny nz" + nx*ny*nz floats; "wxmin wymin wzmin wxmax wymax wzmax";
// "N M"; T; per agent radius; per agent M+1 points.   usage: corridor_cli dump.txt [key=value ...]
#include <cstdio>
#include <fstream>
#include <map>

#include <rbp_corridor.hpp>

using namespace SwarmPlanning;

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    std::ifstream in(argv[1]);
    if (!in) return 2;
    std::map<std::string, std::string> kv;
    for (int i = 2; i < argc; i++) {
        std::string a(argv[i]);
        size_t e = a.find('=');
        if (e != std::string::npos) kv[a.substr(0, e)] = a.substr(e + 1);
    }
    double res;
    int k0[3], n[3];
    in >> res >> k0[0] >> k0[1] >> k0[2] >> n[0] >> n[1] >> n[2];
    std::vector<float> edt((size_t)n[0] * n[1] * n[2]);
    for (auto &v : edt) in >> v;
    Param param;
    in >> param.world_x_min >> param.world_y_min >> param.world_z_min >> param.world_x_max >> param.world_y_max >> param.world_z_max;
    param.setParam(kv);
    int N, M;
    in >> N >> M;
    PlanResult pr;
    pr.T.resize(M + 1);
    for (auto &t : pr.T) in >> t;
    Mission mission;
    mission.qn = N;
    mission.quad_size.resize(N);
    for (auto &r : mission.quad_size) in >> r;
    pr.initTraj.resize(N);
    for (int qi = 0; qi < N; qi++)
        for (int j = 0; j <= M; j++) {
            float x, y, z;
            in >> x >> y >> z;
            pr.initTraj[qi].emplace_back(x, y, z);
        }
    if (!in) return 2;
    auto dm = std::make_shared<GridDistanceMap>(res, k0[0], k0[1], k0[2], n[0], n[1], n[2], std::move(edt));
    Corridor corridor(dm, mission, param);
    const bool sfc_only = kv.count("stage") && kv["stage"] == "sfc";   // host-only stage (no device): SFC without RSFC
    if (!(sfc_only ? corridor.update_sfc(false, &pr) : corridor.update(false, &pr))) { std::printf("update=false\n"); return 1; }
    std::printf("update=true\n");
    for (int qi = 0; qi < N; qi++) {
        std::printf("SFC %d %zu\n", qi, pr.SFC[qi].size());
        for (auto &b : pr.SFC[qi]) std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", b.first[0], b.first[1], b.first[2], b.first[3], b.first[4], b.first[5], b.second);
    }
    for (int qi = 0; qi < N && !sfc_only; qi++)
        for (int qj = qi + 1; qj < N; qj++)
            for (auto &r : pr.RSFC[qi][qj]) std::printf("R %.9g %.9g %.9g %.17g\n", (double)r.first.x(), (double)r.first.y(), (double)r.first.z(), r.second);
    return 0;
}
