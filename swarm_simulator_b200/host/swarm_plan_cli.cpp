// swarm_plan_cli.cpp -- the planner node's whole pipeline (src/swarm_traj_planner_rbp.cpp L70-L123) without ROS:
//   world (.bt) -> clamped distance map -> ECBSPlanner -> Corridor -> RBPPlanner (the B200 engine) -> safety metrics.
// usage: swarm_plan_cli mission.json world.bt out_dir [stage=world|ecbs|sfc|all] [key=value ...]   (keys as in param.hpp)
// stage=world / ecbs need no GPU (map statistics, initial trajectories); stage=all runs Corridor's RSFC kernel and the QP
// engine on the device and writes the coefficient CSVs when log=true.
#include <cstdio>
#include <map>
#include <string>

#include <ecbs_planner.hpp>
#include <octree_bt.hpp>
#include <rbp_corridor.hpp>
#include <rbp_planner.hpp>

using namespace SwarmPlanning;

int main(int argc, char **argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s mission.json world.bt out_dir [stage=world|ecbs|sfc|all] [key=value ...]\n", argv[0]);
        return 2;
    }
    std::map<std::string, std::string> kv;
    for (int i = 4; i < argc; i++) {
        std::string a(argv[i]);
        size_t e = a.find('=');
        if (e != std::string::npos) kv[a.substr(0, e)] = a.substr(e + 1);
    }
    kv["package_path"] = argv[3];
    const std::string stage = kv.count("stage") ? kv["stage"] : "all";
    Param param;
    // launch defaults of plan_rbp_random_forest.launch L29-L65 where they differ from Param's
    param.grid_xy_res = 0.5; param.grid_z_res = 1.0; param.sequential = true; param.batch_size = 4; param.batch_iter = -1;
    param.world_z_min = 0.3; param.world_z_max = 2.5;
    param.setParam(kv);
    Mission mission;
    if (!mission.setMission(argv[1])) { std::fprintf(stderr, "cannot read mission %s\n", argv[1]); return 2; }
    OcTreeBt tree;
    if (!tree.load(argv[2])) { std::fprintf(stderr, "%s\n", tree.error.c_str()); return 2; }
    const double lo[3] = {param.world_x_min, param.world_y_min, param.world_z_min};
    const double hi[3] = {param.world_x_max, param.world_y_max, param.world_z_max};
    int k0[3], n[3];
    std::vector<unsigned char> occ;
    tree.rasterize(lo, hi, k0, n, occ);
    size_t nocc = 0, nleaf_occ = 0;
    for (unsigned char v : occ) nocc += v;
    for (const auto &l : tree.leaves) nleaf_occ += l.occupied;
    std::vector<float> edt = clamped_edt(occ, n, tree.res, 1.0);
    std::printf("world res=%g declared_nodes=%zu inner=%zu leaves=%zu occupied_leaves=%zu grid=%dx%dx%d occupied_cells=%zu\n", tree.res,
                tree.declared_nodes, tree.inner_nodes, tree.leaves.size(), nleaf_occ, n[0], n[1], n[2], nocc);
    if (stage == "world") {
        // occupied columns (x, y cells with any occupied voxel), for the tests
        for (int x = 0; x < n[0]; x++)
            for (int y = 0; y < n[1]; y++) {
                int h = 0;
                for (int z = 0; z < n[2]; z++) h += occ[((size_t)x * n[1] + y) * n[2] + z];
                if (h) std::printf("col %d %d %d\n", x + k0[0], y + k0[1], h);
            }
        return 0;
    }
    auto dm = std::make_shared<GridDistanceMap>(tree.res, k0[0], k0[1], k0[2], n[0], n[1], n[2], std::move(edt));
    PlanResult pr;
    ECBSPlanner ecbs(dm, mission, param);
    if (!ecbs.update(param.log, &pr)) { std::printf("ecbs=false\n"); return 1; }
    const int M = (int)pr.T.size() - 1;
    std::printf("ecbs=true M=%d high_level_nodes=%ld\n", M, ecbs.high_level_expanded);
    if (stage == "ecbs") {
        for (int qi = 0; qi < mission.qn; qi++) {
            std::printf("traj %d", qi);
            for (auto &p : pr.initTraj[qi]) std::printf(" %.9g %.9g %.9g", (double)p.x(), (double)p.y(), (double)p.z());
            std::printf("\n");
        }
        return 0;
    }
    Corridor corridor(dm, mission, param);
    if (stage == "sfc") {   // host only: ECBS + SFC, dumped for offline reproduction (T, initTraj, SFC; RSFC follows from initTraj)
        if (!corridor.update_sfc(param.log, &pr)) { std::printf("corridor=false\n"); return 1; }
        std::printf("dump %d %d\n", mission.qn, M);
        for (auto t : pr.T) std::printf("%.17g ", t);
        std::printf("\n");
        for (int qi = 0; qi < mission.qn; qi++) {
            for (auto &p : pr.initTraj[qi]) std::printf("%.9g %.9g %.9g ", (double)p.x(), (double)p.y(), (double)p.z());
            std::printf("\n%zu\n", pr.SFC[qi].size());
            for (auto &b : pr.SFC[qi]) std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", b.first[0], b.first[1], b.first[2], b.first[3], b.first[4], b.first[5], b.second);
        }
        return 0;
    }
    if (!corridor.update(param.log, &pr)) { std::printf("corridor=false\n"); return 1; }
    size_t nbox = 0;
    for (auto &s : pr.SFC) nbox += s.size();
    std::printf("corridor=true sfc_boxes=%zu\n", nbox);
    RBPPlanner planner(mission, param);
    if (!planner.update(param.log, &pr)) { std::printf("rbp=false\n"); return 1; }
    std::printf("rbp=true time_scale=%.9g makespan=%.9g\n", planner.last_time_scale(), pr.T.back());
    // RBPPublisher's post-hoc checks (rbp_publisher.hpp L117-L127) on the device
    {
        rbpe_handle *h = nullptr;
        rbpe_config cfg = {};
        if (rbpe_create(&cfg, &h) != RBPE_OK) return 1;
        const int N = mission.qn;
        std::vector<double> coef((size_t)N * 18 * M);
        for (int qi = 0; qi < N; qi++)
            for (size_t i = 0; i < (size_t)18 * M; i++) coef[(size_t)qi * 18 * M + i] = pr.msgs_traj_coef[qi].data[i];
        double ratio = 0, t_at = 0, len = 0;
        int rc = rbpe_safety_metrics(h, N, M, 1, coef.data(), pr.T.data(), mission.quad_size.data(), param.downwash, 0.1, &ratio, &t_at, &len);
        rbpe_destroy(h);
        if (rc != RBPE_OK) return 1;
        std::printf("safety_margin_ratio=%.9g at t=%.9g total_flight_length=%.9g\n", ratio, t_at, len);
    }
    return 0;
}
