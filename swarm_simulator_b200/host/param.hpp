// param.hpp -- ROS-free mirror of /root/reference/swarm_planner/include/param.hpp (L8-L75): same member names, same
// defaults as Param::setROSParam (L44-L75); values come from a string map ("plan/batch_size" -> "4") instead of the
// ROS parameter server.  Param::setColor is visualisation only and is not mirrored.
#pragma once
#include <cstdlib>
#include <map>
#include <string>

namespace SwarmPlanning {
class Param {
public:
    bool log = false;
    std::string package_path = ".";

    double world_x_min = -5, world_y_min = -5, world_z_min = 0;
    double world_x_max = 5, world_y_max = 5, world_z_max = 2.5;

    double ecbs_w = 1.3;
    double grid_xy_res = 0.3, grid_z_res = 0.6, grid_margin = 0.2;

    double box_xy_res = 0.1, box_z_res = 0.1;

    bool time_scale = true;
    // Not in the reference: which roots of the acceleration cubic the velocity check of timeScale looks at.
    //   0 = every real root (default here: the true extrema);
    //   2 = the reference rule, "the first two eigenvalues of the companion matrix that are real" (rbp_planner.hpp
    //       L741-L753), with the eigenvalues taken in order of decreasing modulus -- see rbp_planner.hpp (timeScale)
    int time_scale_roots = 0;
    double time_step = 1;
    double downwash = 2.0;  // downwash coefficient
    int iteration = 1;
    bool sequential = false;
    int batch_size = 4;     // the number of agents in a batch
    int batch_iter = 0;     // the number of batches
    int n = 5;              // degree of polynomial
    int phi = 3;            // desired derivatives

    // same keys as the private ROS params of the planner node ("world/x_min", "plan/sequential", ...)
    bool setParam(const std::map<std::string, std::string> &kv) {
        auto d = [&](const char *k, double &v) { auto it = kv.find(k); if (it != kv.end()) v = std::atof(it->second.c_str()); };
        auto i = [&](const char *k, int &v) { auto it = kv.find(k); if (it != kv.end()) v = std::atoi(it->second.c_str()); };
        auto b = [&](const char *k, bool &v) {
            auto it = kv.find(k);
            if (it != kv.end()) v = (it->second == "true" || it->second == "1" || it->second == "True");
        };
        b("log", log);
        d("world/x_min", world_x_min); d("world/y_min", world_y_min); d("world/z_min", world_z_min);
        d("world/x_max", world_x_max); d("world/y_max", world_y_max); d("world/z_max", world_z_max);
        d("grid/xy_res", grid_xy_res); d("grid/z_res", grid_z_res); d("grid/margin", grid_margin);
        d("ecbs/w", ecbs_w);
        d("box/xy_res", box_xy_res); d("box/z_res", box_z_res);
        b("plan/time_scale", time_scale); i("plan/time_scale_roots", time_scale_roots); d("plan/time_step", time_step); d("plan/downwash", downwash);
        i("plan/n", n); i("plan/phi", phi);
        b("plan/sequential", sequential); i("plan/batch_size", batch_size); i("plan/batch_iter", batch_iter);
        i("plan/iteration", iteration);
        auto it = kv.find("package_path");
        if (it != kv.end()) package_path = it->second;
        return true;
    }
};
}  // namespace SwarmPlanning
