// host_selftest.cpp -- CPU-only checks of the host mirror (no engine call): mission JSON loader, Param defaults / keys,
// octomap::point3d stand-in semantics.  Prints key=value lines that tests/test_host_mirror.py compares.
#include <cstdio>
#include <map>

#include <mission.hpp>
#include <param.hpp>
#include <sp_const.hpp>

int main(int argc, char **argv) {
    SwarmPlanning::Param p;
    std::printf("default sequential=%d batch_size=%d batch_iter=%d iteration=%d n=%d phi=%d downwash=%g time_scale=%d grid_xy=%g z_max=%g\n",
                p.sequential, p.batch_size, p.batch_iter, p.iteration, p.n, p.phi, p.downwash, p.time_scale, p.grid_xy_res, p.world_z_max);
    p.setParam({{"plan/sequential", "true"}, {"plan/batch_size", "8"}, {"plan/batch_iter", "-1"}, {"world/z_min", "0.3"}});
    std::printf("set sequential=%d batch_size=%d batch_iter=%d z_min=%g\n", p.sequential, p.batch_size, p.batch_iter, p.world_z_min);
    if (argc > 1) {
        SwarmPlanning::Mission m;
        bool ok = m.setMission(argv[1]);
        std::printf("mission ok=%d qn=%d\n", ok, m.qn);
        for (int qi = 0; ok && qi < m.qn; qi++)
            std::printf("agent %d start=%g,%g,%g goal=%g,%g,%g r=%g vmax=%g amax=%g\n", qi, m.startState[qi][0], m.startState[qi][1],
                        m.startState[qi][2], m.goalState[qi][0], m.goalState[qi][1], m.goalState[qi][2], m.quad_size[qi],
                        m.max_vel[qi][0], m.max_acc[qi][2]);
    }
    octomap::point3d a(1.f, 2.f, 2.f), b(0.f, 0.f, 1.f);
    octomap::point3d c = a - b;
    c.normalize();
    std::printf("vec norm=%.9g dot=%.9g nx=%.9g eq=%d\n", a.norm(), a.dot(b), (double)c.x(), (int)(a == a));
    return 0;
}
