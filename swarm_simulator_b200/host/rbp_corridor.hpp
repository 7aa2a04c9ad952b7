// rbp_corridor.hpp -- host-side mirror of SwarmPlanning::Corridor
//   /root/reference/swarm_planner/include/rbp_corridor.hpp L11-L26 (interface), L44-L147 (box tests and expansion),
//   L149-L243 (updateObsBox), L338-L398 (updateRelBox)
// Same constructor shape `Corridor(distmap, mission, param)` and `bool update(bool log, PlanResult*)`.
// The SFC part stays on the host as in the north-star (it is a sequence of distance-map look-ups); the distance map is
// any type with `float getDistance(const octomap::point3d&) const` -- DynamicEDTOctomap in the reference
// (swarm_traj_planner_rbp.cpp L76-L80), GridDistanceMap below where octomap is absent.  The RSFC part is pure float32
// arithmetic and runs on the device through rbpe_corridor_rsfc() (include/rbpe.h).
// update_flat_box (the IROS-2019 axis-sector variant, L400-L772) is not mirrored.
#pragma once

#include <cmath>
#include <cstdio>
#include <memory>
#include <vector>

#include <mission.hpp>
#include <param.hpp>
#include <sp_const.hpp>

#include "rbpe.h"

namespace SwarmPlanning {

// Clamped Euclidean distance map on a regular voxel grid with DynamicEDTOctomap::getDistance semantics: distance (metres)
// of the voxel holding the point, -1 outside the bounding box.
class GridDistanceMap {
public:
    GridDistanceMap(double res, int kx0, int ky0, int kz0, int nx, int ny, int nz, std::vector<float> edt)
        : res_(res), k0_{kx0, ky0, kz0}, n_{nx, ny, nz}, edt_(std::move(edt)) {}
    float getDistance(const octomap::point3d &p) const {
        int k[3];
        for (int a = 0; a < 3; a++) {
            k[a] = (int)std::floor((double)p(a) / res_) - k0_[a];
            if (k[a] < 0 || k[a] >= n_[a]) return -1.0f;
        }
        return edt_[((size_t)k[0] * n_[1] + k[1]) * n_[2] + k[2]];
    }
private:
    double res_;
    int k0_[3], n_[3];
    std::vector<float> edt_;
};

template <class DistMap>
class CorridorT {
public:
    CorridorT(std::shared_ptr<DistMap> _distmap_obj, Mission _mission, Param _param)
        : distmap_obj(std::move(_distmap_obj)), mission(std::move(_mission)), param(std::move(_param)) {}
    ~CorridorT() {
        if (engine) rbpe_destroy(engine);
    }

    bool update(bool _log, SwarmPlanning::PlanResult *_planResult_ptr) {
        log = _log;
        planResult_ptr = _planResult_ptr;
        makespan = planResult_ptr->T.back();
        return updateObsBox() && updateRelBox();
    }
    // SFC stage alone (host only, no device needed): what tests/test_corridor_properties.py drives through corridor_cli
    bool update_sfc(bool _log, SwarmPlanning::PlanResult *_planResult_ptr) {
        log = _log;
        planResult_ptr = _planResult_ptr;
        makespan = planResult_ptr->T.back();
        return updateObsBox();
    }

private:
    std::shared_ptr<DistMap> distmap_obj;
    Mission mission;
    Param param;
    bool log = false;
    SwarmPlanning::PlanResult *planResult_ptr = nullptr;
    double makespan = 0;
    rbpe_handle *engine = nullptr;

    // ---- SFC construction (reference semantics: rbp_corridor.hpp L44-L243) -------------------------------------------
    // Written in this repository's own structure; what is kept bit-faithful is the ARITHMETIC that decides the boxes:
    //   * the sample lattice of an obstacle test: per axis, coordinates accumulate in double (v += res) from the lower face
    //     while v < upper + 1e-6; a sample sits 1e-6 above its lattice coordinate, except the first one of an axis, which
    //     sits 1e-6 BELOW the lower face unless that face is the world boundary (L47-L63);
    //   * a box passes when no sample's clamped distance is below margin - 1e-6 (L64-L68) and it lies inside the world
    //     with 1e-9 slack (L80-L87);
    //   * greedy growth: faces are tried round-robin in the order -x,-y,-z,+x,+y,+z, one resolution step at a time, only
    //     the newly added slab is tested; a face that fails is retired, and after every retirement the whole box is
    //     re-tested before the next face is tried (L99-L147);
    //   * box <-> time allocation: a box hands over to the next one at the middle of the stretch of path points that lie
    //     in both (L195-L237); the last box ends at the makespan.
    struct Aabb {
        double lo[3], hi[3];
        std::vector<double> as_sfc() const { return {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]}; }
    };
    double axis_res(int a) const { return a == 2 ? param.box_z_res : param.box_xy_res; }
    double world_lo(int a) const { return a == 0 ? param.world_x_min : (a == 1 ? param.world_y_min : param.world_z_min); }
    double world_hi(int a) const { return a == 0 ? param.world_x_max : (a == 1 ? param.world_y_max : param.world_z_max); }

    // sample coordinates of one axis of a box
    void axis_samples(const Aabb &b, int a, std::vector<double> &out) const {
        out.clear();
        const bool nudge_first_down = b.lo[a] > world_lo(a) + SP_EPSILON_FLOAT;
        for (double v = b.lo[a]; v < b.hi[a] + SP_EPSILON_FLOAT; v += axis_res(a))
            out.push_back((out.empty() && nudge_first_down) ? b.lo[a] - SP_EPSILON_FLOAT : v + SP_EPSILON_FLOAT);
    }
    // true when every sample of the box keeps `margin` from the obstacles
    bool clear_of_obstacles(const Aabb &b, double margin) {
        axis_samples(b, 0, sx_); axis_samples(b, 1, sy_); axis_samples(b, 2, sz_);
        for (double x : sx_)
            for (double y : sy_)
                for (double z : sz_)
                    if (distmap_obj->getDistance(octomap::point3d((float)x, (float)y, (float)z)) < margin - SP_EPSILON_FLOAT) return false;
        return true;
    }
    bool inside_world(const Aabb &b) const {
        for (int a = 0; a < 3; a++)
            if (!(b.lo[a] > world_lo(a) - SP_EPSILON && b.hi[a] < world_hi(a) + SP_EPSILON)) return false;
        return true;
    }
    static bool holds(const Aabb &b, const octomap::point3d &p) {
        for (int a = 0; a < 3; a++)
            if (!(p(a) > b.lo[a] - SP_EPSILON && p(a) < b.hi[a] + SP_EPSILON)) return false;
        return true;
    }
    // one resolution step of face f (0..2: lower faces, 3..5: upper faces): `grown` = box + slab, `slab` = the new layer only
    void push_face(const Aabb &box, int f, Aabb &grown, Aabb &slab) const {
        grown = box; slab = box;
        const int a = f % 3;
        if (f < 3) { slab.hi[a] = box.lo[a]; grown.lo[a] = box.lo[a] - axis_res(a); slab.lo[a] = grown.lo[a]; }
        else       { slab.lo[a] = box.hi[a]; grown.hi[a] = box.hi[a] + axis_res(a); slab.hi[a] = grown.hi[a]; }
    }
    void grow(Aabb &box, double margin) {
        std::vector<int> active{0, 1, 2, 3, 4, 5};   // faces still allowed to move, ring order
        int at = -1;                                  // ring position of the face moved last
        while (!active.empty()) {
            Aabb grown = box, probe = box;            // after a retirement the whole box is probed once more
            while (clear_of_obstacles(probe, margin) && inside_world(probe)) {
                at = (at + 1 >= (int)active.size()) ? 0 : at + 1;
                box = grown;                          // the step probed last is accepted
                push_face(box, active[at], grown, probe);
            }
            if (at < 0) at = 0;                       // seed outside the world: nothing moved yet (the reference erases begin() - 1 here)
            active.erase(active.begin() + at);       // the face whose step failed retires
            at = (at > 0) ? at - 1 : (int)active.size() - 1;
        }
    }

    bool updateObsBox() {
        PlanResult &pr = *planResult_ptr;
        pr.SFC.assign(mission.qn, {});
        for (int qi = 0; qi < mission.qn; ++qi) {
            const auto &path = pr.initTraj[qi];
            const int npts = (int)path.size();
            std::vector<Aabb> boxes;
            // 1. one box per path edge that leaves the previous box
            for (int j = 0; j + 1 < npts; j++) {
                if (!boxes.empty() ? holds(boxes.back(), path[j + 1]) : holds(Aabb{{0, 0, 0}, {0, 0, 0}}, path[j + 1])) continue;
                Aabb seed;
                for (int a = 0; a < 3; a++) {
                    const double u = path[j](a), v = path[j + 1](a), r = axis_res(a);
                    seed.lo[a] = std::round(std::min(u, v) / r) * r;
                    seed.hi[a] = std::round(std::max(u, v) / r) * r;
                }
                if (!clear_of_obstacles(seed, mission.quad_size[qi])) {
                    std::fprintf(stderr, "Corridor: Invalid initial trajectory. Obstacle invades initial trajectory.\n");
                    return false;
                }
                grow(seed, mission.quad_size[qi]);
                boxes.push_back(seed);
            }
            const int nbox = (int)boxes.size();
            if (nbox == 0) {
                std::fprintf(stderr, "Corridor: agent %d has no corridor box (degenerate initial trajectory)\n", qi);
                return false;
            }
            // 2. hand-over times.  in_box[i][j]: path point j lies in box i (the reference keeps run lengths in `box_log`
            // but only ever asks whether they are positive)
            std::vector<std::vector<char>> in_box(nbox, std::vector<char>(npts, 0));
            for (int i = 0; i < nbox; i++)
                for (int j = 0; j < npts; j++) in_box[i][j] = holds(boxes[i], path[j]) ? 1 : 0;
            std::vector<double> t_end(nbox, -1.0);
            int cur = 0;   // box being timed
            for (int j = 0; j < npts; j++) {
                if (cur == nbox - 1) {
                    if (in_box[cur][j]) continue;
                    cur--;
                }
                if (cur < 0) {   // the reference reads box_log(-1, .) here (undefined behaviour); reported instead
                    std::fprintf(stderr, "Corridor: box/time allocation ran off the first box (agent %d)\n", qi);
                    return false;
                }
                if (in_box[cur][j] && in_box[cur + 1][j]) {
                    int shared = 1;   // length of the stretch of points lying in both boxes
                    while (j + shared < npts && in_box[cur][j + shared] && in_box[cur + 1][j + shared]) shared++;
                    const int mid = j + shared / 2;
                    t_end[cur] = pr.T[mid];
                    j = mid;
                    cur++;
                } else if (!in_box[cur][j]) {   // fell out of the box: step back one box and look at this point again
                    cur--;
                    j--;
                    if (cur < 0 || j < -1) {
                        std::fprintf(stderr, "Corridor: box/time allocation ran off the first box (agent %d)\n", qi);
                        return false;
                    }
                }
            }
            t_end[nbox - 1] = makespan;
            for (int i = 0; i < nbox; i++) pr.SFC[qi].emplace_back(std::make_pair(boxes[i].as_sfc(), t_end[i]));
        }
        return true;
    }
    std::vector<double> sx_, sy_, sz_;   // scratch of clear_of_obstacles

    bool updateRelBox() {   // L338-L398, arithmetic on the device (bit-identical float32 semantics)
        PlanResult &pr = *planResult_ptr;
        const int N = mission.qn, M = (int)pr.T.size() - 1;
        for (int qi = 0; qi < N; qi++)
            if ((int)pr.initTraj[qi].size() != M + 1) {
                std::fprintf(stderr, "Corridor: size of initial trajectories must be equal\n");
                return false;
            }
        if (!engine) {
            rbpe_config cfg = {};
            if (rbpe_create(&cfg, &engine) != RBPE_OK) {
                std::fprintf(stderr, "Corridor: %s\n", rbpe_last_error(nullptr));
                return false;
            }
        }
        const size_t P = (size_t)N * (N - 1) / 2;
        std::vector<float> traj((size_t)N * (M + 1) * 3), rn((P ? P : 1) * M * 3);
        std::vector<double> rt((P ? P : 1) * M);
        for (int qi = 0; qi < N; qi++)
            for (int j = 0; j <= M; j++)
                for (int k = 0; k < 3; k++) traj[((size_t)qi * (M + 1) + j) * 3 + k] = pr.initTraj[qi][j](k);
        int collided = 0;
        if (rbpe_corridor_rsfc(engine, N, M, 1, traj.data(), pr.T.data(), param.downwash, rn.data(), rt.data(), &collided) != RBPE_OK) {
            std::fprintf(stderr, "Corridor: %s\n", rbpe_last_error(engine));
            return false;
        }
        if (collided) {
            std::fprintf(stderr, "Corridor: initial trajectories are collided with each other\n");
            return false;
        }
        pr.RSFC.assign(N, {});
        size_t it = 0;
        for (int qi = 0; qi < N; qi++) {
            pr.RSFC[qi].resize(N);
            for (int qj = qi + 1; qj < N; qj++, it++)
                for (int ri = 0; ri < M; ri++)
                    pr.RSFC[qi][qj].emplace_back(std::make_pair(
                        octomap::point3d(rn[(it * M + ri) * 3], rn[(it * M + ri) * 3 + 1], rn[(it * M + ri) * 3 + 2]), rt[it * M + ri]));
        }
        return true;
    }
};

#ifdef RBPE_HAVE_ROS
using Corridor = CorridorT<DynamicEDTOctomap>;
#else
using Corridor = CorridorT<GridDistanceMap>;
#endif
}  // namespace SwarmPlanning
