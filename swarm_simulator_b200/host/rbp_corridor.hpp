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

private:
    std::shared_ptr<DistMap> distmap_obj;
    Mission mission;
    Param param;
    bool log = false;
    SwarmPlanning::PlanResult *planResult_ptr = nullptr;
    double makespan = 0;
    rbpe_handle *engine = nullptr;

    bool isObstacleInBox(const std::vector<double> &box, double margin) {   // L44-L78
        double x, y, z;
        int count1 = 0;
        for (double i = box[0]; i < box[3] + SP_EPSILON_FLOAT; i += param.box_xy_res) {
            int count2 = 0;
            for (double j = box[1]; j < box[4] + SP_EPSILON_FLOAT; j += param.box_xy_res) {
                int count3 = 0;
                for (double k = box[2]; k < box[5] + SP_EPSILON_FLOAT; k += param.box_z_res) {
                    x = i + SP_EPSILON_FLOAT;
                    if (count1 == 0 && box[0] > param.world_x_min + SP_EPSILON_FLOAT) x = box[0] - SP_EPSILON_FLOAT;
                    y = j + SP_EPSILON_FLOAT;
                    if (count2 == 0 && box[1] > param.world_y_min + SP_EPSILON_FLOAT) y = box[1] - SP_EPSILON_FLOAT;
                    z = k + SP_EPSILON_FLOAT;
                    if (count3 == 0 && box[2] > param.world_z_min + SP_EPSILON_FLOAT) z = box[2] - SP_EPSILON_FLOAT;
                    octomap::point3d cur_point((float)x, (float)y, (float)z);
                    float dist = distmap_obj->getDistance(cur_point);
                    if (dist < margin - SP_EPSILON_FLOAT) return true;
                    count3++;
                }
                count2++;
            }
            count1++;
        }
        return false;
    }
    bool isBoxInBoundary(const std::vector<double> &box) {   // L80-L87
        return box[0] > param.world_x_min - SP_EPSILON && box[1] > param.world_y_min - SP_EPSILON &&
               box[2] > param.world_z_min - SP_EPSILON && box[3] < param.world_x_max + SP_EPSILON &&
               box[4] < param.world_y_max + SP_EPSILON && box[5] < param.world_z_max + SP_EPSILON;
    }
    static bool isPointInBox(const octomap::point3d &point, const std::vector<double> &box) {   // L89-L97
        return point.x() > box[0] - SP_EPSILON && point.y() > box[1] - SP_EPSILON && point.z() > box[2] - SP_EPSILON &&
               point.x() < box[3] + SP_EPSILON && point.y() < box[4] + SP_EPSILON && point.z() < box[5] + SP_EPSILON;
    }
    void expand_box(std::vector<double> &box, double margin) {   // L99-L147
        std::vector<double> box_cand, box_update;
        std::vector<int> axis_cand{0, 1, 2, 3, 4, 5};
        int i = -1, axis;
        while (!axis_cand.empty()) {
            box_cand = box;
            box_update = box;
            // only the newly added slab is tested: update_box + current_box = cand_box
            while (!isObstacleInBox(box_update, margin) && isBoxInBoundary(box_update)) {
                i++;
                if (i >= (int)axis_cand.size()) i = 0;
                axis = axis_cand[i];
                box = box_cand;
                box_update = box_cand;
                if (axis < 3) {
                    box_update[axis + 3] = box_cand[axis];
                    box_cand[axis] = box_cand[axis] - (axis == 2 ? param.box_z_res : param.box_xy_res);
                    box_update[axis] = box_cand[axis];
                } else {
                    box_update[axis - 3] = box_cand[axis];
                    box_cand[axis] = box_cand[axis] + (axis == 5 ? param.box_z_res : param.box_xy_res);
                    box_update[axis] = box_cand[axis];
                }
            }
            axis_cand.erase(axis_cand.begin() + i);
            if (i > 0) i--;
            else i = (int)axis_cand.size() - 1;
        }
    }

    bool updateObsBox() {   // L149-L243
        PlanResult &pr = *planResult_ptr;
        pr.SFC.assign(mission.qn, {});
        for (int qi = 0; qi < mission.qn; ++qi) {
            std::vector<double> box_prev{0, 0, 0, 0, 0, 0};
            for (int i = 0; i + 1 < (int)pr.initTraj[qi].size(); i++) {
                octomap::point3d state = pr.initTraj[qi][i], state_next = pr.initTraj[qi][i + 1];
                double x = state.x(), y = state.y(), z = state.z();
                double x_next = state_next.x(), y_next = state_next.y(), z_next = state_next.z();
                if (isPointInBox(octomap::point3d((float)x_next, (float)y_next, (float)z_next), box_prev)) continue;
                std::vector<double> box;
                box.emplace_back(std::round(std::min(x, x_next) / param.box_xy_res) * param.box_xy_res);
                box.emplace_back(std::round(std::min(y, y_next) / param.box_xy_res) * param.box_xy_res);
                box.emplace_back(std::round(std::min(z, z_next) / param.box_z_res) * param.box_z_res);
                box.emplace_back(std::round(std::max(x, x_next) / param.box_xy_res) * param.box_xy_res);
                box.emplace_back(std::round(std::max(y, y_next) / param.box_xy_res) * param.box_xy_res);
                box.emplace_back(std::round(std::max(z, z_next) / param.box_z_res) * param.box_z_res);
                if (isObstacleInBox(box, mission.quad_size[qi])) {
                    std::fprintf(stderr, "Corridor: Invalid initial trajectory. Obstacle invades initial trajectory.\n");
                    return false;
                }
                expand_box(box, mission.quad_size[qi]);
                pr.SFC[qi].emplace_back(std::make_pair(box, -1));
                box_prev = box;
            }
            // box <-> time allocation through the run-length table box_log (L195-L237)
            int box_max = (int)pr.SFC[qi].size(), path_max = (int)pr.initTraj[qi].size();
            std::vector<double> box_log((size_t)box_max * path_max, 0.0);
            auto BL = [&](int i, int j) -> double & { return box_log[(size_t)i * path_max + j]; };
            for (int i = 0; i < box_max; i++)
                for (int j = 0; j < path_max; j++)
                    if (isPointInBox(pr.initTraj[qi][j], pr.SFC[qi][i].first)) BL(i, j) = (j == 0) ? 1 : BL(i, j - 1) + 1;
            int box_iter = 0;
            for (int path_iter = 0; path_iter < path_max; path_iter++) {
                if (box_iter == box_max - 1) {
                    if (BL(box_iter, path_iter) > 0) continue;
                    else box_iter--;
                }
                if (box_iter < 0) return false;   // the reference indexes box_log(-1, .) here (undefined behaviour)
                if (BL(box_iter, path_iter) > 0 && BL(box_iter + 1, path_iter) > 0) {
                    int count = 1;
                    while (path_iter + count < path_max && BL(box_iter, path_iter + count) > 0 &&
                           BL(box_iter + 1, path_iter + count) > 0)
                        count++;
                    int obs_index = path_iter + count / 2;
                    pr.SFC[qi][box_iter].second = pr.T[obs_index];
                    path_iter = path_iter + count / 2;
                    box_iter++;
                } else if (BL(box_iter, path_iter) == 0) {
                    box_iter--;
                    path_iter--;
                    if (box_iter < 0 || path_iter < -1) return false;
                }
            }
            pr.SFC[qi][box_max - 1].second = makespan;
        }
        return true;
    }

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
