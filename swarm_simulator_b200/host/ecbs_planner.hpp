// ecbs_planner.hpp -- host-side mirror of the reference's initial-trajectory stage:
//   SwarmPlanning::InitTrajPlanner   /root/reference/swarm_planner/include/init_traj_planner.hpp L9-L38  (grid set-up)
//   SwarmPlanning::ECBSPlanner       /root/reference/swarm_planner/include/ecbs_planner.hpp L9-L137      (obstacles, waypoints,
//                                    T = 0..makespan+2, initTraj = [start, grid states, goal padding])
// and a from-scratch restatement of the vendored search it calls (third_party/ecbs, which needs Boost -- absent here):
//   ECBS high level                  third_party/ecbs/include/ecbs.hpp L108-L262 (focal search over constraint-tree nodes)
//   focal A* low level               third_party/ecbs/include/a_star_epsilon.hpp L84-L277
//   Environment                      third_party/ecbs/include/environment.hpp L347-L681 (7-connected moves incl. wait, radius-
//                                    aware vertex / edge conflicts L656-L681, first conflict L529-L590, constraints L592-L612)
// Same class shape as the reference: ECBSPlanner(distmap, mission, param) + bool update(bool log, PlanResult*).
// Stays on the host (north-star).  The reference's heaps are boost d-ary heaps whose tie-breaking among equal keys is an
// implementation detail; here ties break by insertion order, so paths of equal cost may differ: parity UNPINNED (no golden
// initTraj in the reference either); tests check validity (conflict-free by the rules above, costs >= the individual
// shortest paths, sub-optimality bound on small cases).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <map>
#include <memory>
#include <set>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <mission.hpp>
#include <param.hpp>
#include <sp_const.hpp>

namespace SwarmPlanning {
namespace ecbs {

struct State { int time, x, y, z; };
inline bool same_cell(const State &a, const State &b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline long long key3(int x, int y, int z) { return ((long long)(x + 4096) << 32) | ((long long)(y + 4096) << 16) | (long long)(z + 4096); }
inline long long key4(const State &s) { return ((long long)s.time << 42) ^ key3(s.x, s.y, s.z); }

struct Constraints {
    std::set<std::tuple<int, int, int, int>> vertex;                       // (time, x, y, z)
    std::set<std::tuple<int, int, int, int, int, int, int>> edge;          // (time, x1, y1, z1, x2, y2, z2)
};
struct Path { std::vector<State> states; int cost = 0, fmin = 0; };
struct Conflict { int time, agent1, agent2; bool is_edge; State a1, b1, a2, b2; };

class Environment {
public:
    Environment(int dimx, int dimy, int dimz, std::unordered_set<long long> obstacles, std::vector<State> goals,
                std::vector<double> quad_size, double grid_size)
        : dimx_(dimx), dimy_(dimy), dimz_(dimz), obstacles_(std::move(obstacles)), goals_(std::move(goals)),
          quad_size_(std::move(quad_size)), grid_size_(grid_size) {}

    static State state_at(const Path &p, int t) { return t < (int)p.states.size() ? p.states[t] : p.states.back(); }

    // environment.hpp L656-L664
    bool vertex_conflict(int i, int j, const State &s1, const State &s2) const {
        if (quad_size_[i] + quad_size_[j] < grid_size_) return same_cell(s1, s2);
        const double dx = s2.x - s1.x, dy = s2.y - s1.y, dz = s2.z - s1.z;
        return std::sqrt(dx * dx + dy * dy + dz * dz) * grid_size_ < quad_size_[i] + quad_size_[j];
    }
    // environment.hpp L666-L681 with Vector::min_dist_to_origin L69-L93
    bool edge_conflict(int i, int j, const State &s1a, const State &s1b, const State &s2a, const State &s2b) const {
        if (quad_size_[i] + quad_size_[j] < grid_size_ * 0.5) return same_cell(s1a, s2b) && same_cell(s1b, s2a);
        const double a[3] = {(double)(s2a.x - s1a.x), (double)(s2a.y - s1a.y), (double)(s2a.z - s1a.z)};
        const double b[3] = {(double)(s2b.x - s1b.x), (double)(s2b.y - s1b.y), (double)(s2b.z - s1b.z)};
        auto norm = [](const double *v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
        double min_dist = norm(a);
        if (!(a[0] == b[0] && a[1] == b[1] && a[2] == b[2])) {
            double dist = norm(b);
            if (min_dist > dist) min_dist = dist;
            double n[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
            const double len = norm(n);
            for (double &v : n) v /= len;
            const double adn = a[0] * n[0] + a[1] * n[1] + a[2] * n[2];
            const double c[3] = {a[0] - n[0] * adn, a[1] - n[1] * adn, a[2] - n[2] * adn};
            dist = norm(c);
            const double dot = (c[0] - a[0]) * (c[0] - b[0]) + (c[1] - a[1]) * (c[1] - b[1]) + (c[2] - a[2]) * (c[2] - b[2]);
            if (dot < 0 && min_dist > dist) min_dist = dist;
        }
        return min_dist * grid_size_ <= quad_size_[i] + quad_size_[j];
    }

    // environment.hpp L529-L590 (count_only: focalHeuristic L425-L458)
    int scan_conflicts(const std::vector<Path> &sol, Conflict *first) const {
        int count = 0, max_t = 0;
        for (const Path &p : sol) max_t = std::max<int>(max_t, (int)p.states.size() - 1);
        const int n = (int)sol.size();
        for (int t = 0; t < max_t; t++) {
            for (int i = 0; i < n; i++) {
                const State s1 = state_at(sol[i], t);
                for (int j = i + 1; j < n; j++) {
                    const State s2 = state_at(sol[j], t);
                    if (vertex_conflict(i, j, s1, s2)) {
                        if (first) { *first = Conflict{t, i, j, false, s1, s1, s2, s2}; return 1; }
                        count++;
                    }
                }
            }
            for (int i = 0; i < n; i++) {
                const State s1a = state_at(sol[i], t), s1b = state_at(sol[i], t + 1);
                for (int j = i + 1; j < n; j++) {
                    const State s2a = state_at(sol[j], t), s2b = state_at(sol[j], t + 1);
                    if (edge_conflict(i, j, s1a, s1b, s2a, s2b)) {
                        if (first) { *first = Conflict{t, i, j, true, s1a, s1b, s2a, s2b}; return 1; }
                        count++;
                    }
                }
            }
        }
        return count;
    }

    // focal A* (a_star_epsilon.hpp L84-L277) for one agent under its constraints, other agents' paths in `sol`
    bool low_level(int agent, const State &start, const Constraints &con, const std::vector<Path> &sol, double w, Path &out,
                   long max_expansions = 4000000) const {
        struct Node { State s; int f, g, focal; long long id; };
        typedef std::tuple<int, int, long long> OKey;            // (f, -g, id)
        typedef std::tuple<int, int, int, long long> FKey;       // (focal, f, -g, id)
        std::map<OKey, Node> open;
        std::set<FKey> focal;
        std::unordered_map<long long, OKey> in_open;             // state -> key in open
        std::unordered_set<long long> closed;
        std::unordered_map<long long, std::pair<State, int>> came_from;   // state -> (parent, g)
        const State goal = goals_[agent];
        int last_goal_constraint = -1;
        for (const auto &vc : con.vertex)
            if (std::get<1>(vc) == goal.x && std::get<2>(vc) == goal.y && std::get<3>(vc) == goal.z)
                last_goal_constraint = std::max(last_goal_constraint, std::get<0>(vc));
        auto h = [&](const State &s) { return std::abs(s.x - goal.x) + std::abs(s.y - goal.y) + std::abs(s.z - goal.z); };
        long long next_id = 0;
        {
            Node n{start, h(start), 0, 0, next_id++};
            OKey k(n.f, -n.g, n.id);
            open.emplace(k, n);
            in_open[key4(start)] = k;
            focal.insert(FKey(0, n.f, -n.g, n.id));
        }
        int best_f = h(start);
        long expansions = 0;
        static const int moves[7][3] = {{0, 0, 0}, {-1, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
        while (!open.empty()) {
            {
                const int old_best = best_f;
                best_f = open.begin()->second.f;
                if (best_f > old_best)
                    for (const auto &kv : open) {
                        const int val = kv.second.f;
                        if (val > old_best * w && val <= best_f * w) focal.insert(FKey(kv.second.focal, kv.second.f, -kv.second.g, kv.second.id));
                        if (val > best_f * w) break;
                    }
            }
            const FKey fk = *focal.begin();
            const OKey ok(std::get<1>(fk), std::get<2>(fk), std::get<3>(fk));
            const Node cur = open.at(ok);
            if (same_cell(cur.s, goal) && cur.s.time > last_goal_constraint) {
                out.states.clear();
                State s = cur.s;
                for (;;) {
                    out.states.push_back(s);
                    auto it = came_from.find(key4(s));
                    if (it == came_from.end()) break;
                    s = it->second.first;
                }
                std::reverse(out.states.begin(), out.states.end());
                out.cost = cur.g;
                out.fmin = open.begin()->second.f;
                return true;
            }
            focal.erase(focal.begin());
            open.erase(ok);
            in_open.erase(key4(cur.s));
            closed.insert(key4(cur.s));
            if (++expansions > max_expansions) return false;
            for (const auto &mv : moves) {
                const State nb{cur.s.time + 1, cur.s.x + mv[0], cur.s.y + mv[1], cur.s.z + mv[2]};
                if (nb.x < 0 || nb.x >= dimx_ || nb.y < 0 || nb.y >= dimy_ || nb.z < 0 || nb.z >= dimz_) continue;
                if (obstacles_.count(key3(nb.x, nb.y, nb.z))) continue;
                if (con.vertex.count(std::make_tuple(nb.time, nb.x, nb.y, nb.z))) continue;
                if (con.edge.count(std::make_tuple(cur.s.time, cur.s.x, cur.s.y, cur.s.z, nb.x, nb.y, nb.z))) continue;
                const long long nk = key4(nb);
                if (closed.count(nk)) continue;
                const int tentative_g = cur.g + 1;
                auto it = in_open.find(nk);
                if (it == in_open.end()) {
                    int conflicts = 0;   // focalStateHeuristic + focalTransitionHeuristic (environment.hpp L393-L423)
                    for (int i = 0; i < (int)sol.size(); i++) {
                        if (i == agent || sol[i].states.empty()) continue;
                        const State o_a = state_at(sol[i], cur.s.time), o_b = state_at(sol[i], nb.time);
                        if (vertex_conflict(agent, i, nb, o_b)) conflicts++;
                        if (edge_conflict(agent, i, cur.s, nb, o_a, o_b)) conflicts++;
                    }
                    Node n{nb, tentative_g + h(nb), tentative_g, cur.focal + conflicts, next_id++};
                    OKey k(n.f, -n.g, n.id);
                    open.emplace(k, n);
                    in_open[nk] = k;
                    if (n.f <= best_f * w) focal.insert(FKey(n.focal, n.f, -n.g, n.id));
                } else {
                    Node n = open.at(it->second);
                    if (tentative_g >= n.g) continue;
                    const int last_f = n.f;
                    focal.erase(FKey(n.focal, n.f, -n.g, n.id));
                    open.erase(it->second);
                    n.f -= n.g - tentative_g;
                    n.g = tentative_g;
                    OKey k(n.f, -n.g, n.id);
                    open.emplace(k, n);
                    it->second = k;
                    if (n.f <= best_f * w || last_f <= best_f * w) focal.insert(FKey(n.focal, n.f, -n.g, n.id));
                }
                came_from[nk] = std::make_pair(cur.s, tentative_g);
            }
        }
        return false;
    }

    // ECBS high level (ecbs.hpp L108-L262)
    bool search(const std::vector<State> &starts, double w, std::vector<Path> &solution, bool log, long max_nodes = 200000) {
        struct HL { std::vector<Path> sol; std::vector<Constraints> con; int cost, lb, focal; long long id; };
        typedef std::tuple<int, long long> OKey;            // (cost, id)
        typedef std::tuple<int, int, long long> FKey;       // (focal, cost, id)
        std::map<OKey, HL> open;
        std::set<FKey> focal;
        const int n = (int)starts.size();
        HL root;
        root.sol.resize(n); root.con.resize(n); root.cost = 0; root.lb = 0; root.id = 0;
        for (int i = 0; i < n; i++) {
            if (!low_level(i, starts[i], root.con[i], root.sol, w, root.sol[i])) return false;
            root.cost += root.sol[i].cost;
            root.lb += root.sol[i].fmin;
        }
        root.focal = scan_conflicts(root.sol, nullptr);
        open.emplace(OKey(root.cost, 0), root);
        focal.insert(FKey(root.focal, root.cost, 0));
        int best_cost = root.cost;
        long long id = 1;
        high_level_expanded = 0;
        while (!open.empty()) {
            {
                const int old_best = best_cost;
                best_cost = open.begin()->second.cost;
                if (best_cost > old_best)
                    for (const auto &kv : open) {
                        const int val = kv.second.cost;
                        if (val > old_best * w && val <= best_cost * w) focal.insert(FKey(kv.second.focal, val, kv.second.id));
                        if (val > best_cost * w) break;
                    }
            }
            const FKey fk = *focal.begin();
            focal.erase(focal.begin());
            const OKey ok(std::get<1>(fk), std::get<2>(fk));
            HL P = std::move(open.at(ok));
            open.erase(ok);
            high_level_expanded++;
            Conflict c;
            if (!scan_conflicts(P.sol, &c)) {
                if (log) std::fprintf(stderr, "ECBS: done; cost %d, %ld high-level nodes\n", P.cost, high_level_expanded);
                solution = std::move(P.sol);
                return true;
            }
            if (high_level_expanded > max_nodes) return false;
            for (int side = 0; side < 2; side++) {
                const int i = side == 0 ? c.agent1 : c.agent2;
                const State &sa = side == 0 ? c.a1 : c.a2, &sb = side == 0 ? c.b1 : c.b2;
                HL child = P;
                child.id = id++;
                if (c.is_edge) child.con[i].edge.insert(std::make_tuple(c.time, sa.x, sa.y, sa.z, sb.x, sb.y, sb.z));
                else child.con[i].vertex.insert(std::make_tuple(c.time, sa.x, sa.y, sa.z));
                child.cost -= child.sol[i].cost;
                child.lb -= child.sol[i].fmin;
                const bool ok_ll = low_level(i, starts[i], child.con[i], child.sol, w, child.sol[i]);
                if (!ok_ll) continue;
                child.cost += child.sol[i].cost;
                child.lb += child.sol[i].fmin;
                child.focal = scan_conflicts(child.sol, nullptr);
                const int cc = child.cost, cf = child.focal;
                const long long cid = child.id;
                open.emplace(OKey(cc, cid), std::move(child));
                if (cc <= best_cost * w) focal.insert(FKey(cf, cc, cid));
            }
        }
        return false;
    }

    long high_level_expanded = 0;

private:
    int dimx_, dimy_, dimz_;
    std::unordered_set<long long> obstacles_;
    std::vector<State> goals_;
    std::vector<double> quad_size_;
    double grid_size_;
};

}  // namespace ecbs

// Planner stage with the reference's constructor / update() shape.  Grid conventions (init_traj_planner.hpp L17-L30): the
// planning lattice is the set of multiples of the grid pitch inside the world box; lattice index = round((coord - min) / pitch).
template <class DistMap>
class ECBSPlannerT {
public:
    ECBSPlannerT(std::shared_ptr<DistMap> distmap, Mission mission_in, Param param_in)
        : map_(std::move(distmap)), mission(std::move(mission_in)), param(std::move(param_in)) {
        const double lo[3] = {param.world_x_min, param.world_y_min, param.world_z_min};
        const double hi[3] = {param.world_x_max, param.world_y_max, param.world_z_max};
        for (int a = 0; a < 3; a++) {
            pitch_[a] = (a < 2) ? param.grid_xy_res : param.grid_z_res;
            gmin_[a] = std::ceil((lo[a] - SP_EPSILON) / pitch_[a]) * pitch_[a];
            gmax_[a] = std::floor((hi[a] + SP_EPSILON) / pitch_[a]) * pitch_[a];
            dim_[a] = (int)std::round((gmax_[a] - gmin_[a]) / pitch_[a]) + 1;
        }
        valid = mark_obstacles() && place_agents();
    }

    // ecbs_planner.hpp L21-L74: search, then T = {0, dt, ..., (makespan + 2) dt} and initTraj = exact start, lattice states,
    // exact goal repeated up to makespan + 2 points after the start
    bool update(bool log, SwarmPlanning::PlanResult *out) {
        if (!valid) return false;
        ecbs::Environment env(dim_[0], dim_[1], dim_[2], blocked_, goals_, mission.quad_size, param.grid_xy_res);
        std::vector<ecbs::Path> paths;
        if (!env.search(starts_, param.ecbs_w, paths, log)) {
            std::fprintf(stderr, "ECBSPlanner: ECBS Failed!\n");
            return false;
        }
        high_level_expanded = env.high_level_expanded;
        int makespan = 0;
        for (const ecbs::Path &p : paths) makespan = std::max(makespan, p.cost);
        const int npts = makespan + 3;
        out->T.resize(npts);
        for (int i = 0; i < npts; i++) out->T[i] = i * param.time_step;
        out->initTraj.assign(paths.size(), {});
        for (size_t q = 0; q < paths.size(); q++) {
            auto &traj = out->initTraj[q];
            traj.reserve(npts);
            traj.emplace_back(mission.startState[q][0], mission.startState[q][1], mission.startState[q][2]);
            for (const ecbs::State &st : paths[q].states)
                traj.emplace_back(st.x * pitch_[0] + gmin_[0], st.y * pitch_[1] + gmin_[1], st.z * pitch_[2] + gmin_[2]);
            const octomap::point3d goal(mission.goalState[q][0], mission.goalState[q][1], mission.goalState[q][2]);
            while ((int)traj.size() < npts) traj.push_back(goal);
        }
        return true;
    }

    long high_level_expanded = 0;
    bool valid = false;

private:
    std::shared_ptr<DistMap> map_;
    Mission mission;
    Param param;
    double pitch_[3], gmin_[3], gmax_[3];
    int dim_[3];
    std::unordered_set<long long> blocked_;
    std::vector<ecbs::State> starts_, goals_;

    int cell_of(double coord, int a) const { return (int)std::round((coord - gmin_[a]) / pitch_[a]); }

    // ecbs_planner.hpp L81-L110: a lattice point is blocked when the distance map reads less than the largest radius plus
    // grid_margin there.  The reference walks the lattice with running double sums (k += res); the same sums are formed here
    // so that the points handed to getDistance are the same floats.
    bool mark_obstacles() {
        double rmax = 0;
        for (double r : mission.quad_size) rmax = std::max(rmax, r);
        const double clearance = rmax + param.grid_margin;
        for (double z = gmin_[2]; z < gmax_[2] + SP_EPSILON; z += pitch_[2])
            for (double x = gmin_[0]; x < gmax_[0] + SP_EPSILON; x += pitch_[0])
                for (double y = gmin_[1]; y < gmax_[1] + SP_EPSILON; y += pitch_[1]) {
                    const float d = map_->getDistance(octomap::point3d(x, y, z));
                    if (d < 0) return false;                       // outside the distance map
                    if (d < clearance) blocked_.insert(ecbs::key3(cell_of(x, 0), cell_of(y, 1), cell_of(z, 2)));
                }
        return true;
    }

    // ecbs_planner.hpp L113-L136: nearest lattice point of every start / goal; a blocked one is an error
    bool place_agents() {
        for (int q = 0; q < mission.qn; q++) {
            ecbs::State s{0, 0, 0, 0}, g{0, 0, 0, 0};
            int *sc[3] = {&s.x, &s.y, &s.z}, *gc[3] = {&g.x, &g.y, &g.z};
            for (int a = 0; a < 3; a++) {
                *sc[a] = cell_of(mission.startState[q][a], a);
                *gc[a] = cell_of(mission.goalState[q][a], a);
            }
            if (blocked_.count(ecbs::key3(s.x, s.y, s.z))) {
                std::fprintf(stderr, "ECBSPlanner: start of agent %d is occluded by obstacle\n", q);
                return false;
            }
            if (blocked_.count(ecbs::key3(g.x, g.y, g.z))) {
                std::fprintf(stderr, "ECBSPlanner: goal of agent %d is occluded by obstacle\n", q);
                return false;
            }
            starts_.push_back(s);
            goals_.push_back(g);
        }
        return true;
    }
};

#ifndef RBPE_HAVE_ROS
class GridDistanceMap;
using ECBSPlanner = ECBSPlannerT<GridDistanceMap>;
#endif

}  // namespace SwarmPlanning
