// octree_bt.hpp -- reader for octomap's binary ".bt" occupancy trees and the clamped Euclidean distance map the planner
// node builds from them.  Host stage in front of the hot path: it replaces, where octomap /
// dynamicEDT3D are absent,
//   octomap::OcTree::readBinary            (what `octomap_server` loads from swarm_planner/worlds/*.bt), and
//   DynamicEDTOctomap(maxDist = 1, octree, world_min, world_max, unknownAsOccupied = false) + update()
//     /root/reference/swarm_planner/src/swarm_traj_planner_rbp.cpp L76-L81.
// Both libraries are third-party dependencies of the reference that are not under /root/reference; their file format and
// semantics are restated from the published octomap / dynamicEDT3D sources [recalled]:
//   * header: text lines up to "data" ("id OcTree", "size <nodes>", "res <metres>");
//   * body: depth-first from the root, two bytes per inner node = 2 bits per child, children 0..3 in the first byte from
//     the least significant bits: (bit 2i, bit 2i+1) = (0,0) unknown, (0,1) occupied leaf, (1,0) free leaf, (1,1) child with
//     children; after the two bytes of a node the children flagged (1,1) follow recursively in index order;
//   * child i sits at +x if (i & 1), +y if (i & 2), +z if (i & 4); the root is centred at the origin and 2^16 cells wide;
//   * the distance map lives on the cells floor(coord / res) between the bounding-box keys, distances are Euclidean
//     between cell centres, clamped at maxDist; unknown space counts as free; outside the box getDistance returns -1.
// Parity for this file is unpinned (no golden distance map in the reference); tests check structural invariants on the
// reference's own worlds (pillar count, bounds) and the distance transform against brute force.
#pragma once

#include <cmath>
#include <cstdint>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace SwarmPlanning {

class OcTreeBt {
public:
    struct Leaf { double cx, cy, cz, size; bool occupied; };
    double res = 0.1;
    size_t declared_nodes = 0, inner_nodes = 0;
    std::vector<Leaf> leaves;
    std::string error;

    bool load(const std::string &path) {
        std::ifstream in(path, std::ios::binary);
        if (!in) { error = "cannot open " + path; return false; }
        std::string line;
        bool have_data = false, is_octree = false;
        while (std::getline(in, line)) {
            if (line.empty() || line[0] == '#') continue;
            std::istringstream ls(line);
            std::string key;
            ls >> key;
            if (key == "id") { std::string id; ls >> id; is_octree = (id == "OcTree"); }
            else if (key == "size") ls >> declared_nodes;
            else if (key == "res") ls >> res;
            else if (key == "data") { have_data = true; break; }
        }
        if (!have_data || !is_octree || !(res > 0)) { error = "not an OcTree .bt file: " + path; return false; }
        leaves.clear();
        inner_nodes = 0;
        if (declared_nodes == 0) return true;   // empty tree: no root node in the stream
        return read_node(in, 0.0, 0.0, 0.0, res * 65536.0, 0);
    }

    // occupancy on the cells floor(coord / res) of the box [lo, hi] (inclusive keys, as DynamicEDTOctomap's bounding box)
    void rasterize(const double lo[3], const double hi[3], int k0[3], int n[3], std::vector<unsigned char> &occ) const {
        for (int a = 0; a < 3; a++) {
            k0[a] = (int)std::floor(lo[a] / res + 1e-9);
            n[a] = (int)std::floor(hi[a] / res + 1e-9) - k0[a] + 1;
        }
        occ.assign((size_t)n[0] * n[1] * n[2], 0);
        for (const Leaf &l : leaves) {
            if (!l.occupied) continue;
            const double c[3] = {l.cx, l.cy, l.cz};
            int b0[3], b1[3];
            bool out = false;
            for (int a = 0; a < 3; a++) {
                b0[a] = (int)std::llround((c[a] - l.size / 2) / res) - k0[a];
                b1[a] = (int)std::llround((c[a] + l.size / 2) / res) - k0[a];   // exclusive
                if (b0[a] < 0) b0[a] = 0;
                if (b1[a] > n[a]) b1[a] = n[a];
                if (b0[a] >= b1[a]) out = true;
            }
            if (out) continue;
            for (int x = b0[0]; x < b1[0]; x++)
                for (int y = b0[1]; y < b1[1]; y++)
                    for (int z = b0[2]; z < b1[2]; z++) occ[((size_t)x * n[1] + y) * n[2] + z] = 1;
        }
    }

private:
    bool read_node(std::istream &in, double cx, double cy, double cz, double size, int depth) {
        unsigned char b[2];
        in.read((char *)b, 2);
        if (!in) { error = "truncated .bt stream"; return false; }
        if (depth >= 16) { error = "tree deeper than 16 levels"; return false; }
        inner_nodes++;
        int kind[8];
        for (int i = 0; i < 8; i++) {
            const unsigned char byte = b[i >> 2];
            const int lo = (byte >> (2 * (i & 3))) & 1, hi = (byte >> (2 * (i & 3) + 1)) & 1;
            kind[i] = lo | (hi << 1);   // 0 unknown, 1 = (1,0) free leaf, 2 = (0,1) occupied leaf, 3 = inner
        }
        const double h = size / 2, q = size / 4;
        for (int i = 0; i < 8; i++) {
            if (kind[i] == 0) continue;
            const double x = cx + ((i & 1) ? q : -q), y = cy + ((i & 2) ? q : -q), z = cz + ((i & 4) ? q : -q);
            if (kind[i] == 3) continue;
            leaves.push_back({x, y, z, h, kind[i] == 2});
        }
        for (int i = 0; i < 8; i++) {
            if (kind[i] != 3) continue;
            const double x = cx + ((i & 1) ? q : -q), y = cy + ((i & 2) ? q : -q), z = cz + ((i & 4) ? q : -q);
            if (!read_node(in, x, y, z, h, depth + 1)) return false;
        }
        return true;
    }
};

// Exact Euclidean distance transform (Felzenszwalb & Huttenlocher, one pass per axis) of an occupancy grid
// [nx][ny][nz]: distance in metres from each cell centre to the nearest occupied cell centre, clamped at max_dist.
inline std::vector<float> clamped_edt(const std::vector<unsigned char> &occ, const int n[3], double res, double max_dist) {
    const size_t total = (size_t)n[0] * n[1] * n[2];
    const double INF = 1e18;
    std::vector<double> d(total);
    for (size_t i = 0; i < total; i++) d[i] = occ[i] ? 0.0 : INF;
    const size_t stride[3] = {(size_t)n[1] * n[2], (size_t)n[2], 1};
    int nmax = n[0] > n[1] ? n[0] : n[1];
    if (n[2] > nmax) nmax = n[2];
    std::vector<double> f(nmax), z(nmax + 1), out(nmax);
    std::vector<int> v(nmax);
    for (int axis = 0; axis < 3; axis++) {
        const int len = n[axis], a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
        for (int i = 0; i < n[a1]; i++)
            for (int j = 0; j < n[a2]; j++) {
                const size_t base = (size_t)i * stride[a1] + (size_t)j * stride[a2];
                for (int q = 0; q < len; q++) f[q] = d[base + (size_t)q * stride[axis]];
                int k = -1;
                for (int q = 0; q < len; q++) {
                    if (f[q] >= INF) continue;
                    double s = 0;
                    while (k >= 0) {
                        s = ((f[q] + (double)q * q) - (f[v[k]] + (double)v[k] * v[k])) / (2.0 * q - 2.0 * v[k]);
                        if (s <= z[k]) k--; else break;
                    }
                    if (k < 0) { k = 0; v[0] = q; z[0] = -INF; z[1] = INF; }
                    else { k++; v[k] = q; z[k] = s; z[k + 1] = INF; }
                }
                if (k < 0) continue;   // nothing occupied on this line yet
                int kk = 0;
                for (int q = 0; q < len; q++) {
                    while (z[kk + 1] < q) kk++;
                    out[q] = (double)(q - v[kk]) * (q - v[kk]) + f[v[kk]];
                }
                for (int q = 0; q < len; q++) d[base + (size_t)q * stride[axis]] = out[q];
            }
    }
    std::vector<float> edt(total);
    for (size_t i = 0; i < total; i++) {
        double m = (d[i] >= INF) ? max_dist : std::sqrt(d[i]) * res;
        edt[i] = (float)(m < max_dist ? m : max_dist);
    }
    return edt;
}

}  // namespace SwarmPlanning
