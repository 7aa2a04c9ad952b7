// sp_const.hpp -- ROS-free mirror of /root/reference/swarm_planner/include/sp_const.hpp (L1-L28): the shared types
// that flow between the planner stages.  When the real octomap / std_msgs headers are available define
// RBPE_HAVE_ROS and this header simply forwards to them; otherwise minimal stand-ins with the same member names are
// provided (octomap::point3d is a float32 vector with octomath::Vector3's operator semantics; std_msgs::
// Float64MultiArray keeps {layout.dim[i].size, data}).
#pragma once

#define SP_EPSILON          1e-9
#define SP_EPSILON_FLOAT    1e-6
#define SP_INFINITY         1e+9

#define SP_PT_RBP            0
#define SP_PT_SCP            1
#define SP_IPT_ECBS          0

#include <cmath>
#include <string>
#include <utility>
#include <vector>

#ifdef RBPE_HAVE_ROS
#include <octomap/OcTree.h>
#include <std_msgs/Float64MultiArray.h>
#include <std_msgs/MultiArrayDimension.h>
#else
namespace octomap {
// float32 3-vector; arithmetic in float, norm()/dot() return double of the float result (octomath::Vector3)
class point3d {
public:
    point3d() : d_{0, 0, 0} {}
    point3d(float x, float y, float z) : d_{x, y, z} {}
    float &x() { return d_[0]; }
    float &y() { return d_[1]; }
    float &z() { return d_[2]; }
    const float &x() const { return d_[0]; }
    const float &y() const { return d_[1]; }
    const float &z() const { return d_[2]; }
    float &operator()(unsigned i) { return d_[i]; }
    const float &operator()(unsigned i) const { return d_[i]; }
    point3d operator-(const point3d &o) const { return point3d(d_[0] - o.d_[0], d_[1] - o.d_[1], d_[2] - o.d_[2]); }
    point3d operator+(const point3d &o) const { return point3d(d_[0] + o.d_[0], d_[1] + o.d_[1], d_[2] + o.d_[2]); }
    point3d operator*(float s) const { return point3d(d_[0] * s, d_[1] * s, d_[2] * s); }
    point3d &operator/=(float s) { d_[0] /= s; d_[1] /= s; d_[2] /= s; return *this; }
    bool operator==(const point3d &o) const { return d_[0] == o.d_[0] && d_[1] == o.d_[1] && d_[2] == o.d_[2]; }
    double dot(const point3d &o) const { return d_[0] * o.d_[0] + d_[1] * o.d_[1] + d_[2] * o.d_[2]; }
    double norm_sq() const { return d_[0] * d_[0] + d_[1] * d_[1] + d_[2] * d_[2]; }
    double norm() const { return std::sqrt(norm_sq()); }
    point3d &normalize() {
        double len = norm();
        if (len > 0) *this /= (float)len;
        return *this;
    }
private:
    float d_[3];
};
}  // namespace octomap

namespace std_msgs {
struct MultiArrayDimension { std::string label; unsigned size = 0, stride = 0; };
struct MultiArrayLayout { std::vector<MultiArrayDimension> dim; unsigned data_offset = 0; };
struct Float64MultiArray { MultiArrayLayout layout; std::vector<double> data; };
}  // namespace std_msgs
#endif

typedef std::vector<std::vector<octomap::point3d>> initTraj_t;
typedef std::vector<std::vector<std::pair<std::vector<double>, double>>> SFC_t;
typedef std::vector<std::vector<std::vector<std::pair<octomap::point3d, double>>>> RSFC_t;

namespace SwarmPlanning {
struct PlanResult {
    initTraj_t initTraj;                                       // discrete initial trajectory: pi_0,...,pi_M
    std::vector<double> T;                                     // segment time: T_0,...,T_M
    SFC_t SFC;                                                 // safe flight corridors to avoid obstacles
    RSFC_t RSFC;                                               // relative safe flight corridors (only qi<qj filled)
    std_msgs::Float64MultiArray msgs_traj_info;                // [N, n, T_0, ... , T_M]
    std::vector<std_msgs::Float64MultiArray> msgs_traj_coef;   // per agent, column-major M(n+1) x 3
};
}  // namespace SwarmPlanning
