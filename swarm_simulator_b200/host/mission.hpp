// mission.hpp -- ROS-free mirror of /root/reference/swarm_planner/include/mission.hpp (L10-L88): same members; the
// mission file format is the reference's missions/*.json ({"quadrotors": {name: {max_vel, max_acc, ...}},
// "agents": [{name, start, goal, radius, speed}]}).  A small recursive-descent JSON reader replaces rapidjson.
#pragma once
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace SwarmPlanning {
namespace json {
struct Value {
    enum Kind { Null, Num, Str, Arr, Obj, Bool } kind = Null;
    double num = 0;
    bool b = false;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;
    const Value *find(const std::string &k) const {
        for (auto &kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
};
class Parser {
public:
    explicit Parser(const std::string &s) : s_(s) {}
    bool parse(Value &out) { bool ok = value(out); ws(); return ok && p_ == s_.size(); }
private:
    const std::string &s_;
    size_t p_ = 0;
    void ws() { while (p_ < s_.size() && std::isspace((unsigned char)s_[p_])) p_++; }
    bool lit(const char *w) { size_t n = std::string(w).size(); if (s_.compare(p_, n, w) == 0) { p_ += n; return true; } return false; }
    bool string(std::string &o) {
        if (s_[p_] != '"') return false;
        p_++;
        while (p_ < s_.size() && s_[p_] != '"') { if (s_[p_] == '\\' && p_ + 1 < s_.size()) p_++; o.push_back(s_[p_++]); }
        if (p_ >= s_.size()) return false;
        p_++;
        return true;
    }
    bool value(Value &v) {
        ws();
        if (p_ >= s_.size()) return false;
        char c = s_[p_];
        if (c == '{') {
            v.kind = Value::Obj; p_++; ws();
            if (s_[p_] == '}') { p_++; return true; }
            for (;;) {
                ws();
                std::string k; Value x;
                if (!string(k)) return false;
                ws(); if (s_[p_++] != ':') return false;
                if (!value(x)) return false;
                v.obj.emplace_back(k, x);
                ws();
                if (s_[p_] == ',') { p_++; continue; }
                if (s_[p_] == '}') { p_++; return true; }
                return false;
            }
        }
        if (c == '[') {
            v.kind = Value::Arr; p_++; ws();
            if (s_[p_] == ']') { p_++; return true; }
            for (;;) {
                Value x;
                if (!value(x)) return false;
                v.arr.push_back(x);
                ws();
                if (s_[p_] == ',') { p_++; continue; }
                if (s_[p_] == ']') { p_++; return true; }
                return false;
            }
        }
        if (c == '"') { v.kind = Value::Str; return string(v.str); }
        if (lit("true")) { v.kind = Value::Bool; v.b = true; return true; }
        if (lit("false")) { v.kind = Value::Bool; return true; }
        if (lit("null")) return true;
        char *end = nullptr;
        v.num = std::strtod(s_.c_str() + p_, &end);
        if (end == s_.c_str() + p_) return false;
        v.kind = Value::Num;
        p_ = end - s_.c_str();
        return true;
    }
};
}  // namespace json

class Mission {
public:
    int qn = 0;  // the number of quadrotors
    std::vector<std::vector<double>> startState, goalState, max_vel, max_acc;
    std::vector<double> quad_size, quad_speed;

    // mission.hpp L22-L88 (takes the path directly instead of reading the "mission" ROS param)
    bool setMission(const std::string &mission_addr) {
        std::ifstream ifs(mission_addr);
        if (!ifs) return false;
        std::stringstream ss;
        ss << ifs.rdbuf();
        std::string text = ss.str();
        json::Value doc;
        if (!json::Parser(text).parse(doc)) return false;
        const json::Value *agents = doc.find("agents"), *quads = doc.find("quadrotors");
        if (!agents || agents->kind != json::Value::Arr) return false;
        qn = (int)agents->arr.size();
        startState.assign(qn, std::vector<double>(9, 0));
        goalState.assign(qn, std::vector<double>(9, 0));
        quad_size.assign(qn, 0); quad_speed.assign(qn, 0);
        max_vel.assign(qn, std::vector<double>(3, 0));
        max_acc.assign(qn, std::vector<double>(3, 0));
        for (int qi = 0; qi < qn; qi++) {
            const json::Value &a = agents->arr[qi];
            const json::Value *name = a.find("name"), *start = a.find("start"), *goal = a.find("goal");
            const json::Value *radius = a.find("radius"), *speed = a.find("speed");
            if (!name || !start || !goal || !radius) return false;
            for (size_t i = 0; i < start->arr.size() && i < 9; i++) startState[qi][i] = start->arr[i].num;
            for (size_t i = 0; i < goal->arr.size() && i < 9; i++) goalState[qi][i] = goal->arr[i].num;
            quad_size[qi] = radius->num;
            quad_speed[qi] = speed ? speed->num : 0;
            const json::Value *q = quads ? quads->find(name->str) : nullptr;
            if (!q) return false;
            const json::Value *mv = q->find("max_vel"), *ma = q->find("max_acc");
            for (size_t i = 0; mv && i < mv->arr.size() && i < 3; i++) max_vel[qi][i] = mv->arr[i].num;
            for (size_t i = 0; ma && i < ma->arr.size() && i < 3; i++) max_acc[qi][i] = ma->arr[i].num;
        }
        return true;
    }
};
}  // namespace SwarmPlanning
