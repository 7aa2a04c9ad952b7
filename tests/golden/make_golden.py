#!/usr/bin/env python3
"""Distil the reference's only frozen CPLEX artefacts into small binary fixtures.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py [/root/reference]

Inputs (read-only, never copied verbatim into the repo):
  swarm_planner/log/QPmodel.lp      CPLEX LP export of the LAST batch (l=15, agents 60..63) of a
                                    64-agent sequential run (b=4, M=36) -- written by
                                    rbp_planner.hpp L150-L152 (cplex.exportModel).
  swarm_planner/log/coef{1..64}.csv final monomial coefficients of all 64 agents, written by
                                    generateCoefCSV, rbp_planner.hpp L295-L324.
  swarm_planner/missions/mission_64agents_15.json   the mission of that run.

Outputs (committed):
  tests/golden/qpmodel_batch15.npz  the QP in index form:
        var_names order = reference variable order row = k*offset_dim + bi*offset_quad + m*6 + i
        (rbp_planner.hpp L561); objective triplets exactly as printed ("[ ... ] / 2" convention);
        constraints c1..c59652 as CSR (indptr, indices, values), sense (0:'=',1:'<=',2:'>='), rhs.
  tests/golden/coef_csv.npz         coef[64][36][3][6] lowest power first (CSV column order) + durations
  tests/golden/mission_64agents_15.npz  start/goal/radius/max_vel/max_acc arrays
"""
import json
import os
import re
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

N, M, NB, B0 = 64, 36, 4, 60  # agents, segments, batch size, first agent of batch 15


def var_index(name):
    ax, q, m, i = name.split("_")
    k = "xyz".index(ax)
    bi = int(q) - B0
    assert 0 <= bi < NB
    return k * (NB * M * 6) + bi * (M * 6) + int(m) * 6 + int(i)


TOK = re.compile(r"\s*(\[|\]|/|\^2|\*|\+|-|<=|>=|=|[A-Za-z_][A-Za-z_0-9]*|[0-9.]+(?:[eE][-+]?[0-9]+)?)")


def tokenize(s):
    pos, out = 0, []
    while pos < len(s):
        m = TOK.match(s, pos)
        if not m:
            if s[pos:].strip() == "":
                break
            raise ValueError("bad token at %r" % s[pos:pos + 40])
        out.append(m.group(1))
        pos = m.end()
    return out


def is_num(t):
    return t[0].isdigit() or t[0] == "."


def parse_objective(text):
    """[ c v ^2 | c v * w ... ] / 2  -> triplets (i, j, c) as printed."""
    toks = tokenize(text)
    assert toks[0] == "[" and toks[-3:] == ["]", "/", "2"], (toks[:3], toks[-3:])
    toks = toks[1:-3]
    I, J, V = [], [], []
    p, sign = 0, 1.0
    while p < len(toks):
        t = toks[p]
        if t == "+":
            sign = 1.0; p += 1; continue
        if t == "-":
            sign = -1.0; p += 1; continue
        c = 1.0
        if is_num(t):
            c = float(t); p += 1
        v = toks[p]; p += 1
        if toks[p] == "^2":
            w = v; p += 1
        else:
            assert toks[p] == "*"
            w = toks[p + 1]; p += 2
        I.append(var_index(v)); J.append(var_index(w)); V.append(sign * c)
        sign = 1.0
    return np.array(I, np.int32), np.array(J, np.int32), np.array(V, np.float64)


def parse_constraint(text):
    toks = tokenize(text)
    idx, val = [], []
    p, sign = 0, 1.0
    while toks[p] not in ("=", "<=", ">="):
        t = toks[p]
        if t == "+":
            sign = 1.0; p += 1; continue
        if t == "-":
            sign = -1.0; p += 1; continue
        c = 1.0
        if is_num(t):
            c = float(t); p += 1
        idx.append(var_index(toks[p])); val.append(sign * c); p += 1
        sign = 1.0
    sense = {"=": 0, "<=": 1, ">=": 2}[toks[p]]
    rest = toks[p + 1:]
    s = 1.0
    if rest[0] in "+-":
        s = -1.0 if rest[0] == "-" else 1.0
        rest = rest[1:]
    assert len(rest) == 1
    return idx, val, sense, s * float(rest[0])


def main():
    lp = open(os.path.join(REF, "swarm_planner/log/QPmodel.lp"), encoding="latin-1").read().split("\n")
    i_min = lp.index("Minimize")
    i_st = lp.index("Subject To")
    i_bd = lp.index("Bounds")
    obj_text = " ".join(lp[i_min + 1:i_st])
    assert obj_text.strip().startswith("obj1:")
    qi, qj, qv = parse_objective(obj_text.split(":", 1)[1])

    rows, cur = [], None
    for line in lp[i_st + 1:i_bd]:
        mm = re.match(r"^ c(\d+):(.*)$", line)
        if mm:
            if cur is not None:
                rows.append(cur)
            assert int(mm.group(1)) == len(rows) + 1
            cur = mm.group(2)
        else:
            cur += " " + line
    rows.append(cur)
    indptr, indices, values, sense, rhs = [0], [], [], [], []
    for r in rows:
        idx, val, sn, b = parse_constraint(r)
        indices += idx; values += val; sense.append(sn); rhs.append(b)
        indptr.append(len(indices))
    free = [l.split()[0] for l in lp[i_bd + 1:] if l.strip().endswith("Free")]
    assert len(free) == 3 * NB * M * 6 and len(set(var_index(v) for v in free)) == len(free)
    np.savez_compressed(
        os.path.join(HERE, "qpmodel_batch15.npz"),
        N=N, M=M, batch_size=NB, batch_first_agent=B0,
        obj_i=qi, obj_j=qj, obj_v=qv,
        indptr=np.array(indptr, np.int32), indices=np.array(indices, np.int32),
        values=np.array(values, np.float64), sense=np.array(sense, np.int8),
        rhs=np.array(rhs, np.float64))
    print("LP: %d obj terms, %d rows, %d nnz" % (len(qv), len(rows), len(indices)))

    coef = np.zeros((N, M, 3, 6))
    dur = np.zeros((N, M))
    for q in range(N):
        a = np.genfromtxt(os.path.join(REF, "swarm_planner/log/coef%d.csv" % (q + 1)),
                          delimiter=",", skip_header=1)
        a = a[:, :33]
        assert a.shape == (M, 33)
        dur[q] = a[:, 0]
        for k in range(3):
            coef[q, :, k, :] = a[:, 1 + 8 * k:1 + 8 * k + 6]
            assert np.all(a[:, 1 + 8 * k + 6:1 + 8 * k + 8] == 0)
    np.savez_compressed(os.path.join(HERE, "coef_csv.npz"), coef=coef, duration=dur)

    ms = json.load(open(os.path.join(REF, "swarm_planner/missions/mission_64agents_15.json")))
    ag = ms["agents"]
    start = np.zeros((N, 9)); goal = np.zeros((N, 9))
    for q, a in enumerate(ag):
        start[q, :len(a["start"])] = a["start"]
        goal[q, :len(a["goal"])] = a["goal"]
    np.savez_compressed(
        os.path.join(HERE, "mission_64agents_15.npz"),
        start=start, goal=goal,
        radius=np.array([a["radius"] for a in ag]),
        max_vel=np.array([ms["quadrotors"][a["name"]]["max_vel"] for a in ag]),
        max_acc=np.array([ms["quadrotors"][a["name"]]["max_acc"] for a in ag]))
    print("done")


if __name__ == "__main__":
    main()
