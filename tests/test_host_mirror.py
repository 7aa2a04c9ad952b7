"""Host-side mirror of the reference interface (swarm_simulator_b200/host/*.hpp): Mission / Param / point3d on the CPU,
and RBPPlanner::update() through planner_cli on the GPU."""
import json
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "swarm_simulator_b200", "host")


@pytest.fixture(scope="module")
def built():
    G.build()
    exe = os.path.join(HOST, "host_selftest")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + HOST, "-o", exe, os.path.join(HOST, "host_selftest.cpp")])
    return exe


def test_param_mission_and_point3d(built, tmp_path):
    mission = {"quadrotors": {"crazyflie": {"max_vel": [1.7, 1.7, 1.7], "max_acc": [6.2, 6.2, 6.2], "radius": 0.15, "speed": 1.0}},
               "agents": [{"name": "crazyflie", "start": [4.0, 0.0, 1.0], "goal": [-4.0, 0.0, 1.0], "radius": 0.15, "speed": 1.0},
                          {"name": "crazyflie", "start": [0.0, 4.0, 1.0], "goal": [0.0, -4.0, 1.0], "radius": 0.2, "speed": 1.0}]}
    p = tmp_path / "mission.json"
    p.write_text(json.dumps(mission, indent=2))
    out = subprocess.check_output([built, str(p)], text=True).splitlines()
    # defaults of Param::setROSParam (param.hpp L45-L70)
    assert out[0] == "default sequential=0 batch_size=4 batch_iter=0 iteration=1 n=5 phi=3 downwash=2 time_scale=1 grid_xy=0.3 z_max=2.5"
    assert out[1] == "set sequential=1 batch_size=8 batch_iter=-1 z_min=0.3"
    assert out[2] == "mission ok=1 qn=2"
    assert out[3] == "agent 0 start=4,0,1 goal=-4,0,1 r=0.15 vmax=1.7 amax=6.2"
    assert out[4] == "agent 1 start=0,4,1 goal=0,-4,1 r=0.2 vmax=1.7 amax=6.2"
    assert out[5].startswith("vec norm=3 dot=2 ")
    nx = float(out[5].split("nx=")[1].split()[0])
    assert abs(nx - np.float32(1.0) / np.float32(np.sqrt(np.float32(6.0)))) < 1e-7


def test_planner_cli_is_built(built):
    assert os.path.exists(os.path.join(HOST, "planner_cli"))


@pytest.mark.gpu
def test_rbp_planner_update_drop_in(built, tmp_path):
    """RBPPlanner(mission, param).update(log, &planResult) through the C++ mirror equals the engine called directly,
    writes the reference's CSV format, and applies timeScale (rbp_planner.hpp L209-L266) when the limits ask for it."""
    from swarm_simulator_b200 import engine as E, synth
    m = synth.synth_mission(8, 5, 0.2, 11)
    synth.dump_text(m, str(tmp_path / "dump.txt"))
    (tmp_path / "log").mkdir()
    out = subprocess.check_output([os.path.join(HOST, "planner_cli"), str(tmp_path / "dump.txt"), str(tmp_path),
                                   "plan/sequential=true", "plan/batch_size=2", "plan/batch_iter=-1", "plan/time_scale=false"],
                                  text=True)
    assert out.startswith("update=true time_scale=1 ")
    lines = (tmp_path / "traj_coef.txt").read_text().splitlines()
    info = np.array(lines[0].split(), float)
    assert info[0] == 8 and info[1] == 5 and np.array_equal(info[2:], m["T"])
    eng = E.Engine()
    r = eng.solve_many(E.PackedProblem(synth.pack([m]), sequential=True, batch_size=2))
    for qi in range(8):
        row = np.array(lines[1 + qi].split(), float)
        assert row[0] == 30 and row[1] == 3
        assert np.array_equal(row[2:].reshape(3, 30), r.coef[0, qi])          # column-major M(n+1) x 3, bit for bit
    # CSV: duration, then per axis lowest power first padded to 8, then 8 yaw zeros (L295-L324)
    csv = np.loadtxt(tmp_path / "log" / "coef1.csv", delimiter=",", skiprows=1, usecols=range(33))
    assert csv.shape == (5, 33) and np.all(csv[:, 0] == 1)
    assert np.allclose(csv[:, 1:7], r.coef[0, 0, 0].reshape(5, 6)[:, ::-1], rtol=1e-5, atol=1e-6)
    assert np.all(csv[:, 7:9] == 0) and np.all(csv[:, 25:] == 0)
    # timeScale: tighten the limits so that scaling is needed; scale is a power of 1.1 and T is stretched by it
    m2 = dict(m)
    m2["max_vel"] = np.full((8, 3), 0.3)
    synth.dump_text(m2, str(tmp_path / "dump2.txt"))
    out = subprocess.check_output([os.path.join(HOST, "planner_cli"), str(tmp_path / "dump2.txt"), str(tmp_path),
                                   "plan/sequential=true", "plan/batch_size=2", "plan/batch_iter=-1"], text=True)
    scale = float(out.split("time_scale=")[1].split()[0])
    k = np.log(scale) / np.log(1.1)
    assert scale > 1 and abs(k - round(k)) < 1e-9
    lines = (tmp_path / "traj_coef.txt").read_text().splitlines()
    info = np.array(lines[0].split(), float)
    assert np.allclose(info[2:], m["T"] * scale)
    row = np.array(lines[1].split(), float)[2:].reshape(3, 5, 6)
    want = r.coef[0, 0].reshape(3, 5, 6) * (1.0 / scale) ** np.arange(5, -1, -1)
    assert np.allclose(row, want, rtol=1e-12, atol=1e-15)
    # the scaled trajectory respects the velocity limit at the sampled extrema
    tt = np.linspace(0, scale, 201)
    vel = sum(row[..., j, None] * (5 - j) * tt ** (4 - j) for j in range(5))
    assert np.abs(vel).max() <= 0.3 * 1.1 + 1e-9


@pytest.mark.gpu
def test_corridor_update_drop_in(built, tmp_path):
    """Corridor(distmap, mission, param).update(log, &planResult) through the C++ mirror: SFC boxes and end times equal the
    workload generator's restatement of updateObsBox (exact doubles), RSFC equals the float32 oracle bit for bit."""
    import oracle
    from swarm_simulator_b200 import synth
    m = synth.synth_mission(10, 5, 0.3, 77)
    synth.dump_world_text(m, str(tmp_path / "world.txt"))
    out = subprocess.check_output([os.path.join(HOST, "corridor_cli"), str(tmp_path / "world.txt"), "box/xy_res=0.1", "box/z_res=0.1",
                                   "plan/downwash=2.0"], text=True).splitlines()
    assert out[0] == "update=true"
    i = 1
    for qi in range(10):
        tag, q, nb = out[i].split()
        assert tag == "SFC" and int(q) == qi
        boxes, tend = m["sfc"][qi]
        assert int(nb) == len(tend)
        for b in range(int(nb)):
            vals = np.array(out[i + 1 + b].split(), float)
            assert np.array_equal(vals[:6], boxes[b]) and vals[6] == tend[b]
        i += 1 + int(nb)
    rows = np.array([l.split()[1:] for l in out[i:]], float)
    no, to, _ = oracle.rsfc(m["init_traj"], m["T"], 2.0)
    assert np.array_equal(rows[:, :3].astype(np.float32).reshape(no.shape), no)
    assert np.array_equal(rows[:, 3].reshape(to.shape), to)
