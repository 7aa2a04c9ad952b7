"""The reference's own smoke loop (src/swarm_traj_planner_rbp_test_all.cpp L49-L103 + launch/plan_rbp_test.launch): the
64-agent mission on the 50 random-forest maps worlds/map1..50.bt, every stage must return true.  Runs through
swarm_simulator_b200/host/swarm_plan_cli.
usage: python tools/plan_test_all.py <swarm_planner dir> [stage=ecbs|all] [first last]      (stage=all needs the GPU)"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "swarm_simulator_b200", "host", "swarm_plan_cli")
PARAMS = ["ecbs/w=1.5", "grid/xy_res=0.5", "grid/z_res=1.0", "grid/margin=0.2", "world/z_min=0.3", "plan/sequential=true",
          "plan/batch_size=4", "plan/batch_iter=-1"]          # plan_rbp_test.launch L27-L59


def main():
    pkg = sys.argv[1]
    stage = sys.argv[2] if len(sys.argv) > 2 else "stage=all"
    first, last = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1, 50)
    ok = 0
    for i in range(first, last + 1):
        t = time.time()
        out = subprocess.run([CLI, os.path.join(pkg, "missions", "mission_64agents_15.json"), os.path.join(pkg, "worlds", "map%d.bt" % i),
                              "/tmp", stage] + PARAMS, capture_output=True, text=True)
        good = out.returncode == 0
        ok += good
        tail = " | ".join(l for l in out.stdout.splitlines()[1:] if not l.startswith("traj"))
        print("map%-2d %s %.2fs  %s" % (i, "ok  " if good else "FAIL", time.time() - t, tail), flush=True)
    print("%d / %d maps planned" % (ok, last - first + 1))
    return 0 if ok == last - first + 1 else 1


if __name__ == "__main__":
    sys.exit(main())
