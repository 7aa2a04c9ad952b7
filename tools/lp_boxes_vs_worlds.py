"""Tool (CPU, this container only: reads the reference's data files): is the corridor of the reference's logged run
(`log/QPmodel.lp`, box rows of agents 61..64) reproducible from one of its committed worlds?  For every `worlds/*.bt` the
occupied columns (host/swarm_plan_cli stage=world) are tested against the 10 distinct boxes of the LP.
Result (round 2): apart from `empty.bt`, every world has occupied cells inside at least 5 of the 10 boxes -- the logged run
used a forest generated at run time (random_map_generator, unseeded), so `Corridor::updateObsBox` has no reference output
to be pinned against; tests/test_corridor_properties.py checks brute-force properties instead.
usage: python tools/lp_boxes_vs_worlds.py /root/reference/swarm_planner"""
import glob, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixture_lp as F
ref = sys.argv[1]
rec = F.recover_inputs(F.load_lp(), F.load_csv(), F.load_mission())
boxes = np.unique(rec["seg_box"][60:64].reshape(-1, 6), axis=0)
cli = os.path.join(ROOT, "swarm_simulator_b200", "host", "swarm_plan_cli")
rows = []
for w in sorted(glob.glob(os.path.join(ref, "worlds", "*.bt"))):
    out = subprocess.run([cli, os.path.join(ref, "missions", "mission_64agents_15.json"), w, "/tmp", "stage=world"],
                         capture_output=True, text=True).stdout
    c = np.array([[int(v) for v in l.split()[1:3]] for l in out.splitlines() if l.startswith("col")]).reshape(-1, 2)
    cx, cy = (c[:, 0] + 0.5) * 0.1, (c[:, 1] + 0.5) * 0.1
    bad = sum(bool(((cx > b[0] - 0.1) & (cx < b[3] + 0.1) & (cy > b[1] - 0.1) & (cy < b[4] + 0.1)).any()) for b in boxes)
    rows.append((bad, os.path.basename(w), len(c)))
for bad, name, ncol in sorted(rows):
    print("%-40s occupied columns %4d   LP boxes with an occupied cell inside: %d of %d" % (name, ncol, bad, len(boxes)))
