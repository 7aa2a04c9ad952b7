"""Generates tests/golden/missions_cfg{1..5}.npz: seeded synthetic missions of the BASELINE.json configurations
 from swarm_simulator_b200/synth.py, stored compactly (synth.save_pack).

    python tests/golden/make_missions.py [--jobs 8]

Seeds follow SURVEY 8d: seed = 1000 * config_id + trial.  Runtime ~10 min on 8 cores (the space-time path planner is
pure Python); the packs are committed so that neither the tests nor bench.py pay that at run time.
"""
import argparse
import os
import sys
from multiprocessing import Pool

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from swarm_simulator_b200 import synth  # noqa: E402

CONFIGS = {
    # name: (N, M, [(rho, seed), ...])
    "cfg1": (4, 3, [(0.0, 1000 + i) for i in range(256)]),
    "cfg2": (16, 5, [(0.2, 2000 + i) for i in range(256)]),
    "cfg3": (64, 5, [(0.2, 3000 + i) for i in range(512)]),
    "cfg4": (256, 5, [(0.4, 4000 + i) for i in range(32)]),
    "cfg5": (1024, 5, [(r, 5000 + i) for i, r in enumerate((0.1, 0.2, 0.3, 0.4, 0.5))]),
}


def _job(a):
    N, M, rho, seed = a
    m = synth.synth_mission(N, M, rho, seed)
    m.pop("edt", None)
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=8)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    with Pool(a.jobs) as pool:
        for name, (N, M, items) in CONFIGS.items():
            if a.only and name not in a.only.split(","):
                continue
            ms = pool.map(_job, [(N, M, rho, seed) for rho, seed in items], chunksize=1)
            path = os.path.join(HERE, "missions_%s.npz" % name)
            synth.save_pack(ms, [rho for rho, _ in items], path)
            print(name, len(ms), "missions ->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
