"""TEST INFRASTRUCTURE: oracle.Problem objects from synth missions / packed arrays."""
import numpy as np

import oracle


def oracle_problem(m, sequential=True, batch_size=1, batch_iter=-1, iteration=1):
    offs = [0]
    boxes, tend = [], []
    for b, t in m["sfc"]:
        boxes.append(b); tend.append(t); offs.append(offs[-1] + len(t))
    return oracle.Problem(m["T"], m["start"], m["goal"], m["radius"], np.array(offs, np.int32),
                          np.concatenate(boxes), np.concatenate(tend), m["rsfc_n"], m["rsfc_t"], m["init_traj"],
                          sequential=sequential, batch_size=batch_size, batch_iter=batch_iter, iteration=iteration)
