"""Tool (CPU only): missions of the 64-agent pack through the product kernels under the fiber emulator (tests/cpu_emu) against
the oracle -- iteration counts and control points.  Used to vet kernel variants before spending GPU time.
usage: [BS=1] [TH=256 | -256 (latency kernel)] [SMEM=49152] python tools/emu_missions.py <emulator .so> <mission index>..."""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import emu_util, oracle_util
from swarm_simulator_b200 import engine as E, synth
emu_util._lib = C.CDLL(sys.argv[1])
emu_util._lib.emu_solve_many.argtypes = [C.POINTER(E.RbpeProblem), C.c_int, C.c_int, C.POINTER(E.RbpeResult), C.c_size_t, C.c_int, C.c_double, C.c_double, C.c_int]
emu_util._lib.emu_solve_many.restype = C.c_int
ms = synth.load_pack(os.path.join(ROOT, 'tests', 'golden', 'missions_cfg3.npz'), select=[int(a) for a in sys.argv[2:]])
tot=0; bad=0; worst=0
for m in ms:
    t=time.time()
    prob = E.PackedProblem(synth.pack([m]), sequential=True, batch_size=int(os.environ.get("BS","1")))
    r = emu_util.emu_solve_many(prob, threads=int(os.environ.get("TH","256")), smem_bytes=int(os.environ.get("SMEM","49152")))
    ro = oracle_util.oracle_problem(m, sequential=True, batch_size=int(os.environ.get("BS","1"))).update()
    it_g = r.qp_iters[0]; it_o = np.array(ro["batch_iters"][:len(it_g)])
    d = float(np.abs(r.ctrl[0]-ro["ctrl"]).max())
    worst=max(worst,d); tot+=len(it_g); bad+=int((it_g!=it_o).sum())
    print(m['seed'], 'rc', r.rc, 'oracle status', ro['status'], 'iter mismatches', int((it_g!=it_o).sum()), 'of', len(it_g), 'max|dctrl| %.2e'%d, '%.0fs'%(time.time()-t), flush=True)
print('TOTAL mismatches', bad, 'of', tot, 'worst', worst)
