"""Development aid: event timeline of one steady-state item of CTA 0 (library built with -DFC_T2_TRACE=1)."""
import ctypes, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowchain_iccv2023_b200 import synth, _lib
from flowchain_iccv2023_b200.flow import fastpredNF_CIF_separate_cond
spec = synth.DEFAULT_SPEC
f = fastpredNF_CIF_separate_cond(2, 3, 64, 2, 16, flow_architecture='realNVP', pred_len=12)
f.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_params(spec, 0).items()})
f = f.cuda().eval()
A = 16
inp = synth.make_inputs(spec, A, 1)
grid = torch.from_numpy(synth.make_grid(256)).cuda()
bp, cd = torch.from_numpy(inp['base_pos']).cuda(), torch.from_numpy(inp['cond']).cuda()
lib = _lib.load()
lib.fc_debug_t2trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
prec = _lib.FC_PREC_F16X3
f.grid_log_prob(bp, cd, grid, seed=1, precision=prec); torch.cuda.synchronize()
lib.fc_debug_t2trace(None, 1)
f.grid_log_prob(bp, cd, grid, seed=2, precision=prec); torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (12 * 256))()
lib.fc_debug_t2trace(buf, 0)
a = np.array(buf[:], dtype=np.uint64).reshape(12, 256)
ev = []
for w in range(12):
    for v in a[w]:
        if v: ev.append((int(v & np.uint64(0xFFFFFFFFFFFF)), w, int(v >> np.uint64(48))))
ev.sort()
t0 = ev[0][0]
for t, w, tag in ev:
    if w in (0, 4, int(sys.argv[1]) if len(sys.argv) > 1 else 8): print(f"{t - t0:8d}  warp {w:2d}  tag {tag}")
