import ctypes, os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
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
lib.fc_debug_t2prof.argtypes = [ctypes.c_void_p, ctypes.c_int]
for prec in (_lib.FC_PREC_F16X3, _lib.FC_PREC_F16):
    f.grid_log_prob(bp, cd, grid, seed=1, precision=prec); torch.cuda.synchronize()
    lib.fc_debug_t2prof(None, 1)
    f.grid_log_prob(bp, cd, grid, seed=2, precision=prec); torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 16)()
    lib.fc_debug_t2prof(out, 0)
    items = (A * 65536 // 128) * 12 / 148
    names = ['tail_prev', 'prologue', 'gather', 'r_hidden', 'r_waitL3', 'r_E2', 'cpl_other', 'q_hidden', 'q_waitL3', 'q_E3', 'tail', 'cpl_update', 'cpl_wait_mma0', 'cpl_wait_mma1', 'cpl_epi(hidden)', 'cpl_arrive']
    tot = sum(out[:16])
    print('prec', prec, 'items/CTA', items, 'total cycles/item', tot / items)
    for i, nm in enumerate(names): print(f'  {nm:12s} {out[i] / items:9.0f} cycles/item  {out[i] / tot:.3f}')
