"""Development aid: event timeline of CTA 0 of the two-tile kernel (library built with -DFC_T3_TRACE=1, pointed to by
FLOWCHAIN_B200_LIB).  Prints, for steady-state items, how the epilogue warps of the two tiles and the two MMA issuers
spend their time: busy in an epilogue, waiting for MMAs, waiting for operands."""
import ctypes, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowchain_iccv2023_b200 import synth, _lib
from flowchain_iccv2023_b200.flow import fastpredNF_CIF_separate_cond
# pred_len 8 (default ETH otherwise): every layer is <= 64 wide, so a pass is ONE launch of the HTS layout (a 12-step pass
# is two launches that write the same trace buffer); set T3_TRACE_T=12 to trace the second (80-wide) launch instead
T_ = int(os.environ.get('T3_TRACE_T', '8'))
spec = synth.FlowSpec(pred_len=T_)
f = fastpredNF_CIF_separate_cond(2, 3, 64, 2, 16, flow_architecture='realNVP', pred_len=T_)
f.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_params(spec, 0).items()})
f = f.cuda().eval()
A = 64
inp = synth.make_inputs(spec, A, 1)
grid = torch.from_numpy(synth.make_grid(256)).cuda()
bp, cd = torch.from_numpy(inp['base_pos']).cuda(), torch.from_numpy(inp['cond']).cuda()
lib = _lib.load()
lib.fc_debug_t3trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
prec = _lib.FC_PREC_F16X3
f.grid_log_prob(bp, cd, grid, seed=1, precision=prec); torch.cuda.synchronize()
lib.fc_debug_t3trace(None, 1)
f.grid_log_prob(bp, cd, grid, seed=2, precision=prec); torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (20 * 1024))()
lib.fc_debug_t3trace(buf, 0)
a = np.array(buf[:], dtype=np.uint64).reshape(20, 1024)
# NOTE: two launches per pass write the same buffer; the second launch (steps 8-11) overwrites from index 0 of each warp
def events(w):
    out = []
    for v in a[w]:
        if v: out.append((int(v & np.uint64(0xFFFFFFFFFFFF)), int(v >> np.uint64(48))))
    return out
def summarize_compute(w):
    ev = events(w)
    busy = wait = phase = 0
    n = 0
    for (t0, g0), (t1, g1) in zip(ev, ev[1:]):
        if t1 < t0: continue          # launch boundary
        d = t1 - t0
        if g0 in (0x10, 0x11, 0x14): wait += d          # waiting for MMAs
        elif g0 in (0x12, 0x13): busy += d              # hidden-layer epilogue (until the next wait starts)
        elif g0 == 0x15: phase += d                     # E2 / E3
        elif g0 == 0x16: busy += d                      # coupling tail / gather until next wait
        n += 1
    tot = busy + wait + phase
    return busy, phase, wait, tot, n
print("epilogue warps (lane 0 of the first warp of each column part), share of traced time:")
for w in (0, 4, 8, 12):
    b, p, wt, tot, n = summarize_compute(w)
    if tot: print(f"  warp {w:2d} (tile {w // 8}, part {(w // 4) % 2}): epilogue+tails {100 * b / tot:5.1f} %  E2/E3 {100 * p / tot:5.1f} %  waiting for MMAs {100 * wt / tot:5.1f} %  ({n} events, {tot} cycles)")
for w in (16, 17, 18, 19):      # issuers: (net, tile) = ((w - 16) & 1, (w - 16) >> 1)
    ev = events(w)
    prep = wait_w = wait_o = issue = req = 0
    for (t0, g0), (t1, g1) in zip(ev, ev[1:]):
        if t1 < t0: continue
        d = t1 - t0
        if g1 == 0x22: prep += d                         # loop back + descriptors of the next chunk
        elif g1 == 0x20: wait_w += d                     # until the weight chunk landed
        elif g1 in (0x30, 0x31): wait_o += d             # until the operand was ready
        elif g1 in (0x40, 0x41): issue += d              # issuing MMAs + commits
        elif g1 in (0x50, 0x51, 0x52, 0x53): req += d    # weight-ring requests (last-tile issuers)
    tot = prep + wait_w + wait_o + issue + req
    if tot: print(f"  issuer warp {w} (net {(w - 16) & 1}, tile {(w - 16) >> 1}): bookkeeping {100 * prep / tot:5.1f} %  ring requests {100 * req / tot:5.1f} %  "
                  f"waiting for weights {100 * wait_w / tot:5.1f} %  waiting for operands {100 * wait_o / tot:5.1f} %  issuing {100 * issue / tot:5.1f} %")
# ---- per-phase durations of one epilogue warp (mean cycles between consecutive tags) ----
names = {0x10: "wait MMA net0", 0x11: "wait MMA net1", 0x12: "epilogue net0", 0x13: "epilogue net1", 0x14: "wait both L3",
         0x15: "E2/E3 gaussian phase", 0x16: "coupling tail / gather"}
for w in (0, 8):
    ev = events(w)
    acc, cnt = {}, {}
    for (t0, g0), (t1, g1) in zip(ev, ev[1:]):
        if t1 < t0: continue
        acc[g0] = acc.get(g0, 0) + (t1 - t0); cnt[g0] = cnt.get(g0, 0) + 1
    tot = sum(acc.values())
    print(f"warp {w}: phase means (cycles) and share of time")
    for g in sorted(acc):
        print(f"   {names.get(g, hex(g)):28s} n={cnt[g]:4d}  mean {acc[g] / cnt[g]:8.0f}  share {100 * acc[g] / tot:5.1f} %")

# ---- optional merged raw timeline (T3_TRACE_DUMP=path): warps 0 / 8 (tile 0 / 1, part 0), 4 / 12 (part 1), issuers 17 / 18 ----
dump = os.environ.get("T3_TRACE_DUMP")
if dump:
    allev = []
    for w in range(20):
        allev += [(t, w, g) for t, g in events(w)]
    allev.sort()
    t0 = allev[0][0]
    with open(dump, "w") as fh:
        for t, w, g in allev: fh.write(f"{t - t0:9d} w{w:02d} {g:#04x}\n")
