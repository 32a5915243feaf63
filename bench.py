#!/usr/bin/env python
"""bench.py -- headline benchmark of the FlowChain analytical-density hot path on B200.

Metric (BASELINE.json): log-density evals/s over a dense grid = (agent, timestep, query point)
log-probabilities per second.  A "step" is ONE pass of the hot path over one batch: the per-step
("separate") density of config 2 -- 64 agents x 12 steps x 256x256 lattice, prefix mode "self" (the literal
reference call flow.log_prob(base_pos, g.expand(T), cond)) -- per GPU.  With N > 1 GPUs every
rank evaluates its own 64 agents (weak scaling, no data-path collective) and ONE all-gather assembles the
(N*64, T, G) density maps on every rank inside the timed step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|f16x3|f16]
                    [--workload config2|config4|config5]

Workloads (BASELINE.json configs): config2 (default, the configuration the metric is quoted on; weak scaling: 64 agents
per GPU), config4 (4096 agents x 512x512 lattice in total, STRONG scaling: agents sharded over the ranks, the full
(4096, 12, 262144) maps assembled on every rank), config5 (stress model: 12 steps x 6 coupling blocks x 256-wide
conditioner MLPs, K = 32 importance samples per query point; weak scaling).

`--impl reference` times the reference algorithm's CPU port (oracle/, numpy fp32, all host threads BLAS
uses) on a bounded sample of the same workload; /root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from flowchain_iccv2023_b200 import synth  # noqa: E402

METRIC = "log_density_evals_per_sec"
UNIT = "evals/s"


class _OneLineStdout:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout from C
    code when NCCL_DEBUG or an nccl.conf asks for it), so for the whole run file descriptor 1 points at stderr and the
    JSON line is written to the saved, real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self._real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, obj):
        os.write(self._real, (json.dumps(obj) + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._real, 1)
        os.close(self._real)
        return False


OUT = None      # set by main()


def ncu_traffic(precision: str):
    """DRAM bytes (read + write) of ONE pass of the dominant kernel, from the committed `ncu --set full` capture of this
    same command (profiles/).  The f16x3 kernel (k_separate_tc3) covers the 12 steps with two launches (one per weight
    width class), so the capture holds two rows and the traffic of a pass is their sum.  None when no capture exists
    for the chosen precision."""
    path = None
    for name in {"f16x3": ("r02_ncu_full_k_separate_tc3_f16x3.csv", "r01_ncu_full_k_separate_tc3_f16x3.csv")}.get(precision, ()):
        if os.path.exists(os.path.join(ROOT, "profiles", name)):
            path = os.path.join(ROOT, "profiles", name)
            break
    if path is None:
        return None
    import csv
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return None
    hdr, unit = rows[0], rows[1]
    tot = 0.0
    for val in rows[2:]:
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if key not in hdr:
                return None
            i = hdr.index(key)
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit[i], None)
            if scale is None:
                return None
            tot += float(val[i]) * scale
    return tot


def update_latency(flow, spec, dev, agents_list=(1, 64, 1000), S=10000, t=1, reps=1000):
    """BASELINE.json's second metric (config 3): p50 latency of the rapid update (predict_from_new_obs,
    fastpredNF.py:142-164) on caches produced by the sampler, launched from a captured CUDA graph, CUDA-event timed,
    `reps` replays; the full form (reference output: positions + log-probs, 3.2 MB moved per agent) and the lazy form
    (base_lp' only, 0.36 MB per agent), with the HBM fraction of the full form on its moved bytes."""
    import torch
    from flowchain_iccv2023_b200 import _lib
    h = flow.engine()
    T = spec.pred_len
    res = {"samples": S, "t": t, "reps": reps, "launch": "cuda graph replay, CUDA events", "unit": "us"}
    for A in agents_list:
        inp = synth.make_inputs(spec, A, seed=3)
        bp = torch.from_numpy(inp["base_pos"]).to(dev)
        cd = torch.from_numpy(inp["cond"]).to(dev)
        seq, lp, ldjs, base = torch.ops.flowchain.sample_with_log_prob(h.id, bp, cd, S, None, None, 7, _lib.FC_PREC_F16X3)
        prob0 = torch.cat([seq, lp[..., None]], dim=-1).contiguous()
        norm = torch.ops.flowchain.base_normalizer(h.id, base.contiguous())
        new_pos = (bp + 0.05).contiguous()
        del seq, lp, base
        nrep = reps if A <= 64 else max(reps // 5, 100)
        for name, fn in (("", lambda: torch.ops.flowchain.update(h.id, new_pos, t, prob0, ldjs, norm)),
                         ("_lazy", lambda: torch.ops.flowchain.update_lazy(h.id, new_pos, t, prob0, norm))):
            g = torch.cuda.CUDAGraph()
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=st):
                    out = fn()
            torch.cuda.synchronize()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nrep)]
            for a, b in evs:
                a.record()
                g.replay()
                b.record()
            torch.cuda.synchronize()
            times = np.array([a.elapsed_time(b) for a, b in evs]) * 1e3
            res[f"p50{name}_{A}_agents"] = float(np.median(times))
            res[f"p99{name}_{A}_agents"] = float(np.quantile(times, 0.99))
            del g, out
        moved = A * S * ((T - t) * 12 * 2 + (T - t) * 4 + 32)      # tail rows read + written, ldjs, the position's sector
        res[f"hbm_frac_{A}_agents"] = moved / (res[f"p50_{A}_agents"] * 1e-6) / (peaks()["hbm_gbs"] * 1e9)
        del prob0, ldjs, norm
    res["moved_bytes_note"] = "full form per sample: (T-t) x 12 B read + (T-t) x 12 B written + (T-t) x 4 B ldjs + one 32 B sector of the position"
    return res


def torch_gpu_baseline(flow, spec, dev, A, n, reps=2):
    """north_star's target is stated against the reference's own single-GPU PyTorch log_prob.  The reference cannot travel
    to the GPU box, so this leg times the SAME eager op stream -- the flow's differentiable PyTorch path (flow.py
    _log_prob_separate_torch: nn.Linear / cat / tanh / relu / exp per CIF step, pinned against the unmodified reference's
    update step and the oracle in tests/test_training_cpu.py) -- on this GPU: fp32, TF32 off, rows in chunks of 4 agents
    (262 144 rows), fresh torch.randn noise like the reference."""
    import torch
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    T, G = spec.pred_len, n * n
    inp = synth.make_inputs(spec, A, seed=1000)
    gr = torch.from_numpy(synth.make_grid(n)).to(dev)
    bp = torch.from_numpy(inp["base_pos"]).to(dev)
    cd = torch.from_numpy(inp["cond"]).to(dev)
    chunk = 4 if A >= 4 else A

    def one_pass():
        outs = []
        for a0 in range(0, A, chunk):
            a1 = min(A, a0 + chunk)
            ca = a1 - a0
            rx = gr[None, :, None, :].expand(ca, G, T, 2).reshape(-1, T, 2)
            rb = bp[a0:a1, None].expand(ca, G, 2).reshape(-1, 2)
            rc = cd[a0:a1, None].expand(ca, G, T, -1).reshape(ca * G, T, -1)
            outs.append(flow._log_prob_separate_torch(rb, rx, rc))
        return outs

    with torch.no_grad():
        one_pass()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            one_pass()
        e1.record()
        torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    ms = e0.elapsed_time(e1) / reps
    return {"value": A * G * T / (ms / 1e3), "unit": UNIT, "ms_per_pass": ms, "torch": torch.__version__,
            "what": "the flow's eager PyTorch path (= the reference's op stream), fp32, TF32 off, 262 144-row chunks, same GPU"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc, self.begin = index, [], None, 0

    def mark_begin(self):
        """The timed region starts here: nvidia-smi was launched earlier (its start-up takes longer than a short timed
        region), only the samples that arrive from now on are used."""
        self.begin = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = "timed region"
        lines = self.lines[self.begin:]
        if not lines:                      # region shorter than one sampling period: same load ran during the warm-up
            lines, window = self.lines, "warm-up + timed region"
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def cpu_port_rate(spec, params, seconds: float, agents: int = 1, grid_n: int = 100):
    """Times the numpy port of the reference algorithm (oracle/) on a bounded sample: `agents` x
    grid_n^2 lattice x T steps per pass (config 1 shape), repeated for ~`seconds`.  Returns evals/s."""
    from oracle import flowchain_oracle as orc
    inp = synth.make_inputs(spec, agents, seed=1)
    grid = synth.make_grid(grid_n)
    eps = synth.make_eps(spec, agents * grid.shape[0], 2)
    p32 = orc.cast_params(params, np.float32)
    orc.grid_log_prob_self(p32, inp["base_pos"], inp["cond"], grid, eps, np.float32)   # warm-up
    n, t0 = 0, time.perf_counter()
    times = []
    while True:
        t1 = time.perf_counter()
        orc.grid_log_prob_self(p32, inp["base_pos"], inp["cond"], grid, eps, np.float32)
        times.append(time.perf_counter() - t1)
        n += 1
        if time.perf_counter() - t0 > seconds and n >= 3:
            break
    evals = agents * grid.shape[0] * spec.pred_len
    return evals / float(np.median(times)), float(np.median(times)), n


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([d.get("num_threads", 1) for d in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


WORKLOADS = {
    # name: (spec, agents per GPU or in total, lattice side, K, scaling)
    "config2": dict(spec=synth.DEFAULT_SPEC, agents=64, grid=256, K=1, scaling="weak"),
    "config4": dict(spec=synth.DEFAULT_SPEC, agents=4096, grid=512, K=1, scaling="strong"),
    "config5": dict(spec=synth.FlowSpec(pred_len=12, cond_size=16, hidden_size=256, n_hidden=2, n_blocks=6), agents=8, grid=64, K=32,
                    scaling="weak"),
}


def workload_text(name, spec, A, n, K, world):
    per = "in total (strong scaling)" if WORKLOADS[name]["scaling"] == "strong" else "per GPU"
    base = {"config2": "config2", "config4": "config4", "config5": "config5 (stress)"}[name]
    if name in ("config2", "config4") and (A, n) != (WORKLOADS[name]["agents"], WORKLOADS[name]["grid"]):
        base += " (custom size)"
    k = f", K = {K} importance samples per point" if K > 1 else ""
    return (f"{base}: {A} agents {per} x {spec.pred_len} steps x {n}x{n} lattice{k}, separate density (flow.log_prob), prefix self")


def model_text(spec):
    return (f"CIF_separate_cond (T={spec.pred_len}, C0={spec.cond_size}, H={spec.hidden_size}, {spec.n_blocks} blocks, "
            f"{spec.n_hidden} hidden), random-init weights")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    spec = wl["spec"]
    params = synth.make_params(spec, 0)
    from oracle import flowchain_oracle as orc
    try:        # torchrun exports OMP_NUM_THREADS=1: give BLAS all host cores explicitly so every N times the same baseline
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=os.cpu_count())
    except Exception:
        limiter = None
    K = wl["K"]
    n = 100 if K == 1 else 12
    sample = (f"1 agent x {n}x{n} lattice x {spec.pred_len} steps" + (f" x K = {K}" if K > 1 else "") +
              " per step (bounded sample of the workload), numpy fp32 port of the reference algorithm")
    inp = synth.make_inputs(spec, 1, seed=1)
    grid = synth.make_grid(n)
    p32 = orc.cast_params(params, np.float32)
    if K == 1:
        eps = synth.make_eps(spec, grid.shape[0], 2)
        fn = lambda: orc.grid_log_prob_self(p32, inp["base_pos"], inp["cond"], grid, eps, np.float32)
    else:
        bp, x, cd = orc.grid_rows(inp["base_pos"], inp["cond"], grid)
        eps = synth.make_eps(spec, grid.shape[0] * K, 2).reshape(grid.shape[0], K, -1).transpose(1, 0, 2)
        fn = lambda: orc.log_prob_importance(p32, bp, x, cd, eps, np.float32)
    per = []
    for i in range(args.warmup + args.steps):       # one step = one pass over the bounded sample
        t0 = time.perf_counter()
        fn()
        if i >= args.warmup:
            per.append(time.perf_counter() - t0)
    ms = float(np.mean(per)) * 1e3
    value = grid.shape[0] * spec.pred_len / (ms / 1e3)
    cores = blas_threads()
    A = wl["agents"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_text(args.workload, spec, A, wl["grid"], K, args.gpus), "model": model_text(spec),
                       "sample": sample, "note": "rate metric (evals/s): the CPU arm runs a bounded sample of the same per-row work"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    OUT.emit(line)
    del limiter


def run_ours(args):
    import torch
    import torch.distributed as dist
    from flowchain_iccv2023_b200 import _lib
    from flowchain_iccv2023_b200.flow import fastpredNF_CIF_separate_cond

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # NCCL would print its version banner on STDOUT, next to the JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    wl = WORKLOADS[args.workload]
    spec, K = wl["spec"], wl["K"]
    strong = wl["scaling"] == "strong"
    params = synth.make_params(spec, 0)
    flow = fastpredNF_CIF_separate_cond(input_size=2, n_blocks=spec.n_blocks, hidden_size=spec.hidden_size,
                                        n_hidden=spec.n_hidden, cond_label_size=spec.cond_size,
                                        flow_architecture="realNVP", pred_len=spec.pred_len)
    flow.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    flow = flow.to(dev).eval()                 # version checks stay on (the default): the packed weights follow the parameters
    prec = {"fp32": _lib.FC_PREC_FP32, "f16x3": _lib.FC_PREC_F16X3, "f16": _lib.FC_PREC_F16}[args.precision]
    flow.precision = prec
    handle = flow.engine()

    A_cfg = args.agents if args.agents is not None else wl["agents"]
    n = args.grid if args.grid is not None else wl["grid"]
    G, T = n * n, spec.pred_len
    if strong:      # total agents fixed, contiguous blocks per rank
        from flowchain_iccv2023_b200.dist import shard_range
        A_total = A_cfg
        lo, hi = shard_range(A_total, rank, world)
        A = hi - lo
        if len({shard_range(A_total, r, world)[1] - shard_range(A_total, r, world)[0] for r in range(world)}) != 1:
            raise SystemExit("bench.py: strong-scaling workloads need agents divisible by the number of GPUs")
        inp_all = synth.make_inputs(spec, A_total, seed=1000)
        inp = {k: v[lo:hi] for k, v in inp_all.items()}
    else:           # every rank owns its own agents (weak scaling)
        A, A_total, lo = A_cfg, A_cfg * world, rank * A_cfg
        inp = synth.make_inputs(spec, A, seed=1000 + rank)
    grid_np = synth.make_grid(n)
    # host (pinned) copies for the end-to-end leg, device-resident copies for the kernel-only leg
    h_base = torch.from_numpy(np.ascontiguousarray(inp["base_pos"])).pin_memory()
    h_cond = torch.from_numpy(np.ascontiguousarray(inp["cond"])).pin_memory()
    h_grid = torch.from_numpy(grid_np).pin_memory()
    d_base, d_cond, d_grid = h_base.to(dev), h_cond.to(dev), h_grid.to(dev)
    h_out = torch.empty((A, T, G), dtype=torch.float32).pin_memory()
    # multi-GPU: the density maps are all-gathered by the kernel's own result stores (NVSwitch multicast address of a
    # symmetric buffer) + one barrier; NCCL all-gather is the fallback when the platform has no multicast mapping
    gathered, mm = None, None
    if world > 1:
        if args.gather in ("auto", "multicast") and K == 1:
            try:
                from flowchain_iccv2023_b200.dist import MulticastMaps
                mm = MulticastMaps(A_total, T, G, dev)
            except Exception as e:      # noqa: BLE001 -- any failure to set up symmetric memory means "use NCCL"
                if args.gather == "multicast":
                    raise
                if rank == 0:
                    print(f"[bench] multicast maps unavailable ({type(e).__name__}: {e}); using the NCCL all-gather", file=sys.stderr)
        if mm is None:
            gathered = torch.empty((A_total, T, G), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    seed = [1]

    def hot_path(b=d_base, c=d_cond, g=d_grid):
        seed[0] += 1
        return flow.grid_log_prob(b, c, g, seed=seed[0], map_layout=True, K=K)

    def hot_path_mc(b, c, g, a0=0, na=None):
        """this rank's agents [a0, a0 + na) -> multicast stores into every rank's replica of the maps"""
        seed[0] += 1
        na = A - a0 if na is None else na
        flow.grid_log_prob_multicast(b[a0:a0 + na], c[a0:a0 + na], g, mm.mc_ptr, (lo + a0) * T * G, (T * G, 1, G), seed=seed[0])

    def step():
        if mm is not None:
            mm.barrier()                               # nobody still reads the previous maps
            hot_path_mc(d_base, d_cond, d_grid)
            mm.barrier()
            return mm.maps
        maps = hot_path()
        if world > 1:
            dist.all_gather_into_tensor(gathered, maps)
            return gathered
        return maps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)          # nvidia-smi starts now; only samples taken after mark_begin() count
    sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    barrier()

    # ---- multi-GPU: the fused gather must equal the NCCL all-gather of the same shards, bit for bit (one warm-up step) --
    gather_check = None
    if world > 1 and mm is not None:
        mm.barrier()
        s0 = seed[0]
        hot_path_mc(d_base, d_cond, d_grid)
        mm.barrier()
        seed[0] = s0
        local = hot_path()                            # same seed, same rows: this rank's shard into a plain tensor
        ref = torch.empty((A_total, T, G), dtype=torch.float32, device=dev) if A_total * T * G * 4 < 20e9 else None
        if ref is not None:
            dist.all_gather_into_tensor(ref, local)
            same = torch.equal(ref, mm.maps)
            del ref
        else:       # maps too large for a second full copy: compare this rank's view of every block against its owner's
            same = torch.equal(local, mm.maps[lo:lo + A])
            chk = torch.tensor([mm.maps.view(-1)[::4097].double().sum().item()], dtype=torch.float64, device=dev)
            allc = [torch.zeros_like(chk) for _ in range(world)]
            dist.all_gather(allc, chk)
            same = same and all(bool(torch.equal(c, chk)) for c in allc)       # every replica holds the same maps
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_check = "bit-identical" if int(flag.item()) == 1 else "MISMATCH"
        del local
        barrier()

    # ---- kernel-only leg: inputs resident in HBM, CUDA events on the launching (current) stream ----------
    sampler.mark_begin()
    l0 = handle.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(float(i))                      # flush L2 between timed iterations (outside the events)
        if mm is not None:
            mm.barrier()
        ev[i][0].record()
        if mm is not None:
            hot_path_mc(d_base, d_cond, d_grid)
            ev[i][1].record()                      # end of the dominant kernel (its stores are the all-gather)
            mm.barrier()
        else:
            maps = hot_path()
            ev[i][1].record()                      # end of the dominant kernel
            if world > 1:
                dist.all_gather_into_tensor(gathered, maps)
        ev[i][2].record()
    barrier()
    launches = handle.launch_count - l0
    clocks = sampler.stop()
    step_ms = sum(a.elapsed_time(c) for a, b, c in ev)
    kern_ms = sum(a.elapsed_time(b) for a, b, c in ev) / args.steps
    t = torch.tensor([step_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    evals_per_step = A_total * G * T
    value = evals_per_step / (ms_per_step / 1e3)

    # ---- end-to-end leg: public API with HOST buffers, H2D + D2H inside the timed region -----------------
    # 4 agent chunks per rank; the device->host copy of a finished chunk runs on a side stream under the next chunk's
    # kernel (single GPU: flow.grid_log_prob_to_host; multi-GPU: the same schedule around the multicast launches)
    side = torch.cuda.Stream(device=dev)
    nchunk = max(1, min(4, A))

    def e2e_step():
        if world == 1:
            seed[0] += 1
            if K == 1:
                flow.grid_log_prob_to_host(h_base, h_cond, h_grid, out=h_out, chunks=nchunk, seed=seed[0])
            else:
                maps = flow.grid_log_prob(h_base.to(dev, non_blocking=True), h_cond.to(dev, non_blocking=True),
                                          h_grid.to(dev, non_blocking=True), seed=seed[0], map_layout=True, K=K)
                h_out.copy_(maps, non_blocking=True)
            return
        b = h_base.to(dev, non_blocking=True)
        c = h_cond.to(dev, non_blocking=True)
        g = h_grid.to(dev, non_blocking=True)
        main = torch.cuda.current_stream(dev)
        if mm is not None:
            mm.barrier()                                # nobody still reads the previous maps
            for i in range(nchunk):
                a0, a1 = (A * i) // nchunk, (A * (i + 1)) // nchunk
                hot_path_mc(b, c, g, a0, a1 - a0)
                mm.barrier()                            # the chunk's stores have landed in every replica (incl. this one)
                evc = torch.cuda.Event()
                evc.record(main)
                side.wait_event(evc)
                with torch.cuda.stream(side):
                    h_out[a0:a1].copy_(mm.maps[lo + a0:lo + a1], non_blocking=True)
            main.wait_stream(side)
            return
        for i in range(nchunk):
            a0, a1 = (A * i) // nchunk, (A * (i + 1)) // nchunk
            seed[0] += 1
            maps = flow.grid_log_prob(b[a0:a1], c[a0:a1], g, seed=seed[0], map_layout=True, K=K)
            gathered[lo + a0:lo + a1].copy_(maps)
            evc = torch.cuda.Event()
            evc.record(main)
            side.wait_event(evc)
            with torch.cuda.stream(side):
                h_out[a0:a1].copy_(maps, non_blocking=True)
            maps.record_stream(side)
        dist.all_gather_into_tensor(gathered, gathered[lo:lo + A].clone())
        main.wait_stream(side)
    for _ in range(2):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / args.steps
    e2e_value = evals_per_step / (e2e_ms / 1e3)
    h2d = (h_base.numel() + h_cond.numel() + h_grid.numel()) * 4
    d2h = h_out.numel() * 4

    if rank == 0:
        pk = peaks()
        flops_launch = A * G * K * spec.flops_separate_row()      # algorithmic FLOPs of one pass on this GPU
        achieved = flops_launch / (kern_ms / 1e3) / 1e12
        peak = pk["bf16_tflops_sustained"]
        default_size = (A_cfg, n) == (wl["agents"], wl["grid"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": {"fp32": "f32", "f16": "f16 operands / f32 accumulate",
                      "f16x3": "split-f16 operands (3 tcgen05 passes) / f32 accumulate"}[args.precision], "data": "synthetic",
            "config": {"workload": workload_text(args.workload, spec, A_cfg, n, K, world), "model": model_text(spec),
                       "noise": "in-kernel Philox4x32-10", "precision": args.precision,
                       "parity_bar": "fp32 path: |d| <= 1e-4 + 3e-6 |ref64| nats (the reference's own fp32 is 5.9e-4 off its fp64); tensor-core f16x3: 5e-3 + 3e-6 |ref64|",
                       "l2": "L2 flushed (256 MB write) between timed iterations; output > L2",
                       "multi_gpu": ("single GPU" if world == 1 else
                                     "agents sharded; maps all-gathered by the density kernel's own stores (multimem.st to the NVSwitch multicast address of a symmetric buffer) + 1 barrier, inside the step" if mm is not None else
                                     "agents sharded; one NCCL all-gather of the density maps inside the step")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "schedule": f"{nchunk} agent chunks per rank, device->host copy of a chunk overlapped with the next chunk's kernel"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.precision) if (args.workload == "config2" and default_size) else None,
                         "traffic_note": "dram read+write bytes of one pass (all launches of the dominant kernel), ncu --set full capture in profiles/ (algorithmic output: %d B)" % (A * G * T * 4),
                         "launches_per_pass": int(launches) // max(args.steps, 1),
                         "peak_source": pk["source"] + " bf16 sustained (kernel timed inside a long step)",
                         "kernel_ms": kern_ms, "algorithmic_flops_per_launch": flops_launch},
        }
        if gather_check is not None:
            line["gather_check"] = gather_check
        if world == 1 and args.workload == "config2" and not args.no_update_latency:
            line["update_latency"] = update_latency(flow, spec, dev)
        if world == 1 and args.workload == "config2" and default_size and not args.no_torch_baseline:
            del flush
            torch.cuda.empty_cache()
            tb = torch_gpu_baseline(flow, spec, dev, A, n)
            tb["speedup_kernel"] = value / tb["value"]
            tb["speedup_e2e"] = e2e_value / tb["value"]
            line["gpu_torch_baseline"] = tb
        if world == 1 and not args.no_cpu_baseline:
            if K == 1:
                rate, med, nrep = cpu_port_rate(spec, params, seconds=args.cpu_seconds)
                line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": blas_threads(), "kind": "port",
                                        "sample": f"1 agent x 100x100 lattice x {T} steps (config 1 shape), numpy fp32 port, "
                                                  f"median of {nrep} passes ({med * 1e3:.0f} ms each)"}
        OUT.emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("FLOWCHAIN_PRECISION", "f16x3"), choices=["fp32", "f16x3", "f16"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--agents", type=int, default=None, help="override the workload's agent count (per GPU, or in total for config4)")
    ap.add_argument("--grid", type=int, default=None, help="override the workload's lattice side")
    ap.add_argument("--no-update-latency", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="auto", choices=["auto", "nccl", "multicast"],
                    help="multi-GPU assembly of the maps: multicast stores fused into the kernel, or NCCL all-gather")
    args = ap.parse_args()
    global OUT
    with _OneLineStdout() as OUT:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)


if __name__ == "__main__":
    main()
