"""Informational baseline for north_star's ">= 50x the reference's own single-GPU PyTorch log_prob throughput" target.

The unmodified reference cannot travel to the GPU box (it lives in /root/reference, which does not exist there), so this is a
PyTorch-eager restatement of the same per-step density (flow.log_prob over rows = agents x lattice, fastpredNF.py:450-465,
571-584, 769-774) written against the reference's state_dict keys: the same nn.Linear / cat / tanh / relu / exp / Normal
sequence of small eager kernels the reference launches, fp32, TF32 off, rows processed in chunks.  It is NOT used by the
product, the tests or bench.py; it prints one JSON line: its agreement with the CUDA library on an injected noise tape (so
the baseline demonstrably computes the same function) and its throughput at config 2 (64 agents x 256x256 lattice).
"""
import json, math, os, sys, time
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowchain_iccv2023_b200 import synth, _lib
from flowchain_iccv2023_b200.flow import fastpredNF_CIF_separate_cond

LOG_2PI = math.log(2.0 * math.pi)


def mlp(p, pre, n_lin, x, act):
    for l in range(n_lin):
        x = F.linear(x, p[f"{pre}.{2 * l}.weight"], p[f"{pre}.{2 * l}.bias"])
        if l < n_lin - 1:
            x = act(x)
    return x


def density(p, pre, inp):
    return mlp(p, pre + ".shift_net", 3, inp, torch.relu), mlp(p, pre + ".log_scale_net", 3, inp, torch.relu)


def gauss_lp(w, mu, ls):
    var = torch.exp(ls) ** 2
    return -0.5 * w.shape[-1] * LOG_2PI - ls.sum(-1) - 0.5 * ((w - mu) ** 2 / var).sum(-1)


def cif_forward(p, spec, i, z, y, eps):
    mu, ls = density(p, f"net.{i}.r_u_given_z", torch.cat([z, y], -1))
    u = torch.exp(ls) * eps + mu
    log_r = gauss_lp(u, mu, ls)
    x, total = z, 0
    for b in range(spec.n_blocks):
        pre = f"net.{i}.bijection.{2 * b}"
        mask = p[pre + ".mask"]
        inp = torch.cat([u, x * mask], -1)
        s = mlp(p, pre + ".s_net", spec.n_hidden + 2, inp, torch.tanh)
        t = mlp(p, pre + ".t_net", spec.n_hidden + 2, inp, torch.relu) * (1 - mask)
        log_s = torch.tanh(s) * (1 - mask)
        x = x * torch.exp(log_s) + t
        total = total + log_s
        pre = f"net.{i}.bijection.{2 * b + 1}"
        var = p[pre + ".running_var"]
        x = torch.exp(p[pre + ".log_gamma"]) * (x - p[pre + ".running_mean"]) / torch.sqrt(var + 1e-5) + p[pre + ".beta"]
        total = total + (p[pre + ".log_gamma"] - 0.5 * torch.log(var + 1e-5))
    mu, ls = density(p, f"net.{i}.q_u_given_w", torch.cat([x, y], -1))
    return x, total.sum(-1) + gauss_lp(u, mu, ls) - log_r


def log_prob(p, spec, base_pos, x, cond, eps=None):
    """rows (N, ...) -> (N, T); eps None = torch.randn like the reference."""
    N, T = x.shape[0], x.shape[1]
    flat = x.reshape(N, -1)
    out, off = [], 0
    for i in range(T):
        C = spec.cond_size + 2 * i
        y = torch.cat([cond[:, i], flat[:, :2 * i]], -1)
        e = torch.randn(N, C, device=x.device) if eps is None else eps[:, off:off + C]
        off += C
        w, l = cif_forward(p, spec, i, x[:, i], y, e)
        prev = base_pos if i == 0 else x[:, i - 1]
        std = torch.clamp(p["var"][i], min=1e-4)
        base = (-((w - prev) ** 2) / (2 * std ** 2) - torch.log(std) - math.log(math.sqrt(2 * math.pi))).sum(-1)
        out.append(base + l)
    return torch.stack(out, 1)


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda")
    spec = synth.FlowSpec(); params = synth.make_params(spec, 0)
    p = {k: torch.from_numpy(v).to(dev) for k, v in params.items()}
    T = spec.pred_len
    # agreement with the CUDA library on an injected tape (small lattice)
    f = fastpredNF_CIF_separate_cond(2, spec.n_blocks, spec.hidden_size, spec.n_hidden, spec.cond_size, flow_architecture="realNVP", pred_len=T)
    f.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}); f = f.to(dev).eval()
    A, n = 3, 24
    inp = synth.make_inputs(spec, A, seed=5); grid = synth.make_grid(n); G = grid.shape[0]
    eps = torch.from_numpy(synth.make_eps(spec, A * G, 6)).to(dev)
    bp = torch.from_numpy(inp["base_pos"]).to(dev); cd = torch.from_numpy(inp["cond"]).to(dev); gr = torch.from_numpy(grid).to(dev)
    rows_x = gr[None, :, None, :].expand(A, G, T, 2).reshape(A * G, T, 2)
    rows_b = bp[:, None].expand(A, G, 2).reshape(A * G, 2)
    rows_c = cd[:, None].expand(A, G, T, -1).reshape(A * G, T, -1)
    with torch.no_grad():
        ref = log_prob(p, spec, rows_b, rows_x, rows_c, eps).reshape(A, G, T)
    ours = f.grid_log_prob(bp, cd, gr, eps=eps, precision=_lib.FC_PREC_FP32)
    agree = float((ours - ref).abs().max().item())
    # throughput at config 2, rows in chunks of 2^18 (activations of one chunk ~ 1 GB)
    A, n = 64, 256
    inp = synth.make_inputs(spec, A, seed=1000); gr = torch.from_numpy(synth.make_grid(n)).to(dev); G = n * n
    bp = torch.from_numpy(inp["base_pos"]).to(dev); cd = torch.from_numpy(inp["cond"]).to(dev)
    chunk_agents = 4

    def one_pass():
        outs = []
        for a0 in range(0, A, chunk_agents):
            a1 = a0 + chunk_agents
            rx = gr[None, :, None, :].expand(chunk_agents, G, T, 2).reshape(-1, T, 2)
            rb = bp[a0:a1, None].expand(chunk_agents, G, 2).reshape(-1, 2)
            rc = cd[a0:a1, None].expand(chunk_agents, G, T, -1).reshape(chunk_agents * G, T, -1)
            outs.append(log_prob(p, spec, rb, rx, rc))
        return torch.cat(outs, 0)

    with torch.no_grad():
        one_pass()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 2
        for _ in range(reps):
            one_pass()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"torch_eager_fp32_ms_per_pass": ms, "torch_eager_fp32_evals_per_s": A * G * T / (ms / 1e3),
                      "max_abs_diff_vs_cuda_fp32_path_on_tape": agree, "torch": torch.__version__,
                      "workload": "config 2: 64 agents x 12 steps x 256x256 lattice, rows in chunks of 4 agents (262144 rows)"}))


if __name__ == "__main__":
    main()
