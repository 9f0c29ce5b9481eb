#!/usr/bin/env python
"""Randomised cross-check of the tensor-core kernels against the exact-fp32 CUDA kernels of the same modules (not a test: a
sweep over many random shapes -- ragged tiles, tiny items, odd batch sizes, causal padding -- that prints the worst case per
family and exits non-zero on a violation).  usage: fuzz_parity.py [cases per family] [seed]"""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiotokenization_b200 import ops
from audiotokenization_b200.vq import module as M, activations

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
DEV = "cuda"
TOL = 6e-5
worst = {}
bad = []


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def both(fn):
    M.set_precision("fp32")
    want = fn()
    M.set_precision("bf16x3")
    got = fn()
    again = fn()
    M.set_precision("fp32")
    return got, again, want


def record(family, desc, got, again, want):
    e = rel(got, want)
    ok = e <= TOL and bool(torch.isfinite(got).all()) and torch.equal(got, again)
    if family not in worst or e > worst[family][0]:
        worst[family] = (e, desc)
    if not ok:
        bad.append((family, desc, e, torch.equal(got, again)))


def rand_T(lo, hi):
    return rng.choice([lo, lo + 1, 127, 128, 129, 255, 257, rng.randint(lo, hi), rng.randint(lo, hi)])


with torch.no_grad():
    for i in range(n_cases):
        C, dil, causal = rng.choice([32, 64, 128, 256]), rng.choice([1, 3, 9]), rng.random() < 0.3
        B, T = rng.randint(1, 5), rand_T(1, 3000 if C <= 64 else 1500)
        torch.manual_seed(i)
        ru = M.ResidualUnit(C, dilation=dil, causal=causal).to(DEV)
        x = torch.randn(B, T, C, device=DEV)
        record("ResidualUnit", (C, dil, causal, B, T), *both(lambda: ru.forward_cl(x)))
    for i in range(n_cases):
        ci, co, k, s = rng.choice([(32, 64, 4, 2), (64, 128, 8, 4), (128, 256, 10, 5), (256, 512, 10, 5), (512, 512, 3, 1), (512, 2048, 1, 1), (512, 512, 7, 1)])
        B, T = rng.randint(1, 4), rand_T(max(k, s), 4000 if ci <= 64 else 600)
        pad = (s // 2 + s % 2) if s > 1 else (k - 1) // 2
        torch.manual_seed(1000 + i)
        conv = M.WNConv1d(ci, co, kernel_size=k, stride=s, padding=pad).to(DEV)
        act = activations.SnakeBeta(ci, alpha_logscale=True).to(DEV) if rng.random() < 0.7 else None
        x = torch.randn(B, T, ci, device=DEV)
        record("WNConv1d", (ci, co, k, s, act is not None, B, T), *both(lambda: conv.forward_cl(x, act=act)))
    for i in range(n_cases):
        ci, co, s = rng.choice([(512, 256, 5), (256, 128, 5), (128, 64, 4), (64, 32, 2)])
        B, T = rng.randint(1, 4), rand_T(1, 700 if ci >= 256 else 2500)
        torch.manual_seed(2000 + i)
        m = M.WNConvTranspose1d(ci, co, 2 * s, stride=s, padding=s // 2 + s % 2, output_padding=s % 2).to(DEV)
        act = activations.SnakeBeta(ci, alpha_logscale=True).to(DEV)
        x = torch.randn(B, T, ci, device=DEV)
        record("WNConvTranspose1d", (ci, co, s, B, T), *both(lambda: m.forward_cl(x, act=act)))
    for i in range(max(4, n_cases // 4)):
        H, layers = rng.choice([(512, 2), (512, 1), (256, 2), (128, 1)])
        B, T = rng.choice([1, 2, 7, 64, 127, 128, 129, 200, 256, 257, 300, 384, 500, 512, 513, 600]), rng.randint(1, 40)
        torch.manual_seed(3000 + i)
        lstm = M.ResLSTM(H, num_layers=layers).to(DEV)
        x = torch.randn(B, T, H, device=DEV)
        got, again, want = both(lambda: lstm.forward_cl(x))
        e = rel(got, want)
        if "ResLSTM" not in worst or e > worst["ResLSTM"][0]:
            worst["ResLSTM"] = (e, (H, layers, B, T))
        if not (e <= 2e-4 and torch.isfinite(got).all() and torch.equal(got, again)):
            bad.append(("ResLSTM", (H, layers, B, T), e, torch.equal(got, again)))
for fam, (e, desc) in worst.items():
    print(f"{fam:20s} worst relative error {e:.3e} at {desc}")
print("violations:", len(bad))
for b in bad[:20]:
    print("  ", b)
sys.exit(1 if bad else 0)
