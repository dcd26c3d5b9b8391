"""Cycles per tcgen05.mma (kind::tf32, M=128, N=64, K=8) for the operand formats used by the kernels:
python -m tests.gpu_umma_time"""
import ctypes
import torch
from scgib_b200 import _lib
from tests.gpu_umma_probe2 import FIELDS, G_K

lib = _lib.load()
dev = "cuda:0"
F2 = FIELDS + ["reps"]
S_K = dict(lbo=128 * 128, sbo=512, ltype=1, div=4, adv_lo=32, adv_hi=128 * 128)
G_D = dict(lbo=144, sbo=16 * 144, ltype=0, div=1, adv_lo=0, adv_hi=288)


def run(tag, reps, **kw):
    A = torch.randn(128, 64, device=dev)
    B = torch.randn(64, 64, device=dev)
    p = dict(M=128, N=64, ksteps=8, split=3, a_fmt=0, b_fmt=0, a_mn=0, b_mn=0, RA=128, RB=64, reps=reps)
    for side in "ab":
        for k, v in G_K.items():
            p[side + "_" + k] = v
    p.update(kw)
    arr = (ctypes.c_int32 * (len(F2) + 1))(*([int(p[f]) for f in F2] + [0]))
    out = torch.zeros(128 * 64 + 1, device=dev)
    res = []
    for _ in range(3):
        lib.scgib_debug_umma2(_lib.ptr(A), _lib.ptr(B), _lib.ptr(out), arr, None)
        torch.cuda.synchronize()
        res.append(float(out[-1]))
    n = reps * p["ksteps"] * p["split"]
    print("%-40s reps=%4d mmas=%6d cycles=%9.0f  -> %.1f cycles/mma" % (tag, reps, n, min(res), min(res) / n), flush=True)


S_MN = dict(lbo=128 * 128, sbo=512, ltype=1, div=1, adv_lo=0, adv_hi=1024)
mn = {}
for side in "ab":
    for k, v in S_MN.items():
        mn[side + "_" + k] = v
skw = {"a_" + k: v for k, v in S_K.items()}
for reps in (128,):
    run("A smem G K-major, 3x", reps)
    run("A smem S K-major, 3x", reps, a_fmt=1, **skw)
    run("A in TMEM, 3x", reps, a_fmt=2)
    run("N=128 A smem G (timing only)", reps, N=128)
    run("M=64 N=64 A smem G", reps, M=64)
    run("M=64 N=128 A smem G", reps, M=64, N=128)
    run("M=64 N=64 S MN-major both, 3x", reps, M=64, ksteps=16, a_fmt=1, b_fmt=1, a_mn=1, b_mn=1, RB=128, **mn)
    run("M=64 N=64 S MN-major both, 1x", reps, M=64, ksteps=16, a_fmt=1, b_fmt=1, a_mn=1, b_mn=1, RB=128, split=1, **mn)
    run("M=64 N=128 S MN-major both, 1x", reps, M=64, N=128, ksteps=16, a_fmt=1, b_fmt=1, a_mn=1, b_mn=1, RB=128, split=1, **mn)
    kmn = dict(mn); kmn.update(skw)
    run("M=128 A S K-major, B S MN-major N=64 1x", reps, M=128, a_fmt=1, b_fmt=1, a_mn=0, b_mn=1, split=1, **kmn)
    amn = dict(mn); amn.update({"b_" + k: v for k, v in G_K.items()})
    run("M=64 A S MN-major, B G K-major N=64 1x", reps, M=64, a_fmt=1, b_fmt=0, a_mn=1, b_mn=0, split=1, **amn)
