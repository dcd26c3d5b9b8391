"""Sweep of tcgen05 operand formats on the GPU box: python -m tests.gpu_umma_probe2 [case ...]
Each case runs in its own process when driven by `--all` (a bad descriptor kills the CUDA context).
Finds (1) whether a format-S tile (32B-base 128B swizzle) can be read as a K-major operand, (2) A operand in TMEM."""
import ctypes
import itertools
import subprocess
import sys

FIELDS = ["M", "N", "ksteps", "split", "a_fmt", "b_fmt", "a_mn", "b_mn", "a_lbo", "a_sbo", "a_ltype", "a_div", "a_adv_lo",
          "a_adv_hi", "b_lbo", "b_sbo", "b_ltype", "b_div", "b_adv_lo", "b_adv_hi", "RA", "RB"]
G_K = dict(lbo=144, sbo=16 * 144, ltype=0, div=1, adv_lo=0, adv_hi=288)          # format G, K-major (known good)
S_MN = dict(lbo=128 * 128, sbo=512, ltype=1, div=1, adv_lo=0, adv_hi=1024)       # format S, MN-major (known good)

CASES = [("baseline G/G K-major", "AB", {}),
         ("A in TMEM 3x", "AB", dict(a_fmt=2)),
         ("M=64 A in TMEM", "A64", dict(M=64, a_fmt=2))]
for ltype, sbo, lbo, adv in itertools.product([1, 2, 0], [1024, 512], [16, 128 * 128], [32, 16]):
    CASES.append(("A fmt S K-major ltype=%d sbo=%d lbo=%d adv_lo=%d" % (ltype, sbo, lbo, adv), "AB",
                  dict(a_fmt=1, a_ltype=ltype, a_sbo=sbo, a_lbo=lbo, a_div=4 if adv == 32 else 8, a_adv_lo=adv, a_adv_hi=128 * 128,
                       split=3)))
# single k-step variants: no start-address advance at all (checks the layout independent of the advance rule)
for ltype, sbo, lbo in itertools.product([1, 2], [1024, 512], [16, 128 * 128]):
    CASES.append(("A fmt S K-major 1 kstep ltype=%d sbo=%d lbo=%d" % (ltype, sbo, lbo), "AB8",
                  dict(a_fmt=1, a_ltype=ltype, a_sbo=sbo, a_lbo=lbo, a_div=4, a_adv_lo=32, a_adv_hi=128 * 128, ksteps=1)))
CASES.append(("B fmt S K-major R=64 (dual use)", "AB", dict(b_fmt=1, b_ltype=1, b_sbo=512, b_lbo=64 * 128, b_div=4, b_adv_lo=32, b_adv_hi=64 * 128)))
CASES.append(("A TMEM x B fmt S MN-major R=64", "ABn", dict(a_fmt=2, b_fmt=1, b_mn=1, b_ltype=1, b_sbo=512, b_lbo=64 * 128, b_div=1, b_adv_lo=0, b_adv_hi=1024)))
CASES.append(("A fmt S K-major x B fmt S MN-major R=64", "ABn", dict(a_fmt=1, a_ltype=1, a_sbo=512, a_lbo=128 * 128, a_div=4, a_adv_lo=32, a_adv_hi=128 * 128,
                                                               b_fmt=1, b_mn=1, b_ltype=1, b_sbo=512, b_lbo=64 * 128, b_div=1, b_adv_lo=0, b_adv_hi=1024)))
# 64-row tiles: A K-major R=64 (M=64); M-stacked A = [hi | lo]^T (adjacent tiles, MN-major, M=128) x B MN-major, K = 64 rows
CASES.append(("A fmt S K-major R=64 M=64", "A64", dict(M=64, a_fmt=1, a_ltype=1, a_sbo=512, a_lbo=64 * 128, a_div=4, a_adv_lo=32, a_adv_hi=64 * 128)))
CASES.append(("M-stacked A [hi|lo]^T MN-major M=128 K=64 rows, 1x", "AtB64s", dict(M=128, ksteps=8, split=1, a_fmt=1, b_fmt=1, a_mn=1, b_mn=1, a_lo_off=16384,
                                                                           a_lbo=64 * 128, a_sbo=512, a_ltype=1, a_div=1, a_adv_lo=0, a_adv_hi=1024,
                                                                           b_lbo=64 * 128, b_sbo=512, b_ltype=1, b_div=1, b_adv_lo=0, b_adv_hi=1024)))
kw = {}
for side in "ab":
    for k, v in S_MN.items():
        kw[side + "_" + k] = v
CASES.append(("fmt S MN-major both M=64 K=128", "AtB2", dict(M=64, ksteps=16, a_fmt=1, b_fmt=1, a_mn=1, b_mn=1, **kw)))


def one(idx):
    import torch
    from scgib_b200 import _lib
    from tests import probe_lib
    lib = probe_lib.load()
    dev = "cuda:0"
    tag, kind, kwargs = CASES[idx]
    torch.manual_seed(0)
    A = torch.randn(128, 64, device=dev)
    B = torch.randn(64, 64, device=dev)
    L128 = list(range(128))
    L64 = [(i // 16) * 32 + i % 16 for i in range(64)]
    if kind == "AB":
        ref, lanes = A.double() @ B.double().t(), L128
    elif kind == "AB8":
        ref, lanes = A.double()[:, :8] @ B.double()[:, :8].t(), L128
    elif kind == "ABn":
        ref, lanes = A.double() @ B.double(), L128
    elif kind == "AtB64s":
        A = A[:64].contiguous()
        B = torch.randn(64, 64, device=dev)
        Ah = A.double()
        import struct
        # tf32 round-to-nearest-away of A, emulated on the host
        a32 = A.cpu().numpy().view("uint32").astype("uint64")
        hi = (((a32 + 0x1000) & 0xFFFFE000).astype("uint32")).view("float32")
        hi_t = torch.from_numpy(hi.copy()).double().to(dev)
        lo_t = Ah - hi_t
        ref = torch.cat([hi_t.t() @ B.double(), lo_t.t() @ B.double()], 0)
        lanes = L128
    elif kind == "A64":
        A = A[:64].contiguous()
        ref, lanes = A.double() @ B.double().t(), L64
    else:
        B = torch.randn(128, 64, device=dev)
        ref, lanes = A.double().t() @ B.double(), L64
    p = dict(M=128, N=64, ksteps=8, split=3, a_fmt=0, b_fmt=0, a_mn=0, b_mn=0, RA=A.shape[0], RB=B.shape[0])
    for side in "ab":
        for k, v in G_K.items():
            p[side + "_" + k] = v
    p.update(kwargs)
    arr = (ctypes.c_int32 * (len(FIELDS) + 2))(*([int(p[f]) for f in FIELDS] + [1, int(p.get("a_lo_off", 0))]))
    out = torch.full((129, 64), float("nan"), device=dev)
    rc = lib.scgib_debug_umma2(_lib.ptr(A), _lib.ptr(B), _lib.ptr(out), arr, None)
    torch.cuda.synchronize()
    got = out[lanes].double()[:, :ref.shape[1]]
    err = float((got - ref).abs().max() / ref.abs().max())
    alt = ""
    if kind == "AtB64s":
        lo_err = float((got[64:] - ref[64:]).abs().max() / ref[64:].abs().max())
        alt = " (lo rows 64..127 alone: %.2e)" % lo_err
    if kind == "A64":
        got2 = out[:64].double()
        alt = " (lanes 0..63: %.2e)" % float((got2 - ref).abs().max() / ref.abs().max())
    print("%-62s rc=%d rel_err=%.3e %s%s" % (tag, rc, err, "OK" if err < 3e-6 else ("tf32-ok" if err < 3e-3 else ""), alt), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] != "--all":
        one(int(sys.argv[1]))
    else:
        for i in range(len(CASES)):
            r = subprocess.run([sys.executable, "-m", "tests.gpu_umma_probe2", str(i)], capture_output=True, text=True)
            if r.returncode != 0:
                msg = [ln for ln in r.stderr.splitlines() if "rror" in ln][-1:] or ["failed"]
                print("%-62s FAILED: %s" % (CASES[i][0], msg[0][:100]), flush=True)
            else:
                print(r.stdout.strip(), flush=True)
