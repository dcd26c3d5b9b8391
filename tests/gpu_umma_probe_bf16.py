"""Sweep of tcgen05 kind::f16 (bf16) operand formats on the GPU box: python -m tests.gpu_umma_probe_bf16 [first_case]
Cases run in one process; if a bad descriptor kills the CUDA context the driver restarts after the failing case.
Pins the tile format "B" of csrc/umma.cuh (SWIZZLE_128B rows read K-major and MN-major), M = 64 / 128, N <= 256, the
64-byte swizzle for 32-column tiles, the no-swizzle core layout and the A operand in tensor memory."""
import ctypes
import subprocess
import sys

FIELDS = ["M", "N", "ksteps", "a_src", "a_layout", "b_layout", "a_p0", "a_p1", "b_p0", "b_p1", "a_mn", "b_mn",
          "a_lbo", "a_sbo", "a_ltype", "a_div", "a_adv_lo", "a_adv_hi", "b_lbo", "b_sbo", "b_ltype", "b_div", "b_adv_lo",
          "b_adv_hi", "RA", "CA", "RB", "CB", "reps"]


def k128(R, side):     # SW128 tile [R][C] read K-major: 64-column blocks R*128 bytes apart, 4 K steps per block
    return {side + "_layout": 0, side + "_p0": R * 128, side + "_mn": 0, side + "_lbo": 16, side + "_sbo": 1024,
            side + "_ltype": 2, side + "_div": 4, side + "_adv_lo": 32, side + "_adv_hi": R * 128}


def mn128(R, side, lbo=None, sbo=1024):    # SW128 tile [R = K][C = MN] read MN-major: K step = 16 rows = 2048 B
    return {side + "_layout": 0, side + "_p0": R * 128, side + "_mn": 1, side + "_lbo": R * 128 if lbo is None else lbo,
            side + "_sbo": sbo, side + "_ltype": 2, side + "_div": 1, side + "_adv_lo": 0, side + "_adv_hi": 2048}


def case(tag, kind, A, B, M, N, ksteps, **kw):
    d = dict(M=M, N=N, ksteps=ksteps, a_src=0, a_p1=0, b_p1=0, RA=A[0], CA=A[1], RB=B[0], CB=B[1], reps=1)
    d.update(kw)
    return tag, kind, d


CASES = []
CASES.append(case("K/K SW128 M128 N64 K64", "ABt", (128, 64), (64, 64), 128, 64, 4, **k128(128, "a"), **k128(64, "b")))
CASES.append(case("K/K SW128 M128 N64 K128 (2 blocks)", "ABt", (128, 128), (64, 128), 128, 64, 8, **k128(128, "a"), **k128(64, "b")))
CASES.append(case("K/K SW128 M128 N128 K64", "ABt", (128, 64), (128, 64), 128, 128, 4, **k128(128, "a"), **k128(128, "b")))
CASES.append(case("K/K SW128 M128 N256 K64", "ABt", (128, 64), (256, 64), 128, 256, 4, **k128(128, "a"), **k128(256, "b")))
CASES.append(case("K/K SW128 M64 N64 K64", "ABt", (64, 64), (64, 64), 64, 64, 4, **k128(64, "a"), **k128(64, "b")))
CASES.append(case("K/K SW128 M64 N128 K128", "ABt", (64, 128), (128, 128), 64, 128, 8, **k128(64, "a"), **k128(128, "b")))
CASES.append(case("A MN-major M64 (K=64 rows) x B K-major", "AtBt", (64, 64), (64, 64), 64, 64, 4, **mn128(64, "a"), **k128(64, "b")))
CASES.append(case("A MN-major M128 (2 blocks, lbo=block) x B K-major", "AtBt", (64, 128), (64, 64), 128, 64, 4, **mn128(64, "a"), **k128(64, "b")))
CASES.append(case("A MN-major M128 lbo/sbo swapped", "AtBt", (64, 128), (64, 64), 128, 64, 4, **mn128(64, "a", lbo=1024, sbo=64 * 128), **k128(64, "b")))
CASES.append(case("A K-major x B MN-major N64 (K=64 rows)", "AB", (128, 64), (64, 64), 128, 64, 4, **k128(128, "a"), **mn128(64, "b")))
CASES.append(case("A K-major x B MN-major N128 (2 blocks)", "AB", (128, 64), (64, 128), 128, 128, 4, **k128(128, "a"), **mn128(64, "b")))
CASES.append(case("A K-major x B MN-major N128 lbo/sbo swapped", "AB", (128, 64), (64, 128), 128, 128, 4, **k128(128, "a"), **mn128(64, "b", lbo=1024, sbo=64 * 128)))
CASES.append(case("both MN-major M64 N64 K=128 rows", "AtB", (128, 64), (128, 64), 64, 64, 8, **mn128(128, "a"), **mn128(128, "b")))
CASES.append(case("both MN-major M128 N128 K=64 rows", "AtB", (64, 128), (64, 128), 128, 128, 4, **mn128(64, "a"), **mn128(64, "b")))
CASES.append(case("both MN-major M64 N64 K=64 rows", "AtB", (64, 64), (64, 64), 64, 64, 4, **mn128(64, "a"), **mn128(64, "b")))
CASES.append(case("A in TMEM (lo = even k) x B K-major", "ABt", (128, 64), (64, 64), 128, 64, 4, a_src=2, **k128(128, "a"), **k128(64, "b")))
CASES.append(case("A in TMEM (hi = even k) x B K-major", "ABt", (128, 64), (64, 64), 128, 64, 4, a_src=3, **k128(128, "a"), **k128(64, "b")))
CASES.append(case("A in TMEM K=128 N=128", "ABt", (128, 128), (128, 128), 128, 128, 8, a_src=2, **k128(128, "a"), **k128(128, "b")))
sw64k = lambda R, s: {s + "_layout": 1, s + "_p0": R * 64, s + "_mn": 0, s + "_lbo": 16, s + "_sbo": 512, s + "_ltype": 4,
                      s + "_div": 2, s + "_adv_lo": 32, s + "_adv_hi": R * 64}
CASES.append(case("K/K SW64 (32-col tiles) M128 N64 K32", "ABt", (128, 32), (64, 32), 128, 64, 2, **sw64k(128, "a"), **sw64k(64, "b")))
sw64mn = lambda R, s: {s + "_layout": 1, s + "_p0": R * 64, s + "_mn": 1, s + "_lbo": R * 64, s + "_sbo": 512, s + "_ltype": 4,
                       s + "_div": 1, s + "_adv_lo": 0, s + "_adv_hi": 1024}
CASES.append(case("A K-major SW128 x B MN-major SW64 N32 (K=64 rows)", "AB", (128, 64), (64, 32), 128, 32, 4, **k128(128, "a"), **sw64mn(64, "b")))
CASES.append(case("A MN-major SW128 M64 x B MN-major SW64 N32", "AtB", (64, 64), (64, 32), 64, 32, 4, **mn128(64, "a"), **sw64mn(64, "b")))
core = lambda R, C, s: {s + "_layout": 2, s + "_p0": 128, s + "_p1": (C // 8) * 128, s + "_mn": 0, s + "_lbo": 128,
                        s + "_sbo": (C // 8) * 128, s + "_ltype": 0, s + "_div": 1, s + "_adv_lo": 0, s + "_adv_hi": 256}
CASES.append(case("K/K no-swizzle cores M128 N64 K64", "ABt", (128, 64), (64, 64), 128, 64, 4, **core(128, 64, "a"), **core(64, 64, "b")))
CASES.append(case("A SW128 K-major x B no-swizzle cores", "ABt", (128, 64), (64, 64), 128, 64, 4, **k128(128, "a"), **core(64, 64, "b")))
for N in (64, 128, 256):
    CASES.append(case("timing M128 N%d K64 x64" % N, "time", (128, 64), (N, 64), 128, N, 4, reps=64, **k128(128, "a"), **k128(N, "b")))
CASES.append(case("timing M64 N64 K64 x64", "time", (64, 64), (64, 64), 64, 64, 4, reps=64, **k128(64, "a"), **k128(64, "b")))
CASES.append(case("timing A in TMEM N64 x64", "time", (128, 64), (64, 64), 128, 64, 4, reps=64, a_src=2, **k128(128, "a"), **k128(64, "b")))


def run_from(first):
    import torch
    from tests import probe_lib
    lib = probe_lib.load()
    dev = "cuda:0"
    L64 = [(i // 16) * 32 + i % 16 for i in range(64)]
    for idx in range(first, len(CASES)):
        tag, kind, p = CASES[idx]
        print("CASE %d" % idx, flush=True)
        torch.manual_seed(idx)
        A = torch.randn(p["RA"], p["CA"], device=dev)
        B = torch.randn(p["RB"], p["CB"], device=dev)
        Ab, Bb = A.bfloat16().double(), B.bfloat16().double()
        ref = {"ABt": lambda: Ab @ Bb.t(), "AB": lambda: Ab @ Bb, "AtBt": lambda: Ab.t() @ Bb.t(), "AtB": lambda: Ab.t() @ Bb,
               "time": lambda: Ab @ Bb.t()}[kind]()
        arr = (ctypes.c_int32 * len(FIELDS))(*[int(p[f]) for f in FIELDS])
        out = torch.full((128 * 256 + 1,), float("nan"), device=dev)
        rc = lib.scgib_debug_umma_bf16(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), ctypes.c_void_p(out.data_ptr()), arr, None)
        torch.cuda.synchronize()
        cyc = float(out[-1])
        o = out[:-1].view(128, 256).double()
        lanes = list(range(128)) if p["M"] == 128 else L64
        if kind == "time":
            print("RESULT %-58s rc=%d cycles/MMA %.1f" % (tag, rc, cyc / (p["reps"] * p["ksteps"])), flush=True)
            continue
        got = o[lanes][:, :p["N"]]
        err = float((got - ref).abs().max() / ref.abs().max())
        alt = ""
        if p["M"] == 64:
            alt = " (lanes 0..63: %.2e)" % float((o[:64, :p["N"]] - ref).abs().max() / ref.abs().max())
        print("RESULT %-58s rc=%d rel_err=%.3e %s%s" % (tag, rc, err, "OK" if err < 1e-5 else "", alt), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--from":
        run_from(int(sys.argv[2]))
    else:
        first = 0
        while first < len(CASES):
            r = subprocess.run([sys.executable, "-m", "tests.gpu_umma_probe_bf16", "--from", str(first)], capture_output=True, text=True)
            last = first - 1
            for ln in r.stdout.splitlines():
                if ln.startswith("RESULT"):
                    print(ln[7:], flush=True)
                elif ln.startswith("CASE"):
                    last = int(ln.split()[1])
            if r.returncode == 0:
                break
            msg = [ln for ln in r.stderr.splitlines() if "rror" in ln][-1:] or ["failed"]
            print("%-58s FAILED: %s" % (CASES[last][0] if last >= 0 else "?", msg[0][:100]), flush=True)
            first = max(last, first) + 1
