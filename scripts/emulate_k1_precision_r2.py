"""CPU emulation of the round-2 K1 arithmetic (numpy only, ~1-2 minutes).

Round 1 (scripts/emulate_k1_precision.py) showed that the tensor core's fp32 accumulate TRUNCATES, so a
75-deep 3xTF32 chain into one accumulator carries a biased error proportional to the magnitude of the
PARTIAL sums (which, for a weak bin next to a strong one, are far larger than the final value).

Round 2 removes the truncation instead of shortening the chain. Operands are sliced on a fixed grid:
    t = fold(x) * 256 / S_row          (S_row a power of two >= the row's largest |e|,|o|; |t| <= 256)
    a1 = rint(t)                       (integer, |a1| <= 256: exact in fp16)
    a2 = fp16(t - a1), a3 = fp16(t - a1 - a2)
and the same for the windowed basis (b1, b2, b3 from 256 * w[n] cos / sin in fp64). Then
    Acc0 = sum a1*b1           every product is an integer < 2^16 and the sum of 208 of them is < 2^24:
                               EXACT in an fp32 accumulator whatever the rounding mode;
    Acc1 = sum a1*b2 + a2*b1 + a1*b3 + a3*b1 + a2*b2     terms <= 2^-9 of Acc0's, so truncation there is
                               2^-9 of what it was;
    X = 2^-16 S_row (Acc0 + Acc1)   one fp32 RN add in the epilogue.
kind::f16 has K = 16, so a 208-sample (13 k-step) chain per accumulator.

This script emulates that with truncating accumulation and prints the log10-domain error against the fp64
oracle next to (a) round 1's scheme and (b) the reference's own fp32 torch.stft formulation.
"""
import sys
import numpy as np

sys.path.insert(0, '/root/repo')
from oracle.logmel import mel_filterbank, hann_periodic, power_spectrogram_f64   # noqa: E402


def trunc32(x64):   # round toward zero to fp32
    f = x64.astype(np.float32)
    bad = np.abs(f.astype(np.float64)) > np.abs(x64)
    return np.where(bad, np.nextafter(f, np.float32(0)), f)


def f16(x):
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


def rn_tf32(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xffffe000).astype(np.uint32).view(np.float32)


def survey_signal(seconds, seed=5, amp=1.0):
    rng = np.random.default_rng(seed)
    n = 16000 * seconds
    t = np.arange(n) / 16000.0
    a = 0.1 * rng.standard_normal(n)
    for h in range(1, 11):
        a += (0.3 / h) * np.sin(2 * np.pi * 220 * h * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 3 * t))
    a[int(0.9 * n):] = 0
    return (amp * a).astype(np.float32)


def frames_of(a):
    n = len(a)
    pad = np.pad(a, (200, 200), mode='reflect')
    F = n // 160
    idx = (np.arange(F) * 160)[:, None] + np.arange(400)[None, :]
    return pad[idx]


def basis64():
    nn = np.arange(0, 208)
    w = np.where(nn <= 200, hann_periodic()[np.minimum(nn, 399)], 0.0)
    k = np.arange(201)
    ph = (np.outer(nn, k) % 400)
    C = w[:, None] * np.cos(2 * np.pi * ph / 400)
    S = -(w[:, None] * np.sin(2 * np.pi * ph / 400))
    C[200] *= 0.5           # e[200] = 2 x[200]
    S[200] = 0
    C[201:] = 0
    S[201:] = 0
    C[0] = 0                # w[0] = 0
    S[0] = 0
    return C, S


def slice_basis(B):
    t = B * 256.0
    b1 = np.rint(t)
    r = t - b1
    b2 = f16(r).astype(np.float64)
    b3 = f16(r - b2).astype(np.float64)
    return b1, b2, b3


def fold(fr, twosum):
    """e[n] = x[n] + x[400-n], o[n] = x[n] - x[400-n] for n = 0..207 in fp32 (+ the exact rounding error)."""
    nn = np.arange(0, 208)
    fwd = fr[:, nn]
    rev = fr[:, (400 - nn) % 400] if False else fr[:, np.minimum(400 - nn, 399)]   # n = 0 -> x[400]: basis row is zero
    e = (fwd + rev).astype(np.float32)
    o = (fwd - rev).astype(np.float32)
    if twosum:
        ee = (fwd.astype(np.float64) + rev) - e
        oe = (fwd.astype(np.float64) - rev) - o
        return e, o, ee.astype(np.float32), oe.astype(np.float32)
    return e, o, None, None


def run_r2(fr, C, S, row_scale=True, twosum=False, trunc=True):
    F = fr.shape[0]
    e, o, ee, oe = fold(fr, twosum)
    m = np.abs(fr).max(axis=1) if row_scale else np.full(F, np.abs(fr).max())
    ex = np.floor(np.log2(np.maximum(m, 1e-30))) + 2
    Srow = np.where(m > 0, 2.0 ** ex, 1.0)
    out = []
    for v, verr, B in ((e, ee, C), (o, oe, S)):
        t = (v.astype(np.float64) * (256.0 / Srow)[:, None]).astype(np.float32)       # exact: power-of-two scale
        a1 = np.rint(t)
        r = (t - a1).astype(np.float32)                                               # exact in fp32
        if verr is not None:
            r = (r + (verr.astype(np.float64) * (256.0 / Srow)[:, None]).astype(np.float32)).astype(np.float32)
        a2 = f16(r)
        a3 = f16((r - a2).astype(np.float32))
        assert np.abs(a1).max() <= 256
        b1, b2, b3 = slice_basis(B)
        acc0 = np.zeros((F, 201), np.float32)
        acc1 = np.zeros((F, 201), np.float32)
        A1, A2, A3 = a1.astype(np.float64), a2.astype(np.float64), a3.astype(np.float64)
        exact0 = True
        for j in range(13):
            sl = slice(16 * j, 16 * j + 16)
            s0 = acc0.astype(np.float64) + A1[:, sl] @ b1[sl]
            n0 = trunc32(s0) if trunc else s0.astype(np.float32)
            exact0 &= bool(np.all(n0.astype(np.float64) == s0))
            acc0 = n0
            for X, Y in ((A3, b1), (A1, b3), (A2, b2), (A2, b1), (A1, b2)):
                s1 = acc1.astype(np.float64) + X[:, sl] @ Y[sl]
                acc1 = trunc32(s1) if trunc else s1.astype(np.float32)
        x = (acc0 + acc1).astype(np.float32)                                           # fp32 RN add
        out.append((x.astype(np.float64) * (Srow / 65536.0)[:, None], exact0))
    (re, ex_re), (im, ex_im) = out
    return re, im, ex_re and ex_im


def run_r1(fr, trunc=True):
    nn = np.arange(1, 201)
    e = (fr[:, nn] + fr[:, 400 - nn]).astype(np.float32)
    o = (fr[:, nn] - fr[:, 400 - nn]).astype(np.float32)
    w = hann_periodic()[nn]
    k = np.arange(201)
    ph = (np.outer(nn, k) % 400)
    C = (w[:, None] * np.cos(2 * np.pi * ph / 400)); C[199] *= 0.5
    S = -(w[:, None] * np.sin(2 * np.pi * ph / 400)); S[199] = 0

    def split64(x64):
        hi = rn_tf32(x64.astype(np.float32))
        lo = rn_tf32((x64 - hi.astype(np.float64)).astype(np.float32))
        return hi, lo
    Chi, Clo = split64(C); Shi, Slo = split64(S)
    ehi = rn_tf32(e); elo = rn_tf32(e - ehi); ohi = rn_tf32(o); olo = rn_tf32(o - ohi)

    def gemm(ahi, alo, bhi, blo):
        acc = np.zeros((fr.shape[0], 201), np.float32)
        for j in range(25):
            sl = slice(8 * j, 8 * j + 8)
            for A, B in ((ahi, blo), (alo, bhi), (ahi, bhi)):
                s = acc.astype(np.float64) + (A[:, sl].astype(np.float64) @ B[sl].astype(np.float64))
                acc = trunc32(s) if trunc else s.astype(np.float32)
        return acc
    return gemm(ehi, elo, Chi, Clo).astype(np.float64), gemm(ohi, olo, Shi, Slo).astype(np.float64)


def torch_f32(a):
    import torch
    x = torch.from_numpy(a)
    st = torch.stft(x, 400, 160, window=torch.hann_window(400), return_complex=True)
    return (st[..., :-1].abs() ** 2).numpy().astype(np.float64)      # [201, F]


def report(name, P, want, W, floor):
    got = np.log10(np.maximum(W @ P, 1e-10))
    got = np.maximum(got, floor)
    e_ = np.abs(got - want)
    print(f"{name:34s} max {e_.max():.2e}  p99.99 {np.quantile(e_, 0.9999):.2e}  p99.9 {np.quantile(e_, 0.999):.2e}  "
          f"median {np.median(e_):.2e}  cells>1e-4: {(e_ > 1e-4).sum()}  >2.5e-5: {(e_ > 2.5e-5).sum()} of {e_.size}")


def main():
    W = mel_filterbank().astype(np.float64)
    C, S = basis64()
    for name, a in (("survey 40 s", survey_signal(40)), ("survey 40 s x0.07", survey_signal(40, seed=7, amp=0.07)),
                    ("tone 10 s", (0.8 * np.sin(2 * np.pi * 440.0 * np.arange(160000) / 16000.0)).astype(np.float32))):
        print("====", name)
        fr = frames_of(a)
        want = np.log10(np.maximum(W @ power_spectrogram_f64(a), 1e-10))
        floor = want.max() - 8.0
        want = np.maximum(want, floor)
        report("torch fp32 stft (the reference)", torch_f32(a), want, W, floor)
        re, im = run_r1(fr)
        report("round 1: 3xTF32, truncating acc", (re ** 2 + im ** 2).T, want, W, floor)
        for row_scale in (True, False):
            for twosum in (False, True):
                re, im, exact = run_r2(fr, C, S, row_scale=row_scale, twosum=twosum)
                report(f"round 2: {'row' if row_scale else 'tile'} scale{', twosum fold' if twosum else ''}"
                       f" [Acc0 exact: {exact}]", (re ** 2 + im ** 2).T, want, W, floor)


if __name__ == "__main__":
    main()
