"""CPU model of the digit-plane (Ozaki) GEMM the tcgen05 kernels implement (csrc/ozaki_gemm.cu): slice both operands into
7-bit digits, multiply planes as exact integers, recombine the diagonals in float64 -- checked against a float64 matmul.
Pins the algorithm (digit rule, diagonal truncation, scales) independently of a GPU."""
import numpy as np


def _slice(x, S):
    """Rows of x -> (digits [S][rows][k] int64, exponent e per row) with x = 2^(e-6) * sum_s d_s 128^(-s), |d_s| <= 64:
    d_s = q_s - 128 q_(s-1), q_s = rint(x 2^(6-e) 128^s) (the telescoped form slice_group() uses)."""
    mx = np.abs(x).max(axis=1, keepdims=True)
    _, e = np.frexp(mx)
    e = np.where(mx > 0, e, 0)
    t = np.ldexp(x, 6 - e)
    digits, qprev = [], np.zeros_like(x)
    for s in range(S):
        q = np.rint(t * 128.0 ** s)
        digits.append((q - 128.0 * qprev).astype(np.int64))
        qprev = q
    return np.stack(digits), e


def _ozaki_matmul(x, w, S):
    dx, ex = _slice(x, S)
    dw, ew = _slice(w, S)
    assert np.abs(dx).max() <= 64 and np.abs(dw).max() <= 64
    acc = np.zeros((S, x.shape[0], w.shape[0]), dtype=np.int64)           # diagonal dd = s + t, products with dd >= S dropped
    for s in range(S):
        for t in range(S - s):
            acc[s + t] += dx[s] @ dw[t].T
    assert np.abs(acc).max() < 2 ** 23                                     # int32 accumulators with headroom for the pair merge
    h = np.zeros(acc.shape[1:])
    for dd in range(S - 1, -1, -1):                                        # Horner in float64
        h = h / 128.0 + acc[dd]
    return h * np.ldexp(1.0, ex - 6) * np.ldexp(1.0, (ew - 6).T)


def test_seven_slices_are_float64_faithful():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(96, 128)) * np.exp(rng.normal(size=(96, 1)) * 2)
    x[:, ::7] *= 1e-6
    w = rng.normal(size=(64, 128)) / np.sqrt(128) * np.exp(rng.normal(size=(64, 1)))
    want = x @ w.T
    scale = np.abs(x).max(1, keepdims=True) * np.abs(w).max(1)[None, :] * np.sqrt(128)
    err7 = np.abs(_ozaki_matmul(x, w, 7) - want) / scale
    err6 = np.abs(_ozaki_matmul(x, w, 6) - want) / scale
    assert err7.max() < 2e-13, err7.max()                                  # the bound tests/test_gpu_parity.py enforces on the GPU
    assert err6.max() < 2e-11 and err6.max() > err7.max()


def test_zero_rows_and_exact_powers_of_two():
    x = np.zeros((4, 128)); x[1, 3] = 1.0; x[2, :] = -0.5; x[3, 0] = 2.0 ** -40
    w = np.eye(128)[:8] * 3.0
    got = _ozaki_matmul(x, w, 7)
    assert np.array_equal(got, x @ w.T)


# ----------------------------------------------------------------------------- one-FMA digit cutter (cut_digits16, csrc/ozaki_gemm.cu)

def _cut_digits_onefma(x, e, S):
    """Bit-level model of cut_digits16(): ONE rounded product per value, I = rint(x 2^(6-e) 128^(S-1)), read out of the low
    mantissa bits of x * c + (1.5 * 2^52 + B), B = 64 * sum_k 128^k; plane s = ((I + B) >> 7 (S-1-s)) & 127 (the leading plane keeps
    all eight bits), minus 64 -- done on four bytes of a word at once with a carry-free add of 0xC0. Returns int8 digits [S][n]."""
    B = 64 * (((1 << (7 * S)) - 1) // 127)
    magic = 6755399441055744.0 + float(B)
    assert magic == 6755399441055744 + B                                   # the constant is exact
    v = np.ldexp(x, 6 - e + 7 * (S - 1)) + magic                           # the power-of-two scaling is exact: one rounding, like the FMA
    bits = v.view(np.uint64)
    lo, hi = (bits & np.uint64(0xFFFFFFFF)).astype(np.uint64), (bits >> np.uint64(32)).astype(np.uint64)
    out = []
    for s in range(S):
        sh = 7 * (S - 1 - s)
        if sh == 0:
            a = lo
        elif sh + 8 <= 32:
            a = lo >> np.uint64(sh)
        elif sh < 32:
            a = (((hi << np.uint64(32)) | lo) >> np.uint64(sh)) & np.uint64(0xFFFFFFFF)      # funnel shift
        else:
            a = hi >> np.uint64(sh - 32)
        u = a & np.uint64(0xFF)                                            # PRMT picks byte 0 of each value
        t = (u & np.uint64(0x7F)) + np.uint64(0x40)                        # per byte: no carry into the neighbour (<= 0xBF)
        b = (t ^ ((~u) & np.uint64(0x80))) if s == 0 else (t ^ np.uint64(0x80))
        out.append((b & np.uint64(0xFF)).astype(np.uint8).view(np.int8))
    return np.stack(out)


def test_onefma_digits_represent_the_same_integer_as_the_telescoped_ones():
    rng = np.random.default_rng(3)
    for S in (4, 5, 6, 7):
        x = rng.normal(size=(64, 128)) * np.exp(rng.normal(size=(64, 1)) * 3)
        x[:, ::5] *= 1e-7
        x[0, :] = 0.0
        mx = np.abs(x).max(axis=1, keepdims=True)
        _, e = np.frexp(mx)
        e = np.where(mx > 0, e, 0)
        x[1, 0] = np.ldexp(1.0, int(e[1, 0])) * (1 - 2.0 ** -53)           # |x| -> 2^e from below: the leading digit reaches +-64
        x[2, 0] = -np.ldexp(1.0, int(e[2, 0])) * (1 - 2.0 ** -53)
        tele, _ = _slice(x, S)
        one = np.stack([_cut_digits_onefma(x[r], int(e[r, 0]), S) for r in range(x.shape[0])], axis=1).astype(np.int64)
        assert np.abs(one).max() <= 64
        w = 128 ** np.arange(S - 1, -1, -1, dtype=np.int64)
        val_one = np.tensordot(w, one, axes=1)
        val_tele = np.tensordot(w, tele, axes=1)
        assert np.array_equal(val_one, val_tele)                           # both digit sets are the integer rint(x 2^(6-e) 128^(S-1))
        assert np.array_equal(val_one, np.rint(np.ldexp(x, 6 - e + 7 * (S - 1))).astype(np.int64))
