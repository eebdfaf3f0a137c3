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
