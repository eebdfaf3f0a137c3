"""CPU model of the digit-plane attention engine's number format (csrc/attention_i8.cu): balanced base-256 digits of q and k,
the exact diagonal sums, and the pass-1 bound c_i >= max_j z_ij that lets the probabilities be cut into unsigned 47-bit
fixed point without online rescaling. Pins the bound's constant independently of a GPU."""
import numpy as np

S = 7


def _digits256(x):
    """x (rows, 32) -> (D [S][rows][32] int64 with D_0 most significant, exponent e per row): x = 2^(e-54) sum_s D_s 256^(6-s)."""
    mx = np.abs(x).max(axis=1, keepdims=True)
    _, e = np.frexp(mx)
    e = np.where(mx > 0, e, 0)
    I = np.rint(np.ldexp(x, 54 - e)).astype(np.int64)
    D = np.zeros((S,) + x.shape, dtype=np.int64)
    for s in range(S - 1, 0, -1):
        d = ((I & 0xff) ^ 0x80) - 0x80                      # low byte as a signed digit
        I = (I - d) >> 8
        D[s] = d
    D[0] = I
    return D, e[:, 0]


def _model(q, k):
    Dq, eq = _digits256(q)
    Dk, ek = _digits256(k)
    assert np.abs(Dq[1:]).max() <= 128 and np.abs(Dq[0]).max() <= 65
    acc = np.zeros((S, q.shape[0], k.shape[0]), dtype=np.int64)
    for s in range(S):
        for t in range(S - s):
            acc[s + t] += Dq[s] @ Dk[t].T
    assert np.abs(acc).max() < 2 ** 22                       # two neighbouring diagonals merge exactly in int32
    h = np.zeros(acc.shape[1:])
    for dd in range(S - 1, -1, -1):
        h = h / 256.0 + acc[dd]
    r = np.ldexp(1.0, eq - 12) / np.sqrt(32.0)               # qscale
    ks = np.ldexp(1.0, ek)                                   # kscale
    z = h * ks[None, :] * r[:, None]
    # pass 1 (fp32 like the kernel): the two leading diagonals, then the rigorous slack for the dropped ones
    v = ((acc[0] * 256 + acc[1]).astype(np.float32) * ks[None, :].astype(np.float32)).max(axis=1)
    lead = v.astype(np.float64) * 0.00390625 * r
    c = lead + np.abs(lead) * 4.76837158203125e-07 + 24.2 * ks.max() * r
    return z, c


def test_logits_are_float64_faithful_and_bounded_by_pass1():
    rng = np.random.default_rng(1)
    for scale in (1.0, 7.0, 1e-3):
        q = rng.normal(size=(64, 32)) * scale
        k = rng.normal(size=(96, 32)) * scale * np.exp(rng.normal(size=(96, 1)))
        z, c = _model(q, k)
        want = (q @ k.T) / np.sqrt(32.0)
        assert np.abs(z - want).max() <= 1e-13 * max(1.0, np.abs(want).max())
        zmax = z.max(axis=1)
        assert np.all(c >= zmax), float((zmax - c).max())
        slack = c - zmax                                      # bits of P lost = slack / ln 2
        assert slack.max() < 4.0 * max(1.0, scale * scale), slack.max()


def test_dropped_diagonal_bound_constant():
    # sum_{dd >= 2} (dd + 1) * 32 * 128 * 128 * 256^-dd, the constant 24.2 of the kernel
    bound = sum((dd + 1) * 32 * 128 * 128 * 256.0 ** -dd for dd in range(2, S))
    assert 24.0 < bound < 24.2
