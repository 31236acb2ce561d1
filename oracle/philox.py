"""Philox4x32-10 counter-based RNG and the per-chain stream layout (numpy).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``walnuts_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may use it.

The device code (``walnuts_b200/csrc/philox.cuh``) and the C oracle
(``oracle/c/walnuts_oracle.c``) implement exactly the same integer function, so the
uniforms are bit-identical on every side; normals are Box-Muller on those uniforms
and agree to a few ulp (libm differences only).

Stream layout (one Philox key per run, one counter per draw):
    key     = (seed & 0xffffffff, seed >> 32)
    counter = (block, iteration, chain_id, stream)
    uniform index n of a stream  ->  block n >> 1, words (2*(n&1), 2*(n&1)+1)
    u = ((w_a >> 5) * 2**26 + (w_b >> 6)) * 2**-53            in [0, 1)
    normal pair p of a stream    ->  block p: u1 = words (0,1), u2 = words (2,3)
        r = sqrt(-2 log(1 - u1));  z[2p] = r cos(2 pi u2);  z[2p+1] = r sin(2 pi u2)

Streams per (chain, iteration) -- the reference draws these with numpy's global RNG
(WALNUTSpy/WALNUTS.py:216,236,298,395,426,464,512,554,613 and
adaptiveIntegrators.py:392); SURVEY.md section 8 row A13 lists the order:
    STREAM_DIR   0   the M direction uniforms, drawn up-front (WALNUTS.py:216)
    STREAM_MOM   1   the d momentum normals (WALNUTS.py:236; walnuts.py:325)
    STREAM_SEQ   2   every other scalar uniform, consumed strictly in reference order
    STREAM_INIT  3   initial positions (iteration 0), used by the drivers only
Package mode (walnuts/walnuts.py) keys its scalar draws by purpose instead of by a
running count, so that an extension may stop at its first sub-U-turn
(SURVEY.md appendix D item 7):
    STREAM_PKG_DIR     4   index = depth                    (walnuts.py:330)
    STREAM_PKG_ELL     5   index = 2**depth - 1 + step      (walnuts.py:194,256)
    STREAM_PKG_ACCEPT  6   index = depth                    (walnuts.py:346)
    STREAM_PKG_SELECT  7   index = 2**depth - 1 + step      (walnuts.py:350)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
S32 = np.uint64(32)

STREAM_DIR, STREAM_MOM, STREAM_SEQ, STREAM_INIT = 0, 1, 2, 3
STREAM_PKG_DIR, STREAM_PKG_ELL, STREAM_PKG_ACCEPT, STREAM_PKG_SELECT = 4, 5, 6, 7

TWO_M53 = 2.0 ** -53


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11).  Inputs are array-likes of
    uint32 values (broadcastable); returns four uint64 arrays holding 32-bit words."""
    c0 = np.asarray(c0, dtype=np.uint64) & MASK
    c1 = np.asarray(c1, dtype=np.uint64) & MASK
    c2 = np.asarray(c2, dtype=np.uint64) & MASK
    c3 = np.asarray(c3, dtype=np.uint64) & MASK
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for r in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> S32, p0 & MASK
        hi1, lo1 = p1 >> S32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _to_unit(wa, wb):
    return ((wa >> np.uint64(5)).astype(np.float64) * 67108864.0
            + (wb >> np.uint64(6)).astype(np.float64)) * TWO_M53


def uniforms(seed, chain, iteration, stream, idx):
    """Uniform(s) with index/indices ``idx`` of the given stream, in [0, 1)."""
    idx = np.asarray(idx, dtype=np.uint64)
    w = philox4x32_10(idx >> np.uint64(1), iteration, chain, stream,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    odd = (idx & np.uint64(1)).astype(bool)
    wa = np.where(odd, w[2], w[0])
    wb = np.where(odd, w[3], w[1])
    return _to_unit(wa, wb)


def normals(seed, chain, iteration, stream, d):
    """The first ``d`` standard normals of a stream (Box-Muller on uniform pairs)."""
    npair = (d + 1) // 2
    p = np.arange(npair, dtype=np.uint64)
    w = philox4x32_10(p, iteration, chain, stream, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u1 = _to_unit(w[0], w[1])
    u2 = _to_unit(w[2], w[3])
    r = np.sqrt(-2.0 * np.log(1.0 - u1))
    ang = 2.0 * np.pi * u2
    z = np.empty(2 * npair)
    z[0::2] = r * np.cos(ang)
    z[1::2] = r * np.sin(ang)
    return z[:d]


class ChainStreams:
    """Sequential view of one chain's streams, in the shape the oracles consume.

    ``begin_iteration(it)`` must be called at the start of every transition; ``uniform()``
    then walks STREAM_SEQ.  ``n_seq`` counts the scalars consumed in the iteration."""

    def __init__(self, seed, chain):
        self.seed = int(seed)
        self.chain = int(chain)
        self.iteration = 0
        self.n_seq = 0
        self._buf = None

    def begin_iteration(self, iteration):
        self.iteration = int(iteration)
        self.n_seq = 0
        self._buf = None

    def directions(self, M):
        return uniforms(self.seed, self.chain, self.iteration, STREAM_DIR, np.arange(M))

    def momentum(self, d):
        return normals(self.seed, self.chain, self.iteration, STREAM_MOM, d)

    def uniform(self):
        n = self.n_seq
        if self._buf is None or n >= self._buf_base + len(self._buf):
            self._buf_base = n
            self._buf = uniforms(self.seed, self.chain, self.iteration, STREAM_SEQ,
                                 np.arange(n, n + 256))
        self.n_seq += 1
        return float(self._buf[n - self._buf_base])

    def keyed(self, stream, idx):
        return float(uniforms(self.seed, self.chain, self.iteration, stream, [idx])[0])
