"""TEST INFRASTRUCTURE (never imported by the product): a numpy restatement of the device generator
gcn-max-cut_b200/csrc/postproc.cu::mt19937_uniform_kernel -- numpy's legacy MT19937 (randomkit, the generator behind the
reference's np.random.rand() draws in TestingNeuralNetwork.py:18-46) with the state regeneration written the way the kernel
does it: word i from the OLD state, words i + 227 and i + 454 from the word produced 227 places earlier, then word 623.
tests/test_oracle.py pins it to np.random itself; tests/test_gpu_kernels.py pins the kernel to np.random on the device."""
import numpy as np

N, M = 624, 397
UPPER, LOWER, MATRIX_A = np.uint32(0x80000000), np.uint32(0x7FFFFFFF), np.uint32(0x9908B0DF)


def _mix(cur, nxt, far):
    y = (cur & UPPER) | (nxt & LOWER)
    return far ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), MATRIX_A, np.uint32(0)).astype(np.uint32)


def regenerate(mt: np.ndarray) -> np.ndarray:
    """One in-place regeneration of the 624-word state as three data-parallel sweeps + the last word."""
    old = mt.copy()
    new = mt.copy()
    i = np.arange(227)
    v0 = _mix(old[i], old[i + 1], old[i + M])                    # [0, 227): old words only
    v1 = _mix(old[i + 227], old[i + 228], v0)                    # [227, 454): needs new[i]
    new[i], new[i + 227] = v0, v1
    j = np.arange(169)
    new[j + 454] = _mix(old[j + 454], old[j + 455], v1[j])       # [454, 623): needs new[i + 227]
    new[N - 1] = _mix(old[N - 1:N], new[0:1], new[M - 1:M])[0]   # 623: new[0] and new[396]
    return new


def temper(y: np.ndarray) -> np.ndarray:
    y = y ^ (y >> np.uint32(11))
    y = y ^ ((y << np.uint32(7)) & np.uint32(0x9D2C5680))
    y = y ^ ((y << np.uint32(15)) & np.uint32(0xEFC60000))
    return y ^ (y >> np.uint32(18))


def uniform(state: np.ndarray, pos: int, n: int):
    """(n doubles, new state, new position): what n calls of np.random.rand() return from (state, pos), block by block with the
    kernel's carry logic (a double whose two words straddle a regeneration)."""
    mt = np.asarray(state, dtype=np.uint32).copy()
    out = np.empty(n, dtype=np.float64)
    oi, remaining, carry, cw = 0, n, False, np.uint32(0)
    mk = lambda a, b: ((a >> np.uint32(5)).astype(np.float64) * 67108864.0 + (b >> np.uint32(6)).astype(np.float64)) / 9007199254740992.0
    while remaining > 0:
        if pos == N:
            mt = regenerate(mt)
            pos = 0
        start = pos
        if carry:
            out[oi] = mk(np.asarray([cw]), temper(mt[pos:pos + 1]))[0]
            oi, remaining, start, carry = oi + 1, remaining - 1, pos + 1, False
        if remaining == 0:
            pos = start
            break
        pairs = min(remaining, (N - start) // 2)
        tw = temper(mt[start:start + 2 * pairs])
        out[oi:oi + pairs] = mk(tw[0::2], tw[1::2])
        oi, remaining = oi + pairs, remaining - pairs
        used = start + 2 * pairs
        if remaining > 0 and used < N:
            cw, carry, used = temper(mt[N - 1:N])[0], True, N
        pos = used
    return out, mt, pos
