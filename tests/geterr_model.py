"""Model of csrc/eval.cuh's geterr_block in plain Python integers (test infrastructure): the float32 sum of int32
energies in index order, computed the way the kernel does -- chunk summaries under a guessed binade, applied only when
the real sum is verified to stay inside it, the plain addition chain otherwise -- next to the literal serial loop of
source/triangulation.hpp:653-674.  tests/test_geterr_cpu.py checks one against the other; tests/test_step_gpu.py
checks the kernel against the serial loop."""
import numpy as np


def serial_f32(x):
    s = np.float32(0.0)
    for v in np.asarray(x, np.int64):
        s = np.float32(s + np.float32(v))
    return s


def _f32_int(v):                       # exact integer value of float32(int v)
    return int(np.float32(v))


def _summary(chunk, P0):
    """(inc0, inc1, e): increments in units of U = 2^max(e-23, 0) for entry parity 0 / 1, or None (no summary)"""
    if any(v < 0 for v in chunk):
        return None
    e = P0.bit_length() - 1 if P0 > 0 else 0
    sh = max(e - 23, 0)
    U, half = 1 << sh, (1 << sh) >> 1
    inc = [0, 0]
    for p in (0, 1):
        tot = 0
        for v in chunk:
            a = _f32_int(v)
            q, r = a >> sh, a & (U - 1)
            c = 0 if (sh == 0 or r < half) else 1 if r > half else (q + p + tot) & 1
            tot += q + c
        inc[p] = tot
    if max(inc) >= 1 << 24:
        return None
    return inc[0], inc[1], e


def _apply(s, S):
    inc0, inc1, e = S
    sh = max(e - 23, 0)
    inc = inc1 if (int(np.float32(s).view(np.uint32)) & 1) else inc0
    out = np.float32(np.float32(s) + np.float32(inc) * np.float32(2.0 ** sh))
    if (sh > 0 and not np.float32(s) >= np.float32(2.0 ** e)) or not out < np.float32(2.0 ** (max(e, 23) + 1)):
        return None
    return out


def block_f32(x, nthreads=256, group=8, stats=None):
    x = [int(v) for v in np.asarray(x, np.int64)]
    n = len(x)
    C = -(-n // nthreads) if n else 0
    chunks = [x[t * C:(t + 1) * C] for t in range(nthreads)]
    P, acc = [], 0
    for c in chunks:
        P.append(acc); acc += sum(_f32_int(max(v, 0)) for v in c)
    S = [_summary(c, P[t]) if c else "empty" for t, c in enumerate(chunks)]
    s = np.float32(0.0)
    for g0 in range(0, nthreads, group):
        live = [t for t in range(g0, min(g0 + group, nthreads)) if S[t] != "empty"]
        if not live:
            continue
        G = None
        if all(S[t] is not None for t in live) and len({S[t][2] for t in live}) == 1:
            g = [0, 0]
            for p in (0, 1):
                for t in live:
                    g[p] += S[t][(p + g[p]) & 1]
            if max(g) < 1 << 24:
                G = (g[0], g[1], S[live[0]][2])
        r = _apply(s, G) if G is not None else None
        if r is not None:
            s = r
            if stats is not None: stats["group"] += 1
            continue
        for t in live:
            r = _apply(s, S[t]) if S[t] is not None else None
            if r is not None:
                s = r
                if stats is not None: stats["chunk"] += 1
            else:
                for v in chunks[t]:
                    s = np.float32(s + np.float32(v))
                if stats is not None: stats["chain"] += 1
    return s


def cases(rng, count, nmax=6000):
    """adversarial and plain inputs"""
    for trial in range(count):
        kind = trial % 10
        n = int(rng.integers(1, nmax))
        if kind == 0: x = rng.integers(0, 2_000_000, n)
        elif kind == 1: x = rng.integers(0, 300, n)                        # exact for a long time
        elif kind == 2: x = rng.integers(0, 2 ** 31 - 1, n)                # many binades, int -> float rounding of the addends
        elif kind == 3: x = rng.integers(0, 4000, n) * 64 + 32 * rng.integers(0, 2, n)     # exact ties in some binades
        elif kind == 4: x = rng.integers(0, 2, n) * rng.integers(0, 1 << 20, n)            # zeros mixed in
        elif kind == 5: x = np.full(n, int(rng.integers(1, 1 << 22)))                      # constant
        elif kind == 6: x = rng.integers(-1000, 2_000_000, n)                              # a few negative (wrapped) energies
        elif kind == 7:
            x = rng.integers(0, 600_000, n); x[int(rng.integers(0, n))] = 2 ** 24          # lands on / near a power of two
        elif kind == 8: x = (1 << int(rng.integers(0, 30))) * rng.integers(0, 3, n)        # powers of two only: ties everywhere
        else: x = np.abs(rng.normal(500_000, 250_000, n)).astype(np.int64)                 # like a real frame
        yield np.asarray(x, np.int64).astype(np.int32)
