"""TEST INFRASTRUCTURE: numpy model of the digit arithmetic of the tcgen05 GEMM (torcwa_b200/csrc/tc_gemm.cu).

Mirrors the kernels' integer arithmetic exactly (python ints / int64), so that
  * the algorithm's accuracy can be checked on the CPU, and
  * the GPU kernels can be compared with it digit for digit.
Never imported by the product package.
"""
import numpy as np


def exponent(mx):
    """2^e > mx (0 for a zero vector), as tc_exponent."""
    if not (mx > 0.0):
        return None
    e = int(np.floor(np.log2(mx))) + 1
    while 2.0 ** e <= mx:
        e += 1
    while 2.0 ** (e - 1) > mx:
        e -= 1
    return e


def split_vectors(X, s, conj=False):
    """X [R, K] complex: -> digits int8 [3, s, R, K] (component re, im, re+im; digit 0 most significant), ex int [R]."""
    R, K = X.shape
    dig = np.zeros((3, s, R, K), dtype=np.int8)
    ex = np.zeros(R, dtype=np.int64)
    for r in range(R):
        mx = float(np.max(np.abs(X[r].real) + np.abs(X[r].imag))) if K else 0.0
        e = exponent(mx)
        if e is None:
            continue
        ex[r] = e
        scale = 2.0 ** (8 * s - 2 - e)
        xr = np.rint(X[r].real * scale).astype(np.int64)
        xi = np.rint(X[r].imag * scale).astype(np.int64)
        if conj:
            xi = -xi
        for c, x in enumerate((xr, xi, xr + xi)):
            x = x.copy()
            for d in range(s - 1, -1, -1):
                low = ((x & 0xFF) ^ 0x80) - 0x80          # signed low byte
                dig[c, d, r] = low.astype(np.int8)
                x = (x - low) >> 8
            assert np.all(x == 0)
    return dig, ex


def reconstruct(dig, ex, s):
    """Value represented by the digits: [3, R, K] float (exact in long double for s <= 7)."""
    acc = np.zeros(dig.shape[0:1] + dig.shape[2:], dtype=np.longdouble)
    for d in range(s):
        acc = acc * 256 + dig[:, d].astype(np.longdouble)
    scale = np.exp2((ex + 2 - 8 * s).astype(np.longdouble))
    return acc * scale[None, :, None]


def gemm(A, B, s, opa="N", opb="N"):
    """C = op(A) op(B) by the digit scheme (levels 0 .. s-1 kept), complex128 [M, N]."""
    f = {"N": lambda x: x, "T": lambda x: x.T, "H": lambda x: x.conj().T}
    Ao, Bo = f[opa](A), f[opb](B)
    da, ea = split_vectors(np.ascontiguousarray(Ao), s)
    db, eb = split_vectors(np.ascontiguousarray(Bo.T), s)
    M, N = Ao.shape[0], Bo.shape[1]
    P = []
    for c in range(3):
        acc = np.zeros((M, N), dtype=np.float64)
        for d in range(s):
            S = np.zeros((M, N), dtype=np.int64)
            for p in range(d + 1):
                q = d - p
                S += da[c, p].astype(np.int64) @ db[c, q].astype(np.int64).T
            assert np.max(np.abs(S)) < 2 ** 31
            acc += S.astype(np.float64) * 2.0 ** (-8 * d)
        P.append(acc * np.exp2((ea[:, None] + eb[None, :] - 12).astype(np.float64)))
    return (P[0] - P[1]) + 1j * (P[2] - P[0] - P[1])


def simulate_schedule(ops, groups, s, ring, nkc=3, repeats=3):
    """Walk the op table the way the producer / issuer threads do; returns the multiset of (p, q, level) products of one
    K chunk per group and asserts the ring protocol cannot deadlock (every slot a load needs has been released by an MMA
    that does not itself depend on a later load)."""
    pairs = []
    # global sequences over `repeats` x groups x nkc iterations
    loads = []          # (kind, digit)
    mmas = []           # (ia, ib, level, first, relA, relB) with global load indices
    for _ in range(repeats):
        for g in groups:
            for kc in range(nkc):
                base = len(loads)
                seen = []
                for o in ops[g["op0"]:g["op0"] + g["nops"]]:
                    t = o & 3
                    if t in (0, 1):
                        loads.append((t, (o >> 2) & 15))
                        seen.append((t, (o >> 2) & 15))
                    else:
                        ia, ib = (o >> 2) & 31, (o >> 7) & 31
                        assert ia < len(seen) and ib < len(seen), "MMA uses a load that comes later in the table"
                        assert seen[ia][0] == 0 and seen[ib][0] == 1
                        lvl = (o >> 12) & 3
                        mmas.append((base + ia, base + ib, lvl, (o >> 14) & 1, (o >> 15) & 1, (o >> 16) & 1, len(loads)))
                        if kc == 0 and _ == 0:
                            pairs.append((g["d0"], seen[ia][1], seen[ib][1], lvl, (o >> 14) & 1))
                assert len(seen) == g["nloads"]
    # ring protocol: load i may be issued once load i - ring has been released; MMA j may be issued once its loads are issued.
    released_by = {}
    for j, m in enumerate(mmas):
        if m[4]:
            assert m[0] not in released_by
            released_by[m[0]] = j
        if m[5]:
            assert m[1] not in released_by
            released_by[m[1]] = j
    assert len(released_by) == len(loads), "every load must be released exactly once"
    issued_loads, done_mmas = 0, 0
    progress = True
    while progress:
        progress = False
        while issued_loads < len(loads) and (issued_loads < ring or released_by[issued_loads - ring] < done_mmas):
            issued_loads += 1
            progress = True
        while done_mmas < len(mmas) and max(mmas[done_mmas][0], mmas[done_mmas][1]) < issued_loads:
            # a released slot must not be used afterwards
            done_mmas += 1
            progress = True
    assert issued_loads == len(loads) and done_mmas == len(mmas), "ring deadlock"
    for j, m in enumerate(mmas):
        for idx in (m[0], m[1]):
            assert released_by[idx] >= j, "slot used after its release"
    return pairs


def simulate_compact(groups, ring, nkc=3, repeats=3):
    """Walk the compact per-role tables exactly as the kernel's producer and issuer warps do (slot cursors with wrap,
    acquire counts, releases); returns the (d0, p, q, level, first) products of one K chunk per group and asserts that
    every MMA only touches slots that hold the load it means, that are acquired, and that the ring cannot deadlock."""
    pairs = []
    loads = []                 # global load sequence: (is_b, digit)
    mmas = []                  # (load index a, load index b, relA, relB)
    acquired_total = 0
    for rep in range(repeats):
        for g in groups:
            for kc in range(nkc):
                base = len(loads)
                for w in g["loads"]:
                    loads.append((w & 1, w >> 1))
                acq_iter = 0
                for w in g["mmas"]:
                    ia, ib = w & 31, (w >> 5) & 31
                    acq_iter += (w >> 15) & 15
                    assert ia < acq_iter and ib < acq_iter, "operand not acquired"
                    assert ia < g["nloads"] and ib < g["nloads"]
                    la, lb = loads[base + ia], loads[base + ib]
                    assert la[0] == 0 and lb[0] == 1
                    lvl = (w >> 10) & 3
                    mmas.append((base + ia, base + ib, (w >> 13) & 1, (w >> 14) & 1))
                    if rep == 0 and kc == 0:
                        pairs.append((g["d0"], la[1], lb[1], lvl, (w >> 12) & 1))
                assert acq_iter == g["nloads"], "every load of an iteration is acquired inside it"
                acquired_total += acq_iter
    released_by = {}
    for j, m in enumerate(mmas):
        if m[2]:
            assert m[0] not in released_by
            released_by[m[0]] = j
        if m[3]:
            assert m[1] not in released_by
            released_by[m[1]] = j
    assert len(released_by) == len(loads)
    issued, done = 0, 0
    progress = True
    while progress:
        progress = False
        while issued < len(loads) and (issued < ring or released_by[issued - ring] < done):
            issued += 1
            progress = True
        while done < len(mmas) and max(mmas[done][0], mmas[done][1]) < issued:
            done += 1
            progress = True
    assert issued == len(loads) and done == len(mmas), "ring deadlock"
    for j, m in enumerate(mmas):
        assert released_by[m[0]] >= j and released_by[m[1]] >= j, "slot used after its release"
        # a slot must still hold this load when the MMA runs: the load that overwrites it (index + ring) cannot be issued before the release
    return pairs
