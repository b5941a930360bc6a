"""Numerical study for the round-2 GEMM path (CPU, numpy): fp64 complex GEMM emulated on INTEGER tensor cores
(Ozaki scheme: row/column-scaled operands split into s slices of beta bits, every slice product is an exact
int8 x int8 -> int32 GEMM, recombined in fp64; complex product by the 3-multiplication formula).  Prints the relative
Frobenius error against fp64 on the matrices of the RCWA path (order 7 Example1 cell) for s = 3..9 and beta = 6, 7.
B200: tcgen05 kind::i8 runs at 4.5 PFLOP/s dense against 40 TFLOP/s for fp64 -- see DESIGN.md section 6."""
import numpy as np, sys, torch
sys.path.insert(0,'/root/repo')
def split_rows(A, s, beta):
    # per-row scaling: A[i,:] = 2^e_i * sum_p S_p[i,:] 2^{-beta (p+1)},  S_p integer in [-2^beta, 2^beta]
    e = np.ceil(np.log2(np.maximum(np.abs(A).max(axis=1), 1e-300)))[:, None]
    R = A / 2.0**e                     # |R| <= 1
    slices = []
    for p in range(s):
        S = np.trunc(R * 2.0**beta)    # integer part with beta bits
        slices.append(S)
        R = R * 2.0**beta - S
    return e, slices
BETA = 6


def ozaki_gemm(A, B, s, beta=None):
    beta = BETA if beta is None else beta
    ea, SA = split_rows(A, s, beta)
    eb, SB = split_rows(B.T, s, beta)
    C = np.zeros((A.shape[0], B.shape[1]))
    ngemm = 0
    for p in range(s):
        for q in range(s - p):         # keep terms with p + q <= s - 1
            C += (SA[p] @ SB[q].T) * 2.0**(-beta * (p + q + 2))     # exact in int32 for K*2^(2 beta) < 2^31
            ngemm += 1
    return C * 2.0**ea * 2.0**eb.T, ngemm
def cgemm3m(A, B, s):
    P1, n1 = ozaki_gemm(A.real, B.real, s); P2, _ = ozaki_gemm(A.imag, B.imag, s); P3, _ = ozaki_gemm(A.real + A.imag, B.real + B.imag, s)
    return (P1 - P2) + 1j * (P3 - P1 - P2), 3 * n1
from oracle import cases as C
from oracle.rcwa_oracle import OracleSim
cd = torch.complex128
case = dict(C.CASES["ex1_o15"]); case["order"] = [7, 7]
sim = OracleSim(freq=C.freq_of(case, cd), order=case["order"], L=case["L"], dtype=cd)
sim.add_input_layer(eps=case["eps_in"]); sim.set_incident_angle(0., 0.)
d, e = C.build_layers(case, cd)[0]; sim.add_layer(d, e)
P, Q = sim.P[0].numpy(), sim.Q[0].numpy()
W = sim.E_eigvec[0].numpy()
rng = np.random.default_rng(0)
U = np.linalg.qr(rng.standard_normal((64, 64)) + 1j * rng.standard_normal((64, 64)))[0]
tests = {"P@Q (n=450)": (P, Q), "Q@W": (Q, W), "Z[:,win]@U (unitary update)": (np.linalg.qr(rng.standard_normal((450, 450)) + 1j * rng.standard_normal((450, 450)))[0][:, :64], U),
         "S-matrix product": (sim.layer_S[0][0].numpy(), sim.layer_S[0][1].numpy())}
for BETA in (6, 7):
    print("beta = %d bits per slice (exact int32 accumulation needs 2 beta + log2(K * terms) <= 31)" % BETA)
    for name, (A, B) in tests.items():
        ref = A @ B
        row = []
        for s in (3, 4, 5, 6, 7, 8, 9):
            Cc, ng = cgemm3m(A, B, s)
            row.append("s=%d (%d): %.1e" % (s, ng, np.linalg.norm(Cc - ref) / np.linalg.norm(ref)))
        print("  %-30s" % name, "; ".join(row))
