"""Design check of a kernel candidate (DESIGN.md section 9 -- evaluated, not pursued: the two lanes of a pair share one
instruction stream, so lane-specific work is issued for both and the chain only shrinks to ~0.8x): one backward joint step -- calc_aba's projection and
the congruence to the parent frame (loik-loid-optimized.hxx:60-75) -- split over TWO lanes per instance.

Lane L owns the columns [A; B^T] of H = [[A, B], [B^T, D]] and the linear halves of the force-like vectors, lane A the
columns [B; D] and the angular halves (B is held by both).  Each lane function below only touches its own data, the
batch-uniform constants (S, armature) and the joint transform (R, t), plus what it explicitly receives through `xchg`,
which stands for one `__shfl_xor_sync(.., 1)` of an fp64 value (= two 32-bit shuffles) and counts them.  The result must
equal the plain dense formulas (explicit 6x6 action matrices, as oracle A spells them), and the number of exchanged
scalars is what the cost estimate in DESIGN.md quotes.
"""
import numpy as np


def skew(t):
    return np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0.0]])


def dense_step(H, p, S, arm, r, R, t):
    """Plain formulas: P1 (calc_aba, general 1-DoF form) + P2 (SE3actOn) + SE3::act(Force)."""
    U = H @ S
    Dinv = 1.0 / (S @ U + arm)
    UD = U * Dinv
    r2 = r + S @ p
    Ha = H - np.outer(UD, U)
    pa = p - UD * r2
    X = np.block([[R, np.zeros((3, 3))], [skew(t) @ R, R]])  # dual action matrix of liMi = (R, t)
    return dict(UD=UD, Dinv=Dinv, r=r2, Hc=X @ Ha @ X.T, pc=X @ pa)


class Xchg:
    """Both lanes call xchg(mine) in lock-step and receive the partner's value; counts fp64 scalars per direction."""

    def __init__(self):
        self.count = 0

    def __call__(self, from_L, from_A):
        a, b = np.atleast_1d(np.asarray(from_L, float)), np.atleast_1d(np.asarray(from_A, float))
        assert a.shape == b.shape  # a shuffle moves one value in each direction
        self.count += a.size
        return from_A, from_L  # (what L receives, what A receives)


def two_lane_step(A, B, D, pl, pa, S, arm, r, R, t, xchg):
    Sl, Sa = S[:3], S[3:]
    tx = skew(t)
    # ---- U = H S and the two dot products: partial sums, one exchange of (S^T U, S^T p) partials
    Ul = A @ Sl + B @ Sa            # lane L
    Ua = B.T @ Sl + D @ Sa          # lane A
    dL, dA = np.array([Sl @ Ul, Sl @ pl]), np.array([Sa @ Ua, Sa @ pa])
    gotL, gotA = xchg(dL, dA)
    StU_L, Stp_L = dL + gotL        # both lanes now hold the same two scalars
    StU_A, Stp_A = gotA + dA
    Dinv_L, Dinv_A = 1.0 / (StU_L + arm), 1.0 / (StU_A + arm)
    r_L, r_A = r + Stp_L, r + Stp_A
    UDl, UDa = Ul * Dinv_L, Ua * Dinv_A
    # ---- projection H -= UDinv U^T, p -= UDinv r: the LA block needs the other half of U
    Ua_at_L, Ul_at_A = xchg(Ul, Ua)
    A1 = A - np.outer(UDl, Ul)                      # L
    B1_L = B - np.outer(UDl, Ua_at_L)               # L's copy of B
    B1_A = B - np.outer(Ul_at_A * Dinv_A, Ua)       # A's copy of B
    D1 = D - np.outer(UDa, Ua)                      # A
    pl1, pa1 = pl - UDl * r_L, pa - UDa * r_A
    # ---- congruence X* H X*^T:  A' = R A R^T,  B' = A' tx^T + R B R^T,  D' = R D R^T + tx A' tx^T + tx Rb + (tx Rb)^T
    A2 = R @ A1 @ R.T                               # L
    Rb_L, Rb_A = R @ B1_L @ R.T, R @ B1_A @ R.T     # both (duplicated work instead of 9 exchanges)
    iu = np.triu_indices(3)
    A2_at_A = np.zeros((3, 3))
    _, got = xchg(A2[iu], np.zeros(6))              # L sends the 6 unique entries of A'
    A2_at_A[iu] = got
    A2_at_A = A2_at_A + A2_at_A.T - np.diag(np.diag(A2_at_A))
    B2_L = A2 @ tx.T + Rb_L                         # L
    E_A = A2_at_A @ tx.T
    B2_A = E_A + Rb_A                               # A's copy of B'
    tRb = tx @ Rb_A
    D2 = R @ D1 @ R.T + tx @ E_A + tRb + tRb.T      # A
    # ---- p' = [R pl; R pa + t x (R pl)]
    Rpl = R @ pl1                                   # L
    _, Rpl_at_A = xchg(Rpl, np.zeros(3))
    pa2 = R @ pa1 + np.cross(t, Rpl_at_A)           # A
    return dict(UDl=UDl, UDa=UDa, Dinv=(Dinv_L, Dinv_A), r=(r_L, r_A), A=A2, B_L=B2_L, B_A=B2_A, D=D2, pl=Rpl, pa=pa2)


def _random_case(rng):
    M = rng.normal(size=(6, 6))
    H = M @ M.T + np.eye(6)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    R = q * np.sign(np.linalg.det(q))
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    S = np.concatenate([np.zeros(3), ax]) if rng.random() < 0.7 else np.concatenate([ax, np.zeros(3)])  # revolute / prismatic
    return H, rng.normal(size=6), S, float(rng.uniform(0.1, 10.0)), float(rng.normal()), R, rng.normal(size=3)


def test_two_lane_backward_step_equals_the_dense_formulas():
    rng = np.random.default_rng(0)
    for _ in range(50):
        H, p, S, arm, r, R, t = _random_case(rng)
        ref = dense_step(H, p, S, arm, r, R, t)
        x = Xchg()
        out = two_lane_step(H[:3, :3], H[:3, 3:], H[3:, 3:], p[:3], p[3:], S, arm, r, R, t, x)
        tol = 1e-10 * max(1.0, np.abs(ref["Hc"]).max())
        assert np.abs(np.concatenate([out["UDl"], out["UDa"]]) - ref["UD"]).max() < 1e-12 * max(1.0, np.abs(ref["UD"]).max())
        assert out["Dinv"][0] == out["Dinv"][1] and abs(out["Dinv"][0] - ref["Dinv"]) < 1e-14 * abs(ref["Dinv"]) + 1e-300
        assert out["r"][0] == out["r"][1] and abs(out["r"][0] - ref["r"]) < 1e-12 * max(1.0, abs(ref["r"]))
        assert np.abs(out["A"] - ref["Hc"][:3, :3]).max() < tol
        assert np.abs(out["B_L"] - ref["Hc"][:3, 3:]).max() < tol and np.abs(out["B_A"] - ref["Hc"][:3, 3:]).max() < tol
        assert np.abs(out["D"] - ref["Hc"][3:, 3:]).max() < tol
        assert np.abs(np.concatenate([out["pl"], out["pa"]]) - ref["pc"]).max() < 1e-10 * max(1.0, np.abs(ref["pc"]).max())
        # 2 (dot partials) + 3 (halves of U) + 6 (A') + 3 (R p_lin) fp64 values per direction and joint step
        assert x.count == 14
