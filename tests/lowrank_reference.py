"""TEST INFRASTRUCTURE -- torch restatement of the low-rank adjoint the third-generation chain path computes
(iisan_b200/csrc/san_lr.cu).  Nothing here is product code.

The side-adapter network (CC/model/model.py:300-349) only ever leaves the width-d residual stream through rank-r
bottlenecks (AdapterBlock fc_down, CC/model/modules.py:113-116; fc o pre_fc, CC/model/model.py:340-347 -- two Linear layers
with nothing in between, i.e. one [E, d] matrix M = W_pre W_fc).  Writing beta_s = 1 - g_s (intra-modal towers) or 1
(inter-modal tower), pi(j, s) = prod_{k=s+1..j} beta_k and u_s = the gated hidden-state input of stage s:

    x_s      = sum_{j<=s} pi(s, j) (u_j) + sum_{j<s} pi(s, j) (r_j Wu_j^T + bu_j)          (r_j = relu(z_j))
    d x_s    = sum_{j>=s} pi(j, s) dz_j Wd_j                                               (Wd_A = M, dz_A = dL/dy)
    dz_s     = [z_s > 0] * sum_{j>s} pi(j, s) dz_j (Wd_j Wu_s)                             (rank space only)

so every parameter gradient is a combination of  G_j = dz^T h_j  (one pass over the cached states),  P = dz^T r  (rank
space) and column sums of dz; the gate gradients follow from the additive identity
    R_s = <d last_s, last_s> = R_{s-1} + g_s Q_s - <dWd_s, Wd_s> + <dWu_s, Wu_s> + <dbu_s, bu_s>,   Q_s = <d x_s, h_s>.
"""
import torch

TAU = 0.1      # CC/model/model.py:321


def tower_backward(kind, h, h2, gate_p, Wd, bd, Wu, bu, W_fc, b_fc, W_pre, b_pre, e):
    """One tower.  kind 'intra': x_s = g h_s + (1-g) last_{s-1};  'mm': x_s = last_{s-1} + g h_s + (1-g) h2_s.
    h / h2: lists of [N, d]; gate_p: list of scalar tensors; Wd[s] [r, d]; Wu[s] [d, r]; e = dL/dy [N, E].
    Returns (y, grads dict) with the same keys as the arguments."""
    A = len(Wd)
    g = [torch.sigmoid(p / TAU).reshape(()) for p in gate_p]
    beta = [(1 - g[s]) if kind == "intra" else torch.ones_like(g[s]) for s in range(A)] + [torch.ones_like(g[0])]

    def pi(j, s):                      # prod_{k=s+1..j} beta_k
        out = torch.ones_like(g[0])
        for k in range(s + 1, j + 1):
            out = out * beta[k]
        return out

    # ---- forward (sequential, as the chain kernel runs it); only r_s and y leave the kernel ----
    N, d = h[0].shape
    last = torch.zeros(N, d, dtype=h[0].dtype)
    r = []
    for s in range(A):
        u = g[s] * h[s] if kind == "intra" else g[s] * h[s] + (1 - g[s]) * h2[s]
        x = u + beta[s] * last
        z = torch.relu(x @ Wd[s].T + bd[s])
        r.append(z)
        last = x + z @ Wu[s].T + bu[s]
    M = W_pre @ W_fc
    c = W_pre @ b_fc + b_pre
    y = last @ M.T + c
    WdA = list(Wd) + [M]

    # ---- rank-space backward ----
    dz = [None] * (A + 1)
    dz[A] = e
    for s in range(A - 1, -1, -1):
        pre = sum(pi(j, s) * (dz[j] @ (WdA[j] @ Wu[s])) for j in range(s + 1, A + 1))
        dz[s] = pre * (r[s] > 0)

    # ---- reductions over the items ----
    def inp(j):            # (coefficient, hidden states) pairs that make up u_j
        return [(g[j], h[j])] if kind == "intra" else [(g[j], h[j]), (1 - g[j], h2[j])]
    GT = {}                # GT[(j, s, k)] = dz_s^T h^{(k)}_j  [r, d]  for s >= j
    for j in range(A):
        for s in range(j, A + 1):
            for k, (_, hh) in enumerate(inp(j)):
                GT[(j, s, k)] = dz[s].T @ hh
    P = {(j, s): dz[j].T @ r[s] for s in range(A) for j in range(s + 1, A + 1)}
    cs = [dz[s].sum(0) for s in range(A + 1)]

    # ---- combine ----
    dWd = []
    for s in range(A + 1):
        acc = sum(pi(s, j) * co * GT[(j, s, k)] for j in range(min(s, A - 1) + 1) for k, (co, _) in enumerate(inp(j)))
        for j in range(min(s, A)):
            acc = acc + pi(s, j) * (P[(s, j)] @ Wu[j].T + torch.outer(cs[s], bu[j]))
        dWd.append(acc)
    dWu = [sum(pi(j, s) * (WdA[j].T @ P[(j, s)]) for j in range(s + 1, A + 1)) for s in range(A)]
    dbu = [sum(pi(j, s) * (cs[j] @ WdA[j]) for j in range(s + 1, A + 1)) for s in range(A)]
    dM = dWd[A]
    grads = {"Wd": dWd[:A], "bd": cs[:A], "Wu": dWu, "bu": dbu,
             "W_pre": dM @ W_fc.T + torch.outer(cs[A], b_fc), "W_fc": W_pre.T @ dM, "b_pre": cs[A], "b_fc": W_pre.T @ cs[A]}
    # gates
    dgate = []
    if kind == "intra":
        R = torch.zeros((), dtype=h[0].dtype)
        for s in range(A):
            Q = sum(pi(j, s) * (GT[(s, j, 0)] * WdA[j]).sum() for j in range(s, A + 1))
            dgate.append((g[s] / TAU) * ((1 - g[s]) * Q - R))
            R = R + g[s] * Q - (dWd[s] * Wd[s]).sum() + (dWu[s] * Wu[s]).sum() + (dbu[s] * bu[s]).sum()
    else:
        for s in range(A):
            Q0 = sum((GT[(s, j, 0)] * WdA[j]).sum() for j in range(s, A + 1))
            Q1 = sum((GT[(s, j, 1)] * WdA[j]).sum() for j in range(s, A + 1))
            dgate.append(g[s] * (1 - g[s]) / TAU * (Q0 - Q1))
    grads["gate"] = dgate
    return y, grads
