"""CTC loss / gradient oracle in float64 numpy (alpha-beta over the blank-interleaved lattice).
TEST INFRASTRUCTURE ONLY.

Restates what the reference obtains from torch (trainer/trainer.py:76,167-173:
log_softmax + nn.CTCLoss(blank=0, zero_infinity=True), reduction 'mean') with the conventions
verified in SURVEY.md §A.4: loss = mean_b(nll_b / max(S_b, 1)); infeasible -> 0 loss / 0 grad;
d loss / d logits[b,t,:] = (softmax - occupancy) / (B * max(S_b,1)) for t < L_b, 0 for t >= L_b.
"""
import numpy as np


def _logsumexp(a, axis=None):
    m = np.max(a, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    with np.errstate(divide="ignore"):
        return np.squeeze(m, axis=axis) + np.log(np.sum(np.exp(a - m), axis=axis))


def ctc_loss_and_grad(logits, targets, input_lengths, target_lengths, blank=0):
    """logits (B,T,V); targets (B,Smax) ints; lengths (B,).  Returns loss, nll (B,), dlogits (B,T,V)."""
    logits = np.asarray(logits, dtype=np.float64)
    B, T, V = logits.shape
    lp = logits - _logsumexp(logits, axis=2)[..., None]
    nll = np.zeros(B)
    grad = np.zeros_like(logits)
    for b in range(B):
        L, S = int(input_lengths[b]), int(target_lengths[b])
        lab = [blank]
        for s in range(S):
            lab += [int(targets[b][s]), blank]
        n = len(lab)
        if L == 0:
            nll[b] = 0.0 if S == 0 else np.inf
            continue
        alpha = np.full((L, n), -np.inf)
        beta = np.full((L, n), -np.inf)
        alpha[0, 0] = lp[b, 0, lab[0]]
        if n > 1:
            alpha[0, 1] = lp[b, 0, lab[1]]
        for t in range(1, L):
            for s in range(n):
                c = [alpha[t - 1, s]]
                if s >= 1:
                    c.append(alpha[t - 1, s - 1])
                if s >= 2 and lab[s] != blank and lab[s] != lab[s - 2]:
                    c.append(alpha[t - 1, s - 2])
                alpha[t, s] = _logsumexp(np.array(c)) + lp[b, t, lab[s]]
        beta[L - 1, n - 1] = lp[b, L - 1, lab[n - 1]]
        if n > 1:
            beta[L - 1, n - 2] = lp[b, L - 1, lab[n - 2]]
        for t in range(L - 2, -1, -1):
            for s in range(n):
                c = [beta[t + 1, s]]
                if s + 1 < n:
                    c.append(beta[t + 1, s + 1])
                if s + 2 < n and lab[s] != blank and lab[s] != lab[s + 2]:
                    c.append(beta[t + 1, s + 2])
                beta[t, s] = _logsumexp(np.array(c)) + lp[b, t, lab[s]]
        tail = [alpha[L - 1, n - 1]] + ([alpha[L - 1, n - 2]] if n > 1 else [])
        nll[b] = -_logsumexp(np.array(tail))
        if not np.isfinite(nll[b]):
            continue
        occ = np.zeros((L, V))
        for s in range(n):
            with np.errstate(over="ignore"):
                occ[:, lab[s]] += np.exp(alpha[:, s] + beta[:, s] + nll[b] - lp[b, :L, lab[s]])
        grad[b, :L] = (np.exp(lp[b, :L]) - occ) / (B * max(S, 1))
    feasible = np.isfinite(nll)
    loss = float(np.sum(np.where(feasible, nll, 0.0) / np.maximum(np.asarray(target_lengths, dtype=np.float64), 1)) / B)
    grad[~feasible] = 0.0
    return loss, nll, grad
