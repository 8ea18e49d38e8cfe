"""Oracle: prior push + PUCT select for one tree node.

Test infrastructure (see oracle/__init__.py).  Restates
  ChessPlayer.select_action_q_and_u /root/reference/src/chess_zero/agent/player_chess.py:251-299
and the virtual-loss / backup arithmetic its caller applies
  ChessPlayer.search_my_move        /root/reference/src/chess_zero/agent/player_chess.py:194-215

Arithmetic is IEEE binary64 in the reference's evaluation order (Python floats;
in the numpy-1.x era the reference was written for, float32 priors are promoted
to float64 by the first ``/=``).  One multiply/divide/add at a time, no fused
multiply-add:

    tot   = 1e-8 + sum_{a legal, in order} P[idx(a)]          (left-to-right)
    p_a   = P[idx(a)] / tot
    xx    = sqrt(sum_n + 1)
    p'_a  = (1 - eps) * p_a + eps * noise_i                   (root only)
    s_a   = q_a + ((c_puct * p'_a) * xx) / (1 + n_a)
    pick  = first a with s_a > best, best starting at -999    (strict '>')
"""
import numpy as np


def push_priors(policy_row, legal_label_idx):
    """player_chess.py:267-275.  policy_row: float32[1968]; returns float64 priors per legal move."""
    tot = 1e-8
    vals = []
    for j in legal_label_idx:
        v = float(policy_row[int(j)])
        vals.append(v)
        tot = tot + v
    return np.asarray([v / tot for v in vals], dtype=np.float64)


def push_priors_from_legal(policy_legal):
    """Same as push_priors when the caller has already gathered P[idx(a)] for the legal moves."""
    return push_priors(policy_legal, range(len(policy_legal)))


def select(n, q, p, sum_n, c_puct, noise=None, eps=0.0):
    """player_chess.py:277-299.  n:int[], q:float64[], p:float64[]; returns edge index or -1."""
    xx = float(np.sqrt(np.float64(int(sum_n) + 1)))
    best_s, best_a = -999.0, -1
    for i in range(len(n)):
        p_ = float(p[i])
        if noise is not None:
            p_ = (1 - eps) * p_ + eps * float(noise[i])
        b = float(q[i]) + c_puct * p_ * xx / (1 + int(n[i]))
        if b > best_s:
            best_s, best_a = b, i
    return best_a


def apply_virtual_loss(n, w, sum_n, a, vloss):
    """player_chess.py:199-202: returns updated (n_a, w_a, q_a, sum_n)."""
    n_a = int(n[a]) + vloss
    w_a = float(w[a]) + -vloss
    return n_a, w_a, w_a / n_a, int(sum_n) + vloss


def backup(n_a, w_a, sum_n, vloss, leaf_v):
    """player_chess.py:211-215."""
    n_a = n_a + (-vloss + 1)
    w_a = w_a + (vloss + leaf_v)
    return n_a, w_a, w_a / n_a, sum_n + (-vloss + 1)
