"""Host-side logic of the strip decomposition (BASELINE config #5): how one large world is cut into vertical strips, one
per rank. Pure numpy / torch.distributed — no GPU needed, so it is covered by world_size-2 gloo tests on CPU."""
import numpy as np


def strip_edges(x_min, x_max, n_ranks):
    """n_ranks + 1 float32 edges; rank r owns snapshot x in [edges[r], edges[r+1]). The outer edges are infinite so every
    body has exactly one owner, and consecutive ranks share the SAME float32 value as their common edge."""
    e = np.linspace(np.float64(x_min), np.float64(x_max), n_ranks + 1).astype(np.float32)
    e[0], e[-1] = -np.inf, np.inf
    if n_ranks > 1 and not np.all(np.diff(e[1:-1].astype(np.float64)) > 0) and n_ranks > 2:
        raise ValueError("degenerate strips: extent too small for the number of ranks")
    return e


def owner_of(x, edges):
    """rank that owns snapshot x (vectorised); NaN goes to rank 0 like the device code."""
    x = np.asarray(x, dtype=np.float32)
    r = np.searchsorted(edges[1:-1], x, side="right")
    return np.where(np.isnan(x), 0, r).astype(np.int32)


def agree_edges(dist, local_x, n_ranks):
    """Every rank contributes the x-extent of the bodies it knows; all ranks end up with identical edges."""
    import torch

    lo = torch.tensor([float(np.min(local_x)) if len(local_x) else float("inf")], dtype=torch.float64)
    hi = torch.tensor([float(np.max(local_x)) if len(local_x) else float("-inf")], dtype=torch.float64)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return strip_edges(lo.item(), hi.item(), n_ranks)


def check_strip_width(edges, r_max):
    """Only adjacent ranks exchange ghosts, so an inner strip must be wider than the interaction reach."""
    inner = np.diff(edges[1:-1].astype(np.float64)) if len(edges) > 3 else np.array([np.inf])
    return bool(np.all(inner > 4.0 * r_max))
