"""Multi-GPU host logic: documents shard by doc id across ranks (no exchange during the merge); the one
collective is an all-gather of the 32-byte per-log result headers, after which every rank can run the convergence check
`digest[d][r] == digest[d][0]` (the fuzz harness's `deepStrictEqual(leftText, rightText)`, reference test/fuzz.ts:278)."""
from __future__ import annotations

import numpy as np

from .packing import RESULT_DT


def shard_range(n_docs_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous doc-id range [first, first+count) of `rank` (strong split of a fixed doc set)."""
    base, rem = divmod(n_docs_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def weak_doc_first(docs_per_gpu: int, rank: int) -> int:
    """Weak scaling: every rank holds `docs_per_gpu` documents with globally unique ids."""
    return rank * docs_per_gpu


def all_gather_results(local_headers, world: int, group=None, out=None):
    """`local_headers`: uint8 tensor [max_logs*32] (device tensor with NCCL, CPU tensor with gloo): the rank's result headers,
    zero-padded to the largest per-rank log count (a strong split leaves the first `rem` ranks one document more).  Returns
    the gathered uint8 tensor [world*max_logs*32] (`out` if given); enqueued on the current stream."""
    import torch
    import torch.distributed as dist
    if out is None:
        out = torch.empty(world * local_headers.numel(), dtype=torch.uint8, device=local_headers.device)
    dist.all_gather_into_tensor(out, local_headers, group=group)
    return out


def convergence_report(headers: np.ndarray, replicas: int) -> dict:
    """`headers`: RESULT_DT array [..., n_logs] with logs ordered doc-major (log = doc*R + r)."""
    h = headers.reshape(-1)
    if len(h) == 0:
        return {"all_status_ok": True, "replicas_converged": True, "diverged_docs": []}
    ok = bool((h["status"] == 0).all())
    dig = h["digest"].reshape(-1, replicas, 2)
    same = (dig == dig[:, :1, :]).all(axis=(1, 2))
    return {"all_status_ok": ok, "replicas_converged": bool(same.all()), "diverged_docs": np.nonzero(~same)[0].tolist()[:16]}


def headers_from_bytes(buf) -> np.ndarray:
    return np.frombuffer(bytes(buf), dtype=RESULT_DT) if not isinstance(buf, np.ndarray) else buf.view(RESULT_DT)
