"""world_size-2 `gloo` test of the multi-GPU host logic on CPU: doc-id sharding is a partition, the all-gather of
result headers reassembles the single-process batch, and the convergence check fires on a corrupted digest.
(The per-rank merge itself is stood in for by the oracle here — no GPU in this suite.)"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, docs_per_gpu, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.packed import replay_packed
    from peritext_b200 import sharding, workload
    from peritext_b200.packing import RESULT_DT
    batch = workload.generate("c3", n_docs=docs_per_gpu, ops_per_doc=600, doc_first=sharding.weak_doc_first(docs_per_gpu, rank), threads=1)
    merged, _ = replay_packed(batch)
    headers = merged.results.copy()
    if rank == 1:
        headers[3]["digest"][0] ^= np.uint64(1)       # corrupt one replica's digest on rank 1
    local = torch.from_numpy(headers.view(np.uint8).reshape(-1).copy())
    gathered = sharding.all_gather_results(local, world)
    allh = gathered.numpy().view(RESULT_DT).reshape(world, -1)
    rep = sharding.convergence_report(allh, batch.meta["replicas"])
    q.put((rank, batch.n_logs, allh[:, :]["n_visible"].tolist(), rep))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_and_convergence_check():
    world, docs = 2, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, docs, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # both ranks see the same gathered table
    assert out[0][2] == out[1][2]
    # the gathered table equals a single-process generation of all docs (doc ids partition [0, world*docs))
    from oracle.packed import replay_packed
    from peritext_b200 import workload
    whole = workload.generate("c3", n_docs=world * docs, ops_per_doc=600, threads=1)
    ref, _ = replay_packed(whole)
    assert np.array(out[0][2]).reshape(-1).tolist() == ref.results["n_visible"].tolist()
    # the corrupted digest (rank 1, log 3 -> global doc docs + 3 // R) is reported by every rank
    R = whole.meta["replicas"]
    for _, _, _, rep in out:
        assert rep["all_status_ok"] and not rep["replicas_converged"]
        assert rep["diverged_docs"] == [docs + 3 // R]


def test_shard_range_is_a_partition():
    from peritext_b200.sharding import shard_range
    for n, w in [(10, 3), (100000, 8), (7, 8), (0, 4)]:
        seen = []
        for r in range(w):
            f, c = shard_range(n, r, w)
            seen += list(range(f, f + c))
        assert seen == list(range(n))
