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


def _strong_worker(rank, world, port, n_docs_total, q):
    """bench.py's N > 1 layout: a fixed document set split with shard_range (uneven: 5 docs over 2 ranks), headers padded to
    the largest per-rank count, gathered into a caller-provided tensor."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.packed import replay_packed
    from peritext_b200 import sharding, workload
    first, count = sharding.shard_range(n_docs_total, rank, world)
    batch = workload.generate("c4", n_docs=count, ops_per_doc=300, doc_first=first, threads=1)
    R = batch.meta["replicas"]
    counts = [sharding.shard_range(n_docs_total, r, world)[1] * R for r in range(world)]
    merged, _ = replay_packed(batch)
    stage = torch.zeros(max(counts) * 32, dtype=torch.uint8)
    stage[: batch.n_logs * 32] = torch.from_numpy(merged.results.view(np.uint8).reshape(-1).copy())
    out = torch.empty(world * max(counts) * 32, dtype=torch.uint8)
    sharding.all_gather_results(stage, world, out=out)
    allh = sharding.headers_from_bytes(out.numpy()).reshape(world, max(counts))
    reps = [sharding.convergence_report(allh[r, : counts[r]], R) for r in range(world)]
    vis = [allh[r, : counts[r]]["n_visible"].tolist() for r in range(world)]
    q.put((rank, counts, vis, reps))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_strong_split_with_uneven_shards():
    world, n_docs = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_strong_worker, args=(r, world, port, n_docs, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1] == [9, 6] and out[0][2] == out[1][2]
    from oracle.packed import replay_packed
    from peritext_b200 import workload
    whole = workload.generate("c4", n_docs=n_docs, ops_per_doc=300, threads=1)
    ref, _ = replay_packed(whole)
    assert out[0][2][0] + out[0][2][1] == ref.results["n_visible"].tolist()      # the shards concatenate to the whole job
    for _, _, _, reps in out:
        assert all(r["all_status_ok"] and r["replicas_converged"] for r in reps)
