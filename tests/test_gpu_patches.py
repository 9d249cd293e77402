"""GPU: the device Patch-stream kernel (PT_FLAG_EMIT_PATCHES, csrc/patch_kernel.cuh) against the oracle's applyChange
return values — patch for patch, in order — on seeded fuzz logs (converged and unconverged replicas, zero-width marks,
typing runs, comments) and on the reference's four exact Patch KATs (test/micromerge.ts:915-1029)."""
import pytest

from oracle.oracle import Micromerge as O
from peritext_b200.packing import pack_logs, patch_stream
from tests.harness import fuzz_session, generateDocs, load_kats

pytestmark = pytest.mark.gpu
LIST = "1@doc1"


@pytest.fixture(scope="module")
def pengine():
    from peritext_b200.engine import BatchEngine
    e = BatchEngine(0, emit_patches=True)
    yield e
    e.close()


def list_ops(log):
    return [op for ch in log for op in ch["ops"] if op.get("obj") == LIST]


def oracle_patches(log):
    """Per list op: the patches the oracle's applyChange returned for it."""
    fresh = O("observer")
    out = []
    for ch in log:
        got = [p for p in fresh.applyChange(ch) if p["action"] != "makeList"]
        # split the change's patch list per op: re-apply op by op on a scratch replica is not possible, so compare per change
        out.append(got)
    return out


def device_patches_per_change(batch, dp, i, log):
    per_op = patch_stream(batch, dp, i, list_ops(log))
    out, k = [], 0
    for ch in log:
        n = sum(1 for op in ch["ops"] if op.get("obj") == LIST)
        out.append([p for ps in per_op[k:k + n] for p in ps]); k += n
    return out


def check_logs(pengine, logs):
    batch = pack_logs(logs)
    merged, dp = pengine.run_with_patches(batch)
    assert (merged.results["status"] == 0).all() and (dp.status == 0).all()
    for i, log in enumerate(logs):
        assert device_patches_per_change(batch, dp, i, log) == oracle_patches(log), f"log {i}"


@pytest.mark.parametrize("seed", range(12))
def test_device_patch_stream_equals_oracle_on_fuzz_logs(pengine, seed):
    _, logs, _ = fuzz_session(O, 7000 + seed, 120, sync_prob=0.6 if seed % 3 else 1.0,
                              zero_width_prob=0.1 if seed % 2 else 0.0, full_sync_at_end=bool(seed % 4), remove_comments=bool(seed % 2))
    check_logs(pengine, logs)


@pytest.mark.parametrize("seed", range(4))
def test_device_patch_stream_typing_runs(pengine, seed):
    _, logs, _ = fuzz_session(O, 7500 + seed, 150, replicas=2, max_chars=6, initial="The Peritext editor", sync_prob=0.5)
    check_logs(pengine, logs)


def test_patch_kats_on_the_device(pengine):
    for kat in [k for k in load_kats() if k["kind"] == "script" and any("expectPatches" in st for st in k["steps"])]:
        docs, _, init = generateDocs(O, kat["initialText"])
        logs = [[init], [init]]
        saved = {}
        for st in kat["steps"]:
            d = st["doc"] - 1
            if st["do"] == "change":
                ch = docs[d].change(st["ops"])["change"]
                logs[d].append(ch)
                if "save" in st:
                    saved[st["save"]] = ch
            elif st["do"] == "applyChange":
                docs[d].applyChange(saved[st["change"]])
                logs[d].append(saved[st["change"]])
                batch = pack_logs([logs[d]])
                merged, dp = pengine.run_with_patches(batch)
                per_change = device_patches_per_change(batch, dp, 0, logs[d])
                assert per_change[-1] == st["expectPatches"], kat["name"]


def test_large_logs_are_left_to_the_host(pengine):
    from peritext_b200 import workload
    batch = workload.generate("c2", n_docs=1, ops_per_doc=40000)
    merged, dp = pengine.run_with_patches(batch)
    assert (merged.results["status"] == 0).all() and (dp.status == 1).all()


def test_c4_shaped_batch_patch_items_fit_after_one_retry(pengine):
    from peritext_b200 import workload
    batch = workload.generate("c4", n_docs=20, ops_per_doc=400)
    merged, dp = pengine.run_with_patches(batch)
    assert (merged.results["status"] == 0).all() and (dp.status == 0).all()
    # every insert emits, deletes emit at most once per element
    ins = (batch.insdel["payload"] >> 30) == 0
    assert ((dp.recs["index"] >> 31)[ins] == 1).all()
