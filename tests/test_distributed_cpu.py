"""world_size-2 gloo test of the sharding / gather host logic (CPU tensors stand in for the per-rank step outputs)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gymnasium_b200.distributed import BatchGather, shard_bounds


def test_shard_bounds_cover_range():
    for total in [1, 7, 8, 65536, 65537, 100003]:
        for world in [1, 2, 3, 8]:
            if total < world:
                continue
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


CASES = [(64, None), (65, None), (64, 0), (7, 1)]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for total, dst in CASES:
            start, count = shard_bounds(total, world, rank)
            # stand-in for this shard's step outputs: values that encode the GLOBAL env index
            idx = torch.arange(start, start + count)
            obs = torch.stack([idx.float(), idx.float() * 2, idx.float() * 3, idx.float() * 4], dim=1)
            reward = idx.double() + 0.5
            term = (idx % 3 == 0)
            g = BatchGather(total, world, rank, dst=dst)
            ok = True
            for _ in range(2):  # second call re-uses the cached buffers
                out = g(obs=obs, reward=reward, terminated=term)
                if dst is None or rank == dst:
                    full = torch.arange(total)
                    ok = ok and (torch.equal(out["obs"][:, 0], full.float())
                                 and torch.equal(out["obs"][:, 3], full.float() * 4)
                                 and torch.equal(out["reward"], full.double() + 0.5)
                                 and torch.equal(out["terminated"], full % 3 == 0)
                                 and out["terminated"].dtype == torch.bool)
            q.put((rank, total, dst, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2 * len(CASES))]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[-1] for r in res), res


# ---- HostBatch: the single host-side batch every rank lands its rows in (CPU: memmove stands in for the DMA engine) ------
def _apply_segments(segs):
    import ctypes

    for dst, src, dpitch, spitch, width, height in segs:
        for row in range(height):
            ctypes.memmove(dst + row * dpitch, src + row * spitch, width)


def _hostbatch_worker(rank, world, port, q):
    from gymnasium_b200.distributed import HostBatch

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total, depth, steps = 11, 2, 7  # ragged shards (6 + 5), fewer slots than steps: exercises the ack protocol
        bounds = [shard_bounds(total, world, r) for r in range(world)]
        start, n = bounds[rank]
        layout = [("obs", (n, 3), torch.float64), ("info", (4, n), torch.float64), ("reward", (n,), torch.float64),
                  ("terminated", (n,), torch.bool)]
        hb = HostBatch(layout, bounds, rank, tag="cpu_test", depth=depth, soa_keys=("info",), register=False)
        ok = True
        gi = np.arange(start, start + n)
        for k in range(steps):
            # this rank's "device" outputs of step k: values encode (step, global env index)
            obs = np.ascontiguousarray(np.stack([gi + 100.0 * k, gi * 2.0, gi * 3.0 + k], axis=1))
            info = np.ascontiguousarray(np.stack([gi + 0.25 * c + k for c in range(4)]))
            reward = gi * 0.5 + k
            term = ((gi + k) % 3 == 0)
            src = {"obs": obs, "info": info, "reward": reward, "terminated": term}
            hb.wait_writable(k)
            _apply_segments(hb.segments(k, rank, {key: a.ctypes.data for key, a in src.items()}))
            hb.publish_host(k)
            if rank == 0:  # the consumer lags one step behind, as the bench loop does
                if k >= 1:
                    ok = ok and _check_hostbatch(hb, k - 1, total)
        if rank == 0:
            ok = ok and _check_hostbatch(hb, steps - 1, total)
        dist.barrier()
        hb.close()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _check_hostbatch(hb, k, total):
    v = hb.wait_ready(k)
    g = np.arange(total)
    ok = (v["obs"].shape == (total, 3) and v["info"].shape == (4, total)
          and np.array_equal(v["obs"][:, 0], g + 100.0 * k) and np.array_equal(v["obs"][:, 2], g * 3.0 + k)
          and all(np.array_equal(v["info"][c], g + 0.25 * c + k) for c in range(4))
          and np.array_equal(v["reward"], g * 0.5 + k) and np.array_equal(v["terminated"], (g + k) % 3 == 0)
          and v["terminated"].dtype == np.bool_)
    hb.ack(k)
    return ok


def test_hostbatch_world2_lands_rows_key_major_and_throttles():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_hostbatch_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[-1] for r in res), res


def test_call_recorder_encodes_a_step_as_machine_words():
    """HostBatchPipeline's fast path replays a family's step from C: every argument of a recorded C-ABI call must be one
    machine word (pointer, integer, byref struct); anything else (a float by value) makes the family non-replayable."""
    import ctypes as C

    from gymnasium_b200.distributed import _CallRecorder

    class Cfg(C.Structure):
        _fields_ = [("a", C.c_double), ("b", C.c_int32)]

    seen = []
    proto = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)

    def impl(p0, p1, n, s):
        seen.append((p0, p1, n, s))
        return 0

    class FakeLib:
        b2e_fake_step = proto(impl)
        other = staticmethod(lambda: 7)

    lib = FakeLib()
    rec = _CallRecorder(lib)
    cfg, buf = Cfg(1.5, 3), (C.c_uint8 * 16)()
    assert rec.b2e_fake_step(C.byref(cfg), C.cast(buf, C.c_void_p), -1, None) == 0
    assert seen == []  # recording does not launch
    assert rec.other() == 7  # non-ABI attributes pass through
    (fn, words), = rec.calls
    assert fn == C.cast(lib.b2e_fake_step, C.c_void_p).value
    assert words == [C.addressof(cfg), C.addressof(buf), (1 << 64) - 1, 0]
    # replaying the words through the function pointer reaches the implementation with the same arguments
    replay = C.CFUNCTYPE(C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64)(fn)
    assert replay(*words) == 0 and seen[0][0] == C.addressof(cfg) and seen[0][2] == -1 and seen[0][3] is None
    with pytest.raises(TypeError):
        rec.b2e_fake_step(C.byref(cfg), 0.25, 1, None)
