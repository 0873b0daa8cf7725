"""Multi-GPU reduction path (needs >= 2 GPUs: `gpurun --gpus 2`): one process per GPU, contiguous
slices, local two-pass sums, NCCL all-gather of one scalar per rank, rank-ordered fold on the
device — bit-identical on every rank and equal to the order restated by the oracle."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _rank(rank: int, world: int, uid: bytes, uid_nccl: bytes, n: int, out_dir: str):
    import os
    sys.path.insert(0, str(ROOT))
    from custos_b200 import _native as N
    from custos_b200.raw import Comm, RawDevice, shard_range
    from custos_b200.workloads import CHEAP8

    dev = RawDevice(rank)
    comm = Comm(dev, world, rank, uid)  # fused reduce + exchange over NVLink peer memory
    os.environ["CB_COMM_P2P"] = "0"
    comm_nccl = Comm(dev, world, rank, uid_nccl)  # the fallback: ncclAllGather + fold
    del os.environ["CB_COMM_P2P"]
    assert not comm_nccl.uses_peer_memory
    x = np.random.default_rng(5).uniform(-1, 1, n).astype(np.float32)
    b, e = shard_range(n, 4, world, rank)
    local = x[b:e]
    p = dev.upload(local)
    out = dev.alloc(64)
    res = {}
    for name in ("sum", "mean"):
        if name == "sum":
            comm.sum_into(N.F32, p, local.size, out)
        else:
            comm.mean_into(N.F32, p, local.size, n, out)
        res[name] = dev.d2h(out, 1, N.F32)[0]
    again = []
    for _ in range(5):  # run-to-run determinism
        comm.sum_into(N.F32, p, local.size, out)
        again.append(dev.d2h(out, 1, N.F32)[0].tobytes())
    assert len(set(again)) == 1 and again[0] == res["sum"].tobytes()
    comm_nccl.sum_into(N.F32, p, local.size, out)  # both exchange paths fold in rank order: same bits
    assert dev.d2h(out, 1, N.F32)[0].tobytes() == res["sum"].tobytes()
    # an empty slice still takes part in the exchange
    comm.sum_into(N.F32, p, local.size if rank else 0, out)
    part0 = dev.d2h(out, 1, N.F32)[0]
    np.save(f"{out_dir}/p2p{rank}.npy", np.array([1.0 if comm.uses_peer_memory else 0.0, float(part0)]))
    # other accumulator widths through both exchange paths: f64 (double), bf16 (f32 accumulator), i16 (i64)
    from custos_b200.expr import bf16_from_f32
    extra = {}
    for dt, data in ((N.F64, x.astype(np.float64)), (N.BF16, bf16_from_f32(x)), (N.I16, (x * 30000).astype(np.int16))):
        loc = np.ascontiguousarray(data[b:e])
        pd = dev.upload(loc)
        got = []
        for c in (comm, comm_nccl):
            dev.clear(N.U8, out, 64)
            c.sum_into(dt, pd, loc.size, out)
            got.append(dev.d2h(out, 8, N.U8).tobytes())
        assert got[0] == got[1], f"peer-memory and NCCL exchange disagree for dtype {dt}"
        extra[dt] = got[0]
        dev.free(pd)
    np.save(f"{out_dir}/extra{rank}.npy", np.frombuffer(b"".join(extra[k] for k in sorted(extra)), np.uint8))
    comm_nccl.close()
    # element-wise work on the slice: no communication
    q = dev.alloc(local.nbytes)
    dev.apply(dev.compile(CHEAP8, N.F32), p, q, local.size)
    np.save(f"{out_dir}/out{rank}.npy", dev.d2h(q, local.size, N.F32))
    np.save(f"{out_dir}/r{rank}.npy", np.array([res["sum"], res["mean"]], np.float32))
    comm.close()
    dev.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_sum_over_nccl(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from custos_b200 import _native as N
    from custos_b200.raw import Comm, shard_range, sum_plan
    from custos_b200.workloads import CHEAP8
    from oracle import oracle as orc
    n = (1 << 24) + 1001
    uid, uid_nccl = Comm.unique_id(), Comm.unique_id()
    mp.spawn(_rank, args=(world, uid, uid_nccl, n, str(tmp_path)), nprocs=world, join=True)
    x = np.random.default_rng(5).uniform(-1, 1, n).astype(np.float32)
    partials = []
    for r in range(world):
        b, e = shard_range(n, 4, world, r)
        plan = sum_plan(N.F32, e - b)
        partials.append(orc.sum_two_pass(orc.F32, x[b:e], plan["blocks"], plan["chunk"], plan["threads"], plan["vec"], plan["threads2"]))
    want = partials[0]
    for p in partials[1:]:
        want = np.float32(want + p)
    rows = [np.load(tmp_path / f"r{r}.npy") for r in range(world)]
    for r in range(world):
        assert rows[r][0].tobytes() == want.tobytes(), (r, rows[r][0], want)
        assert rows[r][1].tobytes() == np.float32(want / np.float32(n)).tobytes()
    assert abs(float(want) - orc.sum_f64(orc.F32, x)) <= 1e-6 * float(np.sum(np.abs(x.astype(np.float64))))
    p2p = [np.load(tmp_path / f"p2p{r}.npy") for r in range(world)]
    assert all(v[0] == 1.0 for v in p2p), "the peer-memory exchange was not used (CUDA IPC unavailable?)"
    want_wo0 = np.float32(0)
    for p in partials[1:]:
        want_wo0 = np.float32(want_wo0 + p)
    assert all(np.float32(v[1]) == want_wo0 for v in p2p)
    # the other dtypes: rank-ordered fold of the per-slice oracle partials (floats), exact i64 sum (integers)
    from custos_b200.expr import bf16_from_f32
    datas = {N.F64: x.astype(np.float64), N.BF16: bf16_from_f32(x), N.I16: (x * 30000).astype(np.int16)}
    want_extra = {}
    for dt, data in datas.items():
        if dt == N.I16:
            want_extra[dt] = np.int64(data.astype(np.int64).sum()).tobytes()
            continue
        acc = None
        for r in range(world):
            b, e = shard_range(n, 4, world, r)  # the ranks slice every dtype at the f32 bounds
            plan = sum_plan(dt, e - b)
            part = orc.sum_two_pass(dt, data[b:e], plan["blocks"], plan["chunk"], plan["threads"], plan["vec"], plan["threads2"])
            acc = part if acc is None else type(part)(acc + part)
        want_extra[dt] = acc.tobytes().ljust(8, b"\0")
    for r in range(world):
        raw = np.load(tmp_path / f"extra{r}.npy").tobytes()
        for k, dt in enumerate(sorted(datas)):
            assert raw[8 * k:8 * k + 8] == want_extra[dt], (r, dt)
    full = np.concatenate([np.load(tmp_path / f"out{r}.npy") for r in range(world)])
    assert np.array_equal(full.view(np.uint32), orc.apply_chain(CHEAP8, orc.F32, x).view(np.uint32))


# ------------------------------------------------------------------ the sharded device behind the operator API
def _sharded_rank(rank: int, world: int, uid: bytes, n: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    from custos_b200.device import CUDA
    from custos_b200.workloads import CHAIN8, CHAIN8_GRADS

    x = np.random.default_rng(11).uniform(-4, 4, n).astype(np.float32)
    y = np.random.default_rng(12).uniform(-1, 1, n).astype(np.float32)
    with CUDA("Lazy", "Graph", "Autograd", "Base", ordinal=rank).shard(world, rank, uid) as dev:
        bx, by = dev.buffer_sharded(x).require_grad(), dev.buffer_sharded(y)
        begin, end, glen = bx.shard()
        assert glen == n and len(bx) == end - begin
        cur = bx
        for f, g in zip(CHAIN8, CHAIN8_GRADS):
            cur = dev.unary_ew(cur, f, g)
        assert cur.shard() == (begin, end, n)  # results inherit the slice
        prod = dev.mul(cur, by)
        dev.unary_fusing()
        dev.run()
        cur.backward()
        np.save(f"{out_dir}/s_out{rank}.npy", cur.replace().read())
        np.save(f"{out_dir}/s_grad{rank}.npy", bx.grad().read())
        np.save(f"{out_dir}/s_prod{rank}.npy", prod.replace().read())
        res = np.array([dev.sum(prod.replace()), dev.mean(prod.replace()), dev.sum(by), dev.sum(bx.grad())], np.float32)
        np.save(f"{out_dir}/s_red{rank}.npy", res)
        np.save(f"{out_dir}/s_range{rank}.npy", np.array([begin, end]))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("world", [2, 8])
def test_sharded_device_reassembles_to_the_single_device_bits(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from custos_b200 import _native as N
    from custos_b200.device import CUDA
    from custos_b200.raw import Comm, shard_range, sum_plan
    from custos_b200.workloads import CHAIN8, CHAIN8_GRADS
    from oracle import oracle as orc
    n = (1 << 22) + 777
    mp.spawn(_sharded_rank, args=(world, Comm.unique_id(), n, str(tmp_path)), nprocs=world, join=True)
    x = np.random.default_rng(11).uniform(-4, 4, n).astype(np.float32)
    y = np.random.default_rng(12).uniform(-1, 1, n).astype(np.float32)
    with CUDA("Lazy", "Graph", "Autograd", "Base") as dev:  # the same program on ONE device
        bx, by = dev.buffer(x).require_grad(), dev.buffer(y)
        cur = bx
        for f, g in zip(CHAIN8, CHAIN8_GRADS):
            cur = dev.unary_ew(cur, f, g)
        prod = dev.mul(cur, by)
        dev.unary_fusing()
        dev.run()
        cur.backward()
        one = {"out": cur.replace().read(), "grad": bx.grad().read(), "prod": prod.replace().read()}
    ranges = [np.load(tmp_path / f"s_range{r}.npy") for r in range(world)]
    assert [tuple(r) for r in ranges] == [shard_range(n, 4, world, r) for r in range(world)]
    for name in ("out", "grad", "prod"):
        full = np.concatenate([np.load(tmp_path / f"s_{name}{r}.npy") for r in range(world)])
        assert full.view(np.uint32).tolist() == one[name].view(np.uint32).tolist(), name
    # reductions: rank-ordered fold of the per-slice oracle partials, identical on every rank
    reds = [np.load(tmp_path / f"s_red{r}.npy") for r in range(world)]
    for r in range(1, world):
        assert reds[r].tobytes() == reds[0].tobytes()

    def want_sum(data):
        acc = None
        for r in range(world):
            b, e = shard_range(n, 4, world, r)
            plan = sum_plan(N.F32, e - b)
            part = orc.sum_two_pass(orc.F32, data[b:e], plan["blocks"], plan["chunk"], plan["threads"], plan["vec"], plan["threads2"])
            acc = part if acc is None else np.float32(acc + part)
        return acc
    assert reds[0][0].tobytes() == want_sum(one["prod"]).tobytes()
    assert reds[0][1].tobytes() == np.float32(want_sum(one["prod"]) / np.float32(n)).tobytes()
    assert reds[0][2].tobytes() == want_sum(y).tobytes()
    assert reds[0][3].tobytes() == want_sum(one["grad"]).tobytes()


# ------------------------------------------------------------------ a peer that misses an exchange is an error
def _late_rank(rank: int, world: int, uid: bytes, out_dir: str):
    import os
    import time
    sys.path.insert(0, str(ROOT))
    os.environ["CB_COMM_TIMEOUT_MS"] = "300"
    from custos_b200 import CustosError
    from custos_b200 import _native as N
    from custos_b200.raw import Comm, RawDevice

    dev = RawDevice(rank)
    comm = Comm(dev, world, rank, uid)
    x = np.full(1000, rank + 1, np.float32)
    p = dev.upload(x)
    first = comm.sum(N.F32, p, x.size)  # everybody on time
    if rank == 1:
        time.sleep(2.5)  # rank 1 reaches the second exchange long after rank 0 gave up
    status = "ok"
    try:
        second = comm.sum(N.F32, p, x.size)
    except CustosError as err:
        status, second = f"error {err.code}", np.float32(np.nan)
    later = "ok"
    try:
        comm.check()
    except CustosError as err:
        later = f"error {err.code}"
    np.save(f"{out_dir}/late{rank}.npy", np.array([first, second], np.float32))
    Path(f"{out_dir}/late{rank}.txt").write_text(f"{status}|{later}|{int(comm.uses_peer_memory)}")
    if rank == 0:
        time.sleep(3.5)  # keep the exchange buffer mapped until the late peer has written into it
    comm.close()
    dev.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_a_peer_that_misses_the_exchange_is_reported_not_swallowed(tmp_path):
    from custos_b200 import _native as N
    from custos_b200.raw import Comm
    mp.spawn(_late_rank, args=(2, Comm.unique_id(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (tmp_path / "late0.txt").read_text().split("|"), (tmp_path / "late1.txt").read_text().split("|")
    assert r0[2] == "1" and r1[2] == "1", "peer-memory exchange not in use"
    v0, v1 = np.load(tmp_path / "late0.npy"), np.load(tmp_path / "late1.npy")
    assert v0[0] == v1[0] == 3000.0
    assert r0[0] == f"error {N.CB_ERR_STATE}" and r0[1] == f"error {N.CB_ERR_STATE}"  # sticky: the sum was NOT returned as OK
    assert r1[0] == "ok" and v1[1] == 3000.0  # rank 0 had published its total before it gave up
