"""world_size-2 test of the sharded path's host logic on CPU (gloo): slicing, the single all-reduce, epilogue.

The two compute callables are injected: the CPU oracle stands in for the CUDA kernels (allowed in tests/ only).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import aqlm_oracle as O


def _oracle_partial(x, codes, codebooks):
    y = O.dequantize_gemm(x.float().numpy(), codes.numpy(), codebooks.float().numpy(),
                          np.ones((codes.shape[0], 1, 1, 1), np.float32), None)
    return torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32))


def _oracle_epilogue(partial, scales, bias, dtype):
    y = partial * scales.float().reshape(1, -1)
    if bias is not None:
        y = y + bias.float()
    return y.to(dtype)


def _worker(rank, world, port, K, nbits, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from aqlm_b200.sharded import ShardedQuantizedLinear, shard_bounds

        case = O.make_case(4242, 512, 96, K, nbits, 8, 3, bias=True)
        t = {k: (None if v is None else torch.from_numpy(np.asarray(v))) for k, v in case.items()}
        calls = {"n": 0}
        orig = dist.all_reduce

        def counting_all_reduce(*a, **kw):
            calls["n"] += 1
            return orig(*a, **kw)

        dist.all_reduce = counting_all_reduce
        m = ShardedQuantizedLinear.from_full(t["codes"], t["codebooks"], t["scales"], t["bias"],
                                             partial_fn=_oracle_partial, epilogue_fn=_oracle_epilogue)
        b, e = shard_bounds(64, rank, world)
        assert m.codes.shape == (96, (e - b), K) and m.codes.is_contiguous()
        y_full_in = m(t["x"])                               # full activation: module slices it
        y_slice_in = m(t["x"][:, m.in_begin:m.in_end])      # pre-sliced activation
        dist.all_reduce = orig
        assert calls["n"] == 2, "exactly one all-reduce per forward"
        ref = O.dequantize_gemm(case["x"], case["codes"], case["codebooks"], case["scales"], case["bias"])
        e1 = O.relative_error(y_full_in.float().numpy(), ref)
        e2 = O.relative_error(y_slice_in.float().numpy(), ref)
        ret[rank] = (e1, e2)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("K,nbits", [(1, 16), (2, 8)])
def test_sharded_forward_world2_gloo(K, nbits):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), K, nbits, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        e1, e2 = ret[r]
        assert e1 < 5e-4 and e2 < 5e-4, (r, e1, e2)


def test_shard_bounds_and_divisibility():
    from aqlm_b200.sharded import shard_bounds, shard_codes

    assert [shard_bounds(1024, r, 8) for r in (0, 7)] == [(0, 128), (896, 1024)]
    assert shard_bounds(3584, 3, 4) == (2688, 3584)
    with pytest.raises(ValueError):
        shard_bounds(129, 0, 2)
    codes = torch.arange(4 * 8 * 2, dtype=torch.int16).reshape(4, 8, 2)
    parts = [shard_codes(codes, r, 4) for r in range(4)]
    assert all(p.is_contiguous() and p.shape == (4, 2, 2) for p in parts)
    assert torch.equal(torch.cat(parts, dim=1), codes)


def test_product_defaults_refuse_cpu():
    from aqlm_b200.sharded import ShardedQuantizedLinear

    m = ShardedQuantizedLinear(64, 16, 8, 1, 1, 16, bias=False, rank=0, world_size=1, dtype=torch.float16)
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 64, dtype=torch.float16))
