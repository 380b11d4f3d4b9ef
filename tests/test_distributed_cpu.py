"""world_size-2 gloo test of the multi-GPU host logic: stream sharding + one
all-reduce of the payoff sums.  The GPU shard computation is replaced by the
oracle's stream-convention pricer (tests may use the oracle), so this runs on CPU."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

import oracle_api as oa


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import hestonexotics_b200 as hx
    import oracle_api as oa_

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = oa_.Contract(oa_.ASIAN, [0.5, 1.0], [[95.0, 100.0], [105.0]], 24)

        import torch
        from hestonexotics_b200 import pricing

        def shard(rq, begin, count, world, stats):   # the oracle stands in for the GPU shard
            sm, sq = c.price_stream(int(rq.req.seed), int(rq.req.n_paths), int(rq.req.n_streams),
                                    begin, count)
            return torch.from_numpy(np.concatenate([sm, sq]))

        p = hx.HParams(*oa_.DEFAULT_PARAMS)
        chains = [hx.OptionsChain.from_strikes(0.5, [95.0, 100.0]),
                  hx.OptionsChain.from_strikes(1.0, [105.0])]
        rq = pricing._Request(hx.HQEAnderson(hx.AAsianCallNonAdaptive), p, 100.0, chains, 701, 3,
                              24, 3, "f32", 13)
        res = pricing._reduce_shards(rq, shard)   # what hx.price_distributed runs, GPU shard aside
        ret[rank] = (res.prices.tolist(), res.stderr.tolist(), res.sums.tolist())
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_equals_single_rank():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert set(ret.keys()) == {0, 1}
    assert ret[0] == ret[1]                      # every rank holds the same result
    c = oa.Contract(oa.ASIAN, [0.5, 1.0], [[95.0, 100.0], [105.0]], 24)
    sm, sq = c.price_stream(3, 701, 13)
    prices = np.array(ret[0][0])
    assert np.allclose(prices, sm / 701, rtol=1e-13)
    assert np.allclose(np.array(ret[0][2]), np.concatenate([sm, sq]), rtol=1e-13)
    assert (np.array(ret[0][1]) > 0).all()
