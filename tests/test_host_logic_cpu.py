"""No-GPU tests of the host-side logic: execution plans, variable store / initialisers, batch sharding, and the
world_size-2 gloo path of the one collective (all_gather of per-cloud logits)."""
import math
import os
import socket
import sys

import numpy as np
import pytest
import torch

from common import make_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_capacities_and_cursors():
    from mccnn_b200 import runtime
    with runtime.record(slack=1.25) as plan:
        assert runtime.mode() == "record"
        plan.sizes.extend([1000, 1, 300000])
        plan.flags.extend([True, False])
    assert runtime.mode() == "exact"
    assert plan.capacities == [1280, 256, 375040]
    assert all(c >= s for c, s in zip(plan.capacities, plan.sizes))
    plan._cursor = plan._flag_cursor = 0
    assert [plan.next_capacity() for _ in range(3)] == plan.capacities
    assert [plan.next_flag(), plan.next_flag()] == [True, False]
    with pytest.raises(RuntimeError):
        plan.next_capacity()
    with pytest.raises(RuntimeError):
        plan.next_flag()
    t = torch.zeros(4)
    assert runtime.count_of(t) is None
    c = torch.tensor([3], dtype=torch.int32)
    assert runtime.count_of(runtime.tag(t, c)) is c


def test_variable_store_initialisers_follow_tensorflow():
    from mccnn_b200.variables import VariableStore
    vs = VariableStore(device="cpu", seed=1)
    w = vs.get_variable("w", [3, 64], "glorot")
    lim = math.sqrt(3.0 / ((3 + 64) / 2.0))               # variance_scaling(1.0, FAN_AVG, uniform)
    assert w.shape == (3, 64) and float(w.abs().max()) <= lim and float(w.abs().max()) > 0.8 * lim
    w3 = vs.get_variable("w3", [16, 8, 8], "glorot")        # fan_in = fan_out = 8 * 16 (receptive field 16)
    assert float(w3.abs().max()) <= math.sqrt(3.0 / 128.0)
    assert float(vs.get_variable("b", [7], "zeros").abs().sum()) == 0.0
    assert vs.get_variable("w", [3, 64]) is w               # reuse by name
    with pytest.raises(ValueError):
        vs.get_variable("w", [4, 64])
    sd = vs.state_dict()
    vs2 = VariableStore(device="cpu", seed=9)
    vs2.load_state_dict(sd)
    assert torch.equal(vs2.vars["w3"], w3)


def test_load_state_dict_overwrites_in_place_and_refreshes_the_folded_batch_norm():
    """Loading into a store that already ran: same storage (captured graphs / optimizer references stay valid), and
    the folded inference batch norm sees the new statistics (its cache is keyed on the tensors' version counters)."""
    from mccnn_b200.variables import VariableStore
    from mccnn_b200 import MCNetworkUtils as U
    vs = VariableStore(device="cpu", seed=0)
    scale0, shift0 = U._bn_affine(4, False, "L_BN", vs)
    mean = vs.vars["L_BN/moving_mean"]
    mean.requires_grad_(False)
    ptr = mean.data_ptr()
    vs.load_state_dict({"L_BN/moving_mean": torch.full((4,), 5.0), "L_BN/moving_variance": torch.full((4,), 4.0),
                        "brand_new": torch.ones(2)})
    assert vs.vars["L_BN/moving_mean"] is mean and mean.data_ptr() == ptr
    scale1, shift1 = U._bn_affine(4, False, "L_BN", vs)
    want = 1.0 / np.sqrt(4.0 + U.BN_EPSILON)
    assert np.allclose(scale1.numpy(), want) and np.allclose(shift1.numpy(), -5.0 * want)
    assert not np.allclose(scale0.numpy(), scale1.numpy())
    assert torch.equal(vs.vars["brand_new"], torch.ones(2))
    with pytest.raises(ValueError):
        vs.load_state_dict({"L_BN/moving_mean": torch.zeros(3)})


def test_shard_ranges_partition_the_batch():
    from mccnn_b200.engine import shard_range, shard_batch
    for B in (1, 7, 32, 33):
        for W in (1, 2, 3, 8):
            r = [shard_range(B, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    pts, bids, feats = make_batch(5, 100, "cube", seed=0, ragged=True)
    got = []
    for k in range(2):
        p, b, f, nb, (lo, hi) = shard_batch(pts, bids, feats, 5, k, 2)
        assert nb == hi - lo and b.min() == 0 and b.max() == nb - 1
        assert np.array_equal(p, pts[(bids[:, 0] >= lo) & (bids[:, 0] < hi)])
        got.append(p.shape[0])
    assert sum(got) == pts.shape[0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, B, k, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from mccnn_b200.engine import shard_batch, gather_logits
    from oracle import pipeline as OP
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pts, bids, feats = make_batch(B, 256, "sphere", seed=5)
    params = _params(k)
    p, b, f, nb, _ = shard_batch(pts, bids, feats, B, rank, world)
    local, _ = OP.mcclasss_forward(p, b, f, 1, nb, k, 10, params)        # this rank's clouds only
    full = gather_logits(torch.from_numpy(local), B, world)               # the single collective of the step
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _params(k, numCat=10, seed=0):
    """MCClassS parameter dictionary (reference variable names) for the oracle pipeline."""
    rng = np.random.default_rng(seed)
    p = {}

    def conv(name, nOut):
        nb = (nOut + 7) // 8
        p[name + "_weights"] = rng.uniform(-0.5, 0.5, (3, 8 * nb)).astype(np.float32)
        p[name + "_biases"] = rng.uniform(-0.1, 0.1, (8 * nb,)).astype(np.float32)
        for s in ("2", "3"):
            p[name + "_weights" + s] = rng.uniform(-0.5, 0.5, (nb, 8, 8)).astype(np.float32)
            p[name + "_biases" + s] = rng.uniform(-0.1, 0.1, (nb, 8)).astype(np.float32)

    def bn(name, c):
        p[name + "/gamma"] = rng.uniform(0.8, 1.2, (c,)).astype(np.float32)
        p[name + "/beta"] = rng.uniform(-0.1, 0.1, (c,)).astype(np.float32)
        p[name + "/moving_mean"] = rng.uniform(-0.1, 0.1, (c,)).astype(np.float32)
        p[name + "/moving_variance"] = rng.uniform(0.8, 1.2, (c,)).astype(np.float32)

    def dense(w, b, i, o):
        p[w] = rng.uniform(-0.3, 0.3, (i, o)).astype(np.float32)
        p[b] = rng.uniform(-0.1, 0.1, (o,)).astype(np.float32)

    conv("Conv_1", k)
    conv("Conv_2", 2 * k)
    conv("Conv_3", 4 * k)
    bn("Reduce_1_In_BN_BN", k)
    dense("Reduce_1_weights", "Reduce_1_biases", k, 2 * k)
    bn("Reduce_1_Out_BN_BN", 2 * k)
    bn("Reduce_2_In_BN_BN", 2 * k)
    dense("Reduce_2_weights", "Reduce_2_biases", 2 * k, 4 * k)
    bn("Reduce_2_Out_BN_BN", 4 * k)
    bn("BNRELUDROP_final_BN", 4 * k)
    bn("Final_Logits_BN_Init", 4 * k)
    dense("Final_Logits_weights1", "Final_Logits_biases1", 4 * k, 2 * k)
    bn("Final_Logits_BN_h1", 2 * k)
    dense("Final_Logits_weights2", "Final_Logits_biases2", 2 * k, k)
    bn("Final_Logits_BN_h2", k)
    dense("Final_Logits_weights3", "Final_Logits_biases3", k, numCat)
    return p


@pytest.mark.parametrize("B", [5, 4])  # unequal shards (padded gather) and equal shards (bare all_gather)
def test_sharded_forward_equals_unsharded_gloo_world2(B):
    """Clouds are independent units: computing each rank's clouds separately and all_gathering
    the logits reproduces the single-process result exactly.  Runs the real collective on the gloo backend."""
    import torch.multiprocessing as mp
    from oracle import pipeline as OP
    k, world = 8, 2
    pts, bids, feats = make_batch(B, 256, "sphere", seed=5)
    want, trace = OP.mcclasss_forward(pts, bids, feats, 1, B, k, 10, _params(k))
    assert want.shape == (B, 10) and trace[0]["levels"][-1] == B
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, B, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def _rooms_worker(rank, world, port, B, q):
    """Absolute-radius (MCSegScanNet-like) shard: local box -> global_box all-reduce -> level-1 Poisson sampling of this
    rank's rooms on the UNSHARDED grid; per-point results come back through gather_point_logits."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from mccnn_b200.engine import shard_batch, global_box, gather_point_logits
    from oracle import cpu as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pts, bids, feats = make_batch(B, 3000, "room", seed=9, ragged=True)
    p, b, f, nb, (lo, hi) = shard_batch(pts, bids, feats, B, rank, world)
    mn, mx = O.compute_aabb(p, b, nb, False)
    gmn, gmx = global_box(torch.from_numpy(mn), torch.from_numpy(mx))
    mn, mx = gmn.numpy(), gmx.numpy()
    r = 0.2
    keys, idx = O.sort_points_step1(p, b, mn, mx, nb, r, False)
    sp, sb, sf, cells = O.sort_points_step2(p, b, f, keys, idx, mn, mx, nb, r, False)
    qp, qb, qi = O.poisson_sampling(sp, sb, cells, mn, mx, r, nb, False)
    rows = np.concatenate([qp, (qb + lo).astype(np.float32)], axis=1)      # [M_r, 4], unequal M_r per rank
    full = gather_point_logits(torch.from_numpy(rows), world)
    if rank == 0:
        q.put((full.numpy(), mn[0].copy(), mx[0].copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_absolute_radius_hierarchy_equals_unsharded_gloo_world2():
    """SURVEY.md section 8e caveat: with relativeRadius=False the grid comes from the batch-global box, so shards must
    all-reduce it (engine.global_box) to select the same Poisson samples as the unsharded run."""
    import torch.multiprocessing as mp
    from oracle import cpu as O
    B, world, r = 3, 2, 0.2
    pts, bids, feats = make_batch(B, 3000, "room", seed=9, ragged=True)
    mn, mx = O.compute_aabb(pts, bids, B, False)
    keys, idx = O.sort_points_step1(pts, bids, mn, mx, B, r, False)
    sp, sb, sf, cells = O.sort_points_step2(pts, bids, feats, keys, idx, mn, mx, B, r, False)
    qp, qb, qi = O.poisson_sampling(sp, sb, cells, mn, mx, r, B, False)
    want = np.concatenate([qp, qb.astype(np.float32)], axis=1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rooms_worker, args=(k, world, port, B, q)) for k in range(world)]
    for p in procs:
        p.start()
    got, gmn, gmx = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(gmn, mn[0]) and np.array_equal(gmx, mx[0])       # the merged box IS the unsharded box
    assert got.shape == want.shape and np.array_equal(got, want)           # same samples, same order
