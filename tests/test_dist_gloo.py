"""world_size-2 gloo test of the multi-GPU host logic (data-sharded covariance
sums + one flat SUM all-reduce, nsrunner_roi_replay.py:746-749) on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from nsgp_repre_b200 import dist as D
    from nsgp_repre_b200.covariance import CovarianceHooks, _LayerAcc
    from oracle import restated as O
    r, w = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # every rank accumulates its shard of 6 seeded batches into CPU stand-ins of the
    # accumulators; the reduce must equal the single-process sum
    g = torch.Generator().manual_seed(0)
    batches = [torch.randn(2, 4, 6, 7, generator=g) for _ in range(6)]
    mine = D.shard_batches(len(batches), rank, world)
    acc = torch.zeros(36, 36)
    for i in mine:
        acc += O.cov_conv2d(batches[i], (3, 3), (1, 1), (1, 1))
    hooks = CovarianceHooks(torch.nn.Identity())
    hooks._layers["conv.weight"] = _LayerAcc(None, acc.view(-1).clone())
    hooks._layers["other.weight"] = _LayerAcc(None, torch.full((5,), float(rank + 1)))
    hooks.all_reduce()
    total = sum(O.cov_conv2d(b, (3, 3), (1, 1), (1, 1)) for b in batches)
    ok = torch.allclose(hooks._layers["conv.weight"].acc.view(36, 36), total, rtol=1e-5, atol=1e-4)
    ok &= bool((hooks._layers["other.weight"].acc == 3.0).all())
    # variable-length gather (nsrunner_roi_replay.py:73-105): rank r holds r+2 rows
    from nsgp_repre_b200.rois import all_gather_different_shape, RoIHarvest
    mine_t = torch.arange((rank + 2) * 3, dtype=torch.float32).view(rank + 2, 3) + 100 * rank
    parts = all_gather_different_shape(mine_t)
    ok &= len(parts) == world
    for r in range(world):
        want = torch.arange((r + 2) * 3, dtype=torch.float32).view(r + 2, 3) + 100 * r
        ok &= bool(torch.equal(parts[r], want))
    # cal_rois tail: gather -> list of 6 in rank order
    h = RoIHarvest()
    n = rank + 2
    h.add(torch.full((n, 8), float(rank)), torch.full((n,), rank, dtype=torch.int64),
          torch.ones(n), torch.zeros(n, 4), torch.ones(n, 4), torch.zeros(n, 5))
    res = h.finish()
    ok &= len(res) == 6 and res[0].shape == (5, 8) and res[1].tolist() == [0, 0, 1, 1, 1]
    # EWC importance over ranks (not reduced in the reference; optional here)
    from nsgp_repre_b200.ewc import EWCImportance
    lin = torch.nn.BatchNorm1d(4)
    acc = EWCImportance({"bn.weight": lin.weight, "bn.bias": lin.bias})
    acc.importance["bn.weight"].fill_(float(rank + 1))
    acc.importance["bn.bias"].fill_(10.0 * (rank + 1))
    acc.all_reduce()
    ok &= bool((acc.importance["bn.weight"] == 1.5).all() and (acc.importance["bn.bias"] == 15.0).all())
    # data-parallel coarse prototypes: per-rank class sums -> one all-reduce -> global means
    from nsgp_repre_b200.roi_extract import reduce_class_sums
    gg = torch.Generator().manual_seed(5)
    F_all = torch.randn(12, 6, generator=gg)
    lab_all = torch.tensor([0, 1, 2, 0, 1, 1, 2, 2, 0, 0, 1, 3])
    sel = slice(0, 7) if rank == 0 else slice(7, 12)
    sums = torch.zeros(5, 6); cnt = torch.zeros(5, dtype=torch.int32)
    for f, l in zip(F_all[sel], lab_all[sel]):
        sums[l] += f; cnt[l] += 1
    means, tot = reduce_class_sums(sums, cnt)
    for c in range(4):
        ok &= bool(torch.allclose(means[c], F_all[lab_all == c].mean(0), atol=1e-6))
    ok &= tot.tolist() == [4, 4, 3, 1, 0] and bool(torch.isnan(means[4]).all())
    # layer-sharded projector build: every rank ends up with every owner's result
    dims = [12, 7, 9, 5]
    covs = []
    for i, d in enumerate(dims):
        a = torch.randn(d, d, generator=torch.Generator().manual_seed(40 + i))
        covs.append(a @ a.t())
    owners = D.shard_by_cost([float(d) ** 3 for d in dims], world)
    ok &= owners == [0, 1, 1, 1]                     # LPT: 1728 | 343 + 729 + 125
    calls = []

    def eig(i):
        calls.append(i)
        w, v = torch.linalg.eigh(covs[i].double())
        return w.flip(0).float(), v.flip(1).float()

    res = D.sharded_compute([((d,), (d, d)) for d in dims], [float(d) ** 3 for d in dims], eig,
                            "cpu")
    ok &= calls == [i for i in range(len(dims)) if owners[i] == rank]
    for i, (w, v) in enumerate(res):
        rebuilt = (v * w) @ v.t()
        ok &= bool(torch.allclose(rebuilt, covs[i], rtol=1e-4, atol=1e-4))
    # class-sharded prototype exchange: rows of class c go to rank (index of c) % W in
    # (source rank, original row) order; prototypes come back in class order
    from nsgp_repre_b200.prototypes import exchange_by_class_owner, gather_prototypes
    gg = torch.Generator().manual_seed(60 + rank)
    n_loc = 9 + 3 * rank
    lab_loc = torch.randint(0, 5, (n_loc,), generator=gg)        # class 4 = background here
    f_loc = lab_loc.float().unsqueeze(1) * 100 + rank * 10 + torch.arange(n_loc).float().unsqueeze(1) * 0.01
    f_loc = f_loc.repeat(1, 3)
    f_own, l_own, owned = exchange_by_class_owner(f_loc, lab_loc, [0, 1, 2, 3])
    ok &= owned == ([0, 2] if rank == 0 else [1, 3])
    both = [all_gather_different_shape(t) for t in (f_loc, lab_loc)]
    F_glob, L_glob = torch.cat(both[0]), torch.cat(both[1])
    sel = torch.zeros_like(L_glob, dtype=torch.bool)
    for c in owned:
        sel |= L_glob == c
    # same multiset AND, per class, the reference's gathered order
    for c in owned:
        ok &= bool(torch.equal(f_own[l_own == c], F_glob[L_glob == c]))
    ok &= int(sel.sum()) == f_own.shape[0]
    protos = torch.stack([f_own[l_own == c].mean(0) for c in owned])
    p_all, l_all = gather_prototypes(protos, torch.tensor(owned))
    ok &= l_all.tolist() == [0, 1, 2, 3]
    for c in range(4):
        ok &= bool(torch.allclose(p_all[c], F_glob[L_glob == c].mean(0)))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_covariance_allreduce_world2():
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
