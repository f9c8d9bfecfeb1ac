"""Dictionary sharded over 2 GPUs must equal the unsharded result - with the exchange inside libkdi over
peer-mapped memory (the default) and with torch.distributed collectives, bit for bit the same.
Skipped with fewer than 2 GPUs."""

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist

    import kikuchipy_b200 as kb
    from oracle import di_oracle as orc

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        exp = orc.synthetic_experimental(300, (40, 40), seed=1).reshape(15, 20, 40, 40)
        dic = orc.synthetic_dictionary(7001, (40, 40), seed=2)
        nav = np.random.default_rng(3).random((15, 20)) < 0.2
        smask = orc.circular_signal_mask((40, 40))
        start, end = kb.shard_bounds(7001, world, rank)
        ctx = kb.default_context(rank)
        idx, sc = kb.dictionary_indexing_sharded(exp, dic[start:end], 7001, metric="ncc", keep_n=20,
                                                 navigation_mask=nav, signal_mask=smask, context=ctx)
        # exact ties across the candidate boundary and across shards: certificate -> exact rows -> merge
        dic2 = dic.copy()
        dic2[100:180] = dic2[7]
        dic2[5000:5040] = dic2[7]
        exp2 = np.clip(np.rint(dic2[[7, 9, 6000]] * 255), 0, 255).astype(np.uint8)
        idx2, sc2 = kb.dictionary_indexing_sharded(exp2, dic2[start:end], 7001, metric="ncc", keep_n=20, context=ctx)
        # keep_n beyond the candidate pipeline
        idx3, sc3 = kb.dictionary_indexing_sharded(exp[:2], dic[start:end], 7001, metric="ndp", keep_n=60, context=ctx)
        # dictionary generated on the device: every rank projects its own slice of the rotations
        from oracle import projection_oracle as po

        mu, ml = po.synthetic_master_pattern(201, seed=3)
        dc = po.direction_cosines_fixed_pc([-0.9, 0.85, -0.7, 0.95], 0.5, 40, 40, po.tilted_detector_matrix(70.0))
        rot = po.random_rotations(7001, seed=2)
        gen = kb.get_patterns(mu, ml, rot[start:end], direction_cosines=dc, detector_shape=(40, 40), context=ctx)
        idx4, sc4 = kb.dictionary_indexing_sharded(exp, gen, 7001, metric="ncc", keep_n=20, navigation_mask=nav,
                                                   signal_mask=smask, context=ctx)
        # both forms of the exchange and the single-GPU pipeline on a larger job (several strips and
        # row-block groups, kc = 64, pruned owner rescoring): bit-identical results
        from kikuchipy_b200 import _lib

        expB = orc.synthetic_experimental(2100, (30, 30), seed=11)
        dicB = orc.synthetic_dictionary(30001, (30, 30), seed=12)
        sB, eB = kb.shard_bounds(30001, world, rank)
        big = {}
        for ex in ("peer", "nccl", "peer"):  # (peer twice: the mapped blocks are reused)
            i5, s5 = kb.dictionary_indexing_sharded(expB, dicB[sB:eB], 30001, metric="ncc", keep_n=50, context=ctx,
                                                    exchange=ex)
            if ex in big:
                assert torch.equal(big[ex][0], i5) and torch.equal(big[ex][1], s5)
            big[ex] = (i5, s5)
        i6, s6 = ctx.dictionary_indexing(expB, 2100, dicB, 30001, _lib.KDI_NCC, 50)
        # device-resident shards held as VIEWS of the caller's rows (forced: a shard is too small for the
        # view to pay, so the default copies): the exchange kernels rescore from the raw rows, same bits
        d_exp, d_dic = torch.from_numpy(expB).cuda(), torch.from_numpy(np.ascontiguousarray(dicB[sB:eB])).cuda()
        try:
            for mode in (2, 1):
                ctx.set_option(_lib.OPT_DICT_VIEW, mode)
                for ex in ("peer", "nccl"):
                    i7, s7 = kb.dictionary_indexing_sharded(d_exp, d_dic, 30001, metric="ncc", keep_n=50, context=ctx,
                                                            exchange=ex)
                    assert torch.equal(big["peer"][0], i7) and torch.equal(big["peer"][1], s7), (mode, ex)
        finally:
            ctx.set_option(_lib.OPT_DICT_VIEW, 1)
        # strict certificate (KDI_OPT_CERT_STRICT = 1: worst-case bound, lists one size larger, 2 E pruning margin
        # in the owner rescoring of both exchanges): the same results as the single-GPU pipeline in either mode
        i_def, s_def = ctx.dictionary_indexing(expB, 2100, dicB, 30001, _lib.KDI_NCC, 20)
        ctx.set_option(_lib.OPT_CERT_STRICT, 1)
        try:
            assert ctx.candidate_capacity(20) == 64
            i_one, s_one = ctx.dictionary_indexing(expB, 2100, dicB, 30001, _lib.KDI_NCC, 20)
            assert np.array_equal(i_def, i_one) and np.array_equal(s_def, s_one)
            for ex in ("peer", "nccl"):
                i8, s8 = kb.dictionary_indexing_sharded(expB, dicB[sB:eB], 30001, metric="ncc", keep_n=20, context=ctx,
                                                        exchange=ex)
                assert np.array_equal(i8.cpu().numpy(), i_one) and np.array_equal(s8.cpu().numpy(), s_one), ex
        finally:
            ctx.set_option(_lib.OPT_CERT_STRICT, 2)
        np.savez(os.path.join(tmp, f"big{rank}.npz"), ip=big["peer"][0].cpu().numpy(), sp=big["peer"][1].cpu().numpy(),
                 inn=big["nccl"][0].cpu().numpy(), sn=big["nccl"][1].cpu().numpy(), i1=i6, s1=s6)
        np.savez(os.path.join(tmp, f"r{rank}.npz"), idx=idx.cpu().numpy(), sc=sc.cpu().numpy(),
                 idx2=idx2.cpu().numpy(), sc2=sc2.cpu().numpy(), idx3=idx3.cpu().numpy(), sc3=sc3.cpu().numpy(),
                 idx4=idx4.cpu().numpy(), sc4=sc4.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_two_gpu_shards_equal_unsharded(tmp_path):
    import torch
    import torch.multiprocessing as mp

    from oracle import di_oracle as orc

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    exp = orc.synthetic_experimental(300, (40, 40), seed=1).reshape(15, 20, 40, 40)
    dic = orc.synthetic_dictionary(7001, (40, 40), seed=2)
    nav = np.random.default_rng(3).random((15, 20)) < 0.2
    smask = orc.circular_signal_mask((40, 40))
    ridx, rsc = orc.dictionary_indexing(exp, dic, keep_n=20, navigation_mask=nav, signal_mask=smask)
    z0 = np.load(os.path.join(str(tmp_path), "r0.npz"))
    z1 = np.load(os.path.join(str(tmp_path), "r1.npz"))
    b0 = np.load(os.path.join(str(tmp_path), "big0.npz"))
    b1 = np.load(os.path.join(str(tmp_path), "big1.npz"))
    for k in ("ip", "sp", "inn", "sn"):
        assert np.array_equal(b0[k], b1[k]), k          # identical on every rank
    assert np.array_equal(b0["ip"], b0["inn"]) and np.array_equal(b0["sp"], b0["sn"])  # peer == collectives
    assert np.array_equal(b0["ip"], b0["i1"]) and np.array_equal(b0["sp"], b0["s1"])   # == one GPU
    assert np.array_equal(z0["idx"], z1["idx"]) and np.array_equal(z0["sc"], z1["sc"])
    r = orc.compare_topk(ridx, rsc, z0["idx"], z0["sc"], tie_tol=2e-5)
    assert r["tie_ok"] and r["scores_ok"], r
    assert list(z0["idx2"][0]) == [7] + list(range(100, 119))
    assert np.all(z0["sc2"][0] == z0["sc2"][0, 0]) and z0["idx2"][1, 0] == 9 and z0["idx2"][2, 0] == 6000
    assert np.array_equal(z0["idx2"], z1["idx2"])
    ridx3, rsc3 = orc.dictionary_indexing(exp[:2], dic, metric="ndp", keep_n=60, n_experimental_patterns=40)
    r = orc.compare_topk(ridx3, rsc3, z0["idx3"], z0["sc3"])
    assert r["tie_ok"] and r["scores_ok"], r
    # generated, sharded dictionary == the oracle on the oracle's projection of all rotations
    from oracle import projection_oracle as po

    mu, ml = po.synthetic_master_pattern(201, seed=3)
    dc = po.direction_cosines_fixed_pc([-0.9, 0.85, -0.7, 0.95], 0.5, 40, 40, po.tilted_detector_matrix(70.0))
    patterns = po.project_patterns(po.random_rotations(7001, seed=2), dc, mu, ml).reshape(7001, 40, 40)
    ridx4, rsc4 = orc.dictionary_indexing(exp, patterns, keep_n=20, navigation_mask=nav, signal_mask=smask)
    assert np.array_equal(z0["idx4"], z1["idx4"]) and np.array_equal(z0["sc4"], z1["sc4"])
    r = orc.compare_topk(ridx4, rsc4, z0["idx4"], z0["sc4"], tie_tol=2e-5)
    assert r["tie_ok"] and r["scores_ok"], r
