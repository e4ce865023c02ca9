"""GPU test of the multi-GPU partial-bucket exchange (SURVEY 8e shape 1; include/b2kyber.h: msm_buckets_dev /
msm_reduce_windows_dev / msm_finish_dev) on ONE device: the pairs are split into `world` shards, each shard's partial
buckets are produced by the product pipeline, the all-to-all is spelled out with tensor slices, and the fused
add + reduce, window sums and Horner must give the bytes of the single-call MSM and of the oracle."""
import pytest

from kyber_b200 import workload as wl
from kyber_b200.multi import msm_bucket_exchange, shard_bounds
from oracle import bls12381 as o

pytestmark = pytest.mark.gpu


def _setup(engine, n, tag):
    import torch
    dev = torch.device("cuda", 0)
    a = wl.prng_scalars(tag + "-a", n, o.R)
    s = wl.prng_scalars(tag, n, o.R)
    pts = engine.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
    d_s = torch.frombuffer(bytearray(wl.scalars_to_bytes(s)), dtype=torch.uint8).to(dev)
    d_p = torch.frombuffer(bytearray(pts), dtype=torch.uint8).to(dev)
    torch.cuda.synchronize()
    return a, s, d_s, d_p, dev


# window widths chosen so that the window count of the 127-bit split scalars, (127 + c) / c, is a multiple of the world size
@pytest.mark.parametrize("n,world,c_force", [(4096, 1, 0), (4099, 2, 8), (1 << 16, 4, 11), (1 << 18, 8, 16), (1 << 20, 2, 0)])
def test_bucket_exchange_matches_single_call(engine, n, world, c_force):
    import torch
    a, s, d_s, d_p, dev = _setup(engine, n, "b2k/xchg")
    want = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
    spans = [shard_bounds(n, world, r) for r in range(world)]
    n_max = max(hi - lo for lo, hi in spans)
    # all ranks must run ONE plan: take the plan of the largest shard and pin its window width for the others
    engine.set_msm_window(c_force)
    try:
        plan = engine.bls12381_g1_msm_bucket_plan(n_max)
        c, W, nb, eb = plan["c"], plan["W"], plan["buckets_per_window"], plan["bucket_bytes"]
        assert eb == 192 and nb == 1 << (c - 1) and W % world == 0 and (c_force == 0 or c == c_force)
        engine.set_msm_window(c)
        parts = []
        for lo, hi in spans:
            assert engine.bls12381_g1_msm_bucket_plan(hi - lo) == plan
            b = torch.empty(W * nb * eb, dtype=torch.uint8, device=dev)
            engine.bls12381_g1_msm_buckets_dev(hi - lo, d_s[32 * lo:].data_ptr(), d_p[96 * lo:].data_ptr(), b.data_ptr(), b.numel())
            parts.append(b.view(world, -1))                     # [owner][w_cnt * nb * eb]
        engine.synchronize()                                    # the engine runs on its own stream, torch on the default one
        w_cnt = W // world
        wsums = []
        for g in range(world):                                  # what ncclAllToAll leaves on rank g
            recv = torch.cat([p[g] for p in parts]).contiguous()
            ws = torch.empty(w_cnt * eb, dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()
            engine.bls12381_g1_msm_reduce_windows_dev(c, w_cnt, world, recv.data_ptr(), ws.data_ptr())
            wsums.append(ws)
        engine.synchronize()
        allws = torch.cat(wsums).contiguous()                   # what ncclAllGather leaves everywhere
        out = torch.zeros(96, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        engine.bls12381_g1_msm_finish_dev(c, W, allws.data_ptr(), out.data_ptr())
        engine.synchronize()
        assert bytes(out[:48].cpu().tolist()) == want
        engine.bls12381_g1_msm_finish_dev(c, W, allws.data_ptr(), out.data_ptr(), affine_out=True)
        engine.synchronize()
        assert bytes(out.cpu().tolist()) == o.g1_to_affine_bytes(o.g1_mul(wl.dot_mod(s, a, o.R)))
    finally:
        engine.set_msm_window(0)
    single = torch.zeros(48, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    engine.call_dev("b2k_bls12381_g1_msm_dev", n, d_s.data_ptr(), d_p.data_ptr(), single.data_ptr())
    engine.synchronize()
    assert bytes(single.cpu().tolist()) == want


def test_bucket_exchange_driver_world_1(engine):
    """kyber_b200.multi.msm_bucket_exchange without a process group (world 1): buckets -> reduce -> finish."""
    import torch
    n = 1 << 20
    a, s, d_s, d_p, dev = _setup(engine, n, "b2k/xchg1")
    plan = engine.bls12381_g1_msm_bucket_plan(n)
    c, W, nb, eb = plan["c"], plan["W"], plan["buckets_per_window"], plan["bucket_bytes"]
    buckets = torch.empty(W * nb * eb, dtype=torch.uint8, device=dev)
    ws = torch.empty(W * eb, dtype=torch.uint8, device=dev)
    out = torch.zeros(48, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def local_buckets():
        engine.bls12381_g1_msm_buckets_dev(n, d_s.data_ptr(), d_p.data_ptr(), buckets.data_ptr(), buckets.numel())
        return buckets

    def reduce_windows(recv, parts, w_cnt):
        engine.bls12381_g1_msm_reduce_windows_dev(c, w_cnt, parts, recv.data_ptr(), ws.data_ptr())
        return ws[:w_cnt * eb]

    def finish(allws):
        engine.bls12381_g1_msm_finish_dev(c, W, allws.data_ptr(), out.data_ptr())
        return out

    msm_bucket_exchange(local_buckets, reduce_windows, finish, W, nb, eb)
    engine.synchronize()
    assert bytes(out.cpu().tolist()) == o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
    with pytest.raises(Exception):                               # bucket buffer too small: refused, not overrun
        engine.bls12381_g1_msm_buckets_dev(n, d_s.data_ptr(), d_p.data_ptr(), buckets.data_ptr(), buckets.numel() - 1)


def test_dev_entry_points_report_data_errors_at_wait(engine):
    """*_dev calls only enqueue; an out-of-range scalar surfaces at b2k_wait (sticky status word), once."""
    import torch
    from kyber_b200.capi import B2KError
    n = 2048
    a, s, d_s, d_p, dev = _setup(engine, n, "b2k/xchg-bad")
    bad = d_s.clone()
    bad[32 * 7:32 * 8] = torch.frombuffer(bytearray(o.R.to_bytes(32, "big")), dtype=torch.uint8).to(dev)
    out = torch.zeros(48, dtype=torch.uint8, device=dev)
    plan = engine.bls12381_g1_msm_bucket_plan(n)
    buckets = torch.empty(plan["W"] * plan["buckets_per_window"] * plan["bucket_bytes"], dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    engine.wait()                                                # clean slate
    for call in (lambda sc: engine.call_dev("b2k_bls12381_g1_msm_dev", n, sc.data_ptr(), d_p.data_ptr(), out.data_ptr()),
                 lambda sc: engine.bls12381_g1_msm_buckets_dev(n, sc.data_ptr(), d_p.data_ptr(), buckets.data_ptr(), buckets.numel())):
        call(bad)                                                # accepted ...
        with pytest.raises(B2KError) as ei:
            engine.wait()                                        # ... reported here
        assert ei.value.code == -3
        engine.wait()                                            # cleared
        call(d_s)
        engine.wait()
    assert bytes(out.cpu().tolist()) == o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
