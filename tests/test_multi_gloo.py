"""CPU test of the N>1 path: 2 processes, gloo backend, the sharding/all-gather/combine logic of
kyber_b200.multi with the oracle standing in for the device kernels (the GPU arithmetic itself is covered by
the -m gpu tests; this covers the distributed plumbing that bench.py --gpus N uses)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    from kyber_b200.multi import shard_bounds
    for n in (0, 1, 7, 8, 1000, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kyber_b200 import workload as wl
    from kyber_b200.multi import shard_bounds, msm_sharded
    from oracle import bls12381 as o, cpu_ref
    lib = cpu_ref.load()
    a = wl.prng_scalars("b2k/gloo-a", n, o.R)
    s = wl.prng_scalars("b2k/gloo-s", n, o.R)
    lo, hi = shard_bounds(n, world, rank)
    comp = cpu_ref.g1_mul_batch(lib, wl.scalars_to_bytes(a[lo:hi]), wl.G1_BLS12381_AFFINE * (hi - lo), 1)
    pts = b"".join(o.g1_to_affine_bytes(o.g1_decompress(comp[48 * i:48 * i + 48], False)) for i in range(hi - lo))

    def local_partial():
        c48 = cpu_ref.g1_msm_pippenger(lib, wl.scalars_to_bytes(s[lo:hi]), pts, 1)
        return torch.frombuffer(bytearray(o.g1_to_affine_bytes(o.g1_decompress(c48, False))), dtype=torch.uint8)

    def combine(gathered, w):
        parts = bytes(gathered.tolist())
        acc = None
        for r in range(w):
            acc = o.g1_add(acc, o.g1_from_affine_bytes(parts[96 * r:96 * r + 96]))
        return torch.frombuffer(bytearray(o.g1_compress(acc)), dtype=torch.uint8)

    got = bytes(msm_sharded(local_partial, combine).tolist())
    want = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
    q.put((rank, got == want))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_msm_two_ranks_gloo():
    world, n = 2, 301
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _worker_exchange(rank, world, port, n, q):
    """shape 1: partial-bucket exchange (all-to-all + fused add/reduce + all-gather + Horner); the oracle stands in for
    the three device steps, with 4-bit unsigned windows (W = 64, 15 buckets per window) and 96-byte affine buckets."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kyber_b200 import workload as wl
    from kyber_b200.multi import shard_bounds, msm_bucket_exchange
    from oracle import bls12381 as o
    C_BITS, W, NB, EB = 4, 64, 15, 96
    a = wl.prng_scalars("b2k/gloo-xa", n, o.R)
    s = wl.prng_scalars("b2k/gloo-xs", n, o.R)
    lo, hi = shard_bounds(n, world, rank)
    pts = [o.g1_mul(x) for x in a[lo:hi]]
    seen = {}

    def local_buckets():
        B = [[None] * NB for _ in range(W)]
        for k, p in zip(s[lo:hi], pts):
            for w in range(W):
                d = (k >> (C_BITS * w)) & (NB)
                if d:
                    B[w][d - 1] = o.g1_add(B[w][d - 1], p)
        return torch.frombuffer(bytearray(b"".join(o.g1_to_affine_bytes(x) for row in B for x in row)), dtype=torch.uint8)

    def reduce_windows(recv, parts, w_cnt):
        raw = bytes(recv.tolist())
        seen["parts"], seen["w_cnt"] = parts, w_cnt
        out = []
        for w in range(w_cnt):
            run = acc = None
            for k in range(NB - 1, -1, -1):
                for p in range(parts):
                    off = ((p * w_cnt + w) * NB + k) * EB
                    run = o.g1_add(run, o.g1_from_affine_bytes(raw[off:off + EB]))
                acc = o.g1_add(acc, run)
            out.append(o.g1_to_affine_bytes(acc))
        return torch.frombuffer(bytearray(b"".join(out)), dtype=torch.uint8)

    def finish(wsums):
        raw = bytes(wsums.tolist())
        acc = None
        for w in range(W - 1, -1, -1):
            for _ in range(C_BITS):
                acc = o.g1_add(acc, acc)
            acc = o.g1_add(acc, o.g1_from_affine_bytes(raw[EB * w:EB * w + EB]))
        return torch.frombuffer(bytearray(o.g1_compress(acc)), dtype=torch.uint8)

    got = bytes(msm_bucket_exchange(local_buckets, reduce_windows, finish, W, NB, EB).tolist())
    want = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
    bad_w = False
    try:
        msm_bucket_exchange(local_buckets, reduce_windows, finish, W + 1, NB, EB)
    except ValueError:
        bad_w = True
    q.put((rank, got == want and seen == {"parts": world, "w_cnt": W // world} and bad_w))
    dist.barrier()
    dist.destroy_process_group()


def test_bucket_exchange_two_ranks_gloo():
    world, n = 2, 23
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_exchange, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_bucket_exchange_without_process_group():
    """world 1 (no torch.distributed): local buckets -> reduce -> finish, no collective; window counts that do not split are refused"""
    from kyber_b200.multi import msm_bucket_exchange
    calls = []

    def local_buckets():
        calls.append("buckets")
        return torch.arange(6 * 4 * 2, dtype=torch.uint8)

    def reduce_windows(recv, parts, w_cnt):
        calls.append(("reduce", parts, w_cnt, recv.numel()))
        return recv.view(6, 4, 2)[:, 0, :].reshape(-1).clone()

    def finish(ws):
        calls.append(("finish", ws.numel()))
        return ws

    out = msm_bucket_exchange(local_buckets, reduce_windows, finish, 6, 4, 2)
    assert calls == ["buckets", ("reduce", 1, 6, 48), ("finish", 12)] and out.numel() == 12
