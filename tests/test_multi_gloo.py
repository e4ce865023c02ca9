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
