"""One rank of tests/test_gpu_multi.py::test_sharded_msm_across_processes (launched by torch.distributed.run).

Every rank owns a shard of one MSM; the sum over all ranks must equal the oracle's ((sum s_i a_i) mod r) G on every rank, for the
bucket exchange and the result exchange, device-resident and from host buffers, several steps in flight."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from kyber_b200 import Comm, Engine, workload as wl
    from oracle import bls12381 as o

    transport = sys.argv[1] if len(sys.argv) > 1 else "peer"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("gloo")                     # out-of-band plumbing only (blobs, ids, the check)
    n = 20000 + 7 * rank                                # ragged shards
    a = wl.prng_scalars("b2k/mrank-a", n, o.R, rank * 100000)
    s = wl.prng_scalars("b2k/mrank", n, o.R, rank * 100000)
    NC = 2
    engs = [Engine(local) for _ in range(NC)]
    for e in engs:
        e.set_msm_window(16)                            # one plan on every rank: W = 8 windows
    pts = engs[0].bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
    sb = wl.scalars_to_bytes(s)
    comms = [Comm(e, world, rank) for e in engs]
    for k, c in enumerate(comms):
        if transport == "nccl":
            ids = [Comm.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            c.use_nccl(ids[0])
        else:
            blobs = [None] * world
            dist.all_gather_object(blobs, c.export())
            c.connect(b"".join(blobs))
    dots = [None] * world
    dist.all_gather_object(dots, wl.dot_mod(s, a, o.R))
    want = o.g1_compress(o.g1_mul(sum(dots) % o.R))
    d_s = torch.frombuffer(bytearray(sb), dtype=torch.uint8).to(dev)
    d_p = torch.frombuffer(bytearray(pts), dtype=torch.uint8).to(dev)
    outs = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(NC)]
    torch.cuda.synchronize()
    shapes = (0,) if transport == "nccl" else (0, 1)
    for shape in shapes:
        for it in range(3):
            for k, c in enumerate(comms):
                c.msm_sharded_dev(n, d_s.data_ptr(), d_p.data_ptr(), outs[k].data_ptr(), shape)
        for e in engs:
            e.wait()
        for k in range(NC):
            assert bytes(outs[k][:48].cpu().tolist()) == want, f"rank {rank} shape {shape} context {k}: wrong sum"
    h_s = torch.frombuffer(bytearray(sb), dtype=torch.uint8).pin_memory()
    h_p = torch.frombuffer(bytearray(pts), dtype=torch.uint8).pin_memory()
    h_o = [torch.zeros(64, dtype=torch.uint8).pin_memory() for _ in range(NC)]
    for it in range(2):
        for k, c in enumerate(comms):
            c.msm_sharded_async(n, h_s.data_ptr(), h_p.data_ptr(), h_o[k].data_ptr())
        for e in engs:
            e.wait()
    for k in range(NC):
        assert bytes(h_o[k][:48].tolist()) == want, f"rank {rank}: host-buffer sharded MSM wrong"
    dist.barrier()
    for c in comms:
        c.close()
    for e in engs:
        e.close()
    print("RANK_OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
