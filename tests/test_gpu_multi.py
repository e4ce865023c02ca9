"""GPU tests of the in-library sharded MSM (include/b2kyber.h: b2k_comm_*, b2k_bls12381_g1_msm_sharded_*,
b2k_bls12381_g1_msm_multi_gpu; reference loops share/poly.go:461-473).

On ONE device the ranks are several contexts of the same GPU wired with b2k_comm_connect_local: the whole protocol (slabs,
release/acquire flag words, the peer-pulling fused reduction, window-sum push, Horner) runs exactly as across GPUs, only the
"peer" pointers are local.  With >= 2 GPUs the same checks run one rank per PROCESS under torch.distributed.run (CUDA IPC
mapping of the slabs, and the NCCL transport), which is what bench.py --gpus N does."""
import os
import subprocess
import sys

import pytest

from kyber_b200 import Comm, Engine, workload as wl
from oracle import bls12381 as o

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inputs(engine, n, tag):
    a = wl.prng_scalars(tag + "-a", n, o.R)
    s = wl.prng_scalars(tag, n, o.R)
    pts = engine.bls12381_g1_mul_batch_affine(wl.scalars_to_bytes(a), wl.G1_BLS12381_AFFINE * n)
    return a, s, wl.scalars_to_bytes(s), pts


@pytest.mark.parametrize("n,world,c_force", [(1000, 1, 0), (4099, 2, 8), (1 << 16, 4, 11), (1 << 17, 8, 16), (1 << 18, 2, 0)])
def test_multi_gpu_entry_on_one_device(engine, n, world, c_force):
    """b2k_bls12381_g1_msm_multi_gpu with `world` contexts: result == single-call MSM == oracle; repeated (flag words advance)."""
    a, s, sb, pts = _inputs(engine, n, "b2k/mgpu")
    want = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
    engs = [Engine(0) for _ in range(world)]
    comms = [Comm(e, world, r) for r, e in enumerate(engs)]
    try:
        for e in engs:
            e.set_msm_window(c_force)
        Comm.connect_local(comms)
        for _ in range(3):
            assert Comm.msm_multi_gpu(comms, sb, pts) == want
        plan = comms[0].last_plan()
        assert plan["W"] % world == 0 and plan["bucket_bytes"] == 192
        assert engine.bls12381_g1_msm(sb, pts) == want
    finally:
        for c in comms:
            c.close()
        for e in engs:
            e.close()


@pytest.mark.parametrize("shape", [0, 1])
def test_sharded_dev_both_shapes_on_one_device(engine, shape):
    """device-resident shards, bucket exchange (0) and result exchange (1), 2 ranks x 2 steps in flight on one GPU"""
    import torch
    n, world = 6000, 2
    a, s, sb, pts = _inputs(engine, n, "b2k/mgpu-dev")
    want = o.g1_compress(o.g1_mul(wl.dot_mod(s, a, o.R)))
    dev = torch.device("cuda", 0)
    d_s = torch.frombuffer(bytearray(sb), dtype=torch.uint8).to(dev)
    d_p = torch.frombuffer(bytearray(pts), dtype=torch.uint8).to(dev)
    outs = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    torch.cuda.synchronize()
    engs = [Engine(0) for _ in range(world)]
    comms = [Comm(e, world, r) for r, e in enumerate(engs)]
    try:
        for e in engs:
            e.set_msm_window(8)                      # (127 + 8) / 8 = 16 windows: a multiple of the world size
        Comm.connect_local(comms)
        half = n // 2
        for _ in range(2):
            for r, c in enumerate(comms):
                lo, cnt = (0, half) if r == 0 else (half, n - half)
                c.msm_sharded_dev(cnt, d_s[32 * lo:].data_ptr(), d_p[96 * lo:].data_ptr(), outs[r].data_ptr(), shape)
        for e in engs:
            e.wait()
        for r in range(world):
            assert bytes(outs[r][:48].cpu().tolist()) == want
    finally:
        for c in comms:
            c.close()
        for e in engs:
            e.close()


def test_missing_peer_times_out_instead_of_hanging(engine):
    """a rank whose peer never issues the step: the waiting kernel gives up (B2K_ERR_COMM), the device stays usable"""
    if not os.environ.get("B2K_TEST_COMM_TIMEOUT"):
        pytest.skip("takes the 10 s device-side timeout; set B2K_TEST_COMM_TIMEOUT=1")
    from kyber_b200 import B2KError
    a, s, sb, pts = _inputs(engine, 512, "b2k/mgpu-to")
    engs = [Engine(0) for _ in range(2)]
    comms = [Comm(e, 2, r) for r, e in enumerate(engs)]
    try:
        for e in engs:
            e.set_msm_window(8)
        Comm.connect_local(comms)
        import torch
        h_s = torch.frombuffer(bytearray(sb), dtype=torch.uint8).pin_memory()
        h_p = torch.frombuffer(bytearray(pts), dtype=torch.uint8).pin_memory()
        h_o = torch.zeros(64, dtype=torch.uint8).pin_memory()
        comms[0].msm_sharded_async(512, h_s.data_ptr(), h_p.data_ptr(), h_o.data_ptr())
        with pytest.raises(B2KError) as ei:
            engs[0].wait()
        assert ei.value.code == -6
        assert engine.bls12381_g1_msm(sb, pts)        # the device is still usable
    finally:
        for c in comms:
            c.close()
        for e in engs:
            e.close()


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_sharded_msm_across_processes(transport):
    """>= 2 GPUs: one rank per process (torch.distributed.run), CUDA-IPC-mapped slabs or NCCL; see tests/multi_rank_worker.py"""
    if _ngpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if _ngpus() < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "multi_rank_worker.py"), transport]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("RANK_OK") == world, r.stdout[-3000:] + r.stderr[-3000:]
