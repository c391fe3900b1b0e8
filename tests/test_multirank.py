"""N > 1 host logic on CPU: two gloo ranks take the sample shares the library's own split arithmetic gives them
(csrc/job_split.h, what pbrgpu_render does on a context joined to a job by pbrgpu_nccl_init), render them with the g++
emulation of the device code and add the float4 sums with ONE reduce; count is derived from the alpha sum afterwards,
exactly as the library does over NCCL.  The result must equal the single-process frame: per-path streams are keyed by
(global sample, pixel), so the split does not change any path.  Also checked: the shares of every (ranks x devices)
layout partition the samples, and the NCCL entry points of the library fail cleanly without a device."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {here!r})
import numpy as np, torch, torch.distributed as dist
import pbrlab_b200 as pb, emulbind
from pbrlab_b200 import scenes
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank(); world = dist.get_world_size()
S = pb.Scene([scenes.cornell()], commit_to_device=False)
E = emulbind.Emul(S.flat())
w, h, spp = 24, 24, 7
off, stride, mine = E.job_share(rank, world, spp)
rgba, count, _ = E.render(w, h, spp, seed=77, sample_offset=off, sample_stride=stride)
assert np.all(count == mine) and mine == (4 if rank == 0 else 3)
t_rgba = torch.from_numpy(rgba.copy())
dist.reduce(t_rgba, 0, op=dist.ReduceOp.SUM)            # the one reduce: 16 B per pixel
if rank == 0:
    full, fcount, _ = E.render(w, h, spp, seed=77)
    a, b = t_rgba.numpy(), full
    derived = a[..., 3].astype(np.uint32)                # FinishFrameKernel: count = alpha sum
    assert np.array_equal(derived, fcount), "sample counts differ"
    assert np.all(fcount == spp)
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6), float(np.abs(a - b).max())
    print("MULTIRANK_OK")
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_sample_split_sums_to_the_full_frame(built, tmp_path):
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, here=HERE, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "MULTIRANK_OK" in outs[0]


def test_shares_partition_the_samples(cornell_emul):
    """every (ranks x devices-per-rank) layout deals each sample of the job to exactly one worker"""
    E = cornell_emul
    for spp, job_off, job_stride in [(1024, 0, 1), (7, 0, 1), (64, 3, 5), (5, 0, 1)]:
        want = sorted(range(job_off, spp, job_stride))
        for world, ndev in [(1, 1), (2, 1), (8, 1), (1, 2), (2, 4), (3, 3)]:
            got = []
            for r in range(world):
                for d in range(ndev):
                    off, stride, cnt = E.job_share(r, world, spp, job_off, job_stride, d, ndev)
                    mine = list(range(off, spp, stride))
                    assert len(mine) == cnt
                    got += mine
            assert sorted(got) == want, (spp, job_off, job_stride, world, ndev)


def test_sample_order_starts_every_sample_once(cornell_emul):
    """longest paths first: whatever the pixel permutation, block size and number of probing passes, the
    ids 0 .. npix * spp map onto every (pixel, sample) exactly once — probing samples first, in raster order — and a
    block's samples come sample-major, so consecutive ids are different pixels"""
    E = cornell_emul
    rng = np.random.default_rng(3)
    for npix, probe, rest, block in [(1000, 4, 12, 256), (1000, 1, 1, 1000), (777, 2, 5, 1024), (64, 4, 0, 16),
                                     (4096, 3, 29, 4096), (10, 1, 7, 3)]:
        order = rng.permutation(npix).astype(np.uint32)
        pix, smp = E.sample_order(order, npix, probe, rest, block)
        spp = probe + rest
        key = pix.astype(np.int64) * spp + smp
        assert len(np.unique(key)) == npix * spp and key.min() == 0 and key.max() == npix * spp - 1
        n_probe = npix * probe
        assert np.array_equal(pix[:n_probe], np.tile(np.arange(npix, dtype=np.uint32), probe))
        assert np.array_equal(smp[:n_probe], np.repeat(np.arange(probe, dtype=np.uint32), npix))
        if rest:
            first_block = order[:min(block, npix)]
            got = pix[n_probe:n_probe + len(first_block) * rest].reshape(rest, len(first_block))
            assert np.all(got == first_block[None, :])            # the first block: all its samples, sample-major
            assert np.all(smp[n_probe:] >= probe)
    pix, smp = E.sample_order(None, 50, 0, 6, 16)                  # no order: raster throughout
    assert np.array_equal(pix, np.tile(np.arange(50, dtype=np.uint32), 6))


def test_nccl_entry_points_without_a_device(built):
    """the library loads without libnccl linked in; the job entry points validate their arguments and a unique id can
    be made on a host without a GPU (ncclGetUniqueId needs none)"""
    import ctypes as C
    import pbrlab_b200 as pb
    lib = pb.gpu_lib()
    assert lib.pbrgpu_nccl_unique_id(None) != 0
    ident = pb.nccl_unique_id()
    assert len(ident) == 128 and any(ident)
    assert lib.pbrgpu_nccl_init(None, (C.c_uint8 * 128)(), 0, 1) != 0        # no context
    deps = subprocess.run(["ldd", pb.GPU_LIB], capture_output=True, text=True).stdout
    assert "libnccl" not in deps                                             # reached through dlopen only
