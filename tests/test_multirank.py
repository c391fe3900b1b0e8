"""N > 1 host logic on CPU: two gloo ranks render interleaved sample shares (rank r of R: samples r, r+R, ...) with the
g++ emulation of the device code and sum the accumulators with one reduce, exactly as bench.py does over NCCL.
The sum must equal the single-process frame: per-path streams are keyed by (global sample, pixel), so the split
does not change any path."""
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
w, h, spp = 24, 24, 6
rgba, count, _ = E.render(w, h, spp, seed=77, sample_offset=rank, sample_stride=world)
t_rgba = torch.from_numpy(rgba.copy()); t_count = torch.from_numpy(count.astype(np.int64))
dist.reduce(t_rgba, 0, op=dist.ReduceOp.SUM); dist.reduce(t_count, 0, op=dist.ReduceOp.SUM)
if rank == 0:
    full, fcount, _ = E.render(w, h, spp, seed=77)
    assert np.array_equal(t_count.numpy(), fcount.astype(np.int64)), "sample counts differ"
    assert np.all(fcount == spp)
    a, b = t_rgba.numpy(), full
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6), float(np.abs(a - b).max())
    # each rank alone holds only its share
    assert count.sum() == w * h * (spp // 2)
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
