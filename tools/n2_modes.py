"""One exchange mode of the 2-rank step (tests/test_gpu_text_shard.py::_nccl_worker) as a bounded stand-alone run:
    python tools/n2_modes.py False False     # (shard the text tower?, peer memory: False = NCCL only / None = try peer)
"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.multiprocessing as mp
    from tests.test_gpu_text_shard import _free_port, _nccl_worker
    mode = (sys.argv[1] == "True", None if sys.argv[2] == "None" else sys.argv[2] == "True")
    t0 = time.time()
    out = tempfile.mkdtemp()
    mp.spawn(_nccl_worker, args=(2, _free_port(), out, mode), nprocs=2, join=True)
    for f in sorted(os.listdir(out)):
        d = torch.load(os.path.join(out, f))
        print(f, d["collectives"], d["loss"].tolist())
    print("mode", mode, "ok in %.1f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
