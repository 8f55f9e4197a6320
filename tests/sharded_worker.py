"""TEST INFRASTRUCTURE: one rank of the world_size-2 CPU test of fastf_b200/sharded.py.  Runs under torch.distributed.run with the
gloo backend and the SIMT-emulator build of the C-ABI (FASTF_GPU_LIB), so the whole multi-GPU host logic -- sharding, ordinal
bases, the all-to-all by cell hash, the merge -- is exercised without a GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NCCL = os.environ.get("FASTF_SHARDED_BACKEND") == "nccl"   # the same worker drives the real 2-GPU test (tests/test_gpu_parity.py)
assert NCCL or "libfastf_emu" in os.environ.get("FASTF_GPU_LIB", "")
import torch   # noqa: E402
import torch.distributed as dist   # noqa: E402
from fastf_b200 import sharded   # noqa: E402

d, out, rc_, rd_, seed = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
if NCCL:
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
else:
    dist.init_process_group("gloo")
os.chdir(d)
rc = sharded.bam2db_sharded("in.bam", os.path.join(out, "x.db"), out, "barcodes.tsv.gz", "features.tsv.gz", rc_, rd_, seed, dist=dist, torch=torch, device="cuda" if NCCL else "cpu")
dist.barrier()
dist.destroy_process_group()
sys.exit(rc)
