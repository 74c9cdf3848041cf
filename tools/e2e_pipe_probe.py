"""Two e2e steps in flight (bench.e2e_pipelined) at a small size, errors printed: `python tools/e2e_pipe_probe.py [n_ids]`."""
import sys, time, types
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from vector_db_id_compression_b200 import workloads as W
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
dev = torch.device("cuda:0")
sizes = W.zipf_sizes(n, 65536, 1.0)
off, ids = W.random_partition_lists(n, sizes, 1234, dev)
hin, hout = bench.e2e_buffers(ids)
args = types.SimpleNamespace(max_unit=65536)
try:
    r = bench.e2e_pipelined(args, off, hin.numpy(), hout.numpy(), 3, dev, lambda: None, 1)
    print("pipelined", r)
except Exception as ex:
    print("ERR", repr(ex))
torch.cuda.synchronize()
print("last error check ok; roundtrip", bool(torch.equal(torch.sort(hout.to(dev))[0], torch.sort(ids)[0])))
