"""Run under torchrun: the all-reduced multi-GPU J/K must equal the single-GPU J/K.
   python -m torch.distributed.run --nproc-per-node N tools/check_multi_gpu.py [n_waters basis]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import rchem_b200 as rc
from rchem_b200 import geometry as geo, parallel

rank, world, local = parallel.init_distributed("nccl")
torch.cuda.set_device(local)
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 12
bas = sys.argv[2] if len(sys.argv) > 2 else "6-31G*"
z, x = geo.water_cluster(nw)
b = rc.Basis.new(z, x, bas)
b.set_device(local)
b.set_schwarz_tau(1e-10)
n = b.nbf
D = geo.synthetic_density(n)
dev = torch.device("cuda", local)
Dd = torch.from_numpy(D).to(dev)
b.set_stream(torch.cuda.current_stream().cuda_stream)
full = torch.zeros((2, n, n), dtype=torch.float64, device=dev)
b.jk_direct_device(Dd.data_ptr(), full.data_ptr(), 0, 1)
part = torch.zeros((2, n, n), dtype=torch.float64, device=dev)
parallel.jk_direct_distributed(b, Dd, part, rank, world)
torch.cuda.synchronize()
err = (part - full).abs().max().item()
print(f"rank {rank}/{world}: N={n} max|JK_distributed - JK_single| = {err:.3e}", flush=True)
assert err < 1e-12
b.use_own_stream()
torch.distributed.barrier()
torch.distributed.destroy_process_group()
