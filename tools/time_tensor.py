"""Timing of the dense-tensor path and JK_inmem: python tools/time_tensor.py nw basis"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import rchem_b200 as rc
from rchem_b200 import geometry as geo
nw, bas = int(sys.argv[1]), sys.argv[2]
z, x = geo.water_cluster(nw)
b = rc.Basis.new(z, x, bas)
n = b.nbf
dev = torch.device("cuda", 0)
I = torch.empty((n,) * 4, dtype=torch.float64, device=dev)
b.set_stream(torch.cuda.current_stream().cuda_stream)
for boys in (0, 1):
    b.set_boys(boys)
    b.build_I_device(I.data_ptr()); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        b.build_I_device(I.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    s = b.stats()
    print(f"(H2O){nw} {bas} N={n} boys={boys} build_I {ms:.3f} ms (kernels {s['kernel_ms']:.3f} ms) tensor {8*n**4/1e6:.1f} MB -> {8*n**4/ms/1e6:.1f} GB/s written; quartets {s['shell_quartets']:.3e} model {s['model_flops']/ms/1e9:.2f} TFLOP/s")
D = torch.from_numpy(geo.synthetic_density(n)).to(dev)
JK = torch.empty((2, n, n), dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
rc.jk_inmem_device(n, I.data_ptr(), D.data_ptr(), JK.data_ptr(), st); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    rc.jk_inmem_device(n, I.data_ptr(), D.data_ptr(), JK.data_ptr(), st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"JK_inmem N={n}: {ms:.3f} ms -> {8*n**4/ms/1e6:.1f} GB/s (algorithmic 8 N^4 bytes)")
b.use_own_stream()
