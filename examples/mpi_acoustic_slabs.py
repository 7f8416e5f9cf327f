"""examples/mpi_acoustic_optimized/MPI_forward.jl + MPI_backward.jl of the reference, on libadseis_b200:
   torchrun --nproc-per-node N --master-addr 127.0.0.1 examples/mpi_acoustic_slabs.py [n]
The reference splits the (n*M) x (n*N) grid into M x N blocks, one MPI rank each; here every rank (one per GPU) owns a
slab of rows and the halo rows travel over NVLink inside the step kernels.  Same inputs as the reference's script
(MPI convention: c given as c^2 on the unpadded grid, unpadded 1-based indices): Ricker(100, 500) at (NX/5, NY/2),
receivers on the row NX/5, c^2 = 1000 with a 2000 inclusion; prints the loss and gradient norms."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adseis_b200 as A  # noqa: E402
from adseis_b200 import parallel  # noqa: E402

rank, world, local_rank = parallel.init_process_group("nccl")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
w = A.workloads.c4(nstep=2000, nx=2 * n, ny=2 * n)          # the reference's example: n = 100 per block, 2000 steps
p, shot = w["param"], w["shots"][0]
dd = parallel.DomainDecomposedAcoustic(p, shot["srci"], shot["srcj"], shot["rcvi"], shot["rcvj"], ctx=A.Context(local_rank))
dd.set_model(w["model_obs"]); dd.set_srcv(shot["srcv"])
dd.forward()                                                # MPI_forward.jl: the observed data
obs = dd.rcvv()
dd.set_model(w["model"]); dd.set_obs(obs)
dd.gradient()                                               # MPI_backward.jl: loss and d loss / d c^2
loss, g = dd.loss(), dd.grad_c().cpu().numpy()
if rank == 0:
    print("%d x %d cells on %d GPU(s): loss %.10e  |grad| max %.3e  sum %.10e" %
          (p.NX, p.NY, world, loss, np.abs(g).max(), g.sum()))
dd.close()
