"""Worker for tests/test_gpu_multirank.py (launched by torch.distributed.run, one rank per GPU, NCCL):
runs the row-sharded step on a small cluster and has rank 0 save the results."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_path, nobj, nMax = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    mode = sys.argv[4] if len(sys.argv) > 4 else "pairs"
    import torch
    import torch.distributed as dist
    from optimet_b200 import host as H, sharding, xmlgen
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    xyz = xmlgen.random_sites(nobj, 300.0 * nobj ** (1.0 / 3.0) + 200.0, 150.0, 7)
    belos = [("Solver", "string", "GMRES"), ("Convergence Tolerance", "double", "1.0e-12"),
             ("Maximum Iterations", "int", "600"), ("Num Blocks", "int", "60"), ("Maximum Restarts", "int", "20")]
    case = H.Case(xml=xmlgen.cluster_xml(xyz, 50.0, nMax, 800.0, belos=belos))
    solver = H.Solver(case, device=lr)
    sharding.attach(solver, dist, rank, world)
    if mode == "aca":      # compressed operator, block-rows sharded, all-gather of the y slices
        solver.set_aca_mode(1)
    elif mode == "dense":  # the reference's slab, same sharding
        solver.set_option("operator", 0)
    elif mode == "rot":    # rotated-axial form: pair list sharded, all-reduce of the partial sums
        solver.set_option("operator", 3)
    res = solver.step()
    cs = sharding.sum_partials(dist, torch, [res[k] for k in ("ext", "sca", "abs", "sca_SH", "abs_SH")], device="cuda")
    # replicated vectors must be bit-identical on all ranks
    x = torch.from_numpy(res["X_sca_SH"].view(np.float64)).cuda()
    x0 = x.clone()
    dist.broadcast(x0, src=0)
    same = torch.tensor([1.0 if torch.equal(x, x0) else 0.0], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.savez(out_path, cs=np.array(cs), X_sca=res["X_sca"], X_int=res["X_int"], X_sca_SH=res["X_sca_SH"],
                 X_int_SH=res["X_int_SH"], iters=np.array([res["iters_ff"], res["iters_sh"]]), same=same.item())
    solver.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
