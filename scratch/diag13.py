import numpy as np, sys
sys.path.insert(0, "/root/repo")
import optimet_b200 as ob
from oracle import oracle as O
from tests import util as U
spec = U.Spec("pair13", [[0, 0, 0], [310.0, 40.0, -120.0], [-90.0, 350.0, 200.0]], [120, 100, 90], U.SI, 13, 900.0)
orc = U.oracle_case(spec)
ctx = ob.Context(0)
U.configure_ctx(ctx, spec, orc)
orc.solve(O.SOLVER_DIRECT)
cs = orc.cross_sections()
for name, opts in (("direct", ob.GmresOpts(ob.OB_SOLVE_DIRECT, 0, 0, 0, 0)), ("belos", ob.GmresOpts(ob.OB_GMRES_BELOS, 1e-13, 400, 80, 5)),
                   ("zcomp", ob.GmresOpts(ob.OB_GMRES_ZCOMP, 1e-14, 300, 0, 3))):
    res = ctx.run(opts)
    print(name, res["iters_ff"], res["iters_sh"], {k: res[k] / cs[o] - 1 for k, o in (("ext", "ext"), ("sca", "sca"), ("sca_SH", "sca_SH"), ("abs_SH", "abs_SH"))})
    print("   X_sca", U.relerr(res["X_sca"], orc.vector(0)), "X_int", U.relerr(res["X_int"], orc.vector(1)), "X_sca_SH", U.relerr(res["X_sca_SH"], orc.vector(2)))
    K, K1 = orc.vector(5), orc.vector(6)
# SH source from oracle's X_int
Kg, K1g = ctx.source_sh(np.conj(orc.vector(1)))
print("K", U.relerr(Kg, orc.vector(5)), "K1ana", U.relerr(K1g, orc.vector(6)))
ctx.set_option("operator", 0)
for h in (1, 2):
    ctx.assemble(h)
    S = orc.matrix(h)
    Sg = ctx.fetch_matrix(h)
    print("matrix", h, U.relerr(Sg, S), np.abs(Sg - S).max() / np.abs(S).max())
    b = 390
    for i in range(3):
        for j in range(3):
            if i != j:
                print("   block", i, j, U.relerr(Sg[i*b:(i+1)*b, j*b:(j+1)*b], S[i*b:(i+1)*b, j*b:(j+1)*b]))
# oracle solve of the SH system with the GPU matrix
