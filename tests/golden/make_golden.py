"""Generates the golden fixtures in tests/golden/ from the REFERENCE ITSELF.

Run in the build container (needs /root/reference compiled into oracle/_ref/liblocoref.so by
`make -C oracle ref`):      python tests/golden/make_golden.py

For every case the reference builds the grid with its own builder (LCAdaptiveGrid), the object graph is
dumped and written as the proto file the reference would write (through google.protobuf, independent of
the product's codec), and the reference's LCSampledFunction::evalShapeInfo is run on seeded queries.
Stored per case:  <name>.pb  and  <name>.npz  {q, out, leaf, status, sup_off, sup_ids, n_neighbors, center_err}.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import oracleapi, refapi  # noqa: E402
import protoschema  # noqa: E402

# name -> (N, basis, linear_fn, max_depth, threshold, bootstrap, eval_corners, random_refinements(seed,n,maxLevel) or None)
CASES = {
    # BASELINE.json configs[0]: TestExamples 2-D tree (TestExamples.cpp:457-469) with the linear basis
    "c1_lin2d_boot2": (2, 0, False, 3, 0.1, 2, True, None),
    # TestExamples as shipped (cubic), both parameter sets
    "cub2d_boot2": (2, 1, False, 3, 0.1, 2, True, None),
    "cub2d_boot1_adapt": (2, 1, False, 3, 0.5, 1, True, None),
    # protoTest topology: boot 0 + srand(1) + 12 random refinements (TestExamples.cpp:214-240), linear f
    "cub2d_prototest": (2, 1, True, 0, 0.1, 0, True, (1, 12, 5)),
    # linearPrecisionTestLinear (N=1, cubic; TestExamples.cpp:381-448)
    "cub1d_linprec": (1, 1, True, 0, 0.1, 0, True, (1, 12, 5)),
    "lin2d_prototest": (2, 0, True, 0, 0.1, 0, True, (1, 12, 5)),
    "lin3d_adapt5": (3, 0, False, 5, 0.1, 1, True, None),
    "lin2d_adapt6": (2, 0, False, 6, 0.1, 2, True, None),
    "cub2d_adapt6": (2, 1, False, 6, 0.1, 2, True, None),
    "cub3d_boot2": (3, 1, False, 0, 0.1, 2, True, None),
    "cub3d_boot1_nocorner": (3, 1, False, 0, 0.1, 1, False, None),
    "lin4d_boot1": (4, 0, False, 0, 0.1, 1, True, None),
    # the TestDebug tree (TestDebug.cpp:153-157): 2-D cubic, maxDepth 10, threshold 0.1, bootstrap 2 -> 724 leaves,
    # 54,419 basis functions; ~30 s to build, 2.8 MB of proto -> stored as .pb.gz
    "cub2d_testdebug": (2, 1, False, 10, 0.1, 2, True, None),
}
GZIPPED = {"cub2d_testdebug"}


def queries(N, n, seed):
    rng = np.random.default_rng(seed)
    q = rng.random((n, N))
    q[:64] = np.round(q[:64] * 16) / 16          # exactly on split planes / sample positions (ties -> child2)
    q[64:72] = 1.0 + rng.random((8, N)) * 1e-7    # outside by less than EPSILON: accepted in the boundary leaf
    q[72:80] = -rng.random((8, N)) * 1e-7
    q[80:88] = 1.0 + 1e-5 + rng.random((8, N))    # outside: "parameter ouside cell"
    q[88:96] = -1e-5 - rng.random((8, N))
    q[96] = 0.0
    q[97] = 1.0
    # (a NaN coordinate is not included: the reference build segfaults on it, so there is nothing to pin)
    return q


def main():
    only = set(sys.argv[1:])  # python make_golden.py [case ...]: regenerate only the named cases
    for name, (N, basis, lin, depth, thr, boot, corners, refine) in CASES.items():
        if only and name not in only:
            continue
        g = refapi.RefGrid(N, basis, lin, depth, thr, boot, 0, corners)
        if refine:
            g.refine_random(*refine)
        fn = g.sampled_function()
        ga = oracleapi.graph_from_ref(fn.graph())
        protoschema.write_binary(ga, os.path.join(HERE, name + ".pb"))
        q = queries(N, 1500, 1234)
        out, leaf, status = fn.eval(q)
        sup = [fn.support(l) for l in range(fn.L)]
        sup_off = np.cumsum([0] + [len(s) for s in sup]).astype(np.int64)
        nn = np.array([fn.num_neighbors(l) for l in range(fn.L)], dtype=np.int32)
        cerr = np.array([fn.center_error(l) for l in range(fn.L)])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), q=q, out=out, leaf=leaf, status=status, sup_off=sup_off,
                            sup_ids=np.concatenate(sup).astype(np.int32), n_neighbors=nn, center_err=cerr,
                            true_values=g.true_values(q))
        if name in GZIPPED:
            import gzip
            pb = os.path.join(HERE, name + ".pb")
            with open(pb, "rb") as f, gzip.GzipFile(pb + ".gz", "wb", compresslevel=9, mtime=0) as o:
                o.write(f.read())
        print(f"{name}: N={N} basis={basis} leaves={fn.L} samples={fn.S} basis_fns={fn.B} border={fn.Bc} "
              f"bad={int((status != 0).sum())} pb={os.path.getsize(os.path.join(HERE, name + '.pb'))}B")
        if name in GZIPPED:
            os.unlink(os.path.join(HERE, name + ".pb"))


if __name__ == "__main__":
    main()
