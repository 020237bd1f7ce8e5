"""Adds derivative fixtures <name>.deriv.npz to tests/golden/ from the REFERENCE ITSELF.

Run in the build container (needs oracle/_ref/liblocoref.so):    python tests/golden/make_golden_deriv.py

For every committed <name>.pb the graph is re-materialised as reference objects (ref_adapter.cpp, the loader's code
path without protobuf) and the reference's LCSampledFunction::evalDeriv (LCSampledFunction.cpp:61-67) is run on the
fixture's queries along every cube direction.  Stored: deriv [N, Q] (NaN where the reference returned an LCError).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from oracle import refapi  # noqa: E402
import protoschema  # noqa: E402


def main():
    from conftest import GOLDEN_CASES, golden_proto_path
    only = set(sys.argv[1:])  # python make_golden_deriv.py [case ...]
    for name in GOLDEN_CASES:
        if only and name not in only:
            continue
        pb = golden_proto_path(name)
        z = np.load(os.path.join(HERE, name + ".npz"))
        fn = refapi.RefFn.from_graph(protoschema.read_binary(pb))
        deriv = np.stack([fn.eval_deriv(z["q"], d) for d in range(fn.N)])
        np.savez_compressed(os.path.join(HERE, name + ".deriv.npz"), deriv=deriv)
        print(f"{name}: N={fn.N} deriv {deriv.shape}, {int(np.isnan(deriv[0]).sum())} errors, max |d| {np.nanmax(np.abs(deriv)):.3g}")


if __name__ == "__main__":
    main()
