"""Pins parity against the REAL reference: compares a Torch7 dump ("Pinning parity on a machine that has
Torch7": `th dump_parity.lua` run inside mbchang/dynamics/src/lua) with the CUDA path and with the CPU oracle.

usage: python tests/golden/check_th_dump.py dump.t7        (needs a B200 for the CUDA leg; --oracle-only without one)

Nothing in the test suite runs this (no Torch7 exists in the build image, so no dump is committed): it is the tool a
maintainer with `th` uses to turn "parity unpinned" into a pinned result.  Tolerances: tests/util.py (1e-4 relative on
predictions and losses, 1e-4 of ||g||_inf on the gradient, masks and 12-nearest indices bit-exact).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dynamics_b200 import t7                                   # noqa: E402
from oracle import npe                                          # noqa: E402
from tests.util import grad_err, rel_err                        # noqa: E402


def main(path, oracle_only=False):
    d = t7.load(path)
    this, ctx, y = (np.asarray(d[k], np.float32) for k in ("this", "context", "y"))
    params = t7.params_from_network(d["network"])               # canonical layout from the storage offsets of every nn.Linear
    ref_flat = np.asarray(d["params"], np.float32).ravel()
    print("reference getParameters() order == canonical layout:", bool(np.array_equal(params, ref_flat)))
    pred_ref = np.asarray(d["pred"], np.float32).reshape(len(this), -1)
    masks_ref = np.stack([np.asarray(m, np.float32) for _, m in sorted(d["masks"].items())], axis=1)
    # the reference gradient is in ITS flat order: map it through the same offsets by loading it as a parameter vector
    grad_ref = np.asarray(d["grad"], np.float32).ravel()
    if not np.array_equal(params, ref_flat):
        perm = t7.params_from_network(_with_storage(d["network"], np.arange(ref_flat.size, dtype=np.float32))).astype(np.int64)
        grad_ref = grad_ref[perm]
    report = {}
    batch = (this, ctx, y)
    l, p, lv, lw = npe.fp(params, batch)
    g, _ = npe.bp(params, batch)
    report["oracle"] = dict(pred=rel_err(p[:, :6], pred_ref[:, :6]), loss=rel_err([l, lv, lw], [d["loss"], d["vel_loss"], d["angvel_loss"]], 1e-6),
                            grad=grad_err(g, grad_ref), masks_equal=bool(np.array_equal(npe.neighbor_mask(this, ctx), masks_ref)))
    if not oracle_only:
        from dynamics_b200 import NPEModel
        m = NPEModel()
        gl, gp, glv, glw = m.fp(params, batch)
        gg = m.bp(batch)
        idx, mask, cnt = m.pair_table()
        report["cuda"] = dict(pred=rel_err(gp[:, :6], pred_ref[:, :6]), loss=rel_err([gl, glv, glw], [d["loss"], d["vel_loss"], d["angvel_loss"]], 1e-6),
                              grad=grad_err(gg, grad_ref), masks_equal=bool(np.array_equal(mask[:, :masks_ref.shape[1]], masks_ref.astype(np.uint8))))
        m.close()
    ok = True
    for who, r in report.items():
        print(who, r)
        ok &= r["pred"] < 1e-4 and r["loss"] < 1e-4 and r["grad"] < 1e-4 and r["masks_equal"]
    print("PARITY PINNED" if ok else "PARITY MISMATCH")
    return 0 if ok else 1


def _with_storage(network, values):
    """The same object graph with the shared parameter storage replaced by `values` (here: indices), so that
    params_from_network returns the permutation canonical index -> reference flat index."""
    seen = set()

    def walk(o):
        if id(o) in seen:
            return
        seen.add(id(o))
        if isinstance(o, t7.TensorView) and o.storage is not None and o.storage.size == values.size:
            o[...] = values[o.storage_offset:o.storage_offset + o.size].reshape(o.shape)
        elif isinstance(o, dict):
            for v in o.values():
                walk(v)
    walk(network)
    return network


if __name__ == "__main__":
    sys.exit(main(sys.argv[1], "--oracle-only" in sys.argv))
