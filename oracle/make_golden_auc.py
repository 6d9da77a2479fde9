"""Golden vectors for the AUC of row N4: runs the reference's own, unmodified auc.calc_auc (imported from
/root/reference; numpy + sklearn.metrics.auc) on seeded error arrays -> tests/golden/auc_cases.npz."""
import os
import sys

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "auc_cases.npz")


def main():
    sys.path.insert(0, "/root/reference")
    import auc as ref
    rs = np.random.RandomState(9)
    g = {}
    sizes = [102, 103, 2018, 5, 2, 40, 17, 64]
    for i, n in enumerate(sizes):
        e = np.abs(rs.standard_normal(n)) * [0.05, 0.1, 0.08, 0.3, 0.01, 0.5, 0.02, 0.2][i]
        cutoff = [0.25, 0.25, 0.25, 0.25, 0.25, 0.1, 0.25, 0.5][i]
        a, pts = ref.calc_auc(e.reshape(-1, 1), cutoff=cutoff)
        g["err_%d" % i] = e; g["cutoff_%d" % i] = np.array(cutoff); g["auc_%d" % i] = np.array(a); g["pts_%d" % i] = pts
    g["n_cases"] = np.array(len(sizes))
    np.savez_compressed(OUT, **g)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
