"""Golden batches for tests/test_dataset_cpu.py, produced by the reference's OWN class (utils/DataSet.py, imported from
/root/reference unchanged) over seeded in-memory models:  python tests/golden/make_dataset_golden.py
Stores, per case, every batch of the iteration (points, batch ids, features, labels, categories, paths) and the next
draw of the random stream after the last batch (the stream position must agree too)."""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference/utils")

import numpy as np  # noqa: E402

from common import DATASET_CASES, make_dataset_models  # noqa: E402

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    import DataSet as RefDataSet  # noqa: E402  (the reference module)


class RefInMemory(RefDataSet.DataSet):
    def __init__(self, models, **kw):
        RefDataSet.DataSet.__init__(self, **kw)
        self.models_ = models
        self.fileList_ = [str(i) for i in range(len(models))]
        self.numPts_ = [0 for _ in models]
        self.categories_ = [m[4] for m in models]

    def _load_model_from_disk_(self, modelPath):
        p, n, f, l, _ = self.models_[int(modelPath)]
        return p.copy(), n.copy(), f.copy(), l.copy()


def main():
    models = make_dataset_models()
    for name, case in DATASET_CASES.items():
        ds = RefInMemory(models, **case["ctor"])
        out, nb = {}, 0
        for _ in range(case["epochs"]):
            ds.start_iteration()
            while ds.has_more_batches():
                n, pts, ids, feats, labels, cat, paths = ds.get_next_batch()
                out["b%d_n" % nb] = np.asarray(n)
                out["b%d_pts" % nb] = np.asarray(pts, np.float64)
                out["b%d_ids" % nb] = np.asarray(ids, np.int32)
                out["b%d_feats" % nb] = np.asarray(feats, np.float64)
                if labels is not None:
                    out["b%d_labels" % nb] = np.asarray(labels)
                if cat is not None:
                    out["b%d_cat" % nb] = np.asarray(cat, np.int32)
                out["b%d_paths" % nb] = np.asarray(paths)
                nb += 1
        out["num_batches"] = np.asarray(nb)
        out["next_draw"] = np.asarray(ds.randomState_.random_sample())
        np.savez_compressed(os.path.join(HERE, "dataset_%s.npz" % name), **out)
        print(name, nb, "batches,", sum(int(out["b%d_pts" % b].shape[0]) for b in range(nb)), "points")


if __name__ == "__main__":
    main()
