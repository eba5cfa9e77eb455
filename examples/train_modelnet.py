#!/usr/bin/env python
"""Classification training on ModelNet40-style data -- the reference's ModelNet/ModelNet.py:60-300 on this library.

    python examples/train_modelnet.py --folder data [--model MCClassS] [--grow 64] [--batchSize 32] [--maxEpoch 201]
    python examples/train_modelnet.py --synthetic            # a generated stand-in data set (no files needed)

Same command line, data pipeline (ModelNetDataSet with the reference's sampling protocols and augmentation, assembled
on the GPU), networks (mccnn_b200.models.<model>.create_network, training mode), loss (mean sparse soft-max
cross-entropy + l2 weight decay over the 'weight_decay_loss' collection, ModelNet.py:33-40), optimiser (Adam with
stair-cased exponential decay clipped from below, ModelNet.py:50-57) and per-epoch test pass as the reference script;
TensorFlow's session / summary / checkpoint plumbing is replaced by a torch loop and `VariableStore.state_dict()`."""
import argparse
import importlib
import math
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from mccnn_b200 import variables
from mccnn_b200.datasets import ModelNetDataSet


def write_synthetic_modelnet(root, categories=4, models_per_cat=24, points=600, seed=0):
    """A small ModelNet40-shaped folder of analytic shapes (ellipsoids / boxes / cylinders / cones with noise)."""
    rng = np.random.default_rng(seed)
    names = ["ellipsoid", "box", "cylinder", "cone"][:categories]
    with open(os.path.join(root, "modelnet40_shape_names.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    lists = {"train": [], "test": []}
    for c, name in enumerate(names):
        os.makedirs(os.path.join(root, name), exist_ok=True)
        for m in range(models_per_cat):
            u = rng.normal(size=(points, 3))
            u /= np.linalg.norm(u, axis=1, keepdims=True)
            if c == 0:
                p, nrm = u * rng.uniform(0.3, 1.0, (1, 3)), u
            elif c == 1:
                p = u / np.abs(u).max(axis=1, keepdims=True) * rng.uniform(0.4, 1.0, (1, 3))
                nrm = np.sign(u) * (np.abs(u) == np.abs(u).max(axis=1, keepdims=True))
            elif c == 2:
                r = np.linalg.norm(u[:, :2], axis=1, keepdims=True) + 1e-9
                p = np.concatenate([u[:, :2] / r * 0.5, rng.uniform(-1, 1, (points, 1))], axis=1)
                nrm = np.concatenate([u[:, :2] / r, np.zeros((points, 1))], axis=1)
            else:
                h = rng.uniform(0, 1, (points, 1))
                r = np.linalg.norm(u[:, :2], axis=1, keepdims=True) + 1e-9
                p = np.concatenate([u[:, :2] / r * (1 - h) * 0.6, h * 2 - 1], axis=1)
                nrm = np.concatenate([u[:, :2] / r, np.full((points, 1), 0.3)], axis=1)
            p = p + rng.normal(scale=0.01, size=p.shape)
            mname = "%s_%04d" % (name, m + 1)
            np.savetxt(os.path.join(root, name, mname + ".txt"), np.concatenate([p, nrm], axis=1), fmt="%.6f",
                       delimiter=",")
            lists["train" if m < models_per_cat * 3 // 4 else "test"].append(mname)
    for split, lst in lists.items():
        with open(os.path.join(root, "modelnet40_%s.txt" % split), "w") as f:
            f.write("\n".join(lst) + "\n")


def main():
    ap = argparse.ArgumentParser(description="Train MCCNN for classification of point clouds (ModelNet40 layout)")
    ap.add_argument("--folder", default="data")
    ap.add_argument("--synthetic", action="store_true", help="generate a small stand-in data set in a temp folder")
    ap.add_argument("--logFolder", default="log")
    ap.add_argument("--model", default="MCClassS")
    ap.add_argument("--grow", default=64, type=int)
    ap.add_argument("--batchSize", default=32, type=int)
    ap.add_argument("--maxEpoch", default=201, type=int)
    ap.add_argument("--initLearningRate", default=0.005, type=float)
    ap.add_argument("--learningDeacyFactor", default=0.5, type=float)
    ap.add_argument("--learningDecayRate", default=20, type=int, help="epochs")
    ap.add_argument("--maxLearningRate", default=0.00001, type=float, help="lower bound of the decayed rate")
    ap.add_argument("--useDropOut", action="store_true")
    ap.add_argument("--dropOutKeepProb", default=0.5, type=float)
    ap.add_argument("--useDropOutConv", action="store_true")
    ap.add_argument("--dropOutKeepProbConv", default=0.8, type=float)
    ap.add_argument("--weightDecay", default=0.0, type=float)
    ap.add_argument("--nPoints", default=1024, type=int)
    ap.add_argument("--ptDropOut", default=0.95, type=float)
    ap.add_argument("--augment", action="store_true")
    ap.add_argument("--nonunif", action="store_true")
    ap.add_argument("--gpu", default="0")
    ap.add_argument("--seed", default=None, type=int)
    args = ap.parse_args()

    dev = torch.device("cuda", int(args.gpu))
    torch.cuda.set_device(dev)
    os.makedirs(args.logFolder, exist_ok=True)
    folder = args.folder
    if args.synthetic:
        folder = tempfile.mkdtemp(prefix="modelnet_synth_")
        write_synthetic_modelnet(folder, points=max(600, int(args.nPoints * 1.2)))
    model = importlib.import_module("mccnn_b200.models." + args.model)

    maxStored = 5000 if args.nonunif else int(float(args.nPoints) * (2.0 - args.ptDropOut))   # ModelNet.py:131-141
    train = ModelNetDataSet(True, args.nPoints, args.ptDropOut, maxStored, args.batchSize,
                            [1, 2, 3, 4] if args.nonunif else [0], args.augment, folder=folder, seed=args.seed,
                            device=dev)
    test = ModelNetDataSet(False, args.nPoints, 1.0, maxStored, 1, [0, 1, 2, 3, 4] if args.nonunif else [0], False,
                           folder=folder, seed=args.seed, device=dev)
    cats = train.get_categories()
    batchesPerEpoch = int(math.ceil(train.get_num_models() / float(args.batchSize)))
    print("categories %d | train models %d | test models %d" % (len(cats), train.get_num_models(),
                                                                test.get_num_models()))
    store = variables.VariableStore(device=dev, seed=args.seed or 0)

    def net(points, batchIds, features, B, training):
        return model.create_network(points, batchIds, features, 1, B, args.grow, len(cats), training,
                                    args.dropOutKeepProbConv, args.dropOutKeepProb, args.useDropOutConv,
                                    args.useDropOut, store=store)[:B]

    # one forward creates the variables; everything but the moving statistics trains
    train.start_iteration()
    n, p, b, f, _, c, _ = train.get_next_batch()
    net(p, b, f, args.batchSize, True)
    params = [v.requires_grad_(True) for name, v in store.vars.items() if "moving_" not in name]
    decayed = list(store.collections.get("weight_decay_loss", []))
    opt = torch.optim.Adam(params, lr=args.initLearningRate)
    step = 0
    for epoch in range(args.maxEpoch):
        t0 = time.time()
        train.start_iteration()
        lossAcc, seen, correct = 0.0, 0, 0
        while train.has_more_batches():
            n, p, b, f, _, c, _ = train.get_next_batch()
            lr = max(args.initLearningRate * args.learningDeacyFactor **
                     (step // max(1, args.learningDecayRate * batchesPerEpoch)), args.maxLearningRate)
            for g in opt.param_groups:
                g["lr"] = lr
            logits = net(p, b, f, args.batchSize, True)[:n]
            loss = torch.nn.functional.cross_entropy(logits, c.long())
            if args.weightDecay > 0:   # tf.contrib.layers.l2_regularizer(scale): scale * sum(w^2) / 2
                loss = loss + args.weightDecay * sum((w * w).sum() for w in decayed) * 0.5
            opt.zero_grad()
            loss.backward()
            opt.step()
            lossAcc += float(loss.detach())
            correct += int((logits.argmax(dim=1) == c.long()).sum())
            seen += n
            step += 1
        # test pass: inference path (folded batch norms, own dense kernels), one model per batch as the reference
        test.start_iteration()
        tc, tn = 0, 0
        with torch.no_grad():
            while test.has_more_batches():
                n, p, b, f, _, c, _ = test.get_next_batch()
                tc += int((net(p, b, f, 1, False).argmax(dim=1) == c.long()).sum())
                tn += n
        print("epoch %3d | loss %.4f | train acc %.2f %% | test acc %.2f %% | lr %.2e | %.1f s" %
              (epoch, lossAcc / max(1, batchesPerEpoch), 100.0 * correct / max(1, seen), 100.0 * tc / max(1, tn), lr,
               time.time() - t0), flush=True)
        if epoch % 10 == 0:
            torch.save(store.state_dict(), os.path.join(args.logFolder, "model.pt"))


if __name__ == "__main__":
    main()
