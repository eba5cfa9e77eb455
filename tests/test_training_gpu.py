"""End-to-end learning through the reference-shaped API: MCClassS in training mode (batch statistics, dropout off),
loss.backward() through the dense glue (torch ops) AND the Monte-Carlo ops (own backward kernels), a few Adam steps on
a toy two-class problem (spheres vs cubes) -- the loss must fall and every variable must receive a gradient; the
trained network is then evaluated through the inference path (folded batch norms, own dense kernels)."""
import numpy as np
import pytest

from common import make_cloud

pytestmark = pytest.mark.gpu


def _toy_batch(B, n, seed, dev):
    import torch
    rng = np.random.default_rng(seed)
    pts, bids, labels = [], [], []
    for b in range(B):
        kind = int(rng.integers(0, 2))
        pts.append(make_cloud(n, "sphere" if kind == 0 else "cube", rng))
        bids.append(np.full((n, 1), b, np.int32))
        labels.append(kind)
    pts, bids = np.concatenate(pts), np.concatenate(bids)
    return (torch.from_numpy(pts).to(dev), torch.from_numpy(bids).to(dev),
            torch.ones((pts.shape[0], 1), device=dev), torch.tensor(labels, device=dev))


def test_mcclasss_learns_a_toy_problem():
    import torch
    from mccnn_b200 import variables
    from mccnn_b200.models import MCClassS
    dev = torch.device("cuda:0")
    B, n, k = 8, 256, 8
    store = variables.VariableStore(device=dev, seed=1)

    def net(batch, training):
        p, b, f, _ = batch
        return MCClassS.create_network(p, b, f, 1, B, k, 2, training, 1.0, 1.0, store=store)[:B]

    net(_toy_batch(B, n, 0, dev), True)                      # creates the variables
    trainable = [v for name, v in store.vars.items() if "moving_" not in name]
    for v in trainable:
        v.requires_grad_(True)
    opt = torch.optim.Adam(trainable, lr=5e-3)
    losses = []
    for step in range(30):
        batch = _toy_batch(B, n, 100 + step, dev)
        loss = torch.nn.functional.cross_entropy(net(batch, True), batch[3])
        opt.zero_grad()
        loss.backward()
        if step == 0:
            missing = [name for name, v in store.vars.items() if "moving_" not in name and v.grad is None]
            assert not missing, missing
            assert all(bool(torch.isfinite(v.grad).all()) for v in trainable)
        opt.step()
        losses.append(float(loss.detach()))
    print("loss first5 %.4f last5 %.4f" % (np.mean(losses[:5]), np.mean(losses[-5:])))
    assert np.mean(losses[-5:]) < 0.6 * np.mean(losses[:5]), losses
    # held-out clouds, still normalised with batch statistics (30 steps at momentum 0.99 move the running averages
    # only a quarter of the way, so the inference path is checked for consistency below, not for accuracy)
    with torch.no_grad():
        correct = 0
        for s in range(4):
            batch = _toy_batch(B, n, 900 + s, dev)
            correct += int((net(batch, True).argmax(dim=1) == batch[3]).sum())
    print("held-out correct %d / 32" % correct)
    assert correct >= 24, correct                             # >= 75 % of 32 held-out clouds
    # the moving statistics were updated, and the inference path (folded BN, own dense kernels) consumes them
    mm = store.vars["Reduce_1_In_BN_BN/moving_mean"]
    assert float(mm.abs().max()) > 0
    with torch.no_grad():
        out = net(_toy_batch(B, n, 950, dev), False)
    assert out.shape == (B, 2) and bool(torch.isfinite(out).all())
