"""The loss oracle (oracle/losses.py) against golden values produced by the reference's own
det3d/core/utils/loss_utils.py and head objects (oracle/make_golden.py::gen_losses -> tests/golden/ref_losses.pt)."""
import os

import pytest
import torch

from oracle import losses as ol

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_losses.pt")


@pytest.fixture(scope="module")
def fx():
    return torch.load(GOLD, weights_only=False)


def test_lovasz_and_ce_flat_cases(fx):
    for c in fx["cases"]:
        probas = torch.softmax(c["logits"], -1)
        torch.testing.assert_close(ol.lovasz_softmax(probas, c["labels"], ignore=0), c["lovasz_ignore0"], rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(ol.lovasz_softmax(probas, c["labels"]), c["lovasz_noignore"], rtol=1e-6, atol=1e-7)
        ce = torch.nn.functional.cross_entropy(c["logits"], c["labels"], ignore_index=0)
        torch.testing.assert_close(ce, c["ce_ignore0"], rtol=1e-6, atol=1e-7)


def test_lovasz_dense_and_per_image(fx):
    d = fx["dense"]
    p = torch.softmax(d["logits"], 1)
    torch.testing.assert_close(ol.lovasz_softmax(p, d["labels"]), d["lovasz"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(ol.lovasz_softmax(p, d["labels"], per_image=True), d["lovasz_per_image"], rtol=1e-6, atol=1e-7)


def test_point_head_loss_assembly(fx):
    h = fx["point_head"]
    loss, parts = ol.point_head_loss(h["voxel_logits"], h["voxel_labels"], h["out_logits"], h["point_labels"], h["pcamera"],
                                     h["camera"], ignored_label=0)
    torch.testing.assert_close(loss, h["loss"], rtol=1e-6, atol=1e-6)
    assert set(parts) == set(h["parts"])
    for k, v in h["parts"].items():
        torch.testing.assert_close(parts[k], v, rtol=1e-6, atol=1e-7)


def test_image_head_loss(fx):
    h = fx["image_head"]
    loss, parts = ol.image_head_loss(h["image_logits"], h["image_sem_labels"], loss_weight=0.5, lovasz_loss_weight=-1.0,
                                     ignore_index=0, align_corners=False)
    torch.testing.assert_close(loss, h["loss"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(parts["image_ce_loss"], h["parts"]["image_ce_loss"], rtol=1e-6, atol=1e-7)


def test_lovasz_edge_cases():
    # every label ignored -> zero loss with a graph (loss_utils.py:244-246); a single valid row -> 1-D probas -> zero (:250-251)
    p = torch.softmax(torch.randn(5, 3), -1).requires_grad_()
    z = ol.lovasz_softmax(p, torch.zeros(5, dtype=torch.long), ignore=0)
    assert z.numel() == 0 or float(z.detach().sum()) == 0.0
    one = ol.lovasz_softmax(p, torch.tensor([0, 0, 2, 0, 0]), ignore=0)
    assert float(one.detach().sum()) == 0.0
    # gradient flows and is finite for a regular case
    l = ol.lovasz_softmax(p, torch.tensor([1, 2, 2, 0, 1]), ignore=0)
    l.backward()
    assert torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0
