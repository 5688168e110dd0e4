"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's training losses (the oracle for
the training-step row of 8f).  Nothing under lidarseg3d_b200/ imports this file.

Follows, function by function:
  lovasz_grad / lovasz_softmax_flat / flatten_probas / lovasz_softmax   det3d/core/utils/loss_utils.py:217-333
  point-head loss assembly (voxel CE + Lovasz, point CE + Lovasz, mimic MSE)
                                                                    det3d/models/point_heads/point_seg_mseg3d_head.py:107-115,137-196
  image-head loss (resize logits to the label size, 0.5 * CE, optional Lovasz)
                                                                    det3d/models/img_heads/fcn_mseg3d_head.py:147-153,202-244
Pinned: tests/golden/ref_losses.pt holds inputs and the values the reference's own functions return for them
(oracle/make_golden.py::gen_losses imports det3d/core/utils/loss_utils.py from /root/reference).
"""
import torch
import torch.nn.functional as F


def lovasz_grad(gt_sorted):
    """loss_utils.py:278-291: gradient of the Lovasz extension w.r.t. the sorted errors (Alg. 1 of the paper)."""
    p = gt_sorted.shape[0]
    gts = gt_sorted.sum()
    intersection = gts - gt_sorted.float().cumsum(0)
    union = gts + (1 - gt_sorted).float().cumsum(0)
    jaccard = 1.0 - intersection / union
    if p > 1:
        jaccard = torch.cat([jaccard[:1], jaccard[1:] - jaccard[:-1]])
    return jaccard


def flatten_probas(probas, labels, ignore=None):
    """loss_utils.py:294-330: [P, C] stays, [B, C, H, W] -> [B*H*W, C]; rows whose label == ignore are dropped."""
    if probas.dim() == 4:
        B, C, H, W = probas.shape
        probas = probas.permute(0, 2, 3, 1).contiguous().view(-1, C)
    elif probas.dim() == 5:
        B, C, L, H, W = probas.shape
        probas = probas.contiguous().view(B, C, L, H * W).permute(0, 2, 3, 1).contiguous().view(-1, C)
    labels = labels.view(-1)
    if ignore is None:
        return probas, labels
    valid = labels != ignore
    # the reference indexes with valid.nonzero().squeeze(): a single valid row collapses to a 1-D tensor, which
    # lovasz_softmax_flat then answers with probas * 0 (loss_utils.py:250-251)
    return probas[valid.nonzero().squeeze()], labels[valid]


def lovasz_softmax_flat(probas, labels, classes="present"):
    """loss_utils.py:237-275: mean over the (present) classes of dot(sorted |fg - p_c|, lovasz_grad(sorted fg))."""
    if probas.numel() == 0 or probas.dim() < 2:
        return probas * 0.0
    C = probas.shape[1]
    losses = []
    for c in (range(C) if classes in ("all", "present") else classes):
        fg = (labels == c).float()
        if classes == "present" and fg.sum() == 0:
            continue
        class_pred = probas[:, 0] if C == 1 else probas[:, c]
        errors = (fg - class_pred).abs()
        errors_sorted, perm = torch.sort(errors, 0, descending=True)
        losses.append(torch.dot(errors_sorted, lovasz_grad(fg[perm])))
    if not losses:                      # the reference's mean() of an empty list with empty=0 semantics is never reached by
        return probas.sum() * 0.0       # the heads (a frame always has a labelled point); keep the graph connected
    return sum(losses) / len(losses)


def lovasz_softmax(probas, labels, classes="present", per_image=False, ignore=None):
    """loss_utils.py:217-234 (probas are softmax outputs)."""
    if per_image:
        vals = [lovasz_softmax_flat(*flatten_probas(p.unsqueeze(0), l.unsqueeze(0), ignore), classes=classes)
                for p, l in zip(probas, labels)]
        return sum(vals) / len(vals)
    return lovasz_softmax_flat(*flatten_probas(probas, labels, ignore), classes=classes)


def point_head_loss(voxel_logits, voxel_labels, out_logits, point_labels, pcamera=None, camera=None, ignored_label=0):
    """point_seg_mseg3d_head.py:137-196: (CE(ignore) + Lovasz) on voxel logits + the same on point logits + MSE mimic.
    Returns (loss, dict of the detached terms under the reference's keys)."""
    d = {}
    ce = lambda x, y: F.cross_entropy(x, y.long(), ignore_index=ignored_label)
    lv = lambda x, y: lovasz_softmax(F.softmax(x, dim=-1), y.long(), ignore=ignored_label)
    d["voxel_ce_loss"], d["voxel_lovasz_loss"] = ce(voxel_logits, voxel_labels), lv(voxel_logits, voxel_labels)
    d["out_ce_loss"], d["out_lovasz_loss"] = ce(out_logits, point_labels), lv(out_logits, point_labels)
    loss = d["voxel_ce_loss"] + d["voxel_lovasz_loss"] + d["out_ce_loss"] + d["out_lovasz_loss"]
    if pcamera is not None:
        d["out_mimic_loss"] = F.mse_loss(pcamera, camera)
        loss = loss + d["out_mimic_loss"]
    return loss, {k: v.detach() for k, v in d.items()}


def image_head_loss(image_logits, image_sem_labels, loss_weight=0.5, lovasz_loss_weight=-1.0, ignore_index=0,
                    align_corners=False):
    """fcn_mseg3d_head.py:202-244: logits [B*ncam, C, h, w] resized (bilinear) to the label map [B*ncam, 1, H, W] or
    [B*ncam, H, W], loss_weight * CE(ignore) (+ lovasz_loss_weight * Lovasz over all pixels, no ignore)."""
    if image_sem_labels.dim() == 3:
        image_sem_labels = image_sem_labels.unsqueeze(1)
    logits = F.interpolate(image_logits, size=image_sem_labels.shape[2:], mode="bilinear", align_corners=align_corners)
    target = image_sem_labels.squeeze(1).long()
    d = {"image_ce_loss": loss_weight * F.cross_entropy(logits, target, ignore_index=ignore_index)}
    loss = d["image_ce_loss"]
    if lovasz_loss_weight > 0:
        d["image_lvsz_loss"] = lovasz_loss_weight * lovasz_softmax(torch.softmax(logits, dim=1), target)
        loss = loss + d["image_lvsz_loss"]
    return loss, {k: v.detach() for k, v in d.items()}
