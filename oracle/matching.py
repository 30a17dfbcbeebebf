"""torch restatement of the reference's matching loss (loss.py:32-36, 199-214) -- TEST INFRASTRUCTURE ONLY."""
import torch


def cosine_dists(feats1, feats2):
    """loss.py:32-36"""
    f1 = feats1 / torch.linalg.norm(feats1, dim=1)[:, None]
    f2 = feats2 / torch.linalg.norm(feats2, dim=1)[:, None]
    return 1.0 - torch.matmul(f1, f2.T)


def semantic_nn_loss(image_feat_nc, style_feat_nc, preds_small=None, clusters=None, matching=None, num_classes=0):
    """loss.py:199-214: returns (loss, min_dists, argmin)."""
    dists = cosine_dists(image_feat_nc, style_feat_nc)
    if matching is not None:
        for i in range(num_classes):
            image_mask = (preds_small == i).reshape(-1)
            style_mask = (clusters != matching[i]).reshape(-1)
            invalid = image_mask[:, None] & style_mask[None, :]
            dists[invalid] = float('inf')
    min_dists, arg = torch.min(dists, dim=1)
    return torch.mean(min_dists), min_dists, arg
