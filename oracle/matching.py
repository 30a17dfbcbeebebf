"""torch restatement of the reference's segment-wise matching loss -- TEST INFRASTRUCTURE ONLY.

Follows loss.py:14-36 (compute_centroid, labels_downscale, cosine_dists), :93-112 (NNFMStyleLoss.forward), :147-170
(SemanticStyleLoss.init_feats), :172-185 (update_matching) and :187-214 (forward).  Pinned by tests/golden/matching.npz,
which holds outputs of the reference's own unmodified loss.py (tests/golden/make_matching_golden.py).
"""
import numpy as np
import torch
import torch.nn.functional as F


def compute_centroid(mask):
    """loss.py:14-20"""
    H, W = mask.shape
    N = torch.sum(mask)
    r_mean = torch.sum(torch.sum(mask, dim=1) * torch.arange(H)) / N / H
    c_mean = torch.sum(torch.sum(mask, dim=0) * torch.arange(W)) / N / W
    return torch.stack((r_mean, c_mean))


def labels_downscale(labels, new_dim):
    """loss.py:23-28: nearest sample at linspace(0, H-1, NH) truncated to integers."""
    H, W = labels.shape
    NH, NW = new_dim
    r = torch.linspace(0, H - 1, NH).long()
    c = torch.linspace(0, W - 1, NW).long()
    return labels[r[:, None], c]


def clusters_downscale(clusters, size):
    """loss.py:157-158: F.interpolate (nearest) of the style segmentation to the style feature map."""
    return F.interpolate(clusters[None, None].float(), size)[0, 0].to(torch.long)


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


def hungarian_matching(image_feat, preds, style_feat, clusters_small, num_classes):
    """loss.py:160-185: mean feature per class / cluster + centroid distance -> linear_sum_assignment."""
    from scipy.optimize import linear_sum_assignment
    preds_small = labels_downscale(preds, image_feat.shape[-2:])
    image_mean = torch.stack([torch.mean(image_feat[:, preds_small == i], dim=1) for i in range(num_classes)])
    image_cent = torch.stack([compute_centroid(preds == i) for i in range(num_classes)])
    n_clusters = int(clusters_small.max()) + 1
    style_mean = torch.stack([torch.mean(style_feat[:, clusters_small == i], dim=1) for i in range(n_clusters)])
    style_cent = torch.stack([compute_centroid(clusters_small == i) for i in range(n_clusters)])
    cost = cosine_dists(image_mean, style_mean) + torch.linalg.norm(image_cent[:, None] - style_cent[None], dim=-1)
    return linear_sum_assignment(np.nan_to_num(cost.detach().numpy()))[1]


def semantic_style_loss(image_feat, style_feat, preds, clusters_full, matching, num_classes):
    """SemanticStyleLoss.init_feats + forward (loss.py:147-214) for one feature key.  image_feat [C,h,w], style_feat
    [C,hs,ws], preds [H,W] full-resolution class map, clusters_full [Hs,Ws]; matching None = Hungarian (update_matching).
    Returns (loss, min_dists, argmin, preds_small, clusters_small, matching)."""
    clusters_small = clusters_downscale(clusters_full, style_feat.shape[1:])
    if matching is None:
        matching = hungarian_matching(image_feat, preds, style_feat, clusters_small, num_classes)
    preds_small = labels_downscale(preds, image_feat.shape[-2:])
    C = image_feat.shape[0]
    a = image_feat.reshape(C, -1).t()
    b = style_feat.reshape(C, -1).t()
    loss, md, arg = semantic_nn_loss(a, b, preds_small, clusters_small, matching, num_classes)
    return loss, md, arg, preds_small, clusters_small, matching


def nnfm_loss(image_feat, style_feat):
    """NNFMStyleLoss.forward (loss.py:93-112) for one key: unmasked nearest-neighbour cosine distance."""
    C = image_feat.shape[0]
    return semantic_nn_loss(image_feat.reshape(C, -1).t(), style_feat.reshape(C, -1).t())[0]
