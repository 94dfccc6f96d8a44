"""TEST INFRASTRUCTURE ONLY - CPU restatement of the NSGP-RePRE hot path.

Every function restates one piece of the reference (yyl404/NSGP-RePRE) and cites
the file:line it follows (paths relative to the reference root).  Arithmetic is
torch-CPU fp32 / numpy, the same libraries the reference itself runs on when it
is executed on CPU, so that the restatement can be pinned bit-for-bit (or to
fp32 round-off) against the reference's own functions - see
``oracle/make_golden.py`` and ``tests/test_oracle_golden.py``.

Parity pinning: the reference has no tests / golden vectors for this path
(SURVEY.md 8c); pinned instead against outputs of the reference's own functions
run in the build container (``tests/golden/``).

Must not be imported by the product package.
"""
from __future__ import annotations

import numpy as np
import scipy.ndimage
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- #
# a1 / a2 : per-layer input covariance
# --------------------------------------------------------------------------- #


def conv_rows(x: torch.Tensor, kernel_size, stride, padding) -> torch.Tensor:
    """Rows X (N, d) the reference feeds to ``update_cov`` for a Conv2d input.

    mmdet/engine/runner/nsrunner_roi_replay.py:903-913 - the batch mean is taken
    FIRST (one averaged map), then unfolded; column order is (Cin, kh, kw), i.e.
    the order of ``weight.view(Cout, -1)``.
    """
    mean_map = x.mean(dim=0, keepdim=True)
    cols = F.unfold(mean_map, kernel_size=kernel_size, padding=padding, stride=stride)
    return cols[0].t().contiguous()          # (1,d,N) -> (N,d)


def linear_rows(x: torch.Tensor) -> torch.Tensor:
    """nsrunner_roi_replay.py:900-901 - one (1,d) row: the batch-mean input."""
    return x.mean(dim=0, keepdim=True)


def gram(rows: torch.Tensor) -> torch.Tensor:
    """nsrunner_roi_replay.py:930 - un-centred, un-normalised X^T X."""
    return rows.t() @ rows


def cov_conv2d(x, kernel_size, stride, padding):
    return gram(conv_rows(x, kernel_size, stride, padding))


def cov_linear(x):
    return gram(linear_rows(x))


def accumulate(fea_in: dict, key: str, cov: torch.Tensor) -> None:
    """nsrunner_roi_replay.py:931-934 - first call stores, later calls add."""
    fea_in[key] = cov if key not in fea_in else fea_in[key] + cov


def covariances_of_model(model: torch.nn.Module, batches) -> dict:
    """Restated ``cal_fea_in`` accumulation loop (nsrunner_roi_replay.py:731-744,
    876-916): forward hooks on every module that has a ``weight``; Conv2d and
    Linear inputs contribute, everything else (BatchNorm ...) is a no-op.
    Key = "<named_modules path>.weight" (:893-896)."""
    fea_in: dict = {}
    names = {m: n for n, m in model.named_modules()}

    def hook(module, inputs, _out):
        x = inputs[0]
        key = names[module] + ".weight"
        if isinstance(module, torch.nn.Linear):
            accumulate(fea_in, key, cov_linear(x))
        elif isinstance(module, torch.nn.Conv2d):
            accumulate(fea_in, key, cov_conv2d(
                x, module.kernel_size, module.stride, module.padding))

    handles = [m.register_forward_hook(hook)
               for _, m in model.named_modules() if hasattr(m, "weight")]
    was_training = model.training
    model.eval()
    with torch.no_grad():
        for x in batches:
            model(x)
    for h in handles:
        h.remove()
    model.train(was_training)
    return fea_in


# --------------------------------------------------------------------------- #
# a4 / a5 / a6 : projector build
# --------------------------------------------------------------------------- #


def eigens(cov: torch.Tensor):
    """mmdet/engine/optimizers/SGD_NSCL.py:377 - full SVD of the PSD covariance;
    returns (singular values descending, right singular vectors as columns)."""
    _, s, v = torch.svd(cov, some=False)
    return s, v


def threshold_index(svals: np.ndarray, offset: float = 0.0) -> int:
    """SGD_NSCL.py:134-170 - elbow of the spectrum.

    d >= 128: gaussian-smooth (sigma=10), first and second differences, drop
    int(d*0.03/2) entries at each end of the 2nd difference, take the argmax and
    map it back with + (d - len(valid))//2; d < 128: raw differences.
    ``i_thres`` is the LAST index whose value is >= the elbow value; then the
    offset rule of :164-170.
    """
    pts = np.asarray(svals)
    assert pts.ndim == 1
    d = len(pts)
    if d >= 128:
        smooth = scipy.ndimage.gaussian_filter1d(pts, sigma=10)
        d1 = smooth[:-1] - smooth[1:]
        d2 = d1[:-1] - d1[1:]
        drop = int(d * 0.03 / 2)
        assert d - drop >= 10
        valid = d2[drop:-drop]
        elbow = pts[np.argmax(valid) + int((d - len(valid)) / 2)]
    else:
        d1 = pts[:-1] - pts[1:]
        d2 = d1[:-1] - d1[1:]
        elbow = pts[np.argmax(d2) + int((d - len(d2)) / 2)]
    i_thres = int(np.nonzero(pts >= elbow)[0].max())
    if -1 <= offset <= 1:
        i_thres = min(i_thres + int(offset * i_thres), d - 1)
        i_thres = max(0, i_thres)
    else:
        i_thres = max(min(i_thres + int(offset), d - 1), 0)
    return i_thres


def null_mask(svals: torch.Tensor, offset: float = 0.0) -> torch.Tensor:
    """SGD_NSCL.py:172-177 - boolean mask keeping indices >= i_thres (the
    small-sigma / null-space side; the docstrings there say the opposite)."""
    i = threshold_index(svals.detach().cpu().numpy(), offset)
    m = torch.zeros(svals.shape[0], dtype=torch.bool)
    m[i:] = True
    return m


def transform(svals: torch.Tensor, vecs: torch.Tensor, name: str,
              offset: float = 0.0) -> torch.Tensor:
    """SGD_NSCL.py:254-285 - P = V0 V0^T over the masked columns; names that
    contain 'backbone' are divided by the Frobenius norm of P."""
    basis = vecs[:, null_mask(svals, offset)]
    p = basis @ basis.t()
    if "backbone" in name:
        p = p / torch.norm(p)
    return p


def transforms_from_covariances(fea_in: dict, names, offset: float = 0.0) -> dict:
    """get_eigens + get_transforms (SGD_NSCL.py:360-380, 235-290) for the
    parameter names that have a covariance."""
    out = {}
    for n in names:
        if n in fea_in:
            s, v = eigens(fea_in[n])
            out[n] = transform(s, v, n, offset)
    return out


# --------------------------------------------------------------------------- #
# a7 / a8 : SGD update + projection
# --------------------------------------------------------------------------- #


class SGDNSCLState:
    """Per-parameter state of SGD_NSCL.py:389-397 (``step``, ``previous_grad``)."""

    def __init__(self, p):
        self.step = 0
        self.previous_grad = torch.zeros_like(p)


def sgd_update(state: SGDNSCLState, grad: torch.Tensor, p: torch.Tensor, *, lr,
               momentum=0.0, dampening=0.0, nesterov=False, weight_decay=0.0):
    """SGD_NSCL.py:387-415.  NB side effects kept: ``grad`` is modified in place
    by the weight-decay term (:399-400); first step buf = grad (:405-406)."""
    state.step += 1
    if weight_decay != 0:
        grad.add_(p, alpha=weight_decay)
    if momentum != 0:
        buf = state.previous_grad
        if state.step > 1:
            buf.mul_(momentum).add_(grad, alpha=1 - dampening)
        else:
            buf.add_(grad)
        if nesterov:
            grad.add_(buf, alpha=momentum)
        else:
            grad = buf
    return -(lr * grad)


def sgd_nscl_step(params: dict, grads: dict, states: dict, transforms: dict, *, lr,
                  momentum=0.0, dampening=0.0, nesterov=False, weight_decay=0.0,
                  svd=True) -> None:
    """SGD_NSCL.py:59-96 - for protected names the update is RIGHT-multiplied by
    the transform on its (Cout, -1) view, then added to the weight."""
    for n, p in params.items():
        if n not in states:
            states[n] = SGDNSCLState(p)
        upd = sgd_update(states[n], grads[n], p, lr=lr, momentum=momentum,
                         dampening=dampening, nesterov=nesterov,
                         weight_decay=weight_decay)
        if svd and len(transforms) > 0 and n in transforms:
            if upd.dim() == 4:
                upd = (upd.reshape(upd.shape[0], -1) @ transforms[n]).view_as(upd)
            else:
                upd = upd @ transforms[n]
        p.add_(upd)


# --------------------------------------------------------------------------- #
# a9 / a10 : RePRE prototypes
# --------------------------------------------------------------------------- #


def class_mean(feats: torch.Tensor, labels: torch.Tensor, c: int) -> torch.Tensor:
    """standard_roi_replay_head.py:412-413 - coarse prototype of class c."""
    return feats[labels == c].mean(dim=0, keepdim=True)


def cosine_neighbour_mask(fc: torch.Tensor, thresh: float = 0.6):
    """standard_roi_replay_head.py:417-423 - rows L2-normalised, Gram, >= 0.6,
    row counts."""
    fc = fc.reshape(fc.shape[0], -1)
    fn = fc / fc.norm(dim=-1, keepdim=True)
    sim = fn @ fn.t()
    mask = sim >= thresh
    return sim, mask, mask.long().sum(dim=-1)


def greedy_masks(mask: torch.Tensor, counts: torch.Tensor, max_proto: int,
                 saved=None):
    """standard_roi_replay_head.py:421-448 - density-ordered greedy cover.

    ``order`` is the STABLE descending sort of the counts: the reference's call
    (``.sort(dim=-1, descending=True)``) is a stable sort under its pinned torch 1.12;
    torch >= 2.x leaves the tie order of the non-stable call unspecified
    (oracle/ref_loader.py::stable_sort_ties).  The lowest-density
    third (counts <= sorted[(-n)//3]) starts out as already covered.  Up to
    ``max_proto-1`` picks; a saved mask list (mask.pth from the previous task)
    is replayed first (:432-433)."""
    n = counts.shape[0]
    ordered, order = counts.sort(dim=-1, descending=True, stable=True)
    thr = ordered[-n // 3]
    covered = counts <= thr
    picked = saved if saved is not None else []      # mutated in place, as :426,440
    used = []
    for k in range(max_proto - 1):
        if n == 0:
            break
        if k < len(picked):
            m = picked[k]
        else:
            m = None
            for i in order.tolist():
                if not bool(covered[i]):
                    m = mask[i]
                    break
            if m is None:
                continue
            picked.append(m)
        covered = covered | m
        used.append(m)
    return used, picked


def build_prototypes(feats: torch.Tensor, labels: torch.Tensor, previous_cls,
                     max_proto: int = 10, saved_masks=None, thresh: float = 0.6):
    """standard_roi_replay_head.py:404-452: per old class the coarse mean followed
    by up to max_proto-1 masked means.  Returns (protos (P,D), tmp_label (P,)
    int64, save_idx list[class] of list[mask])."""
    save_idx = list(saved_masks) if saved_masks is not None else []
    protos, tmp_label = [], []
    for c in previous_cls:
        sel = labels == c
        fc = feats[sel]
        protos.append(fc.mean(dim=0, keepdim=True))
        tmp_label.append(c)
        _, mask, counts = cosine_neighbour_mask(fc, thresh)
        saved = save_idx[c] if c < len(save_idx) else None
        used, picked = greedy_masks(mask, counts, max_proto, saved)
        for m in used:
            protos.append(fc[m].mean(dim=0, keepdim=True))
            tmp_label.append(c)
        if c >= len(save_idx):
            save_idx.append(picked)
    return torch.cat(protos, dim=0), torch.tensor(tmp_label, dtype=torch.long), save_idx


# --------------------------------------------------------------------------- #
# a11 : replay staging + loss
# --------------------------------------------------------------------------- #


def replay_gather(protos: torch.Tensor, idx: torch.Tensor | None = None) -> torch.Tensor:
    """standard_roi_replay_head.py:458-463 - every step ALL prototypes are staged
    as the classifier input (idx=None); :58-59 is the sampled variant
    (idx = randperm(M)[:64])."""
    return protos if idx is None else protos[idx]


def replay_loss(cls_score: torch.Tensor, tmp_label: torch.Tensor, pre_idx: int):
    """standard_roi_replay_head.py:497-499 - keep logits of classes < pre_idx plus
    the background column, cross-entropy applied to the SOFTMAX output (double
    softmax, reproduced on purpose)."""
    kept = torch.cat([cls_score[:, :pre_idx], cls_score[:, -1:]], dim=-1)
    return F.cross_entropy(kept.softmax(dim=-1), tmp_label)


# --------------------------------------------------------------------------- #
# Extensions named by BASELINE.json north_star that the reference never calls
# (KMeans / torch.distributions are imported at standard_roi_replay_head.py:18,21
# and unused).  PARITY UNPINNED: defined here as plain torch so the CUDA kernels
# have a checker at all.
# --------------------------------------------------------------------------- #


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Philox4x32-10 (Salmon et al. 2011), counter (n,4) uint32, key (2,) uint32."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = 0x9E3779B9, 0xBB67AE85
    c = counter.astype(np.uint64).copy()
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0 = M0 * c[:, 0]
        p1 = M1 * c[:, 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(0xFFFFFFFF)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(0xFFFFFFFF)
        n0 = hi1 ^ c[:, 1] ^ np.uint64(k0)
        n2 = hi0 ^ c[:, 3] ^ np.uint64(k1)
        c = np.stack([n0, lo1, n2, lo0], axis=1)
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


def gaussian_noise(seed: int, rows: int, cols: int) -> np.ndarray:
    """Counter-based N(0,1) noise keyed by (seed, row, col/4): one Philox block
    per 4 consecutive columns, Box-Muller on (u0,u1) and (u2,u3) with
    u = (x + 0.5) * 2^-32.  cols must be a multiple of 4."""
    assert cols % 4 == 0
    r = np.repeat(np.arange(rows, dtype=np.uint32), cols // 4)
    q = np.tile(np.arange(cols // 4, dtype=np.uint32), rows)
    ctr = np.stack([q, r, np.zeros_like(q), np.zeros_like(q)], axis=1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    x = philox4x32_10(ctr, key).astype(np.float64)
    u = ((x + 0.5) * (1.0 / 4294967296.0)).astype(np.float32)
    rad0 = np.sqrt(-2.0 * np.log(u[:, 0]))
    rad1 = np.sqrt(-2.0 * np.log(u[:, 2]))
    two_pi = np.float32(6.283185307179586)
    z = np.stack([rad0 * np.cos(two_pi * u[:, 1]), rad0 * np.sin(two_pi * u[:, 1]),
                  rad1 * np.cos(two_pi * u[:, 3]), rad1 * np.sin(two_pi * u[:, 3])],
                 axis=1).astype(np.float32)
    return z.reshape(rows, cols)


def replay_gather_gaussian(mu, sigma, idx, seed):
    """out[p] = mu[idx[p]] + sigma[idx[p]] * eps(seed, p, :) - extension."""
    eps = torch.from_numpy(gaussian_noise(seed, idx.shape[0], mu.shape[1]))
    return mu[idx] + sigma[idx] * eps


def class_variance(feats, labels, num_classes):
    """Per-class diagonal covariance (biased, /n) - extension."""
    D = feats.shape[1]
    mean = torch.zeros(num_classes, D)
    var = torch.zeros(num_classes, D)
    cnt = torch.zeros(num_classes, dtype=torch.long)
    for c in range(num_classes):
        fc = feats[labels == c]
        cnt[c] = fc.shape[0]
        if fc.shape[0] > 0:
            mean[c] = fc.mean(0)
            var[c] = ((fc - mean[c]) ** 2).mean(0)
    return mean, var, cnt


def kmeans_assign(x, centres):
    """Lloyd assignment: argmin_k ||x - c_k||^2, ties -> lowest k - extension."""
    d2 = (x * x).sum(1, keepdim=True) - 2.0 * (x @ centres.t()) + (centres * centres).sum(1)[None]
    return d2.argmin(dim=1), d2


def kmeans_update(x, assign, k, old_centres):
    """Lloyd update: segmented mean; empty clusters keep their centre - extension."""
    out = old_centres.clone()
    for j in range(k):
        sel = assign == j
        if sel.any():
            out[j] = x[sel].mean(0)
    return out


# --------------------------------------------------------------------------- #
# SURVEY 8(f)-3 : teacher pseudo-label merge
# --------------------------------------------------------------------------- #


def pseudo_label_merge(gt_boxes, gt_labels, ps_boxes, ps_scores, ps_labels,
                       rpn_thresh: float = 0.5, roi_thresh: float = 0.7,
                       iou_thresh: float = 0.7):
    """faster_rcnn_roi_replay.py:78-108, per image and per teacher box IN ORDER:
    ``max_iou`` = max torchvision ``box_iou`` against the RoI ground-truth set, which
    grows by every teacher box accepted for the RoI head (:104-106) - 0.0 while that set
    is empty (:86-90); ``max_iou > 0.7`` (python float compare, :93) drops the box;
    ``score > rpn_thresh`` / ``score > roi_thresh`` (fp32 tensor compares, :101,:105)
    append it to the RPN / RoI targets.  Returns per image
    (rpn_boxes, rpn_labels, roi_boxes, roi_labels)."""
    from torchvision.ops import box_iou
    out = []
    for gb, gl, pb, psc, pl in zip(gt_boxes, gt_labels, ps_boxes, ps_scores, ps_labels):
        roi_b, roi_l = gb.clone(), gl.clone()
        rpn_b, rpn_l = gb.clone(), gl.clone()
        for k in range(pb.shape[0]):
            box = pb[k:k + 1]
            max_iou = box_iou(box, roi_b).max().item() if roi_b.shape[0] > 0 else 0.0
            if max_iou > iou_thresh:
                continue
            if bool(psc[k] > rpn_thresh):
                rpn_b = torch.cat([rpn_b, box])
                rpn_l = torch.cat([rpn_l, pl[k:k + 1]])
            if bool(psc[k] > roi_thresh):
                roi_b = torch.cat([roi_b, box])
                roi_l = torch.cat([roi_l, pl[k:k + 1]])
        out.append((rpn_b, rpn_l, roi_b, roi_l))
    return out


# --------------------------------------------------------------------------- #
# SURVEY 8(f)-4 : EWC importance + penalty
# --------------------------------------------------------------------------- #


def ewc_register_params(model, must_names=("bn",), ignore_names=("teacher_model",)):
    """nsrunner_roi_replay.py:1006-1031: parameters whose name contains a ``must``
    substring and no ``ignore`` substring."""
    reg = {}
    for n, p in model.named_parameters():
        if any(i in n for i in ignore_names):
            continue
        if len(must_names) == 0 or any(m in n for m in must_names):
            reg[n] = p
    return reg


def ewc_accumulate(importance: dict, grads: dict, len_data_batch: int, len_dataloader: int):
    """:978-981: ``importance[n] += grad**2 * len(data_batch) / len(dataloader)`` for every
    registered parameter that has a gradient (in place, fp32, this operation order)."""
    for n, p in importance.items():
        g = grads.get(n)
        if g is not None:
            p += (g ** 2) * len_data_batch / len_dataloader
    return importance


def ewc_penalty(reg_params: dict, ewc_reg_terms: dict, coeff: float = 1000.0):
    """:1056-1069: ``coeff * sum_n (importance_n * (p_n - old_n)**2).sum()`` with the
    per-task terms concatenated along a new leading axis; differentiable in ``p``.
    Evaluated in fp64 (the checker's value; the reference sums in fp32)."""
    total = 0.0
    for n, p in reg_params.items():
        if not p.requires_grad:
            continue
        imp = torch.cat(ewc_reg_terms["importance"][n], dim=0).double()
        old = torch.cat(ewc_reg_terms["task_param"][n], dim=0).double()
        new = p.double().unsqueeze(0).expand(old.shape)
        total = total + coeff * (imp * (new - old) ** 2).sum()
    return total


# --------------------------------------------------------------------------- #
# SURVEY 8(f)-2 : RoIAlign over the FPN levels (+ per-class sums)
# --------------------------------------------------------------------------- #


def map_roi_levels(rois: torch.Tensor, num_levels: int, finest_scale: float = 56):
    """single_level_roi_extractor.py:45-63."""
    scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
    lv = torch.floor(torch.log2(scale / finest_scale + 1e-6))
    return lv.clamp(min=0, max=num_levels - 1).long()


def roi_extract(feats, rois, featmap_strides=(4, 8, 16, 32), output_size=7, sampling_ratio=0,
                finest_scale=56):
    """:65-118 with mmcv.ops.RoIAlign (absent, un-vendored) restated by
    torchvision.ops.roi_align - the same published algorithm (aligned=True, average
    pooling, sampling grid ceil(roi/output) when sampling_ratio = 0)."""
    from torchvision.ops import roi_align
    C = feats[0].shape[1]
    out = feats[0].new_zeros(rois.shape[0], C, output_size, output_size)
    lv = map_roi_levels(rois, len(feats), finest_scale) if len(feats) > 1 else \
        torch.zeros(rois.shape[0], dtype=torch.long)
    for i, f in enumerate(feats):
        inds = (lv == i).nonzero().squeeze(1)
        if inds.numel():
            out[inds] = roi_align(f, rois[inds], (output_size, output_size),
                                  1.0 / featmap_strides[i], sampling_ratio, True)
    return out, lv


def roi_class_sums(roi_feats: torch.Tensor, labels: torch.Tensor, num_classes: int):
    """Per-class sums and counts of flattened RoI features; ``sums / counts`` is the coarse
    prototype of standard_roi_replay_head.py:411-415.  fp64 accumulation (checker)."""
    F = roi_feats.reshape(roi_feats.shape[0], -1).double()
    sums = torch.zeros(num_classes, F.shape[1], dtype=torch.float64)
    counts = torch.zeros(num_classes, dtype=torch.int32)
    for c in range(num_classes):
        sel = labels == c
        sums[c] = F[sel].sum(0)
        counts[c] = int(sel.sum())
    return sums, counts
