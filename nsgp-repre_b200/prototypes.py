"""RePRE prototype build + replay staging, drop-in for
``StandardMultiPrototypeReplayHead`` (mmdet/models/roi_heads/
standard_roi_replay_head.py:375-501).

Everything runs on the device - per-class segmented means (:412-414), L2-normalise +
cosine Gram + ``>= 0.6`` + neighbour counts (:417-421), the density ordering and the
greedy cover (:421-448: one CTA per class; a stable descending rank reproduces the tie
order of torch's CPU sort), masked means (:443), the device-resident replay gather
(:458-463).  The host only learns the class sizes (to size buffers and the grouped
Gram) and, at the end, the number of prototypes; ``save_idx`` (the ``mask.pth``
payload) is fetched when it is asked for.  With the SIMT bring-up engine the ordering
and the cover run on the host with ``torch.sort`` like the reference.
"""
from __future__ import annotations

import os
import os.path as osp

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import lib, check, ptr


def get_work_dir(previous_path: str) -> str:
    """Path rule of :363-370: '..._N' -> '..._N+1' ('coco' paths -> './')."""
    if "coco" in previous_path:
        return "./"
    parts = previous_path.split("_")
    parts[-1] = str(int(parts[-1]) + 1)
    return "_".join(parts)


class MultiPrototypeReplay:
    """Builds and serves the coarse + fine-grained prototypes of the old classes.

    ``build`` mirrors the loop at :404-449; results: ``bbox_featss`` (P,D) fp32 on
    the device, ``tmp_label`` (P,) int64, ``save_idx`` = the ``mask.pth`` payload
    (list[class] of list[<=max_proto-1] bool masks over that class's rows).
    """

    def __init__(self, max_prototype: int = 10, thresh: float = 0.6):
        self.max_proto = max_prototype
        self.thresh = thresh
        self.bbox_featss = None
        self.tmp_label = None
        self._save_idx = None
        self._lazy_masks = None
        self.sigma = None
        self._out = None
        self._ws = None

    # ------------------------------------------------------------------ build
    @torch.no_grad()
    def build(self, feats: torch.Tensor, cls_targets: torch.Tensor, previous_cls,
              saved_masks=None):
        _lib.require_cuda(feats, "bbox_featss")
        dev = feats.device
        feats = feats.detach()
        feats = feats.reshape(feats.shape[0], -1)
        if feats.dtype != torch.float32 or not feats.is_contiguous():
            feats = feats.float().contiguous()
        M, D = feats.shape
        labels = cls_targets.detach().to(device=dev, dtype=torch.int64).contiguous()
        previous_cls = list(previous_cls)
        stream = _lib.current_stream(dev)
        save_idx = list(saved_masks) if saved_masks is not None else []
        if not previous_cls:
            self.bbox_featss = feats.new_zeros(0, D)
            self.tmp_label = torch.zeros(0, dtype=torch.long, device=dev)
            self.save_idx = save_idx
            return self

        # stable class index over [0, C)
        C = max(previous_cls) + 1
        counts = torch.empty(C, dtype=torch.int32, device=dev)
        offsets = torch.empty(C + 1, dtype=torch.int32, device=dev)
        rows = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
        check(lib.repre_class_index(ptr(labels), M, C, ptr(counts), ptr(offsets), ptr(rows),
                                    stream), "repre_class_index")
        h_off = offsets.cpu().tolist()

        sizes = []
        for c in previous_cls:
            n = h_off[c + 1] - h_off[c]
            if n == 0:
                # the reference dies in sim_sum[-0//3] (:422); keep that contract
                raise IndexError("class %d has no stored RoI feature "
                                 "(index 0 is out of bounds for dimension 0 with size 0)" % c)
            sizes.append(n)
        if lib.nsgp_get_engine() == 0:
            return self._build_device(feats, rows, h_off, previous_cls, sizes, save_idx, stream)

        # neighbour masks / counts of every class, launched back to back into ONE
        # device buffer [masks | counts | rows] that comes back in a single D2H copy
        mask_bytes = sum(n * n for n in sizes)
        mask_pad = (mask_bytes + 15) // 16 * 16
        cnt_elems = sum(sizes)
        pack = torch.empty(mask_pad + 4 * cnt_elems + 4 * max(M, 1), dtype=torch.uint8,
                           device=dev)
        cnt_all = pack[mask_pad:mask_pad + 4 * cnt_elems].view(torch.int32)
        pack[mask_pad + 4 * cnt_elems:].view(torch.int32).copy_(rows)
        if lib.nsgp_get_engine() == 0:
            # all classes at once: normalise, clear, ONE grouped tcgen05 Gram, threshold
            import ctypes
            ncls = len(sizes)
            sizes_arr = (ctypes.c_int32 * ncls)(*sizes)
            need = int(lib.repre_cosine_count_batched_workspace_bytes(sizes_arr, ncls, D))
            ws = self._ws if self._ws is not None and self._ws.numel() >= need and \
                self._ws.device == dev else torch.empty(need, dtype=torch.uint8, device=dev)
            self._ws = ws
            consecutive = all(b == a + 1 for a, b in zip(previous_cls, previous_cls[1:]))
            if consecutive:
                rows_sel = rows[h_off[previous_cls[0]]:h_off[previous_cls[-1] + 1]]
            else:
                rows_sel = torch.cat([rows[h_off[c]:h_off[c + 1]] for c in previous_cls])
            check(lib.repre_cosine_count_batched(
                ptr(feats), D, ptr(rows_sel), sizes_arr, ncls, float(self.thresh),
                pack.data_ptr(), cnt_all.data_ptr(), ptr(ws), ws.numel(), stream),
                "repre_cosine_count_batched")
        else:
            need = max(lib.repre_cosine_count_workspace_bytes(n, D) for n in sizes)
            ws = self._ws if self._ws is not None and self._ws.numel() >= need and \
                self._ws.device == dev else torch.empty(int(need), dtype=torch.uint8, device=dev)
            self._ws = ws
            m_off = c_off = 0
            for c, n in zip(previous_cls, sizes):
                check(lib.repre_cosine_count(
                    ptr(feats), D, rows.data_ptr() + 4 * h_off[c], n, float(self.thresh),
                    pack.data_ptr() + m_off, cnt_all.data_ptr() + 4 * c_off, None, ptr(ws),
                    ws.numel(), stream), "repre_cosine_count")
                m_off += n * n
                c_off += n
        host = pack.cpu().numpy()                                   # the one sync
        h_cnt = host[mask_pad:mask_pad + 4 * cnt_elems].view(np.int32)
        h_rows_np = host[mask_pad + 4 * cnt_elems:].view(np.int32)

        # host: density order + greedy cover (sequential, <= max_proto-1 picks)
        seg_rows, seg_off, seg_label = [], [0], []
        m_off = c_off = 0
        for c, n in zip(previous_cls, sizes):
            sim_mask = host[m_off:m_off + n * n].reshape(n, n).astype(bool)
            cnt_np = h_cnt[c_off:c_off + n]
            m_off += n * n
            c_off += n
            cls_rows = h_rows_np[h_off[c]:h_off[c + 1]]
            # coarse prototype: all rows of the class (:412-414)
            seg_rows.append(cls_rows)
            seg_off.append(seg_off[-1] + n)
            seg_label.append(c)
            # stable, like the reference's sort under its pinned torch 1.12 (equal counts keep
            # ascending row order; the device path ranks the same way)
            sim_sum, idx = torch.from_numpy(cnt_np.astype(np.int64)).sort(
                dim=-1, descending=True, stable=True)                    # :421
            thr = int(sim_sum[-n // 3])                                  # :422
            covered = cnt_np <= thr                                      # :423
            idx_np = idx.numpy()
            tmp_mask = save_idx[c] if c < len(save_idx) else []          # :425-428
            for proto_count in range(self.max_proto - 1):                # :430
                if proto_count < len(tmp_mask):                          # replayed mask.pth entry
                    m = tmp_mask[proto_count].cpu().numpy().astype(bool)
                else:
                    cand = idx_np[~covered[idx_np]]
                    if cand.size == 0:
                        continue
                    m = sim_mask[cand[0]]
                    tmp_mask.append(torch.from_numpy(m.copy()))
                covered = covered | m
                sel = cls_rows[m]
                seg_rows.append(sel)
                seg_off.append(seg_off[-1] + sel.size)
                seg_label.append(c)
            if c >= len(save_idx):
                save_idx.append(tmp_mask)

        # one launch: coarse + fine means of every class
        nseg = len(seg_label)
        idx_host = np.concatenate([np.asarray(seg_off, dtype=np.int32),
                                   np.asarray(seg_label, dtype=np.int32)] +
                                  [r.astype(np.int32, copy=False) for r in seg_rows])
        idx_dev = torch.from_numpy(idx_host).to(dev)                 # one H2D copy
        off_t = idx_dev[:nseg + 1]
        all_rows = idx_dev[2 * nseg + 1:]
        out = torch.empty(nseg, D, dtype=torch.float32, device=dev)
        max_rows = max(b - a for a, b in zip(seg_off[:-1], seg_off[1:]))
        check(lib.repre_segment_mean(ptr(feats), D, ptr(off_t), ptr(all_rows), nseg,
                                     int(max_rows), ptr(out), stream), "repre_segment_mean")
        self.bbox_featss = out
        self.tmp_label = idx_dev[nseg + 1:2 * nseg + 1].to(torch.int64)
        self.save_idx = save_idx
        self._segments = (off_t, all_rows, max_rows)
        self._feats = feats
        return self

    def _build_device(self, feats, rows, h_off, previous_cls, sizes, saved_masks, stream):
        """tcgen05 engine: Gram, ordering, cover, segment table and means without leaving
        the device; one small D2H at the end for the number of prototypes."""
        import ctypes
        dev = feats.device
        M, D = feats.shape
        ncls = len(sizes)
        mp = self.max_proto - 1
        sizes_arr = (ctypes.c_int32 * ncls)(*sizes)
        ids_arr = (ctypes.c_int32 * ncls)(*previous_cls)
        n_tot = sum(sizes)
        mask_bytes = sum(n * n for n in sizes)
        need = int(lib.repre_cosine_count_batched_workspace_bytes(sizes_arr, ncls, D)) + \
            int(lib.repre_greedy_segments_workspace_bytes(sizes_arr, ncls, mp)) + 512
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        ws = self._ws
        ws2_off = (int(lib.repre_cosine_count_batched_workspace_bytes(sizes_arr, ncls, D)) + 255) \
            // 256 * 256
        consecutive = all(b == a + 1 for a, b in zip(previous_cls, previous_cls[1:]))
        if consecutive:
            rows_sel = rows[h_off[previous_cls[0]]:h_off[previous_cls[-1] + 1]]
        else:
            rows_sel = torch.cat([rows[h_off[c]:h_off[c + 1]] for c in previous_cls])
        masks = torch.empty(max(mask_bytes, 1), dtype=torch.uint8, device=dev)
        counts = torch.empty(n_tot, dtype=torch.int32, device=dev)
        check(lib.repre_cosine_count_batched(
            ptr(feats), D, ptr(rows_sel), sizes_arr, ncls, float(self.thresh), ptr(masks),
            ptr(counts), ptr(ws), ws2_off, stream), "repre_cosine_count_batched")
        # masks replayed from mask.pth (:425-433): one H2D of the packed bytes
        saved_dev, n_saved_arr = None, None
        saved_lists = []
        for ci, c in enumerate(previous_cls):
            lst = list(saved_masks[c]) if c < len(saved_masks) else []
            saved_lists.append(lst)
        if any(saved_lists):
            n_saved = [min(len(l), mp) for l in saved_lists]
            chunks = []
            for l, k, n in zip(saved_lists, n_saved, sizes):
                for m in l[:k]:
                    m = m.cpu().to(torch.uint8)
                    if m.numel() != n:
                        raise IndexError("replayed mask has %d entries, the class has %d rows "
                                         "(The shape of the mask does not match the tensor)" %
                                         (m.numel(), n))
                    chunks.append(m)
            if chunks:
                saved_dev = torch.cat(chunks).to(dev)
                n_saved_arr = (ctypes.c_int32 * ncls)(*n_saved)
        nseg_max = ncls * (mp + 1)
        seg = torch.empty(2 * nseg_max + 1 + n_tot * (mp + 1) + 1 + ncls + ncls * mp,
                          dtype=torch.int32, device=dev)
        seg_off = seg[:nseg_max + 1]
        seg_label = seg[nseg_max + 1:2 * nseg_max + 1]
        seg_rows = seg[2 * nseg_max + 1:2 * nseg_max + 1 + n_tot * (mp + 1)]
        info = seg[2 * nseg_max + 1 + n_tot * (mp + 1):]
        check(lib.repre_greedy_segments(
            ptr(masks), ptr(counts), ptr(rows_sel), sizes_arr, ids_arr, ncls, mp,
            ptr(saved_dev), n_saved_arr, ptr(seg_off), ptr(seg_rows), ptr(seg_label), ptr(info),
            ws.data_ptr() + ws2_off, ws.numel() - ws2_off, stream), "repre_greedy_segments")
        out = torch.empty(nseg_max, D, dtype=torch.float32, device=dev)
        max_rows = max(sizes)
        check(lib.repre_segment_mean_dev(ptr(feats), D, ptr(seg_off), ptr(seg_rows), nseg_max,
                                         ptr(info), int(max_rows), ptr(out), stream),
              "repre_segment_mean_dev")
        h_info = info.cpu().tolist()                                 # the one late sync
        nseg = h_info[0]
        self.bbox_featss = out[:nseg]
        self.tmp_label = seg_label[:nseg].to(torch.int64)
        self._segments = (seg_off, seg_rows, max_rows)
        self._feats = feats
        self._lazy_masks = (masks, sizes, list(previous_cls), h_info[1:1 + ncls],
                            h_info[1 + ncls:], saved_lists, list(saved_masks), mp)
        self._save_idx = None
        return self

    @property
    def save_idx(self):
        """``mask.pth`` payload (:450-452): list[class] of list of bool masks over the
        class's rows; built from the device picks on first access."""
        if self._save_idx is None and self._lazy_masks is not None:
            masks, sizes, classes, npicks, picks, saved_lists, save_idx, mp = self._lazy_masks
            moff = 0
            for ci, (c, n) in enumerate(zip(classes, sizes)):
                tmp = list(saved_lists[ci])
                mine = [p for p in picks[ci * mp:ci * mp + npicks[ci]] if p >= 0]
                if mine:
                    idx = torch.tensor(mine, dtype=torch.int64, device=masks.device)
                    sel = masks[moff:moff + n * n].view(n, n)[idx].cpu().bool()
                    tmp.extend(sel[k].clone() for k in range(len(mine)))
                moff += n * n
                if c < len(save_idx):
                    save_idx[c] = tmp
                else:
                    save_idx.append(tmp)
            self._save_idx = save_idx
            self._lazy_masks = None
        return self._save_idx

    @save_idx.setter
    def save_idx(self, value):
        self._save_idx = value
        self._lazy_masks = None

    @torch.no_grad()
    def build_sigma(self):
        """Extension (BASELINE north_star item 4; no reference counterpart): per-
        prototype diagonal standard deviation for Gaussian replay."""
        off_t, all_rows, max_rows = self._segments
        nseg, D = self.bbox_featss.shape
        var = torch.empty_like(self.bbox_featss)
        check(lib.repre_segment_var(ptr(self._feats), D, ptr(off_t), ptr(all_rows), nseg,
                                    int(max_rows), ptr(self.bbox_featss), ptr(var),
                                    _lib.current_stream(var.device)), "repre_segment_var")
        self.sigma = var.sqrt_()
        return self.sigma

    # ----------------------------------------------------------------- replay
    @torch.no_grad()
    def staged(self, idx=None, out=None, sigma=None, seed=0):
        """Classifier input of the replay branch (:458-463): the prototypes are
        device-resident; this gathers them (all of them when ``idx`` is None, as
        the reference stages every prototype every step) into ``out``."""
        protos = self.bbox_featss
        P = protos.shape[0] if idx is None else idx.shape[0]
        D = protos.shape[1]
        if out is None:
            if self._out is None or self._out.shape != (P, D):
                self._out = torch.empty(P, D, dtype=torch.float32, device=protos.device)
            out = self._out
        if idx is not None:
            idx = idx.to(device=protos.device, dtype=torch.int64).contiguous()
        check(lib.repre_replay_gather(ptr(protos), ptr(sigma), ptr(idx), P, D, int(seed),
                                      ptr(out), _lib.current_stream(protos.device)),
              "repre_replay_gather")
        return out


@torch.no_grad()
def kmeans_prototypes(feats: torch.Tensor, init_centres: torch.Tensor, iters: int = 10):
    """Extension (BASELINE north_star item 3; the reference imports sklearn's KMeans at
    standard_roi_replay_head.py:18 and never calls it - parity unpinned, checked against
    ``oracle.restated.kmeans_assign / kmeans_update``): Lloyd iterations from explicit
    initial centres, entirely on the device.  Assignment = argmin_k |x - c_k|^2 (ties ->
    lowest k) through the tcgen05 contraction X C^T, update = segmented mean of the rows of
    every cluster (an empty cluster keeps its centre).  Returns (centres (k,D), labels (n,))."""
    _lib.require_cuda(feats, "features")
    dev = feats.device
    x = feats.detach().reshape(feats.shape[0], -1)
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.float().contiguous()
    n, D = x.shape
    centres = init_centres.detach().to(device=dev, dtype=torch.float32).reshape(-1, D).clone()
    k = centres.shape[0]
    stream = _lib.current_stream(dev)
    ws = torch.empty(int(lib.repre_kmeans_assign_workspace_bytes(n, k, D)), dtype=torch.uint8,
                     device=dev)
    labels = torch.empty(n, dtype=torch.int64, device=dev)
    counts = torch.empty(k, dtype=torch.int32, device=dev)
    offsets = torch.empty(k + 1, dtype=torch.int32, device=dev)
    rows = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    new = torch.empty_like(centres)
    for _ in range(max(1, int(iters))):
        check(lib.repre_kmeans_assign(ptr(x), n, D, ptr(centres), k, ptr(labels), ptr(ws),
                                      ws.numel(), stream), "repre_kmeans_assign")
        check(lib.repre_class_index(ptr(labels), n, k, ptr(counts), ptr(offsets), ptr(rows),
                                    stream), "repre_class_index")
        check(lib.repre_segment_mean(ptr(x), D, ptr(offsets), ptr(rows), k, n, ptr(new), stream),
              "repre_segment_mean")
        centres = torch.where((counts > 0).unsqueeze(1), new, centres)
    check(lib.repre_kmeans_assign(ptr(x), n, D, ptr(centres), k, ptr(labels), ptr(ws),
                                  ws.numel(), stream), "repre_kmeans_assign")
    return centres, labels


class StandardMultiPrototypeReplayHead(nn.Module):
    """Constructor keywords and attributes of the reference head (:377-390):
    ``previous_path, task_id, task_split, max_prototype, work_dir``; attributes
    ``replay``, ``bbox_featss``, ``tmp_label``; ``loss`` adds ``replay_loss_cls``.

    Stand-alone form: ``bbox_head`` is any module mapping (P,12544)-features to
    ``(cls_score, bbox_pred)``.  When mmdet is importable ``registry.py`` builds
    the real subclass of ``StandardRoIHead`` from the same mixin logic.
    """

    def __init__(self, bbox_head: nn.Module = None, previous_path=None, task_id=1,
                 task_split=(0, 10, 20), max_prototype=10, work_dir=None, device=None,
                 **kwargs):
        super().__init__()
        self.bbox_head = bbox_head
        self.replay = False
        self.task_split = list(task_split)
        self.task_id = task_id
        self.max_proto = max_prototype
        self.with_shared_head = False
        self._proto = MultiPrototypeReplay(max_prototype)
        if previous_path is not None and osp.exists(previous_path):
            assert task_id != 1
            self.replay = True
            dev = torch.device(device) if device is not None else torch.device("cuda")
            (bbox_featss, self.cls_targets, self.cls_weights, self.bbox_targets,
             self.bbox_weights, self.roiss) = torch.load(
                osp.join(previous_path, "rois_etc.pth"), map_location=dev)
            previous_cls = range(self.task_split[0], self.task_split[task_id - 1])
            saved = None
            if osp.exists(osp.join(previous_path, "mask.pth")):
                saved = torch.load(osp.join(previous_path, "mask.pth"), map_location="cpu")
            self._proto.build(bbox_featss, self.cls_targets, previous_cls, saved)
            self.bbox_featss = self._proto.bbox_featss
            self.tmp_label = self._proto.tmp_label
            out_dir = work_dir if work_dir is not None else get_work_dir(previous_path)
            torch.save(self._proto.save_idx, osp.join(out_dir, "mask.pth"))

    def replay_loss(self, bbox_feats, sampling_results=None, rois=None) -> dict:
        """:468-501 - logits of classes < task_split[task_id] plus background;
        cross-entropy on the softmax output (double softmax kept on purpose)."""
        cls_score, bbox_pred = self.bbox_head(bbox_feats)
        pre_idx = self.task_split[self.task_id]
        kept = torch.cat([cls_score[:, :pre_idx], cls_score[:, -1:]], dim=-1)
        losses = {"replay_loss_cls": F.cross_entropy(kept.softmax(dim=-1),
                                                     self.tmp_label.to(kept.device))}
        return dict(cls_score=cls_score, bbox_pred=bbox_pred, bbox_feats=bbox_feats,
                    replay_loss=losses)

    def loss(self, x=None, rpn_results_list=None, batch_data_samples=None, base_losses=None):
        """:454-466.  ``base_losses`` stands for ``super().loss(...)``."""
        losses = dict(base_losses or {})
        if self.replay:
            staged = self._proto.staged()
            losses.update(self.replay_loss(staged)["replay_loss"])
        return losses
