"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Restatement of the detection post-processing that follows the forward
in the reference's inference.py:108-142 with NMS() of utils.py:312-317 (torch CPU + the oracle's nms).
Parity note: inference.py itself cannot be imported here (it pulls datasets / pycocotools), so this file is
pinned through its building blocks: bbox_transform_inv, clip_boxes and nms are each pinned to the reference in
tests/test_oracle_pins.py, and the sort + NMS composition (lines marked utils.py below) is pinned to the reference's
own NMS() exec'd from the source text of utils.py:312-317 (oracle/make_golden.py -> tests/golden/postprocess.npz,
tests/test_oracle_pins.py::test_postprocess_vs_reference_NMS)."""
import torch

import dana_oracle as O


def detections(rois, cls_prob, bbox_pred, im_info, score_thresh=0.05, nms_thresh=0.3,
               stds=(0.1, 0.1, 0.2, 0.2), means=(0.0, 0.0, 0.0, 0.0)):
    """rois [B,R,5], cls_prob [B*R,2], bbox_pred [B*R,4], im_info [B,3] -> list over images of dets [M,5]
    (x1,y1,x2,y2,score), sorted by score, after NMS.  (The reference runs batch 1; this loops over images.)"""
    b, r, _ = rois.shape
    out = []
    for i in range(b):
        boxes = rois[i:i + 1, :, 1:5]
        deltas = bbox_pred[i * r:(i + 1) * r].view(-1, 4) * torch.tensor(stds) + torch.tensor(means)   # :114-115
        pred = O.bbox_transform_inv(boxes, deltas.view(1, -1, 4))                                       # :119
        pred = O.clip_boxes(pred, im_info[i:i + 1])                                                     # :120
        pred = pred / im_info[i, 2]                                                                     # :123
        scores = cls_prob[i * r:(i + 1) * r, 1]
        pred = pred.squeeze(0)
        inds = torch.nonzero(scores > score_thresh).view(-1)                                            # :130
        if inds.numel() == 0:
            out.append(torch.zeros(0, 5))
            continue
        cs, cb = scores[inds], pred[inds]
        order = torch.sort(cs, dim=0, descending=True, stable=True)[1]                                    # utils.py:313
        dets = torch.cat((cb, cs.unsqueeze(1)), 1)[order]
        keep = O.nms(cb[order], cs[order], nms_thresh)                                                  # :315
        out.append(dets[keep.view(-1).long()])
    return out
