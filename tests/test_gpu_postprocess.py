"""GPU parity of dana_b200.postprocess.detections (C ABI: dana_detections) against the oracle restatement of
inference.py:108-142.  Kept sets must agree exactly (NMS on identical boxes up to <= 2 ulp of expf); box
coordinates within 1e-3 px; scores are passed through untouched (bit-exact)."""
import numpy as np
import pytest
import torch

import postprocess_oracle as PO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b,r,scale", [(1, 300, 1.0), (4, 300, 1.6), (2, 1000, 0.75)])
def test_detections_vs_oracle(b, r, scale):
    import dana_b200  # noqa: F401
    from dana_b200.postprocess import detections
    rs = np.random.RandomState(b * 1000 + r)
    x1, y1 = rs.uniform(0, 900, (b, r)), rs.uniform(0, 500, (b, r))
    w, h = rs.uniform(8, 300, (b, r)), rs.uniform(8, 300, (b, r))
    rois = np.stack([np.repeat(np.arange(b)[:, None], r, 1), x1, y1, np.minimum(x1 + w, 999), np.minimum(y1 + h, 599)], 2)
    rois[:, r // 2:, 1:] = rois[:, : r - r // 2, 1:] + rs.normal(0, 4, (b, r - r // 2, 4))       # clustered copies
    rois = torch.from_numpy(rois.astype(np.float32))
    cls_prob = torch.from_numpy(rs.uniform(0, 1, (b * r, 2)).astype(np.float32))
    cls_prob[::7, 1] = 0.01                                                                      # below the threshold
    bbox_pred = torch.from_numpy((rs.standard_normal((b * r, 4)) * 1.5).astype(np.float32))
    im_info = torch.tensor([[600.0, 1000.0, scale]] * b)
    want = PO.detections(rois, cls_prob, bbox_pred, im_info, 0.05, 0.3)
    dets, counts = detections(rois.cuda(), cls_prob.cuda(), bbox_pred.cuda(), im_info.cuda(), 0.05, 0.3,
                              stds=(0.1, 0.1, 0.2, 0.2), means=(0.0, 0.0, 0.0, 0.0))
    dets, counts = dets.cpu(), counts.cpu()
    for i in range(b):
        m = int(counts[i])
        assert m == want[i].shape[0]
        assert torch.equal(dets[i, :m, 4], want[i][:, 4])                  # scores: bit-exact, same order
        assert (dets[i, :m, :4] - want[i][:, :4]).abs().max().item() <= 1e-3
        assert dets[i, m:].abs().sum().item() == 0.0
