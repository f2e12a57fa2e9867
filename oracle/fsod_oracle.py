"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Restatement of the attention-RPN feature of the sibling model FSOD
(lib/model/framework/fsod.py:90-112): shot-mean of the positive support features, AvgPool2d(14, stride 1), depth-wise
cross-correlation of the query feature with the pooled 7x7 kernel.  Pinned to the UNMODIFIED reference FSOD module by
oracle/make_golden_fsod.py -> tests/golden/fsod_attention.npz (tests/test_oracle_pins.py)."""
import torch
import torch.nn.functional as F

import dana_oracle as O


def fsod_attention_feature(p, im_data, support_ims, n_shot, num_layers=50):
    """im_data [B,3,H,W], support_ims [B,K,3,Hs,Ws] (positive set) -> correlation_feat [B,1024,h-6,w-6]."""
    base = O.rcnn_base(im_data, p, num_layers)                                            # fsod.py:90
    sup = O.rcnn_base(support_ims.reshape(-1, *support_ims.shape[2:]), p, num_layers)    # :100-101
    sup = sup.view(-1, n_shot, *sup.shape[1:])                                            # :102
    pos = sup[:, :n_shot].mean(1)                                                         # :103
    pos = F.avg_pool2d(pos, 14, 1)                                                        # :104  [B,1024,7,7]
    maps = []
    for kernel, feat in zip(pos.chunk(pos.shape[0], 0), base.chunk(base.shape[0], 0)):    # :107-111
        kernel = kernel.view(pos.shape[1], 1, pos.shape[2], pos.shape[3])
        maps.append(F.conv2d(feat, kernel, groups=pos.shape[1]).squeeze(0))
    return torch.stack(maps, 0)                                                           # :112
