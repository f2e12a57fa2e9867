"""CUDA-graph replay of the eval forward must reproduce the eager launch sequence bit for bit."""
import os

import pytest
import torch

import dana_oracle as O
import make_golden as MG

pytestmark = pytest.mark.gpu


def test_graph_replay_equals_eager():
    import dana_b200  # noqa: F401
    from dana_b200.config import cfg_from_file, cfg_from_list, reset_cfg
    from dana_b200.dana import DAnARCNN
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    reset_cfg()
    cfg_from_file(os.path.join(root, "cfgs", "res50.yml"))
    cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]"])
    fc = MG.FORWARD_CASE
    sd = O.make_params(fc["seed"], attn_std=fc["attn_std"])
    outs = []
    for graph in (False, True):
        net = DAnARCNN(["bg", "fg"], "concat", 256, 256, semantic_enhance=True, num_way=2, num_shot=fc["n_shot"],
                       use_cuda_graph=graph)
        net.create_architecture()
        net.load_state_dict(sd, strict=False)
        net.cuda().eval()
        res = []
        for seed in (fc["seed"], fc["seed"] + 1, fc["seed"]):          # replay with changing inputs
            im, info, sup = O.synth_inputs(seed, 1, fc["height"], fc["width"], fc["n_shot"])
            out = net(im.cuda(), info.cuda(), torch.zeros(1, 1, 5).cuda(), torch.zeros(1).cuda(), sup.cuda())
            res.append([t.clone() for t in out[:3]])
        if graph:
            assert net._graphs and all(bool(g) for g in net._graphs.values()), "graph capture fell back to eager"
        outs.append(res)
    for a, b in zip(*outs):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    # same input again gives the same result as the first call (static buffers are refreshed per call)
    for x, y in zip(outs[1][0], outs[1][2]):
        assert torch.equal(x, y)


def test_prefetcher_and_result_fetcher_pipeline():
    """The e2e loop of bench.py in miniature: inputs prefetched H2D on a side stream, results fetched D2H one step late;
    every step's host results must equal a blocking read of the same step."""
    import dana_b200  # noqa: F401
    from dana_b200.pipeline import EpisodePrefetcher, ResultFetcher
    dev = torch.device("cuda")
    batches = [(torch.full((64, 33), float(i)).pin_memory(), torch.arange(7, dtype=torch.float32).pin_memory() + i)
               for i in range(6)]
    pf, fetch = EpisodePrefetcher(dev), ResultFetcher(dev)
    got, pending = [], None
    for a, b in pf.run(iter(batches)):
        out = (a * 2 + 1, b.sum().reshape(1))
        ticket = fetch.start(out)
        if pending is not None:
            got.append([t.clone() for t in fetch.wait(pending)])
        pending = ticket
    got.append([t.clone() for t in fetch.wait(pending)])
    assert len(got) == 6
    for i, (x, y) in enumerate(got):
        assert torch.equal(x, torch.full((64, 33), 2.0 * i + 1))
        assert float(y) == float(sum(range(7)) + 7 * i)
