"""CPU tests of the host-side mirror of the reference interfaces (config surface, anchors, module
state-dict contract, episode sharding incl. a world_size-2 gloo run)."""
import os
import sys

import numpy as np
import pytest
import torch

import dana_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def cfgmod():
    import dana_b200  # noqa: F401
    from dana_b200 import config
    config.reset_cfg()
    yield config
    config.reset_cfg()


def test_cfg_defaults_and_yaml_merge(cfgmod):
    cfg = cfgmod.cfg
    assert cfg.POOLING_MODE == "crop" and cfg.TEST.RPN_POST_NMS_TOP_N == 300 and cfg.TRAIN.RPN_PRE_NMS_TOP_N == 12000
    cfgmod.cfg_from_file(os.path.join(ROOT, "cfgs", "res50.yml"))
    assert cfg.POOLING_MODE == "align" and cfg.TRAIN.BATCH_SIZE == 128 and cfg.TRAIN.BG_THRESH_LO == 0.0
    cfgmod.cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]", "MAX_NUM_GT_BOXES", "50"])
    assert cfg.ANCHOR_SCALES == [4, 8, 16, 32] and cfg.MAX_NUM_GT_BOXES == 50
    cfgmod.cfg_from_list(["TEST.RPN_POST_NMS_TOP_N", "1000"])
    assert cfg.TEST.RPN_POST_NMS_TOP_N == 1000


def test_cfg_strict_keys_and_types(cfgmod, tmp_path):
    bad = tmp_path / "bad.yml"
    bad.write_text("NOT_A_KEY: 1\n")
    with pytest.raises(KeyError):
        cfgmod.cfg_from_file(str(bad))
    bad.write_text("POOLING_SIZE: seven\n")
    with pytest.raises(ValueError):
        cfgmod.cfg_from_file(str(bad))
    with pytest.raises(AssertionError):
        cfgmod.cfg_from_list(["POOLING_SIZE", "7.5"])
    with pytest.raises(AssertionError):
        cfgmod.cfg_from_list(["POOLING_SIZE"])
    cfgmod.cfg_from_file(os.path.join(ROOT, "cfgs", "res101_ls.yml"))
    assert cfgmod.cfg.TEST.SCALES == (800,) and cfgmod.cfg.TEST.RPN_POST_NMS_TOP_N == 1000


def test_anchors_match_oracle():
    import dana_b200  # noqa: F401
    from dana_b200.anchors import generate_anchors
    np.testing.assert_array_equal(generate_anchors(), O.generate_anchors())
    np.testing.assert_array_equal(generate_anchors(scales=(4, 8, 16, 32)), O.generate_anchors(scales=(4, 8, 16, 32)))
    # 12 anchors of the shipped CLI default, SURVEY.md section 8a
    assert generate_anchors(scales=(4, 8, 16, 32))[0].tolist() == [-38, -16, 53, 31]
    assert generate_anchors(scales=(4, 8, 16, 32))[11].tolist() == [-168, -344, 183, 359]


def test_positional_encoding_matches_oracle():
    import dana_b200  # noqa: F401
    from dana_b200.engine import DanaEngine, positional_encoding
    for n in (49, 196, 400):
        assert torch.equal(positional_encoding(n), O.positional_encoding(n))
    assert DanaEngine._trunk_hw(600, 1000) == (38, 63)
    assert DanaEngine._trunk_hw(600, 800) == (38, 50)
    assert DanaEngine._trunk_hw(800, 1333) == (50, 84)
    assert DanaEngine._trunk_hw(320, 320) == (20, 20)


def test_module_state_dict_contract(cfgmod):
    """346 keys with the reference's names and shapes; trainable / frozen split of dana.py:350-368."""
    cfgmod.cfg_from_file(os.path.join(ROOT, "cfgs", "res50.yml"))
    cfgmod.cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]"])
    from dana_b200.dana import DAnARCNN
    net = DAnARCNN(["bg", "fg"], "concat", 256, 256, pretrained=False, semantic_enhance=True, num_way=2, num_shot=3)
    net.create_architecture()
    sd = net.state_dict()
    assert len(sd) == 346
    shapes = {k: tuple(v.shape) for k, v in sd.items() if not k.endswith("num_batches_tracked")}
    assert shapes == {k: tuple(v) for k, v in O.param_shapes().items()}
    n_all = sum(p.numel() for p in net.parameters())
    n_train = sum(p.numel() for p in net.parameters() if p.requires_grad)
    assert round(n_all / 1e6, 2) == 37.39 and round(n_train / 1e6, 2) == 37.11
    net.train()
    assert not net.RCNN_base[1].training and not net.RCNN_base[4].training and net.RCNN_base[6].training
    assert all(not m.training for m in net.RCNN_top.modules() if isinstance(m, torch.nn.BatchNorm2d))
    with pytest.raises(RuntimeError):  # CUDA only, loudly
        net.eval()(torch.zeros(1, 3, 64, 64), torch.zeros(1, 3), torch.zeros(1, 1, 5), torch.zeros(1),
                   torch.zeros(1, 3, 3, 320, 320))


def test_shard_episodes():
    import dana_b200  # noqa: F401
    from dana_b200.sharding import shard_range
    for n, w in ((32, 8), (10, 4), (3, 8), (0, 2)):
        got = [shard_range(n, r, w) for r in range(w)]
        flat = [i for a, b in got for i in range(a, b)]
        assert flat == list(range(n))
        sizes = [b - a for a, b in got]
        assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import dana_b200  # noqa: F401
    from dana_b200.sharding import aggregate_throughput, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(10, rank, world)
    # rank 1 is "slower": whole-job throughput must use the max time and the summed units
    value, t_max, units = aggregate_throughput(units=b - a, seconds=1.0 + rank, device="cpu")
    if rank == 0:
        out.put((value, t_max, units))
    dist.destroy_process_group()


def test_aggregate_throughput_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    value, t_max, units = q.get(timeout=10)
    assert units == 10 and t_max == pytest.approx(2.0) and value == pytest.approx(5.0)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with the contract's
    keys, on this machine's host cores, without a GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "query-images/sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_episode_host_geometry_matches_oracle():
    """The host-side arithmetic of dana_b200.episode (output sizes, short-side truncation, cvRound) against the oracle's
    restatement of the loaders, without a GPU (the kernel itself is checked in test_gpu_episode.py)."""
    import dana_b200  # noqa: F401
    import episode_oracle as E
    from dana_b200 import episode
    rs = np.random.RandomState(0)
    for _ in range(200):
        bh, bw = int(rs.randint(1, 700)), int(rs.randint(1, 700))
        assert episode._fit(bh, bw, 320) == E._fit(bh, bw, 320)
    for v in (0.5, 1.5, 2.5, 959.5, 960.4999, 1000.0):
        assert episode.cv_round(v) == E.cv_round(v) == int(np.rint(v))
    # prep_im_for_blob output size: cvRound(side * 600 / min side) -- e.g. 375x500 -> 600x800, 333x500 -> 600x901
    for h, w, want in [(375, 500, (600, 800)), (333, 500, (600, 901)), (480, 640, (600, 800)), (600, 1000, (600, 1000))]:
        s = 600.0 / min(h, w)
        assert (episode.cv_round(h * s), episode.cv_round(w * s)) == want
    with pytest.raises(RuntimeError):
        episode.prep_im_for_blob(torch.zeros((8, 8, 3), dtype=torch.uint8), [0, 0, 0], 8)    # CPU tensor: no fallback


def _grad_sync_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import dana_b200  # noqa: F401
    from dana_b200.grad_sync import BucketedGradAllReduce
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                        # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.ReLU(), torch.nn.Linear(64, 19), torch.nn.ReLU(),
                              torch.nn.Linear(19, 3))
    net[2].bias.requires_grad_(False)                            # a frozen parameter (the reference freezes BN / conv1)
    unused = torch.nn.Parameter(torch.ones(5))                   # a parameter that receives no gradient this step
    params = list(net.parameters()) + [unused]
    g = torch.Generator().manual_seed(100 + rank)                # different data per rank
    x, y = torch.randn(8, 37, generator=g), torch.randn(8, 3, generator=g)
    sync = BucketedGradAllReduce(params, bucket_bytes=200)       # several small buckets
    results = {}
    for mode in ("overlap", "after"):
        for p in params:
            p.grad = None
        if mode == "overlap":
            sync.arm()
        loss = ((net(x) - y) ** 2).mean()
        loss.backward()
        local = [None if p.grad is None else p.grad.clone() for p in params]
        if mode == "overlap":
            sync.finish()
        else:
            sync.reduce_now()
        results[mode] = ([None if p.grad is None else p.grad.clone() for p in params], local)
    # reference result: plain all_reduce of every local gradient, averaged
    want = []
    for p, lg in zip(params, results["after"][1]):
        if not p.requires_grad:
            want.append(None)
            continue
        t = lg.clone() if lg is not None else torch.zeros_like(p)
        dist.all_reduce(t)
        want.append(t / world)
    ok = len(sync.buckets) >= 3
    for mode in ("overlap", "after"):
        for got, w, p in zip(results[mode][0], want, params):
            if w is None:
                ok = ok and (got is None)
            else:
                ok = ok and got is not None and torch.allclose(got, w, rtol=0, atol=1e-7)
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_bucketed_grad_allreduce_gloo_world2():
    """The training step's only collective: bucketed, hook-driven gradient averaging equals a plain
    per-parameter all-reduce mean, with frozen and unused parameters, in both the overlapped and the after-backward mode."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_grad_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def _arena_sync_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import dana_b200  # noqa: F401
    from dana_b200.train_step import ArenaGradAllReduce, ParamArena
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                        # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.ReLU(), torch.nn.Linear(64, 19), torch.nn.ReLU(),
                              torch.nn.Linear(19, 3))
    before = [p.detach().clone() for p in net.parameters()]
    named = [(n, p) for n, p in net.named_parameters()]
    named.reverse()
    arenas = [ParamArena([(n, p) for n, p in named if "bias" not in n], 0.1, 0.0, 200),
              ParamArena([(n, p) for n, p in named if "bias" in n], 0.2, 0.0, 64)]
    ok = all(torch.equal(a, b) for a, b in zip(before, net.parameters()))          # values moved into the arena intact
    ok = ok and len(arenas[0].buckets) >= 2 and all(o % 4 == 0 for a in arenas for o in a.offsets)
    sync = ArenaGradAllReduce(arenas)
    g = torch.Generator().manual_seed(100 + rank)                # different data per rank
    x, y = torch.randn(8, 37, generator=g), torch.randn(8, 3, generator=g)
    for step in range(2):                                        # twice: arm / finish are per step, grads re-zeroed
        for a in arenas:
            a.rebind_grads()
            a.grad.zero_()
        sync.arm()
        ((net(x) - y) ** 2).mean().backward()
        sync.finish()
        # every parameter's .grad is still an arena slice and holds the SUM over ranks (1/world goes into the SGD kernel)
        ref = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.ReLU(), torch.nn.Linear(64, 19), torch.nn.ReLU(),
                                  torch.nn.Linear(19, 3))
        ref.load_state_dict(net.state_dict())
        ((ref(x) - y) ** 2).mean().backward()
        for p, q in zip(net.parameters(), ref.parameters()):
            t = q.grad.clone()
            dist.all_reduce(t)
            ok = ok and torch.allclose(p.grad, t, rtol=0, atol=1e-6)
            ok = ok and any(a.grad.data_ptr() <= p.grad.data_ptr() < a.grad.data_ptr() + 4 * a.total for a in arenas)
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_arena_grad_allreduce_gloo_world2():
    """The training step's collective as shipped (train_step.py): gradients accumulate straight into flat arenas, whose
    contiguous bucket ranges are all-reduced from post-accumulate hooks in bucket order -- equals a per-parameter
    all-reduce sum; parameter values survive the move into the arena."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + os.getpid() % 1000
    procs = [ctx.Process(target=_arena_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


# ------------------------------------------------------------------ training target layers (host side, row a15)
@pytest.mark.parametrize("seed,n_gt,hw", [(0, 1, (10, 14)), (1, 3, (38, 63)), (2, 8, (25, 40)), (3, 0, (12, 12))])
def test_target_layers_match_oracle(seed, n_gt, hw):
    """dana_b200.targets (numpy, product path) against oracle/train_oracle.py (torch restatement pinned to the
    unmodified reference): under the same numpy seed both make the same RNG draws -> identical labels / sampled
    rois, targets and weights equal to fp32 rounding."""
    import numpy as np
    import torch

    import dana_oracle as O
    import train_oracle as T
    from dana_b200 import targets as TG
    rs = np.random.RandomState(seed)
    fh, fw = hw
    b = 2
    gt = np.zeros((b, 6 if n_gt <= 6 else 10, 5), dtype=np.float32)
    for i in range(b):
        for j in range(n_gt):
            x1, y1 = rs.uniform(0, fw * 16 - 80), rs.uniform(0, fh * 16 - 80)
            gt[i, j] = [x1, y1, x1 + rs.uniform(20, 78), y1 + rs.uniform(20, 78), 1.0]
    info = np.array([[fh * 16.0, fw * 16.0, 1.0]] * b, dtype=np.float32)
    base = O.generate_anchors(scales=(4, 8, 16, 32)).astype(np.float32)
    np.random.seed(100 + seed)
    want = T.anchor_target_layer(fh, fw, torch.from_numpy(gt), torch.from_numpy(info), torch.from_numpy(base))
    np.random.seed(100 + seed)
    labels, tgt, in_w, out_w = TG.anchor_targets(fh, fw, gt, info, base)
    a = base.shape[0]
    # the oracle returns the reference's NCHW views; bring them to (y, x, a) order
    w_lab = want[0].view(b, a, fh, fw).permute(0, 2, 3, 1).reshape(b, -1).numpy()
    w_tgt = want[1].view(b, a, 4, fh, fw).permute(0, 3, 4, 1, 2).reshape(b, -1, 4).numpy()
    w_in = want[2].view(b, a, 4, fh, fw).permute(0, 3, 4, 1, 2).reshape(b, -1, 4).numpy()[..., 0]
    w_out = want[3].view(b, a, 4, fh, fw).permute(0, 3, 4, 1, 2).reshape(b, -1, 4).numpy()[..., 0]
    np.testing.assert_array_equal(labels, w_lab.astype(np.int8))
    np.testing.assert_allclose(tgt, w_tgt, rtol=0, atol=2e-6)
    np.testing.assert_array_equal(in_w, w_in)
    np.testing.assert_allclose(out_w, w_out, rtol=1e-7, atol=0)
    if n_gt == 0:
        return                                   # no box at all: the proposal target layer raises in both
    rois = np.zeros((b, 300, 5), dtype=np.float32)
    for i in range(b):
        x1, y1 = rs.uniform(0, fw * 16 - 40, 300), rs.uniform(0, fh * 16 - 40, 300)
        rois[i] = np.stack([np.full(300, i), x1, y1, x1 + rs.uniform(8, 200, 300), y1 + rs.uniform(8, 200, 300)], 1)
        rois[i, :40, 1:] = gt[i, rs.randint(0, n_gt, 40), :4] + rs.normal(0, 4, (40, 4))      # some near the gt
    np.random.seed(200 + seed)
    w = T.proposal_target_layer(torch.from_numpy(rois), torch.from_numpy(gt))
    np.random.seed(200 + seed)
    g = TG.proposal_targets(rois, gt)
    np.testing.assert_array_equal(g[0], w[0].numpy())
    np.testing.assert_array_equal(g[1], w[1].numpy())
    np.testing.assert_allclose(g[2], w[2].numpy(), rtol=0, atol=2e-5)
    np.testing.assert_array_equal(g[3], w[3].numpy())
    np.testing.assert_array_equal(g[4], w[4].numpy())


def test_pretrained_trunk_load(tmp_path):
    """pretrained=True (dana.py:337-341): the caffe-converted resnet checkpoint (torchvision-style keys conv1 / bn1 /
    layer1..4 / fc) is loaded into RCNN_base / RCNN_top; fc.* is ignored; a checkpoint lacking a trunk tensor is an error."""
    import torch

    import dana_b200  # noqa: F401
    from dana_b200.config import cfg_from_file, reset_cfg
    from dana_b200.dana import DAnARCNN
    reset_cfg()
    cfg_from_file(os.path.join(ROOT, "cfgs", "res50.yml"))
    ref = DAnARCNN(["bg", "fg"], "concat", 256, 256, pretrained=False, semantic_enhance=True)
    ref.create_architecture()
    key_map = (("RCNN_base.0.", "conv1."), ("RCNN_base.1.", "bn1."), ("RCNN_base.4.", "layer1."),
               ("RCNN_base.5.", "layer2."), ("RCNN_base.6.", "layer3."), ("RCNN_top.0.", "layer4."))
    g = torch.Generator().manual_seed(3)
    ckpt = {}
    for k, v in ref.state_dict().items():
        for src, dst in key_map:
            if k.startswith(src) and not k.endswith("num_batches_tracked"):
                ckpt[dst + k[len(src):]] = torch.randn(v.shape, generator=g)
    ckpt["fc.weight"] = torch.zeros(1000, 2048)
    ckpt["fc.bias"] = torch.zeros(1000)
    path = str(tmp_path / "resnet50_caffe.pth")
    torch.save(ckpt, path)
    net = DAnARCNN(["bg", "fg"], "concat", 256, 256, pretrained=True, semantic_enhance=True)
    net.model_path = path
    net.create_architecture()
    sd = net.state_dict()
    assert torch.equal(sd["RCNN_base.0.weight"], ckpt["conv1.weight"])
    assert torch.equal(sd["RCNN_base.6.5.bn3.running_var"], ckpt["layer3.5.bn3.running_var"])
    assert torch.equal(sd["RCNN_top.0.2.conv3.weight"], ckpt["layer4.2.conv3.weight"])
    assert not net.RCNN_base[0].weight.requires_grad and not net.RCNN_base[4][0].conv1.weight.requires_grad
    del ckpt["layer2.0.conv1.weight"]
    torch.save(ckpt, path)
    bad = DAnARCNN(["bg", "fg"], "concat", 256, 256, pretrained=True)
    bad.model_path = path
    with pytest.raises(RuntimeError):
        bad.create_architecture()


# ------------------------------------------------------------------ training losses in their static-shape form (row a15)
def test_static_shape_losses_match_oracle():
    """train_model.TrainGraph's loss expressions -- written with 0 / 1 weights and device-side ranks instead of
    index_select, so that the step can be captured into a CUDA graph -- against the oracle's index-based restatement
    of rpn.py:96-116 and dana.py:199-215 (values AND gradients), including no-foreground, few-background and
    more-foreground-than-quota cases."""
    import train_oracle as T
    from dana_b200.train_model import TrainGraph
    rs = np.random.RandomState(11)
    tg = TrainGraph({}, num_layers=50)
    # RPN losses: NHWC raw head output [B,H,W,6A] vs the oracle's NCHW layouts
    b, h, w, a = 2, 7, 9, 12
    raw = torch.from_numpy(rs.standard_normal((b, h, w, 6 * a)).astype(np.float32)).requires_grad_(True)
    labels = torch.from_numpy(rs.choice([-1, 0, 1], size=(b, h * w * a), p=[0.7, 0.2, 0.1]).astype(np.int8))
    tgt = torch.from_numpy(rs.standard_normal((b, h * w * a, 4)).astype(np.float32)) * 0.3
    in_w = (labels == 1).float()
    out_w = (labels >= 0).float() / 41.0
    got = tg.rpn_losses(raw, (labels, tgt, in_w, out_w), a)
    (got[0] + got[1]).backward()
    g_got = raw.grad.clone()
    raw2 = raw.detach().clone().requires_grad_(True)
    nchw = raw2.permute(0, 3, 1, 2)
    o_lab = labels.view(b, h, w, a).permute(0, 3, 1, 2).reshape(b, 1, a * h, w).float()
    to4 = lambda t: t.view(b, h, w, a, 1).expand(b, h, w, a, 4).reshape(b, h, w, 4 * a).permute(0, 3, 1, 2)  # noqa: E731
    want = T.rpn_losses(nchw[:, :2 * a].contiguous(), nchw[:, 2 * a:].contiguous(),
                        (o_lab, tgt.view(b, h, w, 4 * a).permute(0, 3, 1, 2), to4(in_w), to4(out_w)), a)
    (want[0] + want[1]).backward()
    assert abs(float(got[0].detach()) - float(want[0].detach())) <= 1e-6 * abs(float(want[0].detach()))
    assert abs(float(got[1].detach()) - float(want[1].detach())) <= 1e-6 * abs(float(want[1].detach()))
    assert torch.allclose(g_got, raw2.grad, rtol=1e-5, atol=1e-8)
    # R-CNN classification loss with hard-negative mining
    for r, n_fg in ((128, 20), (128, 0), (64, 50), (256, 3), (8, 1)):
        scores = torch.from_numpy(rs.standard_normal((2 * r, 2)).astype(np.float32)).requires_grad_(True)
        lab = torch.zeros(r)
        lab[torch.from_numpy(rs.permutation(r)[:n_fg].copy())] = 1
        lab_all = torch.cat([lab, torch.zeros(r)]).long()
        got = TrainGraph.rcnn_cls_loss(scores, lab_all)
        got.backward()
        s2 = scores.detach().clone().requires_grad_(True)
        want = T.rcnn_cls_loss(s2, lab_all)
        want.backward()
        assert abs(float(got.detach()) - float(want.detach())) <= 1e-6 * abs(float(want.detach())), (r, n_fg)
        assert torch.allclose(scores.grad, s2.grad, rtol=1e-5, atol=1e-8), (r, n_fg)


@pytest.mark.parametrize("bucket_bytes", [1, 200, 4096, 1 << 30])
def test_param_arena_layout(bucket_bytes):
    """ParamArena (train_step.py): every parameter and gradient becomes a 16-byte aligned slice of one flat buffer with
    its values intact; the buckets are contiguous ranges of whole tensors that tile the arena exactly once."""
    from dana_b200.train_step import ParamArena
    torch.manual_seed(1)
    shapes = [(7, 3), (5,), (64, 9), (1,), (33, 2, 3, 3), (4,)]
    params = [("p%d" % i, torch.nn.Parameter(torch.randn(*s))) for i, s in enumerate(shapes)]
    before = [p.detach().clone() for _, p in params]
    arena = ParamArena(params, 0.1, 0.0, bucket_bytes)
    covered = 0
    for (i0, i1, begin, end) in arena.buckets:
        assert i1 > i0 and begin == covered and begin % 4 == 0 and end % 4 == 0
        assert begin == arena.offsets[i0]
        last = arena.offsets[i1 - 1] + (params[i1 - 1][1].numel() + 3) // 4 * 4
        assert end == last
        covered = end
    assert covered == arena.total and sum(i1 - i0 for i0, i1, _, _ in arena.buckets) == len(params)
    for (_, p), b, off in zip(params, before, arena.offsets):
        assert torch.equal(p.detach(), b) and off % 4 == 0
        assert p.data_ptr() == arena.param.data_ptr() + 4 * off and p.grad.data_ptr() == arena.grad.data_ptr() + 4 * off
    # a step's worth of autograd accumulates in place, and rebind_grads repairs a dropped .grad
    loss = sum((p * p).sum() for _, p in params)
    loss.backward()
    for (_, p), off in zip(params, arena.offsets):
        assert torch.allclose(arena.grad[off:off + p.numel()].view(p.shape), 2 * p.detach())
    params[2][1].grad = None
    arena.rebind_grads()
    assert params[2][1].grad.data_ptr() == arena.grad.data_ptr() + 4 * arena.offsets[2]
