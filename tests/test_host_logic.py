"""Host-side logic that needs no GPU: BatchNorm folding of `EdgeLineYOLO.fuse` (the counterpart of BaseModel.fuse, nn/tasks.py:214-242,
plus DSConv's BatchNorm) and the K-chunk table of the concat-free 1x1 convolution."""
import pytest
import torch

from edge_yolo_b200 import EdgelineError, ops
from edge_yolo_b200.model import EdgeLineYOLO
from edge_yolo_b200.modules import Conv, DSConv


def _randomise_bn(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.3)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.2)


@torch.no_grad()
def test_fuse_folds_batchnorm_exactly_for_conv_and_dsconv():
    """Every Conv / DWConv / DSConv of the n graph gives the same output (fp32, eval) before and after `fuse(dsconv=True)`; the folded
    modules keep no BatchNorm and the state-dict keys of the convolutions survive (checkpoint compatibility)."""
    model = EdgeLineYOLO("n", 80).float().eval()
    _randomise_bn(model)
    g = torch.Generator().manual_seed(1)
    probes = []
    for name, m in model.named_modules():
        if isinstance(m, (Conv, DSConv)):
            c_in = m.conv.in_channels if isinstance(m, Conv) else m.dw.in_channels
            x = torch.randn(2, c_in, 12, 12, generator=g)
            probes.append((name, m, x, m(x)))
    assert len(probes) > 60 and any(isinstance(m, DSConv) for _, m, _, _ in probes)
    model.fuse(dsconv=True)
    for name, m, x, want in probes:
        got = m(x)
        assert not isinstance(getattr(m, "bn", None), torch.nn.BatchNorm2d), name
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5, msg=lambda s, n=name: f"{n}: {s}")
    keys = model.state_dict().keys()
    assert "model.0.conv.weight" in keys and "model.0.conv.bias" in keys and not any(".bn." in k for k in keys if k.startswith("model.0."))


def test_pw_chunks_cover_the_concatenated_k_in_order():
    """`_pw_chunks`: per source 64-channel TMA boxes (16 / 32 for narrow sources), consecutive in the concatenated K, no box crossing a
    source boundary -- what lets el_pwconv_fwd read `torch.cat([a, b, m1, ...], 1)` without materialising it (block.py:3783-3788)."""
    for src_c in ([64], [16], [32, 32], [24, 64, 8], [128, 128, 64, 64], [40, 200]):
        chunks = ops._pw_chunks(src_c)
        k = 0
        bounds = [sum(src_c[: i + 1]) for i in range(len(src_c))]
        for k0, creal, bw in chunks:
            assert k0 == k and 0 < creal <= bw and bw in (16, 32, 64)
            assert not any(k0 < b < k0 + creal for b in bounds)            # stays inside one source
            k += creal
        assert k == sum(src_c)
    with pytest.raises(EdgelineError):
        ops._pw_chunks([20])                                               # sources are multiples of 8 channels (16-byte vectors)


def test_non_max_suppression_host_branches():
    """The branches of `non_max_suppression` decided on the host, as in utils/ops.py:215-232: threshold asserts (same messages), tuple
    inputs, the end-to-end (B, N, 6) early return with its `classes` filter, and the hard errors outside the detection path."""
    from edge_yolo_b200.nms import non_max_suppression

    with pytest.raises(AssertionError, match="Invalid Confidence threshold 1.5"):
        non_max_suppression(torch.zeros(1, 6, 4), conf_thres=1.5)
    with pytest.raises(AssertionError, match="Invalid IoU -0.1"):
        non_max_suppression(torch.zeros(1, 6, 4), iou_thres=-0.1)
    g = torch.Generator().manual_seed(0)
    e2e = torch.rand(2, 50, 6, generator=g)
    e2e[..., 5] = torch.randint(0, 4, (2, 50), generator=g).float()
    out = non_max_suppression((e2e, None), conf_thres=0.4, max_det=7)      # tuple: (inference output, loss output)
    for o, p in zip(out, e2e):
        assert torch.equal(o, p[p[:, 4] > 0.4][:7])
    out = non_max_suppression(e2e, conf_thres=0.4, max_det=50, classes=[1, 3])
    for o, p in zip(out, e2e):
        keep = p[p[:, 4] > 0.4]
        assert torch.equal(o, keep[(keep[:, 5] == 1) | (keep[:, 5] == 3)])
    dense = torch.rand(1, 4 + 3, 20, generator=g)
    with pytest.raises(NotImplementedError):
        non_max_suppression(dense, rotated=True)
    with pytest.raises(NotImplementedError):
        non_max_suppression(dense, labels=[torch.zeros(1, 5)])
    with pytest.raises(NotImplementedError):
        non_max_suppression(dense, nc=2)                                    # 1 mask channel
    with pytest.raises(EdgelineError):
        non_max_suppression(dense)                                          # CPU tensor: no fallback


def test_engine_fused_model_refuses_load_state_dict():
    """the engine keeps folded biases / packed filters outside the state dict; loading weights after fuse(engine=True) must fail
    loudly instead of pairing new weights with stale side tensors.  Loading BEFORE fusing is the supported order."""
    model = EdgeLineYOLO("n", 8).float().eval()
    state = {k: v.clone() for k, v in model.state_dict().items()}
    model.load_state_dict(state)  # fine: not fused yet
    try:
        model.fuse(engine=True)
    except EdgelineError:
        pytest.skip("engine fuse needs the CUDA library's weight packing")
    with pytest.raises(RuntimeError, match="before fusing"):
        model.load_state_dict(state, strict=False)


@torch.no_grad()
def test_c3_cv12_stacks_the_two_entry_convs_of_dsc3k():
    """`_c3_cv12` (engine: cv1 and cv2 of DSC3k read the same input, C3.forward block.py:394-396): the merged conv's output channels are
    cv1's then cv2's, its bias the two folded-BatchNorm biases in the same order, so a split store at cv1's width reproduces both convs;
    the merged conv is not a registered submodule (state-dict keys unchanged) and is rebuilt when a weight changes."""
    from edge_yolo_b200 import modules as M

    model = EdgeLineYOLO("n", 80).float().eval()
    _randomise_bn(model, seed=3)
    keys_before = set(model.state_dict().keys())
    model.fuse(engine=True)
    blocks = [m for m in model.modules() if isinstance(m, M.DSC3k)]
    assert blocks
    g = torch.Generator().manual_seed(2)
    for blk in blocks:
        x = torch.randn(2, blk.cv1.conv.in_channels, 6, 5, generator=g)
        conv, bias, act = M._c3_cv12(blk, x)
        c1 = blk.cv1.conv.out_channels
        assert conv.out_channels == c1 + blk.cv2.conv.out_channels and bias.numel() == conv.out_channels and act == blk.cv1.el_act
        y = torch.nn.functional.conv2d(x, conv.weight) + bias.view(1, -1, 1, 1)
        want1 = torch.nn.functional.conv2d(x, blk.cv1.conv.weight) + blk.cv1.el_bias.view(1, -1, 1, 1)
        want2 = torch.nn.functional.conv2d(x, blk.cv2.conv.weight) + blk.cv2.el_bias.view(1, -1, 1, 1)
        torch.testing.assert_close(y[:, :c1], want1)
        torch.testing.assert_close(y[:, c1:], want2)
        assert M._c3_cv12(blk, x)[0] is conv                      # cached
        blk.cv2.conv.weight.mul_(2.0)                             # in-place update -> new version -> repacked
        conv2, _, _ = M._c3_cv12(blk, x)
        assert conv2 is not conv
        torch.testing.assert_close(conv2.weight[c1:], blk.cv2.conv.weight)
    assert {k for k in model.state_dict().keys() if "el_cv12" in k} == set() and keys_before  # nothing new registered


def test_dsconv3_weight_packing_is_the_single_tile_pwconv_packing():
    """`pack_dsconv3_weight` = `pack_pw_weight` for one source of C channels with ONE output-channel tile of ceil16(N) rows (what
    el_dsconv3_fwd keeps resident): per 64-channel chunk a tile of ceil16(N) x 128 bytes padded to 1 KiB; 16 / 32-channel inputs one narrow tile."""
    for C, N in ((64, 64), (80, 80), (256, 80), (32, 32), (16, 8), (72, 40)):
        w = torch.randn(N, C)
        wpk = ops.pack_dsconv3_weight(w, torch.bfloat16)
        n_pad = -(-N // 16) * 16
        assert torch.equal(wpk, ops.pack_pw_weight(w, [C], torch.bfloat16, n_tile=n_pad))
        rb = 32 if C <= 16 else (64 if C <= 32 else 128)
        chunks = -(-C // (rb // 2))
        tile = -(-(n_pad * rb) // 1024) * 1024
        assert wpk.numel() * 2 == chunks * tile
