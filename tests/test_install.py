"""Drop-in boundary against the real reference package (dev container only): install() rebinds exactly the hot-path
bodies, keeps classes / signatures / state-dict keys, and uninstall() restores the reference."""
import inspect

import pytest
import torch

from oracle import ref_loader

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")]


def test_install_and_uninstall_round_trip():
    ref_loader.load()
    import ultralytics.nn.modules.block as block
    import ultralytics.nn.modules.head as head
    import ultralytics.nn.tasks as tasks
    import ultralytics.utils.loss as uloss
    import ultralytics.utils.ops as uops

    import edge_yolo_b200.install as el
    from edge_yolo_b200 import modules as M
    from edge_yolo_b200 import nms as el_nms

    orig = {"dwt": block._PywtDWT2D.forward, "enh": block._WaveletEnhancer.forward, "att": block.LinearAttention.forward,
            "head": head.GFLHeadv2_uniH.forward, "nms": uops.non_max_suppression, "qfl": uloss.quality_focal_loss, "dfl": uloss.DFLoss.__call__}
    ref_sig = inspect.signature(uops.non_max_suppression)
    names = el.install()
    try:
        assert len(names) == 16
        assert tasks.v8DetectionLoss.__module__ == "edge_yolo_b200.detection_loss"
        assert block._WaveletEnhancer.forward is M.wavelet_enhancer_forward
        assert block.LinearAttention.forward is M.linear_attention_forward
        assert head.GFLHeadv2_uniH.forward is M.gfl_head_forward
        assert uops.non_max_suppression is el_nms.non_max_suppression
        assert list(inspect.signature(uops.non_max_suppression).parameters) == list(ref_sig.parameters)
        assert {k: v.default for k, v in inspect.signature(uops.non_max_suppression).parameters.items()} == \
               {k: v.default for k, v in ref_sig.parameters.items()}
        # class identities (what parse_model looks up by name, tasks.py:984) are untouched -> checkpoints still unpickle
        assert tasks.DSC3K2_Wavelet is block.DSC3K2_Wavelet and tasks.GFLHeadv2_uniH is head.GFLHeadv2_uniH
        # a reference-built model now carries our forwards and still has the reference's state-dict keys
        m = tasks.DetectionModel(ref_loader.REFERENCE_ROOT + "/ultralytics/cfg/models/11/yolo11n-test.yaml", ch=3, nc=80, verbose=False) \
            if not torch.cuda.is_available() else None
    except Exception as e:  # DetectionModel.__init__ runs a forward (stride probe): with our kernels that needs a GPU
        from edge_yolo_b200 import EdgelineError

        assert isinstance(e, EdgelineError), e
    finally:
        el.uninstall()
    assert block._PywtDWT2D.forward is orig["dwt"] and block._WaveletEnhancer.forward is orig["enh"]
    assert block.LinearAttention.forward is orig["att"] and head.GFLHeadv2_uniH.forward is orig["head"]
    assert uops.non_max_suppression is orig["nms"] and uloss.quality_focal_loss is orig["qfl"] and uloss.DFLoss.__call__ is orig["dfl"]


def test_loss_signatures_match_reference():
    ref_loader.load()
    import ultralytics.utils.loss as uloss

    from edge_yolo_b200 import loss as L

    for name in ("quality_focal_loss", "distribution_focal_loss"):
        a, b = inspect.signature(getattr(uloss, name)), inspect.signature(getattr(L, name))
        assert list(a.parameters) == list(b.parameters)
        assert [p.default for p in a.parameters.values()] == [p.default for p in b.parameters.values()]
    assert list(inspect.signature(uloss.DFLoss.__init__).parameters) == list(inspect.signature(L.DFLoss.__init__).parameters)


def test_module_constructor_signatures_match_reference():
    ref_loader.load()
    import ultralytics.nn.modules.block as block
    import ultralytics.nn.modules.conv as conv
    import ultralytics.nn.modules.head as head

    from edge_yolo_b200 import modules as M

    for ref_cls, mine in ((conv.WTConv2d, M.WTConv2d), (block.DSC3K2_Wavelet, M.DSC3K2_Wavelet), (block.C2PSA_LinearAttention, M.C2PSA_LinearAttention),
                          (block.LinearAttention, M.LinearAttention), (block._WaveletEnhancer, M._WaveletEnhancer),
                          (head.GFLHeadv2_uniH, M.GFLHeadv2_uniH), (block.DSBottleneck, M.DSBottleneck), (block.DSC3k, M.DSC3k)):
        a, b = inspect.signature(ref_cls.__init__), inspect.signature(mine.__init__)
        assert list(a.parameters) == list(b.parameters), ref_cls.__name__
        assert [p.default for p in a.parameters.values()] == [p.default for p in b.parameters.values()], ref_cls.__name__


def test_install_metrics_opt_in_round_trip():
    """metrics=True additionally rebinds box_iou and BaseValidator.match_predictions with the reference's signatures."""
    ref_loader.load()
    import ultralytics.engine.validator as validator
    import ultralytics.models.yolo.detect.val as dval
    import ultralytics.utils.metrics as umetrics

    import edge_yolo_b200.install as el
    from edge_yolo_b200 import metrics as el_metrics

    import ultralytics.utils.ops as uops

    o_iou, o_match, o_ap, o_scale = umetrics.box_iou, validator.BaseValidator.match_predictions, umetrics.ap_per_class, uops.scale_boxes
    assert list(inspect.signature(el_metrics.box_iou).parameters) == list(inspect.signature(o_iou).parameters)
    assert list(inspect.signature(el_metrics.scale_boxes).parameters) == list(inspect.signature(o_scale).parameters)
    assert list(inspect.signature(el_metrics.ap_per_class).parameters)[:10] == list(inspect.signature(o_ap).parameters)
    names = el.install(metrics=True)
    try:
        assert len(names) == 21
        assert umetrics.box_iou is el_metrics.box_iou and dval.box_iou is el_metrics.box_iou
        assert umetrics.ap_per_class is el._ap_per_class and uops.scale_boxes is el._scale_boxes
        # inputs outside the library's domain keep the reference's code: a CPU tensor goes through the reference's scale_boxes
        import torch

        cpu_boxes = torch.tensor([[10.0, 20.0, 700.0, 500.0]])
        want = o_scale((640, 640), cpu_boxes.clone(), (480, 640))
        assert torch.equal(uops.scale_boxes((640, 640), cpu_boxes.clone(), (480, 640)), want)
        assert list(inspect.signature(validator.BaseValidator.match_predictions).parameters) == list(inspect.signature(o_match).parameters)
        assert dval.DetectionValidator.match_predictions is validator.BaseValidator.match_predictions
    finally:
        el.uninstall()
    assert umetrics.box_iou is o_iou and validator.BaseValidator.match_predictions is o_match and dval.box_iou is o_iou
    assert umetrics.ap_per_class is o_ap and uops.scale_boxes is o_scale


def test_install_repairs_the_references_ihaar_dwt2d():
    """`IHaarDWT2D()` raises TypeError in the reference (a stray __init__(self, dim, num_heads) overrides the real one), which
    makes WaveletMixerMultiLevel / C3AW_MLM unbuildable.  install() binds the intended constructor and the synthesis forward onto the
    reference's class; uninstall() restores the (broken) original."""
    ref_loader.load()
    import ultralytics.nn.modules.block as block

    import edge_yolo_b200.install as el

    with pytest.raises(TypeError):
        block.IHaarDWT2D()
    el.install(nms=False, losses=False, criterion=False)
    try:
        m = block.WaveletMixerMultiLevel(8)
        assert type(m.idwt) is block.IHaarDWT2D and sorted(m.idwt.state_dict()) == ["hh", "hl", "lh", "ll", "recon_h", "recon_ll"]
        from edge_yolo_b200 import modules as M

        assert sorted(m.state_dict()) == sorted(M.WaveletMixerMultiLevel(8).state_dict())
    finally:
        el.uninstall()
    with pytest.raises(TypeError):
        block.IHaarDWT2D()
