"""Drop-in boundary on the GPU: `edge_yolo_b200.install()` behind the UNMODIFIED reference's own public API
(`YOLO(cfg, task="detect").predict / val / train`, engine/model.py:501, 609, 742; predictor engine/predictor.py:220;
models/yolo/detect/predict.py:23-41, val.py:92; nn/tasks.py:301-313) compared with the UNINSTALLED reference on the same GPU with the
same weights.  The reference package comes from baseline/_ref (pip-installed by __graft_entry__.build(), shipped with the gpurun
snapshot).  Contract (north_star): fp32 outputs within 1e-5 relative, NMS rows bit-exact on identical inputs, mAP within 0.1 points.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="reference package not found (baseline/_ref is made by __graft_entry__.build())")]

S = 256


@pytest.fixture(scope="module")
def api(tmp_path_factory):
    import _ref_api as R

    tmp = str(tmp_path_factory.mktemp("refapi"))
    R.reference()
    torch.backends.cudnn.allow_tf32 = False  # fp32 contract: cuDNN's default TF32 convolutions would put 1e-3 noise on both arms
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    return R, tmp


def _images(n, seed=1000):
    from tools import synth_data

    x, t = synth_data.synth_batch(n, S, torch.Generator().manual_seed(seed), "cpu")
    return x, t


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def test_model_forward_matches_uninstalled_reference(api):
    """DetectionModel.forward (nn/tasks.py:98-160) of the reference-built model: dense y (B, 4+nc, A) and the raw head maps."""
    R, tmp = api
    m = R.build_yolo(tmp, trained=True)
    net = m.model.to("cuda").eval()
    x = _images(8)[0].to("cuda")
    with torch.no_grad():
        y_ref, maps_ref = net(x)
        with R.installed() as inst:
            y_el, maps_el = net(x)
            assert inst.launches >= 8 + 8 + 8 + 1 + 1, f"only {inst.launches} library kernels ran behind the reference model"
        y_back, _ = net(x)
    assert torch.equal(y_back, y_ref), "uninstall() did not restore the reference forwards"
    assert y_el.shape == y_ref.shape and y_el.dtype == torch.float32
    assert _rel(y_el[:, :4], y_ref[:, :4]) < 1e-5, _rel(y_el[:, :4], y_ref[:, :4])
    assert float((y_el[:, 4:] - y_ref[:, 4:]).abs().max()) < 1e-5
    for a, b in zip(maps_el, maps_ref):
        assert a.shape == b.shape and _rel(a, b) < 1e-5


def test_nms_rows_bit_exact_behind_reference_namespace(api):
    """ultralytics.utils.ops.non_max_suppression (utils/ops.py:167-316) on the SAME decoded tensor: predict and validator settings."""
    R, tmp = api
    import ultralytics.utils.ops as uops

    m = R.build_yolo(tmp, trained=True)
    net = m.model.to("cuda").eval()
    with torch.no_grad():
        y, _ = net(_images(8, seed=7)[0].to("cuda"))
    for kw in (dict(conf_thres=0.25, iou_thres=0.7, max_det=300), dict(conf_thres=0.001, iou_thres=0.7, max_det=300, multi_label=True),
               dict(conf_thres=0.05, iou_thres=0.45, agnostic=True), dict(conf_thres=0.1, iou_thres=0.6, classes=[0, 3, 5])):
        want_gpu = uops.non_max_suppression(y.clone(), max_time_img=1e9, **kw)
        want_cpu = uops.non_max_suppression(y.cpu().clone(), max_time_img=1e9, **kw)
        with R.installed(modules=False, losses=False, criterion=False):
            assert uops.non_max_suppression.__module__ == "edge_yolo_b200.nms"
            got = uops.non_max_suppression(y.clone(), max_time_img=1e9, **kw)
        assert len(got) == len(want_gpu) == 8
        for g, wg, wc in zip(got, want_gpu, want_cpu):
            assert g.shape == wc.shape, (kw, g.shape, wc.shape)
            assert g.cpu().numpy().tobytes() == wc.numpy().tobytes(), f"rows differ from the reference's CPU path for {kw}"
            assert g.cpu().numpy().tobytes() == wg.cpu().numpy().tobytes(), f"rows differ from the reference's CUDA path for {kw}"


def test_predict_through_yolo_api(api):
    """YOLO.predict (engine/model.py:501) -> BasePredictor.stream_inference (engine/predictor.py:220-300) -> DetectionPredictor.postprocess
    (models/yolo/detect/predict.py:23-41) -> Results; tensor source (data/loaders.py:516-570); fp32: boxes / scores within 1e-5."""
    R, tmp = api
    m = R.build_yolo(tmp, trained=True)
    x = _images(8, seed=11)[0]
    kw = dict(device=0, conf=0.25, iou=0.7, max_det=300, half=False, verbose=False)
    res_ref = m.predict(x, **kw)
    with R.installed() as inst:
        res_el = m.predict(x, **kw)
        assert inst.launches > 0
    assert len(res_ref) == len(res_el) == 8
    n_total = 0
    for a, b in zip(res_el, res_ref):
        da, db = a.boxes.data.float().cpu(), b.boxes.data.float().cpu()
        n_total += db.shape[0]
        assert da.shape == db.shape, (da.shape, db.shape)
        assert torch.equal(da[:, 5], db[:, 5]), "classes differ"
        np.testing.assert_allclose(da[:, :4].numpy(), db[:, :4].numpy(), rtol=1e-5, atol=S * 1e-5)
        np.testing.assert_allclose(da[:, 4].numpy(), db[:, 4].numpy(), rtol=0, atol=1e-5)
    assert n_total >= 8, "the synthetic checkpoint should detect at least one object per image"


def test_half_mode_through_yolo_api(api):
    """half=True is the reference's fp16 mode (Q11: AutoBackend casts the model with .half(), engine/predictor.py:306-321).  The
    reference's OWN fp16 outputs sit far from its fp32 outputs on this checkpoint (measured on a B200: confidences of matched detections
    move by up to 0.11, boxes by tens of pixels), so no implementation can meet 2e-2 against fp32 at the detection level; the
    per-op fp16 contract is tested in test_gpu_parity.py.  Here: the installed fp16 model must be at least as close to the fp32
    reference as the reference's own fp16 mode is (mean error of the dense decode output), and YOLO.predict(half=True) must run."""
    R, tmp = api
    # Both arms' fp16 error depends on which algorithms cuDNN picks for the fp16 convolutions, and that depends on what the process ran before:
    # on a B200 the reference's OWN fp16 error on this checkpoint came out as 0.365, 0.466 or 0.638 px depending on which other tests had run
    # (gpurun_out/c46 / c48 / c49 logs: nine orderings), and after the whole of test_gpu_parity.py the installed arm's batched f_h call got a
    # worse algorithm (1.18 px) with the library's kernels unchanged (0.368 px in every other ordering).  The dense comparison therefore runs
    # both arms on PyTorch's native convolution kernels (cuDNN off): what differs between the arms is then only the library's kernels.
    m = R.build_yolo(tmp, trained=True)
    x = _images(8, seed=11)[0]
    net = m.model.to("cuda").eval()
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=False):
        y32, _ = net.float()(x.to("cuda"))
        y16_ref, _ = net.half()(x.to("cuda").half())
        with R.installed() as inst:
            y16_el, _ = net(x.to("cuda").half())
            assert inst.launches > 0
        net.float()
    sel = y32[:, 4:].amax(1) > 0.05  # anchors that matter: background scores are ~1e-4 on either arm
    err = lambda y: (float((y.float()[:, :4] - y32[:, :4]).abs().permute(0, 2, 1)[sel].mean()), float((y.float()[:, 4:] - y32[:, 4:]).abs().permute(0, 2, 1)[sel].mean()))
    (b_ref, s_ref), (b_el, s_el) = err(y16_ref), err(y16_el)
    print(f"\nfp16 model vs fp32 reference, mean |delta| over {int(sel.sum())} foreground anchors: installed box {b_el:.3f} px score {s_el:.5f}; "
          f"the reference's own fp16 mode box {b_ref:.3f} px score {s_ref:.5f}")
    assert b_el <= 1.25 * b_ref + 0.05 and s_el <= 1.25 * s_ref + 1e-3, ((b_el, s_el), (b_ref, s_ref))
    kw = dict(device=0, conf=0.25, iou=0.7, max_det=300, half=True, verbose=False)
    res_ref16 = m.predict(x, **kw)
    m.predictor = None  # the predictor caches its AutoBackend (engine/model.py:545-552)
    with R.installed():
        res_el16 = m.predict(x, **kw)
    m.predictor = None
    n_ref, n_el = sum(len(r.boxes) for r in res_ref16), sum(len(r.boxes) for r in res_el16)
    assert n_el > 0 and abs(n_el - n_ref) <= max(4, n_ref // 4), (n_el, n_ref)
    assert all(r.boxes.data.dtype == torch.float16 or r.boxes.data.dtype == torch.float32 for r in res_el16)


def test_val_map_through_yolo_api(api):
    """YOLO.val (engine/model.py:609) -> DetectionValidator (models/yolo/detect/val.py): dataloader, preprocess, model, NMS with the validator's
    settings (conf 0.001, multi_label), box_iou + match_predictions + ap_per_class.  mAP50-95 within 0.1 points (north_star)."""
    R, tmp = api
    data = R.write_dataset(os.path.join(tmp, "ds64"), 64, S, seed=2000)
    m = R.build_yolo(tmp, trained=True)
    kw = dict(data=data, imgsz=S, batch=32, device=0, workers=0, plots=False, verbose=False, half=False)
    met_ref = m.val(**kw)
    with R.installed() as inst:
        met_el = m.val(**kw)
        assert inst.launches > 0
    with R.installed(metrics=True) as inst:  # + box_iou / match_predictions on the device
        met_el2 = m.val(**kw)
    print(f"\nYOLO.val mAP50-95: reference {100 * met_ref.box.map:.3f}  installed {100 * met_el.box.map:.3f}  installed+metrics {100 * met_el2.box.map:.3f}")
    assert met_ref.box.map > 0.5, f"checkpoint / dataset mismatch: reference mAP {met_ref.box.map}"
    assert abs(met_el.box.map - met_ref.box.map) <= 1e-3 and abs(met_el.box.map50 - met_ref.box.map50) <= 1e-3
    assert abs(met_el2.box.map - met_ref.box.map) <= 1e-3 and abs(met_el2.box.map50 - met_ref.box.map50) <= 1e-3


def test_train_step_loss_and_gradients_through_model_call(api):
    """`model(batch_dict)` (BaseModel.forward -> loss, nn/tasks.py:98-113, 301-313; v8DetectionLoss utils/loss.py:296-420) + backward."""
    R, tmp = api
    from ultralytics.cfg import get_cfg

    m = R.build_yolo(tmp, trained=True)
    net = m.model.to("cuda").train()
    net.args = get_cfg()  # the trainer assigns its namespace here (engine/trainer.py:240); the loss reads .box / .cls / .dfl
    for p in net.parameters():
        p.requires_grad_(True)
    x, t = _images(8, seed=21)
    batch = {"img": x.to("cuda"), "batch_idx": t[:, 0].to("cuda"), "cls": t[:, 1:2].to("cuda"), "bboxes": t[:, 2:6].to("cuda")}

    def run():
        net.zero_grad(set_to_none=True)
        net.criterion = None  # BaseModel.loss (nn/tasks.py:309-310) rebuilds it lazily from the CURRENT nn.tasks.v8DetectionLoss
        loss, items = net(batch)
        loss.backward()
        grads = {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}
        return loss.detach(), items.detach(), grads

    torch.manual_seed(0)
    l_ref, i_ref, g_ref = run()
    with R.installed() as inst:
        torch.manual_seed(0)
        l_el, i_el, g_el = run()
        assert type(net.criterion).__module__ == "edge_yolo_b200.detection_loss"
        assert inst.launches > 20
    net.criterion = None
    assert _rel(l_el, l_ref) < 1e-4 and _rel(i_el, i_ref) < 1e-4, (l_el, l_ref, i_el, i_ref)
    assert set(g_el) == set(g_ref)
    # Per-parameter max |delta| relative to that parameter's largest reference gradient, floored at 1e-4 of the largest gradient of the
    # whole model: several gradients are exactly zero in exact arithmetic (a BatchNorm bias in front of a 1x1 conv + train-mode BatchNorm
    # is cancelled by the mean subtraction, e.g. model.10.m.0.ffn.1.bn.bias) and hold only rounding noise on either arm.
    G = max(float(g.abs().max()) for g in g_ref.values())
    errs = sorted(((float((g_el[n] - g_ref[n]).abs().max()) / (float(g_ref[n].abs().max()) + 1e-4 * G), n) for n in g_ref), reverse=True)
    print("\nworst gradient deviations:", [(f"{e:.2e}", n) for e, n in errs[:5]])
    assert errs[0][0] < 2e-3, errs[:5]  # fp32 sums in another order through ~100 layers of backward


def test_yolo_train_runs_on_the_installed_kernels(api):
    """YOLO.train (engine/model.py:742 -> engine/trainer.py:217-420): one epoch on the synthetic set with install() applied: model built by
    parse_model under the patched names, AMP off, validation at the end, checkpoint pickled (state-dict / qualname compatibility)."""
    R, tmp = api
    data = R.write_dataset(os.path.join(tmp, "ds32"), 32, S, seed=3000)
    m = R.build_yolo(tmp, trained=False)
    with R.installed() as inst:
        m.train(data=data, imgsz=S, batch=16, epochs=1, device=0, workers=0, plots=False, amp=False, verbose=False, project=os.path.join(tmp, "runs"),
                name="t", exist_ok=True, optimizer="SGD", lr0=0.001, warmup_epochs=0)
        n = inst.launches
    assert n > 100, f"only {n} library kernels ran during YOLO.train"
    last = os.path.join(tmp, "runs", "t", "weights", "last.pt")
    assert os.path.exists(last)
    ck = torch.load(last, map_location="cpu", weights_only=False)  # whole module objects are pickled (trainer.py:513-530)
    net = ck["model"] if ck.get("model") is not None else ck["ema"]
    assert type(net).__module__ == "ultralytics.nn.tasks" and all(torch.isfinite(p.float()).all() for p in net.parameters())
