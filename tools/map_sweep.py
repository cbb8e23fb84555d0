"""mAP of the three arms of tests/test_map_parity.py over several validation seeds (how wide is the bf16 gap?)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from edge_yolo_b200.engine import Predictor  # noqa: E402
from edge_yolo_b200.model import EdgeLineYOLO  # noqa: E402
from edge_yolo_b200.nms import non_max_suppression  # noqa: E402
from oracle import metrics_ref, model_ref  # noqa: E402
from tools import synth_data  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
B, S, NC = 64, 256, synth_data.NC
state = {k: (v.float() if v.is_floating_point() else v) for k, v in torch.load("tests/golden/edgeline_n_synth.pt", map_location="cpu").items()}
ref = model_ref.build("n", NC, seed=0); ref.load_state_dict(state)
api = EdgeLineYOLO("n", NC).eval(); api.load_state_dict(state); api = api.to("cuda")
eng = EdgeLineYOLO("n", NC).eval(); eng.load_state_dict(state)
eng = eng.fuse(engine=True).to(device="cuda", dtype=torch.bfloat16).to(memory_format=torch.channels_last)
pred = Predictor(eng, batch=B, imgsz=S, conf=0.001, iou=0.7, max_det=300, multi_label=True)
h16 = EdgeLineYOLO("n", NC).eval(); h16.load_state_dict(state); h16 = h16.to("cuda", torch.bfloat16).to(memory_format=torch.channels_last)
tot = [0, 0, 0, 0]
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    x, t = synth_data.synth_batch(B, S, torch.Generator().manual_seed(1000 + seed), "cpu")
    labels = synth_data.labels_xyxy(t, B, S)
    u8 = (x * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    m_ref = metrics_ref.evaluate(model_ref.predict(ref, x, conf=0.001, iou=0.7, max_det=300, multi_label=True), labels)[0]
    with torch.no_grad():
        y, _ = api(x.to("cuda"))
        m_api = metrics_ref.evaluate([d.cpu().numpy() for d in non_max_suppression(y, conf_thres=0.001, iou_thres=0.7, max_det=300, multi_label=True)], labels)[0]
        y, _ = h16(x.to("cuda", torch.bfloat16).contiguous(memory_format=torch.channels_last))  # unfused module graph in bf16 (cuDNN convs)
        m_h16 = metrics_ref.evaluate([d.cpu().numpy() for d in non_max_suppression(y, conf_thres=0.001, iou_thres=0.7, max_det=300, multi_label=True)], labels)[0]
    m_eng = metrics_ref.evaluate([d.numpy() for d in pred.predict(u8.pin_memory())], labels)[0]
    print(f"seed {seed}: ref {100*m_ref:.3f} api {100*m_api:.3f} bf16-modules {100*m_h16:.3f} engine {100*m_eng:.3f}", flush=True)
    for i, v in enumerate((m_ref, m_api, m_h16, m_eng)):
        tot[i] += v
n = seed + 1
print("mean: ref %.3f api %.3f bf16-modules %.3f engine %.3f" % tuple(100 * v / n for v in tot))
