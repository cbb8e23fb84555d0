"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/tal.npz: outputs of the UNMODIFIED reference's `TaskAlignedAssigner.forward`
(utils/tal.py:14-295, CIoU from utils/metrics.py:74-134) on seeded trained-like inputs.  Dev container only (needs /root/reference):

    python -m oracle.gen_golden_tal

Each case stores the six inputs and the five outputs (target_labels, target_bboxes, target_scores, fg_mask, target_gt_idx).  Seeds are
searched so that every DISCRETE decision of the assigner has a margin far above fp32 rounding: the metric of the last selected anchor of
every ground truth is >= 0.1 % above the first rejected one, and an anchor claimed by several ground truths has a >= 0.1 % gap between
its two best overlaps.  A CUDA implementation whose transcendental functions differ from the CPU's in the last bits must then reproduce
the same assignment exactly; only the target scores carry a tolerance."""
import os

import numpy as np
import torch

from . import ref_loader

CASES = {  # tag: (B, imgsz, nc, n_gt, topk, alpha, beta)
    "multi": (3, 256, 4, 16, 13, 0.5, 6.0),   # crowded: dozens of multi-claimed anchors (what v8DetectionLoss uses: alpha .5, beta 6)
    "sparse": (3, 192, 8, 5, 10, 1.0, 6.0),   # TaskAlignedAssigner's own defaults for alpha / beta
}


def make_case(B, imgsz, nc, M, seed):
    """Predicted boxes scattered around their cells, sparse class probabilities, trailing padded ground truths, last image without targets."""
    g = torch.Generator().manual_seed(seed)
    pts, st = [], []
    for s in (8, 16, 32):
        n = imgsz // s
        gy, gx = torch.meshgrid(torch.arange(n, dtype=torch.float32) + 0.5, torch.arange(n, dtype=torch.float32) + 0.5, indexing="ij")
        pts.append(torch.stack((gx, gy), -1).reshape(-1, 2))
        st.append(torch.full((n * n, 1), float(s)))
    pts, st = torch.cat(pts), torch.cat(st)
    anchors = pts * st
    A = anchors.shape[0]
    ltrb = torch.rand(B, A, 4, generator=g) * 6 * st
    boxes = torch.cat((anchors - ltrb[..., :2], anchors + ltrb[..., 2:]), -1)
    scores = torch.sigmoid(torch.randn(B, A, nc, generator=g) * 2 - 3)
    c = torch.rand(B, M, 2, generator=g) * 0.8 + 0.1
    wh = torch.rand(B, M, 2, generator=g) * 0.35 + 0.05
    gt_boxes = torch.cat((c - wh / 2, c + wh / 2), -1) * imgsz
    gt_labels = torch.randint(0, nc, (B, M, 1), generator=g).float()
    n_valid = torch.randint(1, M + 1, (B,), generator=g)
    n_valid[-1] = 0
    valid = (torch.arange(M)[None, :] < n_valid[:, None]).unsqueeze(-1)
    return scores, boxes, anchors, gt_labels * valid, gt_boxes * valid, valid


def margins_ok(assigner, args, topk, rel=1e-3):
    """Decision margins of the reference's own intermediate tensors (get_pos_mask, tal.py:120-130)."""
    scores, boxes, anchors, gt_labels, gt_boxes, valid = args
    assigner.bs, assigner.n_max_boxes = scores.shape[0], gt_boxes.shape[1]  # what forward() sets before calling get_pos_mask (tal.py:60-61)
    mask_pos, metric, overlaps = assigner.get_pos_mask(scores, boxes, gt_labels, gt_boxes, anchors, valid)
    srt = metric.sort(-1, descending=True).values
    last_in, first_out = srt[..., topk - 1], srt[..., topk]
    v = valid.squeeze(-1).bool()
    if bool(((last_in - first_out) < rel * last_in)[v & (first_out > 0)].any()) or bool((last_in <= 0)[v].any()):
        return False  # a near-tie at the top-k cut, or a ground truth with fewer than topk positive metrics (open tie among zeros)
    multi = mask_pos.sum(1) > 1
    if not bool(multi.any()):
        return True
    two = overlaps.topk(2, dim=1).values                       # (B, 2, A)
    return not bool(((two[:, 0] - two[:, 1]) < rel * two[:, 0])[multi].any())


def main():
    ref_loader.load()
    from ultralytics.utils.tal import TaskAlignedAssigner

    out = {}
    for tag, (B, imgsz, nc, M, topk, alpha, beta) in CASES.items():
        ta = TaskAlignedAssigner(topk=topk, num_classes=nc, alpha=alpha, beta=beta)
        for seed in range(1000):
            args = make_case(B, imgsz, nc, M, seed)
            if margins_ok(ta, args, topk):
                break
        else:
            raise SystemExit(f"{tag}: no seed with safe margins")
        labels, tboxes, tscores, fg, gt_idx = ta(*args)
        n_multi = int((ta.get_pos_mask(args[0], args[1], args[3], args[4], args[2], args[5])[0].sum(1) > 1).sum())
        print(f"{tag}: seed {seed}, A = {args[2].shape[0]}, foreground {int(fg.sum())}, multi-claimed anchors {n_multi}")
        out[f"{tag}_cfg"] = np.array([B, imgsz, nc, M, topk], np.int64)
        out[f"{tag}_alpha_beta"] = np.array([alpha, beta], np.float32)
        for k, t in zip(("scores", "boxes", "anchors", "gt_labels", "gt_boxes", "gt_valid"), args):
            out[f"{tag}_{k}"] = t.numpy()
        for k, t in zip(("labels", "tboxes", "tscores", "fg", "gt_idx"), (labels, tboxes, tscores, fg, gt_idx)):
            out[f"{tag}_out_{k}"] = t.numpy()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tal.npz")
    np.savez_compressed(path, **out)
    print(path, f"{os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
