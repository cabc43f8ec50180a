"""Development helper: print the detections of tests/test_dense_gpu.py::test_end_to_end_detections_vs_fp32_oracle that
differ from the fp32 oracle by more than the tolerance."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import model as omodel  # noqa: E402
from oracle import postprocess as opost  # noqa: E402
from dafne_b200.engine import DafneEngine  # noqa: E402
from dafne_b200.spec import ModelSpec  # noqa: E402
from dafne_b200.weights import synthetic_state_dict  # noqa: E402

spec = ModelSpec(resnet_depth=50, num_classes=15)
sd = synthetic_state_dict(spec, seed=0)
eng = DafneEngine(spec, torch.device("cuda:0"))
eng.load_state_dict(sd)
g = torch.Generator().manual_seed(99)
H, W = 224, 288
imgs = [torch.randint(0, 256, (3, H, W), dtype=torch.uint8, generator=g),
        torch.randint(0, 256, (3, H - 30, W - 50), dtype=torch.uint8, generator=g)]
batch_u8 = torch.zeros(2, 3, H, W, dtype=torch.uint8)
sizes = [(H, W), (H - 30, W - 50)]
for i, im in enumerate(imgs):
    batch_u8[i, :, : sizes[i][0], : sizes[i][1]] = im
eng.forward_dense(batch_u8.cuda(), sizes)
torch.cuda.synchronize()
batch, _ = omodel.preprocess(imgs, spec.pixel_mean, spec.pixel_std)
ref32 = omodel.forward_dense(sd, 50, batch, "fp32")
want = opost.postprocess([t.numpy() for t in ref32["logits"]], [t.numpy() for t in ref32["reg"]],
                         [t.numpy() for t in ref32["ctr"]], spec.fpn_strides, sizes, None)
dets, counts = eng.postprocess(sizes, None, True)
dets, counts = dets.cpu().numpy(), counts.cpu().numpy()
np.set_printoptions(precision=2, suppress=True, linewidth=200)
for i, w in enumerate(want):
    got = {int(c): k for k, c in enumerate(dets[i, : counts[i], 18].view(np.uint32))}
    common = [(k, got[int(c)]) for k, c in enumerate(w["canon"]) if int(c) in got]
    a = np.array([k for k, _ in common])
    b = np.array([k for _, k in common])
    g4 = dets[i, b, 0:8].reshape(-1, 4, 2)
    w4 = w["pred_corners"][a].reshape(-1, 4, 2)
    orders = [np.roll(np.arange(4), s) for s in range(4)] + [np.roll(np.arange(4)[::-1], s) for s in range(4)]
    d = np.stack([np.abs(g4[:, o] - w4).reshape(len(a), -1).max(1) for o in orders], 1)
    tol = np.maximum(1.0, 0.06 * np.array(spec.fpn_strides, np.float32)[dets[i, b, 15].astype(np.int64)])
    print(f"image {i}: want {len(w['canon'])} got {counts[i]} common {len(common)}")
    for r in np.nonzero(d.min(1) > tol)[0]:
        print("  level", dets[i, b[r], 15], "score", dets[i, b[r], 12], w["scores"][a[r]], "dmin", d[r].min())
        print("   got ", dets[i, b[r], 0:8])
        print("   want", w["pred_corners"][a[r]])
