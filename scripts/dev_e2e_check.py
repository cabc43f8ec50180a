"""Development helper: end-to-end check of the dense forward + post-processing against the oracle on one GPU."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dafne_b200.engine import DafneEngine  # noqa: E402
from dafne_b200.spec import ModelSpec  # noqa: E402
from dafne_b200.weights import synthetic_state_dict  # noqa: E402
from oracle import model as omodel  # noqa: E402
from oracle import postprocess as opost  # noqa: E402


def compare_post(spec, eng, heads, sizes, osz, dets, counts):
    logits = [h["logits"].cpu().numpy() for h in heads]
    ctr = [h["ctr_delta"][:, :1].cpu().numpy() for h in heads]
    reg = []
    for l, h in enumerate(heads):
        d = h["ctr_delta"][:, 1:9].cpu().numpy()
        c = h["center"].cpu().numpy()
        reg.append(((np.tile(c, (1, 4, 1, 1)) + d) * np.float32(1.0)).astype(np.float32))
    res = opost.postprocess(logits, reg, ctr, spec.fpn_strides, sizes, osz, score_thresh=spec.score_thresh,
                            pre_nms_topk=spec.pre_nms_topk, nms_thresh=spec.nms_thresh,
                            post_nms_topk=spec.post_nms_topk, sort_corners=spec.sort_corners,
                            thresh_with_ctr=spec.thresh_with_ctr)
    dets = dets.cpu().numpy()
    counts = counts.cpu().numpy()
    for i, r in enumerate(res):
        n = len(r["scores"])
        print(f"  image {i}: oracle n={n} gpu n={counts[i]}")
        if n != counts[i]:
            print("   COUNT MISMATCH")
        m = min(n, counts[i], dets.shape[1])
        g = dets[i, :m]
        canon_g = g[:, 18].view(np.uint32).astype(np.int64)
        same_idx = (canon_g == r["canon"][:m]).all()
        print("   indices equal:", same_idx, " classes equal:", (g[:, 14].astype(np.int64) == r["pred_classes"][:m]).all())
        print("   corners max|d|:", np.abs(g[:, :8] - r["pred_corners"][:m]).max() if m else 0,
              " hbox max|d|:", np.abs(g[:, 8:12] - r["pred_boxes"][:m]).max() if m else 0,
              " score max|d|:", np.abs(g[:, 12] - r["scores"][:m]).max() if m else 0,
              " bit-equal scores:", (g[:, 12].view(np.uint32) == r["scores"][:m].view(np.uint32)).all())
        if not same_idx:
            bad = np.nonzero(canon_g != r["canon"][:m])[0][:5]
            print("   first mismatches at", bad, canon_g[bad], r["canon"][bad])


def main():
    depth = int(os.environ.get("DEPTH", "50"))
    spec = ModelSpec(resnet_depth=depth, num_classes=15, sort_corners=True, thresh_with_ctr=False)
    sd = synthetic_state_dict(spec, seed=0)
    dev = torch.device("cuda:0")
    eng = DafneEngine(spec, dev)
    t0 = time.time()
    eng.load_state_dict(sd)
    print(f"weights loaded in {time.time()-t0:.1f}s")

    # ---- small parity run
    g = torch.Generator().manual_seed(1234)
    N, H, W = 2, 256, 320
    imgs = torch.randint(0, 256, (N, 3, H, W), dtype=torch.uint8, generator=g)
    sizes = [(H, W), (H - 40, W - 24)]
    eng.keep_activations(True)
    eng.forward_dense(imgs.to(dev), sizes)
    torch.cuda.synchronize()
    heads = [eng.head_outputs(l) for l in range(5)]
    batch, _ = omodel.preprocess([imgs[i, :, : sizes[i][0], : sizes[i][1]] for i in range(N)], spec.pixel_mean,
                                 spec.pixel_std)
    assert batch.shape[-2:] == (H, W), batch.shape
    for mode in ("o16", "fp32"):
        t0 = time.time()
        ref = omodel.forward_dense(sd, depth, batch, mode)
        print(f"oracle {mode} forward {time.time()-t0:.1f}s")
        for name, r in ref["named"].items():
            a = eng.activation(name).cpu()
            rel = ((a - r).norm() / (r.norm() + 1e-12)).item()
            if rel > 2e-3 or name in ("stem", "pool", "res2.0", "res5.2", "p3", "p7"):
                print(f"  [{mode}] {name:18s} rel_l2={rel:.3g} max|d|={(a - r).abs().max().item():.3g} absmax={r.abs().max().item():.3g}")
        for l in range(5):
            lg = heads[l]["logits"].cpu()
            cd = heads[l]["ctr_delta"].cpu()
            ce = heads[l]["center"].cpu()
            reg_g = (ce.repeat(1, 4, 1, 1) + cd[:, 1:9])
            e1 = (lg - ref["logits"][l]).abs().max().item()
            e2 = (cd[:, :1] - ref["ctr"][l]).abs().max().item()
            e3 = (reg_g - ref["reg"][l]).abs().max().item()
            print(f"  [{mode}] level {l} {tuple(lg.shape)}: max|dlogit|={e1:.4g} (ref absmax {ref['logits'][l].abs().max():.3g})"
                  f" max|dctr|={e2:.4g} max|dreg|={e3:.4g} (ref absmax {ref['reg'][l].abs().max():.3g})")
    for do_pp in (True,):
        osz = [(H * 2, W * 2), (H - 40, W - 24)]
        dets, counts = eng.postprocess(sizes, osz, do_pp)
        torch.cuda.synchronize()
        print("postprocess counts", counts.tolist())
        compare_post(spec, eng, heads, sizes, osz, dets, counts)

    # ---- timing at the benchmark shape
    eng.keep_activations(False)
    N, H, W = int(os.environ.get("BN", "8")), 1024, 1024
    imgs = torch.randint(0, 256, (N, 3, H, W), dtype=torch.uint8, generator=g).to(dev)
    sizes = [(H, W)] * N
    for _ in range(2):
        eng.forward_dense(imgs, sizes)
    torch.cuda.synchronize()
    print("workspace MB", eng.workspace_bytes / 2**20)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(5):
        eng.forward_dense(imgs, sizes)
    e1.record()
    for _ in range(5):
        dets, counts = eng.postprocess(sizes, None, True)
    e2.record()
    torch.cuda.synchronize()
    tf, tp = e0.elapsed_time(e1) / 5, e1.elapsed_time(e2) / 5
    launches, flops = eng.stats()
    print(f"N={N} forward {tf:.2f} ms ({N/tf*1000:.1f} img/s)  post {tp:.2f} ms  counts {counts.tolist()}")
    lv = [eng.head_outputs(l) for l in range(5)]
    for l in range(5):
        lg = lv[l]["logits"]
        s = torch.sigmoid(lg)
        print(f"  level {l}: logits mean {lg.mean():.3f} std {lg.std():.3f}; frac cls>0.05: {(s > 0.05).float().mean():.5f}")


if __name__ == "__main__":
    main()
