"""Development helper: pass-rate of the score threshold as a function of the synthetic cls bias (1024^2 noise images)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dafne_b200.engine import DafneEngine
from dafne_b200.spec import ModelSpec
from dafne_b200.weights import synthetic_state_dict, SYNTH_CLS_BIAS

for depth in (50, 101):
    spec = ModelSpec(resnet_depth=depth, num_classes=15)
    eng = DafneEngine(spec, torch.device("cuda:0"))
    eng.load_state_dict(synthetic_state_dict(spec, 0))
    g = torch.Generator().manual_seed(1234)
    N = 4
    imgs = torch.randint(0, 256, (N, 3, 1024, 1024), dtype=torch.uint8, generator=g).cuda()
    eng.forward_dense(imgs, [(1024, 1024)] * N)
    torch.cuda.synchronize()
    for l in range(5):
        h = eng.head_outputs(l)
        lg = h["logits"].double() - SYNTH_CLS_BIAS
        ct = torch.sigmoid(h["ctr_delta"][:, :1].double())
        row = []
        for b in (-3.5, -4.0, -4.25, -4.5, -5.0, -5.5, -6.0, -6.5, -7.0):
            cls = torch.sigmoid(lg + b)
            f0 = (cls > 0.05).double().mean().item()
            f1 = ((cls * ct).sqrt() > 0.05).double().mean().item()
            row.append(f"{b}:{f0:.4f}/{f1:.4f}")
        print(f"R{depth} level {l} std {lg.std():.3f} ctr mean {ct.mean():.3f} | " + " ".join(row), flush=True)
    eng.close()
