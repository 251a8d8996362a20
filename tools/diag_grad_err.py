"""Per-parameter gradient error of the bf16 path against the fp64 oracle for the parity-test cases."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_vit_b200 as nv
from oracle import vit_oracle as O
from tests.conftest import rel_err
from tests.test_gpu_model import CASES, _step, _oracle_step
DEV = "cuda"
for tag, kw, batch in CASES:
    kw = dict(kw, dropout=0.0, attention_dropout=0.0, drop_path=0.0)
    cfg = nv.Temporal3DViTConfig(**kw)
    params = O.random_params(O.config_from(cfg), seed=11)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(batch, cfg.n_trials, cfg.freq_size, cfg.time_size, generator=g).to(DEV)
    y = torch.randint(0, 2, (batch,), generator=g).to(DEV)
    rl, rloss, rg = _oracle_step(kw, params, x, y)
    for prec in ("bf16", "bf16_tcgemm", "bf16_simt"):
        m = nv.Temporal3DViT(cfg, precision=prec)
        m.load_state_dict(params)
        m.to(DEV).train()
        logits, loss, grads = _step(m, x, y)
        errs = sorted(((rel_err(grads[k], rg[k]), k, grads[k].numel()) for k in rg), reverse=True)
        flat = torch.cat([grads[k].double().flatten() for k in rg]); fr = torch.cat([rg[k].double().flatten() for k in rg])
        big = [(round(e, 4), k) for e, k, n in errs if n > 4096][:5]
        small = [(round(e, 4), k) for e, k, n in errs if n <= 4096][:3]
        print(f"{tag} {prec}: logits {rel_err(logits, rl):.4f} whole-grad {rel_err(flat, fr):.4f} big {big} small {small}", flush=True)
