import os, sys, torch
sys.path.insert(0, os.getcwd())
from baseline import ref_env
dev = torch.device("cuda:0")
ns = ref_env.load(str(dev))
for flash in (True, False):
    torch.manual_seed(0)
    vae, var = ns.build_vae_var(V=4096, Cvae=32, ch=160, share_quant_resi=4, device=dev, patch_nums=(1,2,3,4,5,6,8,10,13,16), num_classes=1000, depth=16, shared_aln=False, flash_if_available=flash)
    vae.eval().to(dev); var.eval().to(dev)
    var = var.half()
    bad = []
    def hook(name):
        def f(m, i, o):
            t = o[0] if isinstance(o, (tuple, list)) else o
            if torch.is_tensor(t) and not torch.isfinite(t).all() and len(bad) < 5:
                bad.append((name, tuple(t.shape), str(t.dtype), float(torch.nan_to_num(t.float(), nan=0, posinf=0, neginf=0).abs().max())))
        return f
    for n, m in var.named_modules():
        if n and n.count(".") <= 3: m.register_forward_hook(hook(n))
    vq = var.vae_quant_proxy[0]
    orig = vq.get_next_autoregressive_input
    def chk(name, t):
        t = t.float()
        fin = torch.isfinite(t)
        print("   ", name, tuple(t.shape), "finite", bool(fin.all()), "absmax", float(t[fin].abs().max()) if fin.any() else None, flush=True)
    def wrapped(si, SN, f_hat, h):
        chk(f"si={si} h_in", h); chk(f"si={si} f_hat_in", f_hat)
        import torch.nn.functional as F
        up = F.interpolate(h, size=(16, 16), mode='bicubic'); chk(f"si={si} bicubic({h.dtype})", up)
        phi = vq.quant_resi[si/(SN-1)]
        chk("phi.weight", phi.weight); chk("phi.bias", phi.bias)
        cv = torch.nn.Conv2d.forward(phi, up); chk(f"conv out {cv.dtype}", cv)
        with torch.autocast("cuda", enabled=False):
            cv32 = torch.nn.Conv2d.forward(phi, up.float()); chk("conv out fp32", cv32)
        out = orig(si, SN, f_hat, h)
        chk(f"si={si} f_hat_out", out[0]); chk(f"si={si} next", out[1])
        return out
    vq.get_next_autoregressive_input = wrapped
    print("embedding absmax", float(vq.embedding.weight.abs().max()), vq.embedding.weight.dtype, "word_embed w absmax", float(var.word_embed.weight.abs().max()))
    C = var.C
    s = [torch.ones(C, device=dev) for _ in range(16)]
    Q = torch.eye(C, device=dev)
    label = torch.zeros(4, dtype=torch.long, device=dev)
    try:
        with torch.inference_mode(), torch.autocast("cuda", enabled=True, dtype=torch.float16):
            img = var.autoregressive_infer_cfg(B=4, label_B=label, cfg=1.5, top_k=900, top_p=0.96, g_seed=0, more_smooth=False, rotation_matrix=Q, quant_KV=False, kv_bit=8, mat_qkv_best_s=s, fc1_best_s=s)
        torch.cuda.synchronize()
        print("flash", flash, "OK finite", bool(torch.isfinite(img).all()), "bad:", bad)
    except Exception as e:
        print("flash", flash, "EXC", type(e).__name__, str(e)[:200], "bad:", bad)
        break
