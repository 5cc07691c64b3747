"""MEASUREMENT HARNESS (not part of the product): images/sec of the REFERENCE'S OWN MODEL.

Everything around the hot path is the unmodified reference installed by baseline/install_ref.sh (baseline/_ref/FPQVAR):
`models_fp_quant_transform_rotate.build_vae_var` (random init, `init_weights`; there are no checkpoints offline),
`VAR.autoregressive_infer_cfg` (var.py:135-217) under fp16 autocast with the README's sampling arguments
(evaluate_fp_quant_transform_rotate.py:187-199: cfg=1.5, top_k=900, top_p=0.96, g_seed=0, B=50 per class), its VQVAE
decoder, and the GALT factors it ships (learnable_transformation/best_lambda_var30/*.pt).  What differs between the
modes is only who does the fake quantization:

  fp16        no quantization (the model the paper compares against)
  reference   the reference end to end: its transform_model -> rotate_model(block_rotate=True) -> quantize_VAR, its
              AdaLNSelfAttn.forward (dense [C, C] rotation GEMM, basic_var.py:263,266) and its fp_quant_*_cuda functions
              around its own quant_cuda extension compiled for sm_100a
  dropin      INTEGRATION.md level (b): the same model object and forward, with this repository's transform_rotate_model and
              quantize_VAR (fpqvar_b200.quant_utils: one fused launch per quantizer call)
  fused       level (c): additionally AdaLNSelfAttn.forward calls adaln_transform_rotate_quant_activation (adaLN modulate +
              GALT multiply + block rotation + quantizer in ONE launch) instead of the modulate / .mul(s) / matmul / act_quant
              sequence -- the three-line change a maintainer would make at basic_var.py:263,266 -- and FFN.forward calls
              gelu_fake_quant_signsplit (GELU + fc2's activation quantizer in one launch, basic_var.py:120)

  python tools/ref_model_generate.py --depth 30 --batch 50 --modes fp16,reference,dropin,fused --iters 2
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PATCH_256 = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)
PATCH_512 = (1, 2, 3, 4, 6, 9, 13, 18, 24, 32)


def _load_best_s(ns, depth, C, dev, bits):
    """The shipped GALT factors (fp4 only; SURVEY.md 8d config 4: the fp4 ones are used for W6A6 as well)."""
    d = os.path.join(ns.root, "learnable_transformation", f"best_lambda_var{depth}")
    out = []
    for name in ("mat_qkv_best_s_fp4.pt", "fc1_best_s_fp4.pt"):
        p = os.path.join(d, name)
        if os.path.isfile(p):
            s = torch.load(p, map_location=dev)
            s = [t.detach().to(device=dev, dtype=torch.float32).reshape(-1) for t in s]
            if len(s) >= depth and s[0].numel() == C:
                out.append(s[:depth])
                continue
        g = torch.Generator(device="cpu").manual_seed(len(out))
        out.append([torch.exp(0.3 * torch.randn(C, generator=g)).to(dev) for _ in range(depth)])
    return out[0], out[1], os.path.isdir(d)


def _fused_forward(self, x, cond_BD, attn_bias, step_idx, quant_KV, kv_bit, rotation_matrix, mat_qkv_best_s, fc1_best_s):
    """AdaLNSelfAttn.forward (basic_var.py:253-270) with the two modulate / smooth / rotate / act_quant sequences replaced
    by the fused call; mat_qkv / fc1 receive already-quantized inputs (their act_quant is the identity in this mode)."""
    from fpqvar_b200.rotation_utils import adaln_transform_rotate_quant_activation as fused
    if self.shared_aln:
        gamma1, gamma2, scale1, scale2, shift1, shift2 = (self.ada_gss + cond_BD).unbind(2)
    else:
        gamma1, gamma2, scale1, scale2, shift1, shift2 = self.ada_lin(cond_BD).view(-1, 1, 6, self.C).unbind(2)
    x_1 = fused(self.ln_wo_grad(x), scale1, shift1, mat_qkv_best_s, self._fpq_act_fp_type)
    x = x + self.drop_path(self.attn(x_1, attn_bias, self.block_idx, step_idx, quant_KV, kv_bit).mul_(gamma1))
    x_2 = fused(self.ln_wo_grad(x), scale2, shift2, fc1_best_s, self._fpq_act_fp_type)
    x = x + self.drop_path(self.ffn(x_2, self.block_idx, step_idx).mul(gamma2))
    return x


def _init_vqvae(vae):
    """The reference's build_vae_var switches torch's built-in initialisers off (models*/__init__.py:23-25: the VQVAE is
    meant to be filled from a checkpoint) and VAR.init_weights covers the transformer only, so without a checkpoint the VQVAE
    holds uninitialised memory (measured: conv weights of 4e21).  Give it a plain deterministic init (the decoder's cost
    does not depend on the values; they only have to stay finite in fp16)."""
    g = torch.Generator(device="cpu").manual_seed(1)
    with torch.no_grad():
        for name, p in list(vae.named_parameters()) + list(vae.named_buffers()):
            if not p.is_floating_point():
                p.zero_()
            elif p.dim() >= 2:
                fan_in = p[0].numel()
                p.copy_((torch.randn(p.shape, generator=g) * (0.5 / max(1, fan_in) ** 0.5)).to(p.device, p.dtype))
            elif "norm" in name and name.endswith("weight"):
                p.fill_(1.0)
            else:
                p.zero_()
        emb = getattr(getattr(vae, "quantize", None), "embedding", None)
        if emb is not None:
            emb.weight.copy_(torch.randn(emb.weight.shape, generator=g).to(emb.weight.device))


def _fused_ffn_forward(self, x, block_idx, step_idx):
    """FFN.forward (basic_var.py:110-121) with `self.act` and fc2's activation quantizer in one pass
    (fpq_gelu_fake_quant_signsplit) followed by the bare GEMM on fc2's quantized weight."""
    from fpqvar_b200 import ops
    h = self.fc1(x)                                    # QuantizedLinear: its act_quant is the identity in this mode
    hq = ops.gelu_fake_quant_signsplit(h, self._fpq_fc2_split, global_clip=True)
    return self.drop(torch.nn.functional.linear(hq, self.fc2.weight, self.fc2.bias))


def build(ns, dev, depth, res, bits, mode):
    patch_nums = PATCH_512 if res == 512 else PATCH_256
    torch.manual_seed(0)
    vae, var = ns.build_vae_var(V=4096, Cvae=32, ch=160, share_quant_resi=4, device=dev, patch_nums=patch_nums,
                                num_classes=1000, depth=depth, shared_aln=(res == 512))
    _init_vqvae(vae)
    vae.eval().to(dev)
    var.eval().to(dev)
    for p in list(vae.parameters()) + list(var.parameters()):
        p.requires_grad_(False)
    C = var.C
    fmt = {4: ("fp_e2", "fp_e2", "fp_e1m2_neg_e2m1_pos"), 6: ("fp6_e2m3", "fp6_e2m3", "fp6_int_neg_e2m3_pos")}[bits]
    s_qkv, s_fc1, shipped = _load_best_s(ns, depth, C, dev, bits)
    info = {"galt_factors": "shipped best_lambda_var%d/*_fp4.pt" % depth if shipped else "synthetic log-normal"}
    qargs = dict(weight_quant="per_group", act_quant="per_group", quantize_bmm_input=False, w_bit=bits, a_bit=bits,
                 act_quant_sym=True, fc2_act_log2_quant=False, quant_kv=False, kv_bit=8, activation_fp_quant=True,
                 weight_fp_quant=True, act_fp_type=fmt[0], weight_fp_type=fmt[1], fc2_fp_type=fmt[2])
    if mode == "fp16":
        s_qkv = [torch.ones(C, device=dev) for _ in range(depth)]
        s_fc1 = [torch.ones(C, device=dev) for _ in range(depth)]
        Q = torch.eye(C, device=dev)
        var = var.half()
    elif mode == "reference":
        ns.transform_model_utils.transform_model(var, s_qkv, s_fc1)                  # evaluate_fp_quant_transform_rotate.py:99
        ns.rotation_utils.rotate_model(var, dev, True)                               # :106
        ns.rotation_utils.cleanup_memory()
        var = ns.qu.quantize_VAR(var, **qargs)                                       # :114-130
        var = var.half()
        Q = ns.rotation_utils.block_random_hadamard_matrix(total_size=C, block_size=128, device=dev, seed=42).to(torch.float32)
    else:
        from fpqvar_b200 import quant_utils as our_qu
        from fpqvar_b200 import transform_model_utils as our_tm
        from fpqvar_b200 import rotation_utils as our_rot
        our_tm.transform_rotate_model(var, s_qkv, s_fc1)
        var = our_qu.quantize_VAR(var, **qargs)
        var = var.half()
        Q = our_rot.block_random_hadamard_matrix(total_size=C, block_size=128, device=dev, seed=42).to(torch.float32)
        if mode == "fused":
            for b in var.blocks:
                b._fpq_act_fp_type = fmt[0]
                b.forward = types.MethodType(_fused_forward, b)
                b.attn.mat_qkv.act_quant = lambda t: t
                b.ffn.fc1.act_quant = lambda t: t
                b.ffn._fpq_fc2_split = {"fp_e1m2_neg_e2m1_pos": "e1m2_neg_e2m1_pos", "fp6_int_neg_e2m3_pos": "int_neg_e2m3_pos"}[fmt[2]]
                b.ffn.forward = types.MethodType(_fused_ffn_forward, b.ffn)
            Q = None
    return vae, var, Q, s_qkv, s_fc1, info


def measure(ns, dev, depth, batch, res, bits, mode, iters, warmup=1, rank=0, world=1):
    from fpqvar_b200 import ops
    # the reference's "run faster" settings, evaluate_fp_quant_transform_rotate.py:171-175
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.set_float32_matmul_precision("high")
    vae, var, Q, s_qkv, s_fc1, info = build(ns, dev, depth, res, bits, mode)
    times = []
    finite = True
    n0 = ops.launch_count()
    for it in range(warmup + iters):
        label = torch.full((batch,), (rank + world * it) % 1000, device=dev, dtype=torch.long)       # classes sharded over ranks
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.inference_mode():
            with torch.autocast("cuda", enabled=True, dtype=torch.float16, cache_enabled=True):
                img = var.autoregressive_infer_cfg(B=batch, label_B=label, cfg=1.5, top_k=900, top_p=0.96, g_seed=0, more_smooth=False,
                                                   rotation_matrix=Q, quant_KV=False, kv_bit=8, mat_qkv_best_s=s_qkv, fc1_best_s=s_fc1)
        e1.record()
        torch.cuda.synchronize(dev)
        if it == warmup:
            n0 = ops.launch_count()
        if it >= warmup:
            times.append(e0.elapsed_time(e1))
        finite = finite and bool(torch.isfinite(img).all().item()) and tuple(img.shape) == (batch, 3, res, res)
    ms = sum(times) / len(times)
    del vae, var
    torch.cuda.empty_cache()
    return {"mode": mode, "ms_per_batch": ms, "images_per_sec": batch / (ms / 1e3), "finite": finite,
            "fpq_launches_per_batch": (ops.launch_count() - n0) // max(1, iters), **info}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=30)
    ap.add_argument("--batch", type=int, default=50)
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--bits", type=int, default=4)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--modes", default="fp16,reference,dropin,fused")
    args = ap.parse_args()
    from baseline import ref_env
    if not ref_env.available():
        print(json.dumps({"unavailable": ref_env.why_unavailable()}))
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    ns = ref_env.load(str(dev))
    for mode in args.modes.split(","):
        r = measure(ns, dev, args.depth, args.batch, args.res, args.bits, mode, args.iters)
        print(json.dumps({"model": f"reference VAR-d{args.depth} {args.res}x{args.res} B={args.batch} W{args.bits}A{args.bits} random init", **r}), flush=True)


if __name__ == "__main__":
    main()
