"""MEASUREMENT HARNESS (not part of the product): a shape-faithful VAR class-conditional generation pass
— the CALLER of the hot path — used to quote BASELINE.json's second metric, images/sec, with the hot path
in place (SURVEY.md §8d "Images/sec").

The network follows the reference's inference structure (models_fp_quant_transform_rotate/var.py:140-230
`autoregressive_infer_cfg`, basic_var.py:225-284 `AdaLNSelfAttn`, :128-221 `SelfAttention` with l2-normalised
q/k and a KV cache, :99-123 `FFN` with tanh-GELU, quant.py:187-197 next-scale input, basic_vae.py:163-226
decoder) with random-init weights of the named sizes: width = 64*depth, heads = depth, mlp_ratio 4, V=4096,
Cvae=32, decoder ch=160 x (1,1,2,2,4).  It is written for this harness (explicit dtypes, preallocated KV
cache, torch SDPA) and is NOT weight-compatible with the reference's checkpoints.  GEMMs, attention,
convolutions, LayerNorm, GELU and sampling are library (torch) calls; only the quantizers are this repo's.

Modes (the quantized linears, 4 per block):
  fp16      no quantization (the FP16 model the paper compares against)
  modules   level (b) of INTEGRATION.md: fpqvar_b200.quant_utils.quantize_VAR + transform_rotate_model, and the
            reference's online call pattern  matmul(LN(x).mul(1+scale).add_(shift).mul(s), Q) -> QuantizedLinear
  fused     level (c): adaLN modulate + smoothing + block rotation + quantizer in ONE launch
            (adaln_transform_rotate_quant_activation), then F.linear on the quantized weight
  refquant  `modules` with every activation quantizer swapped for the reference's own path: its Python glue
            (tests/ref_glue.py restates it op for op) around its UNMODIFIED quant_cuda extension
            (oracle/_ref, compiled for sm_100a) — the reference arm for this metric

  python tools/var_generate.py --depth 30 --batch 50 --mode fused --iters 3
  torchrun --nproc-per-node 8 tools/var_generate.py ...      # classes sharded over ranks, no collective on the path
Prints one JSON line per mode (rank 0)."""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fpqvar_b200 import kv_cache, ops, quant_utils, rotation_utils, transform_model_utils  # noqa: E402

PATCH_256 = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)
PATCH_512 = (1, 2, 3, 4, 6, 9, 13, 18, 24, 32)


# ------------------------------------------------------------------------------------------ transformer
class SelfAttention(nn.Module):                      # class NAME is what quantize_VAR keys on
    def __init__(self, C, heads):
        super().__init__()
        self.C, self.H, self.hd = C, heads, C // heads
        self.mat_qkv = nn.Linear(C, 3 * C, bias=False)
        self.proj = nn.Linear(C, C)
        self.qkv_bias = nn.Parameter(torch.zeros(3 * C))            # (q_bias, 0, v_bias) of basic_var.py:160
        self.log_scale_mul = nn.Parameter(torch.full((1, 1, heads, 1), 4.0).log())
        self.k_cache = self.v_cache = None
        self.cur = 0
        self.kv = None                                               # IncrementalKVQuant when --quant-kv is on

    def reset_cache(self, rows, L, device, kv_bit=0, kv_incremental=True):
        if kv_bit:
            if self.kv is None or self.kv.max_len != L or self.kv.kv_bit != kv_bit or self.kv.incremental != kv_incremental:
                self.kv = kv_cache.IncrementalKVQuant(kv_bit, L, incremental=kv_incremental)
            self.kv.reset()
            return
        self.kv = None
        if self.k_cache is None or self.k_cache.shape[0] != rows or self.k_cache.shape[1] != L:
            self.k_cache = torch.empty(rows, L, self.H, self.hd, dtype=torch.float16, device=device)
            self.v_cache = torch.empty_like(self.k_cache)
        self.cur = 0

    def attend(self, qkv):                                            # qkv: fp16 [B, l, 3C] (bias already added)
        B, l, _ = qkv.shape
        q, k, v = qkv.view(B, l, 3, self.H, self.hd).unbind(2)
        q = F.normalize(q, dim=-1).mul(self.log_scale_mul.clamp_max(math.log(100)).exp().to(q.dtype))
        k = F.normalize(k, dim=-1)
        if self.kv is not None:                                       # basic_var.py:188-203 with --quant_kv
            kk, vv = self.kv.append(k, v)
        else:
            self.k_cache[:, self.cur:self.cur + l] = k
            self.v_cache[:, self.cur:self.cur + l] = v
            self.cur += l
            kk, vv = self.k_cache[:, :self.cur], self.v_cache[:, :self.cur]
        o = F.scaled_dot_product_attention(q.transpose(1, 2), kk.transpose(1, 2), vv.transpose(1, 2), scale=1.0)
        return o.transpose(1, 2).reshape(B, l, self.C)

    def forward(self, x):                                             # x: fp16, already modulated/rotated
        return self.proj(self.attend(self.mat_qkv(x) + self.qkv_bias.to(x.dtype)))


class FFN(nn.Module):
    def __init__(self, C):
        super().__init__()
        self.fc1 = nn.Linear(C, 4 * C)
        self.fc2 = nn.Linear(4 * C, C)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x), approximate="tanh"))


class Block(nn.Module):
    def __init__(self, C, heads, shared_aln):
        super().__init__()
        self.C = C
        self.attn = SelfAttention(C, heads)
        self.ffn = FFN(C)
        self.shared_aln = shared_aln
        if shared_aln:
            self.ada_gss = nn.Parameter(torch.randn(1, 1, 6, C) / C ** 0.5)
        else:
            self.ada_lin = nn.Linear(C, 6 * C)


class Var(nn.Module):
    def __init__(self, depth, patch_nums, shared_aln, V=4096, Cvae=32, num_classes=1000):
        super().__init__()
        self.depth, self.C, self.V, self.Cvae = depth, 64 * depth, V, Cvae
        self.patch_nums, self.num_classes = patch_nums, num_classes
        self.L = sum(p * p for p in patch_nums)
        C = self.C
        self.class_emb = nn.Embedding(num_classes + 1, C)
        self.pos_start = nn.Parameter(torch.empty(1, patch_nums[0] ** 2, C))
        self.pos_1LC = nn.Parameter(torch.empty(1, self.L, C))
        self.lvl_embed = nn.Embedding(len(patch_nums), C)
        self.word_embed = nn.Linear(Cvae, C)
        self.shared_ada_lin = nn.Linear(C, 6 * C) if shared_aln else None
        self.blocks = nn.ModuleList(Block(C, depth, shared_aln) for _ in range(depth))
        self.head_ada = nn.Linear(C, 2 * C)
        self.head = nn.Linear(C, V)
        self.register_buffer("lvl_1L", torch.cat([torch.full((p * p,), i) for i, p in enumerate(patch_nums)]).view(1, -1))
        # VQ side (quant.py): codebook + 4 partially-shared phi convs
        self.codebook = nn.Embedding(V, Cvae)
        self.phi = nn.ModuleList(nn.Conv2d(Cvae, Cvae, 3, padding=1) for _ in range(4))
        self.mode = "fp16"
        self.fc2_split_fmt = None
        self.kv_bit, self.kv_incremental = 0, True
        self.smooth_qkv = self.smooth_fc1 = None
        self.Q = None

    @torch.no_grad()
    def init_weights(self, seed=0):
        g = torch.Generator(device=self.lvl_1L.device).manual_seed(seed)
        std = math.sqrt(1 / self.C / 3)
        for name, p in self.named_parameters():
            if name.endswith("bias"):
                p.zero_()
            elif name.endswith("log_scale_mul"):
                continue
            elif "ada" in name:
                p.normal_(0, 0.01, generator=g)
            else:
                p.normal_(0, std if ("emb" in name or "pos" in name) else 0.02, generator=g)

    # -- one transformer block, per mode ------------------------------------------------------------
    def _ada(self, b, cond16, shared):
        if b.shared_aln:
            return (b.ada_gss.to(shared.dtype) + shared).unbind(2)                       # 6 x [B, 1, C]
        return b.ada_lin(F.silu(cond16)).view(-1, 1, 6, self.C).unbind(2)

    def _block(self, i, b, x, cond16, shared):
        gamma1, gamma2, scale1, scale2, shift1, shift2 = self._ada(b, cond16, shared)
        ln = F.layer_norm(x, (self.C,), eps=1e-6)                                       # fp32, no affine
        if self.mode == "fused":
            xq = rotation_utils.adaln_transform_rotate_quant_activation(ln, scale1, shift1, self.smooth_qkv[i], self.act_fp_type)
            qkv = F.linear(xq, b.attn.mat_qkv.weight) + b.attn.qkv_bias.to(xq.dtype)
            a = b.attn.proj(b.attn.attend(qkv))
        else:
            x1 = ln.mul(scale1.add(1)).add_(shift1)
            if self.Q is None:
                x1 = x1.half()
            else:
                x1 = torch.matmul(x1.mul(self.smooth_qkv[i]).half(), self.Q)              # autocast: fp16 GEMM (basic_var.py:263)
            a = b.attn(x1)
        x = x + a.mul_(gamma1)
        ln = F.layer_norm(x, (self.C,), eps=1e-6)
        if self.mode == "fused":
            xq = rotation_utils.adaln_transform_rotate_quant_activation(ln, scale2, shift2, self.smooth_fc1[i], self.act_fp_type)
            h1 = F.linear(xq, b.ffn.fc1.weight, b.ffn.fc1.bias.view(-1))
            if self.fc2_split_fmt is not None and h1.dtype == torch.float16:
                # GELU + the fc2 quantizer in one pass (fpq_gelu_fake_quant_signsplit), then the bare GEMM
                hq = ops.gelu_fake_quant_signsplit(h1, self.fc2_split_fmt, global_clip=True)
                f = F.linear(hq, b.ffn.fc2.weight, b.ffn.fc2.bias)
            else:
                f = b.ffn.fc2(F.gelu(h1, approximate="tanh"))
        else:
            x2 = ln.mul(scale2.add(1)).add_(shift2)
            if self.Q is None:
                x2 = x2.half()
            else:
                x2 = torch.matmul(x2.mul(self.smooth_fc1[i]).half(), self.Q)
            f = b.ffn(x2)
        return x + f.mul(gamma2)

    # -- generation ----------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, B, labels, rng, cfg=1.5, top_k=900, top_p=0.96):
        dev = self.lvl_1L.device
        pn, SN = self.patch_nums, len(self.patch_nums)
        cond = self.class_emb(torch.cat((labels, torch.full_like(labels, self.num_classes))))          # [2B, C] fp32
        cond16 = cond.half()
        shared = self.shared_ada_lin(F.silu(cond16)).view(-1, 1, 6, self.C) if self.shared_ada_lin is not None else None
        lvl_pos = self.lvl_embed(self.lvl_1L) + self.pos_1LC
        x = cond.unsqueeze(1) + self.pos_start + lvl_pos[:, :pn[0] ** 2]
        f_hat = cond.new_zeros(B, self.Cvae, pn[-1], pn[-1])
        for b in self.blocks:
            b.attn.reset_cache(2 * B, self.L, dev, self.kv_bit, self.kv_incremental)
        cur = 0
        for si, p in enumerate(pn):
            cur += p * p
            for i, b in enumerate(self.blocks):
                x = self._block(i, b, x, cond16, shared)
            sc, sh = self.head_ada(F.silu(cond)).view(-1, 1, 2, self.C).unbind(2)
            logits = self.head(F.layer_norm(x, (self.C,), eps=1e-6).mul(sc.add(1)).add_(sh))
            t = cfg * si / (SN - 1)
            logits = (1 + t) * logits[:B] - t * logits[B:]
            idx = sample_top_k_top_p(logits, top_k, top_p, rng)
            h = self.codebook(idx).transpose(1, 2).reshape(B, self.Cvae, p, p)
            phi = self.phi[min(3, round(si / (SN - 1) * 3))]
            if si != SN - 1:
                h = F.interpolate(h, size=(pn[-1], pn[-1]), mode="bicubic")
            f_hat += h.mul(0.5) + phi(h).mul_(0.5)
            if si != SN - 1:
                nxt = F.interpolate(f_hat, size=(pn[si + 1], pn[si + 1]), mode="area")
                x = self.word_embed(nxt.view(B, self.Cvae, -1).transpose(1, 2)) + lvl_pos[:, cur:cur + pn[si + 1] ** 2]
                x = x.repeat(2, 1, 1)
        return f_hat


def sample_top_k_top_p(logits, top_k, top_p, rng):
    """helpers.py:6-19 semantics: keep the top_k logits, drop the low tail holding <= 1-top_p of the mass, multinomial."""
    B, l, V = logits.shape
    if top_k > 0:
        kth = logits.topk(top_k, dim=-1, sorted=False)[0].amin(dim=-1, keepdim=True)
        logits = logits.masked_fill(logits < kth, -torch.inf)
    if top_p > 0:
        srt, order = logits.sort(dim=-1, descending=False)
        drop = srt.softmax(dim=-1).cumsum_(dim=-1) <= (1 - top_p)
        drop[..., -1:] = False
        logits = logits.masked_fill(drop.scatter(-1, order, drop), -torch.inf)
    return torch.multinomial(logits.softmax(dim=-1).view(-1, V), 1, replacement=True, generator=rng).view(B, l)


# ------------------------------------------------------------------------------------------ VQVAE decoder
class Res(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.n1, self.c1 = nn.GroupNorm(32, cin, eps=1e-6), nn.Conv2d(cin, cout, 3, padding=1)
        self.n2, self.c2 = nn.GroupNorm(32, cout, eps=1e-6), nn.Conv2d(cout, cout, 3, padding=1)
        self.skip = nn.Conv2d(cin, cout, 1) if cin != cout else nn.Identity()

    def forward(self, x):
        h = self.c1(F.silu(self.n1(x)))
        return self.skip(x) + self.c2(F.silu(self.n2(h)))


class SpatialAttn(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.n, self.qkv, self.o = nn.GroupNorm(32, c, eps=1e-6), nn.Conv2d(c, 3 * c, 1), nn.Conv2d(c, c, 1)

    def forward(self, x):
        B, C, H, W = x.shape
        q, k, v = self.qkv(self.n(x)).view(B, 3, C, H * W).transpose(2, 3).unbind(1)        # [B, HW, C]
        o = F.scaled_dot_product_attention(q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1)).squeeze(1)
        return x + self.o(o.transpose(1, 2).reshape(B, C, H, W))


class Decoder(nn.Module):
    def __init__(self, ch=160, mult=(1, 1, 2, 2, 4), zc=32):
        super().__init__()
        c = ch * mult[-1]
        self.post_quant = nn.Conv2d(zc, zc, 3, padding=1)
        self.inp = nn.Conv2d(zc, c, 3, padding=1)
        layers = [Res(c, c), SpatialAttn(c), Res(c, c)]
        for lvl in reversed(range(len(mult))):
            co = ch * mult[lvl]
            for _ in range(3):
                layers.append(Res(c, co))
                c = co
                if lvl == len(mult) - 1:
                    layers.append(SpatialAttn(c))
            if lvl:
                layers += [nn.Upsample(scale_factor=2, mode="nearest"), nn.Conv2d(c, c, 3, padding=1)]
        self.body = nn.Sequential(*layers)
        self.norm, self.out = nn.GroupNorm(32, c, eps=1e-6), nn.Conv2d(c, 3, 3, padding=1)

    @torch.no_grad()
    def forward(self, f_hat, chunk=10):                               # chunked: ~2 GB of activations per image at 256x256
        outs = []
        with torch.autocast("cuda", dtype=torch.float16):
            for part in f_hat.split(chunk):
                h = self.body(self.inp(self.post_quant(part)))
                outs.append(self.out(F.silu(self.norm(h))).float().clamp_(-1, 1).add_(1).mul_(0.5))
        return torch.cat(outs)


# ------------------------------------------------------------------------------------------ setup per mode
def prepare(model: Var, mode: str, bits: int, seed=0, rotate=True):
    """Quantize / transform / rotate the transformer in place for `mode`; weights end up fp16 on the GPU.
    rotate=False is the plain `models_fp_quant` configuration (BASELINE config 2): no GALT / rotation, fc2 on the
    symmetric grid; only `modules` / `refquant` / `fp16` apply there."""
    dev = model.lvl_1L.device
    model.mode = mode
    C = model.C
    act = "fp_e2" if bits == 4 else "fp6_e2m3"
    fc2 = ("fp_e1m2_neg_e2m1_pos" if bits == 4 else "fp6_int_neg_e2m3_pos") if rotate else act
    model.act_fp_type = act
    model.fc2_split_fmt = {"fp_e1m2_neg_e2m1_pos": "e1m2_neg_e2m1_pos", "fp6_int_neg_e2m3_pos": "int_neg_e2m3_pos"}.get(fc2) if mode == "fused" else None
    if mode != "fp16" and not rotate:
        if mode == "fused":
            raise SystemExit("mode fused needs --rotate (it fuses the online transform + rotation)")
        quant_utils.quantize_VAR(model, weight_quant="per_group", act_quant="per_group", w_bit=bits, a_bit=bits, act_quant_sym=True,
                                 activation_fp_quant=True, weight_fp_quant=True, act_fp_type=act, weight_fp_type=act, fc2_fp_type=fc2)
    elif mode != "fp16":
        g = torch.Generator(device="cpu").manual_seed(seed + 1)
        # GALT factors: the shipped best_lambda fixtures are not on the GPU box; log-normal around 1 like them
        # (best_lambda_var30/*.pt: mean 1.17-1.22, std of log 0.29-0.36, range 0.05-2.8)
        model.smooth_qkv = [torch.empty(C).normal_(0, 0.3, generator=g).exp().to(dev) for _ in range(model.depth)]
        model.smooth_fc1 = [torch.empty(C).normal_(0, 0.3, generator=g).exp().to(dev) for _ in range(model.depth)]
        transform_model_utils.transform_rotate_model(model, model.smooth_qkv, model.smooth_fc1)
        quant_utils.quantize_VAR(model, weight_quant="per_group", act_quant="per_group", w_bit=bits, a_bit=bits, act_quant_sym=True,
                                 activation_fp_quant=True, weight_fp_quant=True, act_fp_type=act, weight_fp_type=act, fc2_fp_type=fc2)
        model.Q = rotation_utils.block_random_hadamard_matrix(C, 128, dev, 42).half()
    for b in model.blocks:                                           # fp16 weights for every mode
        for lin in (b.attn.mat_qkv, b.attn.proj, b.ffn.fc1, b.ffn.fc2):
            for name in ("weight", "bias"):
                t = getattr(lin, name)
                if t is not None:
                    t = t.detach().half()
                    setattr(lin, name, nn.Parameter(t, requires_grad=False) if name in lin._parameters else t)
        if not b.shared_aln:
            b.ada_lin.half()
    if model.shared_ada_lin is not None:
        model.shared_ada_lin.half()
    if mode == "refquant":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import ref_glue
        ext = ref_glue.load_ref_ext()
        if ext is None:
            raise SystemExit("refquant: oracle/_ref/ref_quant_cuda*.so is missing (run __graft_entry__.build() where /root/reference exists)")
        grids = {"fp_e2": quant_utils.fp4_e2m1_grid, "fp6_e2m3": quant_utils.fp6_e2m3_grid}
        sym = partial(ref_glue.sym_group_cuda, ext.quant, grid=grids[act].float(), out_dtype=torch.float16 if bits == 6 else None)
        if not rotate:
            split = sym
        elif bits == 4:
            neg, pos = torch.tensor([-1.75, -1.5, -1.25, -1.0, -0.75, -0.5, -0.25, 0.0]), torch.tensor([0, 0.5, 1, 1.5, 2, 3, 4, 6.0])
            split = partial(ref_glue.signsplit_group_cuda, ext.quant, grid_neg=neg, grid_pos=pos, clipping_strength=1.0)
        else:
            split = partial(ref_glue.signsplit_group_cuda, ext.quant, grid_neg=quant_utils.int_neg_grid.float(),
                            grid_pos=quant_utils.e2m3_pos_grid.float(), clipping_strength=None)
        for b in model.blocks:
            b.attn.mat_qkv.act_quant = b.attn.proj.act_quant = b.ffn.fc1.act_quant = sym
            b.ffn.fc2.act_quant = split
    return model


def measure(dev, depth, batch, res=256, bits=4, mode="fused", iters=3, warmup=1, rank=0, world=1, decode=True, rotate=True, dec=None,
            quant_kv=""):
    """Build the model for `mode`, run `warmup` + `iters` generation passes; returns the JSON-able result (times are
    CUDA-event ms per batch on this rank; the caller takes the max over ranks)."""
    patch = PATCH_256 if res == 256 else PATCH_512
    torch.backends.cuda.matmul.allow_tf32 = True                      # evaluate_fp_quant_transform_rotate.py:172-175
    torch.backends.cudnn.allow_tf32 = True
    if decode and dec is None:
        with torch.device(dev):
            dec = Decoder().eval()
        for _ in range(2):                                            # settle cuDNN's per-shape choices before timing
            dec(torch.zeros(batch, 32, patch[-1], patch[-1], device=dev))
    torch.manual_seed(0)
    with torch.device(dev):
        model = Var(depth, patch, shared_aln=(depth == 36)).eval()
    model.init_weights(0)
    prepare(model, mode, bits, rotate=rotate)
    if quant_kv and mode != "fp16":                                   # --quant_kv: kv_bit follows the activation bits
        model.kv_bit, model.kv_incremental = bits, quant_kv == "incremental"
    rng = torch.Generator(device=dev)
    marks = []
    torch.cuda.reset_peak_memory_stats(dev)

    def one(it):
        labels = torch.full((batch,), (rank + world * it) % 1000, device=dev)     # class i -> GPU i mod G
        rng.manual_seed(0)
        f_hat = model.generate(batch, labels, rng)
        if not decode:
            return f_hat
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        img = dec(f_hat)
        b.record()
        marks.append((a, b))
        return img

    for it in range(warmup):
        img = one(it)
    torch.cuda.synchronize(dev)
    n1 = ops.launch_count()
    marks.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for it in range(iters):
        img = one(warmup + it)
    e1.record()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / iters
    res_d = {"mode": mode, "ms_per_batch": ms, "decode_ms_per_batch": sum(a.elapsed_time(b) for a, b in marks) / iters,
             "wall_ms_per_batch": wall / iters * 1e3, "fpq_launches_per_batch": (ops.launch_count() - n1) // max(1, iters),
             "finite": bool(torch.isfinite(img).all()), "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 2**30, 1),
             "config": {"depth": depth, "width": 64 * depth, "res": res, "batch_per_gpu": batch, "bits": bits, "rotate_transform": rotate, "quant_kv": quant_kv or None,
                        "patch_nums": list(patch), "decode": decode, "iters": iters, "warmup": warmup}}
    del model
    torch.cuda.empty_cache()
    return res_d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=30)
    ap.add_argument("--batch", type=int, default=50)
    ap.add_argument("--res", type=int, default=256, choices=(256, 512))
    ap.add_argument("--bits", type=int, default=4, choices=(4, 6))
    ap.add_argument("--mode", default="fused", help="comma list of fp16,modules,fused,refquant")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-rotate", action="store_true", help="plain models_fp_quant configuration (BASELINE config 2)")
    ap.add_argument("--quant-kv", default="", choices=("", "incremental", "reference"),
                    help="fake-quantize the KV cache (kv_bit = --bits): each row once, or the reference's schedule (whole cache every scale)")
    args = ap.parse_args()

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = True                      # evaluate_fp_quant_transform_rotate.py:172-175
    torch.backends.cudnn.allow_tf32 = True
    dec = None
    if not args.no_decode:
        with torch.device(dev):
            dec = Decoder().eval()
        for _ in range(2):
            dec(torch.zeros(args.batch, 32, (PATCH_256 if args.res == 256 else PATCH_512)[-1], (PATCH_256 if args.res == 256 else PATCH_512)[-1], device=dev))
    for mode in args.mode.split(","):
        if world > 1:
            dist.barrier()
        r = measure(dev, args.depth, args.batch, args.res, args.bits, mode, args.iters, args.warmup, rank, world,
                    decode=not args.no_decode, rotate=not args.no_rotate, dec=dec, quant_kv=args.quant_kv)
        ms = torch.tensor([r["ms_per_batch"]], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        r["ms_per_batch"] = float(ms)
        if rank == 0:
            line = {"metric": "images/sec, class-conditional VAR generation (autoregressive pass + VQVAE decode), random-init weights",
                    "mode": mode, "value": round(world * args.batch / (r["ms_per_batch"] / 1e3), 2), "unit": "images/s", "n_gpus": world}
            line.update({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items() if k != "mode"})
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
