"""Drop-in ``FrozenOpenCLIPEmbedder`` (model/clip.py:12-65): the OpenCLIP text tower that produces ``c_txt``.

The text tower is an INPUT of the accelerated path, not part of it (it runs once per batch on the constant
``""`` prompt, main/det/test_edtr.py:109,122): it is evaluated with plain PyTorch ops in the parameters' dtype, and the
embedding of a prompt list is cached per (prompts, weights version, device) — for EDTR's constant prompt the tower
runs once per process.  What must hold for the drop-in is the interface: same constructor
arguments, same state-dict keys (``model.token_embedding.weight``, ``model.positional_embedding``,
``model.transformer.resblocks.N.{ln_1, attn.in_proj_*, attn.out_proj, ln_2, mlp.c_fc, mlp.c_proj}``,
``model.ln_final``, ``model.text_projection``, ``model.logit_scale``) so that ``load_pretrained_sd`` fills it from
an SD-2.1 checkpoint (model/cldm.py:46-77), and ``encode(list[str]) -> [B, 77, width]``.

Tokenisation: the BPE vocabulary is a data file of the reference (model/open_clip/bpe_simple_vocab_16e6.txt.gz) and
is not shipped here.  The empty prompt needs no vocabulary (start / end token + padding); for any other text the
reference's tokenizer is used when its package is importable, or pass token ids (LongTensor [B, 77]) directly.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple, Union

import torch
import torch.nn.functional as F
from torch import nn

from .nets import _attach, state_version

SOT_TOKEN, EOT_TOKEN = 49406, 49407   # model/open_clip/tokenizer.py: vocab["<start_of_text>"], vocab["<end_of_text>"]


class _Model(nn.Module):
    """Holds the parameters under the ``model.`` prefix of the reference embedder."""


class FrozenOpenCLIPEmbedder(nn.Module):
    LAYERS = ["last", "penultimate"]

    def __init__(self, embed_dim, vision_cfg, text_cfg, layer="last"):
        super().__init__()
        assert layer in self.LAYERS
        t = dict(text_cfg)
        self.context_length = int(t["context_length"])
        self.vocab_size = int(t["vocab_size"])
        self.width = int(t["width"])
        self.heads = int(t["heads"])
        self.layers = int(t["layers"])
        self.layer = layer
        self.layer_idx = 0 if layer == "last" else 1
        self.model = _Model()
        w = self.width
        shapes = [("positional_embedding", (self.context_length, w)), ("text_projection", (w, int(embed_dim))),
                  ("logit_scale", ())]
        for i in range(self.layers):
            p = f"transformer.resblocks.{i}."
            shapes += [(p + "ln_1.weight", (w,)), (p + "ln_1.bias", (w,)),
                       (p + "attn.in_proj_weight", (3 * w, w)), (p + "attn.in_proj_bias", (3 * w,)),
                       (p + "attn.out_proj.weight", (w, w)), (p + "attn.out_proj.bias", (w,)),
                       (p + "ln_2.weight", (w,)), (p + "ln_2.bias", (w,)),
                       (p + "mlp.c_fc.weight", (4 * w, w)), (p + "mlp.c_fc.bias", (4 * w,)),
                       (p + "mlp.c_proj.weight", (w, 4 * w)), (p + "mlp.c_proj.bias", (w,))]
        shapes += [("token_embedding.weight", (self.vocab_size, w)), ("ln_final.weight", (w,)), ("ln_final.bias", (w,))]
        for k, shp in shapes:
            _attach(self.model, k, shp)
        self._init_parameters()
        self._cache: Dict[Tuple, torch.Tensor] = {}

    @torch.no_grad()
    def _init_parameters(self) -> None:
        """model/open_clip/model.py (CLIP.init_parameters): the values only matter until a checkpoint is loaded."""
        sd = dict(self.model.named_parameters())
        w, L = self.width, self.layers
        nn.init.normal_(sd["token_embedding.weight"], std=0.02)
        nn.init.normal_(sd["positional_embedding"], std=0.01)
        sd["logit_scale"].fill_(math.log(1 / 0.07))
        nn.init.normal_(sd["text_projection"], std=w ** -0.5)
        proj_std, attn_std, fc_std = (w ** -0.5) * ((2 * L) ** -0.5), w ** -0.5, (2 * w) ** -0.5
        for k, p in sd.items():
            if k.endswith(("ln_1.weight", "ln_2.weight", "ln_final.weight")):
                p.fill_(1.0)
            elif k.endswith("bias"):          # incl. attn.in_proj_bias
                p.zero_()
            elif k.endswith("attn.in_proj_weight"):
                nn.init.normal_(p, std=attn_std)
            elif k.endswith(("attn.out_proj.weight", "mlp.c_proj.weight")):
                nn.init.normal_(p, std=proj_std)
            elif k.endswith("mlp.c_fc.weight"):
                nn.init.normal_(p, std=fc_std)

    # ------------------------------------------------------------------ tokens
    def tokenize(self, text: Sequence[str]) -> torch.Tensor:
        """[B, context_length] int64 token ids (model/open_clip/tokenizer.py:tokenize)."""
        if all(s == "" for s in text):
            tok = torch.zeros((len(text), self.context_length), dtype=torch.long)
            tok[:, 0], tok[:, 1] = SOT_TOKEN, EOT_TOKEN
            return tok
        try:
            from model.open_clip import tokenize as ref_tokenize   # the reference package, when it is on sys.path
        except ImportError as exc:
            raise NotImplementedError(
                "edtr_b200.clip has no BPE vocabulary: only the empty prompt (EDTR's default_prompt) can be tokenised "
                "without the reference's model.open_clip package; pass token ids [B, 77] instead") from exc
        return ref_tokenize(list(text), self.context_length)

    # ------------------------------------------------------------------ text tower
    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        return self.encode_with_transformer(tokens)

    def encode_with_transformer(self, tokens: torch.Tensor) -> torch.Tensor:
        """model/clip.py:41-55: embedding + positions -> all but the last `layer_idx` residual blocks (causal mask)
        -> ln_final.  [B, 77] -> [B, 77, width]."""
        m = dict(self.model.named_parameters())
        w, h = self.width, self.heads
        x = F.embedding(tokens, m["token_embedding.weight"]) + m["positional_embedding"]
        B, L, _ = x.shape
        mask = torch.full((L, L), float("-inf"), device=x.device, dtype=x.dtype).triu_(1)   # model.py:build_attention_mask
        for i in range(self.layers - self.layer_idx):
            p = f"transformer.resblocks.{i}."
            y = F.layer_norm(x, (w,), m[p + "ln_1.weight"], m[p + "ln_1.bias"], 1e-5)
            q, k, v = F.linear(y, m[p + "attn.in_proj_weight"], m[p + "attn.in_proj_bias"]).chunk(3, dim=-1)
            q, k, v = (t.view(B, L, h, w // h).transpose(1, 2) for t in (q, k, v))
            a = torch.softmax(q @ k.transpose(-1, -2) * (w // h) ** -0.5 + mask, dim=-1) @ v
            a = a.transpose(1, 2).reshape(B, L, w)
            x = x + F.linear(a, m[p + "attn.out_proj.weight"], m[p + "attn.out_proj.bias"])
            y = F.layer_norm(x, (w,), m[p + "ln_2.weight"], m[p + "ln_2.bias"], 1e-5)
            y = F.gelu(F.linear(y, m[p + "mlp.c_fc.weight"], m[p + "mlp.c_fc.bias"]))
            x = x + F.linear(y, m[p + "mlp.c_proj.weight"], m[p + "mlp.c_proj.bias"])
        return F.layer_norm(x, (w,), m["ln_final.weight"], m["ln_final.bias"], 1e-5)

    @torch.no_grad()
    def encode(self, text: Union[List[str], torch.Tensor]) -> torch.Tensor:
        """model/clip.py:57-63.  Results are cached per (prompts, parameter version, device): with EDTR's constant
        prompt the tower runs once; the returned tensor is a fresh clone, so callers may modify it."""
        dev = next(self.model.parameters()).device
        if torch.is_tensor(text):
            return self(text.to(dev))
        key = (tuple(text), state_version(self.model), str(dev))
        hit = self._cache.get(key)
        if hit is None:
            uniq = sorted(set(text))
            emb = self(self.tokenize(uniq).to(dev))
            table = {s: emb[i] for i, s in enumerate(uniq)}
            hit = torch.stack([table[s] for s in text], 0)
            if len(self._cache) > 8:
                self._cache.clear()
            self._cache[key] = hit
        out = hit.clone()
        # Tag for the engine's cross-attention K/V cache (engine.CldmEngine.sample): the projected K/V of all 23 + 23
        # blocks only depend on c_txt, so a batch whose c_txt carries the tag of the previous batch skips the two
        # projection GEMMs.  The tag is void as soon as the tensor is written to (its version counter moves).
        out._edtr_ctx_key = (key, out._version)
        return out
