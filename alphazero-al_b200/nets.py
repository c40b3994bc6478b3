"""Evaluator stand-ins with the reference's CNN I/O contract (the networks themselves are OUT OF SCOPE - SURVEY.md
section 2 row 16 - PyTorch stays the evaluator; only the contract matters to the hot path).

``C4Net`` is an independent PyTorch implementation with the shape of the reference's Connect4 network
(src/environments/Connect4/Network.py:153-253: piece + mirror-orbit position embedding (32) -> 3x3 conv stem (64) ->
3 GroupNorm/SiLU residual conv blocks -> gated multi-head self-attention over the 42 cells -> column policy head and a
WDL + moves-left head; 160 358 parameters) used for random-init throughput runs and for CNN-in-the-loop tests.

Two entry points:
* ``predict(state, action_mask=None)`` - the reference contract (src/environments/Connect4/Network.py:267-288):
  numpy/CPU tensors in, numpy ``(probs[B,A], wdl_rel[B,3], aux[B,1])`` out (bf16 autocast on CUDA, like the reference).
* ``predict_device(planes, action_mask)`` - the device contract of ``batched_mcts.BatchedMCTS``: CUDA tensors in/out,
  no synchronisation, no host copy.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

R, C, CELLS = 6, 7, 42


def _mirror_orbits():
    """cell -> orbit id under the left-right mirror (24 orbits: 6 rows x 4 distinct columns)."""
    return torch.tensor([r * 4 + min(c, C - 1 - c) for r in range(R) for c in range(C)], dtype=torch.long)


class _ConvBlock(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.norm = nn.GroupNorm(1, ch)
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return x + F.silu(self.conv(self.norm(x)))


class _GatedSelfAttention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.h, self.hd = heads, dim // heads
        self.pre = nn.RMSNorm(dim, eps=1e-5)
        self.qkv = nn.Linear(dim, 3 * dim, bias=False)
        self.gate = nn.Linear(dim, heads, bias=False)
        self.out = nn.Linear(dim, dim, bias=False)
        self.qn = nn.RMSNorm(self.hd, eps=1e-5)
        self.kn = nn.RMSNorm(self.hd, eps=1e-5)

    def forward(self, x):                                   # x: [B, 42, dim]
        B, L, D = x.shape
        y = self.pre(x)
        q, k, v = self.qkv(y).view(B, L, 3, self.h, self.hd).unbind(2)
        g = torch.sigmoid(self.gate(y))                     # [B, L, heads]
        o = F.scaled_dot_product_attention(self.qn(q).transpose(1, 2), self.kn(k).transpose(1, 2), v.transpose(1, 2))
        o = (o * g.transpose(1, 2).unsqueeze(-1)).transpose(1, 2).reshape(B, L, D)
        return x + self.out(o)


class C4Net(nn.Module):
    aux_target_offset = 42
    n_actions = 7

    def __init__(self, embed=32, width=64, blocks=3, heads=4, device="cpu", max_batch=32768):
        super().__init__()
        self.device_str = device
        self.max_batch = max_batch
        self.piece = nn.Embedding(2, embed)
        self.pos = nn.Embedding(24, embed)
        self.register_buffer("orbits", _mirror_orbits())
        self.stem = nn.Conv2d(embed, width, 3, padding=1)
        self.body = nn.Sequential(*[_ConvBlock(width) for _ in range(blocks)])
        self.attn = _GatedSelfAttention(width, heads)
        # column policy head
        self.p_norm = nn.RMSNorm(width, eps=1e-5)
        self.p_row = nn.Linear(width, 1)
        self.p_fc = nn.Linear(width, width)
        self.p_out = nn.Linear(width, 1)
        # WDL + moves-left head
        self.v_pool_norm = nn.RMSNorm(width, eps=1e-5)
        self.v_pool_fc = nn.Linear(width, width)
        self.v_norm = nn.RMSNorm(width, eps=1e-5)
        self.v_fc = nn.Linear(width, width)
        self.v_out_norm = nn.RMSNorm(width, eps=1e-5)
        self.v_wdl = nn.Linear(width, 3)
        self.v_aux = nn.Linear(width, 1)
        for m in self.modules():
            if isinstance(m, nn.Embedding):
                nn.init.orthogonal_(m.weight)
            elif isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(m.weight, mode="fan_in", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        for head in (self.p_out, self.v_wdl, self.v_aux):        # like the reference: heads start at zero, so a random-init
            nn.init.zeros_(head.weight)                           # network predicts a uniform policy, WDL = 1/3 and 21 plies
        self.to(device)
        self.eval()

    def forward(self, planes, action_mask=None):
        """planes f32[B,3,6,7] -> (log_probs[B,7], log_wdl[B,3], moves_left_norm[B])."""
        B = planes.shape[0]
        own, opp = planes[:, 0].reshape(B, CELLS, 1), planes[:, 1].reshape(B, CELLS, 1)
        x = own * self.piece.weight[0] + opp * self.piece.weight[1] + self.pos(self.orbits)          # [B, 42, embed]
        x = x.transpose(1, 2).reshape(B, -1, R, C)
        x = self.body(F.silu(self.stem(x)))
        x = self.attn(x.flatten(2).transpose(1, 2))                                               # [B, 42, width]
        # policy: softmax-weighted pooling of each column's 6 cells, then a 2-layer MLP per column
        y = self.p_norm(x).reshape(B, R, C, -1).transpose(1, 2)                                   # [B, 7, 6, width]
        w = torch.softmax(self.p_row(y).squeeze(-1), dim=-1)
        col = (w.unsqueeze(-1) * y).sum(dim=2)
        logits = self.p_out(F.silu(self.p_fc(col))).squeeze(-1)
        if action_mask is not None:
            logits = logits.masked_fill(~action_mask.bool(), -1e9)
        logp = F.log_softmax(logits.float(), dim=-1)
        g = x.mean(dim=1)
        g = g + F.silu(self.v_pool_fc(self.v_pool_norm(g)))
        h = self.v_out_norm(F.silu(self.v_fc(self.v_norm(g))))
        return logp, F.log_softmax(self.v_wdl(h).float(), dim=-1), torch.sigmoid(self.v_aux(h).squeeze(-1).float())

    @torch.no_grad()
    def predict_device(self, planes, action_mask, autocast=True):
        """CUDA tensors f32[B,3,6,7], u8/bool[B,7] -> probs f32[B,7], wdl_rel f32[B,3], aux f32[B] (plies left)."""
        outs = ([], [], [])
        for i in range(0, planes.shape[0], self.max_batch):
            p, m = planes[i:i + self.max_batch], action_mask[i:i + self.max_batch]
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast and planes.is_cuda):
                logp, logw, ml = self(p, m)
            outs[0].append(logp.float().exp()); outs[1].append(logw.float().exp()); outs[2].append(ml.float() * self.aux_target_offset)
        return tuple(torch.cat(o) if len(o) > 1 else o[0] for o in outs)

    @torch.no_grad()
    def predict(self, state, action_mask=None):
        """Reference contract: numpy in, numpy (probs, wdl_rel, aux[B,1]) out."""
        dev = next(self.parameters()).device
        t = torch.as_tensor(np.asarray(state), dtype=torch.float32, device=dev)
        m = torch.ones((t.shape[0], 7), dtype=torch.bool, device=dev) if action_mask is None else \
            torch.as_tensor(np.asarray(action_mask), device=dev).bool().reshape(t.shape[0], 7)
        p, w, a = self.predict_device(t, m, autocast=getattr(self, "autocast", True))
        return p.cpu().numpy(), w.cpu().numpy(), a.reshape(-1, 1).cpu().numpy()
