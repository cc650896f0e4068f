"""torch + cuBLAS + flash-SDPA GPT-2 small (random init) used as the MEASURED library baseline for
the GPT-2 inference config (bs=32, seq=1024): prefill and KV-cached greedy decode, bf16."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class Block(nn.Module):
    def __init__(self, E, H):
        super().__init__()
        self.ln1, self.ln2 = nn.LayerNorm(E), nn.LayerNorm(E)
        self.qkv, self.proj = nn.Linear(E, 3 * E), nn.Linear(E, E)
        self.fc1, self.fc2 = nn.Linear(E, 4 * E), nn.Linear(4 * E, E)
        self.H = H

    def forward(self, x, cache, pos):
        B, L, E = x.shape
        qkv = self.qkv(self.ln1(x)).view(B, L, 3, self.H, E // self.H)
        q, k, v = qkv.unbind(2)
        cache[0][:, pos:pos + L] = k
        cache[1][:, pos:pos + L] = v
        kk, vv = cache[0][:, :pos + L], cache[1][:, :pos + L]
        o = F.scaled_dot_product_attention(q.transpose(1, 2), kk.transpose(1, 2), vv.transpose(1, 2), is_causal=(L > 1))
        x = x + self.proj(o.transpose(1, 2).reshape(B, L, E))
        return x + self.fc2(F.gelu(self.fc1(self.ln2(x)), approximate="tanh"))


def _block_train(blk: "Block", x, p_drop: float):
    B, L, E = x.shape
    q, k, v = blk.qkv(blk.ln1(x)).view(B, L, 3, blk.H, E // blk.H).unbind(2)
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True)
    x = x + F.dropout(blk.proj(o.transpose(1, 2).reshape(B, L, E)), p_drop)
    return x + F.dropout(blk.fc2(F.gelu(blk.fc1(blk.ln2(x)), approximate="tanh")), p_drop)


class GPT2(nn.Module):
    def __init__(self, V=50257, E=768, H=12, layers=12, Lmax=1024):
        super().__init__()
        self.wte, self.wpe = nn.Embedding(V, E), nn.Embedding(Lmax, E)
        self.blocks = nn.ModuleList([Block(E, H) for _ in range(layers)])
        self.ln_f, self.head = nn.LayerNorm(E), nn.Linear(E, V)
        self.E, self.H, self.Lmax = E, H, Lmax

    def new_cache(self, B, dev, dt):
        return [(torch.zeros(B, self.Lmax, self.H, self.E // self.H, device=dev, dtype=dt),
                 torch.zeros(B, self.Lmax, self.H, self.E // self.H, device=dev, dtype=dt)) for _ in self.blocks]

    @torch.no_grad()
    def step(self, ids, caches, pos):
        B, L = ids.shape
        x = self.wte(ids) + self.wpe(torch.arange(pos, pos + L, device=ids.device))
        for b, c in zip(self.blocks, caches):
            x = b(x, c, pos)
        return self.head(self.ln_f(x[:, -1:])).argmax(-1).reshape(B)

    def forward_train(self, ids, p_drop: float = 0.1):
        """Training forward over the whole sequence -> logits [B, L, V] (flash SDPA, no cache)."""
        B, L = ids.shape
        x = F.dropout(self.wte(ids) + self.wpe(torch.arange(L, device=ids.device)), p_drop)
        for b in self.blocks:
            x = _block_train(b, x, p_drop)
        return self.head(self.ln_f(x))
