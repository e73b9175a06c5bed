"""Oracle: pruned-model shape semantics restated on HF-named state dicts (test infrastructure).

* ``prune_heads``   HF 4.7.0 ``PreTrainedModel.prune_heads`` / ``ViTAttention.prune_heads``
  (third-party, removed in the installed 5.5.0); call sites
  ``deit_pruning/vendor/nn_pruning_v1/nn_pruning/inference_model_patcher.py:86``,
  ``are_16_heads/run_classifier.py:41-47``.
* ``optimize_ffn_dense``  ``optimize_model(model, "dense")`` + ``SparseDimensionsLinear.create``
  (``.../inference_model_patcher.py:266-317,124-170``).
* ``heads_to_prune_from_thresholds``  ``BertHeadsPruner.get_pruned_heads`` (``...:22-77``).
* DSL parsers: ``h_{density}_d_{density}-...`` (``.../patch_coordinator.py:396-406``) and
  ``all_head{k}_ffn{r}`` / ``layerwise_h{k}-d{r}_...`` (``modeling/models/vit.py:77-97``).

Pinned against the vendored ``optimize_model`` (imported from /root/reference in
``tests/golden/make_golden.py``) through ``tests/golden/pruned_*.npz``.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

_QKV = ("query", "key", "value")


def prune_heads(sd: Dict[str, torch.Tensor], to_prune: Dict[int, Sequence[int]], head_size: int = 64):
    """Index-select the kept heads' rows of q/k/v and columns of out-proj, original order kept."""
    sd = dict(sd)
    for l, heads in to_prune.items():
        p = f"vit.encoder.layer.{l}.attention."
        n_heads = sd[p + "attention.query.weight"].shape[0] // head_size
        kept = [h for h in range(n_heads) if h not in set(heads)]
        idx = torch.cat([torch.arange(h * head_size, (h + 1) * head_size) for h in kept])
        for n in _QKV:
            sd[p + f"attention.{n}.weight"] = sd[p + f"attention.{n}.weight"][idx].clone()
            sd[p + f"attention.{n}.bias"] = sd[p + f"attention.{n}.bias"][idx].clone()
        sd[p + "output.dense.weight"] = sd[p + "output.dense.weight"][:, idx].clone()
    return sd


def optimize_ffn_dense(sd: Dict[str, torch.Tensor]):
    """Cross-zero then drop all-zero FC1 rows / FC2 columns; keep at least one row."""
    sd = dict(sd)
    l = 0
    while f"vit.encoder.layer.{l}.intermediate.dense.weight" in sd:
        k1 = f"vit.encoder.layer.{l}.intermediate.dense."
        k2 = f"vit.encoder.layer.{l}.output.dense."
        w1, b1, w2 = sd[k1 + "weight"].clone(), sd[k1 + "bias"].clone(), sd[k2 + "weight"].clone()
        out_mask = w1.abs().sum(1) == 0          # inference_model_patcher.py:283
        in_mask = w2.abs().sum(0) == 0           # :286
        w1[in_mask] = 0                          # :288
        w2[:, out_mask] = 0                      # :289
        # SparseDimensionsLinear.get_sparsity (:108-122): keep rows with any non-zero weight
        idx1 = ((w1 != 0).sum(1) != 0).nonzero().squeeze(-1)
        if idx1.numel() == 0:
            idx1 = torch.tensor([0])
        idx2 = ((w2 != 0).sum(0) != 0).nonzero().squeeze(-1)
        if idx2.numel() == 0:
            idx2 = torch.tensor([0])
        sd[k1 + "weight"] = w1[idx1].clone()
        sd[k1 + "bias"] = b1[idx1].clone()
        sd[k2 + "weight"] = w2[:, idx2].clone()
        l += 1
    return sd


def parse_layerwise_thresholds(s: str) -> List[Dict[str, float]]:
    """'h_0.50_d_0.3-h_0.668_d_0.9-...' -> [{'head':0.5,'dense':0.3}, ...] (patch_coordinator.py:396-406)."""
    out = []
    for tok in s.split("-"):
        parts = tok.split("_")
        assert parts[0] == "h" and parts[2] == "d", tok
        out.append({"head": float(parts[1]), "dense": float(parts[3])})
    return out


def parse_prune_encoding(enc: str, depth: int, mlp_dim: int) -> Tuple[List[int], List[int]]:
    """'all_head12_ffn1.0' | 'layerwise_h2-d1.0_h3-d0.5_...' -> (heads[], inter[]) (modeling/models/vit.py:60-97)."""
    toks = enc.split("_")
    assert toks[0] in ("layerwise", "all")
    if toks[0] == "all":
        k = int(toks[1].replace("head", ""))
        r = float(toks[2].replace("ffn", ""))
        return [k] * depth, [int(r * mlp_dim)] * depth
    heads, inter = [], []
    for t in toks[1:]:
        hx, dx = t.split("-")
        heads.append(int(hx.replace("h", "")))
        inter.append(int(float(dx.replace("d", "")) * mlp_dim))
    assert len(heads) == depth
    return heads, inter


def heads_to_prune_from_thresholds(sd, thresholds: List[Dict[str, float]], num_heads: int, head_size: int = 64):
    """BertHeadsPruner.get_pruned_heads: score = #q/k/v blocks with any non-zero, prune the lowest."""
    to_prune = {}
    for l, thr in enumerate(thresholds):
        score = torch.zeros(num_heads, dtype=torch.int32)
        for n in _QKV:
            w = sd[f"vit.encoder.layer.{l}.attention.attention.{n}.weight"]
            score += (w != 0).reshape(num_heads, head_size, w.shape[1]).any(-1).any(-1).int()
        n_prune = num_heads - int(thr["head"] * num_heads)
        _, order = torch.sort(score)
        heads = sorted(int(i) for i in order[:n_prune])
        if len(heads) == num_heads:
            heads.remove(0)
        to_prune[l] = heads
    return to_prune


def synthesize_pruned(sd, heads_kept: Sequence[Sequence[int]], inter_kept: Sequence[int], seed: int = 7,
                      head_size: int = 64):
    """Build (full-size-with-zeros, physically-pruned) state-dict pair from an unpruned one.

    heads_kept[l] = list of surviving head indices; inter_kept[l] = number of surviving FFN rows
    (chosen at random with ``seed``).  The full-size dict is what ``deit_pruning`` writes to disk
    (zero rows / cols); the pruned dict is what ``optimize_model`` + ``prune_heads``
    yield at eval time."""
    g = torch.Generator().manual_seed(seed)
    full = {k: v.clone() for k, v in sd.items()}
    to_prune = {}
    for l, (hk, ik) in enumerate(zip(heads_kept, inter_kept)):
        p = f"vit.encoder.layer.{l}."
        n_heads = full[p + "attention.attention.query.weight"].shape[0] // head_size
        dead = [h for h in range(n_heads) if h not in set(hk)]
        to_prune[l] = dead
        for h in dead:
            sl = slice(h * head_size, (h + 1) * head_size)
            for n in _QKV:
                full[p + f"attention.attention.{n}.weight"][sl] = 0
                full[p + f"attention.attention.{n}.bias"][sl] = 0
            full[p + "attention.output.dense.weight"][:, sl] = 0
        I = full[p + "intermediate.dense.weight"].shape[0]
        drop = torch.randperm(I, generator=g)[: I - ik]
        full[p + "intermediate.dense.weight"][drop] = 0
        full[p + "intermediate.dense.bias"][drop] = 0
        full[p + "output.dense.weight"][:, drop] = 0
    pruned = optimize_ffn_dense(prune_heads(full, to_prune, head_size))
    return full, pruned, to_prune


def pruned_head_str_to_dict(s: str) -> Dict[int, List[int]]:
    """'1:2,3 12:1 ...' (1-based layers and heads) as in draw.py:88-95."""
    rv = {}
    for item in s.split():
        k, v = item.split(":")
        rv[int(k)] = [int(x) for x in v.split(",")]
    return rv


# are16heads DeiT-Tiny, 18 heads pruned (draw.py:104-106); kept heads per layer = [1,1,1,1,2,1,2,2,2,2,1,2]
DEIT_TINY_HEAD18 = "1:2,3 12:1 2:1,3 6:1,3 3:1,2 11:1,2 7:3 10:2 9:2 4:2,3 5:3 8:1"


def kept_heads_from_pruned_str(s: str, layers: int, n_heads: int) -> List[List[int]]:
    d = pruned_head_str_to_dict(s)
    return [[h for h in range(n_heads) if (h + 1) not in d.get(l + 1, [])] for l in range(layers)]
