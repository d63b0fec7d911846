"""Test harness for the two-pass owner-computes graph stage (b200r_expand_rescore_sharded): combine the per-rank records the
way a host that all-gathers them would.  Plumbing only (fancy indexing), kept out of the product package."""
from typing import Dict

import torch

_AUX_KEYS = ("hop", "score_graph_norm", "score_dense_norm", "cos")
_AUX_FILL = {"hop": -1, "score_graph_norm": 0.0, "score_dense_norm": 0.0, "cos": 0.0}


def combine_shard_records(g: Dict[str, torch.Tensor], ncand: torch.Tensor, status: torch.Tensor, k: int):
    """g: {field: [G, B, kin]} per-rank outputs.  Merge (score, id) on the device (b200r_merge_topk_src), then pick each
    winner's other fields from the rank list and position the merge reports."""
    from a_modular_rag_framework_b200.index import merge_topk_src
    G, B, kin = (int(x) for x in g["score"].shape)
    ms, mi, src = merge_topk_src(g["score"], g["id"], k)
    valid = src >= 0
    s = src.clamp_min(0).long()
    gi, ji = s // kin, s % kin
    qi = torch.arange(B, device=s.device).unsqueeze(1).expand_as(s)
    out = {"score": ms, "id": mi}
    for key in _AUX_KEYS:
        picked = g[key][gi, qi, ji]
        out[key] = torch.where(valid, picked, torch.full_like(picked, _AUX_FILL[key]))
    out["ncand"], out["status"] = ncand, status
    return out
