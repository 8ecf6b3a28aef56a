"""Shared plumbing of the dot-product recommenders (CLUSSL, HealthRec, LightGCN)."""
from __future__ import annotations

import torch

from .. import ops
from ..common.abstract_recommender import GeneralRecommender


class DotProductRecommender(GeneralRecommender):
    """Adds what the reference's dot-product models share: candidate scoring (`inference_fast` /
    `inference_by_user`, FoodRec/models/cikm_model.py:283-302) and full-sort scoring over every item
    for any number of users per batch.  Sub-classes implement `_propagate_all()`
    returning the propagated `[n_users + n_items, d]` table plus model-specific extras."""

    def train(self, mode: bool = True):
        """Every mode switch drops the cached evaluation tables: the trainer calls `model.train()` /
        `model.eval()` once per epoch / evaluation (FoodRec/common/trainer.py:156,233,478), so a cache
        can never outlive the parameters it was computed from."""
        self._eval_cache = None
        return super().train(mode)

    def _param_stamp(self):
        """Identity of the current parameter values.  `_version` / `data_ptr` see ATen updates only; a
        CUDA-graph replay (`train.GraphedTrainStep`) changes the values without touching either, so the
        step drivers also bump `_param_generation` (`train.mark_parameters_updated`)."""
        return (getattr(self, "_param_generation", 0),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _tables(self):
        """Propagated (user_all, item_all); cached across calls while in eval mode under no_grad
        and the parameters are unchanged (the reference's by-user / full-sort loops call the model
        once per user)."""
        if self.training or torch.is_grad_enabled():
            all_emb = self._propagate_all()[0]
            return all_emb[:self.n_users], all_emb[self.n_users:]
        stamp = self._param_stamp()
        cache = getattr(self, "_eval_cache", None)
        if cache is None or cache[0] != stamp:
            all_emb = self._propagate_all()[0]
            cache = (stamp, all_emb[:self.n_users], all_emb[self.n_users:])
            self._eval_cache = cache
        return cache[1], cache[2]

    def inference_by_user(self, batch_data):
        user_all, item_all = self._tables()
        return ops.pair_scores(user_all, item_all, batch_data["user_input"], batch_data["item_input"])

    def inference_fast(self, batch_data, user_emb, item_emb):
        return ops.pair_scores(user_emb, item_emb, batch_data["user_input"], batch_data["item_input"])

    def full_sort_predict(self, batch_data):
        """Scores of the batch users against all items, `[n_batch_users, n_items]` (squeezed to
        `[n_items]` for a single user, the shape `Trainer.evaluate` feeds `torch.topk`,
        FoodRec/common/trainer.py:495-497).  Use `foodrec_b200.evaluation.full_sort_topk` for the
        fused score+mask+top-K path that never materialises this matrix."""
        from .. import evaluation
        user_all, item_all = self._tables()
        s = evaluation.full_sort_scores(user_all, item_all, batch_data["u_id"])
        return s[0] if s.shape[0] == 1 else s
