"""Model base classes carrying the contract the reference trainer relies on
(FoodRec/common/abstract_recommender.py:8-91, FoodRec/common/trainer.py:182-190,407,421,495):
epoch hooks, `calculate_loss`, `predict`, `full_sort_predict`, a parameter count in `str(model)`, and
-- for `GeneralRecommender` -- `n_users`, `n_items`, `batch_size`, `device` and the modality feature
tables `v_feat` / `t_feat` as fp32 tensors on the configured device."""
import numpy as np
import torch
from torch import nn


def _abstract(name):
    def method(self, interaction):
        raise NotImplementedError(f"{type(self).__name__}.{name}")
    method.__name__ = name
    return method


class AbstractRecommender(nn.Module):
    calculate_loss = _abstract("calculate_loss")        # batch dict -> loss tensor or tuple of loss tensors
    predict = _abstract("predict")                      # batch dict -> scores [batch]
    full_sort_predict = _abstract("full_sort_predict")  # batch dict -> scores over all items

    def pre_epoch_processing(self):
        return None

    def post_epoch_processing(self):
        return None

    def __str__(self):
        n_params = sum(int(np.prod(p.size())) for p in self.parameters())
        return f"{super().__str__()}\nTrainable parameters: {n_params}"


class GeneralRecommender(AbstractRecommender):
    def __init__(self, config, dataset):
        super().__init__()
        self.n_users, self.n_items = dataset.n_users, dataset.n_items
        self.batch_size, self.device = config["train_batch_size"], config["device"]
        use_features = config["is_multimodal_model"] and not config["end2end"]
        self.v_feat = self._feature_table(dataset.embImage) if use_features else None
        self.t_feat = self._feature_table(dataset.embText) if use_features else None

    def _feature_table(self, array):
        return torch.tensor(np.asarray(array, dtype=np.float32)).to(self.device)
