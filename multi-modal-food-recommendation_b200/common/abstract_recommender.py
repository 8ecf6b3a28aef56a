"""Base classes with the reference's model contract (FoodRec/common/abstract_recommender.py:8-91):
what `Trainer` calls (`pre/post_epoch_processing`, `calculate_loss`, `full_sort_predict`) and what
`GeneralRecommender.__init__` provides (`n_users`, `n_items`, `batch_size`, `device`, `v_feat`,
`t_feat` loaded from `dataset.embImage/embText` as fp32 device tensors)."""
import numpy as np
import torch
import torch.nn as nn


class AbstractRecommender(nn.Module):
    def pre_epoch_processing(self):
        return None

    def post_epoch_processing(self):
        return None

    def calculate_loss(self, interaction):
        raise NotImplementedError

    def predict(self, interaction):
        raise NotImplementedError

    def full_sort_predict(self, interaction):
        raise NotImplementedError

    def __str__(self):
        n = sum(int(np.prod(p.size())) for p in self.parameters())
        return super().__str__() + "\nTrainable parameters: {}".format(n)


class GeneralRecommender(AbstractRecommender):
    def __init__(self, config, dataset):
        super().__init__()
        self.n_users = dataset.n_users
        self.n_items = dataset.n_items
        self.batch_size = config["train_batch_size"]
        self.device = config["device"]
        self.v_feat = self.t_feat = None
        if not config["end2end"] and config["is_multimodal_model"]:
            self.v_feat = torch.tensor(np.asarray(dataset.embImage, dtype=np.float32)).to(self.device)
            self.t_feat = torch.tensor(np.asarray(dataset.embText, dtype=np.float32)).to(self.device)
