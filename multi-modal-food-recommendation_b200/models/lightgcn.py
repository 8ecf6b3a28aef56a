"""LightGCN on the B200 kernels -- drop-in for FoodRec/models/lightgcn.py (items enter the graph
as `image_trs(image_embedding.weight)` where the table is initialised from the TEXT features,
lightgcn.py:73-74,122-132; `item_embedding` exists and is only used by the regulariser)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import graph as G
from .. import ops
from ..common.init import xavier_uniform_initialization
from ..common.loss import BPRLoss, EmbLoss
from ._base import DotProductRecommender


class LightGCN(DotProductRecommender):
    def __init__(self, config, dataset):
        super().__init__(config, dataset)
        self.config = config
        self.dataset = dataset
        self.interaction_matrix = dataset.train_coo_matrix
        self.latent_dim = config["embedding_size"]
        self.n_layers = config["n_layers"]
        self.reg_weight = config["reg_weight"]
        self.user_embedding = nn.Embedding(self.n_users, self.latent_dim)
        self.item_embedding = nn.Embedding(self.n_items, self.latent_dim)
        self.mf_loss = BPRLoss()
        self.reg_loss = EmbLoss()
        self.restore_user_e = None
        self.restore_item_e = None
        self.g_ui = G.norm_adj_user_item(self.interaction_matrix, self.n_users, self.n_items, self.device)
        self.apply(xavier_uniform_initialization)
        self.other_parameter_name = ["restore_user_e", "restore_item_e"]
        self.image_embedding = nn.Embedding.from_pretrained(self.t_feat, freeze=False)
        self.image_trs = nn.Linear(self.t_feat.shape[1], self.latent_dim)

    def get_ego_embeddings(self):
        return torch.cat([self.user_embedding.weight, self.image_trs(self.image_embedding.weight)], dim=0)

    def _propagate_all(self):
        return (ops.propagate_mean(self.g_ui, self.user_embedding.weight, self.n_layers,
                                   bottom=self.image_trs(self.image_embedding.weight)),)

    def forward(self):
        all_emb = self._propagate_all()[0]
        return all_emb[:self.n_users], all_emb[self.n_users:]

    def calculate_loss(self, batch_data):
        user, pos_item, neg_item = batch_data["u_id"], batch_data["pos_i_id"], batch_data["neg_i_id"]
        all_emb = self._propagate_all()[0]
        uw, iw = self.user_embedding.weight, self.item_embedding.weight
        mf_loss, reg = ops.rank_loss(all_emb, self.n_users, user, pos_item, neg_item,
                                     [(uw, user), (iw, pos_item), (iw, neg_item)],
                                     reg_den=float(neg_item.shape[0]), gamma=self.mf_loss.gamma)
        return mf_loss, (self.reg_weight * reg).reshape(1)
