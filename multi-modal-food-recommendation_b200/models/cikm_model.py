"""HealthRec (`CIKM_Model`) on the B200 kernels -- drop-in for FoodRec/models/cikm_model.py.

Hot path (propagation over the recipe-ingredient and user-item graphs, BPR / regulariser gathers,
candidate scoring) runs on the sm_100a kernels.  The ingredient transformer, the two target-attention
blocks and the health head are batch-sized dense torch modules outside the hot path (SURVEY.md
section 2); they are re-stated here with the reference's parameter names, shapes and creation order
so `state_dict`s interchange and a seed reproduces the reference initialisation.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import graph as G
from .. import ops
from ..common.init import xavier_uniform_initialization
from ..common.loss import BPRLoss, EmbLoss
from ._base import DotProductRecommender


class target_attention_layer(nn.Module):
    """Multi-head dot-product attention with per-head LayerNorm on queries and keys and no learned
    projection by default (FoodRec/models/cikm_model.py:311-369).  `q_fc/k_fc/v_fc` exist (and are
    checkpointed) even when `linear_projection` is False, as in the reference."""

    def __init__(self, model_dims, hidden, num_head, linear_projection, atten_mode, padding_idx):
        super().__init__()
        self.linear_projection = linear_projection
        self.num_split = int(hidden / num_head)
        self.num_head = num_head
        self.q_fc = nn.Linear(model_dims, hidden)
        self.k_fc = nn.Linear(model_dims, hidden)
        self.v_fc = nn.Linear(model_dims, hidden)
        self.atten_mode = atten_mode
        self.padding_idx = padding_idx
        if atten_mode == "ln":
            self.ln = nn.LayerNorm(self.num_split, eps=1e-12)

    def _heads(self, x):  # [B, L, H*dh] -> [H*B, L, dh], head-major like chunk(dim=2) + cat(dim=0)
        B, L, _ = x.shape
        return x.reshape(B, L, self.num_head, self.num_split).permute(2, 0, 1, 3).reshape(-1, L, self.num_split)

    def forward(self, target_query, item_vec, seq_ids=None):
        Q, K, V = target_query, item_vec, item_vec
        if self.linear_projection:
            Q, K, V = self.q_fc(Q), self.k_fc(K), self.v_fc(V)
        B, Lq, Lk = Q.shape[0], Q.shape[1], K.shape[1]
        q, k, v = self._heads(Q), self._heads(K), self._heads(V)
        if self.atten_mode == "ln":
            q, k = self.ln(q), self.ln(k)
        att = torch.matmul(q, k.transpose(1, 2)) * (self.num_split ** -0.5)
        if seq_ids is not None:
            pad = (seq_ids == self.padding_idx).float().view(-1, 1, Lk).repeat(self.num_head, Lq, 1)
            att = (1.0 - pad) * att + pad * float(-2 ** 32 + 1)
        att = torch.softmax(att, dim=-1)
        out = torch.matmul(att, v)  # [H*B, Lq, dh]
        out = out.reshape(self.num_head, B, Lq, self.num_split).permute(1, 2, 0, 3).reshape(B, Lq, -1)
        return out.squeeze(), att


class CIKM_Model(DotProductRecommender):
    def __init__(self, config, dataset):
        super().__init__(config, dataset)
        self.config = config
        self.dataset = dataset
        self.n_ingredients = dataset.num_ingredients
        self.n_cal_level = dataset.num_calories_level
        self.n_health_level = (len(dataset.health_level_multi_hot[0]) if config["use_health_level_multi_hot"]
                               else dataset.num_health_level)
        d = config["embedding_size"]
        self.encoder_layer = nn.TransformerEncoderLayer(
            d_model=d, nhead=config["num_attention_heads"], dim_feedforward=4 * d,
            dropout=config["attention_probs_dropout_prob"], activation=config["hidden_act"])
        self.ingr_encoder = nn.TransformerEncoder(self.encoder_layer, num_layers=config["num_hidden_layers"])
        self.mm_target_atten = target_attention_layer(d, d, config["num_attention_heads"], linear_projection=False,
                                                      atten_mode="ln", padding_idx=self.n_ingredients)
        self.ingre_target_atten = target_attention_layer(d, d, config["num_attention_heads"],
                                                         linear_projection=False, atten_mode="ln",
                                                         padding_idx=self.n_ingredients)
        self.health_mlp = nn.Sequential(nn.Linear(d, d), nn.ReLU(), nn.Linear(d, self.n_health_level))
        self.criterion = nn.BCELoss(reduction="none")

        self.interaction_matrix = dataset.train_coo_matrix
        self.latent_dim = d
        self.n_layers = config["n_layers"]
        self.ui_layers = config["ui_layers"]
        self.reg_weight = config["reg_weight"]
        self.loss_kd = config["loss_kd"]
        self.loss_health = config["loss_health"]
        self.kd_threshold = config["kd_threshold"]

        self.user_embedding = nn.Embedding(self.n_users, d)
        self.item_embedding = nn.Embedding(self.n_items, d)
        self.ingre_embedding = nn.Embedding(self.n_ingredients + 1, d, padding_idx=self.n_ingredients)
        self.mf_loss = BPRLoss()
        self.reg_loss = EmbLoss()

        dev = self.device
        self.g_ui = G.norm_adj_user_item(self.interaction_matrix, self.n_users, self.n_items, dev)
        self.g_ri = G.norm_adj_item_side(dataset.rIngre_triples, self.n_items, self.n_ingredients, dev)

        self.apply(xavier_uniform_initialization)
        if self.v_feat is not None:
            self.image_embedding = nn.Embedding.from_pretrained(self.v_feat, freeze=False)
            self.image_trs = nn.Linear(self.v_feat.shape[1], d)
            nn.init.xavier_normal_(self.image_trs.weight)
        if self.t_feat is not None:
            self.text_embedding = nn.Embedding.from_pretrained(self.t_feat, freeze=False)
            self.text_trs = nn.Linear(self.t_feat.shape[1], d)
            nn.init.xavier_normal_(self.text_trs.weight)

    # ------------------------------------------------------------------ propagation
    def _propagate_all(self):
        ir = ops.propagate_mean(self.g_ri, self.item_embedding.weight, self.n_layers, bottom=self.ingre_embedding.weight[:-1])
        all_emb = ops.propagate_mean(self.g_ui, self.user_embedding.weight, self.ui_layers, bottom=ir[:self.n_items])
        return all_emb, ir

    def forward(self):
        all_emb, ir = self._propagate_all()
        return all_emb[:self.n_users], all_emb[self.n_users:], ir[self.n_items:]

    # ------------------------------------------------------------------ training loss
    def calculate_loss(self, batch_data):
        user, pos_item, neg_item = batch_data["u_id"], batch_data["pos_i_id"], batch_data["neg_i_id"]
        pos_ing, neg_ing = batch_data["pos_ingre_code"], batch_data["neg_ingre_code"]
        all_item = torch.cat([pos_item, neg_item], dim=0)
        all_emb, _ = self._propagate_all()

        # ---- knowledge / health branch (dense, batch-sized; outside the hot path)
        ingredients = torch.cat([pos_ing, neg_ing], dim=0)
        ingre_num = torch.cat([batch_data["pos_ingre_num"], batch_data["neg_ingre_num"]], dim=0)
        health_level = torch.cat([batch_data["pos_hl_mh"], batch_data["neg_hl_mh"]], dim=0)
        # (kernel gather + atomic scatter-add backward: torch's sort-based `index_put_` backward serialises on the padding
        #  index, which fills ~11 of the 20 slots of every recipe -- it was 5.1 of the 10.9 ms of a C2 step)
        ingr = ops.gather_rows(self.ingre_embedding.weight, ingredients.reshape(-1)).view(*ingredients.shape, -1)
        encoded = self.ingr_encoder(ingr.permute(1, 0, 2), src_key_padding_mask=(ingredients == self.n_ingredients))
        encoded = encoded.permute(1, 0, 2).contiguous()
        if getattr(self, "project_all_items", False):
            # the reference's formulation (cikm_model.py:240-244): project EVERY item's raw features, then keep 2B rows
            text_feats = self.text_trs(self.text_embedding.weight)
            image_feats = self.image_trs(self.image_embedding.weight)
            query = torch.cat([image_feats[all_item].unsqueeze(1), text_feats[all_item].unsqueeze(1)], dim=1)
        else:
            # Only the rows `all_item` of the projections are consumed (and only those rows of the trainable feature
            # tables receive a non-zero gradient), so the rows are gathered FIRST: [2B, Dv] x [Dv, d] instead of
            # [I, Dv] x [Dv, d] -- the same values and the same dense gradients (zero rows included, which dense Adam
            # needs) for 2B / I of the FLOPs and bytes (C2: 1024 of 45 000 rows).
            image_q = self.image_trs(ops.gather_rows(self.image_embedding.weight, all_item))
            text_q = self.text_trs(ops.gather_rows(self.text_embedding.weight, all_item))
            query = torch.cat([image_q.unsqueeze(1), text_q.unsqueeze(1)], dim=1)
        item_health, _ = self.mm_target_atten(query, encoded, ingredients)
        item_mm, _ = self.ingre_target_atten(encoded, query)
        item_know = F.normalize(item_mm).sum(1) / ingre_num.unsqueeze(1)
        health_pred = torch.sigmoid(self.health_mlp(F.normalize(item_health).mean(dim=1)))
        health_loss = torch.sum(self.criterion(health_pred, health_level))

        # ---- ranking + regulariser (fused kernel)
        uw, iw, gw = self.user_embedding.weight, self.item_embedding.weight, self.ingre_embedding.weight
        pad = self.n_ingredients
        mf_loss, reg = ops.rank_loss(all_emb, self.n_users, user, pos_item, neg_item,
                                     [(uw, user), (iw, pos_item), (iw, neg_item), (gw, pos_ing, pad), (gw, neg_ing, pad)],
                                     reg_den=float(neg_ing.shape[0]), gamma=self.mf_loss.gamma)
        # 1 - mean cos(item_know, [pos_e; neg_e]) against the propagated item rows, gathered inside the kernel
        kd = 1 - ops.cosine_mean(item_know, all_emb, all_item + self.n_users)
        kd_loss = torch.clamp_min(kd - self.kd_threshold, 0.0)
        return mf_loss, self.loss_health * health_loss, self.loss_kd * kd_loss, (self.reg_weight * reg).reshape(1)

    def norm_loss(self, kd_loss, threshold):
        return torch.clamp_min(kd_loss - threshold, 0.0)
