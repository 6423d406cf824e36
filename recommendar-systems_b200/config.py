"""Config: the reference's key names and defaults without its YAML-merge machinery.

Mirrors /root/reference/src/utils/configurator.py:15-149 (dict-like `config[key]` returning
None for unknown keys, `config.get(key, default)`), /root/reference/src/configs/overall.yaml and
the first grid point of each model YAML. The YAML grid search itself is
orchestration glue and out of scope; a caller overrides any key through `config_dict`.
"""
from __future__ import annotations

import torch

OVERALL = {
    # overall.yaml:1-47
    "gpu_id": 0, "use_gpu": True, "seed": 999,
    "data_path": "../data/", "inter_splitting_label": "x_label",
    "filter_out_cod_start_users": True, "is_multimodal_model": True,
    "embedding_size": 64, "weight_decay": 0.0, "req_training": True,
    "epochs": 1000, "stopping_step": 20, "train_batch_size": 2048,
    "learner": "adam", "learning_rate": 0.001, "learning_rate_scheduler": [1.0, 50],
    "eval_step": 1, "training_neg_sample_num": 1, "use_neg_sampling": True,
    "use_full_sampling": False, "NEG_PREFIX": "neg__",
    "USER_ID_FIELD": "userID", "ITEM_ID_FIELD": "itemID", "field_separator": "\t",
    "metrics": ["Recall", "NDCG", "Precision", "MAP"], "topk": [5, 10, 20, 50],
    "valid_metric": "Recall@20", "eval_batch_size": 4096, "end2end": False,
    "vision_feature_file": "image_feat.npy", "text_feature_file": "text_feat.npy",
    "save_recommended_topk": False, "clip_grad_norm": None,
}

MODEL_DEFAULTS = {
    # configs/model/LayerGCN.yaml:1-5 (first grid point)
    "LayerGCN": {"n_layers": 4, "reg_weight": 1e-2, "dropout": 0.0},
    # configs/model/LightGCN.yaml
    "LightGCN": {"is_multimodal_model": False, "n_layers": 4, "reg_weight": 1e-2},
    # configs/model/FREEDOM.yaml
    "FREEDOM": {"feat_embed_dim": 64, "lambda_coeff": 0.9, "reg_weight": 0.0, "n_mm_layers": 1,
                "n_ui_layers": 2, "knn_k": 10, "mm_image_weight": 0.1, "dropout": 0.8,
                "cf_model": None, "degree_ratio": None},
    # configs/model/MGCN.yaml
    "MGCN": {"n_ui_layers": 2, "n_layers": 1, "learning_rate_scheduler": [0.96, 50],
             "reg_weight": 1e-4, "knn_k": 10, "cl_loss": 0.001},
    # configs/model/SMORE.yaml
    "SMORE": {"n_ui_layers": 4, "n_layers": 1, "learning_rate_scheduler": [0.96, 50],
              "reg_weight": 1e-5, "cl_loss": 0.01, "temperature": 0.2, "image_knn_k": 20,
              "text_knn_k": 15, "dropout_rate": 0.1, "mg_enable": True, "mg_interval": 3,
              "mg_alpha": 0.5, "mg_beta": 0.2, "mg_verbose": False, "diag_spectrum": False,
              "diag_gate": False, "diag_grad": False, "mg_target_rel_step": 1e-3},
}


class Config:
    def __init__(self, model=None, dataset=None, config_dict=None):
        d = dict(OVERALL)
        d.update(MODEL_DEFAULTS.get(model, {}))
        d["model"], d["dataset"] = model, dataset
        d.update(config_dict or {})
        valid_metric = d["valid_metric"].split("@")[0]
        d["valid_metric_bigger"] = valid_metric.lower() not in ("rmse", "mae", "logloss")
        if "device" not in d:
            use = d["use_gpu"] and torch.cuda.is_available()
            d["device"] = torch.device("cuda", torch.cuda.current_device()) if use \
                else torch.device("cpu")
        self.final_config_dict = d

    def __setitem__(self, key, value):
        if not isinstance(key, str):
            raise TypeError("index must be a str.")
        self.final_config_dict[key] = value

    def __getitem__(self, item):
        return self.final_config_dict.get(item, None)

    def get(self, key, default=None):
        if not isinstance(key, str):
            raise TypeError("index must be a str.")
        return self.final_config_dict.get(key, default)

    def __contains__(self, key):
        return key in self.final_config_dict

    def __str__(self):
        return "\n" + "\n".join(f"{k}={v}" for k, v in self.final_config_dict.items()) + "\n\n"

    __repr__ = __str__


def init_seed(seed):
    """utils/utils.py:48-54."""
    import random

    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    torch.manual_seed(seed)
