"""Minimal yacs-compatible config for the DiffusionVID hot path.

The reference reads a global yacs CfgNode (mega_core/config/defaults.py) extended by add_diffusiondet_config
(mega_core/modeling/detector/diffusion_det.py:74-179) and merged with configs/vid_*_DiffusionVID.yaml.  yacs is not a
dependency here; this module provides the same surface (attribute access, merge_from_file / merge_from_list / clone /
freeze) for the keys the inference path reads, and `hot_path_params(cfg)` flattens either this CfgNode or a real yacs
node into the plain dict the model uses.  Unknown keys in a yaml file are accepted (the reference yaml files carry
solver/dataset sections that the hot path never reads).
"""
import copy


class CfgNode(dict):
    def __init__(self, init=None):
        super().__init__()
        self.__dict__["_frozen"] = False
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__.get("_frozen"):
            raise AttributeError("Attempted to set {} on a frozen CfgNode".format(name))
        self[name] = value

    def freeze(self):
        self.__dict__["_frozen"] = True
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze()

    def defrost(self):
        self.__dict__["_frozen"] = False
        for v in self.values():
            if isinstance(v, CfgNode):
                v.defrost()

    def clone(self):
        return copy.deepcopy(self)

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self or not isinstance(self[k], CfgNode):
                    self[k] = CfgNode()
                self[k]._merge(v)
            else:
                self[k] = v

    def merge_from_file(self, path):
        import yaml
        with open(path, "r") as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, opts):
        assert len(opts) % 2 == 0, "opts must be KEY VALUE pairs"
        import yaml
        for key, val in zip(opts[0::2], opts[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                if p not in node:
                    node[p] = CfgNode()
                node = node[p]
            node[parts[-1]] = yaml.safe_load(val) if isinstance(val, str) else val


def get_default_cfg():
    """The subset of mega_core/config/defaults.py + add_diffusiondet_config that the inference path reads, with the
    reference's default values."""
    cfg = CfgNode()
    cfg.DTYPE = "float16"
    cfg.MODEL = CfgNode()
    cfg.MODEL.DEVICE = "cuda"
    cfg.MODEL.META_ARCHITECTURE = "DiffusionDet"
    cfg.MODEL.PIXEL_MEAN = [123.675, 116.280, 103.530]
    cfg.MODEL.PIXEL_STD = [58.395, 57.120, 57.375]
    cfg.MODEL.BACKBONE = CfgNode({"NAME": "build_resnet_fpn_backbone", "CONV_BODY": "R-101-torchvision"})
    cfg.MODEL.RESNETS = CfgNode({"DEPTH": 101, "STRIDE_IN_1X1": False, "RES5_DILATION": 1,
                                 "OUT_FEATURES": ["res2", "res3", "res4", "res5"]})
    cfg.MODEL.FPN = CfgNode({"IN_FEATURES": ["res3", "res4", "res5"], "OUT_CHANNELS": 256})
    cfg.MODEL.ROI_HEADS = CfgNode({"IN_FEATURES": ["p3", "p4", "p5"]})
    cfg.MODEL.ROI_BOX_HEAD = CfgNode({"POOLER_TYPE": "ROIAlignV2", "POOLER_RESOLUTION": 7,
                                      "POOLER_SAMPLING_RATIO": 2})
    add_diffusiondet_config(cfg)
    cfg.MODEL.VID = CfgNode({"METHOD": "diffusion", "ROI_BOX_HEAD": {"ATTENTION": {"ENABLE": False, "STAGE": 1}}})
    cfg.MODEL.VID.MEGA = CfgNode({"MIN_OFFSET": 0, "MAX_OFFSET": 7, "ALL_FRAME_INTERVAL": 8, "KEY_FRAME_LOCATION": 0,
                                  "MEMORY_MANAGEMENT_SIZE_TEST": 900,
                                  "GLOBAL": {"ENABLE": True, "RES_STAGE": 1, "SIZE": 24,
                                             "STOP_UPDATE_AFTER_INIT_TEST": True}})
    cfg.INPUT = CfgNode({"INFER_BATCH": 8, "MIN_SIZE_TEST": 600, "MAX_SIZE_TEST": 1000})
    cfg.DATALOADER = CfgNode({"SIZE_DIVISIBILITY": 32})
    return cfg


def add_diffusiondet_config(cfg):
    """Same keys and defaults as mega_core/modeling/detector/diffusion_det.py:74-120 (inference-relevant subset)."""
    if "MODEL" not in cfg:
        cfg.MODEL = CfgNode()
    d = CfgNode()
    d.NUM_CLASSES = 80
    d.NUM_PROPOSALS = 300
    d.NHEADS = 8
    d.DROPOUT = 0.0
    d.DIM_FEEDFORWARD = 2048
    d.ACTIVATION = "relu"
    d.HIDDEN_DIM = 256
    d.NUM_CLS = 1
    d.NUM_REG = 3
    d.NUM_HEADS = 6
    d.NUM_HEADS_LOCAL = 0
    d.NUM_DYNAMIC = 2
    d.DIM_DYNAMIC = 64
    d.USE_FOCAL = True
    d.USE_FED_LOSS = False
    d.PRIOR_PROB = 0.01
    d.SNR_SCALE = 2.0
    d.SAMPLE_STEP = 1
    d.USE_NMS = True
    cfg.MODEL.DiffusionDet = d


def _get(node, path, default=None):
    cur = node
    for p in path.split("."):
        try:
            cur = cur[p] if isinstance(cur, dict) else getattr(cur, p)
        except (KeyError, AttributeError):
            return default
    return cur


# size2config of mega_core/modeling/backbone/swintransformer.py:660-717 (window 7 variants)
SWIN_SIZES = {
    "T": dict(embed=96, depths=(2, 2, 6, 2), heads=(3, 6, 12, 24)),
    "S": dict(embed=96, depths=(2, 2, 18, 2), heads=(3, 6, 12, 24)),
    "B": dict(embed=128, depths=(2, 2, 18, 2), heads=(4, 8, 16, 32)),
    "B-22k": dict(embed=128, depths=(2, 2, 18, 2), heads=(4, 8, 16, 32)),
    "L-22k": dict(embed=192, depths=(2, 2, 18, 2), heads=(6, 12, 24, 48)),
}


def hot_path_params(cfg):
    """Flatten a CfgNode (ours or yacs) - or pass through a plain dict - into the model's parameter dict."""
    if isinstance(cfg, dict) and not isinstance(cfg, CfgNode) and "num_proposals" in cfg:
        return dict(cfg)
    depth = _get(cfg, "MODEL.RESNETS.DEPTH", 101)
    blocks = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}[depth]
    return dict(
        num_proposals=_get(cfg, "MODEL.DiffusionDet.NUM_PROPOSALS", 300),
        num_classes=_get(cfg, "MODEL.DiffusionDet.NUM_CLASSES", 30),
        hidden=_get(cfg, "MODEL.DiffusionDet.HIDDEN_DIM", 256),
        nheads=_get(cfg, "MODEL.DiffusionDet.NHEADS", 8),
        dim_dynamic=_get(cfg, "MODEL.DiffusionDet.DIM_DYNAMIC", 64),
        dim_ff=_get(cfg, "MODEL.DiffusionDet.DIM_FEEDFORWARD", 2048),
        num_heads=_get(cfg, "MODEL.DiffusionDet.NUM_HEADS", 3),
        num_heads_local=_get(cfg, "MODEL.DiffusionDet.NUM_HEADS_LOCAL", 1),
        num_cls=_get(cfg, "MODEL.DiffusionDet.NUM_CLS", 1),
        num_reg=_get(cfg, "MODEL.DiffusionDet.NUM_REG", 3),
        sample_step=_get(cfg, "MODEL.DiffusionDet.SAMPLE_STEP", 1),
        snr_scale=float(_get(cfg, "MODEL.DiffusionDet.SNR_SCALE", 2.0)),
        use_nms=bool(_get(cfg, "MODEL.DiffusionDet.USE_NMS", True)),
        infer_batch=_get(cfg, "INPUT.INFER_BATCH", 8),
        all_frame_interval=_get(cfg, "MODEL.VID.MEGA.ALL_FRAME_INTERVAL", 8),
        key_frame_location=_get(cfg, "MODEL.VID.MEGA.KEY_FRAME_LOCATION", 0),
        global_enable=bool(_get(cfg, "MODEL.VID.MEGA.GLOBAL.ENABLE", True)),
        global_res_stage=int(_get(cfg, "MODEL.VID.MEGA.GLOBAL.RES_STAGE", 1)),
        mem_size=_get(cfg, "MODEL.VID.MEGA.MEMORY_MANAGEMENT_SIZE_TEST", 900),
        mem_size2=150,                      # hard-coded in the reference, diffusion_det.py:487
        topk=(75, 25),                      # hard-coded in the reference, box_head.py:235
        pixel_mean=tuple(_get(cfg, "MODEL.PIXEL_MEAN", [123.675, 116.280, 103.530])),
        pixel_std=tuple(_get(cfg, "MODEL.PIXEL_STD", [58.395, 57.120, 57.375])),
        blocks=blocks,
        swin=SWIN_SIZES[_get(cfg, "MODEL.SWIN.SIZE", "B-22k")]
        if "swin" in str(_get(cfg, "MODEL.BACKBONE.NAME", "")).lower() else None,
        device=_get(cfg, "MODEL.DEVICE", "cuda"),
    )
