"""Registries — the drop-in boundary (mirrors /root/reference/simvg/models/builder.py:1-36).

The reference resolves `{type: "BEIT3", ...}` / `{type: "TextGuidedQuerySelectKDDETRHead", ...}` / `{type: "MIXDETRMB", ...}`
config dicts through mmcv `Registry` objects.  When mmcv is importable the same class is used (so the classes here register
into genuine mmcv registries); otherwise a minimal compatible `Registry` (register_module() decorator / build(cfg,
default_args)) stands in.
"""
try:  # pragma: no cover - mmcv is not installed in the build image
    from mmcv.utils import Registry
except Exception:  # noqa: BLE001

    class Registry:
        def __init__(self, name):
            self.name = name
            self.module_dict = {}

        def get(self, key):
            return self.module_dict.get(key)

        def register_module(self, name=None, force=False, module=None):
            def _register(cls):
                key = name or cls.__name__
                if not force and key in self.module_dict:
                    raise KeyError("%s is already registered in %s" % (key, self.name))
                self.module_dict[key] = cls
                return cls

            if module is not None:
                return _register(module)
            return _register

        def build(self, cfg, default_args=None):
            if not isinstance(cfg, dict) or "type" not in cfg:
                raise KeyError('`cfg` must be a dict containing the key "type", got %r' % (cfg,))
            args = dict(cfg)
            if default_args is not None:
                for k, v in default_args.items():
                    args.setdefault(k, v)
            obj_type = args.pop("type")
            cls = self.get(obj_type) if isinstance(obj_type, str) else obj_type
            if cls is None:
                raise KeyError("%s is not in the %s registry" % (obj_type, self.name))
            return cls(**args)

        def __contains__(self, key):
            return key in self.module_dict

        def __repr__(self):
            return "Registry(name=%s, items=%s)" % (self.name, sorted(self.module_dict))


VIS_ENCODERS = Registry("VIS_ENCS")
LAN_ENCODERS = Registry("LAN_ENCS")
MODELS = Registry("MODELS")
FUSIONS = Registry("FUSIONS")
HEADS = Registry("HEADS")


def build_vis_enc(cfg):
    return VIS_ENCODERS.build(cfg)


def build_lan_enc(cfg, default_args):
    return LAN_ENCODERS.build(cfg, default_args=default_args)


def build_fusion(cfg):
    return FUSIONS.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_model(cfg, word_emb=None, num_token=-1):
    return MODELS.build(cfg, default_args=dict(word_emb=word_emb, num_token=num_token))
