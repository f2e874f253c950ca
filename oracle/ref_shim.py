"""Import shim for the *unmodified* reference (BolinLai/CSTS) — TEST INFRASTRUCTURE ONLY.

The reference at /root/reference is pure Python but imports packages that are not in this
image (fvcore, fairscale, ipdb, iopath, simplejson).  None of them contributes arithmetic to
the hot path: fvcore supplies a registry and an attribute-dict config,
fairscale an (unused) activation-checkpoint wrapper.  This module injects minimal stand-ins
into ``sys.modules`` and puts the reference on ``sys.path`` so that
``slowfast.models.build_model(cfg)`` runs on CPU exactly as written.

It is used by ``oracle/make_golden.py`` (to generate tests/golden/*.pt) and by the
``not gpu`` tests that pin ``oracle/csts_oracle.py`` against the live reference when
``/root/reference`` exists.  Nothing under ``csts_b200/`` imports it, and it is never used on
the GPU box (the reference tree does not exist there).
"""
import ast
import copy
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    """The reference tree: $CSTS_REFERENCE_ROOT, else /root/reference (the builder container), else baseline/_ref —
    the unmodified `slowfast` package as `pip install --no-deps --target baseline/_ref /root/reference` lays it down
    (git-ignored; it travels to the GPU box with the repo snapshot, where /root/reference does not exist)."""
    cands = [os.environ.get("CSTS_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "slowfast", "models")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "slowfast", "models"))


def config_path(config):
    """The reference YAML; the pip-installed package carries no configs/, the repo's copies hold identical keys and
    values (tests/test_host_cpu.py asserts that whenever both are present)."""
    p = os.path.join(REFERENCE_ROOT, "configs", config)
    return p if os.path.exists(p) else os.path.join(_REPO, "configs", config)


class _Registry(dict):
    """Stand-in for fvcore.common.registry.Registry (name -> object)."""

    def __init__(self, name):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self[o.__name__] = o
                return o
            return deco
        self[obj.__name__] = obj
        return obj

    def get(self, name):
        if name not in self:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self[name]

    @property
    def _obj_map(self):          # fvcore keeps its table under this name (INTEGRATION.md's binding line writes to it)
        return self


class _CfgNode(dict):
    """Stand-in for fvcore.common.config.CfgNode / yacs CfgNode (attribute dict)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    @staticmethod
    def _coerce(v):
        if isinstance(v, str):
            try:
                return ast.literal_eval(v)
            except (ValueError, SyntaxError):
                return v
        return v

    def _merge_dict(self, d):
        for k, v in d.items():
            if isinstance(v, dict):
                if k not in self or not isinstance(self[k], _CfgNode):
                    self[k] = _CfgNode()
                self[k]._merge_dict(v)
            else:
                v = self._coerce(v)
                if isinstance(v, tuple):
                    v = list(v)
                self[k] = v

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self._merge_dict(yaml.safe_load(f))

    def merge_from_list(self, lst):
        assert len(lst) % 2 == 0
        for k, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = self._coerce(v)

    def dump(self, **kw):
        import yaml
        return yaml.safe_dump(_to_plain(self), **kw)


def _to_plain(n):
    if isinstance(n, dict):
        return {k: _to_plain(v) for k, v in n.items()}
    return n


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Inject the stand-ins and make ``import slowfast`` resolve to the reference tree."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    fv = _module("fvcore")
    fvc = _module("fvcore.common")
    fv.common = fvc
    fvc.registry = _module("fvcore.common.registry", Registry=_Registry)
    fvc.config = _module("fvcore.common.config", CfgNode=_CfgNode)
    fs = _module("fairscale")
    fsn = _module("fairscale.nn")
    fs.nn = fsn
    fsn.checkpoint = _module("fairscale.nn.checkpoint", checkpoint_wrapper=lambda m, *a, **k: m)
    _module("ipdb")
    import json
    _module("simplejson", dumps=json.dumps, loads=json.loads)

    class _PathMgr:
        def open(self, p, mode="r", **kw):
            return open(p, mode)

        def exists(self, p):
            return os.path.exists(p)

        def mkdirs(self, p):
            os.makedirs(p, exist_ok=True)

        def ls(self, p):
            return os.listdir(p)

    class _PMF:
        @staticmethod
        def get(*a, **k):
            return _PathMgr()

    io = _module("iopath")
    ioc = _module("iopath.common")
    io.common = ioc
    ioc.file_io = _module("iopath.common.file_io", PathManagerFactory=_PMF)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def reference_cfg(config="Ego4D/CSTS_Ego4D_Gaze_Forecast.yaml", overrides=()):
    """get_cfg() + yaml + overrides, exactly as tools/run_net.py does (parser.py:75-81)."""
    install()
    from slowfast.config.defaults import get_cfg
    cfg = get_cfg()
    cfg.merge_from_file(config_path(config))
    base = ["NUM_GPUS", 0, "MODEL.LOSS_FUNC", "kldiv+egonce", "MVIT.DROPPATH_RATE", 0.0]
    cfg.merge_from_list(base + list(overrides))
    return cfg


def reference_model(seed=0, config="Ego4D/CSTS_Ego4D_Gaze_Forecast.yaml", overrides=()):
    """The reference CSTS module on CPU, random init under torch.manual_seed(seed)."""
    import torch
    cfg = reference_cfg(config, overrides)
    from slowfast.models import build_model
    torch.manual_seed(seed)
    model = build_model(cfg)
    return model, cfg


def reference_loss(preds, v_embed, a_embed, labels_hm, alpha=0.05):
    """kldiv+egonce exactly as tools/train_avgaze_net.py:76-88 composes it.

    EgoNCE.forward (losses.py:157-170) hard-codes ``.cuda()`` for its eye mask; on a CPU-only
    box that line is the one thing restated here (eye on x.device), everything else is the
    reference's own code.
    """
    install()
    import torch
    import torch.nn.functional as F
    from slowfast.models import losses
    from slowfast.utils.utils import frame_softmax, sim_matrix
    p = frame_softmax(preds, temperature=2)
    sim = sim_matrix(v_embed, a_embed)
    kld = losses.get_loss_func("kldiv")()(p, labels_hm)
    x = sim
    mask_bool = torch.eye(x.shape[0], device=x.device) > 0       # losses.py:158 without .cuda()
    i_sm = F.softmax(x / 0.05, dim=1)
    j_sm = F.softmax(x.t() / 0.05, dim=1)
    idiag = torch.log(torch.sum(i_sm * mask_bool, dim=1))
    jdiag = torch.log(torch.sum(j_sm * mask_bool, dim=1))
    nce = -idiag.sum() / len(idiag) - jdiag.sum() / len(jdiag)
    return kld + alpha * nce, kld, nce
