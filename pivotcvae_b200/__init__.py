"""pivotcvae_b200 — the PivotCVAE slate-generation hot path on B200 (sm_100a).

Host side mirrors the reference's Python API (models.cvae / models.pivotcvae /
models.listcvae / env.response_model / train_generative); the arithmetic runs in
hand-written CUDA behind the C ABI of include/pcv_b200.h (libpcv_b200.so).
"""
import sys

__all__ = ["install_dropin"]


def install_dropin():
    """Register this package's mirrors under the reference's import paths
    (`models.pivotcvae`, `env.response_model`, `train_generative`, ...) so that
    reference scripts, notebooks and whole-model pickles resolve to the B200 path."""
    import types

    from . import pretrain_env, train_generative
    from .env import response_model
    from .models import cvae, deterministic, listcvae, pivotcvae

    models_pkg = sys.modules.setdefault("models", types.ModuleType("models"))
    env_pkg = sys.modules.setdefault("env", types.ModuleType("env"))
    for name, mod in (("cvae", cvae), ("pivotcvae", pivotcvae), ("listcvae", listcvae), ("deterministic", deterministic)):
        sys.modules["models." + name] = mod
        setattr(models_pkg, name, mod)
    sys.modules["env.response_model"] = response_model
    env_pkg.response_model = response_model
    sys.modules["train_generative"] = train_generative
    sys.modules["pretrain_env"] = pretrain_env          # train_response_model (pretrain_env.py:25-139)
    # slate metrics (analysis.py:5-30): the reference module also holds ranking metrics that are not on the
    # path, so an importable reference `analysis` only gets its two slate metrics replaced
    from . import analysis as slate_metrics
    try:
        import analysis as ref_analysis
    except ImportError:
        sys.modules["analysis"] = slate_metrics
    else:
        ref_analysis.get_coverage = slate_metrics.get_coverage
        ref_analysis.get_ILS = slate_metrics.get_ILS
    return models_pkg, env_pkg
