"""The cfg keys the hot path reads, with the reference's defaults (StackGAN_v2/miscc/config.py:9-69 and
cfg/birds_3stages.yml). Under the reference's own main.py, `utils.install_as_reference_model()` binds the reference's
`miscc.config.cfg` object instead (bind_reference_cfg), so cfg/*.yml files keep working unchanged. The binding is
explicit: merely having `miscc.config` imported somewhere in the process (e.g. by the oracle's reference loader in the
tests) does not change which cfg the sg2b200 modules read."""
import sys
from types import SimpleNamespace


def default_cfg():
    return SimpleNamespace(
        CUDA=True,
        TREE=SimpleNamespace(BRANCH_NUM=3, BASE_SIZE=64),
        GAN=SimpleNamespace(GF_DIM=64, DF_DIM=64, EMBEDDING_DIM=128, Z_DIM=100, R_NUM=2, B_CONDITION=True),
        TEXT=SimpleNamespace(DIMENSION=1024),
        TRAIN=SimpleNamespace(
            BATCH_SIZE=24, GENERATOR_LR=2e-4, DISCRIMINATOR_LR=2e-4,
            COEFF=SimpleNamespace(KL=2.0, UNCOND_LOSS=1.0, COLOR_LOSS=0.0, CAL_LOSS=50.0)),
    )


cfg = default_cfg()
_bound = None          # explicit override (the reference's cfg object), see bind_reference_cfg()
_use_reference = False


def bind_reference_cfg(obj=None):
    """Make G_NET() / D_NETxx() read `obj` (default: the reference's `miscc.config.cfg`, resolved lazily when a
    network is constructed, because main.py imports miscc.config before it builds the networks). `unbind_cfg()` undoes it."""
    global _bound, _use_reference
    _bound = obj
    _use_reference = obj is None


def unbind_cfg():
    global _bound, _use_reference
    _bound, _use_reference = None, False


def active_cfg():
    """The cfg the modules read: an explicitly bound object, the reference's global cfg after
    install_as_reference_model(), else this package's own `cfg`."""
    if _bound is not None:
        return _bound
    if _use_reference:
        m = sys.modules.get("miscc.config")
        if m is not None and hasattr(m, "cfg"):
            return m.cfg
    return cfg


# Arithmetic mode of the engines built from now on: "bf16" (bf16 storage, fp32 accumulation: the fast path) or "fp32"
# (fp32 storage, convolutions as 3-way bf16 splits on the same tensor-core kernels: relative error <= 1e-4 against the
# fp32 reference, see csrc/precise.cu). Networks pick it up when their engine is first built; net.set_precision(p)
# rebuilds it.
PRECISION = "bf16"


def set_precision(p):
    global PRECISION
    if p not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    PRECISION = p
