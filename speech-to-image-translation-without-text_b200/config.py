"""The cfg keys the hot path reads, with the reference's defaults (StackGAN_v2/miscc/config.py:9-69 and
cfg/birds_3stages.yml). When the reference's own `miscc.config.cfg` is importable (running under the reference's
main.py) that object is used instead, so cfg/*.yml files keep working unchanged."""
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


def active_cfg():
    """The reference's global cfg if its miscc.config is already imported, else ours."""
    m = sys.modules.get("miscc.config")
    return m.cfg if m is not None and hasattr(m, "cfg") else cfg
