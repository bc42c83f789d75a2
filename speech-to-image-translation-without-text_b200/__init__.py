"""sg2b200 — B200-native (sm_100a) speech-conditioned StackGAN-v2 train step.

Drop-in for the reference's StackGAN_v2/model.py module API (G_NET, D_NET64/128/256) over
hand-written tcgen05/TMA CUDA kernels behind a C ABI (include/sg2b200.h)."""
__version__ = "0.1.0"
