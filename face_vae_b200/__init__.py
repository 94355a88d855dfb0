"""face_vae_b200: B200-native (sm_100a) implementation of the face-vae training hot path.

Host side is Python/PyTorch (memory, streams, autograd plumbing, torch.distributed); all device work goes through
the C ABI of libfacevae_b200.so (include/facevae_b200.h).  There is no CPU path.
"""
__version__ = "0.1.0"
