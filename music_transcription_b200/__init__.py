"""B200-native audio->piano-roll inference path of cs4247/music-transcription.

Host side: Python/PyTorch (device memory, streams, torch.distributed).
Device side: hand-written sm_100a CUDA kernels behind the C ABI declared in
include/amt.h (libamt_sm100.so, loaded through ctypes).  There is no CPU
fallback: using the compute path without the built library raises.
"""
__version__ = "0.1.0"
