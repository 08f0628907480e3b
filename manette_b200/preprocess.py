"""Stand-alone K3 (reference: atari_emulator.py:79-88 on top of ALE's getScreenGrayscale/getScreenRGB):
max of two raw palette-index screens -> luminance / RGB -> 210x160 -> 84x84 nearest."""
import ctypes as C

import torch

from . import _native


def preprocess(frames, rgb=False, out=None, stream=None):
    """frames: (n, 2, 210, 160) uint8 CUDA palette indices -> (n, 84, 84, depth) uint8."""
    assert frames.is_cuda and frames.dtype == torch.uint8 and tuple(frames.shape[1:]) == (2, 210, 160)
    frames = frames.contiguous()
    n = frames.shape[0]
    d = 3 if rgb else 1
    if out is None:
        out = torch.empty(n, 84, 84, d, dtype=torch.uint8, device=frames.device)
    st = torch.cuda.current_stream(frames.device) if stream is None else stream
    with torch.cuda.device(frames.device):
        _native.check(_native.load().mn_preprocess(frames.data_ptr(), out.data_ptr(), n, int(bool(rgb)),
                                                   C.c_void_p(st.cuda_stream)), "mn_preprocess")
    return out
