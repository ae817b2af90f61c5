"""CPU restatement of the caller-side image plumbing around CodeFormer.forward.

TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
Pinned: tests/test_oracle.py checks it against the reference's own functions when /root/reference is present and
against tests/golden/plumbing.npz (made by oracle/gen_golden.py from those functions) everywhere else.

face_to_input   inference_codeformer.py:199-200 -> basicsr/utils/img_util.py:22-29 (img2tensor) + torchvision normalize
output_to_face  inference_codeformer.py:206,213 -> basicsr/utils/img_util.py:66-67,80-90 (tensor2img, min_max=(-1,1))
"""
import numpy as np


def face_to_input(faces_bgr_u8: np.ndarray) -> np.ndarray:
    """uint8 [B,H,W,3] BGR -> float32 [B,3,H,W] RGB in [-1,1]."""
    f = (faces_bgr_u8 / 255.).astype(np.float32)          # float64 division, then astype('float32')  (img_util.py:24-25)
    f = f[..., ::-1]                                      # cv2.COLOR_BGR2RGB                          (img_util.py:26)
    x = np.ascontiguousarray(f.transpose(0, 3, 1, 2))     # HWC -> CHW                                 (img_util.py:27)
    return ((x - np.float32(0.5)) / np.float32(0.5)).astype(np.float32)   # normalize(mean .5, std .5), fp32


def output_to_face(out_nchw: np.ndarray) -> np.ndarray:
    """float32 [B,3,H,W] RGB -> uint8 [B,H,W,3] BGR (clamp to [-1,1], (x+1)/2, *255, round half to even)."""
    t = np.clip(out_nchw.astype(np.float32), np.float32(-1), np.float32(1))         # img_util.py:66
    t = ((t - np.float32(-1)) / np.float32(2)).astype(np.float32)                  # img_util.py:67
    img = t.transpose(0, 2, 3, 1)[..., ::-1]                                       # CHW -> HWC, RGB2BGR  (:80-85)
    img = (img * np.float32(255.0)).round()                                        # :88-89 (np.round: half to even)
    return np.ascontiguousarray(img).astype(np.uint8)                              # :90, inference_codeformer.py:213
